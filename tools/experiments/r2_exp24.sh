# follow-mode boundary chain: A/B against the chunked resumable chain, timeline, chunk sweep, then the segmentation tests
set -x
HIPPO_PATTERN_FOLLOW=0 CHUNKS=444 python tools/seg_only.py 2>&1 | grep seg_only
CHUNKS=444,588,1176,3599 TIMELINE=588 python tools/seg_only.py 2>&1 | tail -40
CUDA_LAUNCH_BLOCKING=1 CHUNKS=588 python tools/seg_only.py 2>&1 | grep "overlapped"
timeout 900 python -m pytest tests/test_gpu_segmentation.py tests/test_gpu_fullsize.py -m gpu -x -q --tb=short -p no:cacheprovider -k "seg or pattern or ssim or stream or config2 or boundar" 2>&1 | tail -5
