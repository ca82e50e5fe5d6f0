# round 2, call 14: final segmentation structure: SSIM <4> monolithic / <3> in the pipeline; batch modes
set -u
timeout 300 python -m pytest tests/test_gpu_segmentation.py tests/test_gpu_prefilter.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -x 2>&1 | tail -3
CHUNKS=444 BATCH=32 timeout 300 python tools/seg_only.py 2>&1 | grep -E "seg_only"
