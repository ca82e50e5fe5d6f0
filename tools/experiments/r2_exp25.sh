# fused gray + SSIM kernel, follower launched first: tests, timings, A/B against the two-kernel path, batch
set -x
timeout 900 python -m pytest tests/test_gpu_segmentation.py tests/test_gpu_prefilter.py -m gpu -x -q --tb=short -p no:cacheprovider 2>&1 | tail -8
CHUNKS=0 TIMELINE=0 BATCH=32 python tools/seg_only.py 2>&1 | tail -20
HIPPO_FRAMES_UNFUSED=1 CHUNKS=0 python tools/seg_only.py 2>&1 | grep seg_only
for lead in 148 296 1184 2368; do HIPPO_FUSED_LEAD=$lead CHUNKS=0 python tools/seg_only.py 2>&1 | grep "seg_only\] \(over\|segm\)" | sed "s/^/lead $lead: /"; done
