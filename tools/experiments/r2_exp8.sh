# round 2, call 8: single-lane chunks + chain under them; SSIM tail quantisation
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_segmentation.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -x 2>&1 | tail -3
TAIL=1 CHUNKS=444,888,1200,1776 TIMELINE=888 BATCH=32 timeout 300 python tools/seg_only.py 2>&1 | grep -E "seg_only|pattern" | head -80
HIPPO_PATTERN_LANES=2 CHUNKS=888 timeout 300 python tools/seg_only.py 2>&1 | grep -E "seg_only" | head -3
