# round 2, call 10: gray once + SSIM chunks on alternating streams + chain after every chunk
set -u
timeout 300 python -m pytest tests/test_gpu_segmentation.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -x 2>&1 | tail -3
CHUNKS=444,888,222,444 TIMELINE=444 BATCH=32 timeout 300 python tools/seg_only.py 2>&1 | grep -E "seg_only|pattern" | head -80
