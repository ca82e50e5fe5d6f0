# round 2, call 6: hippo_pattern_separation (C-level orchestration): tests, single-stream and batch timings
set -u
mkdir -p gpurun_out
T0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - T0 ))s"; }
timeout 600 python -m pytest tests/test_gpu_segmentation.py tests/test_gpu_prefilter.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -x 2>&1 | tail -15; lap pytest_seg
CHUNKS=222,300,444,600,888,1200 BATCH=32 timeout 300 python tools/seg_only.py 2>&1 | grep seg_only; lap seg_only
