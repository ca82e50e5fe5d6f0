# round 2, call 4: segmentation pipeline (tests + timings), launch list of one shard-size search step (1.25M rows = 10M / 8)
set -u
mkdir -p gpurun_out
T0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - T0 ))s"; }
timeout 600 python -m pytest tests/test_gpu_segmentation.py tests/test_gpu_prefilter.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -x 2>&1 | tail -15; lap pytest_seg
CHUNKS=222,444,888,1776 BATCH=32 timeout 300 python tools/seg_only.py 2>&1 | grep seg_only; lap seg_only
KREGEX='regex:sim_tc_kernel|topk_merge_kernel|bank_build_kernel|topk_single_kernel|exchange'
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" --csv --log-file gpurun_out/r2_launches_shard.csv python bench.py --steps 3 --no-extra --bank-rows 1250000 > gpurun_out/r2_ncu_shard.log 2>&1; echo "ncu shard rc $?"; lap shard
timeout 200 python bench.py --steps 20 --no-extra --bank-rows 1250000 2> /dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('shard-size step (1.25M rows, 1 GPU):', d['ms_per_step'], 'ms; ideal', 61.26/8)"; lap shard_bench
