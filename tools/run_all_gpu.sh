# Full GPU validation + measurement pass (run under gpurun from the repo root): GPU suite, smoke, both bench arms,
# ncu launch list of the bench command.
set -u
mkdir -p gpurun_out
T0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - T0 ))s"; }
timeout 1200 python -m pytest tests -m gpu -q --tb=short --timeout 600 --timeout-method=thread -p no:cacheprovider --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -14 gpurun_out/pytest_gpu.log; lap pytest
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2; lap smoke
timeout 200 python bench.py --impl reference --steps 4 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc $?"; lap reference
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc $?"; grep -E "extra|gate" gpurun_out/bench.err | tail -30; lap bench
KREGEX='regex:sim_tc_kernel|topk_merge_kernel|bank_build_kernel|topk_single_kernel|topk_small|topk_few|exchange|greedy_scan_kernel|recheck_kernel|cons_|ssim_pair|gray_minmax|audio_energy|segment_kernel|ssim_finalize|minmax_init|pattern_init|topk_rows|rescore'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --no-extra > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc $?"; lap launches
