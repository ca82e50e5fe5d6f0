# Full GPU validation + measurement pass (run under gpurun from the repo root).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 500 python bench.py --steps 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc $?"; tail -9 gpurun_out/bench.err
timeout 200 python bench.py --impl reference --steps 4 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc $?"
KREGEX='regex:sim_tc_kernel|topk_merge_kernel|bank_build_kernel|topk_single_kernel|greedy_scan_kernel|recheck_kernel|compact_kernel|ssim_pair_kernel|gray_minmax_kernel|audio_energy|segment_kernel|ssim_finalize|minmax_init'
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sim_tc_kernel -s 2 -c 1 -f -o gpurun_out/sim_tc_topk python bench.py --steps 1 --no-extra --bank-rows 2000000 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc $?"
