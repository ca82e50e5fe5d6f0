# round 2, call 15: whole GPU suite + full N=1 bench with the new extras
set -u
mkdir -p gpurun_out
T0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - T0 ))s"; }
timeout 1200 python -m pytest tests -m gpu -q --tb=short --timeout 600 --timeout-method=thread -p no:cacheprovider --durations=6 -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -14 gpurun_out/r2_pytest_gpu.log; lap pytest
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2; lap smoke
timeout 900 python bench.py --steps 10 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc $?"; grep -E "extra|gate|cpu" gpurun_out/r2_bench.err | tail -30; lap bench
