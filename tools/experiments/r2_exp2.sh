# round 2, call 2: the whole GPU suite with the new full-size parity tests, the direct-rows search, exact re-scoring,
# consolidation overflow retry, read-only bank cache
set -u
mkdir -p gpurun_out
T0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - T0 ))s"; }
timeout 1200 python -m pytest tests -m gpu -q --tb=short --timeout 600 --timeout-method=thread -p no:cacheprovider --durations=12 -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -40 gpurun_out/r2_pytest_gpu.log; lap pytest
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2; lap smoke
