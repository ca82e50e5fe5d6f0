# round 2, call 19: small-batch tcgen05 kernel (bank rows on the M side): tests, batch sweep; consolidation after the T-pairs rule
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_search.py -m gpu -q --tb=short --timeout 120 -p no:cacheprovider -x 2>&1 | tail -8
SWEEP=2,3,4,8,16,32,33,64,65,128,256 timeout 300 python tools/batch_sweep.py 2> gpurun_out/r2_batch_sweep.err > gpurun_out/r2_batch_sweep.json; echo "sweep rc $?"; grep "queries" gpurun_out/r2_batch_sweep.err | cut -c1-150
HIPPO_SMALL_BATCH=0 SWEEP=8,64 timeout 300 python tools/batch_sweep.py 2>&1 | grep "'queries'" | cut -c1-150
