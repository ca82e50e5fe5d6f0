timeout 300 python -m pytest tests/test_gpu_search.py tests/test_gpu_consolidation.py -m gpu -q --tb=short --timeout 120 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -5
for dbg in 0 1; do HIPPO_TC_DEBUG=$dbg timeout 200 python bench.py --steps 5 --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('debug=$dbg', 'ms/step', round(d['ms_per_step'],2), 'TF', round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; done
timeout 400 python bench.py --steps 10 > gpurun_out/r10_bench.json 2> gpurun_out/r10_bench.err; tail -8 gpurun_out/r10_bench.err
