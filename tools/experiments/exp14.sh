timeout 300 python -m pytest tests/test_gpu_search.py tests/test_gpu_consolidation.py -m gpu -q --tb=short --timeout 120 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -4
for dbg in 0 8; do HIPPO_TC_DEBUG=$dbg timeout 200 python bench.py --steps 5 --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('debug=$dbg', 'ms/step', round(d['ms_per_step'],2), 'TF', round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; done
python tools/prof_counters.py 2>&1 | tail -2
