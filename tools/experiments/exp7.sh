for P in 1 2 4; do
  export HIPPO_TC_PAIRS=$P
  timeout 300 python -m pytest tests/test_gpu_search.py tests/test_gpu_consolidation.py -m gpu -q --tb=line --timeout 120 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -4
  timeout 200 python bench.py --steps 5 --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('PAIRS=$P', 'ms/step', round(d['ms_per_step'],2), 'TF', round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
