set -u
for dbg in 0 8 1; do HIPPO_TC_DEBUG=$dbg timeout 200 python bench.py --steps 8 --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('debug=$dbg', 'ms/step', round(d['ms_per_step'],2), 'TF', round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; done
python tools/prof_counters.py 2>&1 | tail -1
QUERIES=gauss python tools/prof_counters.py 2>&1 | tail -1
