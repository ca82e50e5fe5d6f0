export HIPPO_TC_PAIRS=1
for dbg in 1 17 33 49; do HIPPO_TC_DEBUG=$dbg timeout 200 python bench.py --steps 5 --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('debug=$dbg', 'ms/step', round(d['ms_per_step'],2), 'TF', round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; done
