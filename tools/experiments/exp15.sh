# where does the MMA thread wait? counters + no-slow-path + no-epilogue + no-TMA variants
python tools/prof_counters.py 2>&1 | tail -2
for dbg in 0 8 1 16; do HIPPO_TC_DEBUG=$dbg timeout 200 python bench.py --steps 5 --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('debug=$dbg', 'ms/step', round(d['ms_per_step'],2), 'TF', round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; done
