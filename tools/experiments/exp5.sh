for ns in 0 500 2000 8000; do HIPPO_TC_STAGGER_NS=$ns timeout 200 python bench.py --steps 5 --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('stagger_ns=$ns', 'ms/step', round(d['ms_per_step'],2), 'TF', round(d['roofline']['achieved'],1), d['clocks'])"; done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sim_tc_kernel -s 2 -c 1 -f -o gpurun_out/r5_sim_tc python bench.py --steps 1 --no-extra --bank-rows 2000000 > gpurun_out/r5_ncu_full.log 2>&1; echo ncu rc $?
