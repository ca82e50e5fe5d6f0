set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_search.py tests/test_gpu_consolidation.py -m gpu -q --tb=short --timeout 200 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -4
for w in 0 4 8 16 32; do
HIPPO_TC_WINDOW=$w timeout 200 python bench.py --steps 5 --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('window=$w', 'ms/step', round(d['ms_per_step'],2), 'TF', round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
for w in 0 8; do
HIPPO_TC_WINDOW=$w timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:sim_tc_kernel -s 3 -c 1 --csv --log-file gpurun_out/dram_w$w.csv python bench.py --steps 1 --no-extra > /dev/null 2>&1; echo "ncu w$w rc $?"; grep -E "dram__bytes|duration|hit_rate" gpurun_out/dram_w$w.csv | cut -d, -f 15- | tr '\n' ' '; echo
done
timeout 400 python bench.py --steps 3 2>&1 >/dev/null | grep extra
