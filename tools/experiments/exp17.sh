set -u
mkdir -p gpurun_out
./tools/probes/cluster_occ
for dbg in 0 128 384 640; do
HIPPO_TC_DEBUG=$dbg timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:sim_tc_kernel -s 3 -c 1 --csv --log-file gpurun_out/dram_dbg$dbg.csv python bench.py --steps 1 --no-extra > /dev/null 2>&1; echo "ncu dbg$dbg rc $?"; grep -E "dram__bytes|duration|hit_rate" gpurun_out/dram_dbg$dbg.csv | cut -d, -f 13- | tr '\n' ' '; echo
HIPPO_TC_DEBUG=$dbg timeout 200 python bench.py --steps 5 --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('debug=$dbg', 'ms/step', round(d['ms_per_step'],2), 'TF', round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
