# flow-control window of the batched-search kernel: step time and DRAM bytes per launch
for W in 2 3 4 6 8; do
  HIPPO_TC_WINDOW=$W python bench.py --steps 10 --warmup 3 --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('window $W: ms/step', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'])"
  HIPPO_TC_WINDOW=$W timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sim_tc_kernel -s 3 -c 2 --csv --log-file gpurun_out/r2_traffic_w$W.csv python bench.py --steps 2 --warmup 1 --no-extra > /dev/null 2>&1
  python - <<PY
import csv
tot={}
for r in csv.DictReader(l for l in open("gpurun_out/r2_traffic_w$W.csv") if l.startswith('"')):
    v=float(r["Metric Value"].replace(",","")); u=r["Metric Unit"].lower()
    v*={"byte":1,"kbyte":1e3,"mbyte":1e6,"gbyte":1e9,"tbyte":1e12}[u]
    tot[r["ID"]]=tot.get(r["ID"],0)+v
print("window $W: DRAM GB per launch", [round(x/1e9,2) for x in tot.values()])
PY
done
