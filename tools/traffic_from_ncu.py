#!/usr/bin/env python
"""Turn an ncu CSV (`--metrics dram__bytes_read.sum,dram__bytes_write.sum --csv`) of bench.py's top-k kernel
into profiles/traffic.json, which bench.py reports as roofline.traffic (bytes per launch)."""
import csv
import json
import os
import sys

src, rows_arg = sys.argv[1], int(sys.argv[2])
tot = {}
for r in csv.DictReader(l for l in open(src) if l.startswith('"')):
    if "sim_tc_kernel<0>" not in r["Kernel Name"] and "sim_tc_kernel<(int)0>" not in r["Kernel Name"]:
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"].lower()
    v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[unit]
    tot.setdefault(r["ID"], 0.0)
    tot[r["ID"]] += v
vals = sorted(tot.values())
out = {"sim_tc_kernel<EPI_TOPK>": {"bank_rows": rows_arg, "dram_bytes_per_launch": vals[len(vals) // 2],
                                   "launches": len(vals), "source": os.path.basename(src),
                                   "metric": "dram__bytes_read.sum + dram__bytes_write.sum (median over launches)"}}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump(out, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
print(out)
