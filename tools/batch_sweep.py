#!/usr/bin/env python
"""Search latency against batch size on the 10M x 1024 lattice bank (GEMV for 1 query, the two-query GEMV behind
the batched entry, tcgen05 beyond; HIPPO_FEW_QUERIES=0 sends the small batches through tcgen05 too; SWEEP=1,2,4 picks sizes)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from hippomm_b200 import MemoryBank, synth  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
device = torch.device("cuda", 0)
torch.cuda.set_device(0)
bank = MemoryBank(rows, bench.DIM, device=device, row_base=0)
bench.build_bank(bank, rows, 0, rows, device)
q_host, _ = synth.lattice_queries_np(bench.SEED, bench.NQ, bench.DIM, rows)
q_dev = torch.from_numpy(q_host).to(device)
out = []
sizes = [int(x) for x in os.environ.get("SWEEP", "1,2,3,4,8,16,64,256,512,1024,4096").split(",")]
for nq, path in [(1, "single")] + [(x, "batched") for x in sizes]:
    q = q_dev[:nq].contiguous()
    for _ in range(3):
        bank.search_keys(q, 10, path)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        bank.search_keys(q, 10, path)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    gb = (rows * bench.DIM * 2 + rows * 4) / 1e9
    out.append({"queries": nq, "path": path, "ms": ms, "queries_per_s": nq / ms * 1e3, "bank_GBps": gb / ms * 1e3,
                "tflops": 2.0 * nq * rows * bench.DIM / ms / 1e9})
    print(out[-1], file=sys.stderr)
print(json.dumps(out))
