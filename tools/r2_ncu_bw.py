#!/usr/bin/env python
"""Round-2 evidence helper: ONE launch each of bank_build_kernel<float> (whole bank from fp32 rows),
topk_single_kernel<4> and topk_few_kernel<2> at ROWS x 1024 inside a cudaProfilerStart/Stop range, so
`ncu --profile-from-start off --set full` captures exactly these (and nothing of the bank synthesis)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from hippomm_b200 import MemoryBank, synth  # noqa: E402

rows = int(os.environ.get("ROWS", 10_000_000))
device = torch.device("cuda", 0)
torch.cuda.set_device(0)
src = torch.empty((rows, bench.DIM), dtype=torch.float32, device=device)      # 41 GB at 10M rows
for r0 in range(0, rows, 1 << 16):
    m = min(1 << 16, rows - r0)
    synth.lattice_rows_torch(bench.SEED, r0, m, bench.DIM, rows, device, out=src[r0:r0 + m])
bank = MemoryBank(rows, bench.DIM, device=device)
q_host, _ = synth.lattice_queries_np(bench.SEED, 8, bench.DIM, rows)
q = torch.from_numpy(q_host).to(device)
for _ in range(2):                    # warm-up outside the profiled range
    bank.fill(0, src)
    bank.search_keys(q[:1], 10, "single")
    bank.search_keys(q[:2], 10, "batched")
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
torch.cuda.profiler.start()
ev[0].record()
bank.fill(0, src)
ev[1].record()
bank.search_keys(q[:1], 10, "single")
ev[2].record()
bank.search_keys(q[:2], 10, "batched")
ev[3].record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
gb_build = rows * bench.DIM * 6 / 1e9 + rows * 4 / 1e9
gb_pass = rows * bench.DIM * 2 / 1e9 + rows * 4 / 1e9
t = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
print(f"[r2_ncu_bw] rows {rows}: bank_build {t[0]:.3f} ms ({gb_build / t[0] * 1e3:.0f} GB/s of {gb_build:.2f} GB), "
      f"single {t[1]:.3f} ms ({gb_pass / t[1] * 1e3:.0f} GB/s), two-query {t[2]:.3f} ms ({gb_pass / t[2] * 1e3:.0f} GB/s)")
