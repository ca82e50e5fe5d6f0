"""Profiling helper: run the batched search once with HIPPO_TC_DEBUG=64 and print the kernel's counters."""
import os, sys, time
os.environ["HIPPO_TC_DEBUG"] = str(64 | int(os.environ.get("EXTRA_DEBUG", "0")))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hippomm_b200 import MemoryBank, synth, _cuda
import bench
n = int(os.environ.get("ROWS", 10_000_000)); d = 1024; nq = int(os.environ.get("NQ", 4096))
dev = torch.device("cuda", 0)
bank = MemoryBank(n, d, device=dev)
bench.build_bank(bank, n, 0, n, dev)
kind = os.environ.get("QUERIES", "lattice")
if kind == "lattice":
    q, fam = synth.lattice_queries_np(4, nq, d, n)
else:
    q = np.random.default_rng(0).standard_normal((nq, d)).astype(np.float32)
qd = torch.from_numpy(q).to(dev)
for it in range(3):
    ws = _cuda.workspace(1, dev, "topk")
    torch.cuda.synchronize()
    if ws.numel() >= 64: ws[:64].zero_()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    bank.search(qd, 10, "batched"); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ws = _cuda.workspace(1, dev, "topk")
    c = ws[:64].view(torch.int64).cpu().numpy()
    print(f"iter {it}: {dt*1e3:.1f} ms  slow_chunk_calls={c[0]} insertions={c[1]} slow_cycles={c[2]/1e6:.1f}M "
          f"epi_wait_tfull={c[3]/1e6:.1f}M over {c[4]} warp-tiles ({c[3]/max(c[4],1):.0f} clk each)  "
          f"mma_wait_tmem={c[5]/1e6:.1f}M mma_wait_operands={c[6]/1e6:.1f}M (74 leaders)")
