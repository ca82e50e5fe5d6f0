#!/usr/bin/env python
"""Development helper: run bench.py's secondary measurements (single-query search, consolidation 100k,
segmentation of a 1-hour stream) against a small bank, without the 10M-row headline step."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from hippomm_b200 import MemoryBank, _lib, synth  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
device = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = _lib.load()
bank = MemoryBank(rows, bench.DIM, device=device, row_base=0)
bench.build_bank(bank, rows, 0, rows, device)
q_host, _ = synth.lattice_queries_np(bench.SEED, bench.NQ, bench.DIM, rows)
q_dev = torch.from_numpy(q_host).to(device)
print(json.dumps(bench.run_extras(bank, q_dev, bench.load_peaks(), device, lib), indent=1))
