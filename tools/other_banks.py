#!/usr/bin/env python
"""Development helper: bench.py's headline step on the scene-clustered and the Gaussian fp32 bank (10M x 1024), with the
kernel's slow-path / wait counters."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from hippomm_b200 import _lib  # noqa: E402

torch.cuda.set_device(0)
print(json.dumps(bench.run_other_banks(bench.load_peaks(), torch.device("cuda", 0), _lib.load()), indent=1))
