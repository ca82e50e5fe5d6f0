#!/usr/bin/env python
"""Development helper: consolidation of bench.py's 100k video-like rows against the band size (HIPPO_CONS_BAND)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from hippomm_b200.consolidation import select_key_frames_device  # noqa: E402

device = torch.device("cuda", 0)
torch.cuda.set_device(0)
DIM, n_scenes, fps = 1024, 2000, 50
feats = torch.empty((n_scenes * fps, DIM), dtype=torch.float32, device=device)
g = torch.Generator(device=device)
g.manual_seed(3)
for s0 in range(0, n_scenes, 200):
    v = torch.randn((200, DIM), generator=g, device=device)
    for f in range(fps):
        feats[(s0 * fps + f)::fps][:200] = v
        v = v + 0.12 * torch.randn((200, DIM), generator=g, device=device)
feats_bf = feats.to(torch.bfloat16).to(torch.float32).contiguous()
ref = {}
for band in [int(x) for x in os.environ.get("BANDS", "8192,6144,4096,3072,2048").split(",")]:
    os.environ["HIPPO_CONS_BAND"] = str(band)
    for name, f_ in (("bf16_exact", feats_bf), ("fp32", feats)):
        for gamma in (0.9, 0.95):
            for _ in range(2):
                out = select_key_frames_device(f_, gamma)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                out = select_key_frames_device(f_, gamma)
            e1.record()
            torch.cuda.synchronize()
            kept = out[0][: int(out[1].item())].cpu()
            key = (name, gamma)
            same = True if key not in ref else bool(torch.equal(ref[key], kept))
            ref.setdefault(key, kept)
            print(f"band {band:5d} {name:10s} gamma {gamma}: {e0.elapsed_time(e1) / 5:.3f} ms, kept {len(kept)}, same as first band: {same}")
