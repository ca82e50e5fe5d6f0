import sys, os
sys.path.insert(0, os.getcwd())
import torch
from hippomm_b200.consolidation import select_key_frames_device
device = torch.device("cuda", 0)
n_scenes, fps, DIM = 2000, 50, 1024
feats = torch.empty((n_scenes * fps, DIM), dtype=torch.float32, device=device)
g = torch.Generator(device=device); g.manual_seed(3)
for s0 in range(0, n_scenes, 200):
    v = torch.randn((200, DIM), generator=g, device=device)
    for f in range(fps):
        feats[(s0 * fps + f)::fps][:200] = v
        v = v + 0.12 * torch.randn((200, DIM), generator=g, device=device)
feats = feats.to(torch.bfloat16).to(torch.float32).contiguous()
os.environ.pop("HIPPO_SCAN_DEBUG", None)
for _ in range(2): select_key_frames_device(feats, 0.9)
torch.cuda.synchronize()
os.environ["HIPPO_SCAN_DEBUG"] = "1"
select_key_frames_device(feats, 0.9)
torch.cuda.synchronize()
