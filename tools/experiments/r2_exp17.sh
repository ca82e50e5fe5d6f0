# round 2, call 17: consolidation with the triangle of band b+1 under the decision chain of band b
set -u
mkdir -p gpurun_out
HIPPO_CONS_OVERLAP=0 timeout 600 python -m pytest tests/test_gpu_consolidation.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -x 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_consolidation.py "tests/test_gpu_fullsize.py::test_consolidation_100k_equals_blocked_oracle" -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -x 2>&1 | tail -5
python - <<'PY'
import os, sys, time, ctypes
sys.path.insert(0, os.getcwd())
import torch
from hippomm_b200 import synth, _lib
from hippomm_b200.consolidation import select_key_frames_device
dev = torch.device("cuda", 0)
feats = synth.videolike_features_torch(3, 2000, 50, dev)
fb = feats.to(torch.bfloat16).to(torch.float32).contiguous()
def timed(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for name, f in (("bf16_exact", fb), ("fp32", feats)):
    for gamma in (0.9, 0.95):
        for ov in ("1", "0"):
            os.environ["HIPPO_CONS_OVERLAP"] = ov
            t = timed(lambda: select_key_frames_device(f, gamma))
            k, c, s = select_key_frames_device(f, gamma)
            print(f"{name} gamma {gamma} overlap {ov}: {t:.3f} ms kept {int(c.item())} stats {s.tolist()}")
os.environ["HIPPO_CONS_OVERLAP"] = "1"
os.environ["HIPPO_CONS_TIMING"] = "1"
select_key_frames_device(fb, 0.9); torch.cuda.synchronize()
out = (ctypes.c_double * 4)(); _lib.load().hippo_debug_consolidate_timing(out); print("stages (tensor, recheck, scan, build+compact):", list(out))
PY
