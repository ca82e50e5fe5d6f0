python - <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import bench
from hippomm_b200 import synth
from hippomm_b200.events import EventBank
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
torch.zeros(1, device=dev); torch.cuda.synchronize()
events = []
for e in range(2000):
    vis = synth.videolike_features(5000 + e, 5, 50).astype(np.float32)
    at = 300.0 * e + np.arange(250, dtype=np.float64)
    events.append(bench._BenchEvent({"vision": vis}, {"vision_times": at}, ["f"] * 83, at[:83].tolist()))
for keep in (True, False, True, False):
    t0 = time.perf_counter(); eb = EventBank.from_events(events, "vision", device=dev, keep_rows=keep); torch.cuda.synchronize()
    print("from_events keep_rows", keep, round(time.perf_counter() - t0, 3), "s"); del eb
PY
