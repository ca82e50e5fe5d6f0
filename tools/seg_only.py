#!/usr/bin/env python
"""Development helper: segmentation of config 2's synthetic stream-hour only, timed per stage with CUDA events
(gray + SSIM, audio pyramid, boundary state machine).  HIPPO_SEG_DEBUG=1 prints the boundary kernel's cycle split."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from hippomm_b200.segmentation import (audio_energy_device, frame_pair_scores_device,  # noqa: E402
                                       pattern_separation_batch_device, pattern_separation_device,
                                       segment_boundaries_device)

device = torch.device("cuda", 0)
torch.cuda.set_device(0)
frames, pcm, ft = bench.synth_stream_hour(device)
sr = 16000


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ssim, mse = frame_pair_scores_device(frames, range_mode=0)
pyr = audio_energy_device(pcm)
out = segment_boundaries_device(ssim, ft, pcm, pyr, sr, 30.0, 10.0, 0.95, -40.0, 512)
torch.cuda.synchronize()
n = int(out[1].item())
digest = hashlib.sha1(out[0][:n].cpu().numpy().tobytes() + ssim.cpu().numpy().tobytes()).hexdigest()[:16]
dbg = os.environ.pop("HIPPO_SEG_DEBUG", None)      # the debug path synchronises: keep it out of the timings
t_pairs = timed(lambda: frame_pair_scores_device(frames, range_mode=0))
t_audio = timed(lambda: audio_energy_device(pcm))
t_seg = timed(lambda: segment_boundaries_device(ssim, ft, pcm, pyr, sr, 30.0, 10.0, 0.95, -40.0, 512))


def whole():
    s, _ = frame_pair_scores_device(frames, range_mode=0)
    p = audio_energy_device(pcm)
    segment_boundaries_device(s, ft, pcm, p, sr, 30.0, 10.0, 0.95, -40.0, 512)


t_all = timed(whole)
for cp in [int(x) for x in os.environ.get("CHUNKS", "444").split(",")]:
    t_ov = timed(lambda: pattern_separation_device(frames, ft, pcm, sr, 30.0, 10.0, 0.95, -40.0, 512, chunk_pairs=cp))
    o2 = pattern_separation_device(frames, ft, pcm, sr, 30.0, 10.0, 0.95, -40.0, 512, chunk_pairs=cp)
    torch.cuda.synchronize()
    same = int(o2[1].item()) == n and torch.equal(o2[0][:n], out[0][:n])
    print(f"[seg_only] overlapped pipeline, chunks of {cp} pairs: {t_ov:.3f} ms (identical: {same})")
if os.environ.get("TAIL"):
    for nfr in (3553, 3600, 3109, 2665):
        sub = frames[:nfr].contiguous()
        t = timed(lambda: frame_pair_scores_device(sub, range_mode=0))
        print(f"[seg_only] frame pairs on {nfr} frames: {t:.3f} ms")
if os.environ.get("TIMELINE"):
    os.environ["HIPPO_PATTERN_DEBUG"] = "1"
    pattern_separation_device(frames, ft, pcm, sr, 30.0, 10.0, 0.95, -40.0, 512, chunk_pairs=int(os.environ["TIMELINE"]))
    torch.cuda.synchronize()
    del os.environ["HIPPO_PATTERN_DEBUG"]
if os.environ.get("BATCH"):
    nb = int(os.environ["BATCH"])
    for mode in ("stages", "pipeline"):
        for lanes in (1, 2, 3):
            t_b = timed(lambda: pattern_separation_batch_device([(frames, ft, pcm, sr)] * nb, 30.0, 10.0, 0.95, -40.0, 512, lanes=lanes, mode=mode), iters=2, warm=1)
            print(f"[seg_only] batch of {nb}, {mode}, {lanes} lanes: {t_b / nb:.3f} ms per stream-hour")
print(f"[seg_only] segments {n} digest {digest}  frame pairs {t_pairs:.3f} ms  audio pyramid {t_audio:.3f} ms  "
      f"boundaries {t_seg:.3f} ms  whole {t_all:.3f} ms")
if dbg:
    os.environ["HIPPO_SEG_DEBUG"] = dbg
    segment_boundaries_device(ssim, ft, pcm, pyr, sr, 30.0, 10.0, 0.95, -40.0, 512)
    torch.cuda.synchronize()
if os.environ.get("HOSTTIME"):
    import time
    nb = int(os.environ["HOSTTIME"])
    for mode, lanes in (("stages", 3), ("stages", 1), ("pipeline", 2)):
        for _ in range(2):
            pattern_separation_batch_device([(frames, ft, pcm, sr)] * nb, 30.0, 10.0, 0.95, -40.0, 512, lanes=lanes, mode=mode)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pattern_separation_batch_device([(frames, ft, pcm, sr)] * nb, 30.0, 10.0, 0.95, -40.0, 512, lanes=lanes, mode=mode)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"[seg_only] host issue time, batch of {nb}, {mode}, {lanes} lanes: {(t1 - t0) / nb * 1e3:.3f} ms per stream issued, "
              f"{(t2 - t0) / nb * 1e3:.3f} ms per stream until the GPU is done")
