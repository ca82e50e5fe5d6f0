"""Seeded synthetic inputs shared by tests/golden/make_golden.py (reference run, build container) and
the tests (oracle on CPU, CUDA path on the GPU box).  Sizes follow SURVEY.md §8(d) config 1 where the
oracle has to finish in seconds."""
from __future__ import annotations

import functools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from hippomm_b200 import synth  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_outputs.npz")


@functools.lru_cache(maxsize=None)
def golden():
    return dict(np.load(GOLDEN))


# ------------------------------------------------------------------ search ----
@functools.lru_cache(maxsize=None)
def search_config1():
    """Config 1: 2,000 video-like rows (40 scenes x 50 frames), 64 planted queries bank[j] + 0.3 N(0, I)."""
    bank = synth.videolike_features(20250417, 40, 50)
    rng = np.random.default_rng(20250418)
    js = rng.integers(0, len(bank), size=64)
    queries = bank[js] + np.float32(0.3) * rng.standard_normal((64, bank.shape[1])).astype(np.float32)
    return bank, queries.astype(np.float32)


@functools.lru_cache(maxsize=None)
def search_lattice_small():
    """Lattice bank (bit-exact arithmetic): 8,192 rows x 1024, 32 queries."""
    n, d, seed = 8192, 1024, 4
    bank = synth.lattice_rows_np(seed, np.arange(n), d, n)
    q, fam = synth.lattice_queries_np(seed, 32, d, n)
    return bank, q, fam


@functools.lru_cache(maxsize=None)
def search_edge():
    """16 x 64 bank with a zero-norm row (NaN score), exact duplicates (ties) and anti-correlated rows."""
    rng = np.random.default_rng(7)
    b = rng.standard_normal((16, 64)).astype(np.float32)
    q = rng.standard_normal(64).astype(np.float32)
    b[3] = 0.0            # NaN similarity: np.argsort puts it last, i.e. first in the result
    b[9] = b[2]           # exact tie
    b[11] = -q            # similarity -1
    b[5] = 2.0 * q        # similarity +1
    # bf16-representable values: the device bank holds bf16 rows, and at d = 64 the rounding of
    # arbitrary fp32 rows would move a score by up to ~1e-3 (it averages down to ~1e-4 at d = 1024)
    q = synth.round_to_bf16(q)
    b = synth.round_to_bf16(b)
    b[9], b[11], b[5] = b[2], -q, 2.0 * q   # (power-of-two scale keeps the row bf16-exact)
    return b, q


# ----------------------------------------------------------- consolidation ----
@functools.lru_cache(maxsize=None)
def consolidation_cases():
    c1 = synth.videolike_features(20250417, 40, 50)
    rng = np.random.default_rng(11)
    centres = rng.standard_normal((30, 1024)).astype(np.float32)
    clustered = np.repeat(centres, 20, axis=0) + np.float32(0.02) * rng.standard_normal((600, 1024)).astype(np.float32)
    zero = synth.videolike_features(5, 4, 10).copy()
    zero[17] = 0.0        # a zero-norm row is never kept (NaN similarities, hm:960)
    zero0 = synth.videolike_features(6, 3, 10).copy()
    zero0[0] = 0.0        # zero-norm row 0: kept, and nothing after it ever is
    dup = synth.videolike_features(8, 3, 6).copy()
    dup[7] = dup[6]
    return {
        "c1": c1,
        "c1bf": synth.round_to_bf16(c1),
        "clustered": clustered.astype(np.float32),
        "three": synth.videolike_features(9, 1, 3, step=0.5),
        "zero": zero,
        "zero0": zero0,
        "dup": dup,
        "d200": synth.videolike_features(10, 6, 25, d=200),   # d not a multiple of 64
    }


# ------------------------------------------------------------------- audio ----
@functools.lru_cache(maxsize=None)
def audio_case():
    """300 s of int16 PCM at 16 kHz: noise at -20 dBFS with planted silences."""
    return synth.audio_stream_int16(2, 300 * 16000)


def audio_windows(n):
    wins = [(0, 8000), (8000, 8000), (12345, 8000), (160001, 8000), (n - 8000, 8000), (17, 100), (511, 1025),
            (4096, 512), (1000, 15), (0, n), (n - 1, 1)]
    rng = np.random.default_rng(3)
    for _ in range(40):
        s = int(rng.integers(0, n - 8000))
        wins.append((s, 8000))
    return wins


# ------------------------------------------------------------------ frames ----
@functools.lru_cache(maxsize=None)
def frame_case():
    """200 frames of 64 x 64 BGR at 1 fps: piecewise-static scenes of 5-30 s plus sensor noise."""
    return synth.frame_stream(1, 200, 64, 64, min_scene=5, max_scene=30)


def frame_times(n):
    return [float(i) for i in range(n)]


def frame_diff_pairs():
    return [(0, 1), (1, 0), (5, 5), (30, 31), (31, 32), (10, 150), (199, 0)]


@functools.lru_cache(maxsize=None)
def constant_frames():
    a = np.full((32, 40, 3), 10, dtype=np.uint8)
    b = np.full((32, 40, 3), 10, dtype=np.uint8)
    c = np.full((32, 40, 3), 200, dtype=np.uint8)
    return [a, b, c]


def segmentation_variants():
    dflt = dict(max_segment_duration=30.0, min_segment_duration=10.0, frame_similarity_threshold=0.95,
                audio_silence_threshold=-40)
    short = dict(max_segment_duration=10.0, min_segment_duration=5.0, frame_similarity_threshold=0.95,
                 audio_silence_threshold=-40)        # the in-code fallbacks of hm:263-266
    return {
        "av": dict(video=True, audio=True, thresholds=dflt),
        "a": dict(video=False, audio=True, thresholds=dflt),
        "v": dict(video=True, audio=False, thresholds=dflt),
        "av_short": dict(video=True, audio=True, thresholds=short),
        "v_short": dict(video=True, audio=False, thresholds=short),
    }


# ------------------------------------------------------------------ recall ----
class _Event:
    """The fields of the reference's ThetaEvent (hm:95-108) that the recall loops read."""

    def __init__(self, features, feature_times, frames, frame_times):
        self.features = features
        self.feature_times = feature_times
        self.frames = frames
        self.frame_times = frame_times
        self.frame_captions = []                     # falsy: the LLM branch of hm:3156 / hm:3307 is never taken
        self.audio_times = []
        self.audio_transcription = []
        self.holistic_audio_transcription = []
        self.summary = ""
        self.start_time = frame_times[0] if frame_times else 0.0
        self.end_time = frame_times[-1] if frame_times else 0.0


@functools.lru_cache(maxsize=None)
def recall_events():
    """14 events of 3..90 all-frame vision rows (1024-d, video-like) with SHORTER key-frame time tables
    (the index-space quirk of hm:3262), audio rows with their own times, one event without audio and one
    with fewer than 5 rows; two queries planted near rows of events 3 and 9."""
    rng = np.random.default_rng(77)
    events, t0 = [], 0.0
    sizes = [40, 3, 25, 60, 12, 90, 7, 33, 5, 48, 20, 4, 70, 16]
    for e, n in enumerate(sizes):
        vis = synth.videolike_features(1000 + e, max(1, n // 10), min(n, 10))[:n]
        if len(vis) < n:
            vis = np.concatenate([vis, synth.videolike_features(2000 + e, 1, n - len(vis))])
        n_key = max(1, int(n * (0.3 + 0.5 * rng.random())))        # key frames: a subset, so some hits fall outside
        all_times = t0 + np.arange(n, dtype=np.float64)
        key_times = sorted(rng.choice(all_times, size=n_key, replace=False).tolist())
        feats = {"vision": vis.astype(np.float32)}
        ftimes = {"vision_times": all_times.copy()}
        if e != 6:
            na = max(2, n // 3)
            feats["audio"] = synth.videolike_features(3000 + e, 1, na, step=0.4).astype(np.float32)
            ftimes["audio_times"] = t0 + 3.0 * np.arange(na, dtype=np.float64) + 0.5
        frames = [f"/frames/e{e:02d}_{i:04d}.jpg" for i in range(n_key)]
        events.append(_Event(feats, ftimes, frames, key_times))
        t0 += n + 5.0
    d = 1024
    noise = lambda: np.float32(0.2) * rng.standard_normal(d).astype(np.float32)
    queries = {
        # hits inside the key-frame table of one event
        "v_in": ("vision", events[3].features["vision"][7] + noise()),
        # planted at the LAST all-frame row of the 90-row event: beyond its key-frame table, dropped by hm:3262
        "v_out": ("vision", events[5].features["vision"][89] + noise()),
        # between two events: the final five mix events
        "v_mix": ("vision", events[7].features["vision"][4] + events[12].features["vision"][30] + noise()),
        "a_in": ("audio", events[9].features["audio"][2] + noise()),
        "a_mix": ("audio", events[2].features["audio"][1] + events[13].features["audio"][3] + noise()),
    }
    return events, {k: (m, q.astype(np.float32)) for k, (m, q) in queries.items()}


# ----------------------------------------------------------- greedy frame filters ----
@functools.lru_cache(maxsize=None)
def prefilter_frames():
    """360 frames of 96 x 96 BGR 'decoded at 30 fps' (12 s): scenes of 1-3 s with a slow brightness drift
    inside each scene, so that both the single-difference and the cumulative trigger of bp:198-200 fire."""
    base, cuts = synth.frame_stream(21, 12, 96, 96, min_scene=1, max_scene=3)
    rng = np.random.default_rng(22)
    frames = np.empty((360, 96, 96, 3), dtype=np.uint8)
    for i in range(360):
        f = base[i // 30].astype(np.int16)
        drift = int(round(2.5 * (i % 90) / 10.0))
        noise = rng.integers(-2, 3, size=f.shape, dtype=np.int16)
        frames[i] = np.clip(f + drift + noise, 0, 255).astype(np.uint8)
    return frames


PREFILTER_PARAMS = dict(video_fps=30.0, max_diff_threshold=0.3, check_interval=10)


def write_mjpg(frames, path, fps=30.0):
    """Write frames as an MJPG AVI with OpenCV's built-in encoder (no ffmpeg needed)."""
    import cv2

    h, w = frames.shape[1:3]
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), fps, (w, h))
    if not wr.isOpened():
        return False
    for f in frames:
        wr.write(f)
    wr.release()
    return True


def read_video(path):
    import cv2

    cap = cv2.VideoCapture(path)
    out = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        out.append(f)
    cap.release()
    return np.stack(out) if out else np.zeros((0, 1, 1, 3), dtype=np.uint8)


@functools.lru_cache(maxsize=None)
def dedup_windows():
    """Windows of 3-6 frames of 180 x 320 (the size hm:2230 resizes to): near-duplicates and changes."""
    sc = [synth.frame_stream(31 + s, 1, 180, 320)[0][0] for s in range(4)]
    blend = np.clip(0.3 * sc[0].astype(np.float64) + 0.7 * sc[1], 0, 255).astype(np.uint8)   # SSIM to sc[0] ~ 0.36:
    rng = np.random.default_rng(32)                                                           # dropped at 0.3, kept at 0.4
    wins = []
    for order in ([0, 0, 1], [1, 2, 2, 3], [3, 3, 3], [0, 1, 0, 1, 1, 2], [0, "b", 1], [0, "b", "b", 0]):
        fr = []
        for j in order:
            src = blend if j == "b" else sc[j]
            f = src.astype(np.int16) + rng.integers(-1, 2, size=src.shape, dtype=np.int16)
            fr.append(np.clip(f, 0, 255).astype(np.uint8))
        wins.append(np.stack(fr))
    wins.append(np.stack([np.full((180, 320, 3), 9, np.uint8)] * 3))   # constant frames: SSIM is NaN, nothing dropped
    return wins
