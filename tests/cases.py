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
