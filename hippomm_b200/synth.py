"""Counter-based synthetic data shared by the CPU oracle side (NumPy) and the GPU side (torch).

Every generator is a pure function of (seed, row, column), so a 10M- or 80M-row bank is
produced in place on each device and never crosses PCIe or the gpurun snapshot, while
the oracle regenerates exactly the rows it needs on the host.

Lattice bank ("bf16-exact"):  value = k / 128 with integer |k| <= 127.
  * row r belongs to family f = r % F (F families) and is member m = r // F of it;
  * k(r, c) = base(f, c) + noise(r, c), base uniform in [-64, 64], noise uniform in [-L_m, L_m]
    with L_m = 3 m for m <= 9 and L_m = 63 for m >= 10: member 0 is the family centre, members
    1..9 are graded copies (expected cosine to a query of the family about .997, .994, .988,
    .980, .971, .960, .948, .934, .919), members >= 10 are far copies (about .71);
  * a query for family f is the centre plus uniform noise in [-3, 3].
Every product of two lattice values is a multiple of 2^-14 and every dot product / squared
norm stays below 2^24 such units, so fp32 accumulation is exact in ANY order: the reference's
NumPy scores, the GEMV kernel and the tensor-core kernel agree to the last bit.  The true
top-10 of a query is its family's members 0..9 as a SET (8 sigma clear of member 10 and of
unrelated rows, which score about 0 +- .03); the order inside the set follows the grading
only on average.
"""
from __future__ import annotations

import numpy as np

try:  # torch is only needed for the device-side generators
    import torch
except Exception:  # pragma: no cover
    torch = None

_M32 = 0xFFFFFFFF
_LEVEL_STEP = 3
_NEAR_MEMBERS = 10
_FAR_AMP = 63
_BASE_AMP = 64
_QUERY_AMP = 3


# ------------------------------------------------------------------ hashing ----
def _mix_np(x: np.ndarray) -> np.ndarray:
    """murmur3 fmix32 on uint64 arrays holding 32-bit values."""
    x = x & _M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x85EBCA6B)) & _M32
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE35)) & _M32
    x ^= x >> np.uint64(16)
    return x


def _hash_np(seed: int, idx: np.ndarray) -> np.ndarray:
    idx = idx.astype(np.uint64)
    lo = idx & _M32
    hi = (idx >> np.uint64(32)) & _M32
    s = np.uint64((seed * 0x9E3779B1 + 0x7F4A7C15) & _M32)
    return _mix_np(lo ^ _mix_np(hi ^ s))


def _mix_t(x):
    x = x & _M32
    x = x ^ (x >> 16)
    x = (x * 0x85EBCA6B) & _M32
    x = x ^ (x >> 13)
    x = (x * 0xC2B2AE35) & _M32
    x = x ^ (x >> 16)
    return x


def _hash_t(seed: int, idx):
    lo = idx & _M32
    hi = (idx >> 32) & _M32
    s = (seed * 0x9E3779B1 + 0x7F4A7C15) & _M32
    return _mix_t(lo ^ _mix_t(hi ^ s))


def _uniform_int_np(h: np.ndarray, amp: np.ndarray | int) -> np.ndarray:
    """h (32-bit hash) -> integer uniform in [-amp, amp]."""
    span = (2 * np.asarray(amp, dtype=np.int64) + 1).astype(np.uint64)
    return (h % span).astype(np.int64) - np.asarray(amp, dtype=np.int64)


def lattice_families(n_rows: int) -> int:
    """Number of families F for a bank of n_rows (16 members per family, at least 1)."""
    return max(1, n_rows // 16)


# ----------------------------------------------------------- lattice bank ----
def lattice_rows_np(seed: int, rows: np.ndarray, d: int, n_total: int) -> np.ndarray:
    """fp32 [len(rows), d] lattice rows of the bank of `n_total` rows."""
    rows = np.asarray(rows, dtype=np.int64)
    F = lattice_families(n_total)
    fam = rows % F
    mem = rows // F
    cols = np.arange(d, dtype=np.int64)[None, :]
    base = _uniform_int_np(_hash_np(seed, fam[:, None] * d + cols), _BASE_AMP)
    lvl = np.where(mem < _NEAR_MEMBERS, _LEVEL_STEP * mem, _FAR_AMP)[:, None]
    noise = _uniform_int_np(_hash_np(seed + 1, rows[:, None] * d + cols), lvl)
    return ((base + noise).astype(np.float32)) / np.float32(128.0)


def lattice_queries_np(seed: int, n_queries: int, d: int, n_total: int) -> tuple[np.ndarray, np.ndarray]:
    """(queries fp32 [nq, d], family id of each query)."""
    F = lattice_families(n_total)
    qi = np.arange(n_queries, dtype=np.int64)
    fam = (_hash_np(seed + 2, qi) % np.uint64(F)).astype(np.int64)
    cols = np.arange(d, dtype=np.int64)[None, :]
    base = _uniform_int_np(_hash_np(seed, fam[:, None] * d + cols), _BASE_AMP)
    noise = _uniform_int_np(_hash_np(seed + 3, qi[:, None] * d + cols), _QUERY_AMP)
    return ((base + noise).astype(np.float32)) / np.float32(128.0), fam


def lattice_expected_topk(fam: np.ndarray, n_total: int, k: int) -> np.ndarray:
    """Rows of members 0..k-1 of each query's family: for k = 10 (and n_total >= 160) exactly the SET of
    the query's top-10 rows; the order inside the set is only graded on average."""
    F = lattice_families(n_total)
    return fam[:, None] + F * np.arange(k, dtype=np.int64)[None, :]


def lattice_rows_torch(seed: int, row0: int, n_rows: int, d: int, n_total: int, device, out=None):
    """bf16-exact fp32 rows [row0, row0 + n_rows) generated on `device` (same values as the NumPy version)."""
    F = lattice_families(n_total)
    rows = torch.arange(row0, row0 + n_rows, dtype=torch.int64, device=device)
    fam = rows % F
    mem = rows // F
    cols = torch.arange(d, dtype=torch.int64, device=device)[None, :]
    hb = _hash_t(seed, fam[:, None] * d + cols)
    base = hb % (2 * _BASE_AMP + 1) - _BASE_AMP
    lvl = torch.where(mem < _NEAR_MEMBERS, _LEVEL_STEP * mem, torch.full_like(mem, _FAR_AMP))[:, None]
    hn = _hash_t(seed + 1, rows[:, None] * d + cols)
    noise = hn % (2 * lvl + 1) - lvl
    vals = (base + noise).to(torch.float32) / 128.0
    if out is not None:
        out.copy_(vals)
        return out
    return vals


def lattice_queries_torch(seed: int, n_queries: int, d: int, n_total: int, device):
    q, fam = lattice_queries_np(seed, n_queries, d, n_total)
    return torch.from_numpy(q).to(device), fam


# ------------------------------------------------------ video-like features ----
def videolike_features(seed: int, n_scenes: int, frames_per_scene: int, d: int = 1024, step: float = 0.12,
                       bf16_exact: bool = False) -> np.ndarray:
    """Time-ordered fp32 rows: per scene a centre c ~ N(0, I) and a random walk v_t = v_{t-1} + step*N(0, I)
    (SURVEY.md §8d config 1 / 3).  bf16_exact rounds every entry to a bf16-representable fp32."""
    rng = np.random.default_rng(seed)
    out = np.empty((n_scenes * frames_per_scene, d), dtype=np.float32)
    r = 0
    for _ in range(n_scenes):
        v = rng.standard_normal(d).astype(np.float32)
        for _ in range(frames_per_scene):
            out[r] = v
            r += 1
            v = v + np.float32(step) * rng.standard_normal(d).astype(np.float32)
    if bf16_exact:
        out = round_to_bf16(out)
    return out


def round_to_bf16(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even to bf16, returned as fp32."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    rounded = ((u + np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))) >> np.uint64(16)) << np.uint64(16)
    return (rounded & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.float32).reshape(np.shape(x))


# ------------------------------------------------------------ frame stream ----
def frame_stream(seed: int, n_frames: int, h: int = 224, w: int = 224, min_scene: int = 5, max_scene: int = 60,
                 noise_sigma: float = 2.0) -> tuple[np.ndarray, np.ndarray]:
    """uint8 BGR frames [n, h, w, 3]: piecewise-static scenes (smooth random field) + per-frame noise.
    Returns (frames, scene_start flags)."""
    rng = np.random.default_rng(seed)
    frames = np.empty((n_frames, h, w, 3), dtype=np.uint8)
    cuts = np.zeros(n_frames, dtype=bool)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    f = 0
    while f < n_frames:
        length = int(rng.integers(min_scene, max_scene + 1))
        cuts[f] = True
        field = np.zeros((h, w, 3), dtype=np.float32)
        for _ in range(6):  # a few random low-frequency waves per channel
            fx, fy = rng.uniform(0.005, 0.08, size=2)
            ph = rng.uniform(0, 2 * np.pi, size=3)
            amp = rng.uniform(15, 45, size=3)
            arg = (2 * np.pi) * (fx * xx + fy * yy)
            for c in range(3):
                field[:, :, c] += amp[c] * np.sin(arg + ph[c])
        field += rng.uniform(90, 160, size=3).astype(np.float32)
        for _ in range(min(length, n_frames - f)):
            noisy = field + rng.normal(0.0, noise_sigma, size=field.shape).astype(np.float32)
            frames[f] = np.clip(np.rint(noisy), 0, 255).astype(np.uint8)
            f += 1
    return frames, cuts


# ------------------------------------------------------------ audio stream ----
def audio_stream_int16(seed: int, n_samples: int, sample_rate: int = 16000, level_db: float = -20.0,
                       silence_db: float = -80.0) -> np.ndarray:
    """int16 PCM: noise at level_db dBFS with silences of 0.6-3 s at silence_db every 8-40 s."""
    rng = np.random.default_rng(seed)
    amp = 32768.0 * 10.0 ** (level_db / 20.0)
    samp = 32768.0 * 10.0 ** (silence_db / 20.0)
    out = np.empty(n_samples, dtype=np.int16)
    pos = 0
    chunk = 1 << 20
    while pos < n_samples:
        m = min(chunk, n_samples - pos)
        out[pos:pos + m] = np.clip(np.rint(rng.normal(0.0, amp, size=m)), -32768, 32767).astype(np.int16)
        pos += m
    t = 0.0
    dur = n_samples / sample_rate
    while True:
        t += float(rng.uniform(8.0, 40.0))
        if t >= dur:
            break
        length = float(rng.uniform(0.6, 3.0))
        a, b = int(t * sample_rate), min(int((t + length) * sample_rate), n_samples)
        out[a:b] = np.clip(np.rint(rng.normal(0.0, samp, size=b - a)), -32768, 32767).astype(np.int16)
        t += length
    return out


# ------------------------------------------------- device-side generators (torch) ----
# Used where the data set is too large to be born on the host (bench.py, the full-size GPU tests); the oracle
# side simply downloads the tensor, so both sides see the same bytes.
def videolike_features_torch(seed: int, n_scenes: int, frames_per_scene: int, device, d: int = 1024,
                             step: float = 0.12, scene_batch: int = 200):
    """fp32 [n_scenes * frames_per_scene, d] time-ordered video-like rows (scene centre ~ N(0, I), random walk of
    `step` per frame) generated on `device`; the same construction as videolike_features, different stream."""
    feats = torch.empty((n_scenes * frames_per_scene, d), dtype=torch.float32, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    fps = frames_per_scene
    for s0 in range(0, n_scenes, scene_batch):
        m = min(scene_batch, n_scenes - s0)
        v = torch.randn((m, d), generator=g, device=device)
        for f in range(fps):
            feats[(s0 * fps + f)::fps][:m] = v
            v = v + step * torch.randn((m, d), generator=g, device=device)
    return feats


def gaussian_rows_torch(seed: int, n: int, d: int, device, chunk: int = 1 << 18):
    """fp32 [n, d] i.i.d. N(0, 1) rows generated on `device` (SURVEY §8d config 4's second bank, seed 5)."""
    out = torch.empty((n, d), dtype=torch.float32, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    for r0 in range(0, n, chunk):
        m = min(chunk, n - r0)
        out[r0:r0 + m] = torch.randn((m, d), generator=g, device=device)
    return out


def stream_hour_torch(device, nf: int = 3600, h: int = 224, w: int = 224, sr: int = 16000, seed: int = 1):
    """Config 2's synthetic stream, generated on the device: piecewise-static scenes of 5-60 s (smooth field +
    per-frame noise of 2 grey levels), int16 noise at -20 dBFS with silences of 0.6-3 s every 8-40 s.
    Returns (frames uint8 [nf, h, w, 3], pcm int16 [nf * sr, 1], frame_times fp64 [nf])."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    frames = torch.empty((nf, h, w, 3), dtype=torch.uint8, device=device)
    f0 = 0
    while f0 < nf:
        length = int(torch.randint(5, 61, (1,), generator=g, device=device).item())
        m = min(length, nf - f0)
        yy = torch.linspace(0, 6.28, h, device=device)[:, None, None]
        xx = torch.linspace(0, 6.28, w, device=device)[None, :, None]
        ph = torch.rand((1, 1, 3), generator=g, device=device) * 6.28
        fr = torch.rand((2,), generator=g, device=device) * 3 + 0.5
        field = 128 + 40 * torch.sin(fr[0] * xx + ph) + 40 * torch.cos(fr[1] * yy + ph)
        noisy = field[None] + 2.0 * torch.randn((m, h, w, 3), generator=g, device=device)
        frames[f0:f0 + m] = noisy.round().clamp(0, 255).to(torch.uint8)
        f0 += m
    ns = nf * sr
    pcm = (torch.randn((ns, 1), generator=g, device=device) * 3276.8).round().clamp(-32768, 32767).to(torch.int16)
    t_s = 0
    while True:
        t_s += int(torch.randint(8, 41, (1,), generator=g, device=device).item())
        if t_s >= nf - 3:
            break
        ln = int((0.6 + 2.4 * torch.rand((1,), generator=g, device=device).item()) * sr)
        pcm[t_s * sr:t_s * sr + ln] = (pcm[t_s * sr:t_s * sr + ln].float() * 1e-3).round().to(torch.int16)
        t_s += 3
    ft = torch.arange(nf, dtype=torch.float64, device=device)
    return frames, pcm, ft
