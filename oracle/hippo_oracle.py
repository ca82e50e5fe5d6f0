"""TEST INFRASTRUCTURE — CPU restatement (NumPy) of the reference's hot path.

Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py
may import this module; nothing under hippomm_b200/ does, and it is never the thing shipped.

Each function restates one reference function and cites it (paths relative to the reference
checkout; hm = hippomm/core/hippocampal_memory.py, bp = hippomm/core/batch_process.py,
vo = hippomm/utils/vector_ops.py).

PINNING.  The reference ships no tests, fixtures or golden vectors.  The
restatements are pinned instead against OUTPUTS OF THE REFERENCE ITSELF: tests/golden/
make_golden.py imports the unmodified reference functions (oracle/reference_shim.py) in the
build container, runs them on seeded inputs and commits the results as tests/golden/*.npz;
tests/test_oracle.py checks every function here against those files (and, where the reference
checkout is present, against the live reference functions).
  * search, consolidation, audio level, boundary state machine, MSE fallback: pinned.
  * SSIM: PARITY UNPINNED.  The arithmetic lives in scikit-image, which is neither in the
    reference tree nor installed here and is unpinned in requirements.txt:30.
    `structural_similarity` below restates the published algorithm (Wang et al. 2004 as
    implemented by skimage.metrics.structural_similarity with its defaults); the reference's
    own call sites (hm:990, bp:61) are then run on top of it.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

try:
    from scipy.ndimage import uniform_filter as _uniform_filter
except Exception:  # pragma: no cover
    _uniform_filter = None


# =====================================================================  feature search  ==
def top_k_cosine_similarity(a: np.ndarray, b: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """vo:151-188, line for line in NumPy (torch inputs are the caller's business)."""
    a = np.asarray(a).reshape(-1)                                    # vo:172
    b = np.asarray(b)
    if len(b.shape) == 1:                                            # vo:173-174
        b = b.reshape(1, -1)
    a_norm = np.linalg.norm(a)                                       # vo:178
    b_norms = np.linalg.norm(b, axis=1)                              # vo:179
    with np.errstate(divide="ignore", invalid="ignore"):
        similarities = np.dot(b, a) / (b_norms * a_norm)             # vo:182
    top_k_indices = np.argsort(similarities)[-k:][::-1]              # vo:185
    return top_k_indices, similarities[top_k_indices]


def cosine_similarity(a: np.ndarray, b: np.ndarray) -> float:
    """vo:6-20."""
    a = np.asarray(a).reshape(-1)
    b = np.asarray(b).reshape(-1)
    return np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b))


def canonical_topk(similarities: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Top-k of a similarity vector with the GPU's documented tie rule (score descending, NaN first,
    lower row first among equals) -- the deterministic representative of vo:185's unordered ties."""
    s = np.asarray(similarities)
    n = s.shape[0]
    key = np.where(np.isnan(s), np.inf, s.astype(np.float64))
    nanflag = np.isnan(s)
    order = np.lexsort((np.arange(n), -key, ~nanflag))               # NaN first, then score desc, then row asc
    idx = order[:k]
    return idx.astype(np.int64), s[idx]


def top_k_streaming(queries: np.ndarray, row_source, n_rows: int, k: int, chunk_rows: int = 1 << 18,
                    dtype=np.float32) -> Tuple[np.ndarray, np.ndarray]:
    """Streaming restatement of vo:151-188 for banks too large for one array (the reference's
    per-call norm temporary doubles a 41 GB bank at 10M rows): per chunk `norm -> dot/(|b||a|)`
    exactly as vo:178-182, then a running top-k with the canonical tie rule.
    `row_source(r0, r1)` returns rows [r0, r1) as a (r1-r0, d) array.  Returns (idx [nq,k], sims [nq,k])."""
    queries = np.asarray(queries, dtype=dtype)
    if queries.ndim == 1:
        queries = queries[None]
    nq = queries.shape[0]
    a_norms = np.array([np.linalg.norm(q) for q in queries], dtype=dtype)
    best_idx = np.full((nq, 0), -1, dtype=np.int64)
    best_sim = np.zeros((nq, 0), dtype=dtype)
    for r0 in range(0, n_rows, chunk_rows):
        r1 = min(n_rows, r0 + chunk_rows)
        b = np.asarray(row_source(r0, r1), dtype=dtype)
        b_norms = np.linalg.norm(b, axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            sims = np.dot(b, queries.T).T / (b_norms[None, :] * a_norms[:, None])
        cat_sim = np.concatenate([best_sim, sims], axis=1)
        cat_idx = np.concatenate([best_idx, np.broadcast_to(np.arange(r0, r1), (nq, r1 - r0))], axis=1)
        new_idx = np.empty((nq, min(k, cat_sim.shape[1])), dtype=np.int64)
        new_sim = np.empty((nq, new_idx.shape[1]), dtype=dtype)
        for qi in range(nq):
            s = cat_sim[qi]
            key = np.where(np.isnan(s), np.inf, s.astype(np.float64))
            kk = new_idx.shape[1]
            if s.shape[0] > 4 * kk:
                cand = np.argpartition(-key, kk - 1)[:kk]
                thr = key[cand].min()
                cand = np.nonzero(key >= thr)[0]
            else:
                cand = np.arange(s.shape[0])
            order = cand[np.lexsort((cat_idx[qi, cand], -key[cand], ~np.isnan(s[cand])))][:kk]
            new_idx[qi] = cat_idx[qi, order]
            new_sim[qi] = s[order]
        best_idx, best_sim = new_idx, new_sim
    return best_idx, best_sim


# ===================================================================  consolidation  ==
def select_key_frames(features: np.ndarray, times=None, similarity_threshold: float = 0.9) -> np.ndarray:
    """hm:944-967, line for line."""
    if len(features) <= 2:                                           # hm:946-947
        return np.arange(len(features))
    with np.errstate(divide="ignore", invalid="ignore"):
        features_normalized = features / np.linalg.norm(features, axis=1, keepdims=True)   # hm:951
    similarity_matrix = np.dot(features_normalized, features_normalized.T)                # hm:952
    key_indices = [0]                                                # hm:955
    for i in range(1, len(features)):                                # hm:958-961
        similarities = similarity_matrix[i, key_indices]
        if np.all(similarities < similarity_threshold):
            key_indices.append(i)
    if len(features) > 1 and np.all(similarity_matrix[-1, key_indices] < similarity_threshold):  # hm:964-965
        key_indices.append(len(features) - 1)
    return np.array(key_indices)


def select_key_frames_blocked(features: np.ndarray, similarity_threshold: float = 0.9, block: int = 4096,
                              with_moat: bool = False):
    """Blocked restatement of hm:944-967 for N where the N x N matrix does not fit (40 GB at N=100k):
    the same fp32 normalisation, fp32 sgemm per row block against the rows kept so far, the same greedy
    rule.  Validated bit-for-bit against `select_key_frames` at N <= 20k in tests/test_oracle.py.
    with_moat=True also returns min |sim - threshold| over the DECISIVE comparisons (fp64 recomputation),
    i.e. how far the closest call was from flipping."""
    n = len(features)
    if n <= 2:
        return (np.arange(n), np.inf) if with_moat else np.arange(n)
    features = np.asarray(features)
    with np.errstate(divide="ignore", invalid="ignore"):
        fn = features / np.linalg.norm(features, axis=1, keepdims=True)
    thr = np.asarray(similarity_threshold, dtype=fn.dtype) if fn.dtype == np.float32 else similarity_threshold
    kept: List[int] = [0]
    kept_rows = np.empty((n, fn.shape[1]), dtype=fn.dtype)
    kept_rows[0] = fn[0]
    nk = 1
    moat = np.inf
    for b0 in range(1, n, block):
        b1 = min(n, b0 + block)
        blk = fn[b0:b1]
        # similarities of the block rows to everything kept BEFORE the block, one sgemm
        s_prev = np.dot(blk, kept_rows[:nk].T)
        s_in = np.dot(blk, blk.T)
        sup_prev = ~np.all(s_prev < thr, axis=1)
        in_kept: List[int] = []
        for r in range(b1 - b0):
            ok = not sup_prev[r]
            if ok and in_kept:
                ok = bool(np.all(s_in[r, in_kept] < thr))
            if with_moat:
                vals = np.concatenate([s_prev[r], s_in[r, in_kept]]) if in_kept else s_prev[r]
                if vals.size:
                    v64 = vals.astype(np.float64)
                    if ok:
                        moat = min(moat, float(np.min(np.abs(v64 - float(thr)))))
                    else:
                        big = v64[~(v64 < float(thr))]
                        if big.size:
                            moat = min(moat, float(np.max(big) - float(thr)))
            if ok:
                in_kept.append(r)
        for r in in_kept:
            kept.append(b0 + r)
            kept_rows[nk] = blk[r]
            nk += 1
    out = np.array(kept)
    return (out, moat) if with_moat else out


def greedy_valid_under_tolerance(features: np.ndarray, kept: np.ndarray, similarity_threshold: float,
                                 tol: float = 1e-3, block: int = 2048) -> Tuple[bool, str]:
    """SURVEY.md §8d parity rule when the oracle's moat is below tolerance: `kept` must be a valid greedy
    solution under tolerance -- every kept i has sim64(i, j) < thr + tol for all earlier kept j; every
    dropped i has an earlier kept j with sim64(i, j) >= thr - tol."""
    f = np.asarray(features, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        fn = f / np.linalg.norm(f, axis=1, keepdims=True)
    n = len(f)
    kept = np.asarray(kept, dtype=np.int64)
    if n <= 2:
        return (np.array_equal(kept, np.arange(n)), "n<=2")
    if kept.size == 0 or kept[0] != 0 or np.any(np.diff(kept) <= 0):
        return False, "kept must start at 0 and ascend"
    is_kept = np.zeros(n, dtype=bool)
    is_kept[kept] = True
    kf = fn[kept]
    for b0 in range(0, n, block):
        b1 = min(n, b0 + block)
        s = fn[b0:b1] @ kf.T
        for r in range(b0, b1):
            nprev = int(np.searchsorted(kept, r, side="left"))
            row = s[r - b0, :nprev]
            if r == 0:
                continue
            if is_kept[r]:
                if nprev and not np.all(row < similarity_threshold + tol):
                    return False, f"kept row {r} is within threshold of an earlier kept row"
            else:
                if not (nprev and np.any(~(row < similarity_threshold - tol))):
                    return False, f"dropped row {r} has no earlier kept row above threshold"
    return True, "ok"


# ============================================================  temporal pattern separation  ==
def bgr2gray(frame: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY) for uint8 (hm:986-987, bp:45-52): OpenCV's fixed-point
    Y = (3735 B + 19235 G + 9798 R + 16384) >> 15 (verified exhaustively against cv2 4.13 over all 2^24
    colours in tests/test_oracle.py when cv2 is importable)."""
    f = frame.astype(np.uint32)
    return ((3735 * f[..., 0] + 19235 * f[..., 1] + 9798 * f[..., 2] + 16384) >> 15).astype(np.uint8)


def structural_similarity(im1: np.ndarray, im2: np.ndarray, *, data_range=None, win_size: int = 7,
                          K1: float = 0.01, K2: float = 0.03, **_ignored) -> float:
    """Restatement of skimage.metrics.structural_similarity with its defaults (PARITY UNPINNED, see
    module docstring): uniform 7x7 filter, sample covariance, border of (win_size-1)//2 cropped,
    float64 throughout.  Call sites: hm:990 (data_range = gray1.max() - gray1.min() in uint8), bp:61
    (float images in [0,1], data_range = 1.0)."""
    if im1.shape != im2.shape:
        raise ValueError("Input images must have the same dimensions.")
    if min(im1.shape) < win_size:
        raise ValueError("win_size exceeds image extent.")
    if data_range is None:
        raise ValueError("data_range must be given for this restatement")
    im1 = im1.astype(np.float64, copy=False)
    im2 = im2.astype(np.float64, copy=False)
    NP = win_size ** im1.ndim
    cov_norm = NP / (NP - 1)
    ux = _uniform_filter(im1, size=win_size)
    uy = _uniform_filter(im2, size=win_size)
    uxx = _uniform_filter(im1 * im1, size=win_size)
    uyy = _uniform_filter(im2 * im2, size=win_size)
    uxy = _uniform_filter(im1 * im2, size=win_size)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    R = float(data_range)
    C1 = (K1 * R) ** 2
    C2 = (K2 * R) ** 2
    A1, A2, B1, B2 = (2 * ux * uy + C1, 2 * vxy + C2, ux ** 2 + uy ** 2 + C1, vx + vy + C2)
    with np.errstate(divide="ignore", invalid="ignore"):
        S = (A1 * A2) / (B1 * B2)
    pad = (win_size - 1) // 2
    return float(S[pad:-pad, pad:-pad].mean(dtype=np.float64))


def structural_similarity_exact(g1: np.ndarray, g2: np.ndarray, data_range: float) -> float:
    """The same quantity from exact integer window moments (uint8 inputs): an independent route to the
    SSIM value, free of the uniform filter's running-sum rounding; used to bound the restatement's own
    float64 noise in tests."""
    x = g1.astype(np.int64)
    y = g2.astype(np.int64)

    def box(a):
        c = np.cumsum(np.cumsum(np.pad(a, ((1, 0), (1, 0))), axis=0), axis=1)
        return c[7:, 7:] - c[:-7, 7:] - c[7:, :-7] + c[:-7, :-7]

    sx, sy, sxx, syy, sxy = box(x), box(y), box(x * x), box(y * y), box(x * y)
    R = float(data_range)
    C1 = (0.01 * R) ** 2
    C2 = (0.03 * R) ** 2
    a1 = 2.0 * (sx * sy) / 2401.0 + C1
    b1 = (sx * sx + sy * sy) / 2401.0 + C1
    a2 = 2.0 * (49 * sxy - sx * sy) / 2352.0 + C2
    b2 = ((49 * sxx - sx * sx) + (49 * syy - sy * sy)) / 2352.0 + C2
    with np.errstate(divide="ignore", invalid="ignore"):
        S = (a1 * a2) / (b1 * b2)
    return float(S.mean(dtype=np.float64))


def compute_frame_similarity(frame1_bgr: np.ndarray, frame2_bgr: np.ndarray) -> float:
    """hm:980-991 after the two cv2.imread calls (decoding is the caller's I/O)."""
    gray1 = bgr2gray(frame1_bgr) if frame1_bgr.ndim == 3 else frame1_bgr
    gray2 = bgr2gray(frame2_bgr) if frame2_bgr.ndim == 3 else frame2_bgr
    return structural_similarity(gray1, gray2, data_range=gray1.max() - gray1.min())     # hm:990, uint8 range


def compute_frame_difference(frame1: np.ndarray, frame2: np.ndarray) -> float:
    """bp:32-71."""
    g1 = bgr2gray(frame1) if len(frame1.shape) == 3 else frame1                          # bp:44-52
    g2 = bgr2gray(frame2) if len(frame2.shape) == 3 else frame2
    f1 = g1.astype(float) / 255.0                                                        # bp:55-56
    f2 = g2.astype(float) / 255.0
    try:
        score = structural_similarity(f1, f2, data_range=1.0)                           # bp:61
        if np.isfinite(score):                                                           # bp:62-63
            return 1.0 - score
    except Exception:
        pass
    mse = np.mean((f1 - f2) ** 2)                                                        # bp:67
    return min(1.0, mse)                                                                 # bp:71


def compute_audio_level(audio_data: np.ndarray, sample_rate=None) -> float:
    """hm:993-1000."""
    if len(audio_data.shape) > 1:
        audio_data = audio_data.mean(axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        rms = np.sqrt(np.mean(np.square(audio_data))) if audio_data.size else np.float64(np.nan)
    db = 20 * np.log10(rms) if rms > 0 else -100
    return db


def segment_boundaries(frame_ssim: Optional[np.ndarray], frame_times: Optional[List[float]],
                       audio_data: Optional[np.ndarray], audio_sample_rate, *, max_segment_duration=30.0,
                       min_segment_duration=10.0, frame_similarity_threshold=0.95,
                       audio_silence_threshold=-40) -> List[Tuple[float, float]]:
    """The boundary state machine of hm:1002-1114 (everything except attaching frames / audio to the
    segments).  `frame_ssim[p]` stands for `_compute_frame_similarity(frame p+1, frame p)`, the only
    pairs hm:1052-1056 ever asks for when frame_times is sorted."""
    bounds: List[Tuple[float, float]] = []
    has_video = frame_times is not None and len(frame_times) > 0
    if not has_video and audio_data is None:
        return bounds
    if has_video:                                                                        # hm:1027-1032
        total_duration = frame_times[-1] - frame_times[0]
    elif audio_data is not None and audio_sample_rate:
        total_duration = len(audio_data) / audio_sample_rate
    else:
        return bounds
    current_start = 0.0                                                                  # hm:1034
    while current_start < total_duration:                                                # hm:1036
        current_end = min(current_start + max_segment_duration, total_duration)          # hm:1038
        optimal_end = current_end
        if has_video:                                                                    # hm:1043-1059
            frame_indices = [i for i, t in enumerate(frame_times) if current_start <= t <= current_end]
            if len(frame_indices) > 1:
                for i in range(len(frame_indices) - 1, 0, -1):
                    hi, lo = frame_indices[i], frame_indices[i - 1]
                    assert hi - lo == 1, "frame_times must be sorted"
                    if frame_ssim[lo] < frame_similarity_threshold:
                        optimal_end = frame_times[hi]
                        break
        if audio_data is not None and audio_sample_rate:                                 # hm:1061-1077
            start_sample = int(current_start * audio_sample_rate)
            end_sample = int(current_end * audio_sample_rate)
            window_size = int(0.5 * audio_sample_rate)
            for i in range(end_sample - start_sample - window_size, 0, -window_size):
                window_start = start_sample + i
                window_end = window_start + window_size
                level = compute_audio_level(audio_data[window_start:window_end], audio_sample_rate)
                if level < audio_silence_threshold:
                    optimal_end = (window_start / audio_sample_rate)
                    break
        if optimal_end - current_start < min_segment_duration:                           # hm:1080-1084
            optimal_end = min(current_start + min_segment_duration, total_duration)
        bounds.append((current_start, optimal_end))
        current_start = optimal_end                                                      # hm:1111
    return bounds


def adjacent_ssim(frames_bgr: np.ndarray) -> np.ndarray:
    """frame_ssim[p] = compute_frame_similarity(frame p+1, frame p) for a stack [n, h, w, 3]."""
    grays = [bgr2gray(f) if f.ndim == 3 else f for f in frames_bgr]
    out = np.empty(max(len(grays) - 1, 0), dtype=np.float64)
    for p in range(len(grays) - 1):
        g1, g2 = grays[p + 1], grays[p]
        out[p] = structural_similarity(g1, g2, data_range=g1.max() - g1.min())
    return out


# =====================================================================  detailed recall  ==
def find_relevant_segments(query: np.ndarray, events, modality: str = "vision", top: int = 5, k: int = 5,
                           pad: float = 1.0):
    """Similarity path of `_find_relevant_video_segments` (hm:3129-3279, modality "vision") and
    `_find_relevant_audio_segments` (hm:3281-3383, modality "audio"): per event top-k (hm:3153 /
    hm:3304), windows around the hits whose index is inside the event's time table (hm:3258-3272 /
    hm:3366-3377), stable sort by similarity, best `top` (hm:3274-3277 / hm:3379-3381).  The LLM
    branch (taken when the best similarity is below 0.4 AND the event has captions / a transcription)
    is outside the arithmetic path; events here carry neither.

    `events`: objects with .features[modality] (n_e, d), .frame_times / .frames (vision) or
    .feature_times['audio_times'] (audio).  Returns a list of dicts
    {start, end, similarity, event, index, frames, frame_times}."""
    q = np.asarray(query).reshape(-1)                                        # hm:3132-3136
    similarity_segments = []
    for ei, event in enumerate(events):
        if modality not in event.features:                                   # hm:3144 / hm:3296
            continue
        feats = np.asarray(event.features[modality])
        idxs, sims = top_k_cosine_similarity(q, feats, k)
        if modality == "vision":
            times = list(event.frame_times)
        else:
            times = list(event.feature_times["audio_times"])
        for idx, sim in zip(idxs, sims):
            if idx < len(times):                                             # hm:3262 / hm:3367
                t = times[idx]
                seg = dict(start=max(0, t - pad), end=t + pad, similarity=sim, event=ei, index=int(idx))
                if modality == "vision":                                     # hm:3267-3270
                    seg["frames"] = [event.frames[i] for i in range(len(event.frames))
                                     if times[i] >= t - pad and times[i] <= t + pad]
                    seg["frame_times"] = [x for x in times if x >= t - pad and x <= t + pad]
                similarity_segments.append((sim, seg))
    similarity_segments.sort(key=lambda x: x[0], reverse=True)               # hm:3274 / hm:3379
    return [seg for _, seg in similarity_segments[:top]]


# =====================================================================  greedy frame filters  ==
def select_saved_frames(frames: np.ndarray, video_fps: float, max_diff_threshold: float = 0.3,
                        check_interval: int = 30):
    """The save decisions of extract_frames_from_video (bp:179-228) on already decoded frames
    (cv2.VideoCapture / imwrite are the caller's I/O).  Returns (frame numbers, times)."""
    frame_numbers, frame_times = [], []
    last_saved_frame = None
    cumulative_diff = 0.0
    last_save_time = 0.0
    for frame_count, frame in enumerate(frames):                               # bp:179-182
        current_time = frame_count / video_fps                                 # bp:184
        save = False
        if last_saved_frame is None:                                           # bp:187-188
            save = True
        elif current_time - last_save_time >= 1.0:                             # bp:190
            if frame_count % check_interval == 0:                              # bp:192
                diff = compute_frame_difference(frame, last_saved_frame)       # bp:194
                cumulative_diff += diff                                        # bp:195
                if diff > max_diff_threshold or cumulative_diff > max_diff_threshold:   # bp:198-200
                    save = True
        if save:                                                               # bp:202-219
            frame_numbers.append(frame_count)
            frame_times.append(current_time)
            last_saved_frame = frame
            cumulative_diff = 0.0
            last_save_time = current_time
    return frame_numbers, frame_times


def dedup_window_frames(frames: np.ndarray, threshold: float = 0.3) -> List[int]:
    """The de-duplication inside one QA re-decode window (hm:2226-2249 with 0.3; hm:2789-2812 with 0.4):
    a frame is dropped when `_compute_frame_similarity(prev_kept, frame) > threshold` (hm:2237-2238)."""
    kept: List[int] = []
    prev = None
    for i, frame in enumerate(frames):
        if prev is not None:
            with np.errstate(all="ignore"):
                similarity = compute_frame_similarity(frames[prev], frame)
            if similarity > threshold:
                continue
        kept.append(i)
        prev = i
    return kept
