"""Memory consolidation: drop-in for HippocampalMemory._select_key_frames (hm:944-967)."""
from __future__ import annotations

import numpy as np
import torch

from . import _cuda, _lib

# Similarity bands inside which a tensor-core (bf16-input) similarity is not trusted and the pair
# is re-evaluated from the fp32 rows.  Both are rigorous bounds, so a decision outside the band can
# never differ from the fp64 evaluation:
#   * rows that are exactly bf16-representable only suffer fp32 accumulation error,
#     <= d * 2^-24 * sum|a_i b_i| / (|a||b|) <= d * 6e-8  (6.1e-5 at d = 1024);
#   * otherwise each operand carries a relative rounding error <= 2^-8, so
#     |sim_bf16 - sim| <= (2 * 2^-8 + 2^-16) * sum|a_i b_i| / (|a||b|) <= 7.83e-3 (Cauchy-Schwarz).
BAND_EXACT = 1e-4
BAND_INEXACT = 7.9e-3


def _band_exact(d: int) -> float:
    return max(BAND_EXACT, d * 1.2e-7)


class RecheckOverflow(RuntimeError):
    """A band produced more near-threshold pairs than the re-evaluation list holds (stats[1] != 0)."""


def select_key_frames_device(features: torch.Tensor, similarity_threshold: float = 0.9,
                             band_exact: float | None = None, band_inexact: float = BAND_INEXACT,
                             band_rows: int = 0, uncertain_cap: int = 0):
    """Greedy redundancy filter on a device tensor (n, d) fp32, d % 64 == 0.

    Returns (kept int64 [n] device tensor, count int32 [1] device tensor, stats int32 [4] device tensor);
    only the first `count` entries of `kept` are valid.  No host synchronisation.

    stats[1] != 0 means a band held more pairs within the bf16 trust band of gamma than `uncertain_cap`
    (0 = default) has room for; the pairs beyond the capacity kept their tensor-core decision, so the result
    is NOT guaranteed equal to the reference's.  The caller must check it once the stream has run and call
    again with a larger capacity / fewer `band_rows` -- `select_key_frames` and `checked_key_frames` do.
    """
    lib = _lib.load()
    dev = _cuda.require_device(features.device)
    if features.dtype != torch.float32 or features.dim() != 2 or not features.is_contiguous():
        raise ValueError("features must be a contiguous fp32 (n, d) tensor")
    n, d = features.shape
    if d % 64:
        raise ValueError("d must be a multiple of 64 (pad with zero columns)")
    if band_exact is None:
        band_exact = _band_exact(d)
    kept = torch.empty((max(n, 1),), dtype=torch.int64, device=dev)
    count = torch.zeros((1,), dtype=torch.int32, device=dev)
    stats = torch.zeros((4,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws_bytes = lib.hippo_consolidate_ex_workspace_bytes(n, d, int(band_rows), int(uncertain_cap))
        ws = _cuda.workspace(ws_bytes, dev, "consolidate")
        _lib.check(lib.hippo_consolidate_ex(
            features.data_ptr(), n, d, float(np.float32(similarity_threshold)), float(band_exact),
            float(band_inexact), int(band_rows), int(uncertain_cap), kept.data_ptr(), count.data_ptr(),
            stats.data_ptr(), ws.data_ptr(), ws.numel(), _cuda.stream_ptr()))
    return kept, count, stats


def checked_key_frames(features: torch.Tensor, similarity_threshold: float = 0.9, **kw) -> torch.Tensor:
    """`select_key_frames_device` + the overflow check (synchronises): kept row numbers as a device tensor.
    When the near-threshold list overflows, the call is repeated with bands half as long and a list eight times
    as large (at most band x n pairs can be near the threshold); RecheckOverflow if that still does not fit."""
    n = features.shape[0]
    band_rows = int(kw.pop("band_rows", 0))
    cap = int(kw.pop("uncertain_cap", 0))
    for attempt in range(4):
        kept, count, stats = select_key_frames_device(features, similarity_threshold, band_rows=band_rows,
                                                      uncertain_cap=cap, **kw)
        c, overflow = (int(v) for v in torch.stack([count[0], stats[1]]).tolist())
        if not overflow:
            return kept[:c]
        band_rows = max(512, (band_rows or 8192) // 2)
        cap = min(max(cap, 64 * (min(n, band_rows * 2) + 1024) + (1 << 20)) * 8, band_rows * max(n, 1), 0x7FFFFF00)
        _cuda.release_workspaces()          # the retry sizes its own (larger) scratch
    raise RecheckOverflow(
        f"select_key_frames: more than {cap} row pairs of one {band_rows}-row band lie within the bf16 trust band "
        f"of gamma = {similarity_threshold}; refusing to return decisions that were not re-evaluated in fp32")


def select_key_frames(features, times=None, similarity_threshold: float = 0.9) -> np.ndarray:
    """Indices (ascending, int64) of the rows a greedy pass keeps: row 0, then every row whose cosine
    similarity to ALL previously kept rows is below the threshold.  `times` is accepted and unused,
    as in the reference (hm:944-945).  Rows are expected in time order (hm:838)."""
    if isinstance(features, torch.Tensor):
        n = features.shape[0]
        if n <= 2:                                   # hm:946-947
            return np.arange(n)
        f = features.detach().to(torch.float32)
    else:
        features = np.asarray(features)
        n = len(features)
        if n <= 2:
            return np.arange(n)
        f = torch.from_numpy(np.ascontiguousarray(features, dtype=np.float32))
    if f.dim() != 2:
        raise ValueError("features must be 2-D (n, d)")
    dev = _cuda.require_device(None if not f.is_cuda else f.device)
    d = f.shape[1]
    d_pad = (d + 63) // 64 * 64
    if f.is_cuda:
        fd = f
    else:
        fd = _cuda.to_device(f, dev)
    if d_pad != d:
        fp = torch.zeros((n, d_pad), dtype=torch.float32, device=dev)
        fp[:, :d] = fd
        fd = fp
    fd = fd.contiguous()
    return checked_key_frames(fd, similarity_threshold).cpu().numpy()
