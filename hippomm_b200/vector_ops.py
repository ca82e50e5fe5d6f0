"""Drop-in replacements for hippomm/utils/vector_ops.py: `top_k_cosine_similarity` (vo:151-188)
and `cosine_similarity` (vo:6-20), same argument meaning and return types, computed on the GPU.

The feature array arrives with every call (that is the reference's signature), so the default path is
`hippo_topk_rows`: one streaming pass over the rows in their own precision (fp32, or fp64 for ThetaEvents
reloaded from JSON, hm:391) -- nothing is rounded to bf16, scores agree with NumPy's to fp32 / fp64
rounding noise and so do the indices.  With `set_bank_cache(n)` / `install(cache_banks=True)` arrays that
the caller has marked READ-ONLY (`arr.flags.writeable = False`) are kept on the device between calls
(bf16 bank + the original rows, searched with `exact=True`); a writeable array is never cached, because
nothing cheaper than the upload itself can prove it has not been edited in place.
"""
from __future__ import annotations

import weakref
from collections import OrderedDict
from typing import Tuple, Union

import numpy as np
import torch

from . import _cuda
from .bank import MemoryBank, search_rows

ArrayLike = Union[np.ndarray, torch.Tensor]

# Optional reuse of device banks across calls with the same `b` (the reference is called once per
# ThetaEvent per query with the same feature array, hm:3143-3153).  Off unless install(cache_banks=True).
_bank_cache: "OrderedDict[int, tuple]" = OrderedDict()     # id(array) -> (weakref, signature, MemoryBank)
_bank_cache_size = 0


def set_bank_cache(entries: int) -> None:
    """Keep up to `entries` device banks of READ-ONLY host arrays, keyed by object identity (0 = off)."""
    global _bank_cache_size
    _bank_cache_size = max(0, int(entries))
    while len(_bank_cache) > _bank_cache_size:
        _bank_cache.popitem(last=False)


def invalidate_bank_cache(b=None) -> None:
    """Forget the cached device bank of `b` (or all of them)."""
    if b is None:
        _bank_cache.clear()
    else:
        _bank_cache.pop(id(b), None)


def _signature(b: np.ndarray) -> tuple:
    return (b.__array_interface__["data"][0], b.shape, b.dtype.str, b.strides)


def _cacheable(b) -> bool:
    """Only arrays nobody can edit in place through this object: not writeable, and either owning their
    data or a view of a base that is itself read-only."""
    if not isinstance(b, np.ndarray) or b.flags.writeable or b.dtype not in (np.float32, np.float64):
        return False
    base = b.base
    while base is not None:
        if isinstance(base, np.ndarray):
            if base.flags.writeable:
                return False
            base = base.base
        else:
            return False            # foreign buffer (mmap, bytes, ...): cannot vouch for it
    return True


def _cached_bank(b: np.ndarray):
    if _bank_cache_size <= 0 or not _cacheable(b):
        return None
    ent = _bank_cache.get(id(b))
    if ent is not None and ent[0]() is b and ent[1] == _signature(b):
        _bank_cache.move_to_end(id(b))
        return ent[2]
    bank = MemoryBank.from_rows(b, keep_rows=True)
    key = id(b)
    _bank_cache[key] = (weakref.ref(b, lambda _r, key=key: _bank_cache.pop(key, None)), _signature(b), bank)
    while len(_bank_cache) > _bank_cache_size:
        _bank_cache.popitem(last=False)
    return bank


def _np_dtype(x) -> np.dtype:
    if isinstance(x, torch.Tensor):
        return np.dtype(np.float32) if x.dtype == torch.bfloat16 else np.dtype(str(x.dtype).replace("torch.", ""))
    return np.asarray(x).dtype


def _result_dtype(a, b) -> np.dtype:
    dt = np.result_type(_np_dtype(a), _np_dtype(b))
    return dt if dt.kind == "f" else np.dtype(np.float64)


def _float_tensor(x, dev) -> torch.Tensor:
    """Host array / tensor -> device tensor in fp32 or fp64 (whatever NumPy's arithmetic would have used)."""
    if isinstance(x, torch.Tensor):
        t = x.detach()
        if t.dtype not in (torch.float32, torch.float64):
            t = t.to(torch.float32 if t.dtype in (torch.float16, torch.bfloat16) else torch.float64)
        return t.to(dev, non_blocking=True)
    arr = np.asarray(x)
    if arr.dtype not in (np.float32, np.float64):
        arr = arr.astype(np.float32 if arr.dtype == np.float16 else np.float64)
    return _cuda.to_device(arr, dev)


def top_k_cosine_similarity(a: ArrayLike, b: ArrayLike, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Top-k rows of `b` (N, D) by cosine similarity to the single vector `a` (D,).

    Returns (indices int64, similarities), both of length min(k, N), best first -- the contract of
    vo:151-188.  Ties go to the lower row (the reference leaves them unordered); a zero-norm row of
    `b` scores NaN and is returned first, as `np.argsort` places NaN last (vo:185).
    """
    out_dtype = _result_dtype(a, b)
    if isinstance(a, torch.Tensor):
        a = a.detach().reshape(-1)
    else:
        a = np.asarray(a).reshape(-1)
    if isinstance(b, torch.Tensor):
        b = b.detach()
        if b.dim() == 1:
            b = b.reshape(1, -1)
    else:
        b = np.asarray(b)
        if b.ndim == 1:
            b = b.reshape(1, -1)
    n = int(b.shape[0])
    # `argsort(...)[-k:]`: k > 0 keeps min(k, N); k == 0 keeps everything; k < 0 drops the |k| smallest
    if k > 0:
        count = min(int(k), n)
    elif k == 0:
        count = n
    else:
        count = max(n + int(k), 0)
    if count == 0:
        return np.empty((0,), dtype=np.int64), np.empty((0,), dtype=out_dtype)
    dev = _cuda.require_device(b.device if isinstance(b, torch.Tensor) and b.is_cuda else None)
    bank = _cached_bank(b) if isinstance(b, np.ndarray) else None
    if bank is not None:
        idx, score = bank.search(_float_tensor(a, dev).to(torch.float32), count, exact=True)
        idx_h = idx[0].cpu().numpy()
        score_h = score[0].cpu().numpy().astype(out_dtype, copy=False)
        return idx_h, score_h
    idx, score = search_rows(_float_tensor(b, dev), _float_tensor(a, dev), count)
    return idx.cpu().numpy(), score.cpu().numpy().astype(out_dtype, copy=False)


def cosine_similarity(a: ArrayLike, b: ArrayLike) -> float:
    """Cosine similarity of two vectors (vo:6-20): the N = 1 case of the search, in the inputs' own precision."""
    if isinstance(b, torch.Tensor):
        b1 = b.detach().reshape(1, -1)
    else:
        b1 = np.asarray(b).reshape(1, -1)
    _, score = top_k_cosine_similarity(a, b1, 1)
    return score[0]
