"""Drop-in replacements for hippomm/utils/vector_ops.py: `top_k_cosine_similarity` (vo:151-188)
and `cosine_similarity` (vo:6-20), same argument meaning and return types, computed on the GPU.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Tuple, Union

import numpy as np
import torch

from .bank import MemoryBank

ArrayLike = Union[np.ndarray, torch.Tensor]

# Optional reuse of device banks across calls with the same `b` (the reference is called once per
# ThetaEvent per query with the same feature array, hm:3143-3153).  Off unless install(cache_banks=True).
_bank_cache: "OrderedDict[tuple, MemoryBank]" = OrderedDict()
_bank_cache_size = 0


def set_bank_cache(entries: int) -> None:
    """Keep up to `entries` device banks keyed by the identity + fingerprint of the host array (0 = off)."""
    global _bank_cache_size
    _bank_cache_size = max(0, int(entries))
    while len(_bank_cache) > _bank_cache_size:
        _bank_cache.popitem(last=False)


def _fingerprint(b: np.ndarray) -> tuple:
    flat = b.reshape(-1)
    step = max(1, flat.size // 257)
    sample = flat[::step][:257]
    return (b.__array_interface__["data"][0], b.shape, b.dtype.str, b.strides, hash(sample.tobytes()))


def _bank_for(b) -> MemoryBank:
    if _bank_cache_size > 0 and isinstance(b, np.ndarray):
        key = _fingerprint(b)
        bank = _bank_cache.get(key)
        if bank is None:
            bank = MemoryBank.from_rows(b)
            _bank_cache[key] = bank
            while len(_bank_cache) > _bank_cache_size:
                _bank_cache.popitem(last=False)
        else:
            _bank_cache.move_to_end(key)
        return bank
    return MemoryBank.from_rows(b)


def _result_dtype(a, b) -> np.dtype:
    def npdt(x):
        if isinstance(x, torch.Tensor):
            return np.dtype(str(x.dtype).replace("torch.", "")) if x.dtype != torch.bfloat16 else np.dtype(np.float32)
        return np.asarray(x).dtype
    dt = np.result_type(npdt(a), npdt(b))
    return dt if dt.kind == "f" else np.dtype(np.float64)


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
    bank = _bank_for(b)
    idx, score = bank.search(a, count)
    idx_h = idx[0].cpu().numpy()
    score_h = score[0].cpu().numpy().astype(out_dtype, copy=False)
    return idx_h, score_h


def cosine_similarity(a: ArrayLike, b: ArrayLike) -> float:
    """Cosine similarity of two vectors (vo:6-20): the N = 1 case of the search."""
    if isinstance(b, torch.Tensor):
        b1 = b.detach().reshape(1, -1)
    else:
        b1 = np.asarray(b).reshape(1, -1)
    _, score = top_k_cosine_similarity(a, b1, 1)
    return score[0]
