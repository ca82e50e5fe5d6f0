"""Test-side NumPy mirror of the (score,row) order key of hippomm_b200/csrc/common.cuh — the payload the
ranks exchange in the sharded search.  Lives under tests/: the product merges keys with hippo_topk_merge."""
from __future__ import annotations

import numpy as np


def score_to_ord(s: np.ndarray) -> np.ndarray:
    b = np.asarray(s, dtype=np.float32).view(np.uint32).astype(np.uint64)
    nan = (b & np.uint64(0x7FFFFFFF)) > np.uint64(0x7F800000)
    neg = (b & np.uint64(0x80000000)) != 0
    o = np.where(neg, (~b) & np.uint64(0xFFFFFFFF), b | np.uint64(0x80000000))
    return np.where(nan, np.uint64(0xFFFFFFFF), o)


def pack_keys(score: np.ndarray, row: np.ndarray) -> np.ndarray:
    """uint64 keys: larger = better (score desc, NaN above all numbers, lower row first)."""
    lo = (~np.asarray(row, dtype=np.uint64)) & np.uint64(0xFFFFFFFF)
    return (score_to_ord(score) << np.uint64(32)) | lo


def key_rows(keys: np.ndarray) -> np.ndarray:
    k = np.asarray(keys).astype(np.uint64)
    return ((~k) & np.uint64(0xFFFFFFFF)).astype(np.int64)


def local_topk_keys(queries: np.ndarray, rows: np.ndarray, row_base: int, k: int) -> np.ndarray:
    """What one shard contributes: the k best keys per query over its rows (0 = empty slot)."""
    out = np.zeros((len(queries), k), dtype=np.uint64)
    for qi, q in enumerate(queries):
        with np.errstate(all="ignore"):
            sims = np.dot(rows, q) / (np.linalg.norm(rows, axis=1) * np.linalg.norm(q))
        keys = np.sort(pack_keys(sims.astype(np.float32), np.arange(len(rows)) + row_base))[::-1][:k]
        out[qi, : len(keys)] = keys
    return out


def merge_keys(gathered: np.ndarray, k: int) -> np.ndarray:
    """[parts, nq, k_in] -> [nq, k] best keys (what hippo_topk_merge computes)."""
    parts, nq, k_in = gathered.shape
    flat = np.transpose(gathered, (1, 0, 2)).reshape(nq, parts * k_in)
    return np.sort(flat, axis=1)[:, ::-1][:, :k]
