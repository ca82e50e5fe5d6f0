"""Device-resident memory bank and the cosine top-k search over it.

The reference has no bank object: `top_k_cosine_similarity(a, b, k)` (vo:151-188) receives the
whole (N, 1024) array on every call and recomputes all N row norms every time (vo:179, 97 % of
its run time).  `MemoryBank` is the one addition to the API surface: bf16 rows + fp32 norms kept
in HBM, built once by `hippo_bank_build`, searched by `hippo_topk_single` (one query, GEMV) or
`hippo_topk_batched` (many queries, tcgen05).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _cuda, _lib

_NP2ENUM = {np.dtype(np.float32): _lib.HIPPO_F32, np.dtype(np.float64): _lib.HIPPO_F64}
_T2ENUM = {torch.float32: _lib.HIPPO_F32, torch.float64: _lib.HIPPO_F64, torch.bfloat16: _lib.HIPPO_BF16}


def _pad64(d: int) -> int:
    return (d + 63) // 64 * 64


class MemoryBank:
    """n rows of dimension d held on one GPU as bf16 [n, d_pad] plus fp32 norms [n]."""

    def __init__(self, n: int, d: int, device=None, row_base: int = 0):
        if n < 0 or d <= 0:
            raise ValueError("MemoryBank needs n >= 0 and d > 0")
        self.device = _cuda.require_device(device)
        self.n, self.d, self.d_pad = int(n), int(d), _pad64(int(d))
        self.row_base = int(row_base)
        if self.row_base + self.n >= 0xFFFFFFFF:
            raise ValueError("global row numbers must stay below 2^32 - 1")
        alloc = torch.zeros if self.d_pad != self.d else torch.empty
        self.rows = alloc((max(self.n, 1), self.d_pad), dtype=torch.bfloat16, device=self.device)
        self.norm = torch.empty((max(self.n, 1),), dtype=torch.float32, device=self.device)
        self._inexact = torch.zeros((1,), dtype=torch.int32, device=self.device)

    # ------------------------------------------------------------------ build ----
    @classmethod
    def from_rows(cls, rows, device=None, row_base: int = 0, chunk_rows: int = 1 << 18) -> "MemoryBank":
        """Build from a host array or a tensor of shape (n, d); float32 / float64 (/ bfloat16 tensors)."""
        if isinstance(rows, torch.Tensor):
            if rows.dim() == 1:
                rows = rows.reshape(1, -1)
            n, d = rows.shape
        else:
            rows = np.asarray(rows)
            if rows.ndim == 1:
                rows = rows.reshape(1, -1)
            if rows.dtype not in _NP2ENUM:
                rows = rows.astype(np.float32 if rows.dtype.itemsize <= 4 else np.float64)
            n, d = rows.shape
        bank = cls(n, d, device=device, row_base=row_base)
        for r0 in range(0, n, chunk_rows):
            r1 = min(n, r0 + chunk_rows)
            bank.fill(r0, rows[r0:r1])
        return bank

    def fill(self, start: int, rows) -> None:
        """(Re)build rows [start, start + len(rows)) from a host array / tensor on any device."""
        lib = _lib.load()
        if isinstance(rows, torch.Tensor):
            t = rows.detach()
            if t.dtype not in _T2ENUM:
                t = t.to(torch.float32)
            t = t.to(self.device, non_blocking=True)
        else:
            t = _cuda.to_device(rows, self.device)
        if t.dim() != 2 or t.shape[1] != self.d:
            raise ValueError(f"expected rows of shape (m, {self.d}), got {tuple(t.shape)}")
        m = t.shape[0]
        if start < 0 or start + m > self.n:
            raise ValueError("row range outside the bank")
        if m == 0:
            return
        if self.d_pad != self.d:  # zero columns change neither dots nor norms
            tp = torch.zeros((m, self.d_pad), dtype=t.dtype, device=self.device)
            tp[:, : self.d] = t
            t = tp
        if t.stride(1) != 1:
            t = t.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(lib.hippo_bank_build(
                t.data_ptr(), _T2ENUM[t.dtype], m, self.d_pad, t.stride(0),
                self.rows[start:].data_ptr(), self.norm[start:].data_ptr(), self._inexact.data_ptr(),
                _cuda.stream_ptr()))
        # keep `t` alive until the kernel has consumed it
        t.record_stream(torch.cuda.current_stream(self.device))

    @property
    def bf16_exact(self) -> bool:
        """True iff no element changed when the rows were rounded to bf16 (synchronises)."""
        return int(self._inexact.item()) == 0

    # ----------------------------------------------------------------- search ----
    def _prep_queries(self, queries) -> torch.Tensor:
        if isinstance(queries, torch.Tensor):
            q = queries.detach().to(self.device, torch.float32, non_blocking=True)
        else:
            q = _cuda.to_device(np.asarray(queries, dtype=np.float32), self.device)
        if q.dim() == 1:
            q = q.reshape(1, -1)
        if q.dim() != 2 or q.shape[1] != self.d:
            raise ValueError(f"expected queries of shape (nq, {self.d}), got {tuple(q.shape)}")
        if self.d_pad != self.d:
            qp = torch.zeros((q.shape[0], self.d_pad), dtype=torch.float32, device=self.device)
            qp[:, : self.d] = q
            q = qp
        return q.contiguous()

    def search_keys(self, queries, k: int, path: str = "auto"):
        """Device-side search. Returns (idx int64 [nq, k], score fp32 [nq, k], key int64-bits [nq, k]).

        idx is -1 / key 0 where fewer than k rows exist.  `path`: "auto" (GEMV for one query,
        tensor cores otherwise), "single" or "batched".
        """
        if k < 1:
            raise ValueError("k must be >= 1")
        lib = _lib.load()
        q = self._prep_queries(queries)
        nq = q.shape[0]
        dev = self.device
        idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
        score = torch.empty((nq, k), dtype=torch.float32, device=dev)
        key = torch.empty((nq, k), dtype=torch.int64, device=dev)
        if nq == 0:
            return idx, score, key
        use_single = path == "single" or (path == "auto" and nq == 1)
        if path not in ("auto", "single", "batched"):
            raise ValueError(f"unknown path {path!r}")
        kmax = _lib.HIPPO_TOPK_MAX
        with torch.cuda.device(dev):
            stream = _cuda.stream_ptr()
            done = 0
            cursor = None  # uint64 bits [nq]: last key handed out so far
            while done < k:
                kk = min(kmax, k - done)
                o_idx = idx if (done == 0 and kk == k) else torch.empty((nq, kk), dtype=torch.int64, device=dev)
                o_sc = score if o_idx is idx else torch.empty((nq, kk), dtype=torch.float32, device=dev)
                o_key = key if o_idx is idx else torch.empty((nq, kk), dtype=torch.int64, device=dev)
                if use_single:
                    ws_bytes = lib.hippo_topk_single_workspace_bytes(self.n, self.d_pad, kk)
                    ws = _cuda.workspace(ws_bytes, dev, "topk")
                    for qi in range(nq):
                        _lib.check(lib.hippo_topk_single(
                            self.rows.data_ptr(), self.norm.data_ptr(), self.n, self.d_pad, q[qi].data_ptr(), kk,
                            self.row_base, None if cursor is None else cursor[qi:].data_ptr(),
                            o_idx[qi].data_ptr(), o_sc[qi].data_ptr(), o_key[qi].data_ptr(),
                            ws.data_ptr(), ws.numel(), stream))
                else:
                    ws_bytes = lib.hippo_topk_batched_workspace_bytes(self.n, self.d_pad, nq, kk)
                    ws = _cuda.workspace(ws_bytes, dev, "topk")
                    _lib.check(lib.hippo_topk_batched(
                        self.rows.data_ptr(), self.norm.data_ptr(), self.n, self.d_pad, q.data_ptr(), nq, kk,
                        self.row_base, None if cursor is None else cursor.data_ptr(),
                        o_idx.data_ptr(), o_sc.data_ptr(), o_key.data_ptr(), ws.data_ptr(), ws.numel(), stream))
                if o_idx is not idx:
                    idx[:, done:done + kk] = o_idx
                    score[:, done:done + kk] = o_sc
                    key[:, done:done + kk] = o_key
                done += kk
                if done < k:
                    cursor = o_key[:, kk - 1].contiguous()  # 0 once a query is exhausted -> nothing qualifies
        return idx, score, key

    def search(self, queries, k: int, path: str = "auto"):
        """(indices int64 [nq, k], scores fp32 [nq, k]) as device tensors; see search_keys."""
        idx, score, _ = self.search_keys(queries, k, path)
        return idx, score
