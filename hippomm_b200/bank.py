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


_copy_streams: dict = {}
_pinned: dict = {}


def _copy_stream(dev: torch.device) -> "torch.cuda.Stream":
    st = _copy_streams.get(dev.index)
    if st is None:
        st = _copy_streams[dev.index] = torch.cuda.Stream(device=dev)
    return st


def _pinned_pair(numel: int, dtype: torch.dtype):
    """Two grow-only pinned staging buffers per dtype (cudaHostAlloc costs ~0.1 s per call: allocate once)."""
    pair = _pinned.get(dtype)
    if pair is None or pair[0].numel() < numel:
        pair = _pinned[dtype] = [torch.empty((numel,), dtype=dtype, pin_memory=True) for _ in range(2)]
    return pair


# Rigorous bounds on |score_bf16 - score| (Cauchy-Schwarz on the element-wise rounding errors): a bf16 operand
# carries a relative error <= 2^-9 per element (round to nearest, 8 significant bits), fp32 accumulation of d
# products adds <= d * 2^-24.  exact search keeps paging until no row outside its candidate list can still reach
# the k-th exact score under these bounds.
_EPS_BF16 = 2.0 ** -9


class StagedQueries:
    """Queries on their way to the device: `tensor` is valid on any stream that has waited for `event`."""

    def __init__(self, tensor: torch.Tensor, event: "torch.cuda.Event"):
        self.tensor, self.event = tensor, event


def search_rows(rows: torch.Tensor, query: torch.Tensor, k: int, row_base: int = 0):
    """Top-k of ONE query straight over device rows (n, d) fp32 / fp64 -- `hippo_topk_rows`, no bank, nothing rounded
    to bf16.  k beyond HIPPO_TOPK_MAX pages with the cursor.  Returns (idx int64 [k], score fp64 [k]) device tensors
    (idx -1 where fewer than k rows exist); scores are evaluated in the arrays' own precision (vo:178-182)."""
    lib = _lib.load()
    dev = _cuda.require_device(rows.device)
    if rows.dim() != 2 or rows.dtype not in _T2ENUM or rows.dtype == torch.bfloat16:
        raise ValueError("rows must be a 2-D float32 / float64 device tensor")
    if rows.stride(1) != 1:
        rows = rows.contiguous()
    n, d = rows.shape
    q = query.reshape(-1).to(dev)
    if q.dtype not in (torch.float32, torch.float64):
        q = q.to(torch.float32)
    q = q.contiguous()
    if q.numel() != d:
        raise ValueError(f"expected a query of dimension {d}, got {q.numel()}")
    idx = torch.empty((k,), dtype=torch.int64, device=dev)
    key = torch.empty((k,), dtype=torch.int64, device=dev)
    kmax = _lib.HIPPO_TOPK_MAX
    with torch.cuda.device(dev):
        stream = _cuda.stream_ptr()
        ws = _cuda.workspace(lib.hippo_topk_rows_workspace_bytes(n, d, min(k, kmax)), dev, "topk_rows")
        done = 0
        cursor = None
        while done < k:
            kk = min(kmax, k - done)
            _lib.check(lib.hippo_topk_rows(
                rows.data_ptr(), _T2ENUM[rows.dtype], n, d, rows.stride(0), q.data_ptr(), _T2ENUM[q.dtype], kk,
                row_base, None if cursor is None else cursor.data_ptr(), idx[done:].data_ptr(), None,
                key[done:].data_ptr(), ws.data_ptr(), ws.numel(), stream))
            done += kk
            if done < k:
                cursor = key[done - 1:done].clone()
        score = torch.empty((k,), dtype=torch.float64, device=dev)
        _lib.check(lib.hippo_rescore(rows.data_ptr(), _T2ENUM[rows.dtype], n, d, rows.stride(0), row_base, q.data_ptr(),
                                     _T2ENUM[q.dtype], d, 1, idx.data_ptr(), k, None, score.data_ptr(), stream))
    return idx, score


class MemoryBank:
    """n rows of dimension d held on one GPU as bf16 [n, d_pad] plus fp32 norms [n]."""

    def __init__(self, n: int, d: int, device=None, row_base: int = 0, keep_rows: bool = False,
                 rows_dtype: torch.dtype = torch.float32):
        """keep_rows=True also keeps the ORIGINAL rows (`rows_dtype`: float32 / float64) in HBM, which is what
        `search(..., exact=True)` re-scores its bf16 candidates from (41 GB at 10M x 1024 fp32)."""
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
        self._bf16_exact = None
        if keep_rows and rows_dtype not in (torch.float32, torch.float64):
            raise ValueError("keep_rows needs float32 or float64 rows")
        self.src = torch.empty((max(self.n, 1), self.d), dtype=rows_dtype, device=self.device) if keep_rows else None
        self.exact_complete = True      # set by search(exact=True): False if the page budget ran out first

    # ------------------------------------------------------------------ build ----
    @classmethod
    def from_rows(cls, rows, device=None, row_base: int = 0, chunk_rows: int = 1 << 16,
                  keep_rows: bool = False) -> "MemoryBank":
        """Build from a host array or a tensor of shape (n, d); float32 / float64 (/ bfloat16 tensors).
        Host arrays larger than one chunk are uploaded through two pinned staging buffers on a copy stream, so
        the host-side staging copy of chunk i + 1, the DMA of chunk i and the build kernel of chunk i - 1 overlap."""
        if isinstance(rows, torch.Tensor):
            if rows.dim() == 1:
                rows = rows.reshape(1, -1)
            n, d = rows.shape
            rdt = rows.dtype if rows.dtype in (torch.float32, torch.float64) else torch.float32
        else:
            rows = np.asarray(rows)
            if rows.ndim == 1:
                rows = rows.reshape(1, -1)
            if rows.dtype not in _NP2ENUM:
                rows = rows.astype(np.float32 if rows.dtype.itemsize <= 4 else np.float64)
            n, d = rows.shape
            rdt = torch.float32 if rows.dtype == np.float32 else torch.float64
        bank = cls(n, d, device=device, row_base=row_base, keep_rows=keep_rows, rows_dtype=rdt)
        host = not isinstance(rows, torch.Tensor) or not rows.is_cuda
        if host and n > chunk_rows:
            bank._fill_pipelined(rows, chunk_rows)
        else:
            for r0 in range(0, n, chunk_rows):
                r1 = min(n, r0 + chunk_rows)
                bank.fill(r0, rows[r0:r1])
        return bank

    def _fill_pipelined(self, rows, chunk_rows: int) -> None:
        """Host rows -> bank through two pinned buffers + two device staging buffers on a side stream."""
        t_all = rows if isinstance(rows, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(rows))
        if t_all.dtype not in (torch.float32, torch.float64):
            t_all = t_all.to(torch.float32)
        n, d = t_all.shape
        dev = self.device
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream()
            copy = _copy_stream(dev)
            pinned_src = t_all.is_pinned()
            pin = [None, None] if pinned_src else _pinned_pair(chunk_rows * d, t_all.dtype)
            stage = [torch.empty((chunk_rows, d), dtype=t_all.dtype, device=dev) for _ in range(2)]
            h2d_done = [torch.cuda.Event(), torch.cuda.Event()]
            built = [torch.cuda.Event(), torch.cuda.Event()]
            copy.wait_stream(main)
            for i, r0 in enumerate(range(0, n, chunk_rows)):
                b = i & 1
                m = min(chunk_rows, n - r0)
                if pinned_src:
                    src = t_all[r0:r0 + m]
                else:
                    if i >= 2:
                        h2d_done[b].synchronize()            # the DMA that last read this pinned buffer
                    src = pin[b][: m * d].view(m, d)
                    src.copy_(t_all[r0:r0 + m])              # host staging copy (overlaps the DMA of chunk i - 1)
                with torch.cuda.stream(copy):
                    if i >= 2:
                        copy.wait_event(built[b])            # the build kernel that last read this device buffer
                    stage[b][:m].copy_(src, non_blocking=True)
                    h2d_done[b].record(copy)
                main.wait_event(h2d_done[b])
                self.fill(r0, stage[b][:m])
                built[b].record(main)
            for st in stage:
                st.record_stream(main)

    def fill_from_parts(self, parts, chunk_rows: int = 1 << 15) -> None:
        """Rows 0 .. n-1 from a LIST of host arrays laid end to end (the per-event feature arrays of a ThetaEvent store),
        through the same two pinned + two device staging buffers as `_fill_pipelined`: runs of consecutive arrays are
        gathered into one pinned chunk, so the many small arrays cost one DMA and one build launch per chunk."""
        d = self.d
        dt = self.src.dtype if self.src is not None else torch.float32
        for p in parts:
            if np.asarray(p).dtype == np.float64:
                dt = torch.float64
        npdt = np.float64 if dt == torch.float64 else np.float32
        if sum(len(p) for p in parts) != self.n:
            raise ValueError("the parts must add up to the bank's rows")
        dev = self.device
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream()
            copy = _copy_stream(dev)
            pin = _pinned_pair(chunk_rows * d, dt)
            stage = [torch.empty((chunk_rows, d), dtype=dt, device=dev) for _ in range(2)]
            h2d_done = [torch.cuda.Event(), torch.cuda.Event()]
            built = [torch.cuda.Event(), torch.cuda.Event()]
            copy.wait_stream(main)
            # split the parts into pieces of at most chunk_rows rows, then group pieces into chunks
            pieces = []
            for p in parts:
                a = np.asarray(p)
                for r0 in range(0, len(a), chunk_rows):
                    pieces.append(a[r0:r0 + chunk_rows])
            i = c = 0
            start = 0
            while i < len(pieces):
                b = c & 1
                if c >= 2:
                    h2d_done[b].synchronize()                # the DMA that last read this pinned buffer
                view = pin[b][: chunk_rows * d].numpy().reshape(chunk_rows, d)
                m = 0
                while i < len(pieces) and m + len(pieces[i]) <= chunk_rows:
                    k = len(pieces[i])
                    if k:
                        np.copyto(view[m:m + k], pieces[i], casting="same_kind" if pieces[i].dtype.kind == "f" else "unsafe")
                    m += k
                    i += 1
                if m == 0:
                    continue
                src = pin[b][: m * d].view(m, d)
                with torch.cuda.stream(copy):
                    if c >= 2:
                        copy.wait_event(built[b])            # the build kernel that last read this device buffer
                    stage[b][:m].copy_(src, non_blocking=True)
                    h2d_done[b].record(copy)
                main.wait_event(h2d_done[b])
                self.fill(start, stage[b][:m])
                built[b].record(main)
                start += m
                c += 1
            for st in stage:
                st.record_stream(main)
            for e in h2d_done:
                e.synchronize()                              # the pinned pair is shared: done before anyone reuses it

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
        self._bf16_exact = None
        if self.src is not None:
            self.src[start:start + m].copy_(t if t.dtype == self.src.dtype else t.to(self.src.dtype))
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
        """True iff no element changed when the rows were rounded to bf16 (synchronises once per build)."""
        if self._bf16_exact is None:
            self._bf16_exact = int(self._inexact.item()) == 0
        return self._bf16_exact

    # ----------------------------------------------------------------- search ----
    def stage_queries(self, queries) -> StagedQueries:
        """Start the host-to-device copy of a query batch on the bank's COPY stream and return at once; hand the
        result to `search` / `search_keys` later.  With two batches in flight the upload of batch i + 1 runs under
        the search of batch i (pinned host memory makes the copy a plain DMA)."""
        t = queries if isinstance(queries, torch.Tensor) else torch.from_numpy(np.asarray(queries, dtype=np.float32))
        with torch.cuda.device(self.device):
            copy = _copy_stream(self.device)
            with torch.cuda.stream(copy):
                d = t.to(self.device, torch.float32, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
        return StagedQueries(d, ev)

    def _prep_queries(self, queries) -> torch.Tensor:
        if isinstance(queries, StagedQueries):
            with torch.cuda.device(self.device):
                cur = torch.cuda.current_stream()
                cur.wait_event(queries.event)
                queries.tensor.record_stream(cur)
            queries = queries.tensor
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

    def search(self, queries, k: int, path: str = "auto", exact: bool = False, max_pages: int = 16):
        """(indices int64 [nq, k], scores fp32 [nq, k]) as device tensors; see search_keys.

        exact=True (needs keep_rows): the bf16 search only nominates candidates, in pages of HIPPO_TOPK_MAX; every
        candidate is re-scored from the ORIGINAL rows (`hippo_rescore`, fp64 accumulation, the reference's operation
        order) and pages are added until the k-th exact score clears `last bf16 score of the page + eps`, eps being a
        rigorous bound on the bf16 error -- then no row outside the candidates can belong to the exact top-k, and
        indices and order equal the reference's wherever its own scores differ by more than fp32 noise.  One host
        synchronisation per page; `self.exact_complete` tells whether the bound was met within max_pages."""
        if not exact:
            idx, score, _ = self.search_keys(queries, k, path)
            return idx, score
        return self._search_exact(queries, k, path, max_pages)

    def _search_exact(self, queries, k: int, path: str, max_pages: int):
        if self.src is None:
            raise ValueError("exact search needs the original rows: build the bank with keep_rows=True")
        lib = _lib.load()
        dev = self.device
        q = self._prep_queries(queries)
        nq = q.shape[0]
        q_src = q[:, : self.d].contiguous() if self.d_pad != self.d else q
        kmax = _lib.HIPPO_TOPK_MAX
        # bf16 queries only on the tensor-core path (one query and, at d = 1024, two ride the fp32-query GEMV)
        batched = not (path == "single" or (path == "auto" and nq == 1)) and not (nq == 2 and self.d_pad == 1024)
        eps = (0.0 if self.bf16_exact else _EPS_BF16) + (_EPS_BF16 if batched else 0.0)
        eps = eps + eps * eps + self.d * 1.2e-7
        pages_keys = []
        cursor = None
        need_k_pages = (k + kmax - 1) // kmax
        self.exact_complete = True
        idx = score = None
        with torch.cuda.device(dev):
            stream = _cuda.stream_ptr()
            for page in range(max(max_pages, need_k_pages)):
                p_idx, p_score, p_key = self._one_page(q, kmax, path, cursor)
                r_key = torch.empty((nq, kmax), dtype=torch.int64, device=dev)
                _lib.check(lib.hippo_rescore(
                    self.src.data_ptr(), _T2ENUM[self.src.dtype], self.n, self.d, self.src.stride(0), self.row_base,
                    q_src.data_ptr(), _lib.HIPPO_F32, q_src.stride(0), nq, p_idx.data_ptr(), kmax, r_key.data_ptr(), None,
                    stream))
                pages_keys.append(r_key)
                allk = torch.stack(pages_keys).contiguous()
                idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
                score = torch.empty((nq, k), dtype=torch.float32, device=dev)
                _lib.check(lib.hippo_topk_merge(allk.data_ptr(), len(pages_keys), nq, kmax, k, idx.data_ptr(),
                                                score.data_ptr(), None, stream))
                if page + 1 < need_k_pages:
                    cursor = p_key[:, kmax - 1].contiguous()
                    continue
                # rows outside the candidates score at most (last bf16 score of this page) + eps
                full = p_idx[:, kmax - 1] >= 0
                kth = score[:, k - 1]
                have_k = idx[:, k - 1] >= 0
                ok = ~full | (have_k & (torch.isnan(kth) | (kth >= p_score[:, kmax - 1] + eps)))
                if bool(ok.all().item()):
                    break
                cursor = p_key[:, kmax - 1].contiguous()
            else:
                self.exact_complete = False
        return idx, score

    def _one_page(self, q: torch.Tensor, kk: int, path: str, cursor):
        """One page of bf16 candidates for prepared queries q [nq, d_pad]: (idx, score, key) [nq, kk]."""
        lib = _lib.load()
        dev = self.device
        nq = q.shape[0]
        o_idx = torch.empty((nq, kk), dtype=torch.int64, device=dev)
        o_sc = torch.empty((nq, kk), dtype=torch.float32, device=dev)
        o_key = torch.empty((nq, kk), dtype=torch.int64, device=dev)
        stream = _cuda.stream_ptr()
        if path == "single" or (path == "auto" and nq == 1):
            ws = _cuda.workspace(lib.hippo_topk_single_workspace_bytes(self.n, self.d_pad, kk), dev, "topk")
            for qi in range(nq):
                _lib.check(lib.hippo_topk_single(
                    self.rows.data_ptr(), self.norm.data_ptr(), self.n, self.d_pad, q[qi].data_ptr(), kk, self.row_base,
                    None if cursor is None else cursor[qi:].data_ptr(), o_idx[qi].data_ptr(), o_sc[qi].data_ptr(),
                    o_key[qi].data_ptr(), ws.data_ptr(), ws.numel(), stream))
        else:
            ws = _cuda.workspace(lib.hippo_topk_batched_workspace_bytes(self.n, self.d_pad, nq, kk), dev, "topk")
            _lib.check(lib.hippo_topk_batched(
                self.rows.data_ptr(), self.norm.data_ptr(), self.n, self.d_pad, q.data_ptr(), nq, kk, self.row_base,
                None if cursor is None else cursor.data_ptr(), o_idx.data_ptr(), o_sc.data_ptr(), o_key.data_ptr(),
                ws.data_ptr(), ws.numel(), stream))
        return o_idx, o_sc, o_key
