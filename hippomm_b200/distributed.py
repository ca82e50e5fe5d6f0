"""Row-sharded memory bank over the GPUs of one box.

The reference has no distributed code.  Here the bank's rows are partitioned contiguously over
the ranks of a `torch.distributed` group (one process per GPU, NCCL over NVLink); every rank runs
the identical local search with GLOBAL row numbers (`row_base` = shard offset), the per-rank
(score,row) order keys -- k x 8 bytes per query -- are exchanged with ONE all-gather, and every
rank runs the same k-way merge kernel, so all ranks hold identical results, equal to the
single-GPU answer.  Segmentation and consolidation never communicate (replicas only).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _cuda, _lib
from .bank import MemoryBank


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [lo, hi) of `rank`: contiguous, sizes differ by at most one row."""
    if not (0 <= rank < world):
        raise ValueError("rank outside the group")
    return n_total * rank // world, n_total * (rank + 1) // world


def gather_keys(keys: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather of the local order keys [nq, k] (int64 bit patterns) -> [world, nq, k] on every rank."""
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(keys.shape), dtype=keys.dtype, device=keys.device)
    try:
        dist.all_gather_into_tensor(out, keys.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):
        # backends without the flat variant (some gloo builds, used by the CPU tests)
        dist.all_gather(list(out.unbind(0)), keys.contiguous(), group=group)
    return out


def merge_keys(gathered: torch.Tensor, k: int):
    """[parts, nq, k_in] order keys -> (idx int64 [nq, k], score fp32 [nq, k], key [nq, k]) via hippo_topk_merge."""
    lib = _lib.load()
    dev = _cuda.require_device(gathered.device)
    parts, nq, k_in = gathered.shape
    idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
    score = torch.empty((nq, k), dtype=torch.float32, device=dev)
    key = torch.empty((nq, k), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.hippo_topk_merge(gathered.contiguous().data_ptr(), parts, nq, k_in, k, idx.data_ptr(),
                                        score.data_ptr(), key.data_ptr(), _cuda.stream_ptr()))
    return idx, score, key


class PeerExchange:
    """Symmetric gather buffers for `hippo_topk_exchange_merge`: one allocation per rank, mapped into every
    process of the group (torch symmetric memory: CUDA VMM handles exchanged once at rendezvous), so the
    merge kernel stores its keys straight into the peers' HBM over NVLink.  Sized for (nq_cap, k_cap)."""

    def __init__(self, nq_cap: int, k_cap: int, device, group=None):
        import torch.distributed._symmetric_memory as symm

        lib = _lib.load()
        self.group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.nq_cap, self.k_cap = int(nq_cap), int(k_cap)
        self.nbytes = int(lib.hippo_topk_exchange_bytes(self.world, self.nq_cap, self.k_cap))
        with torch.cuda.device(device):
            self.buf = symm.empty(self.nbytes, dtype=torch.uint8, device=device)
            self.buf.zero_()
            self.handle = symm.rendezvous(self.buf, self.group.group_name)
            torch.cuda.synchronize()
        self.handle.barrier()              # every rank's buffer is cleared before anyone pushes
        self.epoch = 0

    def next_epoch(self) -> int:
        # 1, 2, ..., 2^32 - 1, then 2, 3, ...: never 0 (the cleared state of the flags), and the parity -- which of
        # the two gather buffers a call uses -- keeps alternating across the wrap (2^32 - 1 is odd, 2 is even)
        self.epoch = self.epoch + 1 if self.epoch < 0xFFFFFFFF else 2
        return self.epoch

    def fits(self, nq: int, k: int) -> bool:
        return _lib.load().hippo_topk_exchange_bytes(self.world, nq, k) <= self.nbytes

    def exchange_merge(self, keys: torch.Tensor, k: int):
        """keys: this rank's order keys [nq, k_in] (int64 bit patterns). Returns (idx, score, key) [nq, k]."""
        lib = _lib.load()
        dev = keys.device
        nq, k_in = keys.shape
        idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
        score = torch.empty((nq, k), dtype=torch.float32, device=dev)
        key = torch.empty((nq, k), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.hippo_topk_exchange_merge(
                keys.contiguous().data_ptr(), nq, k_in, k, self.handle.buffer_ptrs_dev, self.nbytes, self.rank,
                self.world, self.next_epoch(), idx.data_ptr(), score.data_ptr(), key.data_ptr(), _cuda.stream_ptr()))
        return idx, score, key


def sharded_search_fused(bank: MemoryBank, queries, k: int, peer: "PeerExchange", path: str = "auto"):
    """This rank's whole share of a sharded search in ONE C-ABI call: `hippo_topk_batched_sharded` (tcgen05 pass, then
    one kernel for local merge + NVLink push + global merge) or, for one query, `hippo_topk_single_sharded` (ONE launch:
    the last CTA of the GEMV merges, pushes, waits and merges).  Returns (idx, score, key) [nq, k], identical on all
    ranks.  Every rank of the group must make the matching call."""
    lib = _lib.load()
    if k > _lib.HIPPO_TOPK_MAX:
        raise ValueError(f"sharded search supports k <= {_lib.HIPPO_TOPK_MAX}")
    q = bank._prep_queries(queries)
    nq = q.shape[0]
    dev = bank.device
    if not peer.fits(nq, k):
        raise ValueError("peer exchange buffers too small for this batch")
    idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
    score = torch.empty((nq, k), dtype=torch.float32, device=dev)
    key = torch.empty((nq, k), dtype=torch.int64, device=dev)
    single = path == "single" or (path == "auto" and nq == 1)
    with torch.cuda.device(dev):
        stream = _cuda.stream_ptr()
        if single:
            ws = _cuda.workspace(lib.hippo_topk_single_workspace_bytes(bank.n, bank.d_pad, k), dev, "topk")
            for qi in range(nq):
                _lib.check(lib.hippo_topk_single_sharded(
                    bank.rows.data_ptr(), bank.norm.data_ptr(), bank.n, bank.d_pad, q[qi].data_ptr(), k, bank.row_base,
                    None, peer.handle.buffer_ptrs_dev, peer.nbytes, peer.rank, peer.world, peer.next_epoch(),
                    idx[qi].data_ptr(), score[qi].data_ptr(), key[qi].data_ptr(), ws.data_ptr(), ws.numel(), stream))
        else:
            ws = _cuda.workspace(lib.hippo_topk_batched_workspace_bytes(bank.n, bank.d_pad, nq, k), dev, "topk")
            _lib.check(lib.hippo_topk_batched_sharded(
                bank.rows.data_ptr(), bank.norm.data_ptr(), bank.n, bank.d_pad, q.data_ptr(), nq, k, bank.row_base, None,
                peer.handle.buffer_ptrs_dev, peer.nbytes, peer.rank, peer.world, peer.next_epoch(),
                idx.data_ptr(), score.data_ptr(), key.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    return idx, score, key


class ShardedBank:
    """This rank's shard of an n_total-row bank plus the collective search over all shards.

    exchange = "p2p"  : the fused push + merge kernel over peer memory (`hippo_topk_exchange_merge`);
               "nccl" : one all-gather of the keys, then `hippo_topk_merge` (also what the CPU/gloo tests drive);
               "auto" : p2p on CUDA when the symmetric buffers can be set up, else nccl.
    """

    def __init__(self, n_total: int, d: int, device=None, group=None, rank: Optional[int] = None,
                 world: Optional[int] = None, exchange: str = "auto"):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.n_total, self.d = int(n_total), int(d)
        self.lo, self.hi = shard_range(self.n_total, self.rank, self.world)
        self.local = MemoryBank(self.hi - self.lo, d, device=device, row_base=self.lo)
        if exchange not in ("auto", "p2p", "nccl"):
            raise ValueError(f"unknown exchange {exchange!r}")
        self.exchange = exchange
        self._peer: Optional[PeerExchange] = None
        self._peer_failed: Optional[str] = None

    def _peer_exchange(self, nq: int, k: int) -> Optional[PeerExchange]:
        if self.exchange == "nccl" or self.world == 1 or self._peer_failed:
            return None
        if self._peer is None or not self._peer.fits(nq, k):
            # collective: every rank reaches this with the same (nq, k).  The outcome is agreed on by ALL ranks: a rank
            # that fell back to NCCL on its own would leave the others spinning on flags that never arrive.
            err = None
            try:
                peer = PeerExchange(max(nq, 4096), max(k, _lib.HIPPO_TOPK_MAX), self.local.device, self.group)
            except Exception as e:  # no peer access / symmetric memory unavailable
                peer, err = None, f"{type(e).__name__}: {e}"
            ok = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device=self.local.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                self._peer = None
                self._peer_failed = err or "peer exchange unavailable on another rank"
                if self.exchange == "p2p":
                    raise RuntimeError(f"exchange='p2p' requested but the peer exchange could not be set up on every "
                                       f"rank ({self._peer_failed})")
                return None
            self._peer = peer
        return self._peer

    def fill_local(self, start_local: int, rows) -> None:
        self.local.fill(start_local, rows)

    def search(self, queries, k: int, path: str = "auto"):
        """Identical (idx, score) on every rank: local top-k, one all-gather, replicated merge."""
        if k > _lib.HIPPO_TOPK_MAX:
            raise ValueError(f"sharded search supports k <= {_lib.HIPPO_TOPK_MAX}")
        if self.world > 1 and self.exchange != "nccl" and not self._peer_failed:
            nq = 1 if getattr(queries, "ndim", None) == 1 or (hasattr(queries, "dim") and queries.dim() == 1) else len(queries)
            peer = self._peer_exchange(nq, k)
            if peer is not None:
                idx, score, _ = sharded_search_fused(self.local, queries, k, peer, path)
                return idx, score
        _, _, keys = self.local.search_keys(queries, k, path)
        if self.world == 1:
            gathered = keys.unsqueeze(0)
        else:
            peer = self._peer_exchange(keys.shape[0], k)
            if peer is not None:
                idx, score, _ = peer.exchange_merge(keys, k)
                return idx, score
            gathered = gather_keys(keys, self.group)
        idx, score, _ = merge_keys(gathered, k)
        return idx, score
