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


class ShardedBank:
    """This rank's shard of an n_total-row bank plus the collective search over all shards."""

    def __init__(self, n_total: int, d: int, device=None, group=None, rank: Optional[int] = None,
                 world: Optional[int] = None):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.n_total, self.d = int(n_total), int(d)
        self.lo, self.hi = shard_range(self.n_total, self.rank, self.world)
        self.local = MemoryBank(self.hi - self.lo, d, device=device, row_base=self.lo)

    def fill_local(self, start_local: int, rows) -> None:
        self.local.fill(start_local, rows)

    def search(self, queries, k: int, path: str = "auto"):
        """Identical (idx, score) on every rank: local top-k, one all-gather, replicated merge."""
        if k > _lib.HIPPO_TOPK_MAX:
            raise ValueError(f"sharded search supports k <= {_lib.HIPPO_TOPK_MAX}")
        _, _, keys = self.local.search_keys(queries, k, path)
        if self.world == 1:
            gathered = keys.unsqueeze(0)
        else:
            gathered = gather_keys(keys, self.group)
        idx, score, _ = merge_keys(gathered, k)
        return idx, score
