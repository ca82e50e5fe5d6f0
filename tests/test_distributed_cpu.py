"""CPU, world_size 2 over gloo: the host side of the sharded search -- shard ranges,
global row numbers, the all-gather layout and the order-key payload.  The per-shard top-k is produced
by the test-side NumPy mirror (tests/hostref.py); on the GPU box the same flow runs with the CUDA
kernels (tests/test_gpu_sharded.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import hostref
from oracle import hippo_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hippomm_b200.distributed import gather_keys, shard_range

        bank, queries = cases.search_config1()
        lo, hi = shard_range(len(bank), rank, world)
        local = hostref.local_topk_keys(queries, bank[lo:hi], lo, k)            # global rows via row_base = lo
        gathered = gather_keys(torch.from_numpy(local.view(np.int64)))          # [world, nq, k] on every rank
        assert gathered.shape == (world, len(queries), k)
        merged = hostref.merge_keys(gathered.numpy().view(np.uint64), k)
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), merged)
    finally:
        dist.destroy_process_group()


def test_sharded_topk_merge_equals_global(tmp_path):
    world, k = 2, 5
    port = _free_port()
    mp.spawn(_worker, args=(world, port, k, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npy")
    r1 = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(r0, r1), "ranks disagree after the replicated merge"
    bank, queries = cases.search_config1()
    g = cases.golden()
    rows = hostref.key_rows(r0)
    for qi, q in enumerate(queries):
        ref_idx, _ = O.top_k_cosine_similarity(q, bank, k)
        assert np.array_equal(rows[qi], ref_idx)
        assert np.array_equal(rows[qi], g["search_c1_idx"][qi])
