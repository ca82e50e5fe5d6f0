"""GPU parity: detailed-recall feature search (vo:151-188) through the C ABI."""
import numpy as np
import pytest
import torch

import cases
from oracle import hippo_oracle as O
from parity import check_topk, check_topk_exact

pytestmark = pytest.mark.gpu


def _search(bank, queries, k, path):
    idx, score = bank.search(queries, k, path=path)
    torch.cuda.synchronize()
    return idx.cpu().numpy(), score.cpu().numpy()


@pytest.mark.parametrize("path", ["single", "batched"])
def test_config1_matches_reference_outputs(cuda_device, path):
    """Config 1 (2,000 rows, 64 queries, k=5) against the unmodified reference's committed outputs."""
    from hippomm_b200 import MemoryBank

    bank_h, queries = cases.search_config1()
    g = cases.golden()
    bank = MemoryBank.from_rows(bank_h)
    idx, score = _search(bank, queries, 5, path)
    for qi in range(len(queries)):
        check_topk(idx[qi], score[qi], g["search_c1_idx"][qi], g["search_c1_sim"][qi], what=f"{path} q{qi}")


@pytest.mark.parametrize("path", ["single", "batched"])
def test_lattice_is_bit_exact(cuda_device, path):
    """On the bf16-exact lattice bank every fp32 partial sum is exact, so scores must equal the
    reference's to the last bit and rows must match outright."""
    from hippomm_b200 import MemoryBank

    bank_h, queries, fam = cases.search_lattice_small()
    g = cases.golden()
    bank = MemoryBank.from_rows(bank_h)
    assert bank.bf16_exact
    idx, score = _search(bank, queries, 10, path)
    for qi in range(len(queries)):
        check_topk_exact(idx[qi], score[qi], g["search_lat_idx"][qi], g["search_lat_sim"][qi], what=f"{path} q{qi}")


def test_wrapper_edge_cases(cuda_device):
    """NaN rows first, ties, k > N, 1-D b, cosine_similarity: same contract as vo:151-188 / vo:6-20."""
    from hippomm_b200 import cosine_similarity, top_k_cosine_similarity

    b, q = cases.search_edge()
    g = cases.golden()
    i1, s1 = top_k_cosine_similarity(q, b, 4)
    assert i1.dtype == np.int64 and s1.dtype == np.float32
    assert i1[0] == 3 and np.isnan(s1[0])                     # zero-norm row: NaN sorts first
    check_topk(i1, s1, g["search_edge_idx"], g["search_edge_sim"], what="edge")
    assert i1[1] == 5 and abs(s1[1] - 1.0) < 1e-6
    i2, s2 = top_k_cosine_similarity(q, b[:3], 8)             # k > N -> N results
    check_topk(i2, s2, g["search_kgtn_idx"], g["search_kgtn_sim"], what="k>N")
    i3, s3 = top_k_cosine_similarity(torch.from_numpy(q), torch.from_numpy(b[5]), 3)   # torch in, 1-D b
    check_topk(i3, s3, g["search_1d_idx"], g["search_1d_sim"], what="1-D b")
    assert abs(float(cosine_similarity(q, b[5])) - float(g["cosine_pair"][0])) < 1e-6
    # ties: rows 2 and 9 are identical -> lower row first
    i4, s4 = top_k_cosine_similarity(q, b, 16)
    p2, p9 = list(i4).index(2), list(i4).index(9)
    assert p9 == p2 + 1 and s4[p2] == s4[p9]
    assert i4[-1] == 11 and abs(s4[-1] + 1.0) < 1e-6
    # k == 0 keeps everything (argsort[-0:]), negative k drops the |k| smallest
    i5, _ = top_k_cosine_similarity(q, b, 0)
    assert len(i5) == 16
    i6, _ = top_k_cosine_similarity(q, b, -14)
    assert list(i6) == list(i4[:2])
    # float64 bank (reloaded ThetaEvents, hm:391) keeps the reference's result dtype
    i7, s7 = top_k_cosine_similarity(q, b.astype(np.float64), 4)
    assert s7.dtype == np.float64 and list(i7) == list(i1)


@pytest.mark.parametrize("path", ["single", "batched"])
def test_paging_beyond_kmax(cuda_device, path):
    """k above HIPPO_TOPK_MAX pages with the (score,row) cursor; must equal the full sort."""
    from hippomm_b200 import MemoryBank

    rng = np.random.default_rng(5)
    b = rng.standard_normal((700, 128)).astype(np.float32)
    q = rng.standard_normal((3, 128)).astype(np.float32)
    bank = MemoryBank.from_rows(b)
    idx, score = _search(bank, q, 100, path)
    for qi in range(3):
        ri, rs = O.top_k_cosine_similarity(q[qi], b, 100)
        check_topk(idx[qi], score[qi], ri, rs, what=f"{path} paging q{qi}")
    idx_all, score_all = _search(bank, q[:1], 700, path)
    assert sorted(idx_all[0].tolist()) == list(range(700))
    assert np.all(np.diff(score_all[0]) <= 0)


@pytest.mark.parametrize("n,d,nq,k", [(1, 64, 1, 1), (5, 64, 2, 8), (255, 192, 130, 7), (257, 1024, 129, 32),
                                      (1000, 320, 5, 10), (4099, 1024, 300, 10)])
def test_ragged_shapes_both_paths_agree_with_oracle(cuda_device, n, d, nq, k):
    """Row counts off the 256-row tile, query counts off the 128-row tile, d off 256: both kernels vs oracle."""
    from hippomm_b200 import MemoryBank

    rng = np.random.default_rng(n * 7 + d)
    b = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    bank = MemoryBank.from_rows(b)
    kk = min(k, n)
    for path in ("single", "batched"):
        idx, score = _search(bank, q, k, path)
        for qi in range(0, nq, max(1, nq // 16)):
            ri, rs = O.top_k_cosine_similarity(q[qi], b, k)
            check_topk(idx[qi][:kk], score[qi][:kk], ri, rs, what=f"{path} n={n} d={d} q{qi}")
            assert np.all(idx[qi][kk:] == -1)


@pytest.mark.parametrize("nq", [2, 3, 4, 5, 7, 8, 16, 31, 32, 33, 64, 65, 100, 128, 129])
def test_few_queries_ride_one_bank_pass(cuda_device, nq):
    """A few queries of dimension 1024 behind the batched entry (two take the GEMV with two accumulators per row,
    topk_single.cu; 3 .. 128 the small-batch tcgen05 kernel with the bank rows on the M side, topk_small.cu; more the
    256-query tiles of sim_tc.cu): bit-equal to the single-query kernel on the lattice bank, within the parity rule
    on Gaussian rows, NaN rows first, paging beyond HIPPO_TOPK_MAX, row counts off the 4-row groups / 128-row tiles."""
    from hippomm_b200 import MemoryBank, synth

    bank_h, queries, fam = cases.search_lattice_small()
    if nq > len(queries):
        queries, fam = synth.lattice_queries_np(4, nq, bank_h.shape[1], len(bank_h))
    bank = MemoryBank.from_rows(bank_h)
    bi, bs = _search(bank, queries[:nq], 10, "batched")
    for qi in range(nq):
        si, ss = _search(bank, queries[qi:qi + 1], 10, "single")
        assert np.array_equal(bi[qi], si[0]) and np.array_equal(bs[qi].view(np.uint32), ss[0].view(np.uint32))
    g = cases.golden()
    for qi in range(min(nq, len(g["search_lat_idx"]))):
        check_topk_exact(bi[qi], bs[qi], g["search_lat_idx"][qi], g["search_lat_sim"][qi], what=f"few q{qi}")
    rng = np.random.default_rng(100 + nq)
    for n in (1, 3, 1021, 40000):
        b = rng.standard_normal((n, 1024)).astype(np.float32)
        if n > 2:
            b[2] = 0.0                                        # zero-norm row: NaN, sorts first (vo:185)
        q = rng.standard_normal((nq, 1024)).astype(np.float32)
        bk = MemoryBank.from_rows(b)
        k = 40 if n > 40 else 7                               # 40 pages through the per-query cursor
        fi, fs = _search(bk, q, k, "batched")
        for qi in range(nq):
            si, ss = _search(bk, q[qi:qi + 1], k, "single")
            check_topk(fi[qi][:min(k, n)], fs[qi][:min(k, n)], si[0][:min(k, n)], ss[0][:min(k, n)], what=f"few vs single n={n} q{qi}")
            ri, rs = O.top_k_cosine_similarity(q[qi], b, k)
            kk = min(k, n)
            check_topk(fi[qi][:kk], fs[qi][:kk], ri, rs, what=f"few n={n} q{qi}")
            assert np.all(fi[qi][kk:] == -1)


def test_lattice_1m_planted_families(cuda_device):
    """1M-row lattice bank generated on the device: the top-10 of every query is its planted family
    (members 0..9, spread over the whole bank), identical from both kernels, scores bit-equal to the
    streaming oracle on a sample of queries."""
    from hippomm_b200 import MemoryBank, synth

    n, d, nq, seed = 1_000_000, 1024, 256, 4
    bank = MemoryBank(n, d)
    for r0 in range(0, n, 1 << 17):
        m = min(1 << 17, n - r0)
        bank.fill(r0, synth.lattice_rows_torch(seed, r0, m, d, n, cuda_device))
    q, fam = synth.lattice_queries_np(seed, nq, d, n)
    expect = synth.lattice_expected_topk(fam, n, 10)
    bi, bs = _search(bank, q, 10, "batched")
    assert np.array_equal(np.sort(bi, axis=1), expect)
    si, ss = _search(bank, q[:8], 10, "single")
    assert np.array_equal(si, bi[:8])
    assert np.array_equal(ss.view(np.uint32), bs[:8].view(np.uint32))
    # streaming CPU oracle over the candidate rows of 4 queries + a random slab (full 1M is the bench's job)
    rows = np.unique(np.concatenate([expect[:4].reshape(-1), np.arange(5000, 9000)]))
    sub = synth.lattice_rows_np(seed, rows, d, n)
    for qi in range(4):
        ri, rs = O.top_k_cosine_similarity(q[qi], sub, 10)
        assert np.array_equal(rows[ri], bi[qi])
        assert np.array_equal(rs.view(np.uint32), bs[qi].view(np.uint32))


def test_lattice_10m_full_size(cuda_device):
    """BASELINE config 4 at its full size (10M x 1024, 20.5 GB of bf16 rows, generated in place): every query's
    top-10 is its planted family from BOTH kernels, the two kernels agree bit for bit, and the scores equal the
    oracle's on the candidate rows.  The full 4,096-query batch of the config is the bench's parity gate."""
    from hippomm_b200 import MemoryBank, synth

    n, d, nq, seed = 10_000_000, 1024, 512, 4
    free, _ = torch.cuda.mem_get_info()
    if free < 30 << 30:
        pytest.skip("needs 30 GB of free device memory")
    bank = MemoryBank(n, d)
    for r0 in range(0, n, 1 << 17):
        m = min(1 << 17, n - r0)
        bank.fill(r0, synth.lattice_rows_torch(seed, r0, m, d, n, cuda_device))
    q, fam = synth.lattice_queries_np(seed, nq, d, n)
    expect = synth.lattice_expected_topk(fam, n, 10)
    bi, bs = _search(bank, q, 10, "batched")
    assert np.array_equal(np.sort(bi, axis=1), expect)
    assert np.all(np.diff(bs, axis=1) <= 0)
    si, ss = _search(bank, q[:4], 10, "single")
    assert np.array_equal(si, bi[:4])
    assert np.array_equal(ss.view(np.uint32), bs[:4].view(np.uint32))
    rows = np.unique(np.concatenate([expect[:4].reshape(-1), np.arange(9_990_000, 9_992_000)]))
    sub = synth.lattice_rows_np(seed, rows, d, n)
    for qi in range(4):
        ri, rs = O.top_k_cosine_similarity(q[qi], sub, 10)
        assert np.array_equal(rows[ri], bi[qi])
        assert np.array_equal(rs.view(np.uint32), bs[qi].view(np.uint32))
    del bank
    torch.cuda.empty_cache()


def test_row_base_and_merge(cuda_device, lib):
    """Two half-banks with row offsets + hippo_topk_merge == one full bank (the sharded-search merge, §8e)."""
    from hippomm_b200 import MemoryBank

    bank_h, queries = cases.search_config1()
    full = MemoryBank.from_rows(bank_h)
    lo = MemoryBank.from_rows(bank_h[:1100])
    hi = MemoryBank.from_rows(bank_h[1100:], row_base=1100)
    fi, fs, fk = full.search_keys(queries, 5, "batched")
    _, _, k0 = lo.search_keys(queries, 5, "batched")
    _, _, k1 = hi.search_keys(queries, 5, "batched")
    keys = torch.stack([k0, k1]).contiguous()
    nq = len(queries)
    oi = torch.empty((nq, 5), dtype=torch.int64, device=cuda_device)
    os_ = torch.empty((nq, 5), dtype=torch.float32, device=cuda_device)
    ok = torch.empty((nq, 5), dtype=torch.int64, device=cuda_device)
    from hippomm_b200 import _cuda, _lib
    _lib.check(lib.hippo_topk_merge(keys.data_ptr(), 2, nq, 5, 5, oi.data_ptr(), os_.data_ptr(), ok.data_ptr(),
                                    _cuda.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(oi, fi) and torch.equal(ok, fk)
    assert torch.equal(os_.view(torch.int32), fs.view(torch.int32))


def test_dropin_is_exact_on_config1(cuda_device):
    """The drop-in signature (host arrays in, NumPy out) never rounds the rows to bf16: on config 1's video-like
    fp32 bank the top-5 ROWS AND ORDER equal the unmodified reference's committed outputs for all 64 queries
    (neighbouring frames score within 1e-4 of each other; the bf16 bank alone swaps some of them)."""
    from hippomm_b200 import MemoryBank, top_k_cosine_similarity

    bank_h, queries = cases.search_config1()
    g = cases.golden()
    for qi in range(len(queries)):
        idx, sim = top_k_cosine_similarity(queries[qi], bank_h, 5)
        assert idx.dtype == np.int64 and sim.dtype == np.float32
        assert np.array_equal(idx, g["search_c1_idx"][qi]), f"q{qi}: {idx} vs {g['search_c1_idx'][qi]}"
        assert np.max(np.abs(sim - g["search_c1_sim"][qi])) < 5e-7
    # the persistent bank reaches the same answer through exact=True (bf16 candidates, fp32 re-scoring)
    bank = MemoryBank.from_rows(bank_h, keep_rows=True)
    for path in ("single", "batched"):
        idx, sim = bank.search(queries, 5, path=path, exact=True)
        assert bank.exact_complete
        assert np.array_equal(idx.cpu().numpy(), g["search_c1_idx"])
        assert np.max(np.abs(sim.cpu().numpy() - g["search_c1_sim"])) < 5e-7
    with pytest.raises(ValueError):
        MemoryBank.from_rows(bank_h).search(queries, 5, exact=True)       # no original rows kept


@pytest.mark.parametrize("n,d,k", [(1, 3, 1), (7, 5, 7), (100, 130, 40), (1000, 1024, 10), (513, 2048, 16)])
@pytest.mark.parametrize("bdt,qdt", [(np.float32, np.float32), (np.float64, np.float32), (np.float64, np.float64),
                                     (np.float32, np.float64)])
def test_rows_path_dtypes_and_shapes(cuda_device, n, d, k, bdt, qdt):
    """hippo_topk_rows over fp32 / fp64 rows with fp32 / fp64 queries, dimensions that are not multiples of 4,
    k beyond HIPPO_TOPK_MAX (cursor paging): rows, order and result dtype as NumPy's promotion rules give them
    (vo:178-185), scores to the precision of the arrays."""
    from hippomm_b200 import cosine_similarity, top_k_cosine_similarity

    rng = np.random.default_rng(n * 131 + d)
    b = rng.standard_normal((n, d)).astype(bdt)
    q = rng.standard_normal(d).astype(qdt)
    if n > 4:
        b[3] = 0.0                                             # NaN score: first (vo:185)
        b[4] = b[2]                                            # exact tie: lower row first
    idx, sim = top_k_cosine_similarity(q, b, k)
    ri, rs = O.top_k_cosine_similarity(q, b, k)
    assert sim.dtype == rs.dtype and len(idx) == min(k, n)
    tol = 1e-6 if (bdt == np.float32 and qdt == np.float32) else 1e-7
    check_topk(idx, sim, ri, rs, tol=tol, what=f"rows n={n} d={d}")
    if n > 4 and k >= 5:
        assert idx[0] == 3 and np.isnan(sim[0])
        li = list(idx)
        if 2 in li and 4 in li:
            assert li.index(4) == li.index(2) + 1
    # a short vector pair: the case where a bf16 bank would miss by more than 1e-3
    x = np.array([0.1234567, -2.3456789, 3.4567891], dtype=bdt)
    assert abs(float(cosine_similarity(x, x)) - 1.0) < 1e-6
    assert abs(float(cosine_similarity(x, q[:3].astype(qdt))) - float(O.cosine_similarity(x, q[:3].astype(qdt)))) < 1e-6


def test_bank_cache_only_trusts_read_only_arrays(cuda_device):
    """install(cache_banks=True): a device bank is reused only for arrays the caller marked read-only; a writeable
    array is uploaded on every call, so an in-place edit between two calls is always seen."""
    import hippomm_b200 as hb
    from hippomm_b200 import vector_ops

    rng = np.random.default_rng(9)
    b = rng.standard_normal((2000, 1024)).astype(np.float32)
    q = b[1234] + 0.1 * rng.standard_normal(1024).astype(np.float32)
    hb.set_bank_cache(4)
    try:
        i0, _ = hb.top_k_cosine_similarity(q, b, 3)
        assert i0[0] == 1234 and len(vector_ops._bank_cache) == 0          # writeable: never cached
        b[1234] = -b[1234]                                                 # one row edited in place
        i1, _ = hb.top_k_cosine_similarity(q, b, 3)
        assert i1[0] != 1234
        ri, _ = O.top_k_cosine_similarity(q, b, 3)
        assert np.array_equal(i1, ri)
        b.flags.writeable = False
        i2, _ = hb.top_k_cosine_similarity(q, b, 3)
        assert len(vector_ops._bank_cache) == 1 and np.array_equal(i2, ri)
        bank = next(iter(vector_ops._bank_cache.values()))[2]
        i3, _ = hb.top_k_cosine_similarity(q, b, 3)
        assert next(iter(vector_ops._bank_cache.values()))[2] is bank and np.array_equal(i3, ri)
        v = b[:1000]                                                       # read-only view of a read-only base: fine
        hb.top_k_cosine_similarity(q, v, 3)
        assert len(vector_ops._bank_cache) == 2
        w = np.array(b)                                                    # writeable copy, and a read-only VIEW of it
        wv = w[:500]
        wv.flags.writeable = False
        hb.top_k_cosine_similarity(q, wv, 3)
        assert len(vector_ops._bank_cache) == 2                            # its base can still be edited: not cached
        hb.invalidate_bank_cache(b)
        assert len(vector_ops._bank_cache) == 1
        del v
        import gc
        gc.collect()
        assert len(vector_ops._bank_cache) == 0                            # entries die with their arrays
    finally:
        hb.set_bank_cache(0)
