"""CPU: host-side logic of the drop-in layer (no kernels run here)."""
import numpy as np
import pytest
import torch

import hostref
from hippomm_b200 import synth
from hippomm_b200.distributed import shard_range
from oracle import hippo_oracle as O

needs_no_gpu = pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a CUDA device")


def test_synth_numpy_and_torch_generators_agree():
    for n_total, r0, m, d in ((10_000, 100, 64, 256), (1_000_000, 999_900, 100, 128), (80_000_000, 79_999_990, 10, 64)):
        a = synth.lattice_rows_np(4, np.arange(r0, r0 + m), d, n_total)
        b = synth.lattice_rows_torch(4, r0, m, d, n_total, "cpu").numpy()
        assert np.array_equal(a, b)
        assert np.abs(a * 128).max() <= 127 and np.array_equal(a * 128, np.rint(a * 128))
        assert np.array_equal(synth.round_to_bf16(a), a)              # bf16-exact
    q, fam = synth.lattice_queries_np(4, 16, 64, 10_000)
    assert q.shape == (16, 64) and fam.max() < synth.lattice_families(10_000)
    assert np.array_equal(synth.round_to_bf16(q), q)


def test_lattice_dot_products_are_order_independent_in_fp32():
    """The property the bit-exact parity rests on: every partial sum is an integer below 2^24 (units 2^-14)."""
    rows = synth.lattice_rows_np(4, np.arange(0, 4000, 7), 1024, 8192)
    q, _ = synth.lattice_queries_np(4, 4, 1024, 8192)
    ints_r = np.rint(rows * 128).astype(np.int64)
    ints_q = np.rint(q * 128).astype(np.int64)
    assert (np.abs(ints_r) @ np.abs(ints_q).T).max() < 2 ** 24
    exact = (ints_r @ ints_q.T).astype(np.float64) / 2 ** 14
    assert np.array_equal(np.dot(rows, q.T).astype(np.float64), exact)
    perm = np.random.default_rng(0).permutation(1024)
    assert np.array_equal(np.dot(rows[:, perm], q[:, perm].T).astype(np.float64), exact)


def test_round_to_bf16_matches_torch():
    x = np.random.default_rng(1).standard_normal(10_000).astype(np.float32) * 37.0
    assert np.array_equal(synth.round_to_bf16(x), torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy())


def test_order_keys_sort_like_the_canonical_rule():
    rng = np.random.default_rng(2)
    s = rng.standard_normal(300).astype(np.float32)
    s[[5, 77]] = np.nan
    s[[10, 11, 12]] = s[10]
    s[200] = np.inf
    s[201] = -np.inf
    s[202], s[203] = 0.0, -0.0
    keys = hostref.pack_keys(s, np.arange(300))
    order = np.argsort(keys)[::-1]
    ci, _ = O.canonical_topk(s, 300)
    # -0.0 < +0.0 in key order while they compare equal as floats; every other position must agree
    mism = np.nonzero(order != ci)[0]
    assert set(order[mism].tolist()) <= {202, 203}
    assert order[0] == 5 and order[1] == 77 and order[2] == 200 and order[-1] == 201
    assert np.array_equal(hostref.key_rows(keys), np.arange(300))
    assert keys.min() > 0                                               # 0 is reserved for "empty slot"


@pytest.mark.parametrize("n,world", [(10, 1), (10, 3), (7, 8), (10_000_000, 8), (80_000_000, 8), (0, 4)])
def test_shard_ranges_partition_the_bank(n, world):
    edges = [shard_range(n, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == n
    for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in edges]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(n, world, world)


def test_trivial_inputs_never_touch_the_device():
    import hippomm_b200 as hb

    idx, sims = hb.top_k_cosine_similarity(np.ones(64, np.float32), np.zeros((0, 64), np.float32), 3)
    assert idx.shape == (0,) and idx.dtype == np.int64 and sims.shape == (0,)
    assert np.array_equal(hb.select_key_frames(np.ones((2, 8), np.float32), None), np.arange(2))   # hm:946-947
    assert np.array_equal(hb.select_key_frames(np.ones((0, 8), np.float32), None), np.arange(0))
    assert hb.segment_sequence() == []                                                          # hm:1024-1025
    assert hb.segment_sequence(None, None, np.zeros(0), 16000) == []
    assert hb.segment_sequence(["a"], [], None, None) == []
    with pytest.raises(ValueError):
        hb.segment_sequence(None, None, np.zeros(16000), 16000, min_segment_duration=0.0)


@needs_no_gpu
def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point raises; nothing is silently computed on the host."""
    import hippomm_b200 as hb

    b = np.random.default_rng(0).standard_normal((10, 64)).astype(np.float32)
    for call in (
        lambda: hb.top_k_cosine_similarity(b[0], b, 3),
        lambda: hb.cosine_similarity(b[0], b[1]),
        lambda: hb.select_key_frames(b, None),
        lambda: hb.compute_audio_level(np.zeros(100), 16000),
        lambda: hb.compute_frame_difference(np.zeros((8, 8, 3), np.uint8), np.zeros((8, 8, 3), np.uint8)),
        lambda: hb.segment_sequence(None, None, np.ones(32000), 16000),
        lambda: hb.MemoryBank.from_rows(b),
    ):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()


def test_install_rebinds_the_reference_symbols():
    from oracle import reference_shim

    if not reference_shim.available():
        pytest.skip("reference checkout not present")
    import hippomm_b200 as hb

    ref = reference_shim.load()          # registers the stubs and puts the reference on sys.path
    vo, hm, bp = ref.modules
    orig = (vo.top_k_cosine_similarity, hm.top_k_cosine_similarity, hm.HippocampalMemory._select_key_frames,
            hm.HippocampalMemory._segment_sequence, bp.compute_frame_difference)
    hb.install()
    try:
        assert vo.top_k_cosine_similarity is hb.top_k_cosine_similarity
        assert hm.top_k_cosine_similarity is hb.top_k_cosine_similarity       # the from-import copy (hm:28)
        assert hm.HippocampalMemory._select_key_frames is not orig[2]
        assert hm.HippocampalMemory._segment_sequence is not orig[3]
        assert bp.compute_frame_difference is hb.compute_frame_difference
        m = ref.make_memory()
        assert np.array_equal(m._select_key_frames(np.ones((2, 4), np.float32), None), np.arange(2))
        assert m._segment_sequence() == []
    finally:
        hb.uninstall()
    assert (vo.top_k_cosine_similarity, hm.top_k_cosine_similarity, hm.HippocampalMemory._select_key_frames,
            hm.HippocampalMemory._segment_sequence, bp.compute_frame_difference) == orig


def _window_class_model(e512, s, e, ns, w, db_thr):
    """Host model of csrc/segment.cu:window_class -- the same block selection and the same margins, in NumPy fp64.
    0 = not below the threshold, 1 = below, 2 = undecided (the kernel then evaluates the exact sum)."""
    pow_thr = 10.0 ** (db_thr / 10.0)
    if not (0.0 < pow_thr < 1e300) or not (db_thr > -100.0):
        return 2
    n = e - s
    if n <= 0:
        return 1
    ba, bb = s >> 9, (e - 1) >> 9
    head_in, tail_in = (s & 511) == 0, ((e & 511) == 0 or e == ns)
    head, tail = e512[ba], e512[bb]
    inner = float(np.sum(e512[ba + 1:bb])) if bb > ba + 1 else 0.0
    if ba == bb:
        lower, upper = (head if head_in and tail_in else 0.0), head
    else:
        lower = inner + (head if head_in else 0.0) + (tail if tail_in else 0.0)
        upper = inner + head + tail
    bound = pow_thr * float(n)
    if lower > bound * (1.0 + 1e-9):
        return 0
    if upper < bound * (1.0 - 1e-9):
        return 1
    return 2


def test_block_sum_classifier_never_contradicts_the_exact_level():
    """The boundary kernel decides a half-second window from the 512-sample sums when it can (csrc/segment.cu).
    Property: whenever the model of that classifier decides, the reference's own test `level < threshold`
    (hm:993-1000, hm:1073) on the exact samples says the same -- for windows at every alignment, clipped at the end
    of the stream, all-zero, louder and quieter than the threshold, and levels a hair away from it."""
    rng = np.random.default_rng(7)
    ns = 40_000 + 123
    piece = 500
    amps = rng.choice(np.array([0.0, 0.3, 0.9, 0.999, 1.0, 1.001, 1.1, 3.0]) * 0.01, size=ns // piece + 1)
    x = (rng.standard_normal(ns) * np.repeat(amps, piece)[:ns])
    k = np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)
    x = k.astype(np.float64) / 32768.0                       # what the reference sees for pcm_s16le
    sq = x * x
    e512 = np.add.reduceat(sq, np.arange(0, ns, 512))        # exact for int16-origin samples
    decided = undecided = 0
    for w in (8000, 5512, 4000, 700, 512, 37):
        for _ in range(400):
            s = int(rng.integers(0, ns))
            e = min(s + w, ns)
            if rng.random() < 0.2:
                s = (s >> 9) << 9                            # block-aligned starts
            for db_thr in (-40.0, -36.5, -50.0):
                cls = _window_class_model(e512, s, e, ns, w, db_thr)
                exact = O.compute_audio_level(x[s:e]) < db_thr
                if cls == 2:
                    undecided += 1
                else:
                    decided += 1
                    assert (cls == 1) == bool(exact), (s, e, w, db_thr, cls)
    # the empty slice past the end of the stream: -100 dB, "below" for any threshold above -100
    assert _window_class_model(e512, ns, ns, ns, 8000, -40.0) == 1 and O.compute_audio_level(x[ns:ns]) == -100
    assert _window_class_model(e512, 0, 8000, ns, 8000, -100.0) == 2      # all-zero windows need the exact rule there
    assert decided > 1000 and undecided > 50


def test_committed_bench_lines_carry_the_contract_keys():
    """The bench lines committed under profiles/ (own arm and reference arm) carry every key of the bench contract."""
    import glob
    import json
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    own = sorted(glob.glob(os.path.join(root, "profiles", "r1_v*_bench.json")))[-1]
    line = json.loads(open(own).read().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "cpu_baseline"):
        assert key in line, key
    assert line["warmup"] >= 3 and line["gpu_launches"] > 0 and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
    assert line["roofline"]["bound"] in ("hbm", "tensor")
    assert abs(line["roofline"]["frac"] - line["roofline"]["achieved"] / line["roofline"]["peak"]) < 1e-9
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(line["cpu_baseline"])
    assert line["cpu_baseline"]["kind"] in ("port", "reference")
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(line["clocks"])
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    ref = json.loads(open(own.replace("_bench.json", "_bench_reference.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["metric"] == line["metric"] and ref["unit"] == line["unit"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["d2h_bytes_per_step"] == 0
    assert ref["cpu_baseline"]["value"] == ref["value"] == ref["e2e"]["value"]


def test_frame_filter_speculation_is_decision_neutral(monkeypatch):
    """The greedy frame filters with the device scorer replaced by the oracle's SSIM / MSE: whatever
    the speculative table holds (no table, a band that keeps missing, all pairs) and however many candidates a fallback
    round takes, the decisions equal the restated reference's -- the host logic (candidate gate, table indexing,
    fallback) without a GPU."""
    import cases
    from hippomm_b200 import prefilter

    def scorer(sub, a, b, range_mode=0, out=None):
        fr = sub.numpy()
        gray = [O.bgr2gray(f) if f.shape[-1] == 3 else f[..., 0] for f in fr]
        ss, ms = [], []
        for i, j in zip(a.tolist(), b.tolist()):
            g1, g2 = gray[i].astype(np.float64), gray[j].astype(np.float64)
            ms.append(float(np.mean((g1 / 255.0 - g2 / 255.0) ** 2)))
            if min(g1.shape) < 7:
                ss.append(float("nan"))
            elif range_mode == 1:
                with np.errstate(all="ignore"):
                    ss.append(float(O.structural_similarity(g1 / 255.0, g2 / 255.0, data_range=1.0)))
            else:
                with np.errstate(all="ignore"):
                    ss.append(float(O.structural_similarity(gray[i], gray[j], data_range=float(int(gray[i].max()) - int(gray[i].min())))))
        return torch.tensor(ss, dtype=torch.float64), torch.tensor(ms, dtype=torch.float64)

    monkeypatch.setattr(prefilter, "frame_pair_scores_device", scorer)
    monkeypatch.setattr(prefilter, "_frames_to_device", lambda frames: torch.from_numpy(np.ascontiguousarray(frames)))
    frames = cases.prefilter_frames()[:150, ::2, ::2]                   # 150 frames of 48 x 48: seconds on the host
    for params in (dict(video_fps=30.0, max_diff_threshold=0.3, check_interval=10),
                   dict(video_fps=60.0, max_diff_threshold=0.05, check_interval=7),       # the 1 s gate skips candidates
                   dict(video_fps=10.0, max_diff_threshold=0.9, check_interval=1)):
        want = O.select_saved_frames(frames, **params)
        for band, window in ((0, 3), (1, 8), (2, 1), (24, 8)):
            got = prefilter.select_saved_frames(frames, band=band, window=window, **params)
            assert got[0] == want[0] and got[1] == want[1], (params, band, window)
    for win in cases.dedup_windows()[:5]:
        small = np.ascontiguousarray(win[:, ::3, ::3])
        want = O.dedup_window_frames(small, 0.3)
        for band, window in ((0, 1), (1, 4), (128, 4)):
            assert prefilter.dedup_window_frames(small, 0.3, window=window, band=band) == want, (band, window)


def test_ssim_lane_mappings_cover_every_window_and_every_pixel_exactly_once():
    """A host model of the ownership rules of csrc/frames.cu (ssim_item<4> / <7>, ssim_nchunks): for every width, each
    window column belongs to exactly one (chunk, lane, k), each image column is counted for the squared error by exactly
    one lane, and an owned window only needs columns that its own warp holds (two lanes to the right with 4 columns
    per lane, one lane with 7)."""
    def nchunks(out_cols, cpl):
        if out_cols <= 0:
            return 1
        return max(1, (out_cols - 1 + 216) // 217 if cpl == 7 else (out_cols + 119) // 120)

    for cpl, own_lanes in ((4, 30), (7, 31)):
        for w in list(range(7, 260)) + [320, 433, 434, 440, 441, 442, 600, 651, 652, 1280, 1920]:
            out_cols = w - 6
            pitch = (w + 6) // 7 * 8 if cpl == 7 else (w + 3) & ~3
            nch = nchunks(out_cols, cpl)
            win_owner = np.zeros(out_cols, dtype=np.int32)
            px_owner = np.zeros(w, dtype=np.int32)
            for c in range(nch):
                last = c == nch - 1
                for lane in range(32):
                    g = c * own_lanes + lane
                    col_ok = g * (4 if cpl == 4 else 8) < pitch
                    col0 = g * cpl
                    if col_ok and (lane < own_lanes or last):
                        px_owner[col0:min(col0 + cpl, w)] += 1            # the bytes beyond the row are zero padding
                    for k in range(cpl):
                        if cpl == 7:
                            own = (lane < own_lanes or (k == 0 and last)) and col0 + k < out_cols
                        else:
                            own = lane < own_lanes and col0 + k < out_cols
                        if own:
                            win_owner[col0 + k] += 1
                            last_col = col0 + k + 6                       # the window spans columns col0+k .. col0+k+6
                            lane_of_last = last_col // cpl - c * own_lanes
                            assert lane_of_last <= 31 and lane_of_last - lane <= (2 if cpl == 4 else 1), (cpl, w, c, lane, k)
                            assert last_col < w
            assert (win_owner == 1).all(), (cpl, w, np.nonzero(win_owner != 1)[0][:5])
            assert (px_owner == 1).all(), (cpl, w, np.nonzero(px_owner != 1)[0][:5])


def test_dp2a_permutes_of_the_ssim_row_step_give_the_window_moments():
    """The SSIM row step (csrc/frames.cu) forms, per column, bn = byte_perm(wa, wb, sel) = (x, y, y, x) and
    hn = byte_perm(bn, 0, 0x4140) = x | y << 16, then dp2a.lo(hn, bn) and dp2a.hi(hn, bn).  Emulated with PTX's
    definitions for all byte pairs and all four byte positions: x^2 + y^2, 2xy and the packed sum x | y << 16."""
    def byte_perm(x, y, sel):
        src = [(x >> (8 * i)) & 0xff for i in range(4)] + [(y >> (8 * i)) & 0xff for i in range(4)]
        return sum(src[(sel >> (4 * i)) & 7] << (8 * i) for i in range(4))

    def dp2a(a, b, c, hi):                     # a: two 16-bit halves, b: bytes 0,1 (lo) or 2,3 (hi), unsigned
        a0, a1 = a & 0xffff, a >> 16
        b0, b1 = (b >> (16 if hi else 0)) & 0xff, (b >> (24 if hi else 8)) & 0xff
        return (c + a0 * b0 + a1 * b1) & 0xffffffff

    rng = np.random.default_rng(3)
    for k in range(4):
        sel = k | ((4 + k) << 4) | ((4 + k) << 8) | (k << 12)
        for x, y in [(0, 0), (255, 255), (255, 0), (0, 255), (1, 254)] + [tuple(int(v) for v in rng.integers(0, 256, 2)) for _ in range(200)]:
            wa = int(rng.integers(0, 1 << 32)) & ~(0xff << (8 * k)) | (x << (8 * k))
            wb = int(rng.integers(0, 1 << 32)) & ~(0xff << (8 * k)) | (y << (8 * k))
            bn = byte_perm(wa, wb, sel)
            hn = byte_perm(bn, 0, 0x4140)
            assert bn == x | (y << 8) | (y << 16) | (x << 24)
            assert hn == x | (y << 16)
            assert dp2a(hn, bn, 7, hi=False) == 7 + x * x + y * y
            assert dp2a(hn, bn, 7, hi=True) == 7 + 2 * x * y
