"""CPU: the oracle restatements against the committed outputs of the unmodified reference
(tests/golden/reference_outputs.npz) and, where the reference checkout exists, against the live
reference functions."""
import numpy as np
import pytest

import cases
from oracle import hippo_oracle as O
from oracle import reference_shim
from parity import check_topk, check_topk_exact


def test_search_restatement_equals_reference_outputs():
    g = cases.golden()
    bank, queries = cases.search_config1()
    for qi, q in enumerate(queries):
        idx, sim = O.top_k_cosine_similarity(q, bank, 5)
        assert np.array_equal(idx, g["search_c1_idx"][qi])
        assert np.array_equal(sim, g["search_c1_sim"][qi])
    lbank, lq, fam = cases.search_lattice_small()
    for qi, q in enumerate(lq):
        idx, sim = O.top_k_cosine_similarity(q, lbank, 10)
        check_topk_exact(idx, sim, g["search_lat_idx"][qi], g["search_lat_sim"][qi])
    # the planted family is the top-10 and the gaps dwarf the tolerance
    from hippomm_b200 import synth
    expect = synth.lattice_expected_topk(fam, len(lbank), 10)
    assert np.array_equal(np.sort(g["search_lat_idx"], axis=1), expect)
    assert np.all(g["search_lat_sim"][:, 9] > 0.85)


def test_search_edge_cases_equal_reference_outputs():
    g = cases.golden()
    b, q = cases.search_edge()
    idx, sim = O.top_k_cosine_similarity(q, b, 4)
    assert np.array_equal(idx, g["search_edge_idx"]) and np.isnan(sim[0]) and idx[0] == 3
    idx, sim = O.top_k_cosine_similarity(q, b[:3], 8)
    assert np.array_equal(idx, g["search_kgtn_idx"]) and len(idx) == 3
    idx, sim = O.top_k_cosine_similarity(q, b[5], 3)
    assert np.array_equal(idx, g["search_1d_idx"]) and len(idx) == 1
    assert O.cosine_similarity(q, b[5]) == pytest.approx(g["cosine_pair"][0], abs=1e-7)


def test_canonical_and_streaming_topk_agree_with_full_sort():
    bank, queries = cases.search_config1()
    q = queries[:6]
    idx_s, sim_s = O.top_k_streaming(q, lambda a, b: bank[a:b], len(bank), 5, chunk_rows=300)
    for qi in range(len(q)):
        idx, sim = O.top_k_cosine_similarity(q[qi], bank, 5)
        check_topk(idx_s[qi], sim_s[qi], idx, sim, tol=1e-5)   # sgemm vs sgemv rounding
    b, qq = cases.search_edge()
    with np.errstate(all="ignore"):
        sims = np.dot(b, qq) / (np.linalg.norm(b, axis=1) * np.linalg.norm(qq))
    idx, sc = O.canonical_topk(sims, 16)
    assert idx[0] == 3 and np.isnan(sc[0])
    assert list(idx).index(9) == list(idx).index(2) + 1          # tie: lower row first
    # streaming on the lattice bank: bit-equal to the one-shot reference arithmetic
    lbank, lq, _ = cases.search_lattice_small()
    g = cases.golden()
    idx_s, sim_s = O.top_k_streaming(lq[:8], lambda a, b: lbank[a:b], len(lbank), 10, chunk_rows=1000)
    for qi in range(8):
        check_topk_exact(idx_s[qi], sim_s[qi], g["search_lat_idx"][qi], g["search_lat_sim"][qi])


def test_consolidation_restatements_equal_reference_outputs():
    g = cases.golden()
    for name, feats in cases.consolidation_cases().items():
        for gamma in (0.9, 0.95):
            ref = g[f"cons_{name}_{int(gamma * 100)}"]
            with np.errstate(all="ignore"):
                assert np.array_equal(O.select_key_frames(feats, None, gamma), ref), (name, gamma)
                blocked, moat = O.select_key_frames_blocked(feats, gamma, block=97, with_moat=True)
            assert np.array_equal(blocked, ref), (name, gamma, "blocked")
            if len(feats) > 2 and name not in ("zero", "zero0"):
                ok, why = O.greedy_valid_under_tolerance(feats, ref, gamma)
                assert ok, (name, gamma, why)
    assert np.array_equal(O.select_key_frames(cases.consolidation_cases()["c1"], None), g["cons_default_gamma"])
    assert g["cons_zero0_90"].tolist() == [0]                    # zero-norm row 0 blocks everything after it
    assert 17 not in g["cons_zero_90"].tolist()


def test_greedy_validity_checker_rejects_wrong_answers():
    feats = cases.consolidation_cases()["clustered"]
    ref = cases.golden()["cons_clustered_90"]
    assert ref.tolist() == list(range(0, 600, 20))
    bad = np.delete(ref, 3)
    assert not O.greedy_valid_under_tolerance(feats, bad, 0.9)[0]
    bad = np.sort(np.append(ref, 5))
    assert not O.greedy_valid_under_tolerance(feats, bad, 0.9)[0]


def test_audio_level_restatement_equals_reference_outputs():
    g = cases.golden()
    pcm = cases.audio_case()
    x = pcm.astype(np.float64) / 32768.0
    got = np.array([O.compute_audio_level(x[s:s + n].reshape(-1, 1)) for s, n in cases.audio_windows(len(pcm))])
    assert np.array_equal(got, g["audio_levels"])
    assert O.compute_audio_level(np.zeros(100)) == -100 == g["audio_level_zero"][0]
    assert O.compute_audio_level(np.zeros(0)) == -100             # empty slice (hm:1071 past the end)
    # the planted structure is what config 2 asks for: loud ~ -20 dB, silences ~ -80 dB
    assert -21 < g["audio_levels"][0] < -19


def test_frame_restatements_equal_reference_outputs():
    g = cases.golden()
    frames, cuts = cases.frame_case()
    got = np.array([O.compute_frame_difference(frames[a], frames[b]) for a, b in cases.frame_diff_pairs()])
    assert np.array_equal(got, g["frame_diff"])
    const = cases.constant_frames()
    got = np.array([O.compute_frame_difference(const[a], const[b]) for a, b in ((0, 0), (0, 1), (1, 2))])
    assert np.array_equal(got, g["frame_diff_const"])
    ss = O.adjacent_ssim(frames)
    assert np.array_equal(ss, g["frame_ssim_adjacent"])
    # scene cuts are the low-SSIM pairs, and nothing sits inside the +-2e-3 moat around 0.95
    assert np.array_equal(np.nonzero(ss < 0.95)[0] + 1, np.nonzero(cuts)[0][1:])
    assert np.min(np.abs(ss - 0.95)) > 2e-3
    with np.errstate(all="ignore"):
        c = [O.compute_frame_similarity(const[a], const[b]) for a, b in ((0, 1), (0, 2), (2, 0))]
    assert np.array_equal(np.isnan(c), np.isnan(g["frame_ssim_const"]))


def test_ssim_restatement_against_exact_integer_moments():
    """SSIM is 'vs restated oracle' (scikit-image absent).  Bound the restatement's own fp64 noise with an
    independent evaluation from exact integer window sums."""
    frames, _ = cases.frame_case()
    for p in (0, 29, 30, 31, 100):
        g1, g2 = O.bgr2gray(frames[p + 1]), O.bgr2gray(frames[p])
        r = float(g1.max() - g1.min())
        a = O.structural_similarity(g1, g2, data_range=g1.max() - g1.min())
        b = O.structural_similarity_exact(g1, g2, r)
        assert abs(a - b) < 1e-12


def test_bgr2gray_equals_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(257, 263, 3), dtype=np.uint8)
    assert np.array_equal(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), O.bgr2gray(img))
    # all 2^24 colours in 16 slabs
    b, g = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    for r0 in range(0, 256, 16):
        slab = np.stack([np.stack([b, g, np.full_like(b, r)], axis=-1) for r in range(r0, r0 + 16)])
        slab = slab.reshape(16 * 256, 256, 3)
        assert np.array_equal(cv2.cvtColor(slab, cv2.COLOR_BGR2GRAY), O.bgr2gray(slab))


@pytest.mark.parametrize("tag", ["av", "a", "v", "av_short", "v_short"])
def test_boundary_state_machine_equals_reference_outputs(tag):
    g = cases.golden()
    frames, _ = cases.frame_case()
    kw = cases.segmentation_variants()[tag]
    times = cases.frame_times(len(frames))
    x = cases.audio_case().astype(np.float64) / 32768.0
    audio = x[: int(times[-1] * 16000) + 16000].reshape(-1, 1)
    bounds = O.segment_boundaries(g["frame_ssim_adjacent"] if kw["video"] else None, times if kw["video"] else None,
                                  audio if kw["audio"] else None, 16000 if kw["audio"] else None, **kw["thresholds"])
    assert np.array_equal(np.array(bounds).reshape(-1, 2), g[f"seg_{tag}"])


@pytest.mark.skipif(not reference_shim.available(), reason="reference checkout not present (GPU box)")
def test_live_reference_agrees_with_restatements():
    """Where /root/reference exists: run the unmodified functions again, on fresh seeds."""
    ref = reference_shim.load()
    mem = ref.make_memory()
    rng = np.random.default_rng(99)
    b = rng.standard_normal((500, 1024)).astype(np.float32)
    a = rng.standard_normal(1024).astype(np.float32)
    i1, s1 = ref.top_k_cosine_similarity(a, b, 7)
    i2, s2 = O.top_k_cosine_similarity(a, b, 7)
    assert np.array_equal(i1, i2) and np.array_equal(s1, s2)
    from hippomm_b200 import synth
    f = synth.videolike_features(77, 12, 40)
    assert np.array_equal(mem._select_key_frames(f, None, 0.9), O.select_key_frames_blocked(f, 0.9, block=64))
    x = rng.standard_normal((16000 * 130, 1)) * 0.05
    x[16000 * 40:16000 * 42] *= 1e-4
    x[16000 * 77:int(16000 * 78.3)] *= 1e-4
    segs = mem._segment_sequence(None, None, x, 16000)
    assert [(s.start_time, s.end_time) for s in segs] == O.segment_boundaries(None, None, x, 16000)
    fr, _ = synth.frame_stream(5, 4, 40, 48, min_scene=2, max_scene=2)
    assert ref.compute_frame_difference(fr[0], fr[2]) == O.compute_frame_difference(fr[0], fr[2])


def test_recall_restatement_matches_reference_outputs():
    """hm:3127-3383 (similarity path) restated vs the committed outputs of the reference's own functions."""
    events, queries = cases.recall_events()
    g = cases.golden()
    for name, (modality, q) in queries.items():
        segs = O.find_relevant_segments(q, events, modality)
        bounds = np.array([[s["start"], s["end"]] for s in segs], dtype=np.float64).reshape(-1, 2)
        assert np.array_equal(bounds, g[f"recall_{name}_bounds"]), name
        if modality == "vision":
            assert [len(s["frames"]) for s in segs] == g[f"recall_{name}_nframes"].tolist()
            assert [t for s in segs for t in s["frame_times"]] == g[f"recall_{name}_frame_times"].tolist()


def _decoded_prefilter_frames(tmp_path):
    import hashlib

    avi = str(tmp_path / "v.avi")
    if not cases.write_mjpg(cases.prefilter_frames(), avi, cases.PREFILTER_PARAMS["video_fps"]):
        pytest.skip("OpenCV cannot write MJPG here")
    decoded = cases.read_video(avi)
    sha = np.frombuffer(hashlib.sha1(decoded.tobytes()).digest(), dtype=np.uint8)
    if not np.array_equal(sha, cases.golden()["prefilter_decoded_sha1"]):
        pytest.skip("this OpenCV build decodes the MJPG stream to different pixels than the golden run")
    return decoded


def test_prefilter_restatement_matches_reference_outputs(tmp_path):
    """bp:179-228 restated vs the committed decisions of the reference's own extract_frames_from_video."""
    decoded = _decoded_prefilter_frames(tmp_path)
    g = cases.golden()
    numbers, times = O.select_saved_frames(decoded, **cases.PREFILTER_PARAMS)
    assert numbers == g["prefilter_frame_numbers"].tolist()
    assert times == g["prefilter_frame_times"].tolist()
