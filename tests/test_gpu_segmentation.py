"""GPU parity: temporal pattern separation (hm:980-1114, bp:32-71) through the C ABI."""
import numpy as np
import pytest
import torch

import cases
from oracle import hippo_oracle as O

pytestmark = pytest.mark.gpu

SSIM_TOL = 1e-6      # fp32 per-window ratio + fp64 mean vs the fp64 restatement
DB_TOL = 1e-9        # fp64 on both sides; only log10's last ulp may differ


def _ssim_adjacent(frames, dev, range_mode=0):
    from hippomm_b200.segmentation import frame_pair_scores_device

    fd = torch.from_numpy(frames).to(dev)
    ssim, mse = frame_pair_scores_device(fd, range_mode=range_mode)
    torch.cuda.synchronize()
    return ssim.cpu().numpy(), mse.cpu().numpy()


def test_frame_ssim_matches_reference_outputs(cuda_device):
    frames, _ = cases.frame_case()
    g = cases.golden()
    ssim, _ = _ssim_adjacent(frames, cuda_device)
    assert np.max(np.abs(ssim - g["frame_ssim_adjacent"])) < SSIM_TOL
    # the decisions the state machine takes from them are identical
    assert np.array_equal(ssim < 0.95, g["frame_ssim_adjacent"] < 0.95)


@pytest.mark.parametrize("h,w", [(7, 7), (8, 300), (57, 63), (64, 64), (119, 257), (224, 224), (180, 320), (70, 600),
                                 (62, 126), (63, 127), (9, 247), (118, 246), (61, 13), (20, 223), (15, 225), (30, 441),
                                 (12, 440), (64, 196), (9, 373)])
def test_frame_pairs_shapes_against_oracle(cuda_device, h, w):
    """Frame sizes around the band (56 window rows) and column-chunk boundaries (120 windows per warp with 4 columns
    per lane, 217 (+1) with 7: widths 127-224 and 367-441 take the 7-in-8 gray layout unless only the 4-column layout has
    a vectorised gray path for the size), widths that are / are not
    multiples of 4 or 7 (row padding of the gray buffer) and of 16 / 14 x 32 (vectorised gray paths)."""
    from hippomm_b200 import synth

    frames, _ = synth.frame_stream(h * 1000 + w, 6, h, w, min_scene=2, max_scene=3)
    ssim, mse = _ssim_adjacent(frames, cuda_device)
    ref = O.adjacent_ssim(frames)
    assert np.max(np.abs(ssim - ref)) < SSIM_TOL
    for p in range(len(frames) - 1):
        g1 = O.bgr2gray(frames[p + 1]).astype(np.float64) / 255.0
        g0 = O.bgr2gray(frames[p]).astype(np.float64) / 255.0
        assert abs(mse[p] - np.mean((g1 - g0) ** 2)) < 1e-12
    # gray (single channel) input takes the same path
    gray = np.stack([O.bgr2gray(f) for f in frames])[..., None]
    ssim_g, _ = _ssim_adjacent(np.ascontiguousarray(gray), cuda_device)
    assert np.array_equal(ssim_g, ssim)


@pytest.mark.parametrize("cpl", [4, 7])
def test_frame_pairs_with_either_gray_layout_forced(cuda_device, cpl):
    """HIPPO_SSIM_CPL forces one lane mapping / gray layout for every width (read once per process, hence the child
    process): shapes that cross the chunk boundaries of both mappings, each against the oracle."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np, torch\n"
        f"sys.path.insert(0, {root!r})\n"
        "from hippomm_b200 import synth\n"
        "from hippomm_b200.segmentation import frame_pair_scores_device\n"
        "from oracle import hippo_oracle as O\n"
        "dev = torch.device('cuda', 0)\n"
        "for h, w in [(7, 7), (9, 13), (64, 64), (30, 126), (30, 127), (20, 223), (224, 224), (15, 225), (12, 440),\n"
        "             (30, 441), (9, 442), (70, 600), (180, 320)]:\n"
        "    frames, _ = synth.frame_stream(h * 1000 + w, 4, h, w, min_scene=2, max_scene=2)\n"
        "    ssim, mse = frame_pair_scores_device(torch.from_numpy(frames).to(dev), range_mode=0)\n"
        "    ref = O.adjacent_ssim(frames)\n"
        "    got = ssim.cpu().numpy()\n"
        "    assert np.max(np.abs(got - ref)) < 1e-6, (h, w, got, ref)\n"
        "    for p in range(len(frames) - 1):\n"
        "        g1 = O.bgr2gray(frames[p + 1]).astype(np.float64) / 255.0\n"
        "        g0 = O.bgr2gray(frames[p]).astype(np.float64) / 255.0\n"
        "        assert abs(float(mse[p]) - np.mean((g1 - g0) ** 2)) < 1e-12, (h, w, p)\n"
        "print('layout ok')\n"
    )
    env = dict(os.environ, HIPPO_SSIM_CPL=str(cpl))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "layout ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_frame_similarity_and_difference_wrappers(cuda_device):
    from hippomm_b200 import compute_frame_difference, compute_frame_similarity

    frames, _ = cases.frame_case()
    g = cases.golden()
    for (a, b), want in zip(cases.frame_diff_pairs(), g["frame_diff"]):
        assert abs(compute_frame_difference(frames[a], frames[b]) - want) < SSIM_TOL
    const = cases.constant_frames()
    for (a, b), want in zip(((0, 0), (0, 1), (1, 2)), g["frame_diff_const"]):
        assert abs(compute_frame_difference(const[a], const[b]) - want) < SSIM_TOL
    # constant first frame: data range 0 -> 0/0 -> NaN exactly where the reference gives NaN (hm:990)
    for (a, b), want in zip(((0, 1), (0, 2), (2, 0)), g["frame_ssim_const"]):
        got = compute_frame_similarity(const[a], const[b])
        assert (np.isnan(got) and np.isnan(want)) or abs(got - want) < SSIM_TOL
    assert abs(compute_frame_similarity(frames[31], frames[30]) - g["frame_ssim_adjacent"][30]) < SSIM_TOL
    tiny = np.zeros((5, 5, 3), dtype=np.uint8)
    with pytest.raises(ValueError):
        compute_frame_similarity(tiny, tiny)                    # skimage raises for images below 7x7
    assert compute_frame_difference(tiny, tiny + 255) == 1.0     # bp:64-71: exception -> MSE fallback


def test_audio_levels_match_reference_outputs(cuda_device):
    from hippomm_b200 import compute_audio_level
    from hippomm_b200.segmentation import audio_energy_device, audio_levels_device

    pcm = cases.audio_case()
    g = cases.golden()
    wins = cases.audio_windows(len(pcm))
    ws = torch.tensor([w[0] for w in wins], dtype=torch.int64)
    wl = torch.tensor([w[1] for w in wins], dtype=torch.int64)
    x64 = pcm.astype(np.float64) / 32768.0
    for name, arr in (("i16", pcm), ("f32", x64.astype(np.float32)), ("f64", x64)):
        t = torch.from_numpy(arr).to(cuda_device).reshape(-1, 1)
        pyr = audio_energy_device(t)
        direct = audio_levels_device(t, ws, wl).cpu().numpy()
        viapyr = audio_levels_device(t, ws, wl, pyramid=pyr).cpu().numpy()
        assert np.max(np.abs(direct - g["audio_levels"])) < DB_TOL, name
        assert np.max(np.abs(viapyr - g["audio_levels"])) < DB_TOL, name
        # int16-origin PCM: all sums of squares are exact in fp64, whatever the order
        e16, e512 = pyr
        ref16 = np.add.reduceat(x64 * x64, np.arange(0, len(x64), 16))
        assert np.array_equal(e16.cpu().numpy(), ref16), name
        assert np.array_equal(e512.cpu().numpy(), np.add.reduceat(x64 * x64, np.arange(0, len(x64), 512))), name
    stereo = np.stack([x64[:8000], x64[8000:16000]], axis=1)
    assert abs(compute_audio_level(stereo, 16000) - g["audio_level_stereo"][0]) < DB_TOL
    assert compute_audio_level(np.zeros(100), 16000) == -100
    assert compute_audio_level(x64[:8000].reshape(-1, 1), 16000) == pytest.approx(g["audio_levels"][0], abs=DB_TOL)
    # integer samples: taken at face value by default (what hm:995-998 computes for a 2-D integer array),
    # read as PCM k / 32768 only on request
    raw = pcm[:8000].reshape(-1, 1)
    assert abs(compute_audio_level(raw, 16000) - O.compute_audio_level(raw)) < DB_TOL
    assert compute_audio_level(raw, 16000, int16_pcm=True) == pytest.approx(g["audio_levels"][0], abs=DB_TOL)
    # ragged length, stereo, through the generic pyramid kernel
    st = torch.from_numpy(np.ascontiguousarray(np.stack([x64[:100003], x64[7:100010]], axis=1))).to(cuda_device)
    pyr = audio_energy_device(st)
    lv = audio_levels_device(st, torch.tensor([5, 90000]), torch.tensor([8000, 20000]), pyramid=pyr).cpu().numpy()
    mono = st.cpu().numpy().mean(axis=1)
    assert abs(lv[0] - O.compute_audio_level(mono[5:8005])) < DB_TOL
    assert abs(lv[1] - O.compute_audio_level(mono[90000:100003])) < DB_TOL      # window clipped like a NumPy slice


@pytest.mark.parametrize("tag", ["av", "a", "v", "av_short", "v_short"])
def test_segment_sequence_matches_reference_outputs(cuda_device, tag):
    """Boundaries equal to the unmodified reference's (exact fp64), plus the frames / samples per segment."""
    from hippomm_b200 import segment_sequence

    frames, _ = cases.frame_case()
    g = cases.golden()
    kw = cases.segmentation_variants()[tag]
    times = cases.frame_times(len(frames))
    pcm = cases.audio_case()
    x = pcm.astype(np.float64) / 32768.0
    audio = x[: int(times[-1] * 16000) + 16000].reshape(-1, 1)
    segs = segment_sequence(frames if kw["video"] else None, times if kw["video"] else None,
                            audio if kw["audio"] else None, 16000 if kw["audio"] else None, **kw["thresholds"])
    got = np.array([[s.start_time, s.end_time] for s in segs]).reshape(-1, 2)
    assert np.array_equal(got, g[f"seg_{tag}"]), f"{got} vs {g[f'seg_{tag}']}"
    nfr = [len(s.frames) if s.frames is not None else -1 for s in segs]
    nsm = [len(s.audio_data) if s.audio_data is not None else -1 for s in segs]
    assert nfr == g[f"seg_{tag}_nframes"].tolist()
    assert nsm == g[f"seg_{tag}_nsamples"].tolist()


def test_segment_edge_cases(cuda_device):
    from hippomm_b200 import segment_sequence

    assert segment_sequence() == []
    assert segment_sequence(None, None, np.zeros(0), 16000) == []
    # audio shorter than the video: windows past the end are empty -> -100 dB -> "silent" (NumPy slice clipping)
    frames, _ = cases.frame_case()
    times = cases.frame_times(60)
    x = cases.audio_case()[: 20 * 16000].astype(np.float64) / 32768.0
    segs = segment_sequence(frames[:60], times, x.reshape(-1, 1), 16000)
    ss = O.adjacent_ssim(frames[:60])
    want = O.segment_boundaries(ss, times, x.reshape(-1, 1), 16000)
    assert [(s.start_time, s.end_time) for s in segs] == want
    # int16 PCM on the device side (opt-in) gives the same boundaries as the float64 the reference sees from sf.read
    segs16 = segment_sequence(None, None, cases.audio_case()[: 120 * 16000], 16000, int16_pcm=True)
    x120 = cases.audio_case()[: 120 * 16000].astype(np.float64) / 32768.0
    want = O.segment_boundaries(None, None, x120, 16000)
    assert [(s.start_time, s.end_time) for s in segs16] == want
    # without the opt-in, integer samples are taken at face value like the reference's NumPy arithmetic does
    # (levels ~ +70 dB: nothing is below -40 dB, every segment runs to max_segment_duration)
    k2 = cases.audio_case()[: 120 * 16000].reshape(-1, 1)
    segs_raw = segment_sequence(None, None, k2, 16000)
    want_raw = O.segment_boundaries(None, None, k2, 16000)
    assert [(s.start_time, s.end_time) for s in segs_raw] == want_raw
    assert want_raw != want


def test_one_hour_stream_against_oracle(cuda_device):
    """Config 2 geometry at reduced resolution: 3,600 frames + 57.6M int16 samples; boundaries must equal
    the oracle's exactly (the oracle needs ~20 s for the SSIM restatement at 64x64)."""
    from hippomm_b200 import segment_sequence, synth

    frames, _ = synth.frame_stream(21, 3600, 64, 64)
    pcm = synth.audio_stream_int16(22, 3600 * 16000)
    times = cases.frame_times(3600)
    segs = segment_sequence(frames, times, pcm.reshape(-1, 1), 16000, int16_pcm=True)
    ss = O.adjacent_ssim(frames)
    x = pcm.astype(np.float64) / 32768.0
    want = O.segment_boundaries(ss, times, x.reshape(-1, 1), 16000)
    assert [(s.start_time, s.end_time) for s in segs] == want
    assert 100 <= len(segs) <= 400


def test_batched_streams_match_single_stream_calls(cuda_device):
    """Three different streams (video+audio, audio only, video only) through one launch == one launch each."""
    from hippomm_b200 import synth
    from hippomm_b200.segmentation import (audio_energy_device, frame_pair_scores_device,
                                           segment_boundaries_batch_device, segment_boundaries_device)

    sr = 16000
    streams = []
    for seed, (nsec, video, audio) in enumerate([(150, True, True), (95, False, True), (120, True, False)]):
        ssim = ft = pcm = pyr = None
        if video:
            frames, _ = synth.frame_stream(10 + seed, nsec, 48, 64, min_scene=5, max_scene=25)
            fd = torch.from_numpy(frames).to(cuda_device)
            ssim, _ = frame_pair_scores_device(fd, range_mode=0)
            ft = torch.arange(nsec, dtype=torch.float64, device=cuda_device)
        if audio:
            pcm = torch.from_numpy(synth.audio_stream_int16(20 + seed, nsec * sr).reshape(-1, 1)).to(cuda_device)
            pyr = audio_energy_device(pcm)
        streams.append((ssim, ft, pcm, pyr, sr if audio else None))
    bounds, counts = segment_boundaries_batch_device(streams, 30.0, 10.0, 0.95, -40.0, 64)
    torch.cuda.synchronize()
    for i, st in enumerate(streams):
        b1, c1 = segment_boundaries_device(*st, 30.0, 10.0, 0.95, -40.0, 64)
        c = int(c1.item())
        assert c > 0 and int(counts[i].item()) == c
        assert torch.equal(bounds[i, :c], b1[:c])


@pytest.mark.parametrize("lanes", [1, 2, 3])
def test_pattern_separation_batch_on_alternating_cuda_streams(cuda_device, lanes):
    """Five streams of different shapes through pattern_separation_batch_device (per-stream kernels on `lanes`
    alternating CUDA streams with per-stream scratch, one boundary launch) == the stream-by-stream calls, twice in a
    row (the second call re-uses scratch and side streams)."""
    from hippomm_b200 import synth
    from hippomm_b200.segmentation import (audio_energy_device, frame_pair_scores_device,
                                           pattern_separation_batch_device, segment_boundaries_device)

    sr = 16000
    inputs, want = [], []
    for seed, (nsec, hw, video, audio) in enumerate([(150, (48, 64), True, True), (95, None, False, True),
                                                     (120, (64, 48), True, False), (200, (56, 56), True, True),
                                                     (61, (32, 40), True, True)]):
        fd = ft = pcm = None
        if video:
            frames, _ = synth.frame_stream(30 + seed, nsec, hw[0], hw[1], min_scene=5, max_scene=25)
            fd = torch.from_numpy(frames).to(cuda_device)
            ft = torch.arange(nsec, dtype=torch.float64, device=cuda_device)
        if audio:
            pcm = torch.from_numpy(synth.audio_stream_int16(40 + seed, nsec * sr).reshape(-1, 1)).to(cuda_device)
        inputs.append((fd, ft, pcm, sr if audio else None))
        ssim = frame_pair_scores_device(fd, range_mode=0)[0] if video else None
        pyr = audio_energy_device(pcm) if audio else None
        b1, c1 = segment_boundaries_device(ssim, ft, pcm, pyr, sr if audio else None, 30.0, 10.0, 0.95, -40.0, 64)
        torch.cuda.synchronize()
        want.append(b1[: int(c1.item())].clone())
    for mode in ("stages", "pipeline", "stages", "pipeline"):
        bounds, counts = pattern_separation_batch_device(inputs, 30.0, 10.0, 0.95, -40.0, 64, lanes=lanes, mode=mode)
        torch.cuda.synchronize()
        for i, w in enumerate(want):
            assert int(counts[i].item()) == len(w) > 0, mode
            assert torch.equal(bounds[i, : len(w)], w), mode


@pytest.mark.parametrize("sr,dtype,nch", [(16000, "i16", 1), (8000, "i16", 1), (22050, "f32", 1), (44100, "i16", 2),
                                          (16000, "f64", 1), (11025, "i16", 1)])
def test_audio_scan_near_the_threshold(cuda_device, sr, dtype, nch):
    """Audio-only streams whose level hovers around the -40 dB threshold in quarter-second pieces (0.5x .. 3x the
    threshold amplitude, including 0.99x / 1.01x): most windows cannot be decided from the 512-sample bounds, so
    both the block-sum classifier and the exact four-lane evaluation are exercised, at sample rates whose
    half-second window is / is not a multiple of the 512- and 16-sample blocks.  Boundaries must equal the oracle's."""
    from hippomm_b200 import segment_sequence

    rng = np.random.default_rng(sr + nch)
    nsec = 100
    scales = np.array([3.0, 1.1, 1.01, 1.0, 0.99, 0.9, 0.5, 0.0]) * 0.01
    probs = [0.30, 0.20, 0.20, 0.10, 0.10, 0.05, 0.03, 0.02]
    piece = sr // 4
    parts = []
    for _ in range(nsec * 4):
        amp = rng.choice(scales, p=probs)
        parts.append(rng.standard_normal((piece, nch)) * amp)
    x = np.concatenate(parts)[: nsec * sr - 37]          # ragged end
    if dtype == "i16":
        k = np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)
        given, seen = k, k.astype(np.float64) / 32768.0
    elif dtype == "f32":
        given = x.astype(np.float32)
        seen = given.astype(np.float64)
    else:
        given = seen = x
    for thresholds in ({}, {"max_segment_duration": 45.0, "min_segment_duration": 3.0}):
        segs = segment_sequence(None, None, given, sr, int16_pcm=(dtype == "i16"), **thresholds)
        want = O.segment_boundaries(None, None, seen, sr, **thresholds)
        assert [(s.start_time, s.end_time) for s in segs] == want
        assert len(want) > 3



@pytest.mark.parametrize("chunk_pairs", [1, 7, 50, 444, 100000])
def test_overlapped_pipeline_equals_stage_by_stage(cuda_device, chunk_pairs):
    """pattern_separation_device (frame chunks on alternating streams, audio pyramid and the RESUMABLE boundary chain
    on a third) and pattern_separation_host (chunked upload from host memory) must return exactly the boundaries of
    the stage-by-stage calls and of the oracle, whatever the chunking -- including chunks shorter than one 30 s
    window, where most resume launches find nothing they are allowed to take yet."""
    from hippomm_b200 import synth
    from hippomm_b200.segmentation import (audio_energy_device, frame_pair_scores_device, pattern_separation_device,
                                           pattern_separation_host, segment_boundaries_device)

    nsec, sr = 400, 16000
    frames, _ = synth.frame_stream(31, nsec, 64, 64, min_scene=5, max_scene=40)
    pcm = synth.audio_stream_int16(32, nsec * sr)
    times = np.arange(nsec, dtype=np.float64) * 1.25 + 3.0        # frame_times[0] != 0, non-integer spacing
    fd = torch.from_numpy(frames).to(cuda_device)
    pd = torch.from_numpy(pcm.reshape(-1, 1)).to(cuda_device)
    ft = torch.from_numpy(times).to(cuda_device)
    ssim, _ = frame_pair_scores_device(fd, range_mode=0)
    pyr = audio_energy_device(pd)
    for thr in ((30.0, 10.0, 0.95, -40.0), (10.0, 5.0, 0.95, -40.0)):
        b0, c0 = segment_boundaries_device(ssim, ft, pd, pyr, sr, *thr, 256)
        b1, c1, s1 = pattern_separation_device(fd, ft, pd, sr, *thr, 256, chunk_pairs=chunk_pairs)
        b2, c2, s2 = pattern_separation_host(torch.from_numpy(frames), torch.from_numpy(times), torch.from_numpy(pcm.reshape(-1, 1)),
                                             sr, *thr, 256, chunk_frames=max(2, min(chunk_pairs, 300)))
        torch.cuda.synchronize()
        n = int(c0.item())
        assert n > 5 and int(c1.item()) == n and int(c2.item()) == n
        assert torch.equal(b0[:n], b1[:n]) and torch.equal(b0[:n], b2[:n])
        assert torch.equal(s1, ssim) and torch.equal(s2, ssim)
    want = O.segment_boundaries(O.adjacent_ssim(frames), list(times), pcm.astype(np.float64).reshape(-1, 1) / 32768.0, sr)
    b1, c1, _ = pattern_separation_device(fd, ft, pd, sr, 30.0, 10.0, 0.95, -40.0, 256, chunk_pairs=chunk_pairs)
    n = int(c1.item())
    assert [tuple(x) for x in b1[:n].cpu().numpy().tolist()] == [tuple(w) for w in want]
    # audio-only and video-only streams through the same entry
    ba, ca, _ = pattern_separation_device(None, None, pd, sr, 30.0, 10.0, 0.95, -40.0, 256)
    wa = O.segment_boundaries(None, None, pcm.astype(np.float64).reshape(-1, 1) / 32768.0, sr)
    assert [tuple(x) for x in ba[: int(ca.item())].cpu().numpy().tolist()] == [tuple(w) for w in wa]
    bv, cv, _ = pattern_separation_device(fd, ft, None, None, 30.0, 10.0, 0.95, -40.0, 256, chunk_pairs=chunk_pairs)
    wv = O.segment_boundaries(O.adjacent_ssim(frames), list(times), None, None)
    assert [tuple(x) for x in bv[: int(cv.item())].cpu().numpy().tolist()] == [tuple(w) for w in wv]


@pytest.mark.parametrize("env", [{"HIPPO_FOLLOW_US": "1"}, {"HIPPO_FOLLOW_US": "30"}, {"HIPPO_PATTERN_FOLLOW": "0"},
                                 {"HIPPO_PATTERN_L2PIN": "0"}])
def test_follow_mode_chain_suspends_and_completes(cuda_device, env, monkeypatch):
    """The boundary chain that FOLLOWS the SSIM kernels pair by pair (segment.cu follow mode) gives up when a pair does
    not arrive within its limit -- a serialising profiler, a launch-blocking debug run -- and the final resumable pass
    finishes the stream: with a 1 us / 30 us limit most calls take that path at some segment, and the boundaries, the
    SSIM values and the count must still be those of the stage-by-stage calls.  HIPPO_PATTERN_FOLLOW=0 is the
    chunk-by-chunk chain (also what streams longer than the staged 6,000 frames use); HIPPO_PATTERN_L2PIN=0 runs the
    follower without the persisting-L2 window over the block sums."""
    from hippomm_b200 import synth
    from hippomm_b200.segmentation import (audio_energy_device, frame_pair_scores_device, pattern_separation_device,
                                           segment_boundaries_device)

    nsec, sr = 900, 8000
    frames, _ = synth.frame_stream(41, nsec, 96, 80, min_scene=5, max_scene=40)
    pcm = synth.audio_stream_int16(42, nsec * sr)
    fd = torch.from_numpy(frames).to(cuda_device)
    pd = torch.from_numpy(pcm.reshape(-1, 1)).to(cuda_device)
    ft = torch.arange(nsec, dtype=torch.float64, device=cuda_device)
    ssim, _ = frame_pair_scores_device(fd, range_mode=0)
    pyr = audio_energy_device(pd)
    b0, c0 = segment_boundaries_device(ssim, ft, pd, pyr, sr, 30.0, 10.0, 0.95, -40.0, 256)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for chunk_pairs in (0, 100):
        for _ in range(3):
            b1, c1, s1 = pattern_separation_device(fd, ft, pd, sr, 30.0, 10.0, 0.95, -40.0, 256, chunk_pairs=chunk_pairs)
            torch.cuda.synchronize()
            n = int(c0.item())
            assert n > 20 and int(c1.item()) == n
            assert torch.equal(b0[:n], b1[:n]) and torch.equal(s1, ssim)


def test_stream_longer_than_the_staged_frames(cuda_device):
    """More than 6,000 frames: the chain cannot stage the stream in shared memory, so the pipeline falls back to the
    chunk-by-chunk resumable chain reading global memory; results as stage by stage."""
    from hippomm_b200.segmentation import (frame_pair_scores_device, pattern_separation_device,
                                           segment_boundaries_device)

    g = torch.Generator(device="cpu").manual_seed(7)
    nf = 6500
    scene = torch.randint(0, 255, (nf // 25 + 1, 16, 16, 3), generator=g, dtype=torch.uint8)
    frames = scene.repeat_interleave(25, dim=0)[:nf].clone()
    frames[:, 0, 0, 0] = torch.arange(nf, dtype=torch.int64).remainder(7).to(torch.uint8)
    fd = frames.to(cuda_device)
    ft = torch.arange(nf, dtype=torch.float64, device=cuda_device)
    ssim, _ = frame_pair_scores_device(fd, range_mode=0)
    b0, c0 = segment_boundaries_device(ssim, ft, None, None, None, 30.0, 10.0, 0.95, -40.0, 1024)
    b1, c1, s1 = pattern_separation_device(fd, ft, None, None, 30.0, 10.0, 0.95, -40.0, 1024)
    torch.cuda.synchronize()
    n = int(c0.item())
    assert n > 200 and int(c1.item()) == n and torch.equal(b0[:n], b1[:n]) and torch.equal(s1, ssim)


def test_a_silent_span_does_not_wait_for_video_but_a_loud_one_does(cuda_device):
    """Follow mode releases the video threads of a span as soon as a window of it is KNOWN to be silent (the audio boundary
    overwrites the video boundary, hm:1061-1077) and waits for the pairs otherwise: audio that decides every span, audio
    with no silence at all (every span is decided by the frames), and no audio -- each equal to the stage-by-stage
    calls and to the oracle."""
    from hippomm_b200 import synth
    from hippomm_b200.segmentation import (audio_energy_device, frame_pair_scores_device, pattern_separation_device,
                                           segment_boundaries_device)

    nsec, sr = 600, 8000
    frames, _ = synth.frame_stream(51, nsec, 64, 64, min_scene=4, max_scene=45)
    fd = torch.from_numpy(frames).to(cuda_device)
    ft = torch.arange(nsec, dtype=torch.float64, device=cuda_device)
    ssim, _ = frame_pair_scores_device(fd, range_mode=0)
    rng = np.random.default_rng(52)
    loud = (rng.standard_normal(nsec * sr) * 3000).astype(np.int16)                    # -20 dBFS everywhere
    often = loud.copy()
    for t0 in range(7, nsec, 13):                                                      # a second of silence every 13 s
        often[t0 * sr:(t0 + 1) * sr] = 0
    for pcm in (often, loud, None):
        pd = torch.from_numpy(pcm.reshape(-1, 1)).to(cuda_device) if pcm is not None else None
        pyr = audio_energy_device(pd) if pd is not None else None
        b0, c0 = segment_boundaries_device(ssim, ft, pd, pyr, sr, 30.0, 10.0, 0.95, -40.0, 256)
        for _ in range(3):
            b1, c1, s1 = pattern_separation_device(fd, ft, pd, sr, 30.0, 10.0, 0.95, -40.0, 256)
            torch.cuda.synchronize()
            n = int(c0.item())
            assert n > 15 and int(c1.item()) == n and torch.equal(b0[:n], b1[:n]) and torch.equal(s1, ssim)
        x = pcm.astype(np.float64).reshape(-1, 1) / 32768.0 if pcm is not None else None
        want = O.segment_boundaries(O.adjacent_ssim(frames), [float(i) for i in range(nsec)], x, sr if pcm is not None else None)
        assert [tuple(v) for v in b1[:n].cpu().numpy().tolist()] == [tuple(w) for w in want]
