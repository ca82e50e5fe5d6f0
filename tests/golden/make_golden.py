"""Generate tests/golden/*.npz by running the UNMODIFIED reference functions (imported from
/root/reference through oracle/reference_shim.py) on seeded synthetic inputs.

    python tests/golden/make_golden.py

Run in the build container only (the reference checkout does not exist on the GPU box).  Inputs
are regenerated from their seeds by tests/cases.py, so only the reference's OUTPUTS are stored.
SSIM-dependent outputs come from the reference's own call sites running on the restated
structural_similarity (scikit-image is absent: those entries are "vs restated oracle").
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import cases  # noqa: E402  (tests/cases.py)
from oracle import reference_shim  # noqa: E402


def main() -> None:
    ref = reference_shim.load()
    mem = ref.make_memory()
    out = {}

    # ---- feature search, config 1: 2,000 video-like rows, 64 planted queries, k = 5 ----
    bank, queries = cases.search_config1()
    idx = np.empty((len(queries), 5), dtype=np.int64)
    sim = np.empty((len(queries), 5), dtype=np.float32)
    for i, q in enumerate(queries):
        idx[i], sim[i] = ref.top_k_cosine_similarity(q, bank, 5)
    out["search_c1_idx"], out["search_c1_sim"] = idx, sim

    # ---- feature search, lattice bank (bit-exact scores expected), k = 10 ----
    lbank, lq, _ = cases.search_lattice_small()
    idx = np.empty((len(lq), 10), dtype=np.int64)
    sim = np.empty((len(lq), 10), dtype=np.float32)
    for i, q in enumerate(lq):
        idx[i], sim[i] = ref.top_k_cosine_similarity(q, lbank, 10)
    out["search_lat_idx"], out["search_lat_sim"] = idx, sim

    # ---- edge cases of vo:151-188 ----
    eb, eq = cases.search_edge()
    i1, s1 = ref.top_k_cosine_similarity(eq, eb, 4)
    out["search_edge_idx"], out["search_edge_sim"] = i1, s1
    i2, s2 = ref.top_k_cosine_similarity(eq, eb[:3], 8)        # k > N returns N
    out["search_kgtn_idx"], out["search_kgtn_sim"] = i2, s2
    i3, s3 = ref.top_k_cosine_similarity(eq, eb[5], 3)         # 1-D b
    out["search_1d_idx"], out["search_1d_sim"] = i3, s3
    out["cosine_pair"] = np.array([ref.cosine_similarity(eq, eb[5])], dtype=np.float64)

    # ---- consolidation (hm:944-967) ----
    for name, feats in cases.consolidation_cases().items():
        for g in (0.9, 0.95):
            out[f"cons_{name}_{int(g * 100)}"] = mem._select_key_frames(feats, None, g).astype(np.int64)
    out["cons_default_gamma"] = mem._select_key_frames(cases.consolidation_cases()["c1"], None).astype(np.int64)

    # ---- audio level (hm:993-1000) ----
    pcm = cases.audio_case()
    x = pcm.astype(np.float64) / 32768.0
    wins = cases.audio_windows(len(pcm))
    out["audio_levels"] = np.array([mem._compute_audio_level(x[s:s + n].reshape(-1, 1), 16000) for s, n in wins])
    out["audio_level_stereo"] = np.array([mem._compute_audio_level(np.stack([x[:8000], x[8000:16000]], axis=1), 16000)])
    out["audio_level_zero"] = np.array([mem._compute_audio_level(np.zeros(100), 16000)], dtype=np.float64)

    # ---- frame difference (bp:32-71) ----
    frames, _ = cases.frame_case()
    pairs = cases.frame_diff_pairs()
    out["frame_diff"] = np.array([ref.compute_frame_difference(frames[a], frames[b]) for a, b in pairs])
    const = cases.constant_frames()
    out["frame_diff_const"] = np.array([ref.compute_frame_difference(const[a], const[b]) for a, b in ((0, 0), (0, 1), (1, 2))])

    # ---- frame similarity + segmentation through the reference's own code path (paths on disk) ----
    import cv2

    with tempfile.TemporaryDirectory() as td:
        paths = []
        for i, f in enumerate(frames):
            p = os.path.join(td, f"f{i:05d}.png")          # PNG: lossless, so cv2.imread returns the same pixels
            cv2.imwrite(p, f)
            paths.append(p)
        sims = np.array([mem._compute_frame_similarity(paths[i + 1], paths[i]) for i in range(len(paths) - 1)])
        out["frame_ssim_adjacent"] = sims
        cpaths = []
        for i, f in enumerate(const):
            p = os.path.join(td, f"c{i}.png")
            cv2.imwrite(p, f)
            cpaths.append(p)
        with np.errstate(all="ignore"):
            out["frame_ssim_const"] = np.array([mem._compute_frame_similarity(cpaths[a], cpaths[b])
                                                for a, b in ((0, 1), (0, 2), (2, 0))])
        times = cases.frame_times(len(frames))
        audio = x[: int(times[-1] * 16000) + 16000].reshape(-1, 1)
        for tag, kw in cases.segmentation_variants().items():
            vf = paths if kw["video"] else None
            ft = times if kw["video"] else None
            au = audio if kw["audio"] else None
            sr = 16000 if kw["audio"] else None
            m2 = ref.make_memory(**kw["thresholds"])
            segs = m2._segment_sequence(vf, ft, au, sr)
            out[f"seg_{tag}"] = np.array([[s.start_time, s.end_time] for s in segs], dtype=np.float64).reshape(-1, 2)
            out[f"seg_{tag}_nframes"] = np.array([len(s.frames) if s.frames is not None else -1 for s in segs])
            out[f"seg_{tag}_nsamples"] = np.array([len(s.audio_data) if s.audio_data is not None else -1 for s in segs])

    # ---- detailed recall: the reference's own loops over a synthetic long-term store (hm:3127-3383) ----
    import torch

    events, queries = cases.recall_events()
    rs = ref.make_recall_system(events)
    for name, (modality, q) in queries.items():
        fn = rs._find_relevant_video_segments if modality == "vision" else rs._find_relevant_audio_segments
        segs = fn(torch.from_numpy(q))
        out[f"recall_{name}_bounds"] = np.array([[s.start_time, s.end_time] for s in segs], dtype=np.float64).reshape(-1, 2)
        if modality == "vision":
            out[f"recall_{name}_nframes"] = np.array([len(s.frames) for s in segs], dtype=np.int64)
            out[f"recall_{name}_frame_times"] = np.array([t for s in segs for t in s.frame_times], dtype=np.float64)

    # ---- the same loops over a store that went through the reference's OWN save_theta_event / load_theta_event
    # (JSON text and back: float64 rows, hm:334-335, hm:391) ----
    import pathlib
    import types

    vo_, hm_, bp_ = ref.modules
    with tempfile.TemporaryDirectory() as td:
        m = object.__new__(hm_.HippocampalMemory)
        m.events_dir = pathlib.Path(td) / "events"
        m.events_dir.mkdir()
        m.event_index, m.event_index_file, m.long_term_store = {}, pathlib.Path(td) / "event_index.json", []
        for e in events:
            # the consolidation code keeps the time tables INSIDE `features` under '<modality>_times' keys, which is
            # where to_dict (hm:116-120) looks for them; the reload then files them under feature_times (hm:375-381)
            te = hm_.ThetaEvent(features={**e.features, **e.feature_times}, feature_times=dict(e.feature_times), frames=e.frames,
                                frame_times=e.frame_times, frame_captions=[], audio_times=[], audio_transcription=[],
                                holistic_audio_transcription=[], summary="", start_time=float(e.start_time),
                                end_time=float(e.end_time))
            m.save_theta_event(te, "vid")
        for event_id in list(m.event_index):
            m.load_theta_event(event_id)
        assert len(m.long_term_store) == len(events)
        for e, te in zip(events, m.long_term_store):      # the JSON round trip is exact: float32 values held in float64
            for k_, v in e.features.items():
                assert te.features[k_].dtype == np.float64 and np.array_equal(te.features[k_], v.astype(np.float64))
        rs2 = object.__new__(hm_.QARecallSystem)
        rs2.memory = types.SimpleNamespace(long_term_store=m.long_term_store)
        rs2._current_question = ""
        for name, (modality, q) in queries.items():
            fn = rs2._find_relevant_video_segments if modality == "vision" else rs2._find_relevant_audio_segments
            segs = fn(torch.from_numpy(q))
            out[f"recall_json_{name}_bounds"] = np.array([[s.start_time, s.end_time] for s in segs], dtype=np.float64).reshape(-1, 2)
            if modality == "vision":
                out[f"recall_json_{name}_nframes"] = np.array([len(s.frames) for s in segs], dtype=np.int64)
                out[f"recall_json_{name}_frame_times"] = np.array([t for s in segs for t in s.frame_times], dtype=np.float64)

    # ---- key-frame pre-filter: the reference's own extract_frames_from_video (bp:116-260) on an MJPG AVI ----
    import hashlib
    import pathlib
    import re

    pf = cases.prefilter_frames()
    with tempfile.TemporaryDirectory() as td:
        avi = os.path.join(td, "v.avi")
        assert cases.write_mjpg(pf, avi, cases.PREFILTER_PARAMS["video_fps"])
        decoded = cases.read_video(avi)
        assert len(decoded) == len(pf)
        paths, times, duration = ref.extract_frames_from_video(
            avi, pathlib.Path(td) / "store", "vid", config={}, max_diff_threshold=cases.PREFILTER_PARAMS["max_diff_threshold"],
            check_interval=cases.PREFILTER_PARAMS["check_interval"])
        out["prefilter_frame_numbers"] = np.array([int(re.search(r"frame_(\d+)\.jpg", p).group(1)) for p in paths], dtype=np.int64)
        out["prefilter_frame_times"] = np.array(times, dtype=np.float64)
        # fingerprint of the DECODED frames: the committed decisions only apply if the test box decodes the same pixels
        out["prefilter_decoded_sha1"] = np.frombuffer(hashlib.sha1(decoded.tobytes()).digest(), dtype=np.uint8)

    path = os.path.join(HERE, "reference_outputs.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")
    for k, v in out.items():
        print(f"  {k:28s} {v.dtype} {v.shape}")


if __name__ == "__main__":
    main()
