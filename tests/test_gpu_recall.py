"""GPU parity: detailed recall over ALL events at once (hm:3127-3383) --
EventBank (segmented top-k, one pass), the window tail, the binary bank file -- through the C ABI."""
import numpy as np
import pytest
import torch

import cases
from oracle import hippo_oracle as O
from parity import check_topk

pytestmark = pytest.mark.gpu


def _bank(modality):
    from hippomm_b200.events import EventBank

    events, queries = cases.recall_events()
    return events, queries, EventBank.from_events(events, modality)


@pytest.mark.parametrize("modality", ["vision", "audio"])
def test_segmented_topk_matches_per_event_reference_calls(cuda_device, modality):
    """One hippo_topk_segmented pass == the reference's loop of top_k_cosine_similarity(q, event, 5) calls."""
    events, queries, eb = _bank(modality)
    for name, (m, q) in queries.items():
        if m != modality:
            continue
        idx, score, mx = eb.search(q, 5)
        torch.cuda.synchronize()
        idx, score, mx = idx.cpu().numpy(), score.cpu().numpy(), mx.cpu().numpy()
        for j, ei in enumerate(eb.event_index):
            feats = events[int(ei)].features[modality]
            ri, rs = O.top_k_cosine_similarity(q, feats, 5)
            kk = len(ri)                                           # events with fewer than 5 rows return fewer
            assert (idx[j, kk:] == -1).all()
            check_topk(idx[j, :kk], score[j, :kk], ri, rs, what=f"{name} event {ei}")
            assert abs(mx[j] - np.max(rs)) <= 1e-3


def test_scores_single_matches_reference_formula(cuda_device):
    events, queries, eb = _bank("vision")
    q = queries["v_mix"][1]
    s = eb.scores(q).cpu().numpy()
    rows = np.concatenate([e.features["vision"] for e in events])
    ref = rows @ q / (np.linalg.norm(rows, axis=1) * np.linalg.norm(q))      # vo:178-182
    assert s.shape == ref.shape and np.abs(s - ref).max() <= 1e-3


@pytest.mark.parametrize("name", ["v_in", "v_out", "v_mix", "a_in", "a_mix"])
def test_find_relevant_segments_matches_reference_outputs(cuda_device, name):
    """Against the committed outputs of the reference's own _find_relevant_{video,audio}_segments."""
    from hippomm_b200.events import find_relevant_segments

    events, queries = cases.recall_events()
    modality, q = queries[name]
    g = cases.golden()
    segs = find_relevant_segments(torch.from_numpy(q), events, modality=modality)
    bounds = np.array([[s.start_time, s.end_time] for s in segs], dtype=np.float64).reshape(-1, 2)
    assert np.array_equal(bounds, g[f"recall_{name}_bounds"])
    if modality == "vision":
        assert [len(s.frames) for s in segs] == g[f"recall_{name}_nframes"].tolist()
        assert [t for s in segs for t in s.frame_times] == g[f"recall_{name}_frame_times"].tolist()
        # frames are the event's own paths inside the window (hm:3267-3268)
        want = O.find_relevant_segments(q, events, "vision")
        assert [s.frames for s in segs] == [w["frames"] for w in want]
    else:
        assert all(s.audio_data is None and s.frames is None for s in segs)


def test_low_similarity_callback_entries_take_part_in_the_ranking(cuda_device):
    """hm:3156-3254: an event below 0.4 with captions goes to the LLM path, whose windows enter the final sort
    with similarity 0.6.  The callback stands in for the LLM."""
    from hippomm_b200.events import find_relevant_segments
    from hippomm_b200 import SequenceSegment

    events, queries = cases.recall_events()
    events = list(events)
    modality, q = queries["v_in"]
    base = find_relevant_segments(q, events, modality=modality)
    asked = []

    def llm(event, idx, sims):
        asked.append(event)
        assert len(idx) == len(sims) and np.max(sims) < 0.4
        return [(0.6, [SequenceSegment(start_time=1000.0, end_time=1002.0)])]

    import copy
    ev2 = [copy.copy(e) for e in events]
    ev2[0].frame_captions = ["a caption"]                          # event 0 is far from the query
    segs = find_relevant_segments(q, ev2, modality=modality, low_similarity=llm)
    assert asked == [ev2[0]]
    # every similarity-path hit of the planted event scores > 0.6, so the LLM window cannot displace them ...
    assert [(s.start_time, s.end_time) for s in segs] == [(s.start_time, s.end_time) for s in base]
    # ... but it outranks everything once it claims 2.0
    segs = find_relevant_segments(q, ev2, modality=modality,
                                  low_similarity=lambda e, i, s: [(2.0, [SequenceSegment(1000.0, 1002.0)])])
    assert (segs[0].start_time, segs[0].end_time) == (1000.0, 1002.0) and len(segs) == 5


def test_bank_file_round_trip(cuda_device, tmp_path):
    """save -> load reproduces the device bank bit for bit and the search results exactly."""
    from hippomm_b200.events import EventBank

    events, queries, eb = _bank("vision")
    path = str(tmp_path / "vision.hbank")
    eb.save(path)
    eb2 = EventBank.load(path)
    assert eb2.nev == eb.nev and np.array_equal(eb2.offsets, eb.offsets) and np.array_equal(eb2.times, eb.times)
    assert torch.equal(eb2.bank.rows.view(torch.int16), eb.bank.rows.view(torch.int16))
    assert torch.equal(eb2.bank.norm, eb.bank.norm)
    q = queries["v_mix"][1]
    a, b = eb.search(q, 5), eb2.search(q, 5)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    with open(path, "r+b") as f:
        f.write(b"XXXX")
    with pytest.raises(ValueError):
        EventBank.load(path)


def test_edge_cases(cuda_device):
    """Empty events, a zero-norm row (NaN first, vo:185), k larger than every event, disabled events."""
    from hippomm_b200.events import EventBank

    rng = np.random.default_rng(5)

    class E:
        pass

    evs = []
    for n in (4, 0, 9, 1):
        e = E()
        e.features = {"vision": rng.standard_normal((n, 128)).astype(np.float32)}
        e.frame_times = [float(i) for i in range(n)]
        e.frames = [f"f{i}" for i in range(n)]
        evs.append(e)
    evs[2].features["vision"][5] = 0.0
    eb = EventBank.from_events(evs, "vision")
    q = rng.standard_normal(128).astype(np.float32)
    idx, score, mx = eb.search(q, 6)
    idx, score, mx = idx.cpu().numpy(), score.cpu().numpy(), mx.cpu().numpy()
    assert (idx[1] == -1).all() and np.isnan(mx[1])                # empty event: nothing, np.max of nothing is undefined
    assert idx[2, 0] == 5 and np.isnan(score[2, 0]) and np.isnan(mx[2])
    assert sorted(idx[0, :4].tolist()) == [0, 1, 2, 3] and (idx[0, 4:] == -1).all()
    ev, rows, sims, wins = eb.recall_windows(torch.from_numpy(idx).cuda(), torch.from_numpy(score).cuda(), top=5,
                                             enabled=np.array([1, 1, 0, 1], dtype=np.uint8))
    assert 2 not in ev.tolist() and len(ev) == 5
    assert np.all(np.diff(sims) <= 0)
    assert np.array_equal(wins[:, 0], np.maximum(0.0, rows.astype(np.float64) - 1.0))


def _json_store_events():
    """What the reference's save_theta_event -> load_theta_event round trip yields for cases.recall_events()
    (checked against the real reference in tests/golden/make_golden.py): float64 rows holding the float32 values."""
    events, queries = cases.recall_events()
    out = []
    for e in events:
        feats = {k: v.astype(np.float64) for k, v in e.features.items()}
        out.append(cases._Event(feats, dict(e.feature_times), e.frames, e.frame_times))
    return out, queries


@pytest.mark.parametrize("name", ["v_in", "v_out", "v_mix", "a_in", "a_mix"])
def test_event_store_hook_matches_reference_on_a_json_loaded_store(cuda_device, name):
    """install(event_store=True)'s search path (store.find_segments: ONE cross-event bank, hippo_topk_segmented,
    exact re-scoring from the float64 rows, hippo_recall_windows) against the committed outputs of the reference's
    own _find_relevant_{video,audio}_segments run on a store that went through its JSON save / load."""
    import types

    from hippomm_b200 import store

    events, queries = _json_store_events()
    modality, q = queries[name]
    g = cases.golden()
    rs = types.SimpleNamespace(memory=types.SimpleNamespace(long_term_store=events))
    segs = store.find_segments(rs, torch.from_numpy(q), modality)
    bounds = np.array([[s.start_time, s.end_time] for s in segs], dtype=np.float64).reshape(-1, 2)
    assert np.array_equal(bounds, g[f"recall_json_{name}_bounds"])
    if modality == "vision":
        assert np.array_equal(np.array([len(s.frames) for s in segs]), g[f"recall_json_{name}_nframes"])
        assert np.array_equal(np.array([t for s in segs for t in s.frame_times]), g[f"recall_json_{name}_frame_times"])
    bank = rs.memory.__dict__["_hippo_event_banks"][modality][1]
    assert bank.bank.src.dtype == torch.float64                    # the rows are kept in the store's own precision
    store.find_segments(rs, torch.from_numpy(q), modality)
    assert rs.memory.__dict__["_hippo_event_banks"][modality][1] is bank      # built once per state of the store
    rs.memory.long_term_store = events[:-1]
    store.find_segments(rs, torch.from_numpy(q), modality)
    assert rs.memory.__dict__["_hippo_event_banks"][modality][1] is not bank  # ... and rebuilt when it changes


def test_event_store_hook_wiring_and_llm_delegation(cuda_device):
    """install_event_store on a stand-in module with the reference's class layout: the wrapped recall methods answer
    from the GPU bank, hand the call to the original method exactly when an event would take the LLM branch
    (best similarity < 0.4 and captions present, hm:3156), and uninstall restores the originals."""
    import types

    from hippomm_b200 import SequenceSegment, store

    events, queries = _json_store_events()
    calls = []

    class HippocampalMemory:
        def load_theta_event(self, event_id):
            return None

        def save_theta_event(self, event, video_id):
            return None

    class QARecallSystem:
        def _find_relevant_video_segments(self, query_features, optional_search_query=None):
            calls.append("video")
            return ["from the reference"]

        def _find_relevant_audio_segments(self, query_features):
            calls.append("audio")
            return ["from the reference"]

    fake = types.SimpleNamespace(HippocampalMemory=HippocampalMemory, QARecallSystem=QARecallSystem,
                                 ThetaEvent=cases._Event, SequenceSegment=SequenceSegment)
    saved = {}
    store.install_event_store(fake, saved)
    try:
        rs = QARecallSystem()
        rs.memory = types.SimpleNamespace(long_term_store=events)
        g = cases.golden()
        segs = rs._find_relevant_video_segments(torch.from_numpy(queries["v_mix"][1]))
        assert not calls and all(isinstance(s, SequenceSegment) for s in segs)
        assert np.array_equal(np.array([[s.start_time, s.end_time] for s in segs]), g["recall_json_v_mix_bounds"])
        segs = rs._find_relevant_audio_segments(torch.from_numpy(queries["a_in"][1]))
        assert not calls
        assert np.array_equal(np.array([[s.start_time, s.end_time] for s in segs]), g["recall_json_a_in_bounds"])
        # a query unrelated to every event scores below 0.4 everywhere; captions on one event -> the LLM branch
        rng = np.random.default_rng(3)
        far = torch.from_numpy(rng.standard_normal(1024).astype(np.float32))
        assert rs._find_relevant_video_segments(far) != ["from the reference"] and not calls     # no captions: GPU path
        events[2].frame_captions = ["a caption"]
        assert rs._find_relevant_video_segments(far) == ["from the reference"] and calls == ["video"]
        events[2].frame_captions = []
        assert rs._find_relevant_video_segments(torch.zeros(7)) == []                          # hm:3135-3137
    finally:
        store.uninstall_event_store(saved)
    assert QARecallSystem()._find_relevant_audio_segments(None) == ["from the reference"]


def test_bank_from_many_small_arrays_equals_bank_from_one_array(cuda_device):
    """`MemoryBank.fill_from_parts` (the per-event feature arrays of a store through shared pinned chunks): arrays smaller
    and larger than a chunk, empty ones, fp32 and fp64 mixed -- the same bf16 rows, norms and kept originals as one
    `from_rows` call on the concatenation."""
    from hippomm_b200 import MemoryBank

    rng = np.random.default_rng(5)
    d = 192
    sizes = [3, 0, 700, 1, 2500, 64, 0, 999, 5]
    for dts in ([np.float32] * len(sizes), [np.float32, np.float64] * 4 + [np.float32]):
        parts = [rng.standard_normal((n, d)).astype(dt) for n, dt in zip(sizes, dts)]
        wide = any(dt == np.float64 for dt in dts)
        whole = np.concatenate([p.astype(np.float64 if wide else np.float32) for p in parts])
        ref = MemoryBank.from_rows(whole, keep_rows=True)
        bank = MemoryBank(len(whole), d, keep_rows=True, rows_dtype=torch.float64 if wide else torch.float32)
        bank.fill_from_parts(parts, chunk_rows=1024)
        torch.cuda.synchronize()
        assert torch.equal(bank.rows[: bank.n], ref.rows[: ref.n]) and torch.equal(bank.norm[: bank.n], ref.norm[: ref.n])
        assert torch.equal(bank.src, ref.src)
