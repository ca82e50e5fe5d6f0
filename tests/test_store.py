"""ThetaEvent store hooks (hippomm_b200/store.py): the binary sidecar and the wiring of install(event_store=True).
CPU only: the sidecar and the load / save wrappers touch no GPU code."""
import dataclasses
import os
import pathlib
import sys
import tempfile
import types
from typing import Any, Dict, List, Optional

import numpy as np
import pytest

import cases
from oracle import reference_shim


@dataclasses.dataclass
class _ThetaEvent:
    """Field for field the reference's dataclass (hm:95-108)."""
    features: Dict[str, np.ndarray]
    feature_times: Optional[Dict[str, np.ndarray]]
    frames: List[str]
    frame_times: List[float]
    frame_captions: List[str]
    audio_times: List[float]
    audio_transcription: List[Dict[str, Any]]
    holistic_audio_transcription: List[Dict[str, Any]]
    summary: str
    start_time: float
    end_time: float


def test_sidecar_round_trip_keeps_every_field_bit_for_bit():
    from hippomm_b200 import store

    rng = np.random.default_rng(1)
    ev = _ThetaEvent(
        features={"vision": rng.standard_normal((7, 1024)), "audio": rng.standard_normal((3, 1024)).astype(np.float32),
                  "empty": np.zeros((0, 1024))},
        feature_times={"vision_times": np.arange(7) * 0.5, "audio_times": np.array([1.5, 4.5, 7.5])},
        frames=["/a/b.jpg", "/a/c.jpg"], frame_times=[0.0, 2.5], frame_captions=["a cat", "ünïcode ☃"],
        audio_times=[1.0], audio_transcription=[{"text": "hi", "start": 0.0}], holistic_audio_transcription=[],
        summary="two frames", start_time=0.0, end_time=2.5)
    with tempfile.TemporaryDirectory() as td:
        p = store.sidecar_path(os.path.join(td, "vid_0.json"))
        assert p.name == "vid_0.json.hbin"
        store.write_sidecar(ev, p)
        back = store.read_sidecar(p, _ThetaEvent)
    for k, v in ev.features.items():
        assert back.features[k].dtype == v.dtype and back.features[k].shape == v.shape
        assert back.features[k].tobytes() == v.tobytes()
    for k, v in ev.feature_times.items():
        assert np.array_equal(back.feature_times[k], v) and back.feature_times[k].dtype == v.dtype
    for f in ("frames", "frame_times", "frame_captions", "audio_times", "audio_transcription",
              "holistic_audio_transcription", "summary", "start_time", "end_time"):
        assert getattr(back, f) == getattr(ev, f), f
    with tempfile.TemporaryDirectory() as td:
        bad = os.path.join(td, "x.hbin")
        open(bad, "wb").write(b"not a sidecar at all")
        with pytest.raises(ValueError):
            store.read_sidecar(bad, _ThetaEvent)


@pytest.mark.skipif(not reference_shim.available(), reason="needs the reference checkout (build container only)")
def test_hooked_load_and_save_return_what_the_reference_returns():
    """install(event_store=True) around the REAL reference classes: a store written by the reference's own
    save_theta_event is loaded (a) by the reference's loader, (b) by the hooked loader parsing JSON and writing
    sidecars, (c) by the hooked loader reading the sidecars only -- all three must agree field for field (float64
    rows, hm:391), and an edited JSON file must win over a stale sidecar."""
    import json

    from hippomm_b200 import store

    ref = reference_shim.load()
    _, hm, _ = ref.modules
    events, _ = cases.recall_events()
    saved: dict = {}
    with tempfile.TemporaryDirectory() as td:
        m = object.__new__(hm.HippocampalMemory)
        m.events_dir = pathlib.Path(td) / "events"
        m.events_dir.mkdir()
        m.event_index, m.event_index_file, m.long_term_store = {}, pathlib.Path(td) / "event_index.json", []
        for e in events[:4]:
            te = hm.ThetaEvent(features={**e.features, **e.feature_times}, feature_times=dict(e.feature_times),
                               frames=e.frames, frame_times=e.frame_times, frame_captions=["c"], audio_times=[],
                               audio_transcription=[], holistic_audio_transcription=[], summary="s",
                               start_time=float(e.start_time), end_time=float(e.end_time))
            m.save_theta_event(te, "vid")
        ids = list(m.event_index)
        plain = [m.load_theta_event(i) for i in ids]
        m.long_term_store = []
        store.install_event_store(hm, saved)
        try:
            first = [m.load_theta_event(i) for i in ids]
            sides = [store.sidecar_path(m.event_index[i]["file_path"]) for i in ids]
            assert all(s.exists() for s in sides)
            assert len(m.long_term_store) == len(ids)                      # the side effect of hm:441 is kept
            second = [m.load_theta_event(i) for i in ids]
            for a, b, c in zip(plain, first, second):
                assert type(c) is hm.ThetaEvent
                assert set(a.features) == set(b.features) == set(c.features)
                for k in a.features:
                    assert c.features[k].dtype == np.float64
                    assert a.features[k].tobytes() == b.features[k].tobytes() == c.features[k].tobytes()
                for k in a.feature_times:
                    assert np.array_equal(a.feature_times[k], c.feature_times[k])
                for f in ("frames", "frame_times", "frame_captions", "summary", "start_time", "end_time"):
                    assert getattr(a, f) == getattr(c, f), f
            # a JSON file newer than its sidecar is the truth
            f0 = pathlib.Path(m.event_index[ids[0]]["file_path"])
            doc = json.loads(f0.read_text())
            doc["summary"] = "edited"
            f0.write_text(json.dumps(doc))
            os.utime(f0, ns=(sides[0].stat().st_mtime_ns + 10_000_000, sides[0].stat().st_mtime_ns + 10_000_000))
            assert m.load_theta_event(ids[0]).summary == "edited"
            assert m.load_theta_event("no such event") is None
            # the hooked save writes the sidecar at once, holding what a load of the JSON yields
            e = events[5]
            te = hm.ThetaEvent(features={**e.features, **e.feature_times}, feature_times={}, frames=e.frames,
                               frame_times=e.frame_times, frame_captions=[], audio_times=[], audio_transcription=[],
                               holistic_audio_transcription=[], summary="", start_time=float(e.start_time),
                               end_time=float(e.end_time))
            n_before = len(m.long_term_store)
            m.save_theta_event(te, "vid2")
            assert len(m.long_term_store) == n_before                       # saving does not touch the store
            new_id = [i for i in m.event_index if i not in ids][0]
            side = store.sidecar_path(m.event_index[new_id]["file_path"])
            assert side.exists()
            got = store.read_sidecar(side, hm.ThetaEvent)
            assert got.features["vision"].dtype == np.float64
            assert np.array_equal(got.features["vision"], e.features["vision"].astype(np.float64))
            assert "vision_times" in got.feature_times
        finally:
            store.uninstall_event_store(saved)
        assert hm.HippocampalMemory.load_theta_event.__name__ == "load_theta_event"
        assert "store" not in saved
