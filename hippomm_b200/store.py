"""The reference's ThetaEvent store, served from the GPU.

The reference keeps every embedding as decimal text inside the event's JSON file (`ThetaEvent.to_dict` hm:110-133
-> `json.dump` hm:334-335), parses it back into float64 arrays on load (hm:355-449, hm:391), and detailed recall
then loops over the events calling `top_k_cosine_similarity(query, event_features, k=5)` once per event
(hm:3143-3153 vision, hm:3294-3304 audio).  `install(event_store=True)` changes three things and nothing else:

  * `HippocampalMemory.save_theta_event` / `load_theta_event` (hm:320-353, hm:355-449) also write / prefer a binary
    SIDECAR next to the JSON file (`<event>.json.hbin`: the arrays as raw bytes in their own dtype, the remaining
    fields as one small JSON header).  A load that finds a sidecar at least as new as the JSON file never parses the
    decimal text; the ThetaEvent it returns is field for field what the reference's loader returns (float64 rows).
  * `QARecallSystem._find_relevant_video_segments` / `_find_relevant_audio_segments` (hm:3127-3279, hm:3281-3383)
    search ONE cross-event device bank per modality (`EventBank`, built once per state of `long_term_store`, rows kept
    in their own precision for the exact re-scoring) with one `hippo_topk_segmented` pass instead of a Python loop of
    per-event calls, then `hippo_recall_windows` does the tail.  When an event would take the reference's LLM branch
    (best similarity below 0.4 AND captions / a holistic transcription present, hm:3156 / hm:3307) the call is handed
    to the reference's own method unchanged -- that branch is prompt text and a network client, not arithmetic.
  * the return types stay the reference's own (`ThetaEvent`, `List[SequenceSegment]`).
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import Any, Dict, Optional

import numpy as np

SIDECAR_SUFFIX = ".hbin"
_MAGIC = b"HIPPOEV1"
_ALIGN = 64
_META_FIELDS = ("frames", "frame_times", "frame_captions", "audio_times", "audio_transcription",
                "holistic_audio_transcription", "summary", "start_time", "end_time")


# ------------------------------------------------------------------ sidecar ----
def sidecar_path(event_file) -> Path:
    p = Path(event_file)
    return p.with_name(p.name + SIDECAR_SUFFIX)


def _jsonable(x):
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, (np.floating, np.integer)):
        return x.item()
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    if isinstance(x, dict):
        return {k: _jsonable(v) for k, v in x.items()}
    return x


def write_sidecar(event, path) -> None:
    """ThetaEvent (or any object with its fields) -> binary sidecar.  Arrays keep their dtype and shape bit for bit."""
    arrays: Dict[str, np.ndarray] = {}
    for m, a in (event.features or {}).items():
        arrays["f:" + m] = np.ascontiguousarray(np.asarray(a))
    for m, a in (event.feature_times or {}).items():
        arrays["t:" + m] = np.ascontiguousarray(np.asarray(a))
    directory, pos = {}, 0
    for name, a in arrays.items():
        if a.dtype == object:
            raise TypeError(f"{name}: object arrays cannot be stored")
        directory[name] = dict(dtype=a.dtype.str, shape=list(a.shape), offset=pos, nbytes=int(a.nbytes))
        pos += (a.nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
    meta = {f: _jsonable(getattr(event, f, None)) for f in _META_FIELDS}
    head = json.dumps(dict(version=1, arrays=directory, meta=meta)).encode()
    head_len = (len(_MAGIC) + 8 + len(head) + _ALIGN - 1) // _ALIGN * _ALIGN
    path = Path(path)
    tmp = path.with_name(path.name + ".tmp")
    with open(tmp, "wb") as f:
        f.write(_MAGIC)
        f.write(np.uint64(len(head)).tobytes())
        f.write(head)
        f.write(b"\0" * (head_len - len(_MAGIC) - 8 - len(head)))
        for a in arrays.values():
            f.write(a.tobytes())
            f.write(b"\0" * ((-a.nbytes) % _ALIGN))
    os.replace(tmp, path)


def read_sidecar(path, event_cls):
    """Binary sidecar -> `event_cls(**fields)` (the reference's ThetaEvent dataclass, hm:95-108)."""
    with open(path, "rb") as f:
        if f.read(len(_MAGIC)) != _MAGIC:
            raise ValueError(f"{path}: not a hippomm_b200 event sidecar")
        hl = int(np.frombuffer(f.read(8), dtype=np.uint64)[0])
        doc = json.loads(f.read(hl).decode())
        if doc.get("version") != 1:
            raise ValueError(f"{path}: unsupported sidecar version {doc.get('version')}")
        base = (len(_MAGIC) + 8 + hl + _ALIGN - 1) // _ALIGN * _ALIGN
        features: Dict[str, np.ndarray] = {}
        times: Dict[str, np.ndarray] = {}
        for name, s in doc["arrays"].items():
            f.seek(base + s["offset"])
            a = np.frombuffer(f.read(s["nbytes"]), dtype=np.dtype(s["dtype"])).reshape(s["shape"]).copy()
            (features if name.startswith("f:") else times)[name[2:]] = a
    meta = doc["meta"]
    return event_cls(features=features, feature_times=times, **{k: meta.get(k) for k in _META_FIELDS})


# ------------------------------------------------------------- device banks ----
def _store_signature(store) -> tuple:
    return tuple(id(e) for e in store)


def banks_for(memory, modality: str):
    """The cross-event device bank of `memory.long_term_store` for one modality, rebuilt when the store's events change."""
    from .events import EventBank

    cache = memory.__dict__.setdefault("_hippo_event_banks", {})
    sig = _store_signature(memory.long_term_store)
    ent = cache.get(modality)
    if ent is None or ent[0] != sig:
        try:
            bank = EventBank.from_events(memory.long_term_store, modality, keep_rows=True)
        except ValueError:          # no event carries this modality
            bank = None
        ent = cache[modality] = (sig, bank)
    return ent[1]


class _Delegate(Exception):
    """Raised when the call has to take the reference's own path (LLM branch)."""


def find_segments(recall_system, query_features, modality: str, segment_cls=None):
    """Arithmetic of hm:3127-3279 / hm:3281-3383 for the whole store in one pass; raises _Delegate when an event
    would take the LLM branch."""
    import torch

    from .events import find_relevant_segments

    q = query_features.detach().cpu().numpy() if isinstance(query_features, torch.Tensor) else np.asarray(query_features)
    if q.ndim > 1:
        q = q.flatten()                                                       # hm:3133-3134
    if modality == "vision" and q.shape[0] != 1024:                          # hm:3135-3137
        return []
    memory = recall_system.memory
    store = memory.long_term_store
    bank = banks_for(memory, modality)
    if bank is None:
        return []
    searched = bank.search(q.astype(np.float32), 5, exact=True)
    mx = searched[2].cpu().numpy()
    for j in np.nonzero(mx < 0.4)[0]:                                         # hm:3156 / hm:3307
        ev = store[int(bank.event_index[j])]
        text = getattr(ev, "frame_captions", None) if modality == "vision" else \
            getattr(ev, "holistic_audio_transcription", None)
        if text:
            raise _Delegate()
    return find_relevant_segments(q, store, bank=bank, modality=modality, searched=searched, segment_cls=segment_cls)


# ------------------------------------------------------------------ install ----
def install_event_store(hm_module, saved: Dict[str, Any]) -> None:
    """Wrap the four reference methods (see module docstring).  `saved` receives what uninstall needs."""
    H = hm_module.HippocampalMemory
    R = hm_module.QARecallSystem
    ThetaEvent = hm_module.ThetaEvent
    Segment = hm_module.SequenceSegment
    orig_load, orig_save = H.load_theta_event, H.save_theta_event
    orig_video, orig_audio = R._find_relevant_video_segments, R._find_relevant_audio_segments
    saved["store"] = (hm_module, orig_load, orig_save, orig_video, orig_audio)

    def load_theta_event(self, event_id: str):
        if event_id not in self.event_index:                                  # hm:357-358
            return None
        event_file = Path(self.event_index[event_id]["file_path"])
        side = sidecar_path(event_file)
        if event_file.exists() and side.exists() and side.stat().st_mtime_ns >= event_file.stat().st_mtime_ns:
            try:
                event = read_sidecar(side, ThetaEvent)
                self.long_term_store.append(event)                            # hm:441
                return event
            except Exception:
                pass                                                          # unreadable sidecar: the JSON is the truth
        event = orig_load(self, event_id)
        if event is not None:
            try:
                write_sidecar(event, side)
            except Exception:
                pass
        return event

    def save_theta_event(self, event, video_id: str) -> None:
        orig_save(self, event, video_id)
        try:
            event_id = f"{video_id}_{int(event.start_time * 1000)}"           # hm:323
            # the sidecar holds what a LOAD of this JSON file yields (float64 rows, hm:391), not the in-memory event
            loaded = orig_load_quiet(self, event_id)
            if loaded is not None:
                write_sidecar(loaded, sidecar_path(self.event_index[event_id]["file_path"]))
        except Exception:
            pass

    def orig_load_quiet(self, event_id):
        """The reference's loader without its side effect on long_term_store (hm:441)."""
        n = len(self.long_term_store)
        event = orig_load(self, event_id)
        del self.long_term_store[n:]
        return event

    def _find_relevant_video_segments(self, query_features, optional_search_query=None):
        try:
            return find_segments(self, query_features, "vision", Segment)
        except _Delegate:
            return orig_video(self, query_features, optional_search_query)

    def _find_relevant_audio_segments(self, query_features):
        try:
            return find_segments(self, query_features, "audio", Segment)
        except _Delegate:
            return orig_audio(self, query_features)

    H.load_theta_event, H.save_theta_event = load_theta_event, save_theta_event
    R._find_relevant_video_segments, R._find_relevant_audio_segments = _find_relevant_video_segments, _find_relevant_audio_segments


def uninstall_event_store(saved: Dict[str, Any]) -> None:
    ent = saved.pop("store", None)
    if ent is None:
        return
    hm_module, orig_load, orig_save, orig_video, orig_audio = ent
    H, R = hm_module.HippocampalMemory, hm_module.QARecallSystem
    H.load_theta_event, H.save_theta_event = orig_load, orig_save
    R._find_relevant_video_segments, R._find_relevant_audio_segments = orig_video, orig_audio
