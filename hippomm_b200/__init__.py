"""hippomm_b200 — B200-native (sm_100a) implementation of HippoMM's data-parallel memory hot path.

Temporal pattern separation, memory consolidation and detailed-recall feature search, behind the
reference's own Python call signatures.  `install()` rebinds the reference's symbols so that an
unmodified `hippomm` checkout runs these paths on the GPU; the functions can also be called directly:

    from hippomm_b200 import top_k_cosine_similarity, select_key_frames, segment_sequence, MemoryBank

All arithmetic runs in libhippo_b200.so (hand-written CUDA behind a C ABI, include/hippo_b200.h).
There is no CPU fallback: without the library or without an sm_100 GPU the calls raise.
"""
from __future__ import annotations

from . import _lib
from .bank import MemoryBank, search_rows
from .consolidation import RecheckOverflow, checked_key_frames, select_key_frames, select_key_frames_device
from .events import EventBank, find_relevant_segments
from .prefilter import dedup_window_frames, select_saved_frames
from .segmentation import (SequenceSegment, compute_audio_level, compute_frame_difference,
                           compute_frame_similarity, segment_sequence)
from .vector_ops import cosine_similarity, invalidate_bank_cache, set_bank_cache, top_k_cosine_similarity

__version__ = "0.1.0"

__all__ = [
    "MemoryBank", "SequenceSegment", "top_k_cosine_similarity", "cosine_similarity", "select_key_frames",
    "select_key_frames_device", "segment_sequence", "compute_frame_similarity", "compute_audio_level",
    "compute_frame_difference", "EventBank", "find_relevant_segments", "select_saved_frames", "dedup_window_frames",
    "search_rows", "checked_key_frames", "RecheckOverflow", "set_bank_cache", "invalidate_bank_cache",
    "install", "uninstall", "library_path",
]

_saved = {}


def library_path() -> str:
    return str(_lib.LIB_PATH)


# --- methods bound onto HippocampalMemory (same signatures as hm:944, hm:980, hm:993, hm:1002) ----------
def _m_select_key_frames(self, features, times, similarity_threshold: float = 0.9):
    return select_key_frames(features, times, similarity_threshold)


def _m_compute_frame_similarity(self, frame1_path, frame2_path):
    return compute_frame_similarity(frame1_path, frame2_path)


def _m_compute_audio_level(self, audio_data, sample_rate):
    return compute_audio_level(audio_data, sample_rate)


def _m_segment_sequence(self, video_frames=None, frame_times=None, audio_data=None, audio_sample_rate=None):
    segs = segment_sequence(
        video_frames, frame_times, audio_data, audio_sample_rate,
        max_segment_duration=self.max_segment_duration, min_segment_duration=self.min_segment_duration,
        frame_similarity_threshold=self.frame_similarity_threshold,
        audio_silence_threshold=self.audio_silence_threshold)
    seg_cls = _saved.get("SequenceSegment")
    if seg_cls is None:
        return segs
    return [seg_cls(start_time=s.start_time, end_time=s.end_time, frames=s.frames, audio_data=s.audio_data,
                    frame_times=s.frame_times) for s in segs]


def install(cache_banks: bool = False, event_store: bool = False) -> None:
    """Rebind the reference's hot-path symbols to the GPU implementations:

      hippomm.utils.vector_ops.{top_k_cosine_similarity, cosine_similarity}
      hippomm.core.hippocampal_memory.{top_k_cosine_similarity, cosine_similarity}   (from-import copies, hm:28)
      HippocampalMemory.{_select_key_frames, _segment_sequence, _compute_frame_similarity, _compute_audio_level}
      hippomm.core.batch_process.compute_frame_difference

    `hippomm` must be importable.  Modules that cannot be imported (missing third-party packages) are
    skipped; at least vector_ops must succeed.  cache_banks=True keeps device banks of recently searched
    feature arrays that are READ-ONLY (`arr.flags.writeable = False`; keyed by object identity) so repeated
    queries skip the upload; writeable arrays are searched straight from the upload on every call.

    event_store=True additionally hooks the ThetaEvent store (hippomm_b200/store.py): `save_theta_event` /
    `load_theta_event` (hm:320-449) write / prefer a binary sidecar next to each event's JSON file (same return type,
    no decimal-text parsing on reload), and `QARecallSystem._find_relevant_{video,audio}_segments` (hm:3127-3383) search
    ONE cross-event device bank per modality with `hippo_topk_segmented` + `hippo_recall_windows` instead of looping
    over the events; events that would take the reference's LLM branch hand the call back to the reference's method.
    """
    import importlib

    _lib.load()
    vo = importlib.import_module("hippomm.utils.vector_ops")
    _saved.setdefault("vo", (vo, vo.top_k_cosine_similarity, vo.cosine_similarity))
    vo.top_k_cosine_similarity = top_k_cosine_similarity
    vo.cosine_similarity = cosine_similarity
    set_bank_cache(16 if cache_banks else 0)
    try:
        hm = importlib.import_module("hippomm.core.hippocampal_memory")
    except Exception:
        hm = None
    if hm is not None:
        H = hm.HippocampalMemory
        _saved.setdefault("hm", (hm, hm.top_k_cosine_similarity, hm.cosine_similarity, H._select_key_frames,
                                 H._segment_sequence, H._compute_frame_similarity, H._compute_audio_level))
        _saved["SequenceSegment"] = hm.SequenceSegment
        hm.top_k_cosine_similarity = top_k_cosine_similarity
        hm.cosine_similarity = cosine_similarity
        H._select_key_frames = _m_select_key_frames
        H._segment_sequence = _m_segment_sequence
        H._compute_frame_similarity = _m_compute_frame_similarity
        H._compute_audio_level = _m_compute_audio_level
        if event_store and "store" not in _saved:
            from . import store

            store.install_event_store(hm, _saved)
    elif event_store:
        raise ImportError("install(event_store=True) needs hippomm.core.hippocampal_memory to be importable")
    try:
        bp = importlib.import_module("hippomm.core.batch_process")
    except Exception:
        bp = None
    if bp is not None:
        _saved.setdefault("bp", (bp, bp.compute_frame_difference))
        bp.compute_frame_difference = compute_frame_difference


def uninstall() -> None:
    """Undo install()."""
    if "store" in _saved:
        from . import store

        store.uninstall_event_store(_saved)
    if "vo" in _saved:
        vo, f1, f2 = _saved.pop("vo")
        vo.top_k_cosine_similarity, vo.cosine_similarity = f1, f2
    if "hm" in _saved:
        hm, f1, f2, m1, m2, m3, m4 = _saved.pop("hm")
        hm.top_k_cosine_similarity, hm.cosine_similarity = f1, f2
        H = hm.HippocampalMemory
        H._select_key_frames, H._segment_sequence, H._compute_frame_similarity, H._compute_audio_level = m1, m2, m3, m4
    _saved.pop("SequenceSegment", None)
    if "bp" in _saved:
        bp, f = _saved.pop("bp")
        bp.compute_frame_difference = f
    set_bank_cache(0)
