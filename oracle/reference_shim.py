"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference functions for oracle validation.

Only usable where the reference checkout exists (/root/reference in the build container; it
is absent on the GPU box, where the committed fixtures under tests/golden/ stand in).
Nothing under hippomm_b200/ imports this module.

The reference's hot-path modules fail to import only because of absent NON-arithmetic
third-party packages (librosa, soundfile, decord, token_count, faster_whisper,
qwen_vl_utils, imagebind, skimage).  Empty stub modules are registered for those; the stubs
carry no arithmetic except `skimage.metrics.structural_similarity`, which is supplied by the
restatement in oracle/hippo_oracle.py (scikit-image is not installed and is unpinned in the
reference's requirements.txt:30 -> SSIM parity is "vs restated oracle", i.e. UNPINNED).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("HIPPO_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "librosa", "soundfile", "decord", "token_count", "faster_whisper", "qwen_vl_utils",
    "imagebind", "imagebind.data", "imagebind.models", "imagebind.models.imagebind_model",
    "skimage", "skimage.metrics", "openai", "tiktoken",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "hippomm"))


def _install_stubs() -> None:
    from . import hippo_oracle

    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            __import__(name)
            continue
        except Exception:
            pass
        m = types.ModuleType(name)
        m.__path__ = []  # behave like a package
        sys.modules[name] = m
    sk = sys.modules["skimage.metrics"]
    if not hasattr(sk, "structural_similarity"):
        sk.structural_similarity = hippo_oracle.structural_similarity
    tc = sys.modules["token_count"]
    if not hasattr(tc, "TokenCount"):
        tc.TokenCount = type("TokenCount", (), {"__init__": lambda self, *a, **k: None})
    fw = sys.modules["faster_whisper"]
    if not hasattr(fw, "WhisperModel"):
        fw.WhisperModel = type("WhisperModel", (), {})
    oa = sys.modules["openai"]
    if not hasattr(oa, "OpenAI"):
        oa.OpenAI = type("OpenAI", (), {"__init__": lambda self, *a, **k: None})
    im = sys.modules["imagebind.models.imagebind_model"]
    if not hasattr(im, "ModalityType"):
        im.ModalityType = type("ModalityType", (), {"VISION": "vision", "AUDIO": "audio", "TEXT": "text"})
        im.imagebind_huge = lambda *a, **k: None
    sys.modules["imagebind.models"].imagebind_model = im
    dc = sys.modules["decord"]
    for attr in ("VideoReader", "cpu"):
        if not hasattr(dc, attr):
            setattr(dc, attr, lambda *a, **k: None)
    qv = sys.modules["qwen_vl_utils"]
    if not hasattr(qv, "process_vision_info"):
        qv.process_vision_info = lambda *a, **k: None


_loaded = {}


def load():
    """Returns a namespace with the reference's own functions:
    top_k_cosine_similarity, cosine_similarity, HippocampalMemory (class), compute_frame_difference,
    make_memory(**thresholds) -> instance with only the four threshold attributes set."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    vo = importlib.import_module("hippomm.utils.vector_ops")
    hm = importlib.import_module("hippomm.core.hippocampal_memory")
    bp = importlib.import_module("hippomm.core.batch_process")

    def make_memory(max_segment_duration=30.0, min_segment_duration=10.0, frame_similarity_threshold=0.95,
                    audio_silence_threshold=-40):
        m = object.__new__(hm.HippocampalMemory)  # __init__ loads three foundation models; not needed
        m.max_segment_duration = max_segment_duration
        m.min_segment_duration = min_segment_duration
        m.frame_similarity_threshold = frame_similarity_threshold
        m.audio_silence_threshold = audio_silence_threshold
        return m

    def make_recall_system(events):
        """QARecallSystem (hm:1615) with only the long-term store set: enough for the similarity path of
        _find_relevant_video_segments / _find_relevant_audio_segments (hm:3127-3383)."""
        r = object.__new__(hm.QARecallSystem)
        r.memory = types.SimpleNamespace(long_term_store=list(events))
        r._current_question = ""
        return r

    _loaded.update(
        make_recall_system=make_recall_system,
        top_k_cosine_similarity=vo.top_k_cosine_similarity,
        cosine_similarity=vo.cosine_similarity,
        HippocampalMemory=hm.HippocampalMemory,
        SequenceSegment=hm.SequenceSegment,
        compute_frame_difference=bp.compute_frame_difference,
        extract_frames_from_video=bp.extract_frames_from_video,
        make_memory=make_memory,
        modules=(vo, hm, bp),
    )
    return types.SimpleNamespace(**_loaded)
