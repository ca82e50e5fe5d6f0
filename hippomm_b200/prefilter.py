"""Greedy frame filters around the frame-pair kernel.

  * `select_saved_frames`  -- the key-frame pre-filter decisions of `extract_frames_from_video`
    (bp:179-228): every `check_interval`-th decoded frame that is at least 1 s after the last saved
    frame is compared with the LAST SAVED frame (`compute_frame_difference`, bp:194), the difference
    is accumulated, and the frame is saved when either exceeds `max_diff_threshold` (bp:198-200).
  * `dedup_window_frames`  -- the frame de-duplication of the QA re-decode loops (hm:2226-2249 with
    0.3, hm:2789-2812 with 0.4): a frame is dropped when its SSIM to the last KEPT frame exceeds the
    threshold (`_compute_frame_similarity`, hm:2237).

Both chains are sequential in "the last kept frame".  The arithmetic (gray conversion, SSIM, MSE) runs
in `hippo_frame_pairs`; the host speculates: one launch scores the next `window` candidates against the
current anchor, the decision rule walks them in the reference's order, and the first kept frame becomes
the next anchor.  Decoding / resizing / writing frames stays with the caller (cv2), as in the reference.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from . import _cuda
from .segmentation import _load_frames, frame_pair_scores_device


def _frames_to_device(frames) -> torch.Tensor:
    dev = _cuda.require_device()
    if isinstance(frames, torch.Tensor):
        t = frames.to(dev)
        if t.dim() == 3:
            t = t.unsqueeze(-1)
        return t.contiguous()
    return _cuda.to_device(_load_frames(frames), dev)


def _pair_scores(fd: torch.Tensor, anchor: int, cands: List[int], range_mode: int, anchor_first: bool = False):
    """Scores of (candidate, anchor) pairs -- or (anchor, candidate) with anchor_first, which matters for
    range_mode 0 where the data range comes from the first frame (hm:990).  Only the frames of this round
    are handed to the kernel, so the gray conversion touches 1 + len(cands) frames."""
    sel = torch.tensor([anchor] + list(cands), dtype=torch.int64, device=fd.device)
    sub = fd.index_select(0, sel)
    c = torch.arange(1, len(cands) + 1, dtype=torch.int32)
    z = torch.zeros((len(cands),), dtype=torch.int32)
    a, b = (z, c) if anchor_first else (c, z)
    ssim, mse = frame_pair_scores_device(sub, a, b, range_mode=range_mode)
    both = torch.stack([ssim, mse]).cpu().numpy()
    return both[0], both[1]


def select_saved_frames(frames, video_fps: float, max_diff_threshold: float = 0.3, check_interval: int = 30,
                        window: int = 8) -> Tuple[List[int], List[float]]:
    """Frame numbers and times `extract_frames_from_video` would save (bp:179-228), for decoded frames
    [n, h, w, 3] uint8 (host array, list of arrays, or device tensor).  `min_diff_threshold` of the
    reference is accepted there and never used (bp:121), so it has no counterpart here."""
    fd = _frames_to_device(frames)
    n = fd.shape[0]
    if n == 0:
        return [], []
    small = fd.shape[1] < 7 or fd.shape[2] < 7              # SSIM raises -> the MSE fallback (bp:64-71)
    saved, times = [0], [0 / video_fps]                      # bp:187-188: the first frame is always saved
    last_save_time = 0 / video_fps
    cumulative = 0.0
    # candidates: multiples of check_interval that pass the 1-second gate (bp:190-192); the gate depends on
    # the last save, so the list is rebuilt after every save
    pos = 1
    while pos < n:
        cands = []
        fc = pos
        while fc < n and len(cands) < window:
            if fc % check_interval == 0 and fc / video_fps - last_save_time >= 1.0:
                cands.append(fc)
            fc += 1
        if not cands:
            break
        ssim, mse = _pair_scores(fd, saved[-1], cands, range_mode=1)
        hit = None
        for j, c in enumerate(cands):
            if not small and np.isfinite(ssim[j]):           # bp:59-63
                diff = 1.0 - float(ssim[j])
            else:
                diff = min(1.0, float(mse[j]))               # bp:67-71
            cumulative += diff                               # bp:195
            if diff > max_diff_threshold or cumulative > max_diff_threshold:   # bp:198-200
                hit = c
                break
        if hit is None:
            pos = cands[-1] + 1
            continue
        saved.append(hit)
        times.append(hit / video_fps)
        last_save_time = hit / video_fps
        cumulative = 0.0                                     # bp:218
        pos = hit + 1
    return saved, times


def dedup_window_frames(frames, threshold: float = 0.3, window: int = 4) -> List[int]:
    """Indices of the frames one QA re-decode window keeps (hm:2226-2249 / hm:2789-2812): the first
    frame, then every frame whose SSIM to the last kept frame is NOT above `threshold`."""
    fd = _frames_to_device(frames)
    n = fd.shape[0]
    if n == 0:
        return []
    if fd.shape[1] < 7 or fd.shape[2] < 7:
        raise ValueError("win_size exceeds image extent.")
    kept = [0]
    pos = 1
    while pos < n:
        cands = list(range(pos, min(n, pos + window)))
        ssim, _ = _pair_scores(fd, kept[-1], cands, range_mode=0, anchor_first=True)
        # _compute_frame_similarity(prev, cur): data_range comes from the FIRST argument = the kept frame (hm:990)
        hit = None
        for j, c in enumerate(cands):
            if not (float(ssim[j]) > threshold):             # hm:2238: `if similarity > 0.3: drop`; NaN keeps
                hit = c
                break
        if hit is None:
            pos = cands[-1] + 1
            continue
        kept.append(hit)
        pos = hit + 1
    return kept
