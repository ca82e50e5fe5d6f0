"""Greedy frame filters around the frame-pair kernel.

  * `select_saved_frames`  -- the key-frame pre-filter decisions of `extract_frames_from_video`
    (bp:179-228): every `check_interval`-th decoded frame that is at least 1 s after the last saved
    frame is compared with the LAST SAVED frame (`compute_frame_difference`, bp:194), the difference
    is accumulated, and the frame is saved when either exceeds `max_diff_threshold` (bp:198-200).
  * `dedup_window_frames`  -- the frame de-duplication of the QA re-decode loops (hm:2226-2249 with
    0.3, hm:2789-2812 with 0.4): a frame is dropped when its SSIM to the last KEPT frame exceeds the
    threshold (`_compute_frame_similarity`, hm:2237).

Both chains are sequential in "the last kept frame".  The arithmetic (gray conversion, SSIM, MSE) runs
in `hippo_frame_pairs`; the host speculates.  The anchor of a decision is always an earlier candidate, so ONE
launch per block of candidates scores every (candidate, earlier candidate) pair up to `band` candidates apart
(all pairs for a short window) and the decision rule then walks a table on the host, in the reference's order; a
pair outside the band (a long static stretch) falls back to a launch that scores the next `window` candidates
against the current anchor.  A pair's scores depend on its two frames only, so the decisions do not depend on
the speculation.  Decoding / resizing / writing frames stays with the caller (cv2), as in the reference.
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


class _PairTable:
    """Scores of (node i, node j) pairs, j < i <= j + band, over a sorted list of frame numbers (`nodes`), computed
    block by block on first use: one gather + one `hippo_frame_pairs` call + one readback per block of nodes."""

    def __init__(self, fd: torch.Tensor, nodes: List[int], range_mode: int, anchor_first: bool, band: int,
                 block: int = 2048):
        self.fd, self.nodes, self.range_mode, self.anchor_first = fd, nodes, range_mode, anchor_first
        self.band = max(1, min(int(band), max(len(nodes) - 1, 1)))
        self.block = max(1, min(int(block), 65535 // self.band))          # <= 65,535 pairs and frames per call
        self.pos = {f: i for i, f in enumerate(nodes)}
        self.blocks = {}

    def _build(self, b: int):
        i0, i1 = b * self.block, min((b + 1) * self.block, len(self.nodes))
        j0 = max(0, i0 - self.band)
        sel = torch.tensor(self.nodes[j0:i1], dtype=torch.int64, device=self.fd.device)
        sub = self.fd.index_select(0, sel)
        ci = np.repeat(np.arange(i0, i1), self.band)
        aj = ci - 1 - np.tile(np.arange(self.band), i1 - i0)
        keep = aj >= 0
        c = torch.from_numpy((ci[keep] - j0).astype(np.int32))
        a = torch.from_numpy((aj[keep] - j0).astype(np.int32))
        first, second = (a, c) if self.anchor_first else (c, a)
        ssim = np.full(((i1 - i0) * self.band,), np.nan)
        mse = np.full(((i1 - i0) * self.band,), np.nan)
        if c.numel():
            s_d, m_d = frame_pair_scores_device(sub, first, second, range_mode=self.range_mode)
            both = torch.stack([s_d, m_d]).cpu().numpy()
            ssim[keep], mse[keep] = both[0], both[1]
        self.blocks[b] = (ssim.reshape(i1 - i0, self.band), mse.reshape(i1 - i0, self.band))

    def lookup(self, anchor: int, cands: List[int]):
        """(ssim, mse) arrays of the (candidate, anchor) pairs, or None when one of them lies outside the band."""
        ai = self.pos.get(anchor)
        if ai is None:
            return None
        ssim, mse = np.empty(len(cands)), np.empty(len(cands))
        for n, cf in enumerate(cands):
            ci = self.pos.get(cf)
            if ci is None or not (0 < ci - ai <= self.band):
                return None
            b = ci // self.block
            if b not in self.blocks:
                self._build(b)
            S, M = self.blocks[b]
            ssim[n], mse[n] = S[ci - b * self.block, ci - ai - 1], M[ci - b * self.block, ci - ai - 1]
        return ssim, mse


def select_saved_frames(frames, video_fps: float, max_diff_threshold: float = 0.3, check_interval: int = 30,
                        window: int = 8, band: int = 24) -> Tuple[List[int], List[float]]:
    """Frame numbers and times `extract_frames_from_video` would save (bp:179-228), for decoded frames
    [n, h, w, 3] uint8 (host array, list of arrays, or device tensor).  `min_diff_threshold` of the
    reference is accepted there and never used (bp:121), so it has no counterpart here.
    `band` = how many candidates back the speculative table reaches (0: no table), `window` = candidates per
    fallback launch."""
    fd = _frames_to_device(frames)
    n = fd.shape[0]
    if n == 0:
        return [], []
    small = fd.shape[1] < 7 or fd.shape[2] < 7              # SSIM raises -> the MSE fallback (bp:64-71)
    saved, times = [0], [0 / video_fps]                      # bp:187-188: the first frame is always saved
    last_save_time = 0 / video_fps
    cumulative = 0.0
    # every anchor and every candidate is frame 0 or a multiple of check_interval
    nodes = sorted(set([0] + list(range(check_interval, n, check_interval)))) if check_interval >= 1 else [0]
    table = _PairTable(fd, nodes, 1, False, band) if band > 0 and len(nodes) > 1 else None
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
        got = table.lookup(saved[-1], cands) if table is not None else None
        ssim, mse = got if got is not None else _pair_scores(fd, saved[-1], cands, range_mode=1)
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


def dedup_window_frames(frames, threshold: float = 0.3, window: int = 4, band: int = 128) -> List[int]:
    """Indices of the frames one QA re-decode window keeps (hm:2226-2249 / hm:2789-2812): the first
    frame, then every frame whose SSIM to the last kept frame is NOT above `threshold`.
    `band`: the speculative table holds every pair of frames at most that far apart (a re-decode window is a few
    dozen frames: all pairs, one launch); 0 = fallback launches of `window` candidates only."""
    fd = _frames_to_device(frames)
    n = fd.shape[0]
    if n == 0:
        return []
    if fd.shape[1] < 7 or fd.shape[2] < 7:
        raise ValueError("win_size exceeds image extent.")
    table = _PairTable(fd, list(range(n)), 0, True, band) if band > 0 and n > 1 else None
    kept = [0]
    pos = 1
    while pos < n:
        cands = list(range(pos, min(n, pos + window)))
        got = table.lookup(kept[-1], cands) if table is not None else None
        ssim = got[0] if got is not None else _pair_scores(fd, kept[-1], cands, range_mode=0, anchor_first=True)[0]
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
