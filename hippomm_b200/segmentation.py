"""Temporal pattern separation: drop-ins for HippocampalMemory._segment_sequence (hm:1002-1114),
_compute_frame_similarity (hm:980-991), _compute_audio_level (hm:993-1000) and
batch_process.compute_frame_difference (bp:32-71).

Host code only decodes files and slices Python lists; every arithmetic step (gray conversion,
SSIM / MSE, RMS levels, the boundary state machine) runs in the CUDA library.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _cuda, _lib

_PCM_ENUM = {torch.int16: _lib.HIPPO_I16, torch.float32: _lib.HIPPO_F32, torch.float64: _lib.HIPPO_F64}


@dataclass
class SequenceSegment:
    """Same fields as the reference dataclass (hm:35-42)."""
    start_time: float
    end_time: float
    frames: Optional[List[str]] = None
    audio_data: Optional[np.ndarray] = None
    frame_times: Optional[List[float]] = None


# ------------------------------------------------------------------ frames ----
def _load_frames(video_frames: Sequence) -> np.ndarray:
    """Paths (decoded with cv2 like hm:982-983) or arrays -> one uint8 array [n, h, w, ch]."""
    if isinstance(video_frames, np.ndarray) and video_frames.ndim in (3, 4) and video_frames.dtype == np.uint8:
        arr = video_frames if video_frames.ndim == 4 else video_frames[..., None]
        return np.ascontiguousarray(arr)
    imgs = []
    for f in video_frames:
        if isinstance(f, np.ndarray):
            img = f
        else:
            import cv2  # decode only; all arithmetic happens on the GPU
            img = cv2.imread(str(f))
            if img is None:
                raise ValueError(f"could not read frame {f!r}")
        if img.dtype != np.uint8:
            raise ValueError("frames must be uint8")
        if img.ndim == 2:
            img = img[..., None]
        imgs.append(img)
    shape = imgs[0].shape
    for img in imgs:
        if img.shape != shape:
            # skimage raises for a pair of different shapes (hm:990); a stream mixes none
            raise ValueError("Input images must have the same dimensions.")
    return np.ascontiguousarray(np.stack(imgs))


def frame_pair_scores_device(frames: torch.Tensor, pair_a: Optional[torch.Tensor] = None,
                             pair_b: Optional[torch.Tensor] = None, range_mode: int = 0, out=None):
    """frames uint8 [n, h, w, ch] on the device -> (ssim fp64 [np], mse fp64 [np]) device tensors.
    Without explicit pairs, pair p is (frame p+1, frame p), the order hm:1052-1056 scans them.
    `out` = (ssim, mse) contiguous fp64 tensors of at least np elements to write into."""
    lib = _lib.load()
    dev = _cuda.require_device(frames.device)
    if frames.dtype != torch.uint8 or frames.dim() != 4 or not frames.is_contiguous():
        raise ValueError("frames must be a contiguous uint8 [n, h, w, ch] tensor")
    n, h, w, ch = frames.shape
    if pair_a is None:
        npairs = n - 1
        pa = pb = None
    else:
        pa = pair_a.to(dev, torch.int32).contiguous()
        pb = pair_b.to(dev, torch.int32).contiguous()
        npairs = pa.numel()
    if out is not None:
        ssim, mse = out
        for t in (ssim, mse):
            if t.dtype != torch.float64 or not t.is_contiguous() or t.numel() < npairs or t.device != dev:
                raise ValueError("out tensors must be contiguous fp64 device tensors of at least npairs elements")
    else:
        ssim = torch.empty((max(npairs, 1),), dtype=torch.float64, device=dev)
        mse = torch.empty((max(npairs, 1),), dtype=torch.float64, device=dev)
    if npairs > 0:
        with torch.cuda.device(dev):
            ws_bytes = lib.hippo_frame_pairs_workspace_bytes(n, h, w, npairs)
            ws = _cuda.workspace(ws_bytes, dev, "frames")
            _lib.check(lib.hippo_frame_pairs(
                frames.data_ptr(), n, h, w, ch, _cuda.ptr(pa), _cuda.ptr(pb), npairs, range_mode,
                ssim.data_ptr(), mse.data_ptr(), ws.data_ptr(), ws.numel(), _cuda.stream_ptr()))
    return ssim[:npairs], mse[:npairs]


def compute_frame_similarity(frame1, frame2) -> float:
    """Mean SSIM of two frames given as paths or BGR arrays (hm:980-991)."""
    frames = _load_frames([frame1, frame2])
    if frames.shape[1] < 7 or frames.shape[2] < 7:
        raise ValueError("win_size exceeds image extent. Either ensure that your images are at least 7x7; "
                         "or pass win_size explicitly in the function call, with an odd value less than or "
                         "equal to the smaller side of your images.")
    dev = _cuda.require_device()
    fd = _cuda.to_device(frames, dev)
    a = torch.tensor([0], dtype=torch.int32)
    b = torch.tensor([1], dtype=torch.int32)
    ssim, _ = frame_pair_scores_device(fd, a, b, range_mode=0)
    return float(ssim.item())


def compute_frame_difference(frame1: np.ndarray, frame2: np.ndarray) -> float:
    """1 - SSIM of the frames scaled to [0, 1]; MSE fallback when SSIM is not finite (bp:32-71)."""
    f1 = np.asarray(frame1)
    f2 = np.asarray(frame2)
    c1 = 1 if f1.ndim == 2 else f1.shape[2]
    c2 = 1 if f2.ndim == 2 else f2.shape[2]
    if c1 != c2:
        raise NotImplementedError("compute_frame_difference: frames with different channel counts")
    frames = _load_frames([f1, f2])
    dev = _cuda.require_device()
    fd = _cuda.to_device(frames, dev)
    a = torch.tensor([0], dtype=torch.int32)
    b = torch.tensor([1], dtype=torch.int32)
    ssim, mse = frame_pair_scores_device(fd, a, b, range_mode=1)
    vals = torch.stack([ssim[0], mse[0]]).cpu().numpy()
    score, mse_v = float(vals[0]), float(vals[1])
    if frames.shape[1] >= 7 and frames.shape[2] >= 7 and np.isfinite(score):   # bp:59-63
        return 1.0 - score
    return min(1.0, mse_v)                                                     # bp:67-71


# ------------------------------------------------------------------- audio ----
def _audio_to_device(audio_data, dev, int16_pcm: bool = False):
    """Host audio (n,) / (n, ch) -> device tensor.  float32 / float64 pass through; every other dtype is converted
    to float64 WITHOUT scaling, which is what the reference's arithmetic does with it (`mean(axis=1)` / `np.mean`
    promote integers to float64, hm:995-998; the int16 wrap-around of `np.square` on a 1-D integer array is not
    reproduced).  int16_pcm=True keeps int16 samples as PCM (x = k / 32768, what `sf.read` hands the reference for
    pcm_s16le, bp:331-335): half the upload, exact sums."""
    if isinstance(audio_data, torch.Tensor):
        t = audio_data.detach()
        if not (t.dtype in (torch.float32, torch.float64) or (int16_pcm and t.dtype == torch.int16)):
            t = t.to(torch.float64)
        t = t.to(dev, non_blocking=True)
    else:
        a = np.asarray(audio_data)
        if not (a.dtype in (np.float32, np.float64) or (int16_pcm and a.dtype == np.int16)):
            a = a.astype(np.float64)
        t = _cuda.to_device(a, dev)
    if t.dim() == 1:
        t = t.reshape(-1, 1)
    if t.dim() != 2:
        raise ValueError("audio_data must be (n,) or (n, channels)")
    return t.contiguous()


def audio_energy_device(pcm: torch.Tensor, out=None):
    """pcm [ns, nch] device tensor -> (e16 fp64 [ceil(ns/16)], e512 fp64 [ceil(ns/512)]).
    `out` = (e16, e512) contiguous fp64 tensors of exactly those sizes to write into."""
    lib = _lib.load()
    dev = _cuda.require_device(pcm.device)
    ns, nch = pcm.shape
    n16, n512 = max((ns + 15) // 16, 1), max((ns + 511) // 512, 1)
    if out is not None:
        e16, e512 = out
        if e16.dtype != torch.float64 or e512.dtype != torch.float64 or e16.numel() != n16 or e512.numel() != n512 \
                or not e16.is_contiguous() or not e512.is_contiguous():
            raise ValueError("out must be contiguous fp64 tensors of ceil(ns/16) and ceil(ns/512) elements")
    else:
        e16 = torch.empty((n16,), dtype=torch.float64, device=dev)
        e512 = torch.empty((n512,), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.hippo_audio_energy(pcm.data_ptr(), _PCM_ENUM[pcm.dtype], ns, nch, e16.data_ptr(),
                                          e512.data_ptr(), _cuda.stream_ptr()))
    return e16, e512


def audio_levels_device(pcm: torch.Tensor, win_start: torch.Tensor, win_len: torch.Tensor, pyramid=None):
    """RMS level in dB of windows [start, start+len) of a device pcm [ns, nch]; fp64 [nwin] device tensor."""
    lib = _lib.load()
    dev = _cuda.require_device(pcm.device)
    ns, nch = pcm.shape
    ws_ = win_start.to(dev, torch.int64).contiguous()
    wl_ = win_len.to(dev, torch.int64).contiguous()
    out = torch.empty((max(ws_.numel(), 1),), dtype=torch.float64, device=dev)
    e16, e512 = pyramid if pyramid is not None else (None, None)
    with torch.cuda.device(dev):
        _lib.check(lib.hippo_audio_levels(pcm.data_ptr(), _PCM_ENUM[pcm.dtype], ns, nch, _cuda.ptr(e16),
                                          _cuda.ptr(e512), ws_.data_ptr(), wl_.data_ptr(), ws_.numel(),
                                          out.data_ptr(), _cuda.stream_ptr()))
    return out[: ws_.numel()]


def compute_audio_level(audio_data, sample_rate=None, *, int16_pcm: bool = False) -> float:
    """RMS level of an audio segment in dB, -100 for digital silence (hm:993-1000). `sample_rate` is unused.
    Integer input is taken at face value like the reference does; int16_pcm=True reads int16 as k / 32768."""
    dev = _cuda.require_device()
    pcm = _audio_to_device(audio_data, dev, int16_pcm)
    n = pcm.shape[0]
    lv = audio_levels_device(pcm, torch.tensor([0], dtype=torch.int64), torch.tensor([n], dtype=torch.int64))
    return float(lv.item())


# ------------------------------------------------------------ segmentation ----
def _fill_stream_desc(desc, ssim, frame_times, pcm, pyramid, sample_rate, bounds, count, max_segments):
    desc.ssim = _cuda.ptr(ssim) if ssim is not None and ssim.numel() > 0 else None
    desc.frame_times = _cuda.ptr(frame_times)
    desc.nframes = 0 if frame_times is None else frame_times.numel()
    desc.pcm = _cuda.ptr(pcm)
    desc.e16 = None if pyramid is None else pyramid[0].data_ptr()
    desc.e512 = None if pyramid is None else pyramid[1].data_ptr()
    desc.ns = 0 if pcm is None else pcm.shape[0]
    desc.nch = 1 if pcm is None else pcm.shape[1]
    desc.pcm_dtype = _lib.HIPPO_F64 if pcm is None else _PCM_ENUM[pcm.dtype]
    desc.sample_rate = float(sample_rate) if sample_rate else 0.0
    desc.out_bounds = bounds.data_ptr()
    desc.out_count = count.data_ptr()
    desc.max_segments = max_segments


def segment_boundaries_batch_device(streams, max_segment_duration: float, min_segment_duration: float,
                                    frame_similarity_threshold: float, audio_silence_threshold: float,
                                    max_segments: int):
    """Several independent streams through ONE hippo_segment_boundaries launch (one CTA per stream: the
    sequential boundary chains of different streams run side by side).  `streams` is a list of
    (ssim, frame_times, pcm, pyramid, sample_rate) tuples of device tensors, entries None as for the
    single-stream call.  Returns (bounds fp64 [n, max_segments, 2], counts int32 [n])."""
    lib = _lib.load()
    n = len(streams)
    if n == 0:
        raise ValueError("no streams")
    ref = next(t for st in streams for t in (st[1], st[2]) if t is not None)
    dev = _cuda.require_device(ref.device)
    bounds = torch.empty((n, max_segments, 2), dtype=torch.float64, device=dev)
    counts = torch.zeros((n,), dtype=torch.int32, device=dev)
    descs = (_lib.StreamDesc * n)()
    for i, (ssim, ft, pcm, pyr, sr) in enumerate(streams):
        _fill_stream_desc(descs[i], ssim, ft, pcm, pyr, sr, bounds[i], counts[i:i + 1], max_segments)
    # descriptor table through pinned memory, asynchronously: the call never blocks the host (torch's pinned-memory
    # cache keeps the staging block alive until the copy has run)
    raw = np.frombuffer(ctypes.string_at(ctypes.addressof(descs), ctypes.sizeof(descs)), dtype=np.uint8)
    staged = torch.empty((raw.size,), dtype=torch.uint8, pin_memory=True)
    staged.numpy()[:] = raw
    desc_dev = staged.to(dev, non_blocking=True)
    with torch.cuda.device(dev):
        _lib.check(lib.hippo_segment_boundaries(
            desc_dev.data_ptr(), n, float(max_segment_duration), float(min_segment_duration),
            float(frame_similarity_threshold), float(audio_silence_threshold), _cuda.stream_ptr()))
    return bounds, counts


def pattern_separation_batch_device(streams, max_segment_duration: float, min_segment_duration: float,
                                    frame_similarity_threshold: float, audio_silence_threshold: float,
                                    max_segments: int, lanes: int = 2, mode: str = "stages"):
    """Temporal pattern separation of several independent streams resident on the device.

    `streams`: list of (frames uint8 [nf, h, w, ch] or None, frame_times fp64 [nf] or None, pcm [ns, nch] or None,
    sample_rate).
    mode "stages" (default, the faster one for a batch): the per-stream kernels (gray + SSIM, audio pyramid) of
      consecutive streams are issued on `lanes` alternating CUDA streams, so the HBM-bound gray conversion of one
      stream overlaps the issue-bound SSIM kernel of another; ONE boundary launch (one CTA per stream: the sequential
      chains of all streams side by side) follows on the caller's stream.
    mode "pipeline": every stream goes through `pattern_separation_device` (its own stages overlapped), consecutive
      streams issued from `lanes` alternating CUDA streams.
    Scratch is per CUDA stream (`_cuda.workspace`); results are the same as stream-by-stream calls.
    Returns (bounds fp64 [n, max_segments, 2], counts int32 [n])."""
    if not streams:
        raise ValueError("no streams")
    if mode not in ("stages", "pipeline"):
        raise ValueError(f"unknown mode {mode!r}")
    ref = next(t for st in streams for t in (st[0], st[2]) if t is not None)
    dev = _cuda.require_device(ref.device)
    n = len(streams)
    lanes = max(1, int(lanes))
    with torch.cuda.device(dev):
        main = torch.cuda.current_stream()
        if mode == "stages":
            side = _side_streams(dev, lanes)
            # SSIM values and audio pyramids of all streams (29.7 MB per stream-hour) live in ONE grow-only arena of the
            # caller's stream until the boundary launch has read them: fresh blocks per stream and call made the
            # caching allocator fall back to cudaMalloc whenever the previous batch was still in flight
            layout, total = [], 0
            for frames, ft, pcm, sr in streams:
                npairs = frames.shape[0] - 1 if (frames is not None and frames.shape[0] > 1) else 0
                ns = pcm.shape[0] if pcm is not None else 0
                n16, n512 = (max((ns + 15) // 16, 1), max((ns + 511) // 512, 1)) if pcm is not None else (0, 0)
                offs = []
                for cnt in (npairs, npairs, n16, n512):
                    offs.append((total, cnt))
                    total += (cnt + 31) // 32 * 32                      # fp64 elements, 256-byte aligned
                layout.append(offs)
            arena = _cuda.workspace(total * 8, dev, "batch_arena").view(torch.float64) if total else None
            for sd in side:
                sd.wait_stream(main)
            prepared = []
            for i, (frames, ft, pcm, sr) in enumerate(streams):
                sd = side[i % len(side)]
                (o_s, n_s), (o_m, n_m), (o_16, n_16), (o_512, n_512) = layout[i]
                with torch.cuda.stream(sd):
                    ssim = pyr = None
                    if n_s > 0:
                        ssim, _ = frame_pair_scores_device(frames, range_mode=0,
                                                           out=(arena[o_s:o_s + n_s], arena[o_m:o_m + n_m]))
                    if pcm is not None:
                        p2 = pcm.reshape(-1, 1) if pcm.dim() == 1 else pcm
                        pyr = audio_energy_device(p2, out=(arena[o_16:o_16 + n_16], arena[o_512:o_512 + n_512]))
                prepared.append((ssim, ft, pcm, pyr, sr))
            for sd in side:
                main.wait_stream(sd)
            return segment_boundaries_batch_device(prepared, max_segment_duration, min_segment_duration,
                                                   frame_similarity_threshold, audio_silence_threshold, max_segments)
        bounds = torch.empty((n, max_segments, 2), dtype=torch.float64, device=dev)
        counts = torch.zeros((n,), dtype=torch.int32, device=dev)
        issue = _issue_streams(dev, lanes)
        for sd in issue:
            sd.wait_stream(main)
        for i, (frames, ft, pcm, sr) in enumerate(streams):
            slot = i % lanes
            with torch.cuda.stream(issue[slot]):
                b, c, _ = pattern_separation_device(frames, ft, pcm, sr, max_segment_duration, min_segment_duration,
                                                    frame_similarity_threshold, audio_silence_threshold, max_segments,
                                                    slot=slot)
                bounds[i].copy_(b, non_blocking=True)
                counts[i:i + 1].copy_(c, non_blocking=True)
        for sd in issue:
            main.wait_stream(sd)
            bounds.record_stream(sd)
            counts.record_stream(sd)
    return bounds, counts


def pattern_separation_device(frames: Optional[torch.Tensor], frame_times: Optional[torch.Tensor],
                              pcm: Optional[torch.Tensor], sample_rate, max_segment_duration: float,
                              min_segment_duration: float, frame_similarity_threshold: float,
                              audio_silence_threshold: float, max_segments: int, chunk_pairs: int = 444,
                              slot: int = 0):
    """Temporal pattern separation of ONE stream resident on the device, its stages overlapped
    (`hippo_pattern_separation`: ONE C-ABI call issues every launch).

    The boundary state machine is launched first, in follow mode: one CTA on an SM of its own that polls each pair's
    SSIM as the SSIM warps deliver it, so the sequential chain ends a few microseconds after the last pair instead of
    0.23 ms later.  The audio pyramid runs on a second side stream, gray conversion + one SSIM launch on a third; a
    final resumable pass behind everything completes the chain if the follower gave up waiting (serialising
    profilers, launch-blocking debug runs).  Results are identical to the three stage-by-stage calls.
    `chunk_pairs` is ignored (the first version took the frames in chunks).  `slot` picks one of several independent
    sets of side streams / scratch, so that calls for different streams issued from different CUDA streams can
    overlap each other.
    Returns (bounds fp64 [max_segments, 2], count int32 [1], ssim fp64 [nf - 1] or None)."""
    lib = _lib.load()
    ref = frames if frames is not None else pcm
    dev = _cuda.require_device(ref.device)
    has_video = frames is not None and frame_times is not None and frames.shape[0] > 0
    if has_video and (frames.dtype != torch.uint8 or frames.dim() != 4 or not frames.is_contiguous()):
        raise ValueError("frames must be a contiguous uint8 [n, h, w, ch] tensor")
    nf, h, w, ch = (frames.shape if has_video else (0, 1, 1, 1))
    with torch.cuda.device(dev):
        side = _side_streams(dev, 3 * (slot + 1))[3 * slot: 3 * slot + 3]
        bounds = torch.empty((max_segments, 2), dtype=torch.float64, device=dev)
        count = torch.zeros((1,), dtype=torch.int32, device=dev)
        ssim = mse = e16 = e512 = None
        if nf > 1:
            ssim = torch.empty((nf - 1,), dtype=torch.float64, device=dev)
            mse = torch.empty((nf - 1,), dtype=torch.float64, device=dev)
        ns = nch = 0
        if pcm is not None:
            if pcm.dim() == 1:
                pcm = pcm.reshape(-1, 1)
            ns, nch = pcm.shape
            # the pyramid is scratch of this call: grow-only buffer per (slot, stream), not a fresh 29 MB block per call
            n16, n512 = max((ns + 15) // 16, 1), max((ns + 511) // 512, 1)
            pyr = _cuda.workspace((n16 + n512) * 8, dev, f"pattern_pyr{slot}").view(torch.float64)
            e16, e512 = pyr[:n16], pyr[n16:n16 + n512]
        ft = frame_times.to(dev, torch.float64).contiguous() if has_video else None
        ws = _cuda.workspace(lib.hippo_pattern_separation_workspace_bytes(nf, h, w, int(chunk_pairs)), dev, f"pattern{slot}")
        sides = (ctypes.c_void_p * 3)(*[s.cuda_stream for s in side])
        # every buffer was allocated on the caller's stream, the side streams start behind an event recorded on it
        # after that, and it is joined behind them before the call returns: to the caching allocator this is plain
        # single-stream use (record_stream on the side streams would only force a fresh cudaMalloc per call)
        _lib.check(lib.hippo_pattern_separation(
            _cuda.ptr(frames) if has_video else None, nf, h, w, ch, _cuda.ptr(ft),
            _cuda.ptr(pcm), _PCM_ENUM[pcm.dtype] if pcm is not None else _lib.HIPPO_F64, ns, max(nch, 1),
            float(sample_rate) if (pcm is not None and sample_rate) else 0.0,
            float(max_segment_duration), float(min_segment_duration), float(frame_similarity_threshold),
            float(audio_silence_threshold), int(chunk_pairs), _cuda.ptr(ssim), _cuda.ptr(mse), _cuda.ptr(e16),
            _cuda.ptr(e512), bounds.data_ptr(), count.data_ptr(), int(max_segments), ws.data_ptr(), ws.numel(),
            sides, _cuda.stream_ptr()))
    return bounds, count, ssim


def pattern_separation_host(frames, frame_times, pcm, sample_rate, max_segment_duration: float,
                            min_segment_duration: float, frame_similarity_threshold: float,
                            audio_silence_threshold: float, max_segments: int, chunk_frames: int = 384):
    """Temporal pattern separation of ONE stream whose frames and samples sit in HOST memory (what a caller of
    the reference has after decoding: hm:982-983 reads the frames from disk, bp:331-335 the samples).

    frames: uint8 [nf, h, w, ch] CPU tensor (pinned memory makes the upload a plain DMA), pcm: [ns, nch] CPU tensor
    (int16 PCM / float32 / float64) or None, frame_times: fp64 [nf] CPU or device tensor.  The frames are uploaded
    in chunks of `chunk_frames` (+1 frame of overlap, so every adjacent pair lives in exactly one chunk) through two
    device staging buffers on the copy stream; gray conversion + SSIM of chunk i run under the upload of chunk
    i + 1, the samples follow last, then the boundary kernel.  Returns (bounds fp64 [max_segments, 2] device,
    count int32 [1] device, ssim fp64 [nf - 1] device).  At 224 x 224 the pass is PCIe-bound (542 MB of frames)."""
    from .bank import _copy_stream

    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.is_cuda:
        raise ValueError("frames must be a uint8 [n, h, w, ch] CPU tensor")
    dev = _cuda.require_device()
    nf = frames.shape[0]
    with torch.cuda.device(dev):
        main = torch.cuda.current_stream()
        copy = _copy_stream(dev)
        copy.wait_stream(main)
        ssim = torch.empty((max(nf - 1, 1),), dtype=torch.float64, device=dev)
        mse = torch.empty((max(nf - 1, 1),), dtype=torch.float64, device=dev)
        cf = max(2, int(chunk_frames))
        stage = [torch.empty((min(cf, nf - 1) + 1,) + tuple(frames.shape[1:]), dtype=torch.uint8, device=dev)
                 for _ in range(2)] if nf > 1 else []
        up = [torch.cuda.Event(), torch.cuda.Event()]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        for i, f0 in enumerate(range(0, nf - 1, cf)):
            b = i & 1
            f1 = min(nf - 1, f0 + cf)                      # frames [f0, f1] -> pairs f0 .. f1 - 1
            m = f1 - f0 + 1
            with torch.cuda.stream(copy):
                if i >= 2:
                    copy.wait_event(done[b])
                stage[b][:m].copy_(frames[f0:f1 + 1], non_blocking=True)
                up[b].record(copy)
            main.wait_event(up[b])
            frame_pair_scores_device(stage[b][:m], range_mode=0, out=(ssim[f0:], mse[f0:]))
            done[b].record(main)
        pcm_d = pyr = None
        if pcm is not None:
            with torch.cuda.stream(copy):
                pcm_d = pcm.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            main.wait_event(ev)
            pcm_d.record_stream(main)
            if pcm_d.dim() == 1:
                pcm_d = pcm_d.reshape(-1, 1)
            pyr = audio_energy_device(pcm_d)
        ft_d = frame_times.to(dev, torch.float64, non_blocking=True) if frame_times is not None else None
        for st in stage:
            st.record_stream(main)
        bounds, count = segment_boundaries_device(ssim[: nf - 1] if nf > 1 else None, ft_d, pcm_d, pyr, sample_rate,
                                                  max_segment_duration, min_segment_duration,
                                                  frame_similarity_threshold, audio_silence_threshold, max_segments)
    return bounds, count, ssim[: max(nf - 1, 0)]


_side: dict = {}
_issue: dict = {}


def _issue_streams(dev: torch.device, n: int):
    pool = _issue.setdefault(dev.index, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=dev))
    return pool[:n]


def _side_streams(dev: torch.device, n: int):
    # triples (lane, lane, chain) for hippo_pattern_separation: the chain stream is a high-priority stream, so its one
    # small CTA is placed ahead of the thousands of pending SSIM CTAs when an SM frees up
    pool = _side.setdefault(dev.index, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=dev, priority=-1 if len(pool) % 3 == 2 else 0))
    return pool[:n]


def segment_boundaries_device(ssim: Optional[torch.Tensor], frame_times: Optional[torch.Tensor],
                              pcm: Optional[torch.Tensor], pyramid, sample_rate, max_segment_duration: float,
                              min_segment_duration: float, frame_similarity_threshold: float,
                              audio_silence_threshold: float, max_segments: int):
    """One stream through hippo_segment_boundaries. Returns (bounds fp64 [max_segments, 2], count int32 [1])."""
    bounds, counts = segment_boundaries_batch_device(
        [(ssim, frame_times, pcm, pyramid, sample_rate)], max_segment_duration, min_segment_duration,
        frame_similarity_threshold, audio_silence_threshold, max_segments)
    return bounds[0], counts


def segment_sequence(video_frames=None, frame_times=None, audio_data=None, audio_sample_rate=None, *,
                     max_segment_duration: float = 30.0, min_segment_duration: float = 10.0,
                     frame_similarity_threshold: float = 0.95, audio_silence_threshold: float = -40,
                     int16_pcm: bool = False) -> List[SequenceSegment]:
    """Drop-in for HippocampalMemory._segment_sequence (hm:1002-1114); thresholds default to
    config/default_config.yaml:27-30.  `video_frames` may be paths (as in the reference), BGR arrays,
    or one uint8 array [n, h, w, 3].  Integer audio is taken at face value, as the reference's NumPy
    arithmetic does; int16_pcm=True reads int16 samples as PCM (k / 32768, what sf.read yields, bp:331-335)."""
    segments: List[SequenceSegment] = []
    if video_frames is None and audio_data is None:                                   # hm:1024-1025
        return segments
    has_video = (video_frames is not None and len(video_frames) > 0
                 and frame_times is not None and len(frame_times) > 0)
    has_audio = audio_data is not None and bool(audio_sample_rate)
    if has_video:                                                                     # hm:1027-1032
        total = frame_times[-1] - frame_times[0]
    elif has_audio:
        total = len(audio_data) / audio_sample_rate
    else:
        return segments
    if not (total > 0.0):
        return segments
    if not (min_segment_duration > 0):
        raise ValueError("min_segment_duration must be positive (the reference loop would not terminate)")
    dev = _cuda.require_device()

    max_segments = int(math.ceil(total / min_segment_duration)) + 2
    thresholds = (max_segment_duration, min_segment_duration, frame_similarity_threshold, audio_silence_threshold)
    ft_d = frames_t = None
    if has_video:
        ft = np.asarray(frame_times, dtype=np.float64)
        if len(video_frames) != len(ft):
            raise ValueError("video_frames and frame_times differ in length")
        if np.any(np.diff(ft) < 0):
            raise ValueError("frame_times must be non-decreasing")
        ft_d = _cuda.to_device(ft, dev)
        if len(ft) > 1:
            if isinstance(video_frames, torch.Tensor):
                frames_t = video_frames.contiguous()
            else:
                frames_t = torch.from_numpy(_load_frames(video_frames))
            if frames_t.shape[1] < 7 or frames_t.shape[2] < 7:
                raise ValueError("win_size exceeds image extent.")
    if has_audio and int(0.5 * audio_sample_rate) < 1:
        raise ValueError("range() arg 3 must not be zero")                          # hm:1068 with a tiny rate
    if frames_t is not None and not frames_t.is_cuda:
        # host frames: chunked upload on the copy stream, gray + SSIM of chunk i under the upload of chunk i + 1
        pcm_h = None
        if has_audio:
            a = audio_data.detach().cpu() if isinstance(audio_data, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(audio_data))
            if not (a.dtype in (torch.float32, torch.float64) or (int16_pcm and a.dtype == torch.int16)):
                a = a.to(torch.float64)
            if a.dim() not in (1, 2):
                raise ValueError("audio_data must be (n,) or (n, channels)")
            pcm_h = a.reshape(-1, 1) if a.dim() == 1 else a.contiguous()
        bounds, count, _ = pattern_separation_host(frames_t, ft_d, pcm_h, audio_sample_rate if has_audio else None,
                                                   *thresholds, max_segments)
    else:
        pcm_d = _audio_to_device(audio_data, dev, int16_pcm) if has_audio else None
        bounds, count, _ = pattern_separation_device(frames_t, ft_d, pcm_d, audio_sample_rate if has_audio else None,
                                                     *thresholds, max_segments)
    c = int(count.item())
    if c < 0:
        raise RuntimeError("segment table overflow")
    b = bounds[:c].cpu().numpy()

    # hm:1087-1108: attach the frames / audio of each segment (list slicing, host side)
    ft_list = list(frame_times) if has_video else None
    for i in range(c):
        cs, opt = float(b[i, 0]), float(b[i, 1])
        seg = SequenceSegment(start_time=cs, end_time=opt)
        if has_video:
            seg.frames = [f for f, t in zip(video_frames, ft_list) if cs <= t <= opt]
            seg.frame_times = [t for t in ft_list if cs <= t <= opt]
        if has_audio:
            s0 = int(cs * audio_sample_rate)
            s1 = int(opt * audio_sample_rate)
            seg.audio_data = audio_data[s0:s1]
        segments.append(seg)
    return segments
