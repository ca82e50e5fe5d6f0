"""All ThetaEvents of the long-term store in ONE device bank.

The reference keeps every event's embeddings as decimal text inside the event's JSON file
(`ThetaEvent.to_dict` hm:110-133 -> `json.dump` hm:334-335; reloaded as float64, hm:386-395) and
detailed recall loops over the events, calling `top_k_cosine_similarity(query, event_features, k=5)`
once per event (hm:3143-3153 vision, hm:3295-3304 audio), then turns the hits into +-1 s windows and
keeps the five most similar (hm:3258-3277, hm:3366-3381).

`EventBank` concatenates one modality's rows of all events into a `MemoryBank` (bf16 rows + fp32
norms in HBM) with an event offset table and the events' time tables, so that ONE streaming pass
(`hippo_topk_segmented`) yields every event's top-5 and `hippo_recall_windows` does the tail of the
loop.  `save` / `load` keep the bank as a flat binary file instead of JSON text.
`find_relevant_segments` reproduces the reference's return value (`List[SequenceSegment]`).
"""
from __future__ import annotations

import json
import os
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from . import _cuda, _lib
from .bank import MemoryBank
from .segmentation import SequenceSegment

_MAGIC = b"HIPPOBK1"
_ALIGN = 4096


def _event_rows(event, modality: str):
    feats = event.features if hasattr(event, "features") else event["features"]
    if modality not in feats:
        return None
    f = feats[modality]
    if isinstance(f, torch.Tensor):                      # hm:3149-3150
        f = f.detach().cpu().numpy()
    f = np.asarray(f)
    if f.ndim == 1:                                       # vo:173-174
        f = f.reshape(1, -1)
    return f


def _event_times(event, modality: str):
    """The time table the reference indexes with the top-k rows: the KEY-FRAME times for vision
    (hm:3262-3263; the indices come from the all-frame features, the guard `idx < len(frame_times)`
    is reproduced) and feature_times['audio_times'] for audio (hm:3367-3370)."""
    if modality == "vision":
        t = event.frame_times if hasattr(event, "frame_times") else event["frame_times"]
    else:
        ft = event.feature_times if hasattr(event, "feature_times") else event["feature_times"]
        t = ft[f"{modality}_times"]
    return np.asarray(t, dtype=np.float64).reshape(-1)


class EventBank:
    """One modality of every event: rows concatenated in event order on one GPU."""

    def __init__(self, bank: MemoryBank, offsets: np.ndarray, time_offsets: np.ndarray, times: np.ndarray,
                 event_index: Optional[Sequence[int]] = None, modality: str = "vision"):
        self.bank = bank
        self.modality = modality
        self.offsets = np.asarray(offsets, dtype=np.int64)
        self.time_offsets = np.asarray(time_offsets, dtype=np.int64)
        self.times = np.asarray(times, dtype=np.float64)
        self.nev = len(self.offsets) - 1
        if self.nev < 0 or self.offsets[0] != 0 or self.offsets[-1] != bank.n or np.any(np.diff(self.offsets) < 0):
            raise ValueError("offsets must ascend from 0 to the number of bank rows")
        if len(self.time_offsets) != self.nev + 1 or self.time_offsets[-1] != len(self.times):
            raise ValueError("time_offsets do not match the events / the time table")
        # position of each bank event in the caller's event list (events without the modality are skipped)
        self.event_index = np.arange(self.nev) if event_index is None else np.asarray(event_index, dtype=np.int64)
        dev = bank.device
        self._offsets_d = torch.from_numpy(self.offsets).to(dev)
        self._toffsets_d = torch.from_numpy(self.time_offsets).to(dev)
        self._times_d = torch.from_numpy(self.times if len(self.times) else np.zeros(1)).to(dev)

    # ------------------------------------------------------------------ build ----
    @classmethod
    def from_events(cls, events, modality: str = "vision", device=None, keep_rows: bool = False) -> "EventBank":
        """Events are the reference's ThetaEvent objects (or dicts with the same fields); those without
        `modality` in their features are skipped, as hm:3144-3145 / hm:3296-3297 skip them.
        keep_rows=True also keeps the rows in their own precision (fp32, or the fp64 of a store reloaded from JSON,
        hm:391) on the device: `search(..., exact=True)` re-scores its bf16 candidates from them."""
        rows, times, index = [], [], []
        for i, ev in enumerate(events):
            f = _event_rows(ev, modality)
            if f is None:
                continue
            rows.append(f)
            times.append(_event_times(ev, modality))
            index.append(i)
        if not rows:
            raise ValueError(f"no event carries {modality!r} features")
        d = rows[0].shape[1]
        for f in rows:
            if f.shape[1] != d:
                raise ValueError("events disagree on the feature dimension")
        offsets = np.zeros(len(rows) + 1, dtype=np.int64)
        offsets[1:] = np.cumsum([len(f) for f in rows])
        toffsets = np.zeros(len(rows) + 1, dtype=np.int64)
        toffsets[1:] = np.cumsum([len(t) for t in times])
        wide = any(f.dtype == np.float64 for f in rows)
        bank = MemoryBank(int(offsets[-1]), d, device=device, keep_rows=keep_rows,
                          rows_dtype=torch.float64 if wide else torch.float32)
        # events are small (hundreds of rows): runs of consecutive events share one pinned chunk, one DMA and one build
        # launch (a fill per event is a 1 MB copy + a launch each, ~0.5 ms: a second for a store of 2,000 events)
        bank.fill_from_parts([f if f.dtype in (np.float32, np.float64) else f.astype(np.float32) for f in rows])
        return cls(bank, offsets, toffsets, np.concatenate(times) if times else np.zeros(0), index, modality)

    # ----------------------------------------------------------------- search ----
    def _query(self, query) -> torch.Tensor:
        if isinstance(query, torch.Tensor):
            q = query.detach().reshape(-1).to(self.bank.device, torch.float32)
        else:
            q = _cuda.to_device(np.asarray(query, dtype=np.float32).reshape(-1), self.bank.device)
        if q.numel() != self.bank.d:
            raise ValueError(f"expected a query of dimension {self.bank.d}, got {q.numel()}")
        if self.bank.d_pad != self.bank.d:
            qp = torch.zeros((self.bank.d_pad,), dtype=torch.float32, device=self.bank.device)
            qp[: self.bank.d] = q
            q = qp
        return q.contiguous()

    def scores(self, query) -> torch.Tensor:
        """Cosine similarity of the query to EVERY row (fp32 [n] device tensor), vo:178-182 per row."""
        lib = _lib.load()
        q = self._query(query)
        b = self.bank
        out = torch.empty((max(b.n, 1),), dtype=torch.float32, device=b.device)
        with torch.cuda.device(b.device):
            _lib.check(lib.hippo_scores_single(b.rows.data_ptr(), b.norm.data_ptr(), b.n, b.d_pad, q.data_ptr(),
                                               out.data_ptr(), _cuda.stream_ptr()))
        return out[: b.n]

    def search(self, query, k: int = 5, exact: bool = False):
        """Every event's top-k in one pass: (idx int64 [nev, k] event-local rows, -1 padded;
        score fp32 [nev, k]; maxsim fp32 [nev]) as device tensors.
        exact=True (bank built with keep_rows): the bf16 pass nominates HIPPO_TOPK_MAX candidates per event, every
        candidate is re-scored from the original rows (`hippo_rescore`) and re-ranked; an event whose k-th exact score
        does not clear `last bf16 candidate + eps` (a rigorous bound on the bf16 error, see bank.py) is searched again
        straight from its original rows (`hippo_topk_rows`).  Rows and order then equal the reference's wherever its
        own scores differ by more than rounding noise."""
        if k < 1:
            raise ValueError("k must be >= 1")
        if exact:
            return self._search_exact(query, k)
        lib = _lib.load()
        q = self._query(query)
        b = self.bank
        dev = b.device
        idx = torch.empty((self.nev, k), dtype=torch.int64, device=dev)
        score = torch.empty((self.nev, k), dtype=torch.float32, device=dev)
        mx = torch.empty((max(self.nev, 1),), dtype=torch.float32, device=dev)
        if self.nev:
            with torch.cuda.device(dev):
                ws_bytes = lib.hippo_topk_segmented_workspace_bytes(b.n)
                ws = _cuda.workspace(ws_bytes, dev, "segmented")
                _lib.check(lib.hippo_topk_segmented(
                    b.rows.data_ptr(), b.norm.data_ptr(), b.n, b.d_pad, q.data_ptr(), self._offsets_d.data_ptr(),
                    self.nev, k, idx.data_ptr(), score.data_ptr(), mx.data_ptr(), ws.data_ptr(), ws.numel(),
                    _cuda.stream_ptr()))
        return idx, score, mx[: self.nev]

    def _search_exact(self, query, k: int):
        from .bank import _EPS_BF16, _T2ENUM

        b = self.bank
        if b.src is None:
            raise ValueError("exact search needs the original rows: build the EventBank with keep_rows=True")
        lib = _lib.load()
        dev = b.device
        kc = _lib.HIPPO_TOPK_MAX
        if k > kc:
            raise ValueError(f"exact per-event search supports k <= {kc}")
        c_idx, c_score, _ = self.search(query, kc)                  # bf16 candidates, event-local rows
        q = self._query(query)[: b.d].contiguous()
        nev = self.nev
        glob = torch.where(c_idx >= 0, c_idx + self._offsets_d[:nev, None], c_idx).contiguous()
        r_key = torch.empty((nev, kc), dtype=torch.int64, device=dev)
        idx = torch.empty((nev, k), dtype=torch.int64, device=dev)
        score = torch.empty((nev, k), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = _cuda.stream_ptr()
            _lib.check(lib.hippo_rescore(b.src.data_ptr(), _T2ENUM[b.src.dtype], b.n, b.d, b.src.stride(0), 0, q.data_ptr(),
                                         _lib.HIPPO_F32, 0, nev, glob.data_ptr(), kc, r_key.data_ptr(), None, stream))
            _lib.check(lib.hippo_topk_merge(r_key.data_ptr(), 1, nev, kc, k, idx.data_ptr(), score.data_ptr(), None, stream))
        idx = torch.where(idx >= 0, idx - self._offsets_d[:nev, None], idx)       # back to event-local rows
        # events with more rows than candidates: is the candidate list provably sufficient?
        eps = (0.0 if b.bf16_exact else _EPS_BF16)
        eps = eps + b.d * 1.2e-7
        sizes = self._offsets_d[1:] - self._offsets_d[:-1]
        kth = score[:, k - 1]
        ok = (sizes <= kc) | ((idx[:, k - 1] >= 0) & (torch.isnan(kth) | (kth >= c_score[:, kc - 1] + eps)))
        # the others: EVERY row of the event re-scored from the original rows, all such events in one launch pair
        # (video-like events hold runs of near-duplicate frames, so this is not rare: a Python loop of per-event
        # passes cost milliseconds per query on a store of 2,000 events)
        bad = torch.nonzero(~ok).reshape(-1)                       # the one host synchronisation of this path
        sizes_h = np.diff(self.offsets)
        bad_h = bad.cpu().numpy()
        step = max(1, (1 << 24) // max(int(sizes_h[bad_h].max()) if len(bad_h) else 1, 1))     # <= 16M candidates a round
        for c0 in range(0, len(bad_h), step):
            sel = bad[c0:c0 + step]
            sz = sizes[sel]
            m = int(sizes_h[bad_h[c0:c0 + step]].max())
            col = torch.arange(m, dtype=torch.int64, device=dev)[None, :]
            cand = torch.where(col < sz[:, None], self._offsets_d[sel][:, None] + col, torch.full_like(col, -1)).contiguous()
            keys = torch.empty((sel.numel(), m), dtype=torch.int64, device=dev)
            i2 = torch.empty((sel.numel(), k), dtype=torch.int64, device=dev)
            s2 = torch.empty((sel.numel(), k), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(lib.hippo_rescore(b.src.data_ptr(), _T2ENUM[b.src.dtype], b.n, b.d, b.src.stride(0), 0,
                                             q.data_ptr(), _lib.HIPPO_F32, 0, sel.numel(), cand.data_ptr(), m,
                                             keys.data_ptr(), None, stream))
                _lib.check(lib.hippo_topk_merge(keys.data_ptr(), 1, sel.numel(), m, k, i2.data_ptr(), s2.data_ptr(),
                                                None, stream))
            idx[sel] = torch.where(i2 >= 0, i2 - self._offsets_d[sel][:, None], i2)
            score[sel] = s2
        valid = idx >= 0
        # np.max over the event's top-k as the reference takes it (hm:3156): NaN wins
        mx = torch.where(valid, score, torch.full_like(score, float("-inf"))).max(dim=1).values
        mx = torch.where((valid & torch.isnan(score)).any(dim=1), torch.full_like(mx, float("nan")), mx)
        return idx, score, mx

    def recall_windows(self, idx: torch.Tensor, score: torch.Tensor, top: int = 5, pad: float = 1.0,
                       enabled: Optional[np.ndarray] = None):
        """Tail of the recall loop on the device. Returns host arrays
        (event [c] positions in this bank, row [c], similarity [c] fp32, window [c, 2] fp64), c <= top."""
        lib = _lib.load()
        dev = self.bank.device
        k = idx.shape[1]
        ev = torch.empty((top,), dtype=torch.int32, device=dev)
        oi = torch.empty((top,), dtype=torch.int64, device=dev)
        sc = torch.empty((top,), dtype=torch.float32, device=dev)
        win = torch.empty((top, 2), dtype=torch.float64, device=dev)
        cnt = torch.zeros((1,), dtype=torch.int32, device=dev)
        en = None
        if enabled is not None:
            en = torch.from_numpy(np.ascontiguousarray(enabled, dtype=np.uint8)).to(dev)
            if en.numel() != self.nev:
                raise ValueError("enabled must have one entry per bank event")
        with torch.cuda.device(dev):
            _lib.check(lib.hippo_recall_windows(
                idx.contiguous().data_ptr(), score.contiguous().data_ptr(), self.nev, k, self._toffsets_d.data_ptr(),
                self._times_d.data_ptr(), _cuda.ptr(en), float(pad), top, ev.data_ptr(), oi.data_ptr(), sc.data_ptr(),
                win.data_ptr(), cnt.data_ptr(), _cuda.stream_ptr()))
        # one readback: everything as fp64 columns (event and row numbers are far below 2^53)
        packed = torch.cat([cnt.to(torch.float64), ev.to(torch.float64), oi.to(torch.float64), sc.to(torch.float64),
                            win.reshape(-1)]).cpu().numpy()
        c = int(packed[0])
        ev_h = packed[1:1 + top][:c].astype(np.int32)
        oi_h = packed[1 + top:1 + 2 * top][:c].astype(np.int64)
        sc_h = packed[1 + 2 * top:1 + 3 * top][:c].astype(np.float32)
        win_h = packed[1 + 3 * top:].reshape(top, 2)[:c].copy()
        return ev_h, oi_h, sc_h, win_h

    # ------------------------------------------------------------ persistence ----
    def save(self, path: str) -> None:
        """Flat binary file: magic, JSON directory, 4 KiB-aligned raw sections (bf16 rows, fp32 norms,
        offsets, time tables).  Replaces the decimal-text embeddings of hm:110-133 / hm:334-335."""
        b = self.bank
        sections = {
            "rows_bf16": b.rows[: b.n].contiguous().view(torch.int16).cpu().numpy().view(np.uint16),
            "norm_f32": b.norm[: b.n].cpu().numpy(),
            "offsets": self.offsets,
            "time_offsets": self.time_offsets,
            "times": self.times,
            "event_index": self.event_index.astype(np.int64),
        }
        directory, pos = {}, 0
        for name, arr in sections.items():
            directory[name] = dict(dtype=arr.dtype.str, shape=list(arr.shape), offset=pos, nbytes=int(arr.nbytes))
            pos += (arr.nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        head = json.dumps(dict(version=1, n=b.n, d=b.d, d_pad=b.d_pad, modality=self.modality,
                               sections=directory)).encode()
        head_len = (len(_MAGIC) + 8 + len(head) + _ALIGN - 1) // _ALIGN * _ALIGN
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:
            f.write(_MAGIC)
            f.write(np.uint64(len(head)).tobytes())
            f.write(head)
            f.write(b"\0" * (head_len - len(_MAGIC) - 8 - len(head)))
            for name, arr in sections.items():
                f.write(np.ascontiguousarray(arr).tobytes())
                f.write(b"\0" * ((-arr.nbytes) % _ALIGN))
        os.replace(tmp, path)

    @classmethod
    def load(cls, path: str, device=None) -> "EventBank":
        with open(path, "rb") as f:
            if f.read(len(_MAGIC)) != _MAGIC:
                raise ValueError(f"{path}: not a hippomm_b200 bank file")
            hl = int(np.frombuffer(f.read(8), dtype=np.uint64)[0])
            meta = json.loads(f.read(hl).decode())
        if meta.get("version") != 1:
            raise ValueError(f"{path}: unsupported bank file version {meta.get('version')}")
        base = (len(_MAGIC) + 8 + hl + _ALIGN - 1) // _ALIGN * _ALIGN

        def section(name):
            s = meta["sections"][name]
            count = int(np.prod(s["shape"])) if s["shape"] else 1
            return np.memmap(path, dtype=np.dtype(s["dtype"]), mode="r", offset=base + s["offset"],
                             shape=tuple(s["shape"])) if count else np.zeros(s["shape"], dtype=np.dtype(s["dtype"]))

        bank = MemoryBank(meta["n"], meta["d"], device=device)
        if meta["n"]:
            rows = torch.from_numpy(np.array(section("rows_bf16")).view(np.int16))
            bank.rows[: meta["n"]].view(torch.int16).copy_(rows.to(bank.device))
            bank.norm[: meta["n"]].copy_(torch.from_numpy(np.array(section("norm_f32"))).to(bank.device))
            bank._inexact.fill_(1)     # the fp32 originals are gone: treat as not bf16-exact
        return cls(bank, np.array(section("offsets")), np.array(section("time_offsets")), np.array(section("times")),
                   np.array(section("event_index")), meta.get("modality", "vision"))


def find_relevant_segments(query_features, events, bank: Optional[EventBank] = None, modality: str = "vision",
                           low_similarity: Optional[Callable] = None, top: int = 5, k: int = 5,
                           pad: float = 1.0, searched=None, segment_cls=None) -> List[SequenceSegment]:
    """Drop-in for the arithmetic of `_find_relevant_video_segments` (hm:3129-3279, modality
    "vision") and `_find_relevant_audio_segments` (hm:3281-3383, modality "audio").

    `low_similarity(event, top_k_indices, top_k_similarities)` is consulted for an event whose best
    similarity is below 0.4 and that has captions / a holistic transcription (hm:3156, hm:3307: the
    reference asks an LLM there); it returns a list of (similarity, [SequenceSegment]) entries that
    take part in the final ranking, or None to fall through to the similarity path, which is also
    what the reference does when the LLM call fails (hm:3256-3272).  With no callback every event takes
    the similarity path.
    """
    events = list(events)
    if bank is None:
        bank = EventBank.from_events(events, modality)
    Seg = segment_cls or SequenceSegment                          # install() hands in the reference's own dataclass
    idx_d, score_d, max_d = searched if searched is not None else bank.search(query_features, k)
    extra = []                                     # (similarity, position of the event, segments) from the callback
    enabled = np.ones(bank.nev, dtype=np.uint8)
    if low_similarity is not None:
        mx = max_d.cpu().numpy()
        idx_h = sc_h = None
        for j in np.nonzero(mx < 0.4)[0]:          # NaN < 0.4 is False, as in hm:3156
            ev = events[int(bank.event_index[j])]
            text = getattr(ev, "frame_captions", None) if modality == "vision" else \
                getattr(ev, "holistic_audio_transcription", None)
            if not text:
                continue
            if idx_h is None:
                idx_h, sc_h = idx_d.cpu().numpy(), score_d.cpu().numpy()
            valid = idx_h[j] >= 0
            res = low_similarity(ev, idx_h[j][valid], sc_h[j][valid])
            if res is not None:
                enabled[j] = 0
                extra.extend((float(s), int(j), segs) for s, segs in res)
    ev_pos, rows, sims, wins = bank.recall_windows(idx_d, score_d, top, pad, enabled)

    ranked = []                                    # (similarity, event position, order inside the event, segments)
    for r in range(len(ev_pos)):
        j = int(ev_pos[r])
        ev = events[int(bank.event_index[j])]
        t0, t1 = float(wins[r, 0]), float(wins[r, 1])
        if modality == "vision":                   # hm:3264-3271
            t = float(bank.times[bank.time_offsets[j] + rows[r]])
            ft = list(ev.frame_times)
            seg = Seg(start_time=t0, end_time=t1,
                      frames=[ev.frames[i] for i in range(len(ev.frames)) if t - pad <= ft[i] <= t + pad],
                      frame_times=[x for x in ft if t - pad <= x <= t + pad])
        else:                                      # hm:3368-3372
            seg = Seg(start_time=t0, end_time=t1, audio_data=None)
        ranked.append((float(sims[r]), j, r, [seg]))
    if extra:
        # the callback's entries sit where the event sits in the reference's list: stable sort by similarity
        merged = [(s, j, 0, o, segs) for s, j, o, segs in ranked] + \
                 [(s, j, 1, o, segs) for o, (s, j, segs) in enumerate(extra)]
        merged.sort(key=lambda x: (x[1], x[2], x[3]))          # the reference's append order: by event
        merged.sort(key=lambda x: x[0], reverse=True)          # hm:3274 / hm:3379 (stable)
        ranked = [(m[0], m[1], m[3], m[4]) for m in merged[:top]]
    out: List[SequenceSegment] = []
    for _, _, _, segs in ranked[:top]:
        out.extend(segs)
    return out
