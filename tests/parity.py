"""Parity rules of SURVEY.md §8(d), shared by the tests."""
from __future__ import annotations

import numpy as np

TOL = 1e-3


def _eq(a, b, tol):
    if np.isnan(a) and np.isnan(b):
        return True
    if np.isnan(a) or np.isnan(b):
        return False
    return abs(float(a) - float(b)) <= tol


def check_topk(idx, score, ref_idx, ref_score, tol=TOL, what=""):
    """Scores within `tol` rank by rank; the index SET must agree inside every band of reference scores
    that is separated from its neighbours by more than `tol` (ties inside a band are unordered in the
    reference, vo:185).  The last band may continue past k, so there only membership is checked loosely:
    each of our rows in it must carry a score within tol of the band."""
    idx = np.asarray(idx).reshape(-1)
    score = np.asarray(score).reshape(-1)
    ref_idx = np.asarray(ref_idx).reshape(-1)
    ref_score = np.asarray(ref_score).reshape(-1)
    assert idx.shape == ref_idx.shape, f"{what}: length {idx.shape} vs reference {ref_idx.shape}"
    k = len(ref_idx)
    for r in range(k):
        assert _eq(score[r], ref_score[r], tol), f"{what}: rank {r} score {score[r]} vs reference {ref_score[r]}"
    r = 0
    while r < k:
        e = r + 1
        while e < k and _eq(ref_score[e], ref_score[e - 1], tol):
            e += 1
        if e < k:  # closed band: exact set equality
            assert set(idx[r:e].tolist()) == set(ref_idx[r:e].tolist()), (
                f"{what}: ranks {r}..{e - 1} rows {idx[r:e].tolist()} vs reference {ref_idx[r:e].tolist()}")
        r = e
    assert len(set(idx.tolist())) == len(idx), f"{what}: duplicate rows {idx.tolist()}"


def check_topk_exact(idx, score, ref_idx, ref_score, what=""):
    """Bit-exact scores (lattice data) and identical rows up to exact score ties."""
    score = np.asarray(score, dtype=np.float32).reshape(-1)
    ref_score = np.asarray(ref_score, dtype=np.float32).reshape(-1)
    same = (score.view(np.uint32) == ref_score.view(np.uint32)) | (np.isnan(score) & np.isnan(ref_score))
    assert same.all(), f"{what}: scores differ in bits: {score} vs {ref_score}"
    check_topk(idx, score, ref_idx, ref_score, tol=0.0, what=what)
