"""Seeded synthetic ORB-descriptor workloads.

The reference's own precedent for fabricated descriptors is ``utils.int2orb`` (reference
``utils.py:53-55``: 32 seeded random bytes per id).  All generators use
``np.random.default_rng(seed)`` and return C-contiguous ``uint8[N, 32]`` arrays, the layout
``cv2.ORB.detectAndCompute`` gives boslam (``camera.py:141``).
"""
from __future__ import annotations

import numpy as np

DESC_BYTES = 32
WIDTH, HEIGHT = 640, 480  # reference config.py:40-41


def uniform(n: int, seed: int = 0) -> np.ndarray:
    """Uniform random descriptors: distances ~ Binomial(256, 1/2); throughput-only data."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (n, DESC_BYTES), dtype=np.uint8)


def correlated(nq: int, nt: int, seed: int = 0, p_flip: float = 0.04, frac_true: float = 0.7):
    """Train uniform; 70 % of queries are a train row with each bit flipped w.p. ``p_flip``
    (mean distance ~10), the rest fresh uniform rows.  Returns (query, train, truth) where
    truth[i] is the source train row or -1."""
    rng = np.random.default_rng(seed)
    train = rng.integers(0, 256, (nt, DESC_BYTES), dtype=np.uint8)
    query = rng.integers(0, 256, (nq, DESC_BYTES), dtype=np.uint8)
    truth = np.full(nq, -1, dtype=np.int64)
    if nt == 0 or nq == 0:
        return query, train, truth
    is_true = rng.random(nq) < frac_true
    src = rng.integers(0, nt, nq)
    flips = np.packbits(rng.random((nq, DESC_BYTES * 8)) < p_flip, axis=1)
    query[is_true] = train[src[is_true]] ^ flips[is_true]
    truth[is_true] = src[is_true]
    return np.ascontiguousarray(query), train, truth


def tie_stress(n: int, seed: int = 0) -> np.ndarray:
    """Rows with only the first two bytes non-zero, drawn from {0..3}: many exact ties."""
    rng = np.random.default_rng(seed)
    a = np.zeros((n, DESC_BYTES), dtype=np.uint8)
    a[:, :2] = rng.integers(0, 4, (n, 2), dtype=np.uint8)
    return a


def duplicate_rows(n_unique: int, seed: int = 0, max_rep: int = 4) -> np.ndarray:
    """Each unique row repeated 1..max_rep times, models the per-(keyframe, map point) edge
    stacking of ``slam/tracking.py:97-110`` (SURVEY finding 4)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (n_unique, DESC_BYTES), dtype=np.uint8)
    reps = rng.integers(1, max_rep + 1, n_unique)
    return np.ascontiguousarray(np.repeat(base, reps, axis=0))


def window_scene(nq: int, nt: int, seed: int = 0, sigma_px: float = 3.0):
    """Tracking-style scene: descriptors from :func:`correlated`, pixel positions such that true
    pairs project within N(0, sigma) px of each other and everything else is uniform over the
    640x480 image.  Returns (query, train, q_xy float32[nq,2], t_xy float32[nt,2], truth)."""
    query, train, truth = correlated(nq, nt, seed)
    rng = np.random.default_rng(seed + 1000003)
    t_xy = np.stack([rng.uniform(0, WIDTH, nt), rng.uniform(0, HEIGHT, nt)], axis=1)
    q_xy = np.stack([rng.uniform(0, WIDTH, nq), rng.uniform(0, HEIGHT, nq)], axis=1)
    has = truth >= 0
    q_xy[has] = t_xy[truth[has]] + rng.normal(0.0, sigma_px, (int(has.sum()), 2))
    return query, train, q_xy.astype(np.float32), t_xy.astype(np.float32), truth


def keyframe_pairs(n_pairs: int, n_desc: int, seed: int = 0, shared_query: bool = False):
    """Local-mapping / loop-closing style batch: ``n_pairs`` independent (query KF, train KF)
    problems of ``n_desc`` descriptors each.  With ``shared_query`` the same query keyframe is
    matched against every candidate (loop closing: current KF vs BoW candidates,
    reference ``slam/loop_closing.py:13-15``)."""
    qs, ts = [], []
    q0 = None
    for p in range(n_pairs):
        q, t, _ = correlated(n_desc, n_desc, seed * 100003 + p)
        if shared_query:
            if q0 is None:
                q0 = q
            q = q0
        qs.append(q)
        ts.append(t)
    return qs, ts


def keyframe_pair_batch(n_pairs: int, n_desc: int, seed: int = 0, p_flip: float = 0.04, frac_true: float = 0.7):
    """Vectorised :func:`keyframe_pairs`: packed (query uint8[P*n,32], train uint8[P*n,32]) with
    problem p in rows [p*n, (p+1)*n) of both arrays.  Same statistics as :func:`correlated`
    (70 % of queries are a noisy copy of a train row of the same pair)."""
    rng = np.random.default_rng(seed)
    total = n_pairs * n_desc
    train = rng.integers(0, 256, (total, DESC_BYTES), dtype=np.uint8)
    query = rng.integers(0, 256, (total, DESC_BYTES), dtype=np.uint8)
    if total == 0:
        return query, train
    is_true = rng.random(total) < frac_true
    src = rng.integers(0, n_desc, total) + np.repeat(np.arange(n_pairs) * n_desc, n_desc)
    # sparse bit flips: ~p_flip*256 flipped bits per true row, drawn as positions
    rows = np.nonzero(is_true)[0]
    query[rows] = train[src[rows]]
    n_flips = rng.binomial(DESC_BYTES * 8, p_flip, len(rows))
    r_idx = np.repeat(rows, n_flips)
    bit = rng.integers(0, DESC_BYTES * 8, len(r_idx))
    np.bitwise_xor.at(query, (r_idx, bit >> 3), (1 << (bit & 7)).astype(np.uint8))
    return query, train
