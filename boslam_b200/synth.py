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


def local_map_scene(n_points: int, n_edges: int, n_frame: int, seed: int = 0, p_flip: float = 0.04):
    """Synthetic local map + frame for the tracking step (reference ``slam/tracking.py:91-128``).

    Returns a dict: store arrays ``desc uint8[n_points,32]``, ``pt3d float64[n_points,3]``,
    ``normal float64[n_points,3]`` (unit, ``frame.t - point`` direction jittered, as
    ``slam/covisibility_graph.py:127`` builds it), ``edges int32[n_edges]`` (map-point slot per (keyframe,
    map point) edge; points seen by several local keyframes repeat - SURVEY.md finding 4), a pose
    ``R, t, see_vector`` (``camera.py:24-29``) and a frame ``des uint8[n_frame,32]``, ``kp float64[n_frame,2]``
    (integer-valued pixels like ``Frame.kp_arr``) whose first ~70 % rows observe random visible points.
    """
    rng = np.random.default_rng(seed)
    ang = rng.uniform(-0.3, 0.3, 3)
    cx_, sx = np.cos(ang[0]), np.sin(ang[0])
    cy_, sy = np.cos(ang[1]), np.sin(ang[1])
    cz_, sz = np.cos(ang[2]), np.sin(ang[2])
    Rx = np.array([[1, 0, 0], [0, cx_, -sx], [0, sx, cx_]])
    Ry = np.array([[cy_, 0, sy], [0, 1, 0], [-sy, 0, cy_]])
    Rz = np.array([[cz_, -sz, 0], [sz, cz_, 0], [0, 0, 1]])
    R = Rz @ Ry @ Rx
    t = rng.uniform(-0.5, 0.5, 3)
    # points in camera coordinates: a frustum a little wider than the image so some fall outside,
    # a few behind the camera; world = R^T (Xc - t)
    fx, fy, cx, cy = 384.239013671875, 384.239013671875, 322.432373046875, 239.6533203125
    z = rng.uniform(0.5, 8.0, n_points)
    z[rng.random(n_points) < 0.03] *= -1.0
    u = rng.uniform(-120.0, WIDTH + 120.0, n_points)
    v = rng.uniform(-90.0, HEIGHT + 90.0, n_points)
    Xc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], axis=1)
    pt3d = (Xc - t) @ R  # R^T applied to rows
    see = R @ np.array([0.0, 0.0, 1.0])
    see = see / np.linalg.norm(see)
    normal = rng.normal(0.0, 1.0, (n_points, 3))
    normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    desc = rng.integers(0, 256, (n_points, DESC_BYTES), dtype=np.uint8)
    edges = np.concatenate([rng.permutation(n_points)[:min(n_points, n_edges)],
                            rng.integers(0, n_points, max(0, n_edges - n_points))]).astype(np.int32)
    rng.shuffle(edges)
    # frame: noisy observations of points that are inside the image
    inside = np.nonzero((u >= 0) & (u < WIDTH) & (v >= 0) & (v < HEIGHT) & (z > 0))[0]
    n_true = min(int(0.7 * n_frame), len(inside))
    src = rng.choice(inside, n_true, replace=False) if n_true else np.zeros(0, np.int64)
    des = rng.integers(0, 256, (n_frame, DESC_BYTES), dtype=np.uint8)
    kp = np.stack([rng.integers(0, WIDTH, n_frame), rng.integers(0, HEIGHT, n_frame)], axis=1).astype(np.float64)
    if n_true:
        flips = np.packbits(rng.random((n_true, DESC_BYTES * 8)) < p_flip, axis=1)
        des[:n_true] = desc[src] ^ flips
        kp[:n_true, 0] = np.clip(np.round(u[src] + rng.normal(0, 2.0, n_true)), 0, WIDTH - 1)
        kp[:n_true, 1] = np.clip(np.round(v[src] + rng.normal(0, 2.0, n_true)), 0, HEIGHT - 1)
    return {"desc": desc, "pt3d": pt3d, "normal": normal, "edges": edges, "R": R, "t": t, "see_vector": see,
            "des": des, "kp": kp, "fx": fx, "fy": fy, "cx": cx, "cy": cy, "width": WIDTH, "height": HEIGHT}


def rgbd_frame_pair(n_desc: int = 1000, seed: int = 0, p_flip: float = 0.04):
    """BASELINE config 1: two synthetic 640x480 RGB-D frames with ``n_desc`` ORB descriptors each, shaped
    like the reference's ``Frame`` (``camera.py:6-43``): ``des uint8[N,32]``, ``kp_arr int[N,2]``,
    ``cloud_kp float64[N,3]``.  ~70 % of the previous frame's features reappear in the current frame
    (noisy descriptor, pixel re-projected through a small known motion), the rest are new.
    Returns (prev, cur, R, t) with prev / cur dicts of those three arrays."""
    rng = np.random.default_rng(seed)
    fx = fy = 384.239013671875
    cx, cy = 322.432373046875, 239.6533203125
    ang = rng.uniform(-0.03, 0.03, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(ang[0]), -np.sin(ang[0])], [0, np.sin(ang[0]), np.cos(ang[0])]])
    Ry = np.array([[np.cos(ang[1]), 0, np.sin(ang[1])], [0, 1, 0], [-np.sin(ang[1]), 0, np.cos(ang[1])]])
    Rz = np.array([[np.cos(ang[2]), -np.sin(ang[2]), 0], [np.sin(ang[2]), np.cos(ang[2]), 0], [0, 0, 1]])
    R = Rz @ Ry @ Rx
    t = rng.uniform(-0.05, 0.05, 3)

    def frame(n):
        u = rng.integers(8, WIDTH - 8, n)
        v = rng.integers(8, HEIGHT - 8, n)
        z = rng.uniform(0.6, 6.0, n)
        cloud = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], axis=1)
        return {"des": rng.integers(0, 256, (n, DESC_BYTES), dtype=np.uint8), "kp_arr": np.stack([u, v], 1).astype(np.int64),
                "cloud_kp": cloud}

    prev, cur = frame(n_desc), frame(n_desc)
    n_true = int(0.7 * n_desc)
    src = rng.permutation(n_desc)[:n_true]
    dst = rng.permutation(n_desc)[:n_true]
    Xc = prev["cloud_kp"][src] @ R.T + t                      # the same 3-D points in the current camera
    u = Xc[:, 0] / Xc[:, 2] * fx + cx + rng.normal(0, 0.5, n_true)
    v = Xc[:, 1] / Xc[:, 2] * fy + cy + rng.normal(0, 0.5, n_true)
    ok = (u >= 0) & (u < WIDTH) & (v >= 0) & (v < HEIGHT)
    src, dst, u, v, Xc = src[ok], dst[ok], u[ok], v[ok], Xc[ok]
    flips = np.packbits(rng.random((len(src), DESC_BYTES * 8)) < p_flip, axis=1)
    cur["des"][dst] = prev["des"][src] ^ flips
    cur["kp_arr"][dst] = np.stack([np.round(u), np.round(v)], 1).astype(np.int64)
    cur["cloud_kp"][dst] = Xc
    return prev, cur, R, t
