"""The reference's own implementation of the path: live ``cv2.BFMatcher``.  TEST INFRASTRUCTURE ONLY.

boslam does not contain matcher code; every call site hands two uint8[N,32] arrays to
``cv2.BFMatcher_create(cv2.NORM_HAMMING, crossCheck=True).match`` (reference
``slam/tracking.py:45,56,121``; ``slam/local_mapping.py:21``; ``slam/covisibility_graph.py:34``;
``experiments/pnp_one_way_tracking.py:11,30``).  OpenCV is an installed third-party package in
this image (opencv-python-headless 4.13.0; also present on the GPU box), so "running the
reference" for this path means calling it.  These wrappers turn its ``DMatch`` tuples into the
structure-of-arrays form the parity tests compare against, and are what
``bench.py --impl reference`` / ``cpu_baseline.kind == "reference"`` time.

Importers: tests/, tests/golden/make_golden.py, bench.py (reference arm + cpu_baseline),
__graft_entry__.smoke().  Never the product package.
"""
from __future__ import annotations

import numpy as np

try:  # cv2 is part of the image; keep the import soft so the numpy/C port can stand in
    import cv2  # type: ignore
    HAVE_CV2 = True
except Exception:  # pragma: no cover
    cv2 = None
    HAVE_CV2 = False


def version() -> str:
    return cv2.__version__ if HAVE_CV2 else "absent"


def threads() -> int:
    return int(cv2.getNumThreads()) if HAVE_CV2 else 0


def matcher(cross_check: bool = False):
    """Exactly the constructor boslam uses (slam/tracking.py:45)."""
    return cv2.BFMatcher_create(cv2.NORM_HAMMING, crossCheck=bool(cross_check))


def dmatches_to_arrays(ms):
    n = len(ms)
    qi = np.fromiter((m.queryIdx for m in ms), dtype=np.int32, count=n)
    ti = np.fromiter((m.trainIdx for m in ms), dtype=np.int32, count=n)
    d = np.fromiter((m.distance for m in ms), dtype=np.float64, count=n)
    di = d.astype(np.int32)
    assert np.all(di == d), "cv2 Hamming distances are integer-valued"
    return qi, ti, di


def match(q, t, cross_check: bool = True, mask=None):
    """``BFMatcher.match`` -> (queryIdx, trainIdx, distance) int32 arrays."""
    m = matcher(cross_check)
    ms = m.match(q, t) if mask is None else m.match(q, t, mask)
    return dmatches_to_arrays(ms)


def knn(q, t, k: int, mask=None):
    """``BFMatcher.knnMatch`` -> dense (idx int32[Q,k], dist int32[Q,k]), -1 padded (rule R3)."""
    m = matcher(False)
    rows = m.knnMatch(q, t, k) if mask is None else m.knnMatch(q, t, k, mask)
    Q = len(q)
    idx = np.full((Q, k), -1, dtype=np.int32)
    dist = np.full((Q, k), -1, dtype=np.int32)
    # knnMatch keeps empty rows (compactResult=False), one row per query in order
    assert len(rows) in (Q, 0)
    for i, row in enumerate(rows):
        for c, dm in enumerate(row):
            assert dm.queryIdx == i
            idx[i, c] = dm.trainIdx
            dist[i, c] = int(dm.distance)
            assert dist[i, c] == dm.distance
    return idx, dist


def ratio_match(q, t, ratio: float, mask=None):
    """Lowe ratio test written the way a boslam-style call site would (Python floats, fp64)."""
    m = matcher(False)
    rows = m.knnMatch(q, t, 2) if mask is None else m.knnMatch(q, t, 2, mask)
    good = [r[0] for r in rows if len(r) == 2 and r[0].distance < ratio * r[1].distance]
    return dmatches_to_arrays(good)
