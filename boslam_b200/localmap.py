"""Frame-to-local-map tracking with the map resident on the GPU.

Mirrors what reference ``slam/tracking.py:91-128`` (``Tracker._track_local_map``) does between
"we have a pose guess" and "hand matched (pixel, 3-D point) pairs to CamOnlyBA":

* ``MapStore`` is the device copy of the per-map-point data the reference keeps in
  ``MapPoint.feat / pt3d / n`` (``slam/nodes.py:115-118``), addressed by a caller-assigned slot
  (the ``MapPoint.id``).  ``update`` is the upsert of ``slam/covisibility_graph.py:128-134`` and the
  descriptor / normal refresh of ``slam/nodes.py:153-154``.
* ``MapStore.track`` takes the local map as the list of (keyframe, map point) *edges* in the order the
  reference's double loop visits them (``slam/tracking.py:97-98``) and returns the visible edges
  (``:103-104``), the cross-check / gated matches (``:121``) and the gathered arrays of ``:128``.

All arithmetic happens in ``libbfm_b200.so``; nothing here computes a projection or a distance.
"""
from __future__ import annotations

import ctypes
import math
import weakref
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _ffi
from .engine import Engine

COS60 = math.cos(math.pi / 3)  # slam/tracking.py:20


@dataclass
class CameraModel:
    """Pinhole intrinsics; defaults are the reference's ``config.py:36-41``."""
    fx: float = 384.239013671875
    fy: float = 384.239013671875
    cx: float = 322.432373046875
    cy: float = 239.6533203125
    width: int = 640
    height: int = 480


def quaternion_from_rotation(R) -> tuple:
    """(w, x, y, z) as ``g2o.SE3Quat(R, t)`` stores it: Eigen's matrix -> quaternion conversion, then
    ``normalizeRotation`` (w >= 0, unit norm).  Plain Python floats (IEEE fp64)."""
    m = [[float(R[i][j]) for j in range(3)] for i in range(3)]
    tr = m[0][0] + m[1][1] + m[2][2]
    q = [0.0, 0.0, 0.0, 0.0]  # x, y, z, w (Eigen coefficient order)
    if tr > 0.0:
        t = math.sqrt(tr + 1.0)
        q[3] = 0.5 * t
        t = 0.5 / t
        q[0] = (m[2][1] - m[1][2]) * t
        q[1] = (m[0][2] - m[2][0]) * t
        q[2] = (m[1][0] - m[0][1]) * t
    else:
        i = 0
        if m[1][1] > m[0][0]:
            i = 1
        if m[2][2] > m[i][i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = math.sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        q[3] = (m[k][j] - m[j][k]) * t
        q[j] = (m[j][i] + m[i][j]) * t
        q[k] = (m[k][i] + m[i][k]) * t
    if q[3] < 0.0:
        q = [-c for c in q]
    n = math.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])
    return q[3] / n, q[0] / n, q[1] / n, q[2] / n


@dataclass
class TrackResult:
    visible_edges: np.ndarray   # int32[V]   edge numbers that passed (train row j = visible_edges[j])
    visible_pixels: Optional[np.ndarray]  # float64[V, 2] their projections (only with want_pixels)
    inds_frame: np.ndarray      # int32[M]   queryIdx  (slam/tracking.py:126)
    inds: np.ndarray            # int32[M]   trainIdx into the visible list
    distance: np.ndarray        # float32[M]
    edges: np.ndarray           # int32[M]   = visible_edges[inds]
    pts3d: np.ndarray           # float64[M, 3] = pts3d[inds]          (:128)
    kp: np.ndarray              # float64[M, 2] = frame.kp_arr[inds_frame]


class MapStore:
    """Device-resident map points (descriptor, 3-D point, normal per slot)."""

    def __init__(self, capacity: int, engine: Optional[Engine] = None, device: int = 0):
        self.engine = engine if engine is not None else Engine(device)
        self.capacity = int(capacity)
        self._lib = _ffi.lib()
        m = ctypes.c_void_p()
        _ffi.check(self.engine._h, self._lib.bfm_map_create(self.engine._h, self.capacity, ctypes.byref(m)))
        self._m = m
        self._fin = weakref.finalize(self, self._lib.bfm_map_destroy, m)
        self._keep_engine = self.engine  # the store must not outlive its handle

    def close(self):
        self._fin()

    def update(self, slots, desc=None, pt3d=None, normal=None):
        """Upsert: ``slots`` int[n] (each at most once per call); any of the fields may be omitted."""
        slots = np.ascontiguousarray(slots, np.int32).ravel()
        n = slots.shape[0]
        if len(np.unique(slots)) != n:
            raise ValueError("a slot may appear only once per update")

        def prep(a, dtype, width, name):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype)
            if a.shape != (n, width):
                raise ValueError(f"{name} must be {np.dtype(dtype).name}[{n}, {width}]")
            return a

        d = prep(desc, np.uint8, 32, "desc")
        p = prep(pt3d, np.float64, 3, "pt3d")
        nn = prep(normal, np.float64, 3, "normal")
        ptr = lambda a: a.ctypes.data if a is not None else None
        with self.engine._lock:
            rc = self._lib.bfm_map_update(self._m, n, slots.ctypes.data, ptr(d), ptr(p), ptr(nn))
            _ffi.check(self.engine._h, rc)

    def track(self, frame_des, frame_kp, R, t, see_vector, edges, cam: Optional[CameraModel] = None,
              cos_max: float = COS60, cross_check: bool = True, max_distance=30, strict: bool = False,
              k: int = 1, ratio=None, window_radius=None, want_pixels: bool = False) -> TrackResult:
        """One ``_track_local_map`` candidate + match step.  Defaults are the reference's: crossCheck
        matcher (``slam/tracking.py:45``), ``distance <= 30`` (``:121``).  ``window_radius`` adds the
        projection-guided window of the north star (|kp - projected pixel| < r in both axes).
        ``want_pixels`` also returns the projections of the visible points (16 bytes each over PCIe)."""
        cam = cam or CameraModel()
        q = np.ascontiguousarray(frame_des)
        if q.dtype != np.uint8:
            raise TypeError("frame descriptors must be uint8")
        if q.ndim != 2 or (q.shape[0] and q.shape[1] != 32):
            raise ValueError("frame descriptors must be uint8[N, 32]")
        nq = q.shape[0]
        kp = np.ascontiguousarray(frame_kp, np.float64)
        if kp.shape != (nq, 2):
            raise ValueError("frame_kp must be [N, 2]")
        edges = np.ascontiguousarray(edges, np.int32).ravel()
        ne = edges.shape[0]
        tp = _ffi.TrackParams()
        tp.q[:] = quaternion_from_rotation(R)
        tp.t[:] = [float(x) for x in np.asarray(t, np.float64).ravel()[:3]]
        tp.see_vector[:] = [float(x) for x in np.asarray(see_vector, np.float64).ravel()[:3]]
        tp.fx, tp.fy, tp.cx, tp.cy = float(cam.fx), float(cam.fy), float(cam.cx), float(cam.cy)
        tp.cos_max, tp.width, tp.height = float(cos_max), int(cam.width), int(cam.height)
        opts, none_pass = self.engine._options(k, ratio, cross_check, max_distance, strict)
        if window_radius is not None:
            opts.mask_kind, opts.window_radius = _ffi.MASK_WINDOW, float(window_radius)
        vis_e = np.empty(max(ne, 1), np.int32)
        vis_p = np.empty((max(ne, 1), 2), np.float64) if want_pixels else None
        mq, mt, md, me = (np.empty(max(nq, 1), np.int32) for _ in range(4))
        mp3 = np.empty((max(nq, 1), 3), np.float64)
        mkp = np.empty((max(nq, 1), 2), np.float64)
        nv, nm = ctypes.c_int32(0), ctypes.c_int32(0)
        with self.engine._lock:
            rc = self._lib.bfm_track_local_map(self._m, ctypes.byref(tp), edges.ctypes.data, ne, q.ctypes.data, kp.ctypes.data,
                                               nq, ctypes.byref(opts), vis_e.ctypes.data,
                                               vis_p.ctypes.data if want_pixels else None, mq.ctypes.data,
                                               mt.ctypes.data, md.ctypes.data, me.ctypes.data, mp3.ctypes.data,
                                               mkp.ctypes.data, ctypes.byref(nv), ctypes.byref(nm))
            _ffi.check(self.engine._h, rc)
        v, m = nv.value, (0 if none_pass else nm.value)
        return TrackResult(vis_e[:v], vis_p[:v] if want_pixels else None, mq[:m], mt[:m], md[:m].astype(np.float32), me[:m], mp3[:m], mkp[:m])


    def vote(self, edge_kf, inliers, top: int = 100):
        """``Counter(ids_matching_kfs[inds[inliers]]).most_common(top)`` (reference ``slam/tracking.py:154``) over the
        match list of the last :meth:`track` call, which is still on the device: ``edge_kf[e]`` = id of the keyframe
        of edge ``e`` (one per edge passed to ``track``), ``inliers`` = positions in the returned match list that the
        pose optimisation kept.  Returns ``[(keyframe id, count), ...]``, most frequent first, ties in order of first
        appearance - the order ``Counter.most_common`` yields.  Call it before the next ``track`` / ``update``."""
        edge_kf = np.ascontiguousarray(edge_kf, np.int32).ravel()
        inl = np.ascontiguousarray(inliers, np.int32).ravel()
        top = int(top)
        ids = np.empty(max(top, 1), np.int32)
        cnt = np.empty(max(top, 1), np.int32)
        n = ctypes.c_int32(0)
        with self.engine._lock:
            rc = self._lib.bfm_keyframe_vote(self._m, edge_kf.ctypes.data, edge_kf.shape[0], inl.ctypes.data, inl.shape[0], top,
                                             ids.ctypes.data, cnt.ctypes.data, ctypes.byref(n))
            _ffi.check(self.engine._h, rc)
        return list(zip(ids[:n.value].tolist(), cnt[:n.value].tolist()))


def select_representative(obs, counts, engine: Optional[Engine] = None, device: int = 0) -> np.ndarray:
    """Batched ``MapPoint.add_observation`` descriptor choice (reference ``slam/nodes.py:146-153``):
    ``obs`` uint8[P, max_obs, 32] (each point's stored observations, first ``counts[p]`` rows valid),
    returns int32[P]: the index of the observation with the least median Hamming distance to the
    others (numpy median / argmin conventions), -1 for a point without observations."""
    from .engine import default_engine
    eng = engine if engine is not None else default_engine(device)
    obs = np.ascontiguousarray(obs)
    if obs.dtype != np.uint8 or obs.ndim != 3 or obs.shape[2] != 32:
        raise ValueError("obs must be uint8[P, max_obs, 32]")
    P, max_obs = obs.shape[0], obs.shape[1]
    if not 1 <= max_obs <= 16:
        raise ValueError("max_obs must be 1..16 (the reference keeps at most 10 observations)")
    counts = np.ascontiguousarray(counts, np.int32).ravel()
    if counts.shape[0] != P:
        raise ValueError("counts must be int[P]")
    out = np.full(P, -1, np.int32)
    if P:
        with eng._lock:
            rc = eng._lib.bfm_select_representative(eng._h, obs.ctypes.data, counts.ctypes.data, P, max_obs, out.ctypes.data)
            _ffi.check(eng._h, rc)
    return out
