"""Host side of the matcher: a thin, typed layer over the C ABI (include/bfm.h).

``Engine`` owns one ``bfm_handle_t`` (own CUDA stream + workspace).  Inputs are either numpy
arrays (host path: the upload overlaps the call's single kernel launch - pinned arrays are streamed by the
kernel's own feeder CTAs, pageable ones are staged by a small host thread pool - and the kernel writes the
results into pinned host memory) or CUDA ``torch.Tensor`` s (device
path: zero-copy via ``data_ptr`` on torch's current stream).  All arithmetic happens in the CUDA
library; nothing here computes distances.

The array API mirrors what the reference's call sites do with cv2 (reference
``slam/tracking.py:56-60,119-126``): ``match`` -> (queryIdx, trainIdx, distance) in ascending
queryIdx order, with the caller-side filters (distance gate, ratio test) optionally fused.
"""
from __future__ import annotations

import ctypes
import threading
import weakref
from typing import Optional, Sequence

import numpy as np

from . import _ffi

DESC_BYTES = 32


def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


def _addr(a: np.ndarray) -> int:
    """Address of a numpy array's first byte (0.7 us; ``a.ctypes.data`` builds a helper object: 1.9 us)."""
    try:
        return ctypes.addressof(ctypes.c_char.from_buffer(a))
    except (TypeError, ValueError):      # read-only or empty arrays
        return a.ctypes.data


def _check_desc_np(a, name: str) -> np.ndarray:
    """cv2 raises on non-uint8 descriptors; so do we.  Non-contiguous is accepted."""
    a = np.asarray(a)
    if a.dtype != np.uint8:
        raise TypeError(f"{name}: descriptors must be uint8 (got {a.dtype}), as cv2.NORM_HAMMING requires")
    if a.ndim != 2 or a.shape[1] != DESC_BYTES:
        if a.ndim == 2 and a.shape[0] == 0:
            return np.zeros((0, DESC_BYTES), np.uint8)
        raise ValueError(f"{name}: expected uint8[N, {DESC_BYTES}] ORB descriptors, got shape {a.shape}")
    if not a.flags.c_contiguous or (a.ctypes.data & 15):
        a = np.ascontiguousarray(a)
        if a.ctypes.data & 15:  # pragma: no cover - numpy allocations are 16-byte aligned
            b = np.empty(a.shape[0] * DESC_BYTES + 16, np.uint8)
            off = (-b.ctypes.data) & 15
            b = b[off:off + a.size].reshape(a.shape)
            b[...] = a
            a = b
    return a


class _PinnedBlock:
    """Owns one cudaMallocHost allocation; freed when the last reference (buffer object or numpy view) dies."""

    def __init__(self, lib, nbytes: int):
        p = ctypes.c_void_p()
        _ffi.check(None, lib.bfm_host_alloc(max(int(nbytes), 1), ctypes.byref(p)))
        self.ptr = p.value
        self._fin = weakref.finalize(self, lib.bfm_host_free, ctypes.c_void_p(self.ptr))


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy array (``.array``).  Views of ``.array`` (for instance
    the slices a :class:`BatchResult` hands out) keep the allocation alive on their own."""

    def __init__(self, shape, dtype=np.uint8):
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        self._block = _PinnedBlock(_ffi.lib(), n)
        self._ptr = self._block.ptr
        raw = (ctypes.c_uint8 * max(n, 1)).from_address(self._ptr)
        raw._block = self._block  # numpy view -> ctypes array -> block: no view can outlive the memory
        self.array = np.frombuffer(raw, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        """Release now.  Only safe when no view of ``.array`` is used afterwards."""
        self.array = None
        self._block._fin()


class BatchResult:
    """Matches of a batch of problems, packed.  ``counts[p]`` matches of problem p start at
    ``offsets[p]`` in ``query_idx`` / ``train_idx`` / ``distance`` (problem-local indices; only the
    first ``counts[p]`` entries of a problem's slice are meaningful).  ``distance`` is kept as the
    integer Hamming distance; ``result[p]`` yields cv2-style float32 distances."""

    __slots__ = ("query_idx", "train_idx", "distance", "counts", "offsets")

    def __init__(self, query_idx, train_idx, distance, counts, offsets):
        self.query_idx, self.train_idx, self.distance = query_idx, train_idx, distance
        self.counts, self.offsets = counts, offsets

    def __len__(self):
        return len(self.counts)

    def __getitem__(self, p):
        b, n = int(self.offsets[p]), int(self.counts[p])
        return self.query_idx[b:b + n], self.train_idx[b:b + n], self.distance[b:b + n].astype(np.float32)

    def __iter__(self):
        for p in range(len(self)):
            yield self[p]


class HostBatchBuffers:
    """Reusable pinned output arrays for :meth:`Engine.match_batched` (``out=``): results are DMA-ed
    straight into them, so a steady-state batched call allocates nothing."""

    def __init__(self, n_out: int, n_problems: int, k: int = 1, want_knn: bool = False):
        self.n_out, self.n_problems, self.k = int(n_out), int(n_problems), int(k)
        self._m = PinnedBuffer((3, max(self.n_out, 1)), np.int32)
        self._c = PinnedBuffer((max(self.n_problems, 1),), np.int32)
        self.m, self.count = self._m.array, self._c.array
        self.knn_idx = self.knn_dist = None
        if want_knn:
            self._ki = PinnedBuffer((max(self.n_out, 1), self.k), np.int32)
            self._kd = PinnedBuffer((max(self.n_out, 1), self.k), np.int32)
            self.knn_idx, self.knn_dist = self._ki.array, self._kd.array


def make_problems(q_counts: Sequence[int], t_counts: Sequence[int], shared_query: bool = False) -> np.ndarray:
    """Problem table int32[P, 6] (q_begin, q_count, t_begin, t_count, out_begin, 0) for problems
    stored back to back.  With ``shared_query`` every problem reads query rows [0, q_counts[0])."""
    P = len(t_counts)
    tab = np.zeros((P, 6), np.int32)
    qc = np.asarray(q_counts, np.int64)
    tc = np.asarray(t_counts, np.int64)
    tab[:, 1] = qc
    tab[:, 3] = tc
    tab[:, 2] = np.concatenate([[0], np.cumsum(tc)[:-1]]) if P else 0
    out = np.concatenate([[0], np.cumsum(qc)[:-1]]) if P else np.zeros(0, np.int64)
    tab[:, 4] = out
    tab[:, 0] = 0 if shared_query else out
    return tab


class Engine:
    """One matcher instance = one CUDA stream + workspace on one B200."""

    def __init__(self, device: int = 0):
        self._lib = _ffi.lib()
        h = ctypes.c_void_p()
        rc = self._lib.bfm_create(int(device), ctypes.byref(h))
        if rc != _ffi.BFM_OK:
            _ffi.check(None, rc)
        self._h = h
        self.device = int(device)
        self._lock = threading.Lock()  # the C handle is not re-entrant
        self._fin = weakref.finalize(self, self._lib.bfm_destroy, h)
        self._opt_cache = {}           # option sets of the plain match call -> (struct, byref, gate nothing passes)
        self._cnt = np.zeros(4, np.int32)
        self._cnt_ptr = self._cnt.ctypes.data

    # -- housekeeping -------------------------------------------------------------------------
    def close(self):
        self._fin()

    def set_tuning(self, **knobs):
        for k, v in knobs.items():
            _ffi.check(self._h, self._lib.bfm_set_tuning(self._h, k.encode(), int(v)))

    def launch_info(self) -> dict:
        li = _ffi.LaunchInfo()
        _ffi.check(self._h, self._lib.bfm_get_launch_info(self._h, ctypes.byref(li)))
        return {f: getattr(li, f) for f, _ in li._fields_ if f != "reserved"}

    def kernel_launch_count(self) -> int:
        return int(self._lib.bfm_kernel_launch_count(self._h))

    # -- option marshalling ---------------------------------------------------------------------
    @staticmethod
    def _gate(max_distance, strict):
        if max_distance is None:
            return -1
        import math
        # distances are integers: d < x  <=>  d <= ceil(x) - 1 ;  d <= x  <=>  d <= floor(x)
        g = math.ceil(max_distance) - 1 if strict else math.floor(max_distance)
        return max(int(g), -1) if g >= 0 else -2  # -2: nothing can pass

    def _options(self, k, ratio, cross_check, max_distance, strict):
        if k < 1:
            raise ValueError("k must be >= 1")
        if cross_check and k != 1:
            raise ValueError("cross_check requires k == 1 (cv2 asserts the same, batch_distance.cpp:303)")
        if cross_check and ratio is not None:
            raise ValueError("cross_check and ratio are exclusive")
        if ratio is not None and k > 2:
            raise ValueError("the ratio test is defined on the two nearest neighbours: use k <= 2")
        o = _ffi.Options()
        o.k = int(k)
        o.cross_check = 1 if cross_check else 0
        o.mask_kind = _ffi.MASK_NONE
        o.ratio = float(ratio) if ratio is not None else -1.0
        g = self._gate(max_distance, strict)
        o.max_distance = g if g != -2 else 0
        return o, (g == -2)

    # -- core call --------------------------------------------------------------------------------
    def _call(self, mem, q_ptr, nq, t_ptr, nt, problems: np.ndarray, n_out, opts, knn, matches, stream=None):
        P = problems.shape[0]
        probs = np.ascontiguousarray(problems, np.int32)
        pp = probs.ctypes.data_as(ctypes.POINTER(_ffi.Problem))
        ki, kd = knn if knn is not None else (None, None)
        mq, mt, md, mc = matches if matches is not None else (None, None, None, None)
        with self._lock:
            rc = self._lib.bfm_match_batched(self._h, mem, q_ptr, nq, t_ptr, nt, pp, P, n_out, ctypes.byref(opts),
                                             ki, kd, mq, mt, md, mc, stream)
            _ffi.check(self._h, rc)

    # -- numpy (host) path --------------------------------------------------------------------------
    def _mask_args_np(self, opts, mask, window, nq, nt, keep):
        if mask is not None and window is not None:
            raise ValueError("give mask or window, not both")
        if mask is not None:
            m = np.asarray(mask)
            if m.dtype != np.uint8 or m.shape != (nq, nt):
                raise ValueError(f"mask must be uint8[{nq}, {nt}] (cv2 convention), got {m.dtype}{m.shape}")
            m = np.ascontiguousarray(m)
            keep.append(m)
            opts.mask_kind = _ffi.MASK_DENSE
            opts.mask = m.ctypes.data
            opts.mask_row_stride = nt
        if window is not None:
            q_xy, t_xy, radius = window
            q_xy = np.ascontiguousarray(q_xy, np.float32)
            t_xy = np.ascontiguousarray(t_xy, np.float32)
            if q_xy.shape != (nq, 2) or t_xy.shape != (nt, 2):
                raise ValueError("window = (q_xy float[Q,2], t_xy float[T,2], radius)")
            keep += [q_xy, t_xy]
            opts.mask_kind = _ffi.MASK_WINDOW
            opts.q_xy = q_xy.ctypes.data
            opts.t_xy = t_xy.ctypes.data
            opts.window_radius = float(radius)

    def knn(self, query, train, k: int = 1, mask=None, window=None):
        """cv2 ``knnMatch`` as arrays: (idx int32[Q,k], dist int32[Q,k]); -1 marks a missing neighbour."""
        if _is_torch(query):
            return self._knn_torch(query, train, k, mask, window)
        q = _check_desc_np(query, "query")
        t = _check_desc_np(train, "train")
        nq, nt = q.shape[0], t.shape[0]
        self._check_limits(nq, nt, k)
        idx = np.full((nq, k), -1, np.int32)
        dist = np.full((nq, k), -1, np.int32)
        if nq == 0 or nt == 0:
            return idx, dist
        opts, _ = self._options(k, None, False, None, False)
        keep = []
        self._mask_args_np(opts, mask, window, nq, nt, keep)
        probs = np.array([[0, nq, 0, nt, 0, 0]], np.int32)
        self._call(_ffi.MEM_HOST, q.ctypes.data, nq, t.ctypes.data, nt, probs, nq, opts,
                   (idx.ctypes.data, dist.ctypes.data), None)
        return idx, dist

    def match(self, query, train, k: int = 1, ratio: Optional[float] = None, cross_check: bool = False,
              mask=None, window=None, max_distance=None, strict: bool = False):
        """(queryIdx int32[M], trainIdx int32[M], distance float32[M]), ascending queryIdx.

        cross_check -> cv2 ``BFMatcher(crossCheck=True).match``; ratio -> Lowe test on the 2-NN;
        max_distance (+ strict) -> the gates of slam/tracking.py:121 (``<=``) and :57 (``<``).
        """
        if type(query) is np.ndarray and type(train) is np.ndarray and mask is None and window is None:
            # the call shape of the reference (slam/tracking.py:56,121): two plain descriptor arrays.  Everything that
            # does not depend on the data - options, the ctypes plumbing - is cached per option set: ~8 us of Python
            # instead of ~27
            q, t = query, train
            if (q.dtype == np.uint8 and t.dtype == np.uint8 and q.ndim == 2 and t.ndim == 2 and q.shape[1] == DESC_BYTES and
                    t.shape[1] == DESC_BYTES and q.strides == (DESC_BYTES, 1) and t.strides == (DESC_BYTES, 1)):
                nq, nt = q.shape[0], t.shape[0]
                key = (k, ratio, cross_check, max_distance, strict)
                hit = self._opt_cache.get(key)
                if hit is None:
                    self._check_limits(0, 0, k)
                    opts, none_pass = self._options(k, ratio, cross_check, max_distance, strict)
                    hit = self._opt_cache[key] = (opts, ctypes.byref(opts), none_pass)
                    if len(self._opt_cache) > 64:
                        self._opt_cache.clear()
                if nq == 0 or nt == 0 or hit[2]:
                    e = np.zeros(0, np.int32)
                    return e, e.copy(), np.zeros(0, np.float32)
                if nt >= _ffi.MAX_TRAIN_ROWS or nq >= _ffi.MAX_QUERY_ROWS:
                    self._check_limits(nq, nt, k)
                qa, ta = _addr(q), _addr(t)
                if not ((qa | ta) & 15):
                    buf = np.empty((3, nq), np.int32)
                    b = _addr(buf)
                    with self._lock:
                        rc = self._lib.bfm_match(self._h, _ffi.MEM_HOST, qa, nq, ta, nt, hit[1], b, b + 4 * nq, b + 8 * nq,
                                                 self._cnt_ptr, None)
                        if rc:
                            _ffi.check(self._h, rc)
                        n = int(self._cnt[0])
                    return buf[0, :n], buf[1, :n], buf[2, :n].astype(np.float32)
        if _is_torch(query):
            return self._match_torch(query, train, k, ratio, cross_check, mask, window, max_distance, strict)
        q = _check_desc_np(query, "query")
        t = _check_desc_np(train, "train")
        nq, nt = q.shape[0], t.shape[0]
        self._check_limits(nq, nt, k)
        opts, none_pass = self._options(k, ratio, cross_check, max_distance, strict)
        if nq == 0 or nt == 0 or none_pass:
            e = np.zeros(0, np.int32)
            return e, e.copy(), np.zeros(0, np.float32)
        keep = []
        self._mask_args_np(opts, mask, window, nq, nt, keep)
        mq = np.empty(nq, np.int32)
        mt = np.empty(nq, np.int32)
        md = np.empty(nq, np.int32)
        mc = np.zeros(1, np.int32)
        probs = np.array([[0, nq, 0, nt, 0, 0]], np.int32)
        self._call(_ffi.MEM_HOST, q.ctypes.data, nq, t.ctypes.data, nt, probs, nq, opts, None,
                   (mq.ctypes.data, mt.ctypes.data, md.ctypes.data, mc.ctypes.data))
        n = int(mc[0])
        return mq[:n], mt[:n], md[:n].astype(np.float32)

    @staticmethod
    def _check_limits(nq, nt, k):
        if nt >= _ffi.MAX_TRAIN_ROWS or nq >= _ffi.MAX_QUERY_ROWS:
            raise ValueError("at most 2^22 - 1 rows per problem (cv2 itself stops at 2^18 - 1 train rows)")
        if k > _ffi.MAX_K:
            raise NotImplementedError(f"k > {_ffi.MAX_K} is not supported by this build")

    # -- batched (keyframe pairs) -----------------------------------------------------------------------
    def match_batched(self, q_packed, t_packed, problems: np.ndarray, k: int = 1, ratio=None,
                      cross_check: bool = False, max_distance=None, strict: bool = False, window=None,
                      want_knn: bool = False, out: "HostBatchBuffers | None" = None):
        """Batched form over packed descriptor arrays + a problem table (see :func:`make_problems`).

        Returns a :class:`BatchResult` (and the dense (idx, dist) tables first if ``want_knn``).
        This is the local-mapping / loop-closing shape: P independent (query KF, train KF) problems
        in one launch.
        """
        if _is_torch(q_packed):
            return self._match_batched_torch(q_packed, t_packed, problems, k, ratio, cross_check, max_distance,
                                             strict, window, want_knn)
        q = _check_desc_np(q_packed, "q_packed")
        t = _check_desc_np(t_packed, "t_packed")
        probs = np.ascontiguousarray(problems, np.int32)
        if probs.ndim != 2 or probs.shape[1] != 6:
            raise ValueError("problems must be int32[P, 6]")
        P = probs.shape[0]
        n_out = int((probs[:, 4] + probs[:, 1]).max()) if P else 0
        if P and (int(probs[:, 3].max()) >= _ffi.MAX_TRAIN_ROWS or int(probs[:, 1].max()) >= _ffi.MAX_QUERY_ROWS):
            raise ValueError("at most 2^22 - 1 rows per problem")
        if k > _ffi.MAX_K:
            raise NotImplementedError(f"k > {_ffi.MAX_K} is not supported by this build")
        opts, none_pass = self._options(k, ratio, cross_check, max_distance, strict)
        keep = []
        if window is not None:
            self._mask_args_np(opts, None, window, q.shape[0], t.shape[0], keep)
        if out is not None:
            if out.n_out < n_out or out.n_problems < P or (want_knn and (out.knn_idx is None or out.k != k)):
                raise ValueError("out= buffers are too small for this batch")
            mq, mt, md, mc = out.m[0], out.m[1], out.m[2], out.count[:P]
            idx, dist = (out.knn_idx, out.knn_dist) if want_knn else (None, None)
        else:
            mq = np.empty(n_out, np.int32)
            mt = np.empty(n_out, np.int32)
            md = np.empty(n_out, np.int32)
            mc = np.zeros(P, np.int32)
            idx = np.full((n_out, k), -1, np.int32) if want_knn else None
            dist = np.full((n_out, k), -1, np.int32) if want_knn else None
        if P and n_out:
            self._call(_ffi.MEM_HOST, q.ctypes.data, q.shape[0], t.ctypes.data, t.shape[0], probs, n_out, opts,
                       (idx.ctypes.data, dist.ctypes.data) if want_knn else None,
                       (mq.ctypes.data, mt.ctypes.data, md.ctypes.data, mc.ctypes.data))
        if none_pass:
            mc[:] = 0
        res = BatchResult(mq, mt, md, mc, probs[:, 4].copy())
        return (idx, dist, res) if want_knn else res

    def plan_batch(self, problems: np.ndarray, k: int = 1, ratio=None, cross_check: bool = False, max_distance=None,
                   strict: bool = False) -> "BatchPlan":
        """Validate a problem table and its options once; :meth:`BatchPlan.run` is then the bare library call
        (for loops that match the same batch shape every step: loop closing, the bench's e2e leg)."""
        return BatchPlan(self, problems, k, ratio, cross_check, max_distance, strict)

    def plan_device(self, q, t, problems: np.ndarray, **kw) -> "DevicePlan":
        """Bind a device-resident batch (CUDA tensors) once; :meth:`DevicePlan.run` is then the bare library call."""
        return DevicePlan(self, q, t, problems, **kw)

    def match_pairs(self, queries: Sequence, trains: Sequence, **kw):
        """Convenience over :meth:`match_batched` for lists of per-keyframe arrays.  Identical query
        objects are stored once (loop closing: one keyframe against many candidates)."""
        if len(queries) != len(trains):
            raise ValueError("queries and trains must have the same length")
        P = len(trains)
        if P == 0:
            return BatchResult(*(np.zeros(0, np.int32),) * 5)
        qs = [_check_desc_np(a, "query") for a in queries]
        ts = [_check_desc_np(a, "train") for a in trains]
        tab = np.zeros((P, 6), np.int32)
        q_rows, q_seen, q_list = 0, {}, []
        for p, (a, src) in enumerate(zip(qs, queries)):
            key = id(src)
            if key not in q_seen:
                q_seen[key] = q_rows
                q_list.append(a)
                q_rows += a.shape[0]
            tab[p, 0], tab[p, 1] = q_seen[key], a.shape[0]
        tc = np.array([a.shape[0] for a in ts], np.int64)
        tab[:, 3] = tc
        tab[:, 2] = np.concatenate([[0], np.cumsum(tc)[:-1]])
        tab[:, 4] = np.concatenate([[0], np.cumsum(tab[:, 1].astype(np.int64))[:-1]])
        qp = np.concatenate(q_list) if q_list else np.zeros((0, DESC_BYTES), np.uint8)
        tp = np.concatenate(ts)
        return self.match_batched(qp, tp, tab, **kw)

    # -- torch (device) path --------------------------------------------------------------------------------
    def _torch_prep(self, x, name):
        import torch
        if x.dtype != torch.uint8:
            raise TypeError(f"{name}: descriptors must be uint8")
        if not x.is_cuda or x.device.index != self.device:
            raise ValueError(f"{name}: tensor must live on cuda:{self.device}")
        if x.dim() != 2 or x.shape[1] != DESC_BYTES:
            raise ValueError(f"{name}: expected uint8[N, {DESC_BYTES}]")
        x = x.contiguous()
        if x.data_ptr() & 15:
            x = x.clone()
        return x

    def _stream(self):
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _mask_args_torch(self, opts, mask, window, nq, nt, keep):
        import torch
        if mask is not None and window is not None:
            raise ValueError("give mask or window, not both")
        if mask is not None:
            if mask.dtype != torch.uint8 or tuple(mask.shape) != (nq, nt) or not mask.is_cuda:
                raise ValueError(f"mask must be a CUDA uint8[{nq}, {nt}] tensor")
            m = mask.contiguous()
            keep.append(m)
            opts.mask_kind, opts.mask, opts.mask_row_stride = _ffi.MASK_DENSE, m.data_ptr(), nt
        if window is not None:
            q_xy, t_xy, radius = window
            q_xy = q_xy.to(torch.float32).contiguous()
            t_xy = t_xy.to(torch.float32).contiguous()
            if tuple(q_xy.shape) != (nq, 2) or tuple(t_xy.shape) != (nt, 2) or not q_xy.is_cuda or not t_xy.is_cuda:
                raise ValueError("window = (q_xy float[Q,2], t_xy float[T,2], radius) as CUDA tensors")
            keep += [q_xy, t_xy]
            opts.mask_kind, opts.q_xy, opts.t_xy = _ffi.MASK_WINDOW, q_xy.data_ptr(), t_xy.data_ptr()
            opts.window_radius = float(radius)

    def _knn_torch(self, query, train, k, mask, window):
        import torch
        q = self._torch_prep(query, "query")
        t = self._torch_prep(train, "train")
        nq, nt = q.shape[0], t.shape[0]
        self._check_limits(nq, nt, k)
        idx = torch.full((nq, k), -1, dtype=torch.int32, device=q.device)
        dist = torch.full((nq, k), -1, dtype=torch.int32, device=q.device)
        if nq == 0 or nt == 0:
            return idx, dist
        opts, _ = self._options(k, None, False, None, False)
        keep = []
        self._mask_args_torch(opts, mask, window, nq, nt, keep)
        probs = np.array([[0, nq, 0, nt, 0, 0]], np.int32)
        self._call(_ffi.MEM_DEVICE, q.data_ptr(), nq, t.data_ptr(), nt, probs, nq, opts,
                   (idx.data_ptr(), dist.data_ptr()), None, self._stream())
        return idx, dist

    def _match_torch(self, query, train, k, ratio, cross_check, mask, window, max_distance, strict):
        import torch
        q = self._torch_prep(query, "query")
        t = self._torch_prep(train, "train")
        nq, nt = q.shape[0], t.shape[0]
        self._check_limits(nq, nt, k)
        opts, none_pass = self._options(k, ratio, cross_check, max_distance, strict)
        dev = q.device
        if nq == 0 or nt == 0 or none_pass:
            e = torch.zeros(0, dtype=torch.int32, device=dev)
            return e, e.clone(), torch.zeros(0, dtype=torch.float32, device=dev)
        keep = []
        self._mask_args_torch(opts, mask, window, nq, nt, keep)
        out = torch.empty((3, nq), dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        probs = np.array([[0, nq, 0, nt, 0, 0]], np.int32)
        self._call(_ffi.MEM_DEVICE, q.data_ptr(), nq, t.data_ptr(), nt, probs, nq, opts, None,
                   (out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), cnt.data_ptr()), self._stream())
        n = int(cnt.item())  # the one host sync of this call: the match count
        return out[0, :n], out[1, :n], out[2, :n].to(torch.float32)

    def match_batched_device(self, q, t, problems: np.ndarray, k=1, ratio=None, cross_check=False,
                             max_distance=None, strict=False, window=None, want_knn=False, out=None,
                             replicas=None):
        """Fully asynchronous device form: returns padded CUDA tensors, no host sync.  Queued on torch's current
        stream; calls on one engine run in issue order even across streams (the library chains them with events,
        include/bfm.h "streams"), because they share the engine's workspace.

        out = dict(m=int32[3, n_out], count=int32[P], knn_idx=..., knn_dist=...) may be passed to
        reuse buffers.  Matches of problem p sit at m[:, out_begin[p] : out_begin[p] + count[p]].

        replicas = up to 7 more dicts of the same layout whose tensors (or raw device pointers, as
        ints) live on NVLink peers: the kernel's epilogue writes every result there as well (the
        multi-GPU gather fused into the match, see :class:`boslam_b200.distributed.FusedGather`).
        """
        import torch
        q = self._torch_prep(q, "q_packed")
        t = self._torch_prep(t, "t_packed")
        probs = np.ascontiguousarray(problems, np.int32)
        P = probs.shape[0]
        n_out = int((probs[:, 4] + probs[:, 1]).max()) if P else 0
        if k > _ffi.MAX_K:
            raise NotImplementedError(f"k > {_ffi.MAX_K} is not supported by this build")
        opts, none_pass = self._options(k, ratio, cross_check, max_distance, strict)
        keep = []
        if window is not None:
            self._mask_args_torch(opts, None, window, q.shape[0], t.shape[0], keep)
        dev = q.device
        if out is None:
            out = {"m": torch.empty((3, max(n_out, 1)), dtype=torch.int32, device=dev),
                   "count": torch.zeros(max(P, 1), dtype=torch.int32, device=dev)}
            if want_knn:
                out["knn_idx"] = torch.empty((max(n_out, 1), k), dtype=torch.int32, device=dev)
                out["knn_dist"] = torch.empty((max(n_out, 1), k), dtype=torch.int32, device=dev)
        if P and n_out:
            if replicas or "m" not in out:
                replicas = replicas or []
                dests = (_ffi.Outputs * (1 + len(replicas)))()
                for d, o in zip(dests, [out] + list(replicas)):
                    self._fill_outputs(d, o, want_knn)
                pp = probs.ctypes.data_as(ctypes.POINTER(_ffi.Problem))
                with self._lock:
                    rc = self._lib.bfm_match_batched_multi(self._h, q.data_ptr(), q.shape[0], t.data_ptr(), t.shape[0],
                                                           pp, P, n_out, ctypes.byref(opts), dests, len(dests),
                                                           self._stream())
                    _ffi.check(self._h, rc)
            else:
                m = out["m"]
                self._call(_ffi.MEM_DEVICE, q.data_ptr(), q.shape[0], t.data_ptr(), t.shape[0], probs, n_out, opts,
                           (out["knn_idx"].data_ptr(), out["knn_dist"].data_ptr()) if want_knn else None,
                           (m[0].data_ptr(), m[1].data_ptr(), m[2].data_ptr(), out["count"].data_ptr()), self._stream())
            if none_pass:   # a gate nothing can pass (e.g. distance < 0): no matches, on the same stream
                for o in [out] + list(replicas or []):
                    if not isinstance(o["count"], int):
                        o["count"].zero_()
        return out

    def bind_device_multi(self, problems, dests, want_knn=False, k=1, ratio=None, cross_check=False, max_distance=None,
                          strict=False):
        """Marshal a device-resident multi-destination call once (problem table, options, destination structs - raw
        device pointers as :meth:`boslam_b200.distributed.FusedGather.destinations` builds them);
        :meth:`run_device_multi` then takes only the two descriptor tensors."""
        probs = np.ascontiguousarray(problems, np.int32)
        P = probs.shape[0]
        n_out = int((probs[:, 4] + probs[:, 1]).max()) if P else 0
        if k > _ffi.MAX_K:
            raise NotImplementedError(f"k > {_ffi.MAX_K} is not supported by this build")
        opts, none_pass = self._options(k, ratio, cross_check, max_distance, strict)
        if none_pass:
            raise ValueError("a gate nothing can pass has no bound form: use match_batched_device")
        arr = (_ffi.Outputs * len(dests))()
        for d, o in zip(arr, dests):
            self._fill_outputs(d, o, want_knn)
        return (probs, probs.ctypes.data_as(ctypes.POINTER(_ffi.Problem)), P, n_out, opts, ctypes.byref(opts), arr, len(dests))

    def run_device_multi(self, q, t, bound):
        _probs, pp, P, n_out, _opts, opts_ref, arr, n = bound
        if q.dtype != self._u8() or t.dtype != self._u8() or not q.is_cuda or not t.is_cuda or not q.is_contiguous() or not t.is_contiguous():
            raise ValueError("run_device_multi needs contiguous CUDA uint8[rows, 32] tensors")
        if P and n_out:
            with self._lock:
                rc = self._lib.bfm_match_batched_multi(self._h, q.data_ptr(), q.shape[0], t.data_ptr(), t.shape[0], pp, P, n_out,
                                                       opts_ref, arr, n, self._stream())
                if rc:
                    _ffi.check(self._h, rc)

    @staticmethod
    def _u8():
        import torch
        return torch.uint8

    @staticmethod
    def _fill_outputs(d, o, want_knn):
        """One bfm_outputs_t from a dict of tensors (m int32[3, n], count, knn_idx, knn_dist) or of raw
        device pointers (m_query, m_train, m_dist, count, knn_idx, knn_dist as ints)."""
        ptr = lambda x: int(x) if isinstance(x, int) else x.data_ptr()
        if "m" in o:
            m = o["m"]
            d.m_query, d.m_train, d.m_dist = m[0].data_ptr(), m[1].data_ptr(), m[2].data_ptr()
        else:
            d.m_query, d.m_train, d.m_dist = ptr(o["m_query"]), ptr(o["m_train"]), ptr(o["m_dist"])
        d.m_count = ptr(o["count"])
        d.multicast = 1 if o.get("multicast") else 0
        if want_knn:
            d.knn_idx, d.knn_dist = ptr(o["knn_idx"]), ptr(o["knn_dist"])

    def _match_batched_torch(self, q, t, problems, k, ratio, cross_check, max_distance, strict, window, want_knn):
        out = self.match_batched_device(q, t, problems, k, ratio, cross_check, max_distance, strict, window, want_knn)
        probs = np.ascontiguousarray(problems, np.int32)
        m = out["m"].cpu().numpy()
        cnt = out["count"].cpu().numpy()[:probs.shape[0]]
        res = BatchResult(m[0], m[1], m[2], cnt, probs[:, 4].copy())
        if want_knn:
            return out["knn_idx"], out["knn_dist"], res
        return res


class DevicePlan:
    """A device-resident call bound once: descriptor tensors, problem table, options and output tensors are validated
    and allocated at construction, ``run()`` is then the bare library call (one kernel launch on torch's current
    stream, no host synchronisation, a few microseconds of host time).  For loops that match the same shapes every
    step - a tracking loop against a resident keyframe, the bench's device-time legs.  Results: ``plan.out`` (the
    dict :meth:`Engine.match_batched_device` returns), valid once the stream has drained."""

    def __init__(self, engine: "Engine", q, t, problems, k=1, ratio=None, cross_check=False, max_distance=None,
                 strict=False, window=None, want_knn=False):
        import torch
        self.engine = engine
        self.q, self.t = engine._torch_prep(q, "q_packed"), engine._torch_prep(t, "t_packed")
        self.probs = np.ascontiguousarray(problems, np.int32)
        if self.probs.ndim != 2 or self.probs.shape[1] != 6:
            raise ValueError("problems must be int32[P, 6]")
        self.P = self.probs.shape[0]
        self.n_out = int((self.probs[:, 4] + self.probs[:, 1]).max()) if self.P else 0
        if k > _ffi.MAX_K:
            raise NotImplementedError(f"k > {_ffi.MAX_K} is not supported by this build")
        self.opts, self.none_pass = engine._options(k, ratio, cross_check, max_distance, strict)
        self._keep = []
        if window is not None:
            engine._mask_args_torch(self.opts, None, window, self.q.shape[0], self.t.shape[0], self._keep)
        dev = self.q.device
        self.out = {"m": torch.empty((3, max(self.n_out, 1)), dtype=torch.int32, device=dev),
                    "count": torch.zeros(max(self.P, 1), dtype=torch.int32, device=dev)}
        ki = kd = None
        if want_knn:
            self.out["knn_idx"] = torch.empty((max(self.n_out, 1), k), dtype=torch.int32, device=dev)
            self.out["knn_dist"] = torch.empty((max(self.n_out, 1), k), dtype=torch.int32, device=dev)
            ki, kd = self.out["knn_idx"].data_ptr(), self.out["knn_dist"].data_ptr()
        m = self.out["m"]
        self._args = (engine._h, _ffi.MEM_DEVICE, self.q.data_ptr(), self.q.shape[0], self.t.data_ptr(), self.t.shape[0],
                      self.probs.ctypes.data_as(ctypes.POINTER(_ffi.Problem)), self.P, self.n_out, ctypes.byref(self.opts),
                      ki, kd, m[0].data_ptr(), m[1].data_ptr(), m[2].data_ptr(), self.out["count"].data_ptr())
        self._fn = engine._lib.bfm_match_batched
        self._torch = torch

    def run(self, stream=None):
        """``stream``: a raw cudaStream_t (int) to skip the current-stream lookup; default torch's current stream."""
        if self.P and self.n_out:
            eng = self.engine
            st = ctypes.c_void_p(stream if stream is not None else self._torch.cuda.current_stream(eng.device).cuda_stream)
            with eng._lock:
                rc = self._fn(*self._args, st)
                if rc:
                    _ffi.check(eng._h, rc)
            if self.none_pass:
                self.out["count"].zero_()
        return self.out


class BatchPlan:
    """A validated (problem table, options) pair bound to an engine.  ``run(q, t, out)`` takes C-contiguous
    uint8[rows, 32] host arrays and a :class:`HostBatchBuffers`; everything else was checked at construction."""

    def __init__(self, engine: Engine, problems, k, ratio, cross_check, max_distance, strict):
        probs = np.ascontiguousarray(problems, np.int32)
        if probs.ndim != 2 or probs.shape[1] != 6:
            raise ValueError("problems must be int32[P, 6]")
        if k > 2:
            raise NotImplementedError("BatchPlan covers the match-list form (k <= 2)")
        self.engine, self.probs, self.P = engine, probs, probs.shape[0]
        self.n_out = int((probs[:, 4] + probs[:, 1]).max()) if self.P else 0
        self.nq_min = int((probs[:, 0] + probs[:, 1]).max()) if self.P else 0
        self.nt_min = int((probs[:, 2] + probs[:, 3]).max()) if self.P else 0
        self.opts, self.none_pass = engine._options(k, ratio, cross_check, max_distance, strict)
        self.k = k
        self._pp = probs.ctypes.data_as(ctypes.POINTER(_ffi.Problem))
        self._opts_ref = ctypes.byref(self.opts)
        self._offsets = probs[:, 4].copy()

    def run(self, q: np.ndarray, t: np.ndarray, out: HostBatchBuffers, replicas=None) -> BatchResult:
        """``replicas``: up to 7 dicts of raw DEVICE pointers (``m_query, m_train, m_dist, count`` [+ ``multicast``], as
        :meth:`boslam_b200.distributed.FusedGather.destinations` builds them): the epilogue that writes ``out`` (host)
        also writes every result there - host copies, match and the multi-GPU exchange in one call."""
        if (q.dtype != np.uint8 or t.dtype != np.uint8 or not q.flags.c_contiguous or not t.flags.c_contiguous or
                q.ndim != 2 or t.ndim != 2 or q.shape[1] != DESC_BYTES or t.shape[1] != DESC_BYTES or
                q.shape[0] < self.nq_min or t.shape[0] < self.nt_min or (q.ctypes.data | t.ctypes.data) & 15):
            raise ValueError("BatchPlan.run needs C-contiguous, 16-byte aligned uint8[rows, 32] arrays covering the planned rows")
        if out.n_out < self.n_out or out.n_problems < self.P:
            raise ValueError("out= buffers are too small for this batch")
        eng = self.engine
        if self.P and self.n_out:
            m = out.m
            if replicas:
                host = _ffi.Outputs()
                host.m_query, host.m_train, host.m_dist = m[0].ctypes.data, m[1].ctypes.data, m[2].ctypes.data
                host.m_count = out.count.ctypes.data
                dests = (_ffi.Outputs * len(replicas))()
                for d, o in zip(dests, replicas):
                    Engine._fill_outputs(d, o, False)
                with eng._lock:
                    rc = eng._lib.bfm_match_batched_host_multi(eng._h, q.ctypes.data, q.shape[0], t.ctypes.data, t.shape[0],
                                                               self._pp, self.P, self.n_out, self._opts_ref,
                                                               ctypes.byref(host), dests, len(replicas))
                    _ffi.check(eng._h, rc)
            else:
                with eng._lock:
                    rc = eng._lib.bfm_match_batched(eng._h, _ffi.MEM_HOST, q.ctypes.data, q.shape[0], t.ctypes.data, t.shape[0],
                                                    self._pp, self.P, self.n_out, self._opts_ref, None, None, m[0].ctypes.data,
                                                    m[1].ctypes.data, m[2].ctypes.data, out.count.ctypes.data, None)
                    _ffi.check(eng._h, rc)
        cnt = out.count[:self.P]
        if self.none_pass:
            cnt[:] = 0
        return BatchResult(out.m[0], out.m[1], out.m[2], cnt, self._offsets)

    def bind_replicas(self, out: HostBatchBuffers, replicas):
        """Marshal the destinations of :meth:`run` with ``replicas`` once (the structs of a sharded step's slot never
        change); :meth:`run_bound` is then the bare library call."""
        if out.n_out < self.n_out or out.n_problems < self.P:
            raise ValueError("out= buffers are too small for this batch")
        m = out.m
        host = _ffi.Outputs()
        host.m_query, host.m_train, host.m_dist = m[0].ctypes.data, m[1].ctypes.data, m[2].ctypes.data
        host.m_count = out.count.ctypes.data
        dests = (_ffi.Outputs * len(replicas))()
        for d, o in zip(dests, replicas):
            Engine._fill_outputs(d, o, False)
        return (out, host, ctypes.byref(host), dests, len(replicas))

    def run_bound(self, q: np.ndarray, t: np.ndarray, bound) -> BatchResult:
        out, _host, host_ref, dests, n = bound
        if (q.dtype != np.uint8 or t.dtype != np.uint8 or q.strides != (DESC_BYTES, 1) or t.strides != (DESC_BYTES, 1) or
                q.shape[0] < self.nq_min or t.shape[0] < self.nt_min):
            raise ValueError("BatchPlan.run_bound needs C-contiguous uint8[rows, 32] arrays covering the planned rows")
        qa, ta = _addr(q), _addr(t)
        if (qa | ta) & 15:
            raise ValueError("descriptor arrays must be 16-byte aligned")
        eng = self.engine
        if self.P and self.n_out:
            with eng._lock:
                rc = eng._lib.bfm_match_batched_host_multi(eng._h, qa, q.shape[0], ta, t.shape[0], self._pp, self.P, self.n_out,
                                                           self._opts_ref, host_ref, dests, n)
                if rc:
                    _ffi.check(eng._h, rc)
        cnt = out.count[:self.P]
        if self.none_pass:
            cnt[:] = 0
        return BatchResult(out.m[0], out.m[1], out.m[2], cnt, self._offsets)


_default_engines = {}
_default_lock = threading.Lock()


def default_engine(device: int = 0) -> Engine:
    """Per-(thread, device) engine so concurrent callers never share a handle.  Engines of threads that have
    exited are closed (stream, pinned buffers, worker threads) the next time a new one is created."""
    key = (threading.get_ident(), int(device))
    with _default_lock:
        hit = _default_engines.get(key)
        if hit is not None and hit[1].is_alive():
            return hit[0]
        for k, (eng, th) in list(_default_engines.items()):
            if not th.is_alive() or k == key:   # thread idents are reused: a dead thread's entry must go first
                del _default_engines[k]
                eng.close()
        e = Engine(device)
        _default_engines[key] = (e, threading.current_thread())
        return e
