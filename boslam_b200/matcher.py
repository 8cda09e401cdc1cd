"""Drop-in for the ``cv2.BFMatcher`` object boslam holds (reference ``slam/tracking.py:45``,
``slam/local_mapping.py:21``, ``slam/covisibility_graph.py:34``, ``experiments/pnp_*_tracking.py:11``).

    matcher = BFMatcher_create(NORM_HAMMING, crossCheck=True)        # slam/tracking.py:45
    matches = matcher.match(frame.des, kf_ref.desf())                 # slam/tracking.py:56
    matches = [m for m in matches if m.distance < d_hamming_max]      # slam/tracking.py:57

Same constructor arguments, same ``match`` / ``knnMatch`` signatures, same result objects
(``DMatch`` with ``queryIdx, trainIdx, imgIdx, distance``), same errors in the
same places (non-uint8 input, crossCheck with k > 1 or with a mask ... except that this engine
accepts crossCheck + mask, which cv2 refuses).  The work happens in the CUDA library.
"""
from __future__ import annotations

from itertools import repeat

import numpy as np

from .engine import Engine

NORM_HAMMING = 6  # == cv2.NORM_HAMMING
NORM_HAMMING2 = 7


class DMatch:
    """Mirror of ``cv2.DMatch``: the four attributes the reference reads
    (``slam/tracking.py:57,60,121,126``)."""

    __slots__ = ("queryIdx", "trainIdx", "imgIdx", "distance")

    def __init__(self, queryIdx=-1, trainIdx=-1, imgIdx=-1, distance=float("inf")):
        self.queryIdx = queryIdx
        self.trainIdx = trainIdx
        self.imgIdx = imgIdx
        self.distance = distance

    def __lt__(self, other):  # cv2.DMatch orders by distance
        return self.distance < other.distance

    def __repr__(self):
        return f"DMatch(queryIdx={self.queryIdx}, trainIdx={self.trainIdx}, imgIdx={self.imgIdx}, distance={self.distance})"


def _dmatches(qi, ti, d):
    return tuple(map(DMatch, qi.tolist(), ti.tolist(), repeat(0), d.tolist()))


class BFMatcher:
    """``cv2.BFMatcher`` for ``NORM_HAMMING`` on a B200."""

    def __init__(self, normType: int = NORM_HAMMING, crossCheck: bool = False, device: int = 0):
        if normType != NORM_HAMMING:
            raise ValueError("boslam_b200.BFMatcher implements NORM_HAMMING only (the norm boslam uses)")
        self.normType = normType
        self.crossCheck = bool(crossCheck)
        self._device = device
        self._engine = None
        self._train = []  # DescriptorMatcher.add() collection
        self._train_dev = None  # (device tensor of the concatenated collection, image bounds), built on first use

    # engine is created lazily so constructing a matcher (as covisibility_graph.py:34 does and never
    # uses) costs nothing
    @property
    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self._device)
        return self._engine

    # -- cv2.DescriptorMatcher collection API ---------------------------------------------------------
    def add(self, descriptors):
        self._train.extend(list(descriptors))
        self._train_dev = None

    def clear(self):
        self._train = []
        self._train_dev = None

    def train(self):
        """cv2: "trains" the matcher; for a brute-force matcher that means nothing on the CPU.  Here it
        uploads the collection so the following match calls move only the query (the collection stays
        resident on the GPU until add() / clear() - SURVEY 8(f) row 2)."""
        self._resident()

    def empty(self):
        return len(self._train) == 0

    def getTrainDescriptors(self):
        return list(self._train)

    def isMaskSupported(self):
        return True

    def _resident(self):
        if self._train_dev is None and self._train:
            import torch
            sizes = [len(a) for a in self._train]
            host = np.ascontiguousarray(np.concatenate([np.asarray(a) for a in self._train]))
            if host.dtype != np.uint8:
                raise TypeError("descriptors must be uint8, as cv2.NORM_HAMMING requires")
            dev = torch.from_numpy(host).to(torch.device("cuda", self._device))
            self._train_dev = (dev, np.cumsum([0] + sizes))
        return self._train_dev

    def _resolve_train(self, trainDescriptors):
        if trainDescriptors is not None:
            return trainDescriptors, None
        if not self._train:
            return np.zeros((0, 32), np.uint8), None
        return self._resident()

    def _to_engine(self, query, train):
        """A resident (device) train set needs the query on the device too; results come back as numpy."""
        if type(train).__module__.split(".")[0] != "torch":
            return query, train, (lambda x: x)
        import torch
        q = np.ascontiguousarray(query)
        if q.dtype != np.uint8:
            raise TypeError("descriptors must be uint8, as cv2.NORM_HAMMING requires")
        if q.ndim != 2 or q.shape[1] != 32:
            q = q.reshape(-1, 32) if q.size == 0 else q
        qd = torch.from_numpy(q).to(train.device)
        return qd, train, (lambda x: x.cpu().numpy())

    @staticmethod
    def _img_index(ti, bounds):
        img = np.searchsorted(bounds, ti, side="right") - 1
        return img, ti - bounds[img]

    # -- matching ----------------------------------------------------------------------------------------
    def match(self, queryDescriptors, trainDescriptors=None, mask=None):
        """tuple[DMatch], ascending queryIdx; queries without a candidate are omitted (rule R3)."""
        train, bounds = self._resolve_train(trainDescriptors)
        query, train, back = self._to_engine(queryDescriptors, train)
        if mask is not None and bounds is not None:
            import torch
            mask = torch.from_numpy(np.ascontiguousarray(mask)).to(train.device)
        qi, ti, d = (back(x) for x in self.engine.match(query, train, k=1, cross_check=self.crossCheck, mask=mask))
        if bounds is None:
            return _dmatches(qi, ti, d)
        img, loc = self._img_index(ti, bounds)
        return tuple(map(DMatch, qi.tolist(), loc.tolist(), img.tolist(), d.tolist()))

    def knnMatch(self, queryDescriptors, trainDescriptors=None, k=None, mask=None, compactResult=False):
        """tuple[tuple[DMatch]]: one row per query (empty rows kept unless compactResult)."""
        if k is None:
            raise TypeError("knnMatch() missing required argument 'k'")
        train, bounds = self._resolve_train(trainDescriptors)
        queryDescriptors_in = queryDescriptors
        queryDescriptors, train, back = self._to_engine(queryDescriptors, train)
        if mask is not None and bounds is not None:
            import torch
            mask = torch.from_numpy(np.ascontiguousarray(mask)).to(train.device)
        if self.crossCheck:
            if k != 1:
                raise ValueError("crossCheck=True requires k == 1 (cv2: batch_distance.cpp:303 assertion)")
            qi, ti, d = (back(x) for x in self.engine.match(queryDescriptors, train, k=1, cross_check=True, mask=mask))
            queryDescriptors = queryDescriptors_in
            nq = len(queryDescriptors)
            rows = [()] * nq
            if bounds is None:
                img_l, loc_l = repeat(0), ti.tolist()
            else:   # add()/train() collection: per-image trainIdx + imgIdx, as match() and the plain knn path
                img, loc = self._img_index(ti, bounds)
                img_l, loc_l = img.tolist(), loc.tolist()
            for a, b, im, c in zip(qi.tolist(), loc_l, img_l, d.tolist()):
                rows[a] = (DMatch(a, b, im, c),)
            if compactResult:
                rows = [r for r in rows if r]
            return tuple(rows)
        idx, dist = (back(x) for x in self.engine.knn(queryDescriptors, train, k=k, mask=mask))
        idx_l, dist_l = idx.tolist(), dist.tolist()
        if bounds is not None:
            img_a, loc_a = self._img_index(np.maximum(idx, 0), bounds)
            img_l, loc_l = img_a.tolist(), loc_a.tolist()
        rows = []
        for i, (ri, rd) in enumerate(zip(idx_l, dist_l)):
            if bounds is None:
                row = tuple(DMatch(i, j, 0, float(dd)) for j, dd in zip(ri, rd) if j >= 0)
            else:
                row = tuple(DMatch(i, loc_l[i][c], img_l[i][c], float(rd[c])) for c, j in enumerate(ri) if j >= 0)
            if row or not compactResult:
                rows.append(row)
        return tuple(rows)


def BFMatcher_create(normType: int = NORM_HAMMING, crossCheck: bool = False, device: int = 0) -> BFMatcher:
    """Same spelling as ``cv2.BFMatcher_create`` (slam/tracking.py:45)."""
    return BFMatcher(normType, crossCheck, device)


def install(cv2_module, device: int = 0):
    """Route ``cv2_module.BFMatcher_create`` / ``cv2_module.BFMatcher`` to this engine for ``NORM_HAMMING`` (other
    norms keep going to OpenCV).  boslam builds its matchers with exactly that call (``slam/tracking.py:45``,
    ``slam/local_mapping.py:21``, ``slam/covisibility_graph.py:34``), so two lines at the top of ``slam/main.py``
    switch the whole program over without touching the call sites::

        import boslam_b200
        boslam_b200.install(cv2)          # cv2 being the module slam/main.py already imported

    The module is passed in (this package never imports cv2 itself).  :func:`uninstall` restores the originals."""
    if getattr(cv2_module, "_boslam_b200_originals", None) is not None:
        return
    originals = (cv2_module.BFMatcher_create, cv2_module.BFMatcher)

    def _create(normType=4, crossCheck=False):  # cv2's default norm is NORM_L2 (4)
        if normType == NORM_HAMMING:
            return BFMatcher(normType, crossCheck, device)
        return originals[0](normType, crossCheck)

    class _BFMatcherDispatch:
        def __new__(cls, normType=4, crossCheck=False):
            if normType == NORM_HAMMING:
                return BFMatcher(normType, crossCheck, device)
            return originals[1](normType, crossCheck)

        create = staticmethod(_create)

    cv2_module._boslam_b200_originals = originals
    cv2_module.BFMatcher_create = _create
    cv2_module.BFMatcher = _BFMatcherDispatch


def uninstall(cv2_module):
    originals = getattr(cv2_module, "_boslam_b200_originals", None)
    if originals is not None:
        cv2_module.BFMatcher_create, cv2_module.BFMatcher = originals
        cv2_module._boslam_b200_originals = None
