"""boslam_b200 - B200-native ORB-descriptor Hamming matching for boslam's hot path.

Drop-in for the ``cv2.BFMatcher`` calls of reference ``slam/tracking.py:45,56,121`` (and the
batched keyframe-pair workloads of local mapping / loop closing).  The compute path is the CUDA
library ``libbfm_b200.so`` (C ABI: ``include/bfm.h``); importing this package never falls back to
a CPU implementation.
"""
from . import synth  # noqa: F401
from ._ffi import BfmError, LIB_PATH  # noqa: F401
from .bank import KeyframeBank  # noqa: F401
from .engine import BatchPlan, BatchResult, DevicePlan, Engine, HostBatchBuffers, PinnedBuffer, default_engine, make_problems  # noqa: F401
from .localmap import CameraModel, MapStore, TrackResult, quaternion_from_rotation, select_representative  # noqa: F401
from .matcher import NORM_HAMMING, BFMatcher, BFMatcher_create, DMatch, install, uninstall  # noqa: F401

__all__ = ["BFMatcher", "BFMatcher_create", "DMatch", "NORM_HAMMING", "Engine", "BatchPlan", "DevicePlan", "BatchResult", "PinnedBuffer", "HostBatchBuffers",
           "make_problems", "default_engine", "KeyframeBank", "MapStore", "CameraModel", "TrackResult", "select_representative", "match", "knn_match", "match_pairs", "BfmError", "synth", "install", "uninstall"]


def match(query, train, k=1, ratio=None, cross_check=False, mask=None, window=None, max_distance=None,
          strict=False, device=0):
    """Array form of the hot path: (queryIdx, trainIdx, distance).  See :meth:`Engine.match`."""
    return default_engine(device).match(query, train, k=k, ratio=ratio, cross_check=cross_check, mask=mask,
                                        window=window, max_distance=max_distance, strict=strict)


def knn_match(query, train, k=2, mask=None, window=None, device=0):
    """Dense k-NN table (idx[Q,k], dist[Q,k]).  See :meth:`Engine.knn`."""
    return default_engine(device).knn(query, train, k=k, mask=mask, window=window)


def match_pairs(queries, trains, device=0, **kw):
    """Batched keyframe-pair matching.  See :meth:`Engine.match_pairs`."""
    return default_engine(device).match_pairs(queries, trains, **kw)
