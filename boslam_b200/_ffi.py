"""ctypes binding of libbfm_b200.so (the C ABI in include/bfm.h).

There is no CPU implementation behind this module: if the CUDA library is missing, or no B200 is
visible, the calls raise.  (The CPU oracle lives in ``oracle/`` and is test infrastructure only.)
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BFM_LIB_PATH") or os.path.join(_HERE, "libbfm_b200.so")  # the override is for A/B builds of the kernels

BFM_OK, BFM_ERR_INVALID, BFM_ERR_CUDA, BFM_ERR_NOMEM, BFM_ERR_UNSUPPORTED = range(5)
MEM_HOST, MEM_DEVICE = 0, 1
MASK_NONE, MASK_DENSE, MASK_WINDOW = 0, 1, 2
MAX_TRAIN_ROWS = 1 << 22
MAX_QUERY_ROWS = 1 << 22
MAX_K = 16
ABI_VERSION = 5

# every symbol include/bfm.h declares; tests/test_abi.py checks the library exports all of them
EXPORTED_SYMBOLS = (
    "bfm_abi_version", "bfm_create", "bfm_destroy", "bfm_last_error", "bfm_match_batched", "bfm_knn",
    "bfm_match", "bfm_match_batched_multi", "bfm_match_batched_host_multi", "bfm_get_launch_info", "bfm_set_tuning", "bfm_kernel_launch_count", "bfm_microbench",
    "bfm_device_info", "bfm_host_alloc", "bfm_host_free", "bfm_map_create", "bfm_map_destroy", "bfm_map_update",
    "bfm_track_local_map", "bfm_select_representative", "bfm_plan_preview", "bfm_debug_timeline", "bfm_plan_preview_tiles", "bfm_keyframe_vote", "bfm_synchronize",
    "bfm_plan_preview_tensor", "bfm_plan_preview_host_chunks",
)


class BfmError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libbfm_b200 error {code}: {message}")
        self.code = code


class Problem(ctypes.Structure):
    _fields_ = [("q_begin", ctypes.c_int32), ("q_count", ctypes.c_int32), ("t_begin", ctypes.c_int32),
                ("t_count", ctypes.c_int32), ("out_begin", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class Options(ctypes.Structure):
    _fields_ = [("k", ctypes.c_int32), ("cross_check", ctypes.c_int32), ("mask_kind", ctypes.c_int32),
                ("max_distance", ctypes.c_int32), ("ratio", ctypes.c_double), ("window_radius", ctypes.c_float),
                ("reserved0", ctypes.c_int32), ("mask", ctypes.c_void_p), ("mask_row_stride", ctypes.c_int64),
                ("q_xy", ctypes.c_void_p), ("t_xy", ctypes.c_void_p)]


class Outputs(ctypes.Structure):
    _fields_ = [("knn_idx", ctypes.c_void_p), ("knn_dist", ctypes.c_void_p), ("m_query", ctypes.c_void_p),
                ("m_train", ctypes.c_void_p), ("m_dist", ctypes.c_void_p), ("m_count", ctypes.c_void_p),
                ("multicast", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class TrackParams(ctypes.Structure):
    _fields_ = [("q", ctypes.c_double * 4), ("t", ctypes.c_double * 3), ("see_vector", ctypes.c_double * 3),
                ("fx", ctypes.c_double), ("fy", ctypes.c_double), ("cx", ctypes.c_double), ("cy", ctypes.c_double),
                ("cos_max", ctypes.c_double), ("width", ctypes.c_int32), ("height", ctypes.c_int32)]


class LaunchInfo(ctypes.Structure):
    _fields_ = [("kernels_launched", ctypes.c_int32), ("scan_grid", ctypes.c_int32), ("scan_block", ctypes.c_int32),
                ("queries_per_thread", ctypes.c_int32), ("popc_mode", ctypes.c_int32), ("segments", ctypes.c_int32),
                ("train_rows_per_segment", ctypes.c_int32), ("copy_chunks", ctypes.c_int32),
                ("scan_ms", ctypes.c_float), ("total_ms", ctypes.c_float)]


_lib = None
_lib_lock = threading.Lock()


def lib():
    """Load the library once.  Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C boslam_b200/csrc` (nvcc, sm_100a). boslam_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
        L.bfm_abi_version.restype = ctypes.c_int
        L.bfm_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
        L.bfm_destroy.argtypes = [vp]
        L.bfm_last_error.argtypes = [vp]
        L.bfm_last_error.restype = ctypes.c_char_p
        L.bfm_match_batched.argtypes = [vp, ctypes.c_int, vp, i32, vp, i32, ctypes.POINTER(Problem), i32, i32,
                                        ctypes.POINTER(Options), vp, vp, vp, vp, vp, vp, vp]
        L.bfm_match_batched_multi.argtypes = [vp, vp, i32, vp, i32, ctypes.POINTER(Problem), i32, i32,
                                              ctypes.POINTER(Options), ctypes.POINTER(Outputs), i32, vp]
        L.bfm_match_batched_host_multi.argtypes = [vp, vp, i32, vp, i32, ctypes.POINTER(Problem), i32, i32,
                                                   ctypes.POINTER(Options), ctypes.POINTER(Outputs),
                                                   ctypes.POINTER(Outputs), i32]
        L.bfm_map_create.argtypes = [vp, i32, ctypes.POINTER(vp)]
        L.bfm_map_destroy.argtypes = [vp]
        L.bfm_map_update.argtypes = [vp, i32, vp, vp, vp, vp]
        L.bfm_track_local_map.argtypes = [vp, ctypes.POINTER(TrackParams), vp, i32, vp, vp, i32, ctypes.POINTER(Options),
                                          vp, vp, vp, vp, vp, vp, vp, vp, ctypes.POINTER(i32), ctypes.POINTER(i32)]
        L.bfm_select_representative.argtypes = [vp, vp, vp, i32, i32, vp]
        L.bfm_keyframe_vote.argtypes = [vp, vp, i32, vp, i32, i32, vp, vp, vp]
        L.bfm_synchronize.argtypes = [vp]
        L.bfm_knn.argtypes = [vp, ctypes.c_int, vp, i32, vp, i32, ctypes.POINTER(Options), vp, vp, vp]
        L.bfm_match.argtypes = [vp, ctypes.c_int, vp, i32, vp, i32, ctypes.POINTER(Options), vp, vp, vp, vp, vp]
        L.bfm_get_launch_info.argtypes = [vp, ctypes.POINTER(LaunchInfo)]
        L.bfm_set_tuning.argtypes = [vp, ctypes.c_char_p, i32]
        L.bfm_kernel_launch_count.argtypes = [vp]
        L.bfm_kernel_launch_count.restype = i64
        L.bfm_debug_timeline.argtypes = [vp, vp, i32]
        L.bfm_plan_preview_tiles.argtypes = [ctypes.POINTER(Problem), i32, i32, vp, vp, i32, vp]
        L.bfm_microbench.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                     ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        L.bfm_device_info.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                      ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.c_char_p,
                                      ctypes.c_int]
        L.bfm_host_alloc.argtypes = [ctypes.c_uint64, ctypes.POINTER(vp)]
        L.bfm_host_free.argtypes = [vp]
        L.bfm_plan_preview.argtypes = [ctypes.POINTER(Problem), i32, i32, i32, i32, i32, i32, i32, vp, i32,
                                       ctypes.POINTER(i32), ctypes.POINTER(i32)]
        if L.bfm_abi_version() != ABI_VERSION:
            raise ImportError(f"{LIB_PATH}: ABI version {L.bfm_abi_version()} != {ABI_VERSION}; rebuild")
        _lib = L
    return _lib


def check(handle, rc: int):
    if rc != BFM_OK:
        msg = lib().bfm_last_error(handle)
        raise BfmError(rc, msg.decode() if msg else "unknown error")


MICROBENCH_TESTS = {
    "popc": 0, "lop3": 1, "iadd": 2, "popc+lop3": 3, "popc+2lop3": 4, "redux_min": 5, "imad": 6,
    "vimnmx": 7, "popc+imad": 8, "pair_mix_popc": 9,
}


def microbench(device: int = 0, iters: int = 2000, tests=None):
    """Integer-pipe issue rates (thread ops / clk / SM, ops / s, implied SM MHz) per probe."""
    L = lib()
    out = {}
    for name, t in MICROBENCH_TESTS.items():
        if tests is not None and name not in tests:
            continue
        a, b, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        rc = L.bfm_microbench(device, t, iters, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        if rc != BFM_OK:
            raise BfmError(rc, f"bfm_microbench({name}) failed")
        out[name] = {"ops_per_clk_per_sm": a.value, "ops_per_s": b.value, "sm_mhz": c.value}
    return out


def device_info(device: int = 0):
    L = lib()
    sm, ma, mi, khz = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    name = ctypes.create_string_buffer(128)
    rc = L.bfm_device_info(device, ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi), ctypes.byref(khz), name, 128)
    if rc != BFM_OK:
        raise BfmError(rc, "bfm_device_info failed (no CUDA device?)")
    return {"sm_count": sm.value, "cc": (ma.value, mi.value), "clock_khz": khz.value, "name": name.value.decode()}


def plan_preview(problems, queries_per_thread: int = 4, slots: int = 148 * 8, segment_rows: int = 0, waves: int = 0,
                 taper: int = 0, taper_pct: int = 0):
    """The work items (one per CTA) the planner cuts a batch into - host only, no GPU needed.
    ``problems``: int32[P, 6] as :func:`boslam_b200.make_problems` builds it.  Returns (items int32[n, 8] =
    {q_row0, q_valid, q_local0, out_row0, t_row0, t_count, t_local0, problem}, segment_rows)."""
    import numpy as np
    L = lib()
    tab = np.ascontiguousarray(problems, dtype=np.int32).reshape(-1, 6)
    pp = tab.ctypes.data_as(ctypes.POINTER(Problem))
    n, rows = ctypes.c_int32(), ctypes.c_int32()
    args = (pp, len(tab), queries_per_thread, slots, segment_rows, waves, taper, taper_pct)
    rc = L.bfm_plan_preview(*args, None, 0, ctypes.byref(n), ctypes.byref(rows))
    if rc != BFM_OK:
        raise BfmError(rc, "bfm_plan_preview: invalid arguments")
    items = np.empty((n.value, 8), np.int32)
    rc = L.bfm_plan_preview(*args, items.ctypes.data, n.value, ctypes.byref(n), ctypes.byref(rows))
    if rc != BFM_OK:
        raise BfmError(rc, "bfm_plan_preview failed")
    return items, rows.value


def plan_preview_tiles(problems, n_ctas: int = 148 * 8):
    """Finalize tiles of the persistent form - host only.  Returns (tiles int32[m, 5] = {problem, row0, tile number,
    tiles of the problem, look-back slot of tile 0}, tile_cta int32[m])."""
    import numpy as np
    L = lib()
    tab = np.ascontiguousarray(problems, dtype=np.int32).reshape(-1, 6)
    pp = tab.ctypes.data_as(ctypes.POINTER(Problem))
    m = ctypes.c_int32()
    rc = L.bfm_plan_preview_tiles(pp, len(tab), n_ctas, None, None, 0, ctypes.byref(m))
    if rc != BFM_OK:
        raise BfmError(rc, "bfm_plan_preview_tiles: invalid arguments")
    tiles, tile_cta = np.empty((m.value, 5), np.int32), np.empty(m.value, np.int32)
    rc = L.bfm_plan_preview_tiles(pp, len(tab), n_ctas, tiles.ctypes.data, tile_cta.ctypes.data, m.value, ctypes.byref(m))
    if rc != BFM_OK:
        raise BfmError(rc, "bfm_plan_preview_tiles failed")
    return tiles, tile_cta


def plan_preview_tensor(problems, n_sms: int = 148):
    """Work items of the tensor form - host only.  Returns (items int32[n, 8] as in :func:`plan_preview`, but with
    q_row0 / t_row0 in rows of the expanded planes; train rows per item; (query plane rows, train plane rows))."""
    import numpy as np
    L = lib()
    tab = np.ascontiguousarray(problems, dtype=np.int32).reshape(-1, 6)
    pp = tab.ctypes.data_as(ctypes.POINTER(Problem))
    n, rows = ctypes.c_int32(), ctypes.c_int32()
    planes = (ctypes.c_int32 * 2)()
    L.bfm_plan_preview_tensor.argtypes = [ctypes.POINTER(Problem), ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    rc = L.bfm_plan_preview_tensor(pp, len(tab), n_sms, None, 0, ctypes.byref(n), ctypes.byref(rows), planes)
    if rc != BFM_OK:
        raise BfmError(rc, "bfm_plan_preview_tensor: invalid arguments")
    items = np.empty((n.value, 8), np.int32)
    rc = L.bfm_plan_preview_tensor(pp, len(tab), n_sms, items.ctypes.data, n.value, ctypes.byref(n), ctypes.byref(rows), planes)
    if rc != BFM_OK:
        raise BfmError(rc, "bfm_plan_preview_tensor failed")
    return items, rows.value, (planes[0], planes[1])


def plan_preview_host_chunks(problems, n_query_rows: int, n_train_rows: int, n_sms: int = 148, forced: int = 0):
    """(number of copy chunks, problems per chunk) of a host batch that takes the tensor form - host only."""
    import numpy as np
    L = lib()
    tab = np.ascontiguousarray(problems, dtype=np.int32).reshape(-1, 6)
    pp = tab.ctypes.data_as(ctypes.POINTER(Problem))
    c, per = ctypes.c_int32(), ctypes.c_int32()
    L.bfm_plan_preview_host_chunks.argtypes = [ctypes.POINTER(Problem)] + [ctypes.c_int32] * 5 + [ctypes.c_void_p, ctypes.c_void_p]
    rc = L.bfm_plan_preview_host_chunks(pp, len(tab), n_query_rows, n_train_rows, n_sms, forced, ctypes.byref(c), ctypes.byref(per))
    if rc != BFM_OK:
        raise BfmError(rc, "bfm_plan_preview_host_chunks: invalid arguments")
    return c.value, per.value
