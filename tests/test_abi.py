"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/bfm.h declares, and fails loudly (no CPU fallback) when no GPU is present."""
import ctypes
import os
import re

import numpy as np
import pytest

import boslam_b200 as bb
from boslam_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "bfm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bfm_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported():
    L = _ffi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/bfm.h but not exported"
    assert sorted(_ffi.EXPORTED_SYMBOLS) == declared


def test_abi_version_and_struct_layout():
    L = _ffi.lib()
    assert L.bfm_abi_version() == _ffi.ABI_VERSION
    assert ctypes.sizeof(_ffi.Problem) == 24
    assert ctypes.sizeof(_ffi.Options) == 64
    assert _ffi.Options.ratio.offset == 16 and _ffi.Options.mask.offset == 32
    assert ctypes.sizeof(_ffi.LaunchInfo) == 40
    assert ctypes.sizeof(_ffi.Outputs) == 56 and _ffi.Outputs.multicast.offset == 48        # bfm_outputs_t
    assert ctypes.sizeof(_ffi.TrackParams) == 128 and _ffi.TrackParams.width.offset == 120  # bfm_track_params_t


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(bb.BfmError) as ei:
        bb.Engine(0)
    assert "no CPU path" in str(ei.value)
    with pytest.raises(bb.BfmError):
        bb.match(np.zeros((4, 32), np.uint8), np.zeros((4, 32), np.uint8))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "boslam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "import cv2" not in src, f


def test_make_problems():
    tab = bb.make_problems([3, 4], [5, 6])
    assert tab.tolist() == [[0, 3, 0, 5, 0, 0], [3, 4, 5, 6, 3, 0]]
    tab = bb.make_problems([3, 3], [5, 6], shared_query=True)
    assert tab[:, 0].tolist() == [0, 0] and tab[:, 4].tolist() == [0, 3]


def test_dmatch_surface():
    m = bb.DMatch(1, 2, 0, 7.0)
    assert (m.queryIdx, m.trainIdx, m.imgIdx, m.distance) == (1, 2, 0, 7.0)
    assert bb.NORM_HAMMING == 6
    mm = bb.BFMatcher_create(bb.NORM_HAMMING, crossCheck=True)  # construction needs no GPU
    assert mm.crossCheck is True
    with pytest.raises(ValueError):
        bb.BFMatcher_create(4)


def test_struct_layout_against_the_c_header(tmp_path):
    """Compile include/bfm.h with gcc and compare every struct's size / key offsets with the ctypes mirror."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "bfm.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(bfm_problem_t), sizeof(bfm_options_t),'
                   ' offsetof(bfm_options_t, mask), sizeof(bfm_outputs_t), offsetof(bfm_outputs_t, multicast),'
                   ' sizeof(bfm_track_params_t), offsetof(bfm_track_params_t, cos_max), sizeof(bfm_launch_info_t),'
                   ' offsetof(bfm_launch_info_t, scan_ms)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(_ffi.Problem), ctypes.sizeof(_ffi.Options), _ffi.Options.mask.offset, ctypes.sizeof(_ffi.Outputs),
            _ffi.Outputs.multicast.offset, ctypes.sizeof(_ffi.TrackParams), _ffi.TrackParams.cos_max.offset,
            ctypes.sizeof(_ffi.LaunchInfo), _ffi.LaunchInfo.scan_ms.offset]
    assert got == want


def test_install_routes_cv2_constructors():
    """boslam_b200.install(cv2): NORM_HAMMING matchers come from this package (construction needs no GPU), other
    norms still come from OpenCV; uninstall restores the originals."""
    cv2 = pytest.importorskip("cv2")
    orig_create, orig_cls = cv2.BFMatcher_create, cv2.BFMatcher
    bb.install(cv2)
    try:
        m = cv2.BFMatcher_create(cv2.NORM_HAMMING, crossCheck=True)          # slam/tracking.py:45, verbatim
        assert isinstance(m, bb.BFMatcher) and m.crossCheck is True
        assert isinstance(cv2.BFMatcher(cv2.NORM_HAMMING), bb.BFMatcher)
        assert isinstance(cv2.BFMatcher.create(cv2.NORM_HAMMING, True), bb.BFMatcher)
        l2 = cv2.BFMatcher_create(cv2.NORM_L2)
        assert not isinstance(l2, bb.BFMatcher)
        a = np.random.default_rng(0).random((5, 8)).astype(np.float32)
        assert len(l2.match(a, a)) == 5                                       # OpenCV still does the float norms
        bb.install(cv2)                                                       # idempotent
    finally:
        bb.uninstall(cv2)
    assert cv2.BFMatcher_create is orig_create and cv2.BFMatcher is orig_cls
