"""The C ABI used from plain C (no Python binding in the loop): tests/c/abi_smoke.c is compiled against
include/bfm.h, linked with boslam_b200/libbfm_b200.so and run on the GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = tmp_path / "abi_smoke"
    lib_dir = os.path.join(ROOT, "boslam_b200")
    subprocess.check_call(["gcc", "-O1", "-std=gnu11", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
                           "-o", str(exe), "-L", lib_dir, "-l:libbfm_b200.so", "-Wl,-rpath," + lib_dir])
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_c_caller_links_against_the_library(tmp_path):
    """CPU part: the C caller compiles and links against the header + shared library."""
    assert _build(tmp_path).exists()


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_c_caller_runs_on_the_gpu(tmp_path):
    out = subprocess.run([str(_build(tmp_path))], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "c abi ok" in out.stdout
