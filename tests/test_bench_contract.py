"""bench.py's output contract, checked on CPU through the reference arm (cv2 on the host cores): exactly one JSON
line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    pytest.importorskip("cv2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "hamming_pairs_per_s" and d["unit"] == "pairs/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k
    assert d["vs_baseline"] is None and d["higher_is_better"] is True and d["value"] > 0 and "workload" in d["config"]
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.bench_config() and d["scaling"] == "strong"   # both arms print the same config dict
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_b200_arm_refuses_to_run_without_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("only meaningful on a box without a GPU")
    except Exception:
        pytest.skip("torch not importable")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CPU path" in (out.stderr + out.stdout)


@pytest.mark.gpu
def test_b200_arm_line_has_every_contract_key():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--no-extras"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert "impl" not in d and d["metric"] == "hamming_pairs_per_s" and d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3
    assert d["scaling"] == "strong" and d["vs_baseline"] is None and d["dtype"] == "s8" and d["data"] == "synthetic"
    assert d["config"]["workload"] == "loop_closing" and "l2" in d["config"] and d["config"]["pairs"] == 256
    assert d["gpu_launches"] == 3 * 3                                # expansion, tcgen05 scan, finalize tiles: three kernels per step
    assert d["launch"]["form"].startswith("tensor")
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "tensor" and r["unit"] == "TOP/s" and r["frac"] > 0.3 and r["hbm"]["peak"] > 0
    assert 0 < r["kernel_ms"] < d["ms_per_step"] and 0.5 < r["kernel_share_of_step"] < 1.0   # the scan alone, measured live
    assert d["verify"]["tables_crc32"] > 0 and d["cpu_baseline"]["tables_equal_cv2"] is True
    assert d["sustained"]["seconds"] >= 2.0 and d["sustained"]["value"] > 0
    assert 0.2 < r["epilogue_alu"]["frac"] <= 1.0                    # the pipe that binds the scan in practice
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 30e6 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    c = d["clocks"]
    assert "sm_mhz" in c and "reasons" in c
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0
    assert d["value"] / cb["value"] > 50                             # a GPU run that is not far above the host is not a GPU run
