"""Integer-pipe issue-rate probes: writes gpurun_out/popc_peak.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boslam_b200 import _ffi  # noqa: E402


def main():
    info = _ffi.device_info(0)
    res = {"device": info, "probes": {}}
    for it in (2000, 8000):
        res["probes"][f"iters{it}"] = _ffi.microbench(0, it)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/popc_peak.json", "w") as f:
        json.dump(res, f, indent=1)
    for k, v in res["probes"]["iters8000"].items():
        print(f"{k:16s} {v['ops_per_clk_per_sm']:8.2f} ops/clk/SM  {v['ops_per_s'] / 1e12:8.3f} Tops/s  sm_mhz~{v['sm_mhz']:.0f}")


if __name__ == "__main__":
    main()
