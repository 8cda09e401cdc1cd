"""Kernel-variant sweep on device-resident data: pairs/s per (shape, mode, R, popc_mode)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import boslam_b200 as bb  # noqa: E402
from boslam_b200 import synth  # noqa: E402


def time_call(eng, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best_scan, best_total = 1e9, 1e9
    for _ in range(reps):
        fn()
        li = eng.launch_info()
        best_scan = min(best_scan, li["scan_ms"])
        best_total = min(best_total, li["total_ms"])
    return best_scan, best_total, eng.launch_info()


def main():
    eng = bb.Engine(0)
    eng.set_tuning(timing=1)
    out = []
    shapes = [("loop256x2000", 256, 2000, 2000), ("track2000x20000", 1, 2000, 20000), ("sq16k", 1, 16384, 16384),
              ("cfg1_1000", 1, 1000, 1000), ("lm20x2000", 20, 2000, 2000)]
    if len(sys.argv) > 1:
        shapes = [s for s in shapes if s[0] in sys.argv[1:]]
    for name, P, nq, nt in shapes:
        q = torch.from_numpy(synth.uniform(P * nq, 1)).cuda()
        t = torch.from_numpy(synth.uniform(P * nt, 2)).cuda()
        tab = bb.make_problems([nq] * P, [nt] * P)
        pairs = P * nq * nt
        for mode, kw in (("k2", dict(k=2)), ("k1", dict(k=1)), ("cross", dict(k=1, cross_check=True))):
            for r in (4, 2, 1):
                for pm in (8, 5, 4, 50, 40):
                    for waves in (0,):
                        eng.set_tuning(queries_per_thread=r, popc_mode=pm, waves=waves)
                        obuf = {}
                        fn = lambda: eng.match_batched_device(q, t, tab, **kw)
                        scan, total, li = time_call(eng, fn)
                        rec = dict(shape=name, mode=mode, R=r, pm=pm, scan_ms=scan, total_ms=total,
                                   gpairs_scan=pairs / scan / 1e6, gpairs_total=pairs / total / 1e6,
                                   grid=li["scan_grid"], seg_rows=li["train_rows_per_segment"])
                        out.append(rec)
                        print(json.dumps(rec), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/sweep.json", "w"), indent=1)


if __name__ == "__main__":
    main()
