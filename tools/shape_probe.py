"""Kernel time (CUDA events around the one launch) of the small BASELINE shapes for each register tile
and a few segment lengths: what the planner should pick."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import boslam_b200 as bb  # noqa: E402
from boslam_b200 import synth  # noqa: E402

eng = bb.Engine(0)
eng.set_tuning(timing=1)


def probe(name, call, pairs):
    for r in (0, 1, 2, 4):
        for seg in (0, 64, 128, 256, 512):
            eng.set_tuning(queries_per_thread=r, segment_rows=seg)
            ts = []
            for _ in range(12):
                call()
                ts.append(eng.launch_info()["scan_ms"])
            li = eng.launch_info()
            t = float(np.median(ts[2:]))
            print(f"{name:28s} R={r} seg={seg:4d} -> R={li['queries_per_thread']} grid={li['scan_grid']:5d} rows/seg={li['train_rows_per_segment']:5d} "
                  f"{t * 1e3:7.1f} us  {pairs / t / 1e6:7.1f} Gpair/s", flush=True)
    eng.set_tuning(queries_per_thread=0, segment_rows=0)


q, t, _ = synth.correlated(1000, 1000, 11)
qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
tab1 = bb.make_problems([1000], [1000])
probe("f2f 1000x1000 cc", lambda: eng.match_batched_device(qd, td, tab1, cross_check=True, max_distance=29), 1e6)

q, t, qxy, txy, _ = synth.window_scene(2000, 20000, 12)
qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
qxyd, txyd = torch.from_numpy(qxy).cuda(), torch.from_numpy(txy).cuda()
tab2 = bb.make_problems([2000], [20000])
probe("track 2000x20000 win+ratio", lambda: eng.match_batched_device(qd, td, tab2, k=2, ratio=0.8, window=(qxyd, txyd, 15.0)), 4e7)
probe("track 2000x20000 cc", lambda: eng.match_batched_device(qd, td, tab2, cross_check=True, max_distance=30), 4e7)
probe("track 2000x20000 k2", lambda: eng.match_batched_device(qd, td, tab2, k=2, ratio=0.8), 4e7)

qb, tb = synth.keyframe_pair_batch(20, 2000, 13)
tab3 = bb.make_problems([2000] * 20, [2000] * 20)
qbd, tbd = torch.from_numpy(qb).cuda(), torch.from_numpy(tb).cuda()
probe("localmap 20x2000x2000 k2", lambda: eng.match_batched_device(qbd, tbd, tab3, k=2, ratio=0.8), 8e7)
