"""Single-problem calls against the planner knobs (waves / segment_rows / queries_per_thread): is the >= 3-wave
cut right for mid-size problems, or does one exact wave win?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import boslam_b200 as bb
from boslam_b200 import synth
eng = bb.Engine(0)
eng.set_tuning(timing=1)
base, other = synth.uniform(32768, 7), synth.uniform(32768, 8)
for nq, nt in ((2000, 20000), (2000, 2000), (1000, 1000), (4096, 4096), (600, 600)):
    q, t = torch.from_numpy(base[:nq]).cuda(), torch.from_numpy(other[:nt]).cuda()
    tab = bb.make_problems([nq], [nt])
    for mode, kw in (("k2+ratio", dict(k=2, ratio=0.8)), ("cross+gate", dict(cross_check=True, max_distance=30))):
        line = f"{nq:5d} x {nt:5d} {mode:10s}"
        for knobs in (dict(), dict(waves=1), dict(waves=2), dict(queries_per_thread=1), dict(queries_per_thread=1, waves=1),
                      dict(queries_per_thread=2, waves=1), dict(queries_per_thread=4, waves=1), dict(queries_per_thread=4, waves=2)):
            eng.set_tuning(waves=0, queries_per_thread=0)
            eng.set_tuning(**knobs)
            ts = []
            for _ in range(14):
                eng.match_batched_device(q, t, tab, **kw)
                ts.append(eng.launch_info()["scan_ms"])
            li = eng.launch_info()
            tag = ",".join(f"{k[0]}{v}" for k, v in knobs.items()) or "auto"
            line += f" | {tag}: R{li['queries_per_thread']} g{li['scan_grid']} {np.median(ts[3:]) * 1e3:5.1f}"
        print(line, flush=True)
