"""Single-problem kernel time against the finalize threshold (single finalizing CTA vs the tile-parallel
fin_count / fin_write kernels), at the small BASELINE shapes."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import boslam_b200 as bb
from boslam_b200 import synth
eng = bb.Engine(0)
eng.set_tuning(timing=1)
base, other = synth.uniform(65536, 7), synth.uniform(65536, 8)
for nq, nt in ((1000, 1000), (2000, 2000), (2000, 20000), (4096, 1024), (4096, 4096), (4096, 20000), (6000, 6000)):
    q, t = torch.from_numpy(base[:nq]).cuda(), torch.from_numpy(other[:nt]).cuda()
    tab = bb.make_problems([nq], [nt])
    for mode, kw in (("k2+ratio", dict(k=2, ratio=0.8)), ("cross+gate", dict(cross_check=True, max_distance=30))):
        line = f"{nq:6d} x {nt:6d} {mode:10s}"
        for thr in (0, 512):
            eng.set_tuning(finalize_rows=thr)
            ts = []
            for _ in range(12):
                eng.match_batched_device(q, t, tab, **kw)
                ts.append(eng.launch_info()["scan_ms"])
            line += f"   thr={thr:4d}: {np.median(ts[2:]) * 1e3:7.1f} us ({eng.launch_info()['kernels_launched']} kernels)"
        print(line, flush=True)
