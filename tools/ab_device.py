import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import boslam_b200 as bb
from boslam_b200 import synth
eng = bb.Engine(0)
eng.set_tuning(timing=1)
P, N = 256, 2000
sets = [tuple(torch.from_numpy(a).cuda() for a in synth.keyframe_pair_batch(P, N, s)) for s in range(6)]
tab = bb.make_problems([N] * P, [N] * P)
out = {"m": torch.empty((3, P * N), dtype=torch.int32, device="cuda"), "count": torch.zeros(P, dtype=torch.int32, device="cuda")}
import sys as _s
R = int(os.environ.get("AB_R", "0"))
eng.set_tuning(queries_per_thread=R)
for kw in (dict(k=2, ratio=0.8), dict(cross_check=True), dict(k=1)):
    ts = []
    for i in range(25):
        eng.match_batched_device(sets[i % 6][0], sets[i % 6][1], tab, out=out, **kw)
        ts.append(eng.launch_info()["scan_ms"])
    t = float(np.median(ts[5:]))
    print(f"{os.environ.get('BFM_LIB_PATH','default')[-12:]} R={R} {str(kw):32s} {t:.4f} ms {P*N*N/t/1e6:.0f} Gp/s", flush=True)
