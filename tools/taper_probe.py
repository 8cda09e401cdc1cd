"""Tapered tail of a batch plan (taper / taper_pct knobs): device-resident kernel time and the pinned host path,
headline batch (256 x 2000 x 2000, k = 2 + ratio) and the local-mapping batch (20 x 2000 x 2000)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import boslam_b200 as bb
from boslam_b200 import synth
from boslam_b200.engine import PinnedBuffer
eng = bb.Engine(0)
for P, N in ((256, 2000), (20, 2000)):
    host = [synth.keyframe_pair_batch(P, N, s) for s in range(6)]
    sets = [tuple(torch.from_numpy(a).cuda() for a in h) for h in host]
    pinned = []
    for q, t in host:
        pq, pt = PinnedBuffer(q.shape), PinnedBuffer(t.shape)
        pq.array[...] = q; pt.array[...] = t
        pinned.append((pq, pt))
    tab = bb.make_problems([N] * P, [N] * P)
    out = {"m": torch.empty((3, P * N), dtype=torch.int32, device="cuda"), "count": torch.zeros(P, dtype=torch.int32, device="cuda")}
    hb = bb.HostBatchBuffers(P * N, P, k=2)
    ref = None
    for taper, pct in ((1, 0), (2, 10), (2, 20), (2, 35), (4, 10), (4, 20), (4, 35), (8, 20), (8, 35), (1, 0)):
        eng.set_tuning(taper=taper, taper_pct=pct, timing=1)
        ts = []
        for i in range(30):
            eng.match_batched_device(sets[i % 6][0], sets[i % 6][1], tab, out=out, k=2, ratio=0.8)
            ts.append(eng.launch_info()["scan_ms"])
        grid = eng.launch_info()["scan_grid"]
        cnt = out["count"].cpu().numpy().copy()
        eng.set_tuning(timing=0)
        plan = eng.plan_batch(tab, k=2, ratio=0.8)
        for i in range(5):
            plan.run(pinned[i % 6][0].array, pinned[i % 6][1].array, hb)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(30):
            res = plan.run(pinned[i % 6][0].array, pinned[i % 6][1].array, hb)
        e2e = (time.perf_counter() - t0) / 30 * 1e3
        if ref is None:
            ref = cnt
        same = bool(np.array_equal(cnt, ref))
        t = float(np.median(ts[5:]))
        print(f"P={P:3d} taper={taper} pct={pct:2d}: grid {grid:5d}  kernel {t * 1e3:7.1f} us  {P * N * N / t / 1e6:6.0f} Gp/s   e2e {e2e * 1e3:7.1f} us   same_counts={same}", flush=True)
