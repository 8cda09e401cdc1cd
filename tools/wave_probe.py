import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import boslam_b200 as bb
from boslam_b200 import synth
eng = bb.Engine(0)
eng.set_tuning(timing=1)
def probe(name, call, pairs):
    for r in (0, 1, 2, 4):
        for waves in (0, 1, 2, 3):
            eng.set_tuning(queries_per_thread=r, waves=waves)
            ts = []
            for _ in range(15):
                call(); ts.append(eng.launch_info()["scan_ms"])
            li = eng.launch_info()
            t = float(np.median(ts[3:]))
            print(f"{name:26s} R={r} waves={waves} -> R={li['queries_per_thread']} grid={li['scan_grid']:5d} rows/seg={li['train_rows_per_segment']:5d} {t*1e3:7.1f} us {pairs/t/1e6:7.1f} Gp/s", flush=True)
    eng.set_tuning(queries_per_thread=0, waves=0)
q, t, _ = synth.correlated(2000, 20000, 12)
qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
tab2 = bb.make_problems([2000], [20000])
probe("track 2000x20000 cc", lambda: eng.match_batched_device(qd, td, tab2, cross_check=True, max_distance=30), 4e7)
probe("track 2000x20000 k2", lambda: eng.match_batched_device(qd, td, tab2, k=2, ratio=0.8), 4e7)
q, t, _ = synth.correlated(1000, 1000, 11)
qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
tab1 = bb.make_problems([1000], [1000])
probe("f2f 1000x1000 cc", lambda: eng.match_batched_device(qd, td, tab1, cross_check=True), 1e6)
qb, tb = synth.keyframe_pair_batch(20, 2000, 13)
tab3 = bb.make_problems([2000] * 20, [2000] * 20)
qbd, tbd = torch.from_numpy(qb).cuda(), torch.from_numpy(tb).cuda()
probe("localmap 20x2000x2000 k2", lambda: eng.match_batched_device(qbd, tbd, tab3, k=2, ratio=0.8), 8e7)
qb, tb = synth.keyframe_pair_batch(64, 2000, 13)
tab4 = bb.make_problems([2000] * 64, [2000] * 64)
qbd, tbd = torch.from_numpy(qb).cuda(), torch.from_numpy(tb).cuda()
probe("loop 64x2000x2000 k2", lambda: eng.match_batched_device(qbd, tbd, tab4, k=2, ratio=0.8), 64*4e6)
