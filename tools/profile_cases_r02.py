"""The launches ncu looks at in round 2 (device-resident inputs, fixed order, two launches per case after one warm-up):
  0-1  loop closing 256 x 2000^2 k = 2 + ratio          static form   (the headline batch)
  2-3  a rank's share at 8 GPUs: 32 x 2000^2            persistent form
  4-5  tracking 2000 x 20000 cross-check + gate         persistent form
  6-7  frame to frame 1000 x 1000 cross-check + gate    persistent form
  8-9  sweep 16k x 16k k = 2                            persistent form
 10-11 sweep 64k x 64k k = 2                            static form
`ncu -k regex:bfm_scan` sees exactly these (the warm-up launches are skipped with -s)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import boslam_b200 as bb
from boslam_b200 import synth
eng = bb.Engine(0)
plans = []
for P in (256, 32):
    q, t = synth.keyframe_pair_batch(P, 2000, seed=100 + P)
    plans.append(eng.plan_device(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), bb.make_problems([2000] * P, [2000] * P), k=2, ratio=0.8))
q, t, _ = synth.correlated(2000, 20000, 12)
plans.append(eng.plan_device(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), bb.make_problems([2000], [20000]), cross_check=True, max_distance=30))
q, t, _ = synth.correlated(1000, 1000, 11)
plans.append(eng.plan_device(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), bb.make_problems([1000], [1000]), cross_check=True, max_distance=30, strict=True))
base, other = synth.uniform(65536, 7), synth.uniform(65536, 8)
for n in (16384, 65536):
    plans.append(eng.plan_device(torch.from_numpy(base[:n]).cuda(), torch.from_numpy(other[:n]).cuda(), bb.make_problems([n], [n]), k=2))
for p in plans:      # warm-up: one launch per case (12 launches follow)
    p.run()
torch.cuda.synchronize()
for p in plans:
    for _ in range(2):
        p.run()
    torch.cuda.synchronize()
    print(eng.launch_info())
print("done")
