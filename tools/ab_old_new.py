"""A/B of two builds of the package on one box: the headline batch, device-resident, back to back.
usage: ab_old_new.py <path holding a boslam_b200 package> [knob=value ...]"""
import os, sys
root = sys.argv[1]
sys.path.insert(0, root)
import numpy as np
import torch
import boslam_b200 as bb
from boslam_b200 import synth
assert os.path.abspath(bb.__file__).startswith(os.path.abspath(root)), bb.__file__
knobs = {k: int(v) for k, v in (a.split("=") for a in sys.argv[2:])}
eng = bb.Engine(0)
eng.set_tuning(**knobs)
P, N = 256, 2000
tab = bb.make_problems([N] * P, [N] * P)
sets = []
for s in range(6):   # rotating inputs, > L2 in total
    q, t = synth.keyframe_pair_batch(P, N, seed=100 + s)
    sets.append((torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()))
out = eng.match_batched_device(sets[0][0], sets[0][1], tab, k=2, ratio=0.8)
for i in range(6):
    eng.match_batched_device(*sets[i % 6], tab, k=2, ratio=0.8, out=out)
torch.cuda.synchronize()
res = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30):
        eng.match_batched_device(*sets[i % 6], tab, k=2, ratio=0.8, out=out)
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 30 * 1e3)
print(f"{root} {knobs}: {[round(r, 1) for r in res]} us per call; launch {eng.launch_info()}; matches {int(out['count'].sum())}")
