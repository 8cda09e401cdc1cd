"""BASELINE config 5's two ncu points: one brute-force problem of 16k x 16k and one of 64k x 64k uniform
descriptors, k = 2 (and k = 1 / cross-check at 64k), device-resident; a few launches each for
`ncu --set full -k regex:bfm_scan_kernel`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import boslam_b200 as bb
from boslam_b200 import synth
eng = bb.Engine(0)
base, other = synth.uniform(65536, 7), synth.uniform(65536, 8)
for n, kws in ((16384, (dict(k=2),)), (65536, (dict(k=2), dict(k=1), dict(cross_check=True)))):
    q, t = torch.from_numpy(base[:n]).cuda(), torch.from_numpy(other[:n]).cuda()
    tab = bb.make_problems([n], [n])
    for kw in kws:
        for _ in range(1 if "--once" in sys.argv else 3):
            eng.match_batched_device(q, t, tab, **kw)
        torch.cuda.synchronize()
print("done")
