import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import boslam_b200 as bb
from boslam_b200 import synth
P, N = 256, 2000
eng = bb.Engine(0)
sets = [tuple(torch.from_numpy(a).cuda() for a in synth.keyframe_pair_batch(P, N, s)) for s in range(6)]
tab = bb.make_problems([N] * P, [N] * P)
ob = bb.HostBatchBuffers(P * N, P, k=2)
host_dest = {"m_query": ob.m[0].ctypes.data, "m_train": ob.m[1].ctypes.data, "m_dist": ob.m[2].ctypes.data, "count": ob.count.ctypes.data}
dev_dest = {"m": torch.empty((3, P * N), dtype=torch.int32, device="cuda"), "count": torch.zeros(P, dtype=torch.int32, device="cuda")}
def run(dest, n=30):
    for i in range(5):
        eng.match_batched_device(sets[i % 6][0], sets[i % 6][1], tab, k=2, ratio=0.8, out=dest); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        eng.match_batched_device(sets[i % 6][0], sets[i % 6][1], tab, k=2, ratio=0.8, out=dest); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
print(f"resident inputs, device outputs, sync per step: {run(dev_dest):.4f} ms")
print(f"resident inputs, pinned-host outputs, sync per step: {run(host_dest):.4f} ms")
