import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import boslam_b200 as bb
from boslam_b200 import synth
P, N = 256, 2000
eng = bb.Engine(0)
sets = []
for s in range(6):
    q, t = synth.keyframe_pair_batch(P, N, s)
    pq, pt = bb.PinnedBuffer(q.shape), bb.PinnedBuffer(t.shape)
    pq.array[...] = q; pt.array[...] = t
    sets.append((pq, pt, q, t))
tab = bb.make_problems([N] * P, [N] * P)
out = bb.HostBatchBuffers(P * N, P, k=2)
def run(n=20):
    for i in range(3): eng.match_batched(sets[i % 6][0].array, sets[i % 6][1].array, tab, k=2, ratio=0.8, out=out)
    t0 = time.perf_counter()
    for i in range(n): eng.match_batched(sets[i % 6][0].array, sets[i % 6][1].array, tab, k=2, ratio=0.8, out=out)
    return (time.perf_counter() - t0) / n * 1e3
print(f"before torch: {run():.4f} {run():.4f} ms", flush=True)
import torch
torch.cuda.set_device(0)
x = torch.zeros(4, device="cuda")
print(f"after torch import + context: {run():.4f} {run():.4f} ms", flush=True)
dev_sets = [(torch.from_numpy(s[2]).cuda(), torch.from_numpy(s[3]).cuda()) for s in sets]
print(f"after device sets: {run():.4f} {run():.4f} ms", flush=True)
o = {"m": torch.empty((3, P * N), dtype=torch.int32, device="cuda"), "count": torch.zeros(P, dtype=torch.int32, device="cuda")}
for i in range(23):
    eng.match_batched_device(dev_sets[i % 6][0], dev_sets[i % 6][1], tab, k=2, ratio=0.8, out=o)
torch.cuda.synchronize()
print(f"after device loop: {run():.4f} {run():.4f} ms", flush=True)
from boslam_b200 import _ffi
_ffi.microbench(0, 4000, tests=("popc",))
print(f"after microbench: {run():.4f} {run():.4f} ms", flush=True)
