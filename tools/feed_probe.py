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
    sets.append((pq, pt))
tab = bb.make_problems([N] * P, [N] * P)
out = bb.HostBatchBuffers(P * N, P, k=2)
def run(n=30):
    for i in range(5): eng.match_batched(sets[i % 6][0].array, sets[i % 6][1].array, tab, k=2, ratio=0.8, out=out)
    t0 = time.perf_counter()
    for i in range(n): eng.match_batched(sets[i % 6][0].array, sets[i % 6][1].array, tab, k=2, ratio=0.8, out=out)
    return (time.perf_counter() - t0) / n * 1e3
eng.set_tuning(feeders=-1)
print(f"copy engine gate: {run():.4f} ms", flush=True)
for ramp in (0, -1):
    for feeders in (16, 24, 32):
        for rows in (4096, 8192, 16384):
            eng.set_tuning(feeders=feeders, feed_rows=rows, ramp=ramp)
            print(f"ramp={ramp:2d} feeders={feeders:2d} rows/round={rows:5d}: {run():.4f} ms", flush=True)
