import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import boslam_b200 as bb
from boslam_b200 import synth
P, N = 256, 4000
eng = bb.Engine(0)
q, t = synth.keyframe_pair_batch(P, N, 0)
q = q[:P * 16]
pq, pt = bb.PinnedBuffer(q.shape), bb.PinnedBuffer(t.shape)
pq.array[...] = q; pt.array[...] = t
tab = bb.make_problems([16] * P, [N] * P)
out = bb.HostBatchBuffers(P * 16, P, k=2)
mb = (q.nbytes + t.nbytes) / 1e6
def run(n=30):
    for i in range(5): eng.match_batched(pq.array, pt.array, tab, k=2, ratio=0.8, out=out)
    t0 = time.perf_counter()
    for i in range(n): eng.match_batched(pq.array, pt.array, tab, k=2, ratio=0.8, out=out)
    return (time.perf_counter() - t0) / n * 1e3
eng.set_tuning(feeders=-1)
ms = run(); print(f"copy engine gate: {ms:.4f} ms  {mb/ms:.1f} GB/s", flush=True)
eng.set_tuning(pipeline_chunks=1)
ms = run(); print(f"single copy then kernel: {ms:.4f} ms  {mb/ms:.1f} GB/s", flush=True)
eng.set_tuning(pipeline_chunks=0)
for feeders in (8, 16, 32):
    for rows in (4096, 8192, 32768):
        eng.set_tuning(feeders=feeders, feed_rows=rows)
        ms = run(); print(f"feeders={feeders:2d} rows/round={rows:5d}: {ms:.4f} ms  {mb/ms:.1f} GB/s", flush=True)
