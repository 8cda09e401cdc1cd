import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import boslam_b200 as bb
from boslam_b200 import synth
P, N = 256, 2000
eng = bb.Engine(0)
q, t = synth.keyframe_pair_batch(P, N, 1)
pq, pt = bb.PinnedBuffer(q.shape), bb.PinnedBuffer(t.shape)
pq.array[...] = q; pt.array[...] = t
tab = bb.make_problems([N] * P, [N] * P)
out = bb.HostBatchBuffers(P * N, P, k=2)
for c in (0, 3, 5, 6):
    eng.set_tuning(pipeline_chunks=c)
    for i in range(4):
        if i == 3: os.environ['BFM_TRACE'] = '1'
        t0 = time.perf_counter()
        eng.match_batched(pq.array, pt.array, tab, k=2, ratio=0.8, out=out)
        dt = time.perf_counter() - t0
    os.environ.pop('BFM_TRACE')
    print(f"chunks={c} wall {dt*1e3:.3f} ms", flush=True)
