import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import boslam_b200 as bb
from boslam_b200 import synth
eng = bb.Engine(0)
def timeit(f, n=100):
    for _ in range(10): f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e6
for P in (20, 40, 64):
    qb, tb = synth.keyframe_pair_batch(P, 2000, 13)
    tab = bb.make_problems([2000] * P, [2000] * P)
    pq, pt = bb.PinnedBuffer(qb.shape), bb.PinnedBuffer(tb.shape)
    pq.array[...] = qb; pt.array[...] = tb
    ob = bb.HostBatchBuffers(P * 2000, P, k=2)
    for kb in (0, 512, 1024, 2048):
        eng.set_tuning(pipeline_min_kb=kb)
        a = timeit(lambda: eng.match_batched(qb, tb, tab, k=2, ratio=0.8))
        b = timeit(lambda: eng.match_batched(pq.array, pt.array, tab, k=2, ratio=0.8, out=ob))
        print(f"P={P} ({qb.nbytes*2/1e6:.1f} MB) pipeline_min_kb={kb:5d}: pageable {a:7.1f} us   pinned {b:7.1f} us", flush=True)
