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
print(f"pinned in/out: {run():.4f} ms", flush=True)
def run_mix(pin_in, pin_out, n=20):
    def call(i):
        a = sets[i % 6]
        qq, tt = (a[0].array, a[1].array) if pin_in else (a[2], a[3])
        return eng.match_batched(qq, tt, tab, k=2, ratio=0.8, out=out if pin_out else None)
    for i in range(3): call(i)
    t0 = time.perf_counter()
    for i in range(n): call(i)
    return (time.perf_counter() - t0) / n * 1e3
for pin_in in (True, False):
    for pin_out in (True, False):
        print(f"pinned_in={pin_in} pinned_out={pin_out}: {run_mix(pin_in, pin_out):.4f} ms", flush=True)
os.environ["BFM_TRACE"] = "1"
run_mix(False, True, 2)
