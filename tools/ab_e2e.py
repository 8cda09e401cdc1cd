"""A/B of two builds on one box, host path: Engine.match / match_batched end to end (numpy in -> numpy out).
usage: ab_e2e.py <path holding a boslam_b200 package>"""
import os, sys, time
root = sys.argv[1]
sys.path.insert(0, root)
import numpy as np
import boslam_b200 as bb
from boslam_b200 import synth
assert os.path.abspath(bb.__file__).startswith(os.path.abspath(root)), bb.__file__
eng = bb.Engine(0)

def clock(f, n=200):
    for _ in range(20):
        f()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        for _ in range(n):
            f()
        best = min(best, (time.perf_counter() - t0) / n)
    return best * 1e6

q1, t1, _ = synth.correlated(1000, 1000, 11)
q2, t2, qxy, txy, _ = synth.window_scene(2000, 20000, 12)
qb, tb = synth.keyframe_pair_batch(20, 2000, 13)
tab = bb.make_problems([2000] * 20, [2000] * 20)
print(f"{root}: f2f 1000^2 cross {clock(lambda: eng.match(q1, t1, cross_check=True, max_distance=30, strict=True)):6.1f} us | "
      f"track 2000x20000 cross {clock(lambda: eng.match(q2, t2, cross_check=True, max_distance=30)):6.1f} | "
      f"track window+ratio {clock(lambda: eng.match(q2, t2, k=2, ratio=0.8, window=(qxy, txy, 15.0))):6.1f} | "
      f"local mapping 20 pairs k2 {clock(lambda: eng.match_batched(qb, tb, tab, k=2, ratio=0.8), 100):6.1f} | "
      f"cross {clock(lambda: eng.match_batched(qb, tb, tab, cross_check=True, max_distance=30), 100):6.1f}")
