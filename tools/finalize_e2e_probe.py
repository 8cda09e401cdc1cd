"""End-to-end (numpy in -> numpy out) times of the single-frame calls against the finalize threshold:
the single finalizing CTA (finalize_rows = never) vs the tile-parallel fin_count / fin_write kernels."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import boslam_b200 as bb
from boslam_b200 import synth

eng = bb.Engine(0)


def timeit(f, n=300):
    for _ in range(30):
        f()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    return (time.perf_counter() - t0) / n * 1e6


sc = synth.local_map_scene(20000, 20000, 2000, seed=14)
store = bb.MapStore(20000, engine=eng)
store.update(np.arange(20000), sc["desc"], sc["pt3d"], sc["normal"])
targs = (sc["des"], sc["kp"], sc["R"], sc["t"], sc["see_vector"], sc["edges"])
cases = []
for name, (nq, nt), kw in (("2000x2000 cc gate", (2000, 2000), dict(cross_check=True, max_distance=30)),
                           ("2000x2000 k2 ratio", (2000, 2000), dict(k=2, ratio=0.8)),
                           ("2000x20000 cc gate", (2000, 20000), dict(cross_check=True, max_distance=30)),
                           ("2000x20000 k2 ratio", (2000, 20000), dict(k=2, ratio=0.8))):
    q, t, _ = synth.correlated(nq, nt, 3)
    cases.append((name + " Engine.match", lambda q=q, t=t, kw=kw: eng.match(q, t, **kw)))
    cases.append((name + " Engine.knn  ", lambda q=q, t=t, kw=kw: eng.knn(q, t, kw.get("k", 1))))
cases.append(("MapStore.track cc gate30", lambda: store.track(*targs)))
cases.append(("MapStore.track window+ratio", lambda: store.track(*targs, cross_check=False, k=2, ratio=0.8, max_distance=None, window_radius=15.0)))
for name, f in cases:
    line = f"{name:36s}"
    for rep in range(2):
        for thr in (1 << 30, 1792):
            eng.set_tuning(finalize_rows=thr)
            line += f"  thr={'never' if thr > 1 << 20 else thr}: {timeit(f):7.1f} us"
    print(line, flush=True)
