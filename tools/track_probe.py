import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import boslam_b200 as bb
from boslam_b200 import synth
eng = bb.Engine(0)
sc = synth.local_map_scene(20000, 20000, 2000, seed=14)
store = bb.MapStore(20000, engine=eng)
store.update(np.arange(20000), sc["desc"], sc["pt3d"], sc["normal"])
targs = (sc["des"], sc["kp"], sc["R"], sc["t"], sc["see_vector"], sc["edges"])
def timeit(f, n=200):
    for _ in range(20): f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e6
print("track cc gate:", timeit(lambda: store.track(*targs)))
print("track window+ratio:", timeit(lambda: store.track(*targs, cross_check=False, k=2, ratio=0.8, max_distance=None, window_radius=15.0)))
print("track window+cc:", timeit(lambda: store.track(*targs, window_radius=15.0)))
os.environ["BFM_TRACE"] = "1"
for _ in range(3): store.track(*targs)
for _ in range(3): store.track(*targs, cross_check=False, k=2, ratio=0.8, max_distance=None, window_radius=15.0)
