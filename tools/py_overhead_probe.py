import cProfile, pstats, io, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import boslam_b200 as bb
from boslam_b200 import synth
eng = bb.Engine(0)
sc = synth.local_map_scene(20000, 20000, 2000, seed=14)
store = bb.MapStore(20000, engine=eng)
store.update(np.arange(20000), sc["desc"], sc["pt3d"], sc["normal"])
targs = (sc["des"], sc["kp"], sc["R"], sc["t"], sc["see_vector"], sc["edges"])
q, t, _ = synth.correlated(1000, 1000, 3)
for name, f in (("MapStore.track", lambda: store.track(*targs)), ("Engine.match 1000x1000 cc", lambda: eng.match(q, t, cross_check=True, max_distance=30, strict=True))):
    for _ in range(50): f()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(1000): f()
    pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(12)
    print("=====", name); print("\n".join(s.getvalue().splitlines()[:30]))
