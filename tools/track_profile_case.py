"""Tracking step (MapStore.track) a few times, both variants, for an ncu launch list."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import boslam_b200 as bb
from boslam_b200 import synth
eng = bb.Engine(0)
sc = synth.local_map_scene(20000, 20000, 2000, seed=14)
store = bb.MapStore(20000, engine=eng)
store.update(np.arange(20000), sc["desc"], sc["pt3d"], sc["normal"])
targs = (sc["des"], sc["kp"], sc["R"], sc["t"], sc["see_vector"], sc["edges"])
for _ in range(4):
    store.track(*targs)                                                                             # reference-faithful: cross-check + gate
for _ in range(4):
    store.track(*targs, cross_check=False, k=2, ratio=0.8, max_distance=None, window_radius=15.0)   # north star: window + ratio
q, t, qxy, txy, _ = synth.window_scene(2000, 20000, 12)
for _ in range(3):
    eng.match(q, t, k=2, ratio=0.8, window=(qxy, txy, 15.0))
print("done")
