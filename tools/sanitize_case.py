"""Small pass over every kernel for compute-sanitizer (memcheck).  The copy-engine-gated host path is switched
off (pipeline_chunks=1): a sanitizer serialises launches, which the gate's design (kernel running while the
rest of the upload is queued) deliberately does not tolerate beyond its 4 s time-out."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import boslam_b200 as bb  # noqa: E402
from boslam_b200 import synth  # noqa: E402

eng = bb.Engine(0)
eng.set_tuning(pipeline_chunks=1)
q, t, qxy, txy, _ = synth.window_scene(300, 700, 1)
for r in (1, 2, 4):
    eng.set_tuning(queries_per_thread=r)
    eng.knn(q, t, 2)
    eng.knn(q, t, 1)
    eng.match(q, t, cross_check=True, max_distance=40)
eng.set_tuning(queries_per_thread=0)
eng.knn(q, t, 5)
eng.match(q, t, k=2, ratio=0.8, window=(qxy, txy, 15.0))
eng.set_tuning(window_bins=1)
eng.match(q, t, cross_check=True, window=(qxy, txy, 15.0))
eng.set_tuning(window_bins=0)
mask = (np.random.default_rng(0).random((300, 700)) < 0.3).astype(np.uint8)
eng.knn(q, t, 2, mask=mask)
qs, ts = synth.keyframe_pairs(5, 260, seed=2)
eng.match_pairs(qs, ts, k=2, ratio=0.8)
eng.match_pairs([qs[0]] * 3 + [np.zeros((0, 32), np.uint8)], ts[:3] + [ts[3]], cross_check=True)
big = synth.uniform(8300, 3)
eng.knn(big, t[:200], 2)
eng.match(big, t[:200], cross_check=True)
sc = synth.local_map_scene(900, 1300, 200, seed=4)
store = bb.MapStore(1000, engine=eng)
store.update(np.arange(900), sc["desc"], sc["pt3d"], sc["normal"])
store.track(sc["des"], sc["kp"], sc["R"], sc["t"], sc["see_vector"], sc["edges"])
store.track(sc["des"], sc["kp"], sc["R"], sc["t"], sc["see_vector"], sc["edges"], cross_check=False, k=2, ratio=0.8,
            max_distance=None, window_radius=15.0)
obs = np.random.default_rng(1).integers(0, 256, (100, 10, 32), dtype=np.uint8)
bb.select_representative(obs, np.random.default_rng(2).integers(0, 11, 100), engine=eng)
bank = bb.KeyframeBank(2048, engine=eng)
for i in range(4):
    bank.add(i, ts[i])
bank.match_pairs([(0, 1), (2, 3), (1, 1)], cross_check=True)
print("sanitize case done", flush=True)
