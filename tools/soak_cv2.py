"""Differential soak against live cv2 on random shapes / data kinds / modes (run on the GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import boslam_b200 as bb
from boslam_b200 import synth

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
eng = bb.Engine(0)
rng = np.random.default_rng(99)
bad = 0
for it in range(iters):
    nq, nt = int(rng.integers(1, 3000)), int(rng.integers(1, 6000))
    kind = it % 4
    if kind == 0:
        q, t, _ = synth.correlated(nq, nt, it)
    elif kind == 1:
        q, t = synth.tie_stress(nq, it), synth.tie_stress(nt, it + 1)
    elif kind == 2:
        q, t = synth.uniform(nq, it), synth.duplicate_rows(max(1, nt // 3), it)
    else:
        q, t = synth.uniform(nq, it), synth.uniform(nt, it + 5)
    k = int(rng.choice([1, 2, 3, 5, 8]))
    mask = (rng.random((nq, len(t))) < 0.4).astype(np.uint8) if it % 5 == 0 else None
    rows = cv2.BFMatcher_create(cv2.NORM_HAMMING).knnMatch(q, t, k, mask=mask)
    idx, dist = eng.knn(q, t, k, mask=mask)
    ok = True
    for i, r in enumerate(rows):
        want_i = [m.trainIdx for m in r] + [-1] * (k - len(r))
        want_d = [int(m.distance) for m in r] + [-1] * (k - len(r))
        if idx[i].tolist() != want_i or dist[i].tolist() != want_d:
            ok = False
            break
    if mask is None:
        cm = cv2.BFMatcher_create(cv2.NORM_HAMMING, crossCheck=True).match(q, t)
        qi, ti, d = eng.match(q, t, cross_check=True)
        ok = ok and [(m.queryIdx, m.trainIdx, m.distance) for m in cm] == list(zip(qi.tolist(), ti.tolist(), d.astype(float).tolist()))
    if not ok:
        bad += 1
        print(f"iter {it}: MISMATCH nq={nq} nt={len(t)} kind={kind} k={k} mask={mask is not None}", flush=True)
print(f"soak vs cv2 {cv2.__version__}: {iters} iterations, {bad} mismatches", flush=True)
sys.exit(1 if bad else 0)
