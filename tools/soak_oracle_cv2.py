"""CPU-only differential soak that pins the oracle: the numpy and the plain-C restatements against live
cv2.BFMatcher (the reference's own matcher) on random shapes / data kinds / k / masks / cross-check / ratio.
Runs in the authoring container (no GPU): `python tools/soak_oracle_cv2.py [iterations]`."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
from boslam_b200 import synth
from oracle import c_oracle, cv2_reference as ref, hamming_oracle as orc

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(2026)
bad, checks, t0 = 0, 0, time.time()
for it in range(iters):
    nq, nt = int(rng.integers(1, 1500)), int(rng.integers(1, 2500))
    kind = it % 4
    if kind == 0:
        q, t = synth.correlated(nq, nt, it)[:2]
    elif kind == 1:
        q, t = synth.tie_stress(nq, it), synth.tie_stress(nt, it + 1)
    elif kind == 2:
        q, t = synth.uniform(nq, it), synth.duplicate_rows(max(1, nt // 3), it)
    else:
        q, t = synth.uniform(nq, it), synth.uniform(nt, it + 5)
    k = int(rng.choice([1, 2, 3, 5, 8]))
    mask = (rng.random((len(q), len(t))) < 0.3).astype(np.uint8) * int(rng.choice([1, 255])) if it % 3 == 0 else None
    ri, rd = ref.knn(q, t, k, mask)
    impls = (orc, c_oracle) if len(q) * len(t) <= 400_000 else (c_oracle,)     # the numpy restatement builds the full matrix
    for impl in impls:
        oi, od = impl.knn(q, t, k, mask)
        ok = np.array_equal(oi, ri) and np.array_equal(od, rd)
        checks += 1
        bad += not ok
    rq, rt, rdd = ref.match(q, t, cross_check=True)                              # cv2 refuses cross-check + mask
    for impl in impls:
        oq, ot, od = impl.cross_check(q, t)
        ok = np.array_equal(oq, rq) and np.array_equal(ot, rt) and np.array_equal(od, rdd)
        checks += 1
        bad += not ok
    a, b = orc.match(q, t, k=2, ratio=0.8) if orc in impls else None, ref.ratio_match(q, t, 0.8)
    if a is not None:
        checks += 1
        bad += not all(np.array_equal(x, y) for x, y in zip(a, b))
    gate = int(rng.integers(5, 60))
    if orc in impls:                                                             # the reference's caller-side gates
        m = cv2.BFMatcher_create(cv2.NORM_HAMMING, crossCheck=True).match(q, t)
        want = [(x.queryIdx, x.trainIdx) for x in m if x.distance < gate]        # slam/tracking.py:57
        got = orc.match(q, t, cross_check_=True, max_distance=gate, strict=True)
        checks += 1
        bad += list(zip(got[0].tolist(), got[1].tolist())) != want
print(f"oracle soak vs cv2 {cv2.__version__}: {iters} iterations, {checks} comparisons, {bad} mismatches, {time.time() - t0:.0f} s", flush=True)
sys.exit(1 if bad else 0)
