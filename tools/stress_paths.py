"""Randomised cross-path stress: every way of running the same batch must return the same bits.
host path (1 chunk / gated with 2..24 chunks, pinned and pageable buffers), device path, KeyframeBank;
modes k=1 / k=2+ratio / cross-check / k=3; ragged problem sizes including empty ones."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import boslam_b200 as bb  # noqa: E402
from boslam_b200 import synth  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 60
eng = bb.Engine(0)
rng = np.random.default_rng(2024)
bad = 0
for it in range(iters):
    P = int(rng.integers(1, 40))
    qn = rng.integers(0, 900, P)
    tn = rng.integers(0, 1200, P)
    if rng.random() < 0.3:
        qn[:] = int(rng.integers(1, 2500))
        tn[:] = int(rng.integers(1, 2500))
    q = synth.uniform(int(qn.sum()) + 1, it)[:int(qn.sum())]
    t = synth.uniform(int(tn.sum()) + 1, it + 7777)[:int(tn.sum())]
    # plant some true matches
    for p in range(P):
        if qn[p] and tn[p]:
            n = min(int(qn[p]), int(tn[p])) // 2
            qo, to = int(qn[:p].sum()), int(tn[:p].sum())
            q[qo:qo + n] = t[to:to + n]
            q[qo:qo + n, 0] ^= rng.integers(0, 4, n).astype(np.uint8)
    tab = bb.make_problems(qn.tolist(), tn.tolist())
    mode = it % 4
    kw = [dict(k=1, max_distance=40), dict(k=2, ratio=0.8), dict(cross_check=True), dict(k=3)][mode]
    want_knn = mode != 2
    ref = None
    runs = []
    for chunks in (1, int(rng.integers(2, 25))):
        knobs = dict(pipeline_chunks=chunks, feeders=[0, -1, int(rng.integers(1, 33))][it % 3], host_threads=[0, 0, -1, 3][it % 4])
        if os.environ.get("STRESS_VERBOSE"):
            print(f"iter {it}: P={P} mode={mode} rows={int(qn.sum())}x{int(tn.sum())} {knobs}", flush=True)
        eng.set_tuning(**knobs)
        runs.append(("host pageable chunks=%d" % chunks, eng.match_batched(q, t, tab, want_knn=want_knn, **kw)))
        n_out = int(qn.sum())
        if n_out and len(q) and len(t):
            pq, pt = bb.PinnedBuffer(q.shape), bb.PinnedBuffer(t.shape)
            pq.array[...] = q
            pt.array[...] = t
            ob = bb.HostBatchBuffers(n_out, P, k=kw.get("k", 1), want_knn=want_knn)
            r = eng.match_batched(pq.array, pt.array, tab, want_knn=want_knn, out=ob, **kw)
            runs.append(("host pinned chunks=%d" % chunks, r))
    eng.set_tuning(pipeline_chunks=0, feeders=0, host_threads=0)
    if len(q) and len(t):
        runs.append(("device", eng.match_batched(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), tab, want_knn=want_knn, **kw)))

    def norm(r):
        if want_knn:
            idx, dist, res = r
            idx = idx.cpu().numpy() if hasattr(idx, "cpu") else np.asarray(idx)
            dist = dist.cpu().numpy() if hasattr(dist, "cpu") else np.asarray(dist)
            n = int(qn.sum())
            head = (idx[:n].tobytes(), dist[:n].tobytes())
        else:
            res, head = r, ()
        return head + tuple(b"".join(a.tobytes() for a in res[p]) for p in range(P)) + (res.counts[:P].tobytes(),)

    sigs = [(name, norm(r)) for name, r in runs]
    for name, s in sigs[1:]:
        if s != sigs[0][1]:
            bad += 1
            print(f"iter {it}: {name} differs from {sigs[0][0]} (P={P}, mode={mode})", flush=True)
print(f"stress: {iters} iterations, {bad} mismatches", flush=True)
sys.exit(1 if bad else 0)
