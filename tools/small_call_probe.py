"""Where the end-to-end time of a small (single-frame) call goes: Python layer vs C call vs device."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import boslam_b200 as bb  # noqa: E402
from boslam_b200 import _ffi, synth  # noqa: E402

eng = bb.Engine(0)


def timeit(f, n=200):
    for _ in range(20):
        f()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    return (time.perf_counter() - t0) / n * 1e6


for name, (nq, nt), kw in (("f2f 1000x1000 cc gate", (1000, 1000), dict(cross_check=True, max_distance=30, strict=True)),
                           ("track 2000x20000 cc gate", (2000, 20000), dict(cross_check=True, max_distance=30)),
                           ("track 2000x20000 k2 ratio", (2000, 20000), dict(k=2, ratio=0.8))):
    q, t, _ = synth.correlated(nq, nt, 3)
    full = timeit(lambda: eng.match(q, t, **kw))
    # C call only, pageable in / pageable out
    opts, _ = eng._options(kw.get("k", 1), kw.get("ratio"), kw.get("cross_check", False), kw.get("max_distance"), kw.get("strict", False))
    mq, mt, md, mc = (np.empty(nq, np.int32) for _ in range(3)), None, None, None
    mq = np.empty(nq, np.int32); mt = np.empty(nq, np.int32); md = np.empty(nq, np.int32); mc = np.zeros(1, np.int32)
    probs = np.array([[0, nq, 0, nt, 0, 0]], np.int32)
    args = (_ffi.MEM_HOST, q.ctypes.data, nq, t.ctypes.data, nt, probs, nq, opts, None,
            (mq.ctypes.data, mt.ctypes.data, md.ctypes.data, mc.ctypes.data))
    ccall = timeit(lambda: eng._call(*args))
    # pinned in / pinned out
    pq, pt = bb.PinnedBuffer(q.shape), bb.PinnedBuffer(t.shape)
    pq.array[...] = q; pt.array[...] = t
    hb = bb.HostBatchBuffers(nq, 1)
    args2 = (_ffi.MEM_HOST, pq.array.ctypes.data, nq, pt.array.ctypes.data, nt, probs, nq, opts, None,
             (hb.m[0].ctypes.data, hb.m[1].ctypes.data, hb.m[2].ctypes.data, hb.count.ctypes.data))
    pinned = timeit(lambda: eng._call(*args2))
    eng.set_tuning(timing=1)
    import torch
    qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    tab = bb.make_problems([nq], [nt])
    ks = []
    for _ in range(20):
        eng.match_batched_device(qd, td, tab, **kw)
        ks.append(eng.launch_info()["scan_ms"] * 1e3)
    eng.set_tuning(timing=0)
    print(f"{name:28s} Engine.match {full:7.1f} us | C call pageable {ccall:7.1f} us | C call pinned {pinned:7.1f} us | kernel {np.median(ks):6.1f} us", flush=True)
os.environ["BFM_TRACE"] = "1"
q, t, _ = synth.correlated(2000, 20000, 3)
for _ in range(3):
    eng.match(q, t, cross_check=True, max_distance=30)

# window (binned) vs brute force
os.environ.pop("BFM_TRACE", None)
q, t, qxy, txy, _ = synth.window_scene(2000, 20000, 12)
import torch
qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
qxyd, txyd = torch.from_numpy(qxy).cuda(), torch.from_numpy(txy).cuda()
tab = bb.make_problems([2000], [20000])
for wb in (0, 1):
    eng.set_tuning(window_bins=wb, timing=1)
    ks = []
    for _ in range(20):
        eng.match_batched_device(qd, td, tab, k=2, ratio=0.8, window=(qxyd, txyd, 15.0))
        ks.append(eng.launch_info()["scan_ms"] * 1e3)
    eng.set_tuning(timing=0)
    e2e = timeit(lambda: eng.match(q, t, k=2, ratio=0.8, window=(qxy, txy, 15.0)))
    print(f"track 2000x20000 window r=15 window_bins={wb}: kernels {np.median(ks):6.1f} us, Engine.match {e2e:6.1f} us", flush=True)
eng.set_tuning(window_bins=0)
sc = synth.local_map_scene(20000, 20000, 2000, seed=14)
store = bb.MapStore(20000, engine=eng)
store.update(np.arange(20000), sc["desc"], sc["pt3d"], sc["normal"])
targs = (sc["des"], sc["kp"], sc["R"], sc["t"], sc["see_vector"], sc["edges"])
print(f"MapStore.track cross-check gate30: {timeit(lambda: store.track(*targs)):6.1f} us;  window r=15 + ratio: "
      f"{timeit(lambda: store.track(*targs, cross_check=False, k=2, ratio=0.8, max_distance=None, window_radius=15.0)):6.1f} us")
