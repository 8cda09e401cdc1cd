"""The batched host path of the tensor form (copy-engine chunks + per-chunk launches, bfm_pipeline.cuh): time per 256-pair
step against the SM-fed POPC path (tensor=1) and the resident call, results compared; BFM_TRACE=1 prints the timeline."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import boslam_b200 as bb  # noqa: E402
from boslam_b200 import synth  # noqa: E402

P, N = 256, 2000
eng = bb.Engine(0)
q, t = synth.keyframe_pair_batch(P, N, 1)
pq, pt = bb.PinnedBuffer(q.shape), bb.PinnedBuffer(t.shape)
pq.array[...] = q
pt.array[...] = t
tab = bb.make_problems([N] * P, [N] * P)
outs = {}
for name, knob, chunks in (("sm-fed popc", 1, 0), ("tensor chunks", 0, 0)):
    eng.set_tuning(tensor=knob, tensor_chunks=chunks)
    out = bb.HostBatchBuffers(P * N, P, k=2)
    for _ in range(3):
        eng.match_batched(pq.array, pt.array, tab, k=2, ratio=0.8, out=out)
    t0 = time.perf_counter()
    for _ in range(20):
        eng.match_batched(pq.array, pt.array, tab, k=2, ratio=0.8, out=out)
    dt = (time.perf_counter() - t0) / 20
    li = eng.launch_info()
    print(f"{name:14s}: {dt * 1e3:.3f} ms/step  {P * N * N / dt / 1e9:.0f} G pairs/s  chunks={li['copy_chunks']} launches={li['kernels_launched']}", flush=True)
    outs[name] = (out.count.copy(), [m.copy() for m in out.m])
a, b = outs["sm-fed popc"], outs["tensor chunks"]
same = np.array_equal(a[0], b[0])
for p in range(P):
    n = int(a[0][p])
    for x, y in zip(a[1], b[1]):
        same = same and np.array_equal(x[p * N:p * N + n], y[p * N:p * N + n])
print("identical match lists:", same, "matches:", int(a[0].sum()))
# how fast does the copy engine read these very buffers (NUMA placement of the pinned pages matters)?
import glob
tq, tt = torch.from_numpy(pq.array), torch.from_numpy(pt.array)
dq, dt_ = torch.empty_like(tq, device="cuda"), torch.empty_like(tt, device="cuda")
for chunks in (1, 7):
    rows = (P * N + chunks - 1) // chunks
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for c in range(chunks):
            dq[c * rows:(c + 1) * rows].copy_(tq[c * rows:(c + 1) * rows], non_blocking=True)
            dt_[c * rows:(c + 1) * rows].copy_(tt[c * rows:(c + 1) * rows], non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"plain copies of the same pinned arrays, {chunks} chunk(s): {best * 1e3:.3f} ms  {(tq.numel() + tt.numel()) / best / 1e9:.1f} GB/s")
print("affinity", len(os.sched_getaffinity(0)), "nodes", [open(f).read().strip() for f in sorted(glob.glob("/sys/devices/system/node/node*/cpulist"))],
      "gpu numa", [open(f).read().strip() for f in glob.glob("/sys/bus/pci/devices/*/numa_node") if open(f.replace("numa_node", "class")).read().startswith("0x0302")][:8])
os.environ["BFM_TRACE"] = "1"
for chunks in (4, 4, 7, 2):
    eng.set_tuning(tensor_chunks=chunks)
    eng.match_batched(pq.array, pt.array, tab, k=2, ratio=0.8, out=out)
