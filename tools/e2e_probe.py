"""Where does the end-to-end time of the batched host path go?  (H2D bandwidth, chunking sweep.)"""
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
out = bb.HostBatchBuffers(P * N, P, k=2)

# raw copy bandwidth, torch pinned
hp = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
for direction in ("h2d", "d2h"):
    for _ in range(3):
        (d.copy_(hp, non_blocking=True) if direction == "h2d" else hp.copy_(d, non_blocking=True))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        (d.copy_(hp, non_blocking=True) if direction == "h2d" else hp.copy_(d, non_blocking=True))
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"{direction} 64 MiB pinned: {dt * 1e3:.3f} ms  {(64 << 20) / dt / 1e9:.1f} GB/s")

for chunks in (1, 2, 4, 8, 16):
    eng.set_tuning(pipeline_chunks=chunks)
    for _ in range(3):
        eng.match_batched(pq.array, pt.array, tab, k=2, ratio=0.8, out=out)
    t0 = time.perf_counter()
    for _ in range(20):
        eng.match_batched(pq.array, pt.array, tab, k=2, ratio=0.8, out=out)
    dt = (time.perf_counter() - t0) / 20
    print(f"pipeline_chunks={chunks:2d}: {dt * 1e3:.3f} ms/step  {P * N * N / dt / 1e9:.0f} Gpair/s  launches/call={eng.launch_info()['kernels_launched']}")
# pageable inputs / outputs for comparison
eng.set_tuning(pipeline_chunks=0)
for _ in range(3):
    eng.match_batched(q, t, tab, k=2, ratio=0.8)
t0 = time.perf_counter()
for _ in range(10):
    eng.match_batched(q, t, tab, k=2, ratio=0.8)
dt = (time.perf_counter() - t0) / 10
print(f"pageable in/out: {dt * 1e3:.3f} ms/step")

# --- split: python wrapper vs C call ------------------------------------------------------------
import ctypes
from boslam_b200 import _ffi
opts, _ = eng._options(2, 0.8, False, None, False)
probs = np.ascontiguousarray(tab, np.int32)
args = (_ffi.MEM_HOST, pq.array.ctypes.data, P * N, pt.array.ctypes.data, P * N, probs, P * N, opts, None,
        (out.m[0].ctypes.data, out.m[1].ctypes.data, out.m[2].ctypes.data, out.count.ctypes.data))
for chunks in (1, 8):
    eng.set_tuning(pipeline_chunks=chunks)
    for _ in range(3):
        eng._call(*args)
    t0 = time.perf_counter()
    for _ in range(20):
        eng._call(*args)
    dt = (time.perf_counter() - t0) / 20
    print(f"C call only, chunks={chunks}: {dt * 1e3:.3f} ms")
# device-only time of the same batch for reference (data resident)
qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
for npairs in (256, 64, 32, 16):
    tb = bb.make_problems([N] * npairs, [N] * npairs)
    for waves in (0, 2, 1):
        eng.set_tuning(waves=waves)
        for _ in range(3):
            eng.match_batched_device(qd, td, tb, k=2, ratio=0.8)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eng.match_batched_device(qd, td, tb, k=2, ratio=0.8)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        li = eng.launch_info()
        print(f"device {npairs:3d} pairs waves={waves}: {ms:.3f} ms  {npairs * N * N / ms / 1e6:.0f} Gpair/s grid={li['scan_grid']} seg_rows={li['train_rows_per_segment']}")
eng.set_tuning(waves=0)
