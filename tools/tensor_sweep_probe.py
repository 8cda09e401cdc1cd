"""BASELINE config 5 on the diagonal, both forms of the kernel: back-to-back time per call of a bound device call
(Engine.plan_device) with tensor=1 (POPC kernel) and tensor=0 (auto: the tensor form from 8 M pairs on), k = 1 and k = 2;
the two forms' tables are compared at every point."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import boslam_b200 as bb  # noqa: E402
from boslam_b200 import synth  # noqa: E402

eng = bb.Engine(0)
base, other = synth.uniform(65536, 7), synth.uniform(65536, 8)
st = torch.cuda.current_stream().cuda_stream
print("| n x n | k | POPC kernel G pairs/s (us) | tensor form G pairs/s (us) | launches | same tables |")
print("|---|---|---|---|---|---|")
for n in (2048, 4096, 8192, 16384, 32768, 65536):
    q, t = torch.from_numpy(base[:n]).cuda(), torch.from_numpy(other[:n]).cuda()
    tab = bb.make_problems([n], [n])
    for k in (1, 2):
        res, ms, li = {}, {}, {}
        for knob in (1, 0):
            eng.set_tuning(tensor=knob)
            o = eng.match_batched_device(q, t, tab, k=k, want_knn=True)
            torch.cuda.synchronize()
            res[knob] = (o["knn_idx"].clone(), o["knn_dist"].clone())
            li[knob] = eng.launch_info()
            plan = eng.plan_device(q, t, tab, k=k)
            reps = 30 if n <= 16384 else 5
            for _ in range(2):
                plan.run(st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                plan.run(st)
            e1.record()
            torch.cuda.synchronize()
            ms[knob] = e0.elapsed_time(e1) / reps
        same = torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
        print(f"| {n} | {k} | {n * n / ms[1] / 1e6:.0f} ({ms[1] * 1e3:.1f}) | {n * n / ms[0] / 1e6:.0f} ({ms[0] * 1e3:.1f}) | "
              f"{li[0]['kernels_launched']} ({'tensor' if li[0]['popc_mode'] == 0 else 'POPC'}, {li[0]['segments']} items) | {same} |", flush=True)
eng.set_tuning(tensor=0)
