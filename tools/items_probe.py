"""Persistent form against the static form (persistent = 1), and against item lengths / register tiles."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import boslam_b200 as bb
from boslam_b200 import synth

eng = bb.Engine(0)
cases = [("batch 32 x 2000^2 k2", [2000] * 32, [2000] * 32, dict(k=2, ratio=0.8)),
         ("batch 64 x 2000^2 k2", [2000] * 64, [2000] * 64, dict(k=2, ratio=0.8)),
         ("batch 20 x 2000^2 k2", [2000] * 20, [2000] * 20, dict(k=2, ratio=0.8)),
         ("batch 20 x 2000^2 cross", [2000] * 20, [2000] * 20, dict(cross_check=True, max_distance=30)),
         ("single 2000 x 20000 k2", [2000], [20000], dict(k=2, ratio=0.8)),
         ("single 2000 x 20000 cross", [2000], [20000], dict(cross_check=True, max_distance=30)),
         ("single 4096^2 k2", [4096], [4096], dict(k=2, ratio=0.8)),
         ("single 8192^2 k2", [8192], [8192], dict(k=2, ratio=0.8)),
         ("single 2048^2 k2", [2048], [2048], dict(k=2, ratio=0.8)),
         ("single 1000^2 cross", [1000], [1000], dict(cross_check=True, max_distance=30))]
knobsets = [dict(persistent=1), dict(persistent=2), dict(persistent=2, queries_per_thread=4), dict(persistent=2, queries_per_thread=2), dict(persistent=2, queries_per_thread=1),
            dict(persistent=2, ctas_per_sm=6), dict(persistent=2, ctas_per_sm=4)]
for name, qs, ts, kw in cases:
    q = torch.from_numpy(synth.uniform(sum(qs), 7)).cuda()
    t = torch.from_numpy(synth.uniform(sum(ts), 8)).cuda()
    tab = bb.make_problems(qs, ts)
    line = f"{name:26s}"
    for knobs in knobsets:
        eng.set_tuning(segment_rows=0, persistent=0, queries_per_thread=0, taper=0, waves=0, ctas_per_sm=0)
        eng.set_tuning(**knobs)
        plan = eng.plan_device(q, t, tab, **kw)
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            plan.run(st)
        torch.cuda.synchronize()
        n = 20 if len(qs) > 1 else 100
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            plan.run(st)
        e1.record()
        torch.cuda.synchronize()
        li = eng.launch_info()
        tag = ",".join(f"{k[0]}={v}" for k, v in knobs.items())
        line += f" | {tag}: {e0.elapsed_time(e1) / n * 1e3:7.1f} (R{li['queries_per_thread']} {li['segments']})"
    print(line, flush=True)
