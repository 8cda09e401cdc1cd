"""Does a concurrent H2D copy slow the scan kernel?  (debugging the pipelined host path)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import boslam_b200 as bb
from boslam_b200 import synth

N = 2000
eng = bb.Engine(0)
hp = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
dd = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
copy_stream = torch.cuda.Stream()
for npairs in (32, 64, 256):
    q, t = synth.keyframe_pair_batch(npairs, N, 1)
    qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    tb = bb.make_problems([N] * npairs, [N] * npairs)
    for concurrent in (False, True):
        for _ in range(3):
            eng.match_batched_device(qd, td, tb, k=2, ratio=0.8)
        torch.cuda.synchronize()
        if concurrent:
            with torch.cuda.stream(copy_stream):
                dd.copy_(hp, non_blocking=True)   # ~4.9 ms of H2D traffic
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            eng.match_batched_device(qd, td, tb, k=2, ratio=0.8)
        e1.record()
        torch.cuda.synchronize()
        print(f"{npairs:3d} pairs, concurrent H2D={concurrent}: {e0.elapsed_time(e1) / 4:.3f} ms per call")
