"""Two engines on two threads, each running persistent launches in which EVERY CTA owns a finalize tile (a batch of many tiny
problems), at the same time: the CTAs of the two kernels share the SMs, so neither grid is fully resident.  Every call must
return the right lists, and none may take anywhere near the 2 s of a finalize tile's time-out."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import boslam_b200 as bb
from boslam_b200 import synth

P, N = 1300, 96
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 150
worst, bad = [0.0, 0.0], [0, 0]


def work(tid):
    eng = bb.Engine(0)
    eng.set_tuning(persistent=2)
    q = torch.from_numpy(synth.uniform(P * N, 10 + tid)).cuda()
    t = torch.from_numpy(synth.uniform(P * N, 20 + tid)).cuda()
    t[:P * N // 2] = q[:P * N // 2]          # half of the rows match themselves at distance 0
    tab = bb.make_problems([N] * P, [N] * P)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        plan = eng.plan_device(q, t, tab, cross_check=True, max_distance=10)
        for it in range(iters):
            t0 = time.perf_counter()
            out = plan.run()
            st.synchronize()
            worst[tid] = max(worst[tid], time.perf_counter() - t0)
            cnt = out["count"].cpu().numpy()[:P]
            want = np.where(np.arange(P) < P // 2, N, cnt)      # the first half: every row is its own mutual match
            if not np.array_equal(cnt, want) or eng.launch_info()["scan_grid"] < 1000:
                bad[tid] += 1


ths = [threading.Thread(target=work, args=(i,)) for i in range(2)]
for th in ths:
    th.start()
for th in ths:
    th.join()
print(f"shared GPU: {iters} launches per thread, wrong results {bad}, slowest call {max(worst) * 1e3:.1f} ms")
sys.exit(1 if sum(bad) or max(worst) > 0.5 else 0)
