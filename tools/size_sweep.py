"""BASELINE config 5: brute-force sweep Q, T in {1k .. 64k}^2 on uniform synthetic descriptors, k = 1, k = 2 and
cross-check; device-resident inputs.  Two device times per point: `kernel_ms` = CUDA events around ONE launch on an idle
stream (the round-1 figure; it includes ~9 us of event / launch latency) and `b2b_ms` = per call in a back-to-back loop of
bound calls (Engine.plan_device), what a loop of such calls costs.  Also checks size-
independent properties at every point (self-match: knn(t, t) finds row i at distance 0; cross-check of a
set with itself is the identity on de-duplicated rows)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import boslam_b200 as bb  # noqa: E402
from boslam_b200 import synth  # noqa: E402

SIZES = [1024, 4096, 16384, 65536]
if "--full" in sys.argv:      # every power of two, as BASELINE config 5 words it
    SIZES = [1024 << i for i in range(7)]


def main():
    eng = bb.Engine(0)
    base = synth.uniform(65536, 7)
    other = synth.uniform(65536, 8)
    out = []
    try:
        import cv2
    except Exception:
        cv2 = None
    for nq in SIZES:
        for nt in SIZES:
            q, t = torch.from_numpy(base[:nq]).cuda(), torch.from_numpy(other[:nt]).cuda()
            tab = bb.make_problems([nq], [nt])
            for mode, kw in (("k1", dict(k=1)), ("k2", dict(k=2)), ("cross", dict(cross_check=True))):
                ts = []
                eng.set_tuning(timing=1)
                for _ in range(5):
                    eng.match_batched_device(q, t, tab, **kw)
                    ts.append(eng.launch_info()["scan_ms"])
                eng.set_tuning(timing=0)     # (timing synchronises after every launch: off for the back-to-back loop)
                li = eng.launch_info()
                ms = float(np.median(ts[1:]))
                plan = eng.plan_device(q, t, tab, **kw)
                reps = 40 if nq * nt <= (1 << 28) else 6
                st = torch.cuda.current_stream().cuda_stream
                for _ in range(2):
                    plan.run(st)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    plan.run(st)
                e1.record()
                torch.cuda.synchronize()
                b2b = e0.elapsed_time(e1) / reps
                rec = dict(Q=nq, T=nt, mode=mode, kernel_ms=ms, gpairs=nq * nt / ms / 1e6, b2b_ms=b2b, gpairs_b2b=nq * nt / b2b / 1e6,
                           grid=li["scan_grid"], items=li["segments"], R=li["queries_per_thread"], seg_rows=li["train_rows_per_segment"],
                           kernels=li["kernels_launched"])
                if cv2 is not None and nq == nt and nq <= 16384 and mode in ("k2", "cross"):
                    m = cv2.BFMatcher_create(cv2.NORM_HAMMING, crossCheck=(mode == "cross"))
                    t0 = time.perf_counter()
                    (m.match(base[:nq], other[:nt]) if mode == "cross" else m.knnMatch(base[:nq], other[:nt], 2))
                    rec["cv2_ms"] = (time.perf_counter() - t0) * 1e3
                    rec["cv2_threads"] = cv2.getNumThreads()
                out.append(rec)
                print(json.dumps(rec), flush=True)
            if nq == nt:  # properties
                idx, dist = eng.knn(q, q, 1)
                ok_self = bool((dist[:, 0] == 0).all().item()) and bool((idx[:, 0] == torch.arange(nq, device=idx.device)).all().item())
                qi, ti, d = eng.match(q, q, cross_check=True)
                ok_cc = len(qi) == nq and bool((qi == ti).all().item()) and bool((d == 0).all().item())
                print(json.dumps(dict(Q=nq, T=nt, self_match_ok=ok_self, self_cross_check_identity=ok_cc)), flush=True)
                out.append(dict(Q=nq, T=nt, self_match_ok=ok_self, self_cross_check_identity=ok_cc))
                assert ok_self and ok_cc
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/size_sweep_full.json" if "--full" in sys.argv else "gpurun_out/size_sweep.json", "w"), indent=1)
    with open("gpurun_out/size_sweep.md", "w") as f:
        f.write("G pairs/s per call, back to back (in brackets: from CUDA events around one launch on an idle stream, the round-1 figure).  Rows: Q, columns: T.\n")
        for mode, title in (("k1", "k = 1"), ("k2", "k = 2"), ("cross", "cross-check")):
            f.write(f"\n## {title}\n\n| Q \\ T | " + " | ".join(str(n) for n in SIZES) + " |\n|---|" + "---|" * len(SIZES) + "\n")
            for nq in SIZES:
                row = [next(r for r in out if r.get("mode") == mode and r["Q"] == nq and r["T"] == nt) for nt in SIZES]
                f.write(f"| {nq} | " + " | ".join(f"{r['gpairs_b2b']:.0f} ({r['gpairs']:.0f})" for r in row) + " |\n")
        f.write("\n## cv2 4.13 on the same box, same arrays, one call\n\n| Q = T | mode | cv2 ms | engine ms (back to back) | ratio |\n|---|---|---|---|---|\n")
        for r in out:
            if "cv2_ms" in r:
                f.write(f"| {r['Q']} | {r['mode']} | {r['cv2_ms']:.1f} | {r['b2b_ms']:.3f} | {r['cv2_ms'] / r['b2b_ms']:.0f}x |\n")


if __name__ == "__main__":
    main()
