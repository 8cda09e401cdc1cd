"""Where does the time of ONE single-problem call go?  bfm_debug_timeline makes every CTA of the matching kernel stamp
%globaltimer at the points of its life; this prints, per shape and mode, the span of the call on the GPU clock and
the phase medians, next to the CUDA-event time of the same call.

  entry   : CTA start, relative to the first CTA's start (ramp of the launch)
  in      : entry -> inputs requested (segment table read, TMA issued, queries loaded)
  land    : -> first train chunk in shared memory
  scan    : -> scan loop done
  commit  : -> row keys committed (atomics issued)
  count   : -> completion counter bumped (threadfence + atomic round trip)
  final   : -> finalize done (only the finalizing CTA / the tile-parallel kernels)
"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import boslam_b200 as bb
from boslam_b200 import synth

CAP = 16384
eng = bb.Engine(0)
buf = torch.zeros((CAP, 8), dtype=torch.int64, device="cuda")
base, other = synth.uniform(32768, 7), synth.uniform(32768, 8)
shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]] or [(600, 600), (1000, 1000), (2000, 2000), (2000, 8000), (2000, 20000), (4096, 4096)]


def med(a):
    return float(np.median(a)) if len(a) else float("nan")


for nq, nt in shapes:
    q, t = torch.from_numpy(base[:nq]).cuda(), torch.from_numpy(other[:nt]).cuda()
    tab = bb.make_problems([nq], [nt])
    for mode, kw in (("k2+ratio", dict(k=2, ratio=0.8)), ("cross+gate", dict(cross_check=True, max_distance=30))):
        eng._lib.bfm_debug_timeline(eng._h, None, 0)
        eng.set_tuning(timing=1)
        ts = []
        for _ in range(14):
            eng.match_batched_device(q, t, tab, **kw)
            ts.append(eng.launch_info()["scan_ms"])
        ev_us = np.median(ts[3:]) * 1e3
        eng.set_tuning(timing=0)
        # back-to-back calls without a host sync: the steady-state cost per call as the stream sees it
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            eng.match_batched_device(q, t, tab, **kw)
        e1.record()
        torch.cuda.synchronize()
        b2b_us = e0.elapsed_time(e1) * 1e3 / 50
        li = eng.launch_info()
        spans, rows = [], []
        for rep in range(5):
            buf.zero_()
            eng._lib.bfm_debug_timeline(eng._h, ctypes.c_void_p(buf.data_ptr()), CAP)
            eng.match_batched_device(q, t, tab, **kw)
            torch.cuda.synchronize()
            a = buf.cpu().numpy()
            used = a[:, 0] > 0
            a = a[used]
            t0 = a[:, 0].min()
            end = a[a > 0].max()
            spans.append((end - t0) / 1e3)
            rows.append((a, t0))
        a, t0 = rows[len(rows) // 2]
        n_scan = li["scan_grid"]
        scan = a[:n_scan]
        fin = a[n_scan:]
        ph = lambda i, j: med((scan[:, j] - scan[:, i])[(scan[:, j] > 0) & (scan[:, i] > 0)]) / 1e3
        entry = (scan[:, 0] - t0) / 1e3
        last_scan_end = (scan[:, 1:6].max() - t0) / 1e3
        finalizer = scan[scan[:, 6] > 0]
        line = (f"{nq:5d} x {nt:5d} {mode:10s} events {ev_us:6.1f} us  back-to-back {b2b_us:6.1f} us | GPU span {np.median(spans):6.1f} us "
                f"(min {min(spans):.1f}) | R{li['queries_per_thread']} grid {n_scan} rows/seg {li['train_rows_per_segment']} kernels {li['kernels_launched']} | "
                f"entry med {med(entry):.1f} max {entry.max():.1f} | in {ph(0, 1):.2f} land {ph(1, 2):.2f} scan {ph(2, 3):.2f} commit {ph(3, 4):.2f} "
                f"count {ph(4, 5):.2f} | scan CTAs end {last_scan_end:.1f}")
        if len(finalizer):
            f = finalizer[0]
            line += f" | finalizer {(f[5] - t0) / 1e3:.1f} -> {(f[6] - t0) / 1e3:.1f}"
        if len(fin):
            line += f" | fin kernels: first entry {(fin[:, 0].min() - t0) / 1e3:.1f}, last exit {(fin[:, 6].max() - t0) / 1e3:.1f} ({len(fin)} CTAs)"
        print(line, flush=True)
eng._lib.bfm_debug_timeline(eng._h, None, 0)
