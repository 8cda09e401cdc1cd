"""Raw timeline of one call (bfm_debug_timeline): per-CTA stamps relative to the first entry, CTAs per SM, histograms.
usage: timeline_dump.py single NQ NT [k2|cross]   |   timeline_dump.py batch PAIRS N [k2|cross]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import boslam_b200 as bb
from boslam_b200 import synth

kind, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
mode = sys.argv[4] if len(sys.argv) > 4 else "k2"
knobs = dict(kv.split("=") for kv in sys.argv[5:])
kw = dict(k=2, ratio=0.8) if mode == "k2" else dict(cross_check=True, max_distance=30)
eng = bb.Engine(0)
eng.set_tuning(**{k: int(v) for k, v in knobs.items()})
if kind == "single":
    q, t = synth.uniform(a, 7), synth.uniform(b, 8)
    tab = bb.make_problems([a], [b])
else:
    q, t = synth.uniform(a * b, 7), synth.uniform(a * b, 8)
    tab = bb.make_problems([b] * a, [b] * a)
q, t = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
CAP = 16384
buf = torch.zeros((CAP, 8), dtype=torch.int64, device="cuda")
for _ in range(5):
    eng.match_batched_device(q, t, tab, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = eng.match_batched_device(q, t, tab, **kw)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    eng.match_batched_device(q, t, tab, out=out, **kw)
e1.record()
torch.cuda.synchronize()
print(f"{kind} {a} {b} {mode} {knobs}: back-to-back {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call; launch {eng.launch_info()}")
eng._lib.bfm_debug_timeline(eng._h, ctypes.c_void_p(buf.data_ptr()), CAP)
eng.match_batched_device(q, t, tab, out=out, **kw)
torch.cuda.synchronize()
eng._lib.bfm_debug_timeline(eng._h, None, 0)
x = buf.cpu().numpy()
x = x[x[:, 0] > 0]
t0 = x[:, 0].min()
rel = np.where(x[:, :7] > 0, (x[:, :7] - t0) / 1e3, np.nan)
smid = x[:, 7]
print(f"{len(x)} CTAs on {len(np.unique(smid))} SMs; CTAs per SM: min {np.bincount(smid).min()} max {np.bincount(smid).max()}")
names = ["entry", "inputs", "landed", "scan1", "commit1", "items_done", "tiles_done"]
for i, n in enumerate(names):
    c = rel[:, i][~np.isnan(rel[:, i])]
    if len(c):
        print(f"{n:11s} n={len(c):5d} min {c.min():8.1f} p10 {np.percentile(c, 10):8.1f} med {np.median(c):8.1f} p90 {np.percentile(c, 90):8.1f} max {c.max():8.1f} us")
ent = rel[:, 0]
print("entry histogram (us):", np.histogram(ent, bins=10)[0].tolist(), "edges", np.round(np.histogram(ent, bins=10)[1], 1).tolist())
late = np.argsort(ent)[-8:]
print("latest entries: CTA", late.tolist(), "at", np.round(ent[late], 1).tolist(), "on SM", smid[late].tolist())
end = np.nanmax(rel, axis=1)
print("end histogram (us):", np.histogram(end, bins=10)[0].tolist(), "edges", np.round(np.histogram(end, bins=10)[1], 1).tolist())
# concurrency: CTAs alive per SM at the median time
mid = np.median(end) / 2
alive = np.bincount(smid[(ent <= mid) & (end >= mid)], minlength=148)
print(f"CTAs alive per SM at t = {mid:.1f} us: min {alive.min()} med {np.median(alive)} max {alive.max()}")
