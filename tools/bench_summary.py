"""One-screen summary of bench.py JSON lines (files given on the command line)."""
import json, sys
for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    e, r = d.get("e2e") or {}, d.get("e2e_resident") or {}
    print(f"{f}: N={d['n_gpus']} value {d['value'] / 1e9:.1f} G pairs/s ({d['ms_per_step'] * 1e3:.1f} us/step) roofline {d['roofline']['bound']} frac {d['roofline']['frac']:.3f} | "
          f"e2e {e.get('value', 0) / 1e9:.1f} G ({e.get('ms_per_step', 0) * 1e3:.1f} us) | resident {r.get('value', 0) / 1e9:.1f} G ({r.get('ms_per_step', 0) * 1e3:.1f} us) | "
          f"verified {d.get('gather_verified_full')} launch {d['launch']}")
    if d.get("weak"):
        print("   weak:", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["weak"].items() if k != "note"})
    if d.get("sustained"):
        print("   sustained: %.1f G pairs/s over %.1f s" % (d["sustained"]["value"] / 1e9, d["sustained"]["seconds"]))
    for k, v in (d.get("frames") or {}).items():
        keys = ("ms_e2e", "ms_device", "ms_device_idle_stream", "us_gpu_span", "kernels_per_call", "grid", "work_items", "queries_per_thread")
        print("   ", k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a in keys})
