import numpy as np, threading, time, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import boslam_b200 as bb
n = 32 << 20
src = np.random.default_rng(0).integers(0, 256, n, dtype=np.uint8)
pin = bb.PinnedBuffer((n,), np.uint8).array
dst2 = np.empty(n, np.uint8)
for target, name in ((pin, "pageable->pinned"), (dst2, "pageable->pageable")):
    for nt in (1, 2, 4, 8, 16):
        sl = [(i * n // nt, (i + 1) * n // nt) for i in range(nt)]
        def work(a, b): np.copyto(target[a:b], src[a:b])
        best = 1e9
        for rep in range(5):
            th = [threading.Thread(target=work, args=s) for s in sl]
            t0 = time.perf_counter(); [x.start() for x in th]; [x.join() for x in th]; dt = time.perf_counter() - t0
            best = min(best, dt)
        print(f"{name} threads={nt:2d}: {best*1e3:.3f} ms  {n/best/1e9:.1f} GB/s", flush=True)
print("cpus", os.cpu_count())
