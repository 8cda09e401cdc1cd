import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import boslam_b200 as bb
from boslam_b200 import synth
from oracle import c_oracle

eng = bb.Engine(0)
qp, tp = synth.keyframe_pair_batch(24, 1500, seed=77)
q0 = qp[:1500]
tab2 = bb.make_problems([1500] * 24, [1500] * 24, shared_query=True)
want = [c_oracle.cross_check(q0, tp[p * 1500:(p + 1) * 1500]) for p in range(24)]
for chunks in (1, 4, 2, 8):
    for r in (0, 2, 4):
        eng.set_tuning(pipeline_chunks=chunks, queries_per_thread=r)
        bad = {}
        for it in range(15):
            res = eng.match_batched(q0, tp, tab2, cross_check=True)
            for p in range(24):
                g = res[p]
                if not (np.array_equal(g[0], want[p][0]) and np.array_equal(g[1], want[p][1])):
                    bad.setdefault(p, 0)
                    bad[p] += 1
        print(f"chunks={chunks} R={r}: bad problems {bad} info={eng.launch_info()['train_rows_per_segment']}", flush=True)
