"""Generates tests/golden/bfmatcher_golden.npz from the live reference implementation.

The reference's matcher IS cv2.BFMatcher (slam/tracking.py:45,56,121), so golden vectors are the
outputs of the installed cv2 (4.13.0 at generation time) on seeded synthetic descriptors.  Run
from the repo root:  python tests/golden/make_golden.py
Inputs and outputs are both stored so the fixtures travel to boxes without cv2.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from boslam_b200 import synth  # noqa: E402
from oracle import cv2_reference as ref  # noqa: E402


def cases():
    q, t, _ = synth.correlated(96, 130, seed=11)
    yield "correlated", q, t
    yield "uniform", synth.uniform(70, 12), synth.uniform(45, 13)
    yield "tie_stress", synth.tie_stress(80, 14), synth.tie_stress(120, 15)
    d = synth.duplicate_rows(40, 16)
    yield "dup_train", synth.correlated(64, len(d), 17)[0], d
    yield "dup_both", d[::-1].copy(), d
    yield "single_train", synth.uniform(9, 18), synth.uniform(1, 19)
    yield "two_train", synth.uniform(9, 20), synth.uniform(2, 21)
    # the reference's own fabricated descriptors: utils.int2orb(i) (utils.py:53-55)
    def int2orb(i):
        np.random.seed(i)
        return np.random.randint(256, size=32).astype(np.uint8)
    a = np.stack([int2orb(i) for i in range(54)])
    b = np.stack([int2orb(i) for i in range(20, 74)])
    yield "int2orb", a, b


def main():
    out = {"cv2_version": np.array(ref.version())}
    names = []
    for name, q, t in cases():
        names.append(name)
        out[f"{name}/q"] = q
        out[f"{name}/t"] = t
        qi, ti, d = ref.match(q, t, cross_check=True)
        out[f"{name}/cc_q"], out[f"{name}/cc_t"], out[f"{name}/cc_d"] = qi, ti, d
        for k in (1, 2, 3):
            idx, dist = ref.knn(q, t, k)
            out[f"{name}/knn{k}_idx"], out[f"{name}/knn{k}_dist"] = idx, dist
        rng = np.random.default_rng(len(name) * 7919)
        mask = (rng.random((len(q), len(t))) < 0.3).astype(np.uint8) * rng.choice([1, 255], (len(q), len(t))).astype(np.uint8)
        mask[0, :] = 0  # a fully masked query (rule R4)
        out[f"{name}/mask"] = mask
        idx, dist = ref.knn(q, t, 2, mask)
        out[f"{name}/mknn2_idx"], out[f"{name}/mknn2_dist"] = idx, dist
        rq, rt, rd = ref.ratio_match(q, t, 0.8)
        out[f"{name}/ratio_q"], out[f"{name}/ratio_t"], out[f"{name}/ratio_d"] = rq, rt, rd
    out["names"] = np.array(names)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bfmatcher_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(names), "cases; cv2", ref.version())


if __name__ == "__main__":
    main()
