"""The tensor form of the matching kernel (boslam_b200/csrc/bfm_tensor.cuh: distances as s8 dot products on tcgen05) against
the oracle and against the POPC kernel: ragged shapes around the tile sizes (256 query rows, 128 train rows, 8-row
alignment of the expanded planes), ties, empty problems, plan reuse with other rows, every finalize option it shares
with the POPC form.  `tensor=2` forces the form wherever the call is eligible (resident or plainly copied inputs,
k <= 2, no mask, no cross-check); `tensor=1` turns it off."""
import numpy as np
import pytest

import boslam_b200 as bb
from boslam_b200 import synth
from oracle import hamming_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = bb.Engine(0)
    yield e
    e.set_tuning(tensor=0)
    e.close()


def _eq(a, b, what=""):
    for x, y in zip(a, b):
        assert np.array_equal(np.asarray(x), np.asarray(y)), what


SHAPES = [(1, 1), (1, 9), (7, 300), (255, 127), (256, 128), (257, 129), (300, 1000), (513, 1025), (1000, 1000), (2000, 2000), (130, 5000)]


@pytest.mark.parametrize("nq,nt", SHAPES)
def test_single_problems_match_the_oracle(eng, nq, nt):
    q, t, _ = synth.correlated(nq, nt, 11 + nq + nt)
    eng.set_tuning(tensor=2)
    try:
        for kw, okw in ((dict(k=2, ratio=0.8), dict(k=2, ratio=0.8)), (dict(k=1), dict(k=1)), (dict(k=1, max_distance=60), dict(k=1, max_distance=60)),
                        (dict(k=2, ratio=0.95, max_distance=90), dict(k=2, ratio=0.95, max_distance=90))):
            _eq(eng.match(q, t, **kw), orc.match(q, t, **okw), (nq, nt, kw))
            li = eng.launch_info()
            assert li["popc_mode"] == 0 and li["scan_block"] == 608, li     # it WAS the tensor form
        for k in (1, 2):
            if nt >= k:
                _eq(eng.knn(q, t, k=k), orc.knn(q, t, k=k), (nq, nt, "knn", k))
    finally:
        eng.set_tuning(tensor=0)


def test_ties_go_to_the_lowest_train_index(eng):
    """Many identical train rows: cv2 keeps the first of equal distances; the keys carry the train index below the
    distance, so both the nearest and the second nearest must be the lowest indices - across tile and item boundaries."""
    rng = np.random.default_rng(3)
    base = synth.uniform(40, 5)
    t = base[rng.integers(0, 40, 3000)]                     # 3000 rows, only 40 distinct descriptors
    q = base[rng.integers(0, 40, 700)] ^ np.packbits(rng.random((700, 256)) < 0.01, axis=1)
    eng.set_tuning(tensor=2)
    try:
        _eq(eng.knn(q, t, k=2), orc.knn(q, t, k=2), "ties")
        _eq(eng.match(q, t, k=2, ratio=0.99), orc.match(q, t, k=2, ratio=0.99), "ties, ratio")
        # all-equal distances: every train row equals every query row
        one = np.repeat(base[:1], 600, axis=0)
        _eq(eng.knn(one[:300], one, k=2), orc.knn(one[:300], one, k=2), "all equal")
        # the extreme distances 0 and 256
        inv = np.bitwise_not(one)
        _eq(eng.knn(one[:50], inv, k=2), orc.knn(one[:50], inv, k=2), "distance 256")
    finally:
        eng.set_tuning(tensor=0)


def test_ragged_batch_equals_the_popc_form_and_the_oracle(eng):
    import torch
    nq = [300, 0, 257, 2000, 5, 640, 1, 900, 0, 1300]
    nt = [500, 40, 0, 2000, 2500, 129, 1, 7, 0, 1111]
    q = synth.uniform(sum(nq), 21)
    t = synth.uniform(sum(nt), 22)
    # plant near matches so ratio / distance gates keep a sensible share
    rng = np.random.default_rng(4)
    qo = np.cumsum([0] + nq)
    to = np.cumsum([0] + nt)
    for p in range(len(nq)):
        if nq[p] and nt[p]:
            src = rng.integers(0, nt[p], nq[p])
            q[qo[p]:qo[p + 1]] = t[to[p] + src] ^ np.packbits(rng.random((nq[p], 256)) < 0.04, axis=1)
    tab = bb.make_problems(nq, nt)
    dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    outs = {}
    try:
        for tensor in (1, 2):
            eng.set_tuning(tensor=tensor)
            for name, kw in (("ratio", dict(k=2, ratio=0.8)), ("gate", dict(k=1, max_distance=50))):
                o = eng.match_batched_device(dq, dt, tab, want_knn=True, **kw)
                torch.cuda.synchronize()
                outs[(tensor, name)] = {k: v.cpu().numpy() for k, v in o.items() if hasattr(v, "cpu")}
                assert (eng.launch_info()["popc_mode"] == 0) == (tensor == 2)
    finally:
        eng.set_tuning(tensor=0)
    for name in ("ratio", "gate"):
        a, b = outs[(1, name)], outs[(2, name)]
        assert np.array_equal(a["count"], b["count"]), name
        for key in a:
            if key == "m":
                for p in range(len(nq)):
                    n = int(a["count"][p])
                    assert np.array_equal(a["m"][:, qo[p]:qo[p] + n], b["m"][:, qo[p]:qo[p] + n]), (name, p)
            elif key != "count":
                assert np.array_equal(a[key], b[key]), (name, key)
    o = outs[(2, "ratio")]
    for p in range(len(nq)):
        n = int(o["count"][p])
        want = orc.match(q[qo[p]:qo[p + 1]], t[to[p]:to[p + 1]], k=2, ratio=0.8)
        _eq((o["m"][0, qo[p]:qo[p] + n], o["m"][1, qo[p]:qo[p] + n], o["m"][2, qo[p]:qo[p] + n]), want, p)


def test_same_shapes_other_rows_reuse_the_tensor_plan(eng):
    sizes = [300, 300, 0, 450, 300, 450, 300, 300, 450, 300]
    kfs = {i: synth.correlated(max(n, 1), max(n, 1), 900 + i)[1][:n] for i, n in enumerate(sizes)}
    bank = bb.KeyframeBank(capacity_rows=8192, engine=eng)
    for i, d in kfs.items():
        bank.add(i, d)
    step1 = [(0, 1), (3, 5), (2, 0), (4, 6), (8, 3), (7, 9)]
    step2 = [(9, 7), (5, 8), (2, 4), (6, 1), (3, 5), (0, 4)]
    eng.set_tuning(tensor=2)
    try:
        for step in (step1, step2, step1, step2[::-1], step1):
            res = bank.match_pairs(step, k=2, ratio=0.9)
            assert eng.launch_info()["popc_mode"] == 0
            for p, (a, b) in enumerate(step):
                _eq(res[p], orc.match(kfs[a], kfs[b], k=2, ratio=0.9), (a, b))
    finally:
        eng.set_tuning(tensor=0)


def test_large_batches_take_the_tensor_form_by_default(eng):
    """32 keyframe pairs of 2000 descriptors: the default choice is the tensor form, and its match lists are the POPC
    form's, bit for bit (sampled against the oracle)."""
    import torch
    P, N = 32, 2000
    q, t = synth.keyframe_pair_batch(P, N, seed=77)
    dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    tab = bb.make_problems([N] * P, [N] * P)
    res = {}
    try:
        for tensor in (0, 1):
            eng.set_tuning(tensor=tensor)
            o = eng.match_batched_device(dq, dt, tab, k=2, ratio=0.8, want_knn=True)
            torch.cuda.synchronize()
            res[tensor] = {k: v.cpu().numpy() for k, v in o.items() if hasattr(v, "cpu")}
            assert (eng.launch_info()["popc_mode"] == 0) == (tensor == 0)
    finally:
        eng.set_tuning(tensor=0)
    assert np.array_equal(res[0]["count"], res[1]["count"])
    for key in res[0]:
        if key == "m":
            for p in range(P):
                n = int(res[0]["count"][p])
                assert np.array_equal(res[0]["m"][:, p * N:p * N + n], res[1]["m"][:, p * N:p * N + n]), p
        elif key != "count":
            assert np.array_equal(res[0][key], res[1][key]), key
    for p in (0, 13, P - 1):
        n = int(res[0]["count"][p])
        want = orc.match(q[p * N:(p + 1) * N], t[p * N:(p + 1) * N], k=2, ratio=0.8)
        _eq((res[0]["m"][0, p * N:p * N + n], res[0]["m"][1, p * N:p * N + n], res[0]["m"][2, p * N:p * N + n]), want, p)
