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


def _pinned(a):
    b = bb.PinnedBuffer(a.shape, a.dtype)
    b.array[...] = a
    return b


@pytest.mark.parametrize("chunks", [0, 3, 8])
def test_host_batches_in_pinned_memory_take_copy_chunks_and_equal_the_popc_path(eng, chunks):
    """bfm_pipeline.cuh: a large unmasked batch from pinned host arrays is uploaded by the copy engine in chunks of whole
    problems, every chunk is matched by the tensor launches, results come back through a device block and a third
    stream.  Ragged shapes (chunks are not whole rounds), an empty problem inside, match lists and knn tables, into
    pinned result buffers and into plain arrays - against the one-launch POPC path (tensor=1) and the oracle."""
    rng = np.random.default_rng(31)
    P = 56
    nq = [int(x) for x in rng.integers(1200, 2100, P)]
    nt = [int(x) for x in rng.integers(1500, 2300, P)]
    nq[17] = 0
    nt[40] = 1
    qo, to = np.cumsum([0] + nq), np.cumsum([0] + nt)
    q, t = synth.uniform(int(qo[-1]), 51), synth.uniform(int(to[-1]), 52)
    for p in range(P):
        if nq[p] and nt[p]:
            src = rng.integers(0, nt[p], nq[p])
            q[qo[p]:qo[p + 1]] = t[to[p] + src] ^ np.packbits(rng.random((nq[p], 256)) < 0.04, axis=1)
    tab = bb.make_problems(nq, nt)
    pq, pt = _pinned(q), _pinned(t)
    n_out = int(qo[-1])
    res = {}
    try:
        for tensor in (1, 0):
            eng.set_tuning(tensor=tensor, tensor_chunks=chunks)
            idx, dist, r = eng.match_batched(pq.array, pt.array, tab, k=2, ratio=0.8, want_knn=True)       # plain result arrays
            li = eng.launch_info()
            if tensor == 0:
                assert li["popc_mode"] == 0 and li["copy_chunks"] >= 2, li
                assert chunks == 0 or li["copy_chunks"] == chunks, li
            else:
                assert li["popc_mode"] != 0, li
            res[tensor] = (idx.copy(), dist.copy(), r.counts.copy(), [np.concatenate(x) for x in zip(*[r[p] for p in range(P)])])
            hb = bb.HostBatchBuffers(n_out, P, k=2)                                                         # pinned result buffers
            r2 = eng.match_batched(pq.array, pt.array, tab, k=2, ratio=0.8, out=hb)
            assert np.array_equal(r2.counts, r.counts)
            for p in (0, 16, 17, 18, 40, P - 1):
                _eq(r2[p], r[p], ("pinned out", p))
    finally:
        eng.set_tuning(tensor=0, tensor_chunks=0)
    a, b = res[1], res[0]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), "knn tables"
    assert np.array_equal(a[2], b[2]), "counts"
    _eq(a[3], b[3], "match lists")
    for p in (0, 17, 33, 40, P - 1):
        want = orc.match(q[qo[p]:qo[p + 1]], t[to[p]:to[p + 1]], k=2, ratio=0.8)
        n0 = int(np.sum(b[2][:p]))
        n = int(b[2][p])
        _eq([x[n0:n0 + n] for x in b[3]], want, ("oracle", p))


def test_one_frame_against_many_keyframes_from_pinned_memory(eng):
    """Shared query rows (every problem reads the same frame): the chunked upload copies them once, with the first chunk."""
    P, n = 70, 2000
    q, _, _ = synth.correlated(n, 10, 7)
    t = synth.uniform(P * n, 8)
    rng = np.random.default_rng(9)
    for p in range(P):
        rows = rng.integers(0, n, 600)
        t[p * n + rows] = q[rng.integers(0, n, 600)] ^ np.packbits(rng.random((600, 256)) < 0.03, axis=1)
    tab = bb.make_problems([n] * P, [n] * P, shared_query=True)
    pq, pt = _pinned(q), _pinned(t)
    try:
        eng.set_tuning(tensor=0)
        r = eng.match_batched(pq.array, pt.array, tab, k=2, ratio=0.8)
        li = eng.launch_info()
        assert li["popc_mode"] == 0 and li["copy_chunks"] >= 2, li
        eng.set_tuning(tensor=1)
        r1 = eng.match_batched(pq.array, pt.array, tab, k=2, ratio=0.8)
    finally:
        eng.set_tuning(tensor=0)
    assert np.array_equal(r.counts, r1.counts) and int(r.counts.sum()) > 1000
    for p in range(P):
        _eq(r[p], r1[p], p)
    for p in (0, 37, P - 1):
        _eq(r[p], orc.match(q, t[p * n:(p + 1) * n], k=2, ratio=0.8), ("oracle", p))
