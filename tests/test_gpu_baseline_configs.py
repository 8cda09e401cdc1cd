"""BASELINE.json configs[0..4] at FULL size, through the public API, against live cv2 (the reference's own
implementation of the path, oracle/cv2_reference.py) - every (queryIdx, trainIdx, distance) must be identical.

    configs[0]  frame <-> frame, 1000 x 1000, crossCheck + `< 30`          slam/tracking.py:56-57
    configs[1]  tracking, 2000 x 20000: crossCheck + `<= 30` (slam/tracking.py:121, reference-faithful) and
                projection window 15 px + ratio 0.8 (north star; cv2 knnMatch(mask=dense) is the check)
    configs[2]  local mapping, 20 pairs x (2000 x 2000): k = 2 + ratio 0.8 and crossCheck, full tables
    configs[3]  loop closing, 256 pairs x (2000 x 2000): k = 2 + ratio 0.8, full knn tables and match lists
    configs[4]  brute-force sweep points 16k x 16k and 64k x 64k: k = 1, k = 2, crossCheck

Where cv2 has no answer (crossCheck + mask asserts in cv2) the numpy/C oracle is the check.
"""
import numpy as np
import pytest

import boslam_b200 as bb
from boslam_b200 import synth
from oracle import c_oracle, cv2_reference as ref, hamming_oracle as orc

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.HAVE_CV2, reason="cv2 not importable")]

RATIO = 0.8


@pytest.fixture(scope="module")
def eng():
    e = bb.Engine(0)
    yield e
    e.close()


def _eq3(got, want, what=""):
    for name, a, b in zip(("queryIdx", "trainIdx", "distance"), got, want):
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape, f"{what}: {name} has {a.shape[0]} entries, reference has {b.shape[0]}"
        assert np.array_equal(a.astype(np.int64), b.astype(np.int64)), f"{what}: {name} differs"


def _cv2_knn_tables(q, t, k, mask=None):
    """cv2 knnMatch -> dense int32 tables, vectorised over the DMatch rows (rows are full here: T >= k)."""
    m = ref.matcher(False)
    rows = m.knnMatch(q, t, k) if mask is None else m.knnMatch(q, t, k, mask)
    idx = np.full((len(q), k), -1, np.int32)
    dist = np.full((len(q), k), -1, np.int32)
    for i, r in enumerate(rows):
        for c, dm in enumerate(r):
            idx[i, c] = dm.trainIdx
            dist[i, c] = int(dm.distance)
    return idx, dist


def _ratio_from_tables(idx, dist, ratio):
    """The caller-side Lowe test exactly as Python evaluates it (fp64), from dense tables."""
    keep = (idx[:, 1] >= 0) & (dist[:, 0].astype(np.float64) < ratio * dist[:, 1].astype(np.float64))
    qi = np.nonzero(keep)[0].astype(np.int32)
    return qi, idx[qi, 0], dist[qi, 0]


def test_config0_frame_to_frame_1000(eng):
    for seed in (0, 1):
        q, t, _ = synth.correlated(1000, 1000, seed)
        rq, rt, rd = ref.match(q, t, cross_check=True)
        keep = rd < 30                                           # slam/tracking.py:57
        _eq3(eng.match(q, t, cross_check=True, max_distance=30, strict=True), (rq[keep], rt[keep], rd[keep]), f"seed {seed}")
        _eq3(eng.match(q, t, cross_check=True), (rq, rt, rd), f"seed {seed} ungated")


def test_config1_tracking_2000x20000_crosscheck_gate(eng):
    """The reference-faithful variant: duplicate train rows (one per (keyframe, map point) edge), crossCheck, `<= 30`."""
    q, t, _ = synth.correlated(2000, 5000, 21)
    t = np.ascontiguousarray(np.repeat(t, 4, axis=0))          # 20000 rows, every descriptor four times
    assert t.shape[0] == 20000
    rq, rt, rd = ref.match(q, t, cross_check=True)
    keep = rd <= 30                                              # slam/tracking.py:121
    _eq3(eng.match(q, t, cross_check=True, max_distance=30), (rq[keep], rt[keep], rd[keep]), "duplicates")
    q, t, _, _, _ = synth.window_scene(2000, 20000, 22)
    rq, rt, rd = ref.match(q, t, cross_check=True)
    keep = rd <= 30
    _eq3(eng.match(q, t, cross_check=True, max_distance=30), (rq[keep], rt[keep], rd[keep]), "scene")


@pytest.mark.parametrize("binned", [True, False])
def test_config1_tracking_2000x20000_window_ratio(eng, binned):
    """North-star variant: 15 px projection window + ratio 0.8; cv2.knnMatch with the dense mask built from the
    same fp32 predicate is the reference.  Both the binned search and the brute-force window kernel."""
    q, t, qxy, txy, _ = synth.window_scene(2000, 20000, 23)
    r = np.float32(15.0)
    mask = ((np.abs(qxy[:, None, 0] - txy[None, :, 0]) < r) & (np.abs(qxy[:, None, 1] - txy[None, :, 1]) < r)).astype(np.uint8)
    ci, cd = _cv2_knn_tables(q, t, 2, mask)
    eng.set_tuning(window_bins=0 if binned else 1)
    try:
        idx, dist = eng.knn(q, t, 2, window=(qxy, txy, 15.0))
        got = eng.match(q, t, k=2, ratio=RATIO, window=(qxy, txy, 15.0))
        assert eng.launch_info()["kernels_launched"] >= 1
    finally:
        eng.set_tuning(window_bins=0)
    assert np.array_equal(idx, ci) and np.array_equal(dist, cd), "windowed knn table differs from cv2 knnMatch(mask)"
    _eq3(got, _ratio_from_tables(ci, cd, RATIO), "window + ratio")
    # the same search with device-resident inputs (the form MapStore.track uses)
    import torch
    dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    dqxy, dtxy = torch.from_numpy(qxy).cuda(), torch.from_numpy(txy).cuda()
    di, dd = eng.knn(dq, dt, 2, window=(dqxy, dtxy, 15.0))
    assert np.array_equal(di.cpu().numpy(), ci) and np.array_equal(dd.cpu().numpy(), cd)
    # crossCheck + window: cv2 asserts (R6); oracle = masked distances -> mutual argmin
    _eq3(eng.match(q, t, cross_check=True, max_distance=30, window=(qxy, txy, 15.0)),
         orc.match(q, t, cross_check_=True, max_distance=30, mask=mask), "window + crossCheck")


def test_config2_local_mapping_20_pairs_full_tables(eng):
    P, N = 20, 2000
    qb, tb = synth.keyframe_pair_batch(P, N, seed=31)
    tab = bb.make_problems([N] * P, [N] * P)
    idx, dist, res = eng.match_batched(qb, tb, tab, k=2, ratio=RATIO, want_knn=True)
    resx = eng.match_batched(qb, tb, tab, cross_check=True, max_distance=30)
    import torch
    dres = eng.match_batched(torch.from_numpy(qb).cuda(), torch.from_numpy(tb).cuda(), tab, cross_check=True, max_distance=30)
    for p in range(P):
        q, t = qb[p * N:(p + 1) * N], tb[p * N:(p + 1) * N]
        ci, cd = _cv2_knn_tables(q, t, 2)
        assert np.array_equal(idx[p * N:(p + 1) * N], ci) and np.array_equal(dist[p * N:(p + 1) * N], cd), f"pair {p}: knn table"
        _eq3(res[p], _ratio_from_tables(ci, cd, RATIO), f"pair {p}: ratio list")
        rq, rt, rd = ref.match(q, t, cross_check=True)           # slam/local_mapping.py:21 builds crossCheck=True
        keep = rd <= 30
        _eq3(resx[p], (rq[keep], rt[keep], rd[keep]), f"pair {p}: crossCheck list")
        _eq3(dres[p], (rq[keep], rt[keep], rd[keep]), f"pair {p}: crossCheck list (device inputs)")


def test_config3_loop_closing_256_pairs_full_tables(eng):
    """The bench's headline batch, every entry: knn tables (512k x 2) and ratio-filtered match lists of all 256
    pairs against cv2, through the pinned host path (SM-fed upload), plain numpy arrays and device tensors."""
    P, N = 256, 2000
    qb, tb = synth.keyframe_pair_batch(P, N, seed=1000)          # the bench's rank-0 input set 0
    tab = bb.make_problems([N] * P, [N] * P)
    idx, dist, res = eng.match_batched(qb, tb, tab, k=2, ratio=RATIO, want_knn=True)
    m = ref.matcher(False)
    total = 0
    for p in range(P):
        q, t = qb[p * N:(p + 1) * N], tb[p * N:(p + 1) * N]
        ci, cd = _cv2_knn_tables(q, t, 2)
        assert np.array_equal(idx[p * N:(p + 1) * N], ci) and np.array_equal(dist[p * N:(p + 1) * N], cd), f"pair {p}: knn table"
        want = _ratio_from_tables(ci, cd, RATIO)
        _eq3(res[p], want, f"pair {p}: ratio list")
        total += len(want[0])
    assert total == int(res.counts.sum()) and total > 0
    # the two other ways into the same kernel must return the same bits
    pq, pt = bb.PinnedBuffer(qb.shape), bb.PinnedBuffer(tb.shape)
    pq.array[...] = qb
    pt.array[...] = tb
    out = bb.HostBatchBuffers(P * N, P, k=2, want_knn=True)
    i2, d2, r2 = eng.match_batched(pq.array, pt.array, tab, k=2, ratio=RATIO, want_knn=True, out=out)
    assert np.array_equal(i2, idx) and np.array_equal(d2, dist) and np.array_equal(r2.counts, res.counts)
    import torch
    i3, d3, r3 = eng.match_batched(torch.from_numpy(qb).cuda(), torch.from_numpy(tb).cuda(), tab, k=2, ratio=RATIO, want_knn=True)
    assert np.array_equal(i3.cpu().numpy(), idx) and np.array_equal(d3.cpu().numpy(), dist)
    for p in (0, 100, 255):
        _eq3(r2[p], res[p], f"pinned path pair {p}")
        _eq3(r3[p], res[p], f"device path pair {p}")


@pytest.mark.parametrize("n", [16384, 65536])
def test_config4_sweep_points_against_cv2(eng, n):
    """configs[4] at its two ncu profile points: uniform descriptors, k = 1, k = 2 and crossCheck against cv2
    (64k < 2^18, so cv2 accepts it; ~10 s of host time at 64k)."""
    q, t = synth.uniform(n, 41), synth.uniform(n, 42)
    ci, cd = _cv2_knn_tables(q, t, 2)
    idx, dist = eng.knn(q, t, 2)
    assert np.array_equal(idx, ci) and np.array_equal(dist, cd), "k = 2"
    i1, d1 = eng.knn(q, t, 1)
    assert np.array_equal(i1[:, 0], ci[:, 0]) and np.array_equal(d1[:, 0], cd[:, 0]), "k = 1"
    _eq3(eng.match(q, t, cross_check=True), ref.match(q, t, cross_check=True), "crossCheck")
    # device-resident form (the sweep's timed form)
    import torch
    di, dd = eng.knn(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), 2)
    assert np.array_equal(di.cpu().numpy(), ci) and np.array_equal(dd.cpu().numpy(), cd)
