"""Parity of the CUDA path (through the C ABI) against the CPU oracle, the committed cv2 golden
vectors and live cv2.  Bit-exact: indices and integer distances must be identical."""
import numpy as np
import pytest

import boslam_b200 as bb
from boslam_b200 import synth
from oracle import c_oracle, cv2_reference as ref, hamming_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = bb.Engine(0)
    yield e
    e.close()


def _eq(a, b, what=""):
    for x, y in zip(a, b):
        assert np.array_equal(np.asarray(x), np.asarray(y)), what


def _names(golden):
    return [str(n) for n in golden["names"]]


def test_golden_vectors(eng, golden):
    for name in _names(golden):
        q, t = golden[f"{name}/q"], golden[f"{name}/t"]
        for k in (1, 2):
            idx, dist = eng.knn(q, t, k)
            assert np.array_equal(idx, golden[f"{name}/knn{k}_idx"]), (name, k)
            assert np.array_equal(dist, golden[f"{name}/knn{k}_dist"]), (name, k)
        qi, ti, d = eng.match(q, t, cross_check=True)
        _eq((qi, ti, d), (golden[f"{name}/cc_q"], golden[f"{name}/cc_t"], golden[f"{name}/cc_d"]), name)
        idx, dist = eng.knn(q, t, 2, mask=golden[f"{name}/mask"])
        assert np.array_equal(idx, golden[f"{name}/mknn2_idx"]), name
        assert np.array_equal(dist, golden[f"{name}/mknn2_dist"]), name
        rq, rt, rd = eng.match(q, t, k=2, ratio=0.8)
        _eq((rq, rt, rd), (golden[f"{name}/ratio_q"], golden[f"{name}/ratio_t"], golden[f"{name}/ratio_d"]), name)


@pytest.mark.parametrize("pm", [8, 6, 5, 4, 50, 40])
@pytest.mark.parametrize("r", [1, 2, 4])
def test_kernel_variants_bit_exact(eng, pm, r):
    """Every POPC / register-tile variant must return identical results."""
    eng.set_tuning(popc_mode=pm, queries_per_thread=r)
    try:
        q, t, _ = synth.correlated(700, 1300, seed=pm * 10 + r)
        oi, od = c_oracle.knn(q, t, 2)
        idx, dist = eng.knn(q, t, 2)
        assert np.array_equal(idx, oi) and np.array_equal(dist, od)
        _eq(eng.match(q, t, cross_check=True), c_oracle.cross_check(q, t))
        qt, tt = synth.tie_stress(500, pm), synth.tie_stress(900, r + 100)
        oi, od = c_oracle.knn(qt, tt, 2)
        idx, dist = eng.knn(qt, tt, 2)
        assert np.array_equal(idx, oi) and np.array_equal(dist, od)
        _eq(eng.match(qt, tt, cross_check=True), c_oracle.cross_check(qt, tt))
    finally:
        eng.set_tuning(popc_mode=0, queries_per_thread=0)


@pytest.mark.parametrize("seed", range(6))
def test_random_shapes_vs_oracle(eng, seed):
    rng = np.random.default_rng(seed)
    nq, nt = int(rng.integers(1, 1500)), int(rng.integers(1, 3000))
    gens = [
        lambda: synth.correlated(nq, nt, seed)[:2],
        lambda: (synth.tie_stress(nq, seed), synth.tie_stress(nt, seed + 50)),
        lambda: (synth.uniform(nq, seed), synth.duplicate_rows(max(1, nt // 3), seed)),
    ]
    for g in gens:
        q, t = g()
        for k in (1, 2):
            oi, od = c_oracle.knn(q, t, k)
            idx, dist = eng.knn(q, t, k)
            assert np.array_equal(idx, oi) and np.array_equal(dist, od), (nq, nt, k)
        _eq(eng.match(q, t, cross_check=True), c_oracle.cross_check(q, t), (nq, nt))
        _eq(eng.match(q, t, k=2, ratio=0.8), orc.match(q, t, k=2, ratio=0.8))
        _eq(eng.match(q, t, cross_check=True, max_distance=30, strict=True),
            orc.match(q, t, cross_check_=True, max_distance=30, strict=True))
        _eq(eng.match(q, t, cross_check=True, max_distance=30),
            orc.match(q, t, cross_check_=True, max_distance=30))


@pytest.mark.skipif(not ref.HAVE_CV2, reason="cv2 not importable")
def test_live_cv2_reference_call_shapes(eng):
    """The reference's two live call shapes (slam/tracking.py:56-57 and :121) against live cv2."""
    q, t, _ = synth.correlated(1000, 1000, 3)                       # BASELINE config 1
    rq, rt, rd = ref.match(q, t, cross_check=True)
    keep = rd < 30                                                   # slam/tracking.py:57 (strict)
    _eq(eng.match(q, t, cross_check=True, max_distance=30, strict=True), (rq[keep], rt[keep], rd[keep]))
    feats = synth.duplicate_rows(7000, 4)                            # local map with duplicate rows (finding 4)
    q2 = synth.correlated(2000, len(feats), 5)[0]
    q2[:1400] = feats[np.random.default_rng(6).integers(0, len(feats), 1400)]
    rq, rt, rd = ref.match(q2, feats, cross_check=True)
    keep = rd <= 30                                                  # slam/tracking.py:121 (non-strict)
    _eq(eng.match(q2, feats, cross_check=True, max_distance=30), (rq[keep], rt[keep], rd[keep]))
    ri, rdd = ref.knn(q2, feats, 2)
    idx, dist = eng.knn(q2, feats, 2)
    assert np.array_equal(idx, ri) and np.array_equal(dist, rdd)


def test_drop_in_bfmatcher_object(eng):
    """BFMatcher_create(...).match returns DMatch objects the call sites can consume unchanged."""
    q, t, _ = synth.correlated(400, 600, 8)
    m = bb.BFMatcher_create(bb.NORM_HAMMING, crossCheck=True)
    ms = m.match(q, t)
    ms = [_ for _ in ms if _.distance < 30]                          # slam/tracking.py:57
    inds_f, inds_kf = zip(*((_.queryIdx, _.trainIdx) for _ in ms))   # slam/tracking.py:60
    oq, ot, od = orc.match(q, t, cross_check_=True, max_distance=30, strict=True)
    assert list(inds_f) == oq.tolist() and list(inds_kf) == ot.tolist()
    assert [x.distance for x in ms] == od.astype(float).tolist() and all(x.imgIdx == 0 for x in ms)
    assert isinstance(ms[0].distance, float)
    m2 = bb.BFMatcher_create(bb.NORM_HAMMING)
    rows = m2.knnMatch(q, t, k=2)
    oi, odd = orc.knn(q, t, 2)
    assert len(rows) == len(q)
    assert [[d.trainIdx for d in r] for r in rows] == oi.tolist()
    assert [[int(d.distance) for d in r] for r in rows] == odd.tolist()
    if ref.HAVE_CV2:
        import cv2
        cm = cv2.BFMatcher_create(cv2.NORM_HAMMING, crossCheck=True).match(q, t)
        mine = m.match(q, t)
        assert [(a.queryIdx, a.trainIdx, a.imgIdx, a.distance) for a in cm] == \
               [(a.queryIdx, a.trainIdx, a.imgIdx, a.distance) for a in mine]


def test_masks_and_window(eng):
    q, t, qxy, txy, _ = synth.window_scene(900, 2500, 9)
    dense = orc.window_mask(qxy, txy, 15.0)
    for k in (1, 2):
        oi, od = c_oracle.knn(q, t, k, dense)
        idx, dist = eng.knn(q, t, k, window=(qxy, txy, 15.0))
        assert np.array_equal(idx, oi) and np.array_equal(dist, od), k
        idx, dist = eng.knn(q, t, k, mask=dense)
        assert np.array_equal(idx, oi) and np.array_equal(dist, od), k
    _eq(eng.match(q, t, k=2, ratio=0.8, window=(qxy, txy, 15.0)), orc.match(q, t, k=2, ratio=0.8, mask=dense))
    # cross-check + mask: cv2 refuses it (R6); the numpy restatement is the oracle
    _eq(eng.match(q, t, cross_check=True, window=(qxy, txy, 15.0)), orc.match(q, t, cross_check_=True, mask=dense))
    _eq(eng.match(q, t, cross_check=True, mask=dense), c_oracle.cross_check(q, t, dense))
    m255 = dense * 255                                               # R4: any non-zero allows
    _eq(eng.knn(q, t, 2, mask=m255), eng.knn(q, t, 2, mask=dense))
    dense[5, :] = 0                                                  # fully masked query -> no match
    idx, _ = eng.knn(q, t, 2, mask=dense)
    assert (idx[5] == -1).all()


def test_edge_cases(eng):
    e = np.zeros((0, 32), np.uint8)
    t = synth.uniform(5, 1)
    assert all(len(x) == 0 for x in eng.match(e, t))
    assert all(len(x) == 0 for x in eng.match(t, e, cross_check=True))
    idx, dist = eng.knn(t, e, 2)
    assert idx.shape == (5, 2) and (idx == -1).all() and (dist == -1).all()
    idx, dist = eng.knn(t, t[:1], 2)                                 # fewer than k candidates (R3)
    assert (idx[:, 0] == 0).all() and (idx[:, 1] == -1).all() and (dist[:, 1] == -1).all()
    assert len(eng.match(t, t[:1], k=2, ratio=0.8)[0]) == 0          # rows with < 2 neighbours are dropped
    z = np.zeros((3, 32), np.uint8)
    f = np.full((2, 32), 255, np.uint8)
    idx, dist = eng.knn(z, f, 2)                                     # distance 256, the maximum
    assert dist.tolist() == [[256, 256]] * 3 and idx.tolist() == [[0, 1]] * 3
    with pytest.raises(TypeError):
        eng.match(t.astype(np.float32), t)
    with pytest.raises(ValueError):
        eng.match(np.zeros((4, 16), np.uint8), t)
    with pytest.raises(ValueError):
        eng.match(t, t, k=2, cross_check=True)
    nc = np.asfortranarray(synth.uniform(40, 2))                     # non-contiguous input is accepted (R9)
    _eq(eng.knn(nc, t, 1), orc.knn(np.ascontiguousarray(nc), t, 1))
    assert len(eng.match(t, t, max_distance=0, strict=True)[0]) == 0  # nothing is < 0


def test_batched_pairs(eng):
    rng = np.random.default_rng(21)
    qs, ts = [], []
    for p in range(13):                                              # ragged batch, one empty problem
        nq, nt = int(rng.integers(1, 700)), int(rng.integers(1, 900))
        if p == 4:
            nt = 0
        q, t, _ = synth.correlated(nq, max(nt, 1), 100 + p)
        qs.append(q)
        ts.append(t[:nt])
    for kw, okw in (({"cross_check": True}, {"cross_check_": True}),
                    ({"k": 2, "ratio": 0.8}, {"k": 2, "ratio": 0.8}),
                    ({"k": 1, "max_distance": 40}, {"k": 1, "max_distance": 40})):
        res = eng.match_pairs(qs, ts, **kw)
        assert len(res) == 13
        for p in range(13):
            _eq(res[p], orc.match(qs[p], ts[p], **okw), (p, kw))
    # shared query keyframe (loop closing shape): stored once, matched against every candidate
    q0 = qs[0]
    res = eng.match_pairs([q0] * 5, ts[:5], k=2, ratio=0.8)
    for p in range(5):
        _eq(res[p], orc.match(q0, ts[p], k=2, ratio=0.8), p)
    res = eng.match_pairs([q0] * 5, ts[:5], cross_check=True)
    for p in range(5):
        _eq(res[p], orc.match(q0, ts[p], cross_check_=True), p)


def test_batched_knn_table_config3_shape(eng):
    """BASELINE config 3 shape, reduced: 6 pairs x (2000 x 2000), dense k=2 table + ratio list."""
    qs, ts = synth.keyframe_pairs(6, 2000, seed=3)
    qp, tp = np.concatenate(qs), np.concatenate(ts)
    tab = bb.make_problems([2000] * 6, [2000] * 6)
    idx, dist, res = eng.match_batched(qp, tp, tab, k=2, ratio=0.8, want_knn=True)
    for p in range(6):
        oi, od = c_oracle.knn(qs[p], ts[p], 2)
        assert np.array_equal(idx[p * 2000:(p + 1) * 2000], oi)
        assert np.array_equal(dist[p * 2000:(p + 1) * 2000], od)
        _eq(res[p], orc.match(qs[p], ts[p], k=2, ratio=0.8), p)


def test_device_tensor_path(eng):
    torch = pytest.importorskip("torch")
    q, t, _ = synth.correlated(800, 1700, 31)
    qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    idx, dist = eng.knn(qd, td, 2)
    oi, od = c_oracle.knn(q, t, 2)
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(dist.cpu().numpy(), od)
    got = eng.match(qd, td, cross_check=True, max_distance=30)
    _eq([g.cpu().numpy() for g in got], orc.match(q, t, cross_check_=True, max_distance=30))
    tab = bb.make_problems([800], [800])                             # a gate nothing can pass (d < 0), device form
    assert int(eng.match_batched_device(qd, qd, tab, max_distance=0, strict=True)["count"][0]) == 0
    assert len(eng.match(qd, qd, max_distance=0, strict=True)[0]) == 0
    assert int(eng.match_batched_device(qd, qd, tab, max_distance=0)["count"][0]) == 800   # d <= 0: the self matches
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):                                       # honours torch's current stream
        idx2, dist2 = eng.knn(qd, td, 2)
    s.synchronize()
    assert torch.equal(idx2, idx) and torch.equal(dist2, dist)


def test_split_invariance_properties(eng):
    """Size-independent properties at a larger size: any work split gives the same answer, the
    cross-check list is the mutual subset of the 1-NN table, and swapping roles mirrors it."""
    q, t, _ = synth.correlated(3000, 9000, 41)
    base = eng.knn(q, t, 2)
    for rows in (32, 100, 1000, 100000):
        eng.set_tuning(segment_rows=rows)
        _eq(eng.knn(q, t, 2), base, rows)
    eng.set_tuning(segment_rows=0)
    qi, ti, d = eng.match(q, t, cross_check=True)
    assert np.array_equal(base[0][qi, 0], ti) and np.array_equal(base[1][qi, 0], d.astype(np.int32))
    ti2, qi2, d2 = eng.match(t, q, cross_check=True)                 # roles swapped
    o = np.argsort(qi2, kind="stable")
    assert np.array_equal(qi2[o], qi) and np.array_equal(ti2[o], ti) and np.array_equal(d2[o], d)
    rev = eng.knn(t, q, 1)
    assert np.array_equal(rev[0][ti, 0], qi)


def test_threads_own_handles():
    """Two threads, two matchers (tracking + local mapping, slam/main.py:37-47), concurrently."""
    import threading
    q, t, _ = synth.correlated(600, 900, 51)
    want = orc.match(q, t, cross_check_=True)
    errs = []

    def work():
        try:
            m = bb.BFMatcher_create(bb.NORM_HAMMING, crossCheck=True)
            for _ in range(10):
                ms = m.match(q, t)
                assert [x.queryIdx for x in ms] == want[0].tolist()
                assert [x.trainIdx for x in ms] == want[1].tolist()
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work) for _ in range(2)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errs, errs


def test_pipelined_host_path_and_reused_buffers(eng):
    """Large batches take the chunked copy/compute-overlap path; results must not depend on it."""
    qp, tp = synth.keyframe_pair_batch(24, 1500, seed=77)
    tab = bb.make_problems([1500] * 24, [1500] * 24)
    out = bb.HostBatchBuffers(24 * 1500, 24, k=2, want_knn=True)
    ref_res = None
    for chunks in (1, 0, 3, 16):                                     # 1 = pipeline off
        eng.set_tuning(pipeline_chunks=chunks)
        idx, dist, res = eng.match_batched(qp, tp, tab, k=2, ratio=0.8, want_knn=True, out=out)
        snap = (idx.copy(), dist.copy(), res.counts.copy(), [tuple(a.copy() for a in res[p]) for p in range(24)])
        if ref_res is None:
            ref_res = snap
            for p in (0, 7, 23):
                oi, od = c_oracle.knn(qp[p * 1500:(p + 1) * 1500], tp[p * 1500:(p + 1) * 1500], 2)
                assert np.array_equal(idx[p * 1500:(p + 1) * 1500], oi) and np.array_equal(dist[p * 1500:(p + 1) * 1500], od)
                _eq(res[p], orc.match(qp[p * 1500:(p + 1) * 1500], tp[p * 1500:(p + 1) * 1500], k=2, ratio=0.8), p)
        else:
            assert np.array_equal(snap[0], ref_res[0]) and np.array_equal(snap[1], ref_res[1])
            assert np.array_equal(snap[2], ref_res[2])
            for a, b in zip(snap[3], ref_res[3]):
                _eq(a, b)
    eng.set_tuning(pipeline_chunks=0)
    # shared query keyframe + cross-check through the pipelined path (pageable outputs)
    q0 = qp[:1500]
    tab2 = bb.make_problems([1500] * 24, [1500] * 24, shared_query=True)
    eng.set_tuning(pipeline_chunks=4)
    res = eng.match_batched(q0, tp, tab2, cross_check=True)
    eng.set_tuning(pipeline_chunks=0)
    for p in (0, 11, 23):
        _eq(res[p], c_oracle.cross_check(q0, tp[p * 1500:(p + 1) * 1500]), p)


def test_self_cleaning_workspace_and_plan_cache(eng):
    """Back-to-back calls of different modes / shapes reuse one workspace without a memset, and a
    repeated shape reuses the cached plan: results must stay exact."""
    shapes = [(300, 700), (300, 700), (1200, 400), (50, 5000), (300, 700), (1200, 400)]
    for rep, (nq, nt) in enumerate(shapes * 2):
        q, t, _ = synth.correlated(nq, nt, 500 + rep)
        mode = rep % 3
        if mode == 0:
            _eq(eng.match(q, t, cross_check=True), c_oracle.cross_check(q, t), rep)
        elif mode == 1:
            _eq(eng.knn(q, t, 2), c_oracle.knn(q, t, 2), rep)
        else:
            _eq(eng.knn(q, t, 1), c_oracle.knn(q, t, 1), rep)
    q, t, _ = synth.correlated(640, 480, 3)
    want = c_oracle.cross_check(q, t)
    for _ in range(5):                                               # identical call: plan-cache hits
        _eq(eng.match(q, t, cross_check=True), want)


def test_empty_problems_inside_a_batch(eng):
    """Problems with no query or no train rows still get their count / knn rows written (one kernel:
    the CTA that completes a problem finalizes it, so every problem owns at least one work item)."""
    q, t, _ = synth.correlated(300, 400, 61)
    tab = np.array([[0, 100, 0, 0, 0, 0],        # no train rows
                    [100, 0, 0, 50, 100, 0],     # no query rows
                    [100, 200, 50, 350, 100, 0],
                    [0, 100, 0, 0, 300, 0]], np.int32)
    idx, dist, res = eng.match_batched(q, t, tab, k=2, want_knn=True)
    assert res.counts.tolist()[:2] == [0, 0] and res.counts[3] == 0
    assert (idx[:100] == -1).all() and (dist[:100] == -1).all() and (idx[300:400] == -1).all()
    oi, od = c_oracle.knn(q[100:300], t[50:400], 2)
    assert np.array_equal(idx[100:300], oi) and np.array_equal(dist[100:300], od)
    _eq(res[2], orc.match(q[100:300], t[50:400], k=2))


def test_multi_destination_epilogue(eng):
    """bfm_match_batched_multi: the epilogue writes identical results to every destination (the
    fused multi-GPU gather; here all destinations are buffers on the same GPU)."""
    torch = pytest.importorskip("torch")
    qs, ts = synth.keyframe_pairs(5, 700, seed=9)
    qp, tp = np.concatenate(qs), np.concatenate(ts)
    tab = bb.make_problems([700] * 5, [700] * 5)
    qd, td = torch.from_numpy(qp).cuda(), torch.from_numpy(tp).cuda()
    n = 5 * 700

    def bufs():
        return {"m": torch.full((3, n), -7, dtype=torch.int32, device="cuda"),
                "count": torch.full((5,), -7, dtype=torch.int32, device="cuda"),
                "knn_idx": torch.full((n, 2), -7, dtype=torch.int32, device="cuda"),
                "knn_dist": torch.full((n, 2), -7, dtype=torch.int32, device="cuda")}

    out, reps = bufs(), [bufs(), bufs(), bufs()]
    eng.match_batched_device(qd, td, tab, k=2, ratio=0.8, want_knn=True, out=out, replicas=reps)
    torch.cuda.synchronize()
    for p in range(5):
        oi, od = c_oracle.knn(qs[p], ts[p], 2)
        want = orc.match(qs[p], ts[p], k=2, ratio=0.8)
        for o in [out] + reps:
            assert np.array_equal(o["knn_idx"][p * 700:(p + 1) * 700].cpu().numpy(), oi)
            assert np.array_equal(o["knn_dist"][p * 700:(p + 1) * 700].cpu().numpy(), od)
            c = int(o["count"][p])
            assert c == len(want[0])
            m = o["m"][:, p * 700:p * 700 + c].cpu().numpy()
            assert np.array_equal(m[0], want[0]) and np.array_equal(m[1], want[1])


@pytest.mark.parametrize("k", [3, 4, 5, 8, 16])
def test_knn_k_above_two(eng, golden, k):
    """cv2 knnMatch with k > 2 (rules R2, R3): rows ascend by (distance, trainIdx); short rows when
    fewer than k candidates exist.  One kernel pass per two neighbours."""
    q, t, _ = synth.correlated(600, 900, 70 + k)
    oi, od = c_oracle.knn(q, t, k)
    idx, dist = eng.knn(q, t, k)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)
    qt, tt = synth.tie_stress(300, k), synth.duplicate_rows(40, k)      # ties and duplicate rows
    oi, od = c_oracle.knn(qt, tt, k)
    idx, dist = eng.knn(qt, tt, k)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)
    idx, dist = eng.knn(q[:50], t[:k - 1], k)                            # fewer than k candidates
    assert (idx[:, k - 1] == -1).all() and (idx[:, :k - 1] >= 0).all()
    q2, t2, qxy, txy, _ = synth.window_scene(400, 1200, k)               # masked
    dense = orc.window_mask(qxy, txy, 25.0)
    oi, od = c_oracle.knn(q2, t2, k, dense)
    for got in (eng.knn(q2, t2, k, window=(qxy, txy, 25.0)), eng.knn(q2, t2, k, mask=dense)):
        assert np.array_equal(got[0], oi) and np.array_equal(got[1], od)
    if k == 3:
        for name in _names(golden):
            gq, gt = golden[f"{name}/q"], golden[f"{name}/t"]
            idx, dist = eng.knn(gq, gt, 3)
            assert np.array_equal(idx, golden[f"{name}/knn3_idx"]) and np.array_equal(dist, golden[f"{name}/knn3_dist"]), name
    # batched + the drop-in object
    qs, ts = synth.keyframe_pairs(3, 300, seed=k)
    tab = bb.make_problems([300] * 3, [300] * 3)
    bi, bd, _ = eng.match_batched(np.concatenate(qs), np.concatenate(ts), tab, k=k, want_knn=True)
    for p in range(3):
        oi, od = c_oracle.knn(qs[p], ts[p], k)
        assert np.array_equal(bi[p * 300:(p + 1) * 300], oi) and np.array_equal(bd[p * 300:(p + 1) * 300], od)
    rows = bb.BFMatcher_create(bb.NORM_HAMMING).knnMatch(q[:20], t, k=k)
    oi, od = c_oracle.knn(q[:20], t, k)
    assert [[m.trainIdx for m in r] for r in rows] == oi.tolist()


@pytest.mark.skipif(not ref.HAVE_CV2, reason="cv2 not importable")
def test_resident_collection_api_vs_cv2(eng):
    """cv2.DescriptorMatcher collection API (add / train / clear, imgIdx): the collection is uploaded once
    and stays on the GPU; results equal cv2's for match and knnMatch."""
    import cv2
    # every image has >= k rows: with a shorter image cv2's per-image update re-creates its k-column buffers and
    # returns truncated rows (probed on 4.13: sizes (400, 250, 2), k = 3 -> every row has 2 entries); this
    # engine returns the k nearest over the whole collection instead (documented divergence)
    imgs = [synth.correlated(300, n, 90 + i)[1] for i, n in enumerate((400, 5, 250))]
    q = synth.correlated(300, 400, 90)[0]
    mine, theirs = bb.BFMatcher_create(bb.NORM_HAMMING), cv2.BFMatcher_create(cv2.NORM_HAMMING)
    mine.add(imgs)
    theirs.add(imgs)
    mine.train()
    assert not mine.empty() and len(mine.getTrainDescriptors()) == 3
    for _ in range(2):                                                 # second call reuses the resident copy
        a, b = mine.match(q), theirs.match(q)
        assert [(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in a] == \
               [(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in b]
    a, b = mine.knnMatch(q, k=3), theirs.knnMatch(q, k=3)
    assert [[(m.trainIdx, m.imgIdx, m.distance) for m in r] for r in a] == \
           [[(m.trainIdx, m.imgIdx, m.distance) for m in r] for r in b]
    mine.clear()
    assert mine.empty() and mine.match(q) == ()
    # crossCheck over a multi-image collection: cv2 asserts (batch_distance.cpp:303, update != 0); here it is
    # the mutual-nearest test over the concatenated collection
    mc = bb.BFMatcher_create(bb.NORM_HAMMING, crossCheck=True)
    mc.add(imgs)
    cat, bounds = np.concatenate(imgs), np.cumsum([0] + [len(i) for i in imgs])
    oq, ot, od = orc.match(q, cat, cross_check_=True)
    img = np.searchsorted(bounds, ot, side="right") - 1
    assert [(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in mc.match(q)] == \
           list(zip(oq.tolist(), (ot - bounds[img]).tolist(), img.tolist(), od.astype(float).tolist()))


def test_keyframe_bank_pairs_without_descriptor_traffic(eng):
    """KeyframeBank: keyframes uploaded once, pair batches name them by id (local mapping: new keyframe
    vs covisible ones; loop closing: current keyframe vs candidates)."""
    bank = bb.KeyframeBank(capacity_rows=1024, engine=eng)             # small: forces growth + compaction
    kfs = {i: synth.correlated(10, 200 + 37 * i, 300 + i)[1] for i in range(12)}
    for i, d in kfs.items():
        bank.add(i, d)
    bank.erase(3)
    bank.add(3, kfs[5][:50])                                           # re-added with other content
    kfs[3] = kfs[5][:50]
    bank.add(20, np.zeros((0, 32), np.uint8))                          # a keyframe without features
    kfs[20] = np.zeros((0, 32), np.uint8)
    assert len(bank) == 13 and np.array_equal(bank.descriptors(7), kfs[7])
    pairs = [(0, j) for j in (1, 2, 3, 4, 20)] + [(5, 6), (6, 5), (20, 1), (11, 11)]
    res = bank.match_pairs(pairs, cross_check=True, max_distance=60)
    for p, (a, b) in enumerate(pairs):
        _eq(res[p], orc.match(kfs[a], kfs[b], cross_check_=True, max_distance=60), (a, b))
    idx, dist, res = bank.match_pairs(pairs, k=2, ratio=0.9, want_knn=True)
    o = 0
    for p, (a, b) in enumerate(pairs):
        oi, od = c_oracle.knn(kfs[a], kfs[b], 2)
        n = len(kfs[a])
        assert np.array_equal(idx[o:o + n], oi) and np.array_equal(dist[o:o + n], od)
        _eq(res[p], orc.match(kfs[a], kfs[b], k=2, ratio=0.9), (a, b))
        o += n
    assert len(bank.match_pairs([])) == 0


@pytest.mark.parametrize("case", ["image", "wide", "negative", "tiny_radius", "big_radius", "nan"])
def test_binned_window_search_equals_brute_force(eng, case):
    """The binned projection-window path (bfm_window.cuh) and the brute-force window kernel return the
    same bits for any coordinate range (cells wrap, so nothing is assumed about the extent)."""
    rng = np.random.default_rng(len(case))
    nq, nt = 700, 5000
    q, t, qxy, txy, _ = synth.window_scene(nq, nt, 17)
    radius = 15.0
    if case == "wide":
        qxy, txy = qxy * 300.0, txy * 300.0
        radius = 4000.0
    elif case == "negative":
        qxy, txy = qxy - 5000.0, txy - 5000.0
    elif case == "tiny_radius":
        radius = 0.75
        txy[:400] = qxy[rng.integers(0, nq, 400)] + rng.uniform(-1, 1, (400, 2)).astype(np.float32)
    elif case == "big_radius":
        radius = 900.0                                                   # every pair admissible
    elif case == "nan":
        qxy[::13, 0] = np.nan
        txy[::7, 1] = np.nan
    dense = orc.window_mask(qxy, txy, radius)
    win = (qxy, txy, radius)
    for kw, okw in (({"k": 2, "ratio": 0.8}, {"k": 2, "ratio": 0.8}), ({"cross_check": True}, {"cross_check_": True}),
                    ({"k": 1, "max_distance": 50}, {"k": 1, "max_distance": 50})):
        want = orc.match(q, t, mask=dense, **okw)
        eng.set_tuning(window_bins=0)
        got_binned = eng.match(q, t, window=win, **kw)
        binned_kernels = eng.launch_info()["kernels_launched"]
        eng.set_tuning(window_bins=1)
        got_brute = eng.match(q, t, window=win, **kw)
        eng.set_tuning(window_bins=0)
        _eq(got_binned, want, (case, kw))
        _eq(got_brute, want, (case, kw))
        assert binned_kernels == 3                                       # rank + scatter + search
    oi, od = c_oracle.knn(q, t, 2, dense)
    _eq(eng.knn(q, t, 2, window=win), (oi, od), case)


def test_finalize_paths_agree(eng):
    """Resident inputs take the persistent form of the matching kernel (at most one wave of CTAs, each walking a run
    of work items, then finalizing the tiles the plan gave it); the gated host path keeps one work item per CTA and
    the CTA that completes a problem finalizes it.  Forced both ways (persistent knob) every mode must return the same
    bits as the oracle: plain, ratio, gate, cross-check, dense mask, window, k = 3, device tensors, and the
    local-map step (train count decided on the device) - always ONE launch per two neighbours."""
    import torch
    q, t, qxy, txy, _ = synth.window_scene(2100, 2600, 77)
    win = (qxy, txy, 15.0)
    dense = orc.window_mask(qxy, txy, 15.0)
    rng = np.random.default_rng(3)
    mask = (rng.random((2100, 2600)) < 0.3).astype(np.uint8) * 255
    want = dict(knn3=c_oracle.knn(q, t, 3), cc=c_oracle.cross_check(q, t), ratio=orc.match(q, t, k=2, ratio=0.8),
                gate=orc.match(q, t, cross_check_=True, max_distance=40), mknn=c_oracle.knn(q, t, 2, mask),
                ccm=orc.match(q, t, cross_check_=True, mask=mask), wknn=c_oracle.knn(q, t, 2, dense),
                wratio=orc.match(q, t, k=2, ratio=0.8, mask=dense))
    sc = synth.local_map_scene(6000, 6000, 2000, seed=5)
    store = bb.MapStore(6000, engine=eng)
    store.update(np.arange(6000), sc["desc"], sc["pt3d"], sc["normal"])
    targs = (sc["des"], sc["kp"], sc["R"], sc["t"], sc["see_vector"], sc["edges"])
    tracks = []
    try:
        for thr, kernels in ((2, 1), (1, 1)):
            eng.set_tuning(persistent=thr)
            _eq(eng.knn(q, t, 3), want["knn3"], thr)
            assert eng.launch_info()["kernels_launched"] == 2 * kernels
            _eq(eng.match(q, t, cross_check=True), want["cc"], thr)
            assert eng.launch_info()["kernels_launched"] == kernels
            _eq(eng.match(q, t, k=2, ratio=0.8), want["ratio"], thr)
            _eq(eng.match(q, t, cross_check=True, max_distance=40), want["gate"], thr)
            _eq(eng.knn(q, t, 2, mask=mask), want["mknn"], thr)
            _eq(eng.match(q, t, cross_check=True, mask=mask), want["ccm"], thr)
            eng.set_tuning(window_bins=1)                                # brute-force window kernel
            _eq(eng.knn(q, t, 2, window=win), want["wknn"], thr)
            _eq(eng.match(q, t, k=2, ratio=0.8, window=win), want["wratio"], thr)
            eng.set_tuning(window_bins=0)
            qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
            idx, dist = eng.knn(qd, td, 2)
            _eq((idx.cpu().numpy(), dist.cpu().numpy()), (want["knn3"][0][:, :2], want["knn3"][1][:, :2]), thr)
            r = store.track(*targs)
            r2 = store.track(*targs, cross_check=False, k=2, ratio=0.8, max_distance=None, window_radius=15.0)
            tracks.append([np.array(x) for x in (r.visible_edges, r.inds_frame, r.inds, r.distance, r.pts3d, r.kp,
                                                 r2.inds_frame, r2.inds, r2.distance)])
            _eq(eng.match(q[:300], t[:500], cross_check=True), c_oracle.cross_check(q[:300], t[:500]), thr)   # clean workspace
        _eq(tracks[0], tracks[1], "local-map step")
        assert len(tracks[0][1]) > 0 and len(tracks[0][6]) > 0
    finally:
        eng.set_tuning(persistent=0, window_bins=0)


def test_large_single_problem_tile_parallel_finalize(eng):
    """A single problem with many query rows is finalized tile by tile, in parallel, by the CTAs the plan names
    (inside the one launch): same results, workspace left clean for the next call."""
    q, t, _ = synth.correlated(9001, 700, 123)
    oi, od = c_oracle.knn(q, t, 3)
    for k in (1, 2, 3):
        idx, dist = eng.knn(q, t, k)
        assert np.array_equal(idx, oi[:, :k]) and np.array_equal(dist, od[:, :k]), k
        assert eng.launch_info()["kernels_launched"] == (k + 1) // 2
    _eq(eng.match(q, t, cross_check=True), c_oracle.cross_check(q, t))
    _eq(eng.match(q, t, k=2, ratio=0.8), orc.match(q, t, k=2, ratio=0.8))
    _eq(eng.match(q, t, cross_check=True, max_distance=25), orc.match(q, t, cross_check_=True, max_distance=25))
    q2, t2, _ = synth.correlated(500, 800, 124)                       # a normal call right after: clean workspace
    _eq(eng.match(q2, t2, cross_check=True), c_oracle.cross_check(q2, t2))
    _eq(eng.knn(q2, t2, 2), c_oracle.knn(q2, t2, 2))


def test_pinned_views_keep_their_memory_alive(eng):
    """A BatchResult built on HostBatchBuffers stays valid after the buffers object is dropped."""
    import gc
    qs, ts = synth.keyframe_pairs(4, 300, seed=5)
    tab = bb.make_problems([300] * 4, [300] * 4)
    ob = bb.HostBatchBuffers(1200, 4, k=2)
    res = eng.match_batched(np.concatenate(qs), np.concatenate(ts), tab, k=2, ratio=0.8, out=ob)
    want = [orc.match(qs[p], ts[p], k=2, ratio=0.8) for p in range(4)]
    del ob
    gc.collect()
    junk = [bb.PinnedBuffer((1200, 3), np.int32) for _ in range(4)]   # would reuse the freed block
    for j in junk:
        j.array[...] = -5
    for p in range(4):
        _eq(res[p], want[p], p)


def test_sm_fed_upload_from_pinned_arrays(eng):
    """Pinned caller arrays: the first CTAs of the matching kernel stream them into HBM themselves (no
    copy-engine chunks).  Same bits as the copy-engine gate and as the plain single-copy path; window
    coordinate arrays with an odd number of rows exercise the 8-byte tail."""
    P, N = 9, 1111                                                     # odd row counts everywhere
    qp, tp = synth.keyframe_pair_batch(P, N, seed=31)
    rng = np.random.default_rng(5)
    qxy = rng.uniform(0, 640, (P * N, 2)).astype(np.float32)
    txy = (qxy + rng.normal(0, 6, (P * N, 2))).astype(np.float32)      # row i of t near row i of q
    tab = bb.make_problems([N] * P, [N] * P)
    pin = []
    for a in (qp, tp, qxy, txy):
        b = bb.PinnedBuffer(a.shape, a.dtype)
        b.array[...] = a
        pin.append(b)
    pq, pt, pqxy, ptxy = (b.array for b in pin)
    cases = [dict(k=2, ratio=0.8), dict(cross_check=True, max_distance=60), dict(k=2, ratio=0.9, window=(pqxy, ptxy, 12.0))]
    for kw in cases:
        results = []
        for chunks, feeders in ((1, 0), (5, 0), (5, 3), (5, 32), (5, -1)):   # plain, SM-fed (16 / 3 / 32 feeders), copy engine
            eng.set_tuning(pipeline_chunks=chunks, feeders=feeders)
            out = bb.HostBatchBuffers(P * N, P, k=kw.get("k", 1), want_knn=True)
            idx, dist, res = eng.match_batched(pq, pt, tab, want_knn=True, out=out, **kw)
            results.append((idx.copy(), dist.copy(), res.counts.copy(), [tuple(x.copy() for x in res[p]) for p in range(P)]))
        eng.set_tuning(pipeline_chunks=0, feeders=0)
        for r in results[1:]:
            assert np.array_equal(r[0], results[0][0]) and np.array_equal(r[1], results[0][1]) and np.array_equal(r[2], results[0][2])
            for a, b in zip(r[3], results[0][3]):
                _eq(a, b)
        if "window" not in kw:
            okw = {("cross_check_" if k_ == "cross_check" else k_): v for k_, v in kw.items()}
            for p in (0, 4, 8):
                _eq(results[1][3][p], orc.match(qp[p * N:(p + 1) * N], tp[p * N:(p + 1) * N], **okw), p)


@pytest.mark.skipif(not ref.HAVE_CV2, reason="cv2 not importable")
def test_maximum_train_size_against_cv2(eng):
    """Rule R7: cv2 accepts up to 2^18 - 1 train rows; this engine up to 2^22 - 1.  At cv2's maximum the
    results must agree (trainIdx needs 18 bits of the packed key), and the limits raise where documented."""
    nt = (1 << 18) - 1
    t = synth.uniform(nt, 77)
    q = synth.uniform(48, 78)
    q[:24] = t[np.random.default_rng(1).integers(0, nt, 24)]          # exact hits, some far into the array
    q[0] = t[nt - 1]
    ri, rd = ref.knn(q, t, 2)
    idx, dist = eng.knn(q, t, 2)
    assert np.array_equal(idx, ri) and np.array_equal(dist, rd)
    assert idx[0, 0] == nt - 1 and dist[0, 0] == 0
    rq, rt, rdd = ref.match(q, t, cross_check=True)
    _eq(eng.match(q, t, cross_check=True), (rq, rt, rdd))
    with pytest.raises(ValueError):
        eng.knn(q, np.zeros((1 << 22, 32), np.uint8), 1)                # beyond the packed-key index range
    big_t = np.zeros(((1 << 22) - 1, 32), np.uint8)                     # the engine's own maximum
    big_t[-1] = 255
    idx, dist = eng.knn(np.full((2, 32), 255, np.uint8), big_t, 2)
    assert idx.tolist() == [[(1 << 22) - 2, 0]] * 2 and dist.tolist() == [[0, 256]] * 2


def test_input_gate_times_out_instead_of_hanging(eng):
    """If the upload never arrives the waiting CTAs give up after 4 s, the call fails with a clear error and the
    engine stays usable (the workspace is re-initialised on the next call)."""
    P, N = 6, 600
    qp, tp = synth.keyframe_pair_batch(P, N, seed=41)
    tab = bb.make_problems([N] * P, [N] * P)
    pq, pt = bb.PinnedBuffer(qp.shape), bb.PinnedBuffer(tp.shape)
    pq.array[...] = qp
    pt.array[...] = tp
    eng.set_tuning(pipeline_chunks=4, test_stall=1)
    try:
        with pytest.raises(bb.BfmError) as ei:
            eng.match_batched(pq.array, pt.array, tab, k=2, ratio=0.8)
        assert "timed out" in str(ei.value)
    finally:
        eng.set_tuning(pipeline_chunks=0, test_stall=0)
    res = eng.match_batched(pq.array, pt.array, tab, k=2, ratio=0.8)
    for p in range(P):
        _eq(res[p], orc.match(qp[p * N:(p + 1) * N], tp[p * N:(p + 1) * N], k=2, ratio=0.8), p)
    q, t, _ = synth.correlated(300, 500, 42)
    _eq(eng.match(q, t, cross_check=True), c_oracle.cross_check(q, t))


def test_two_threads_run_gated_batches_concurrently():
    """Tracking thread + local-mapping thread (slam/main.py:37-47), each with its own engine, both on the host
    path with the SM-fed upload: two gated kernels share the GPU without starving each other's feeders."""
    import threading
    P, N = 12, 900
    jobs = []
    for s in range(2):
        qp, tp = synth.keyframe_pair_batch(P, N, seed=50 + s)
        want = [orc.match(qp[p * N:(p + 1) * N], tp[p * N:(p + 1) * N], k=2, ratio=0.8) for p in (0, 5, 11)]
        jobs.append((qp, tp, want))
    tab = bb.make_problems([N] * P, [N] * P)
    errs = []

    def work(job):
        try:
            qp, tp, want = job
            eng = bb.Engine(0)
            eng.set_tuning(pipeline_chunks=6)
            pq, pt = bb.PinnedBuffer(qp.shape), bb.PinnedBuffer(tp.shape)
            pq.array[...] = qp
            pt.array[...] = tp
            out = bb.HostBatchBuffers(P * N, P, k=2)
            for _ in range(25):
                res = eng.match_batched(pq.array, pt.array, tab, k=2, ratio=0.8, out=out)
                for w, p in zip(want, (0, 5, 11)):
                    _eq(res[p], w, p)
            eng.close()
        except Exception as e:  # pragma: no cover
            errs.append(repr(e))

    th = [threading.Thread(target=work, args=(j,)) for j in jobs]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errs, errs


def test_misuse_is_reported_through_the_abi(eng):
    """Bad problem tables and options come back as BfmError with a message (no crash, no exception across the ABI)."""
    q, t, _ = synth.correlated(64, 64, 1)
    for tab, what in ((np.array([[0, 64, 0, 64, 0, 0], [0, 64, 0, 64, 32, 0]], np.int32), "overlap"),
                      (np.array([[0, 65, 0, 64, 0, 0]], np.int32), "out of range"),
                      (np.array([[0, 64, 10, 64, 0, 0]], np.int32), "out of range"),
                      (np.array([[-1, 64, 0, 64, 0, 0]], np.int32), "out of range")):
        with pytest.raises(bb.BfmError) as ei:
            eng.match_batched(q, t, tab)
        assert what in str(ei.value)
    with pytest.raises(ValueError):
        eng.match(q, t, k=3, ratio=0.8)
    with pytest.raises((ValueError, bb.BfmError)):
        eng.match(q, t, window=(np.zeros((3, 2), np.float32), np.zeros((64, 2), np.float32), 5.0))
    _eq(eng.match(q, t, cross_check=True), c_oracle.cross_check(q, t))   # still usable


def test_batch_plan_fast_path(eng):
    """Engine.plan_batch validates once; BatchPlan.run returns what match_batched returns."""
    P, N = 10, 700
    qp, tp = synth.keyframe_pair_batch(P, N, seed=61)
    tab = bb.make_problems([N] * P, [N] * P)
    out = bb.HostBatchBuffers(P * N, P, k=2)
    for kw in (dict(k=2, ratio=0.8), dict(cross_check=True, max_distance=40), dict(k=1, max_distance=0, strict=True)):
        plan = eng.plan_batch(tab, **kw)
        want = eng.match_batched(qp, tp, tab, **kw)
        for _ in range(3):
            got = plan.run(qp, tp, out)
            assert np.array_equal(got.counts, want.counts)
            for p in range(P):
                _eq(got[p], want[p], p)
    with pytest.raises(ValueError):
        plan.run(qp[:100], tp, out)
    with pytest.raises(ValueError):
        plan.run(qp.astype(np.int8), tp, out)


@pytest.mark.skipif(not ref.HAVE_CV2, reason="cv2 not importable")
def test_installed_into_cv2_call_sites_run_verbatim():
    """boslam_b200.install(cv2): the reference's own lines (slam/tracking.py:45,56-60) run unchanged on this
    engine and produce what OpenCV produces."""
    import cv2
    d_hamming_max = 30                                                # config.py:13
    frame_des, kf_des, _ = synth.correlated(600, 600, 71)             # camera.py:50: nfeatures=600

    def call_site():
        matcher = cv2.BFMatcher_create(cv2.NORM_HAMMING, crossCheck=True)           # slam/tracking.py:45
        matches = matcher.match(frame_des, kf_des)                                  # :56
        matches = [_ for _ in matches if _.distance < d_hamming_max]                # :57
        inds_f, inds_kf = zip(*((_.queryIdx, _.trainIdx) for _ in matches))         # :60
        return type(matcher), inds_f, inds_kf, [m.distance for m in matches]

    theirs = call_site()
    bb.install(cv2)
    try:
        mine = call_site()
    finally:
        bb.uninstall(cv2)
    assert mine[0] is bb.BFMatcher and theirs[0] is not bb.BFMatcher
    assert mine[1:] == theirs[1:] and len(mine[1]) > 100
