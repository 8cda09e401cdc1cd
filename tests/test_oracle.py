"""Pins the CPU oracle (numpy + C restatements) against the reference implementation:
(a) the committed cv2-generated golden vectors, (b) live cv2.BFMatcher on seeded inputs,
(c) the known-answer rules R1-R10 of SURVEY.md section 8(c)."""
import numpy as np
import pytest

from boslam_b200 import synth
from oracle import c_oracle, cv2_reference as ref, hamming_oracle as orc

needs_cv2 = pytest.mark.skipif(not ref.HAVE_CV2, reason="cv2 not importable")


def _names(golden):
    return [str(n) for n in golden["names"]]


def test_golden_numpy_oracle(golden):
    for name in _names(golden):
        q, t = golden[f"{name}/q"], golden[f"{name}/t"]
        for k in (1, 2, 3):
            idx, dist = orc.knn(q, t, k)
            assert np.array_equal(idx, golden[f"{name}/knn{k}_idx"]), (name, k)
            assert np.array_equal(dist, golden[f"{name}/knn{k}_dist"]), (name, k)
        qi, ti, d = orc.cross_check(q, t)
        assert np.array_equal(qi, golden[f"{name}/cc_q"]), name
        assert np.array_equal(ti, golden[f"{name}/cc_t"]), name
        assert np.array_equal(d, golden[f"{name}/cc_d"]), name
        idx, dist = orc.knn(q, t, 2, golden[f"{name}/mask"])
        assert np.array_equal(idx, golden[f"{name}/mknn2_idx"]), name
        assert np.array_equal(dist, golden[f"{name}/mknn2_dist"]), name
        rq, rt, rd = orc.match(q, t, k=1, ratio=0.8)
        assert np.array_equal(rq, golden[f"{name}/ratio_q"]), name
        assert np.array_equal(rt, golden[f"{name}/ratio_t"]), name
        assert np.array_equal(rd, golden[f"{name}/ratio_d"]), name


def test_golden_c_oracle(golden):
    for name in _names(golden):
        q, t = golden[f"{name}/q"], golden[f"{name}/t"]
        for k in (1, 2, 3):
            idx, dist = c_oracle.knn(q, t, k)
            assert np.array_equal(idx, golden[f"{name}/knn{k}_idx"]), (name, k)
            assert np.array_equal(dist, golden[f"{name}/knn{k}_dist"]), (name, k)
        qi, ti, d = c_oracle.cross_check(q, t)
        assert np.array_equal(qi, golden[f"{name}/cc_q"])
        assert np.array_equal(ti, golden[f"{name}/cc_t"])
        assert np.array_equal(d, golden[f"{name}/cc_d"])
        idx, dist = c_oracle.knn(q, t, 2, golden[f"{name}/mask"])
        assert np.array_equal(idx, golden[f"{name}/mknn2_idx"])
        assert np.array_equal(dist, golden[f"{name}/mknn2_dist"])


@needs_cv2
@pytest.mark.parametrize("seed", range(4))
def test_live_cv2_differential(seed):
    rng = np.random.default_rng(seed)
    nq, nt = int(rng.integers(1, 300)), int(rng.integers(1, 400))
    gens = [
        lambda: synth.correlated(nq, nt, seed)[:2],
        lambda: (synth.tie_stress(nq, seed), synth.tie_stress(nt, seed + 50)),
        lambda: (synth.uniform(nq, seed), synth.duplicate_rows(max(1, nt // 3), seed)),
    ]
    for g in gens:
        q, t = g()
        for k in (1, 2, 4):
            ri, rd = ref.knn(q, t, k)
            for impl in (orc, c_oracle):
                oi, od = impl.knn(q, t, k)
                assert np.array_equal(oi, ri) and np.array_equal(od, rd)
        rq, rt, rd = ref.match(q, t, cross_check=True)
        for impl in (orc, c_oracle):
            oq, ot, od = impl.cross_check(q, t)
            assert np.array_equal(oq, rq) and np.array_equal(ot, rt) and np.array_equal(od, rd)
        mask = (rng.random((len(q), len(t))) < 0.2).astype(np.uint8)
        ri, rd = ref.knn(q, t, 2, mask)
        for impl in (orc, c_oracle):
            oi, od = impl.knn(q, t, 2, mask)
            assert np.array_equal(oi, ri) and np.array_equal(od, rd)
        a = orc.match(q, t, ratio=0.8)
        b = ref.ratio_match(q, t, 0.8)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


@needs_cv2
def test_live_cv2_config1_shape():
    """BASELINE config 1: 1000 x 1000 crossCheck (the reference's live call shape)."""
    q, t, _ = synth.correlated(1000, 1000, 5)
    rq, rt, rd = ref.match(q, t, cross_check=True)
    oq, ot, od = c_oracle.cross_check(q, t)
    assert np.array_equal(oq, rq) and np.array_equal(ot, rt) and np.array_equal(od, rd)
    nq, nt, nd = orc.cross_check(q, t)
    assert np.array_equal(nq, rq) and np.array_equal(nt, rt) and np.array_equal(nd, rd)


def test_rule_r1_distance():
    a = np.zeros((1, 32), np.uint8)
    b = np.full((1, 32), 255, np.uint8)
    assert orc.hamming_matrix(a, b)[0, 0] == 256
    b[0, 31] = 0x0F
    assert orc.hamming_matrix(a, b)[0, 0] == 252  # the last byte counts too
    assert orc.hamming_matrix(a, a)[0, 0] == 0


def test_rule_r2_ties_lowest_index():
    a = synth.uniform(1, 1)[0]
    b = synth.uniform(1, 2)[0]
    train = np.stack([a, a, b, b, b])
    d = int(np.bitwise_count(a ^ b).sum())
    idx, dist = orc.knn(a[None], train, 3)
    assert idx.tolist() == [[0, 1, 2]] and dist.tolist() == [[0, 0, d]]
    idx, dist = c_oracle.knn(a[None], train, 3)
    assert idx.tolist() == [[0, 1, 2]] and dist.tolist() == [[0, 0, d]]


def test_rule_r3_short_rows_and_r8_empty():
    q = synth.uniform(3, 3)
    t = synth.uniform(2, 4)
    idx, dist = orc.knn(q, t, 3)
    assert (idx[:, 2] == -1).all() and (dist[:, 2] == -1).all() and (idx[:, :2] >= 0).all()
    e = np.zeros((0, 32), np.uint8)
    assert orc.knn(e, t, 2)[0].shape == (0, 2)
    assert (orc.knn(q, e, 2)[0] == -1).all()
    assert all(len(x) == 0 for x in orc.cross_check(q, e))
    assert all(len(x) == 0 for x in orc.match(e, t))


def test_rule_r4_mask_and_r5_r6_crosscheck():
    q = synth.uniform(6, 5)
    t = synth.uniform(7, 6)
    m1 = np.ones((6, 7), np.uint8)
    m255 = m1 * 255
    assert np.array_equal(orc.knn(q, t, 2, m1)[0], orc.knn(q, t, 2, m255)[0])
    m1[2, :] = 0
    assert (orc.knn(q, t, 2, m1)[0][2] == -1).all()
    # duplicates: only the lowest-index duplicate can be a mutual match (R5)
    t2 = np.concatenate([t[:1], t[:1], t[1:]])
    q2 = np.concatenate([t[:1], t[:1]])
    qi, ti, d = orc.cross_check(q2, t2)
    assert qi.tolist() == [0] and ti.tolist() == [0] and d.tolist() == [0]
    # masked cross-check: masking the mutual pair frees both sides (R6 definition)
    m = np.ones((2, len(t2)), np.uint8)
    m[0, 0] = 0
    qi, ti, d = orc.cross_check(q2, t2, m)
    cq, ct, cd = c_oracle.cross_check(q2, t2, m)
    assert np.array_equal(qi, cq) and np.array_equal(ti, ct) and np.array_equal(d, cd)
    assert (0, 1) in set(zip(qi.tolist(), ti.tolist())) or (1, 0) in set(zip(qi.tolist(), ti.tolist()))


def test_ratio_fp64_boundary():
    """80 < 0.8*100 is False in fp64 (SURVEY 8(c) caller-side filters)."""
    q = np.zeros((1, 32), np.uint8)
    t = np.zeros((2, 32), np.uint8)
    t[0, :10] = 255  # distance 80
    t[1, :12] = 255
    t[1, 12] = 0x0F  # distance 100
    assert orc.knn(q, t, 2)[1].tolist() == [[80, 100]]
    assert len(orc.match(q, t, ratio=0.8)[0]) == 0
    assert len(orc.match(q, t, ratio=0.81)[0]) == 1


def test_window_equals_dense_mask():
    q, t, qxy, txy, _ = synth.window_scene(120, 300, 7)
    m = orc.window_mask(qxy, txy, 15.0)
    a = orc.match(q, t, k=2, ratio=0.8, window=(qxy, txy, 15.0))
    b = orc.match(q, t, k=2, ratio=0.8, mask=m)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert 0 < m.mean() < 0.05


def test_gates_strict_and_nonstrict():
    q, t, _ = synth.correlated(200, 200, 9)
    qi, ti, d = orc.match(q, t, cross_check_=True)
    a = orc.match(q, t, cross_check_=True, max_distance=10, strict=True)
    b = orc.match(q, t, cross_check_=True, max_distance=10, strict=False)
    assert np.array_equal(a[0], qi[d < 10]) and np.array_equal(b[0], qi[d <= 10])


# ---- property tests: the two restatements (numpy, plain C) must agree with each other on anything -------------
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=60, deadline=None)
@given(nq=st.integers(1, 40), nt=st.integers(1, 60), k=st.integers(1, 5), seed=st.integers(0, 10 ** 6),
       kind=st.sampled_from(["uniform", "ties", "dups"]), masked=st.booleans())
def test_numpy_and_c_oracles_agree(nq, nt, k, seed, kind, masked):
    from oracle import c_oracle
    if kind == "uniform":
        q, t = synth.uniform(nq, seed), synth.uniform(nt, seed + 1)
    elif kind == "ties":
        q, t = synth.tie_stress(nq, seed), synth.tie_stress(nt, seed + 1)
    else:
        q, t = synth.uniform(nq, seed), synth.duplicate_rows(max(1, nt // 2), seed + 1)
    mask = None
    if masked:
        mask = (np.random.default_rng(seed).random((nq, len(t))) < 0.5).astype(np.uint8)
    ai, ad = orc.knn(q, t, k, mask)
    bi, bd = c_oracle.knn(q, t, k, mask)
    assert np.array_equal(ai, bi) and np.array_equal(ad, bd)
    # rows ascend by (distance, index); unfilled slots are -1 and trail
    for i in range(nq):
        keys = [(int(d), int(j)) for d, j in zip(ad[i], ai[i]) if j >= 0]
        assert keys == sorted(keys) and all(j == -1 for j in ai[i][len(keys):])
    a = orc.cross_check(q, t, mask) if hasattr(orc, "cross_check") else orc.match(q, t, cross_check_=True, mask=mask)
    b = c_oracle.cross_check(q, t, mask)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    # a cross-check match is mutual: the reverse problem contains the mirrored pair
    ra = orc.match(t, q, cross_check_=True, mask=None if mask is None else np.ascontiguousarray(mask.T))
    assert set(zip(a[0].tolist(), a[1].tolist())) == set(zip(ra[1].tolist(), ra[0].tolist()))
