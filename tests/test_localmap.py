"""Frame-to-local-map tracking: the oracle's own consistency on CPU, and the
CUDA path (bfm_map_* / bfm_track_local_map) against it, bit for bit, on the GPU."""
import math

import numpy as np
import pytest

import boslam_b200 as bb
from boslam_b200 import synth
from oracle import localmap_oracle as lmo

COS60 = math.cos(math.pi / 3)


def _oracle(sc, **kw):
    return lmo.track_local_map(sc["desc"], sc["pt3d"], sc["normal"], sc["edges"], sc["des"], sc["kp"], sc["R"], sc["t"],
                               sc["see_vector"], sc["fx"], sc["fy"], sc["cx"], sc["cy"], sc["width"], sc["height"], COS60, **kw)


def test_quaternion_restatements_agree_and_rotate_like_R():
    rng = np.random.default_rng(0)
    for i in range(200):
        A = rng.normal(size=(3, 3))
        Q, _ = np.linalg.qr(A)
        if np.linalg.det(Q) < 0:
            Q[:, 0] *= -1
        q1 = bb.quaternion_from_rotation(Q)
        q2 = lmo.quaternion_from_rotation(Q)
        assert q1 == q2                                   # product mirror == oracle, bit for bit
        assert q1[0] >= 0 and abs(sum(c * c for c in q1) - 1) < 1e-14
        X = rng.normal(size=(5, 3))
        x, y, z = lmo.project(Q, np.zeros(3), X)
        assert np.allclose(np.stack([x, y, z], 1), X @ Q.T, atol=1e-12)


def test_oracle_visibility_rules_known_answers():
    R, t, see = np.eye(3), np.zeros(3), np.array([0.0, 0.0, 1.0])
    fx = fy = 100.0
    cx, cy = 320.0, 240.0
    pts = np.array([[0, 0, 1.0],        # centre -> (320, 240): visible
                    [-3.2, 0, 1.0],     # u = 0 exactly: visible (0 <= u)
                    [3.2, 0, 1.0],      # u = 640 exactly: not visible (u < width)
                    [0, 2.4, 1.0],      # v = 480 exactly: not visible
                    [0, 0, -1.0],       # behind the camera but projects to the centre: the reference keeps it
                    [0, 0, 0.0]])       # z = 0 -> nan: dropped
    n_ok = np.tile([1.0, 0, 0], (6, 1))           # dot = 0 < cos60
    ok, pix = lmo.visible(R, t, see, pts, n_ok, fx, fy, cx, cy, 640, 480, COS60)
    assert ok.tolist() == [True, True, False, False, True, False]
    assert pix[0].tolist() == [320.0, 240.0]
    n_bad = np.tile([0, 0, 1.0], (6, 1))          # dot = 1 >= cos60: nothing passes (as written at :104)
    assert not lmo.visible(R, t, see, pts, n_bad, fx, fy, cx, cy, 640, 480, COS60)[0].any()


def test_oracle_track_matches_a_literal_loop():
    """The vectorised oracle equals a literal per-edge loop in the shape of slam/tracking.py:97-121."""
    sc = synth.local_map_scene(300, 450, 120, seed=3)
    want = _oracle(sc)
    feats, pts3d, kept = [], [], []
    for e, slot in enumerate(sc["edges"]):
        ok, _ = lmo.visible(sc["R"], sc["t"], sc["see_vector"], sc["pt3d"][slot:slot + 1], sc["normal"][slot:slot + 1],
                            sc["fx"], sc["fy"], sc["cx"], sc["cy"], sc["width"], sc["height"], COS60)
        if ok[0]:
            kept.append(e)
            feats.append(sc["desc"][slot])
            pts3d.append(sc["pt3d"][slot])
    assert kept == want["visible_edges"].tolist() and 0 < len(kept) < len(sc["edges"])
    from oracle import hamming_oracle as orc
    mq, mt, md = orc.match(sc["des"], np.stack(feats), cross_check_=True, max_distance=30)
    assert np.array_equal(mq, want["inds_frame"]) and np.array_equal(mt, want["inds"])
    assert np.array_equal(np.stack(pts3d)[mt], want["pts3d"]) and len(mq) > 20


# ---------------------------------------------------------------------------------------------- GPU
def _same(got, want):
    assert np.array_equal(got.visible_edges, want["visible_edges"])
    if got.visible_pixels is not None:
        assert np.array_equal(got.visible_pixels, want["visible_pixels"])    # bit-exact fp64
    assert np.array_equal(got.inds_frame, want["inds_frame"]) and np.array_equal(got.inds, want["inds"])
    assert np.array_equal(got.distance, want["distance"]) and np.array_equal(got.edges, want["edges"])
    assert np.array_equal(got.pts3d, want["pts3d"]) and np.array_equal(got.kp, want["kp"])


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(300, 450, 120), (5000, 8000, 1000), (20000, 20000, 2000), (1, 1, 1), (700, 257, 0)])
def test_track_local_map_bit_exact(shape):
    n_points, n_edges, n_frame = shape
    sc = synth.local_map_scene(n_points, n_edges, n_frame, seed=n_points)
    store = bb.MapStore(n_points)
    store.update(np.arange(n_points), sc["desc"], sc["pt3d"], sc["normal"])
    args = (sc["des"], sc["kp"], sc["R"], sc["t"], sc["see_vector"], sc["edges"])
    _same(store.track(*args, want_pixels=True), _oracle(sc))                          # the reference's call shape
    assert store.track(*args).visible_pixels is None
    _same(store.track(*args, cross_check=False, k=2, ratio=0.8, max_distance=None),
          _oracle(sc, cross_check=False, k=2, ratio=0.8, max_distance=None))
    _same(store.track(*args, cross_check=False, k=2, ratio=0.8, max_distance=None, window_radius=15.0),
          _oracle(sc, cross_check=False, k=2, ratio=0.8, max_distance=None, window_radius=15.0))  # north-star shape
    _same(store.track(*args, window_radius=15.0, want_pixels=True), _oracle(sc, window_radius=15.0))
    _same(store.track(*args, max_distance=30, strict=True), _oracle(sc, max_distance=30, strict=True))


@pytest.mark.gpu
def test_map_store_updates_and_poses():
    sc = synth.local_map_scene(2000, 3000, 400, seed=11)
    store = bb.MapStore(4096)
    half = np.arange(0, 2000, 2)
    store.update(half, sc["desc"][half], sc["pt3d"][half], sc["normal"][half])      # two batches, partial fields
    rest = np.arange(1, 2000, 2)
    store.update(rest, desc=sc["desc"][rest])
    store.update(rest, pt3d=sc["pt3d"][rest], normal=sc["normal"][rest])
    args = (sc["des"], sc["kp"], sc["R"], sc["t"], sc["see_vector"], sc["edges"])
    _same(store.track(*args), _oracle(sc))
    # descriptor refresh of some map points (slam/nodes.py:153): results follow the store
    rng = np.random.default_rng(1)
    ch = rng.choice(2000, 300, replace=False)
    sc["desc"][ch] = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    store.update(ch, desc=sc["desc"][ch])
    _same(store.track(*args), _oracle(sc))
    # other poses, including one whose rotation takes the non-positive-trace branch of the conversion
    for R in (np.diag([1.0, -1.0, -1.0]), np.diag([-1.0, 1.0, -1.0]), np.array([[0, -1.0, 0], [1.0, 0, 0], [0, 0, 1.0]])):
        sc2 = dict(sc, R=R, see_vector=R @ np.array([0, 0, 1.0]))
        _same(store.track(sc2["des"], sc2["kp"], R, sc2["t"], sc2["see_vector"], sc2["edges"]), _oracle(sc2))
    with pytest.raises(bb.BfmError):
        store.update([5000], desc=np.zeros((1, 32), np.uint8))
    with pytest.raises(ValueError):
        store.update([1, 1], desc=np.zeros((2, 32), np.uint8))


# ------------------------------------------------------------------- representative descriptor (8(f) row 4)
def _obs_batch(P, max_obs, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (P, 1, 32), dtype=np.uint8)
    flips = np.packbits(rng.random((P, max_obs, 256)) < 0.06, axis=2)
    obs = base ^ flips                                            # noisy views of one descriptor per point
    counts = rng.integers(0, max_obs + 1, P).astype(np.int32)
    obs[::7] = obs[::7, :1]                                       # all-identical observations: every median ties -> index 0
    if max_obs > 1:
        obs[3::11, 1] = obs[3::11, 0]                             # duplicate pairs
    return obs, counts


def test_representative_oracle_known_answers():
    from oracle import representative_oracle as ro
    z, o = np.zeros(32, np.uint8), np.full(32, 255, np.uint8)
    assert ro.hamming(z, o) == 256 and ro.hamming(z, z) == 0
    assert ro.select([z]) == 0 and ro.select([z, o]) == 0            # n = 2: both medians equal -> lowest index
    far = z.copy(); far[:5] = 255                                    # 40 bits from z
    near = z.copy(); near[31] = 1                                    # 1 bit from z
    assert ro.select([far, z, near]) == 1                            # medians 40, 1, 1 -> lowest index of the tie
    assert ro.select([far, near, z, near]) == 1                      # medians 40.5, 0.5, 1, 0.5
    assert ro.select([o, z, z, z]) == 1


@pytest.mark.gpu
@pytest.mark.parametrize("max_obs", [1, 2, 5, 10, 16])
def test_select_representative_matches_reference_loop(max_obs):
    from oracle import representative_oracle as ro
    obs, counts = _obs_batch(777, max_obs, seed=max_obs)
    got = bb.select_representative(obs, counts)
    assert np.array_equal(got, ro.select_batch(obs, counts))
    assert (got[counts == 0] == -1).all()
    full = np.full(777, max_obs, np.int32)
    assert np.array_equal(bb.select_representative(obs, full), ro.select_batch(obs, full))


@pytest.mark.gpu
def test_keyframe_vote_matches_counter_most_common():
    """SURVEY 8(f) row 3, reference slam/tracking.py:154: the vote for the next reference keyframe over the inliers of
    the pose optimisation - ids, counts AND the order Counter.most_common yields (count, then first appearance)."""
    eng = bb.Engine(0)
    sc = synth.local_map_scene(4000, 6000, 1500, seed=21)
    store = bb.MapStore(4000, engine=eng)
    store.update(np.arange(4000), sc["desc"], sc["pt3d"], sc["normal"])
    r = store.track(sc["des"], sc["kp"], sc["R"], sc["t"], sc["see_vector"], sc["edges"])
    assert len(r.inds) > 50
    rng = np.random.default_rng(5)
    ne = len(sc["edges"])
    for trial, n_kf in enumerate((1, 3, 17, 400)):
        # keyframe ids as boslam makes them (a counter), negative and large ones too; edges grouped by keyframe or not
        ids = rng.choice(np.concatenate([np.arange(-3, 50), [2 ** 31 - 1, -2 ** 31]]), n_kf, replace=n_kf > 50)
        edge_kf = np.sort(rng.integers(0, n_kf, ne)) if trial % 2 == 0 else rng.integers(0, n_kf, ne)
        edge_kf = ids[edge_kf].astype(np.int32)
        for inliers in (np.arange(len(r.inds)), rng.permutation(len(r.inds))[:len(r.inds) // 2], np.zeros(0, np.int64), np.array([0, 0, 1, 0])):
            for top in (100, 2, 1):
                want = lmo.keyframe_vote(edge_kf, r.visible_edges, r.inds, inliers, top)
                assert store.vote(edge_kf, inliers, top) == want, (n_kf, len(inliers), top)
    # misuse: the vote reads the match list of the last track call on the device
    store.update(np.arange(2), sc["desc"][:2])
    with pytest.raises(bb.BfmError):
        store.vote(np.zeros(ne, np.int32), [0])
    eng.close()
