"""Round-2 robustness cases: stream hand-over between calls on one handle, coherent loads on the SM-fed path,
option validation of the tracking entry point, outputs of all-empty batches, the documented divergences from
cv2 / the reference pinned to the behaviour this engine chose."""
import ctypes
import threading

import numpy as np
import pytest

import boslam_b200 as bb
from boslam_b200 import _ffi, synth
from oracle import c_oracle, cv2_reference as ref, hamming_oracle as orc, representative_oracle as ro

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = bb.Engine(0)
    yield e
    e.close()


def _eq(a, b, what=""):
    for x, y in zip(a, b):
        assert np.array_equal(np.asarray(x), np.asarray(y)), what


def test_calls_on_different_streams_share_one_workspace_safely(eng):
    """Device calls are asynchronous and all calls of a handle share its workspace and device tables: a call that
    arrives on another stream (or a host call on the handle's own stream) must wait for the previous one
    (bfm_api.cu).  Alternate two torch streams and the host path, with batches long enough that the
    kernels would overlap without the hand-over."""
    import torch
    P, N = 24, 1500
    sets = []
    for s in range(3):
        q, t = synth.keyframe_pair_batch(P, N, seed=50 + s)
        sets.append((q, t, torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()))
    tab = bb.make_problems([N] * P, [N] * P)
    small_q, small_t, _ = synth.correlated(700, 900, 3)
    want_small = orc.match(small_q, small_t, cross_check_=True, max_distance=40)
    want = []
    for q, t, _, _ in sets:
        want.append([orc.match(q[p * N:(p + 1) * N], t[p * N:(p + 1) * N], k=2, ratio=0.8) for p in (0, P // 2, P - 1)])
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for rep in range(6):
        outs = []
        for i, st in enumerate((s1, s2, s1)):
            with torch.cuda.stream(st):
                outs.append(eng.match_batched_device(sets[i][2], sets[i][3], tab, k=2, ratio=0.8))   # queued, not waited for
        got_small = eng.match(small_q, small_t, cross_check=True, max_distance=40)                   # host path, own stream
        torch.cuda.synchronize()
        _eq(got_small, want_small, f"rep {rep}: host call after device calls")
        for i, o in enumerate(outs):
            m, c = o["m"].cpu().numpy(), o["count"].cpu().numpy()
            for j, p in enumerate((0, P // 2, P - 1)):
                n = int(c[p])
                _eq((m[0, p * N:p * N + n], m[1, p * N:p * N + n], m[2, p * N:p * N + n]), want[i][j], f"rep {rep} set {i} pair {p}")


def test_sm_fed_window_batch_with_half_sector_row_counts(eng):
    """Window + pinned batched input where the train row count is 16 mod 32 (a 32-byte sector of the pixel
    coordinate array then straddles two feed rounds).  Feeder-written arrays are read with coherent (.cg) loads and
    round boundaries are 128-byte aligned; every feeder configuration must return the single-copy path's bits."""
    P = 7
    nq, nt = 500, 1200 + 16                      # 1216 % 32 == 0 ... use 1232: 1232 % 32 == 16
    nt = 1232
    assert nt % 32 == 16
    rng = np.random.default_rng(9)
    q = synth.uniform(P * nq, 60)
    t = synth.uniform(P * nt, 61)
    for p in range(P):                           # plant true matches so the window decides something
        src = rng.integers(0, nt, nq)
        q[p * nq:(p + 1) * nq] = t[p * nt + src] ^ np.packbits(rng.random((nq, 256)) < 0.03, axis=1)
    qxy = rng.uniform(0, 640, (P * nq, 2)).astype(np.float32)
    txy = rng.uniform(0, 640, (P * nt, 2)).astype(np.float32)
    tab = bb.make_problems([nq] * P, [nt] * P)
    pins = []
    for a in (q, t, qxy, txy):
        b = bb.PinnedBuffer(a.shape, a.dtype)
        b.array[...] = a
        pins.append(b)
    pq, pt, pqxy, ptxy = (b.array for b in pins)
    results = []
    try:
        for chunks, feeders, rows in ((1, 0, 0), (4, 0, 0), (4, 5, 256), (4, 32, 128), (4, 2, 1024)):
            eng.set_tuning(pipeline_chunks=chunks, feeders=feeders, feed_rows=rows, pipeline_min_kb=1)
            for rep in range(3):
                out = bb.HostBatchBuffers(P * nq, P, k=2, want_knn=True)
                idx, dist, res = eng.match_batched(pq, pt, tab, k=2, ratio=0.9, window=(pqxy, ptxy, 80.0), want_knn=True, out=out)
                results.append((idx.copy(), dist.copy(), res.counts.copy()))
    finally:
        eng.set_tuning(pipeline_chunks=0, feeders=0, feed_rows=0, pipeline_min_kb=0)
    for r in results[1:]:
        assert np.array_equal(r[0], results[0][0]) and np.array_equal(r[1], results[0][1]) and np.array_equal(r[2], results[0][2])
    for p in (0, P - 1):
        mask = orc.window_mask(qxy[p * nq:(p + 1) * nq], txy[p * nt:(p + 1) * nt], 80.0)
        oi, od = c_oracle.knn(q[p * nq:(p + 1) * nq], t[p * nt:(p + 1) * nt], 2, mask=mask)
        assert np.array_equal(results[0][0][p * nq:(p + 1) * nq], oi) and np.array_equal(results[0][1][p * nq:(p + 1) * nq], od)


def test_track_local_map_validates_options(eng):
    """bfm_track_local_map skipped option validation (k <= 0 ran no kernel and left the match count
    unwritten).  Bad options are refused through the ABI; a valid call afterwards still works."""
    sc = synth.local_map_scene(300, 400, 120, seed=5)
    store = bb.MapStore(300, engine=eng)
    store.update(np.arange(300), sc["desc"], sc["pt3d"], sc["normal"])
    args = (sc["des"], sc["kp"], sc["R"], sc["t"], sc["see_vector"], sc["edges"])
    good = store.track(*args)
    L, h = _ffi.lib(), eng._h
    tp = _ffi.TrackParams()
    tp.q[0] = 1.0
    tp.fx = tp.fy = 300.0
    tp.cx, tp.cy, tp.width, tp.height, tp.cos_max = 320.0, 240.0, 640, 480, 0.5
    edges = np.ascontiguousarray(sc["edges"], np.int32)
    des = np.ascontiguousarray(sc["des"])
    kp = np.ascontiguousarray(sc["kp"], np.float64)
    nv, nm = ctypes.c_int32(123), ctypes.c_int32(456)

    def call(**kw):
        o = _ffi.Options()
        o.k, o.cross_check, o.mask_kind, o.max_distance, o.ratio = 1, 1, 0, 30, -1.0
        for k_, v in kw.items():
            setattr(o, k_, v)
        return L.bfm_track_local_map(store._m, ctypes.byref(tp), edges.ctypes.data, len(edges), des.ctypes.data, kp.ctypes.data,
                                     len(des), ctypes.byref(o), None, None, None, None, None, None, None, None,
                                     ctypes.byref(nv), ctypes.byref(nm))
    for bad in (dict(k=0), dict(k=-3), dict(k=3, cross_check=0), dict(k=2, cross_check=1), dict(cross_check=1, ratio=0.8),
                dict(mask_kind=1), dict(mask_kind=7)):
        assert call(**bad) == _ffi.BFM_ERR_INVALID, bad
        assert L.bfm_last_error(h)
    assert call() == _ffi.BFM_OK and 0 <= nm.value <= len(des)
    again = store.track(*args)
    assert np.array_equal(again.inds, good.inds) and np.array_equal(again.visible_edges, good.visible_edges)


def test_all_empty_batches_still_write_their_counts(eng):
    """with every query set empty no kernel runs, but m_count is an output per problem: it must read 0,
    on the host and on the device path.  A dense mask with a row stride below t_count is refused."""
    import torch
    t = synth.uniform(50, 1)
    q0 = np.zeros((0, 32), np.uint8)
    tab = np.array([[0, 0, 0, 20, 0, 0], [0, 0, 20, 30, 0, 0]], np.int32)
    L, h = _ffi.lib(), eng._h
    o = _ffi.Options()
    o.k, o.ratio, o.max_distance = 1, -1.0, -1
    cnt = np.full(2, 77, np.int32)
    mq = np.zeros(4, np.int32)
    pp = tab.ctypes.data_as(ctypes.POINTER(_ffi.Problem))
    rc = L.bfm_match_batched(h, _ffi.MEM_HOST, None, 0, t.ctypes.data, 50, pp, 2, 0, ctypes.byref(o), None, None,
                             mq.ctypes.data, mq.ctypes.data, mq.ctypes.data, cnt.ctypes.data, None)
    assert rc == _ffi.BFM_OK and cnt.tolist() == [0, 0]
    dt = torch.from_numpy(t).cuda()
    dcnt = torch.full((2,), 77, dtype=torch.int32, device="cuda")
    dm = torch.zeros(4, dtype=torch.int32, device="cuda")
    rc = L.bfm_match_batched(h, _ffi.MEM_DEVICE, None, 0, dt.data_ptr(), 50, pp, 2, 0, ctypes.byref(o), None, None,
                             dm.data_ptr(), dm.data_ptr(), dm.data_ptr(), dcnt.data_ptr(), ctypes.c_void_p(-1))
    torch.cuda.synchronize()
    assert rc == _ffi.BFM_OK and dcnt.cpu().tolist() == [0, 0]
    res = eng.match_batched(q0, t, tab, k=1)
    assert res.counts.tolist() == [0, 0]
    # dense mask whose rows are shorter than the train set
    q = synth.uniform(8, 2)
    mask = np.ones((8, 50), np.uint8)
    o2 = _ffi.Options()
    o2.k, o2.ratio, o2.max_distance, o2.mask_kind, o2.mask, o2.mask_row_stride = 1, -1.0, -1, _ffi.MASK_DENSE, mask.ctypes.data, 49
    one = np.array([[0, 8, 0, 50, 0, 0]], np.int32)
    idx = np.zeros((8, 1), np.int32)
    rc = L.bfm_match_batched(h, _ffi.MEM_HOST, q.ctypes.data, 8, t.ctypes.data, 50, one.ctypes.data_as(ctypes.POINTER(_ffi.Problem)),
                             1, 8, ctypes.byref(o2), idx.ctypes.data, idx.ctypes.data, None, None, None, None, None)
    assert rc == _ffi.BFM_ERR_INVALID and b"mask_row_stride" in L.bfm_last_error(h)


def test_divergence_multi_image_collection_with_a_short_image(eng):
    """Documented divergence 1: a collection in which one image has fewer than k rows.  cv2 4.13
    returns truncated rows; this engine returns the k nearest over the WHOLE collection, mapped back to
    (imgIdx, per-image trainIdx).  Pins the chosen behaviour against the oracle over the concatenation."""
    imgs = [synth.correlated(200, n, 70 + i)[1] for i, n in enumerate((300, 2, 150))]
    q = synth.correlated(200, 300, 70)[0]
    m = bb.BFMatcher_create(bb.NORM_HAMMING)
    m.add(imgs)
    rows = m.knnMatch(q, k=3)
    cat, bounds = np.concatenate(imgs), np.cumsum([0] + [len(i) for i in imgs])
    oi, od = c_oracle.knn(q, cat, 3)
    assert len(rows) == len(q) and all(len(r) == 3 for r in rows)
    img = np.searchsorted(bounds, oi, side="right") - 1
    assert [[(x.trainIdx, x.imgIdx, x.distance) for x in r] for r in rows] == \
           [[(int(oi[i, c] - bounds[img[i, c]]), int(img[i, c]), float(od[i, c])) for c in range(3)] for i in range(len(q))]
    # crossCheck knnMatch over a collection carries imgIdx / per-image trainIdx like match() does (matcher.py)
    mc = bb.BFMatcher_create(bb.NORM_HAMMING, crossCheck=True)
    mc.add(imgs)
    a = mc.match(q)
    b = mc.knnMatch(q, k=1, compactResult=True)
    assert [(x.queryIdx, x.trainIdx, x.imgIdx, x.distance) for x in a] == [(r[0].queryIdx, r[0].trainIdx, r[0].imgIdx, r[0].distance) for r in b]
    assert {x.imgIdx for x in a} - {0} and len(mc.knnMatch(q, k=1)) == len(q)


def test_divergence_representative_distance_256_does_not_wrap(eng):
    """Documented divergence 2: reference slam/nodes.py:148 stores distances in a uint8 array, so a Hamming distance of
    exactly 256 (complementary descriptors) wraps to 0 there; this engine (and oracle/representative_oracle.py) keep
    integer distances.  Pins the chosen behaviour with observations that contain exact complements."""
    rng = np.random.default_rng(12)
    base = rng.integers(0, 256, (30, 32), dtype=np.uint8)
    obs = np.zeros((30, 10, 32), np.uint8)
    obs[:, 0] = ~base                            # distance exactly 256 to observation 1
    obs[:, 1] = base
    obs[:, 2] = base ^ np.packbits(rng.random((30, 256)) < 0.02, axis=1)
    cnt = np.full(30, 3, np.int32)
    got = bb.select_representative(obs, cnt, engine=eng)
    want = ro.select_batch(obs, cnt)
    assert np.array_equal(got, want)
    # true medians per column: [~251, d12, d12] -> observation 1 (lowest index of the tie)
    assert got.tolist() == [1] * 30
    # the literal uint8 arithmetic of the reference sees D[0][1] = D[1][0] = 0, medians [0, 0, d12] -> observation 0
    wrapped = []
    for p in range(30):
        d = np.zeros((3, 3), np.uint8)
        for i in range(3):
            for j in range(3):
                d[i, j] = np.uint8(int(np.unpackbits(obs[p, i] ^ obs[p, j]).sum()) & 0xFF)
        wrapped.append(int(np.argmin(np.median(d, axis=0))))
    assert wrapped == [0] * 30


def test_default_engines_of_dead_threads_are_closed(eng):
    """default_engine cached one handle per thread id forever."""
    from boslam_b200 import engine as E
    q, t, _ = synth.correlated(64, 64, 1)
    made = []

    def work():
        made.append(bb.default_engine(0))
        assert len(bb.match(q, t, cross_check=True)[0]) > 0
    for _ in range(3):
        th = threading.Thread(target=work)
        th.start()
        th.join()
    bb.match(q, t, cross_check=True)             # this thread's own default engine: creation evicts the dead ones
    th = threading.Thread(target=work)
    th.start()
    th.join()
    with E._default_lock:
        alive = [k for k, (e, th_) in E._default_engines.items() if th_.is_alive()]
        dead = [k for k, (e, th_) in E._default_engines.items() if not th_.is_alive()]
    assert len(alive) >= 1 and len(dead) <= 1


@pytest.mark.parametrize("persistent", [1, 2])
def test_same_shapes_other_rows_reuse_the_plan(eng, persistent):
    """A loop-closing step names other keyframes than the step before, of the same sizes: the cached cut of the work is
    re-based on the new rows instead of being planned again.  Results must follow the rows - in both forms of the
    kernel, with an empty keyframe in the list, for k = 2 + ratio and for cross-check."""
    sizes = [300, 300, 0, 450, 300, 450, 300, 300, 450, 300]
    kfs = {i: synth.correlated(max(n, 1), max(n, 1), 700 + i)[1][:n] for i, n in enumerate(sizes)}
    bank = bb.KeyframeBank(capacity_rows=8192, engine=eng)
    for i, d in kfs.items():
        bank.add(i, d)
    same = lambda a: [j for j in kfs if len(kfs[j]) == len(kfs[a])]
    step1 = [(0, 1), (3, 5), (2, 0), (4, 6), (8, 3), (7, 9)]
    # the same shapes, every keyframe replaced by another one of its size (so every row offset differs)
    step2 = [(9, 7), (5, 8), (2, 4), (6, 1), (3, 5), (0, 4)]
    assert [(len(kfs[a]), len(kfs[b])) for a, b in step1] == [(len(kfs[a]), len(kfs[b])) for a, b in step2]
    try:
        eng.set_tuning(persistent=persistent)
        for kw, okw in ((dict(k=2, ratio=0.9), dict(k=2, ratio=0.9)), (dict(cross_check=True, max_distance=70), dict(cross_check_=True, max_distance=70))):
            for step in (step1, step2, step1, step2[::-1]):
                res = bank.match_pairs(step, **kw)
                for p, (a, b) in enumerate(step):
                    _eq(res[p], orc.match(kfs[a], kfs[b], **okw), (persistent, kw, a, b))
        # the bound form (copy=False: views of the bank's pinned buffers, arguments marshalled once): same lists
        for step in (step1, step2, step1, step2[::-1], step1):
            res = bank.match_pairs(step, k=2, ratio=0.9, copy=False)
            for p, (a, b) in enumerate(step):
                _eq(res[p], orc.match(kfs[a], kfs[b], k=2, ratio=0.9), (persistent, "bound", a, b))
    finally:
        eng.set_tuning(persistent=0)


def test_window_batches_take_the_binned_search(eng):
    """A batch of projection-window problems (ragged, with an empty train set, an empty query set and a one-row problem)
    goes through the binned search like a single one does: three launches for the whole batch, and the same bits as
    the brute-force window kernel and the oracle (dense mask built from the same predicate) - k = 2 + ratio, k = 1 +
    gate, cross-check, device tensors and host arrays."""
    import torch
    shapes = [(300, 900), (1, 1), (450, 0), (0, 200), (700, 2500), (64, 64), (500, 1300)]
    qs, ts, qx, tx = [], [], [], []
    for i, (nq, nt) in enumerate(shapes):
        q, t, qxy, txy, _ = synth.window_scene(max(nq, 1), max(nt, 1), 40 + i)
        qs.append(q[:nq]); ts.append(t[:nt]); qx.append(qxy[:nq]); tx.append(txy[:nt])
    qp, tp = np.concatenate(qs), np.concatenate(ts)
    qxy, txy = np.concatenate(qx), np.concatenate(tx)
    tab = bb.make_problems([a for a, _ in shapes], [b for _, b in shapes])
    radius = 14.0
    dense = [orc.window_mask(qx[i], tx[i], radius) if shapes[i][0] and shapes[i][1] else None for i in range(len(shapes))]
    for kw, okw in ((dict(k=2, ratio=0.85), dict(k=2, ratio=0.85)), (dict(k=1, max_distance=50), dict(k=1, max_distance=50)),
                    (dict(cross_check=True), dict(cross_check_=True))):
        want = []
        for i, (nq, nt) in enumerate(shapes):
            want.append(orc.match(qs[i], ts[i], mask=dense[i], **okw) if nq and nt else (np.zeros(0, np.int32),) * 3)
        try:
            got = {}
            for bins in (0, 1):
                eng.set_tuning(window_bins=bins)
                got[bins] = eng.match_batched(qp, tp, tab, window=(qxy, txy, radius), **kw)
                if bins == 0:
                    assert eng.launch_info()["kernels_launched"] == 3, "count + scatter + search for the whole batch"
                dev = eng.match_batched(torch.from_numpy(qp).cuda(), torch.from_numpy(tp).cuda(), tab,
                                        window=(torch.from_numpy(qxy).cuda(), torch.from_numpy(txy).cuda(), radius), **kw)
                for p in range(len(shapes)):
                    _eq(got[bins][p], want[p], (kw, bins, p))
                    _eq(dev[p], want[p], (kw, bins, p, "device"))
        finally:
            eng.set_tuning(window_bins=0)
    # the workspace is clean afterwards: an ordinary call right behind
    q2, t2, _ = synth.correlated(500, 800, 124)
    _eq(eng.match(q2, t2, cross_check=True), c_oracle.cross_check(q2, t2))


def test_first_gated_multi_pass_call_does_not_stall_on_a_kernel_load():
    """CUDA loads a kernel at its first launch and may wait for the device to go idle to do so.  A gated knnMatch with
    k > 2 launches its second pass while the first one spins on the upload; in a fresh process whose earlier calls used
    other kernels that stalled the host for the gate's 4 s time-out and failed the call.  The library loads the kernels
    of all passes before the first launch (tools/repro_gate.py, run in its own process)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "repro_gate.py")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("gated k=3 call:")][-1]
    assert "same tables" in line, line
    assert float(line.split()[-2]) < 1.0, line


def test_two_persistent_launches_share_the_gpu():
    """Two engines on two threads run persistent launches in which the finalize tiles outnumber their owners, at the same
    time: neither grid is fully resident, tiles wait for work items of CTAs that start later - every call must still
    return the right lists, and none may come near the 2 s of a tile's time-out (tools/share_gpu_probe.py)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "share_gpu_probe.py"), "80"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-1500:])
    assert "wrong results [0, 0]" in r.stdout
