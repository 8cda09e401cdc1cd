"""Host logic of the CUDA library, on a box without a GPU: the work-item planner (bfm_plan_preview).

Every CTA of the matching kernel takes one work item = (block of 128 x R query rows) x (contiguous train range) of
one problem.  The result is a min over packed (distance, index) keys, so ANY exact tiling gives the same bits;
what the planner must guarantee is that the tiling IS exact - every (query row, train row) pair of every problem
is covered exactly once - and the balance properties the kernel's throughput relies on."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import boslam_b200 as bb
from boslam_b200 import _ffi

SLOTS = 148 * 8          # CTAs resident on a B200 (148 SMs x 8 per SM)
Q_ROW0, Q_VALID, Q_LOCAL0, OUT_ROW0, T_ROW0, T_COUNT, T_LOCAL0, PROBLEM = range(8)


def _check_exact_tiling(tab, items, r):
    bq = 128 * r
    assert (np.diff(items[:, PROBLEM]) >= 0).all(), "work items are in problem order (the upload chases them)"
    for p, (q_begin, q_count, t_begin, t_count, out_begin, _) in enumerate(tab.tolist()):
        mine = items[items[:, PROBLEM] == p]
        if q_count == 0 or t_count == 0:
            # an empty problem still gets ONE item: the CTA that completes a problem finalizes it
            assert len(mine) == 1 and mine[0, T_COUNT] == 0 and mine[0, Q_VALID] == 0
            assert mine[0, OUT_ROW0] == out_begin
            continue
        n_blocks = (q_count + bq - 1) // bq
        assert sorted(set(mine[:, Q_LOCAL0].tolist())) == [b * bq for b in range(n_blocks)]
        for b in range(n_blocks):
            blk = mine[mine[:, Q_LOCAL0] == b * bq]
            assert (blk[:, Q_ROW0] == q_begin + b * bq).all() and (blk[:, OUT_ROW0] == out_begin + b * bq).all()
            assert (blk[:, Q_VALID] == min(bq, q_count - b * bq)).all()
            # the train ranges of one query block tile [0, t_count) exactly, in order, without gaps or overlaps
            assert blk[0, T_LOCAL0] == 0
            assert (blk[1:, T_LOCAL0] == blk[:-1, T_LOCAL0] + blk[:-1, T_COUNT]).all()
            assert blk[-1, T_LOCAL0] + blk[-1, T_COUNT] == t_count
            assert (blk[:, T_COUNT] > 0).all()
            assert (blk[:, T_ROW0] == t_begin + blk[:, T_LOCAL0]).all()


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 5000), st.integers(0, 30000)), min_size=1, max_size=40),
       st.sampled_from([1, 2, 4]), st.sampled_from([0, 1, 2, 4, 8]), st.sampled_from([0, 5, 10, 35, 90]),
       st.sampled_from([0, 0, 0, 33, 128, 1000]), st.sampled_from([0, 0, 1, 3]))
def test_plan_tiles_every_problem_exactly(shapes, r, taper, pct, seg_rows, waves):
    tab = bb.make_problems([q for q, _ in shapes], [t for _, t in shapes])
    items, L = _ffi.plan_preview(tab, r, SLOTS, seg_rows, waves, taper, pct)
    assert L >= 1 and len(items) >= len(shapes)
    _check_exact_tiling(tab, items, r)
    if seg_rows:
        assert L == seg_rows and items[:, T_COUNT].max() <= seg_rows


def test_headline_batch_plan():
    """256 x (2000 x 2000), R = 4: ~7 waves of equal CTAs, then a tapered tail (1/2, then 1/4 of the length over
    the last 10 % of the work) - the numbers bench.py reports as launch.scan_grid."""
    tab = bb.make_problems([2000] * 256, [2000] * 256)
    flat, L = _ffi.plan_preview(tab, 4, SLOTS, taper=1)
    assert len(flat) == 8192 and set(flat[:, T_COUNT].tolist()) == {250}
    assert abs(len(flat) / SLOTS - round(len(flat) / SLOTS)) < 0.1           # close to whole waves
    items, L2 = _ffi.plan_preview(tab, 4, SLOTS)                              # auto: tapered
    assert L2 == L and len(items) == 9564
    _check_exact_tiling(tab, items, 4)
    per_problem = np.array([items[items[:, PROBLEM] == p][:, T_COUNT].max() for p in range(256)])
    assert (per_problem[:230] == 250).all()                                   # the first 90 % of the work: full length
    assert (np.diff(per_problem) <= 0).all()                                  # never longer towards the end
    assert per_problem[-1] <= L // 4 and per_problem[235] <= L // 2                 # L = 285 rows asked, 250 cut
    # work is conserved
    assert int((items[:, Q_VALID].astype(np.int64) * items[:, T_COUNT]).sum()) == 256 * 2000 * 2000


def test_small_and_single_problem_plans_are_not_tapered():
    for shapes in ([(2000, 20000)], [(1000, 1000)], [(2000, 2000)] * 7):
        tab = bb.make_problems([q for q, _ in shapes], [t for _, t in shapes])
        a, _ = _ffi.plan_preview(tab, 2, SLOTS)
        b, _ = _ffi.plan_preview(tab, 2, SLOTS, taper=1)
        assert np.array_equal(a, b)
        _check_exact_tiling(tab, a, 2)


def test_balance_of_the_wave_aware_cut():
    """All work items of a flat plan cost about the same (within one row per segment), whatever the shape."""
    rng = np.random.default_rng(1)
    for _ in range(20):
        n = int(rng.integers(8, 300))
        tab = bb.make_problems([2000] * n, [int(rng.integers(1500, 2500))] * n)
        items, L = _ffi.plan_preview(tab, 4, SLOTS, taper=1)
        full = items[items[:, Q_VALID] == 512]
        assert full[:, T_COUNT].max() - full[:, T_COUNT].min() <= 1
        assert full[:, T_COUNT].max() <= L


def test_plan_preview_rejects_bad_arguments():
    tab = bb.make_problems([10], [10])
    for kw in (dict(queries_per_thread=3), dict(slots=0), dict(taper=3), dict(taper_pct=91), dict(segment_rows=-1)):
        with pytest.raises(bb.BfmError):
            _ffi.plan_preview(tab, **kw)
    bad = tab.copy()
    bad[0, 1] = -1
    with pytest.raises(bb.BfmError):
        _ffi.plan_preview(bad)


# ---- the persistent form (resident inputs): finalize tiles and their owners ---------------------------------------

def _check_tiles(tab, tiles, tile_cta, n_ctas):
    assert (np.diff(tile_cta) >= 0).all() and tile_cta.min() >= 0 and tile_cta.max() < n_ctas, "a CTA's tiles are one contiguous run"
    slot = 0
    for p, (_, q_count, _, _, _, _) in enumerate(tab.tolist()):
        sel = tiles[:, 0] == p
        mine, owners = tiles[sel], tile_cta[sel]
        order = np.argsort(mine[:, 2], kind="stable")
        mine, owners = mine[order], owners[order]
        nt = max(1, -(-q_count // 512))
        # every problem's rows exactly once, 512 rows per tile (an empty problem still gets the tile that writes its count)
        assert len(mine) == nt and mine[:, 2].tolist() == list(range(nt)) and (mine[:, 3] == nt).all()
        assert mine[:, 1].tolist() == [512 * j for j in range(nt)]
        assert (mine[:, 4] == slot).all(), "look-back slots are dense and disjoint"
        slot += nt
        # no CTA waits for a CTA dispatched after it: owners do not decrease along the look-back chain
        assert (np.diff(owners) >= 0).all()
    # within one CTA the tiles come in (problem, tile) order, so a look-back never waits for a later tile of the same CTA
    for c in np.unique(tile_cta):
        mine = tiles[tile_cta == c]
        key = mine[:, 0].astype(np.int64) * (1 << 20) + mine[:, 2]
        assert (np.diff(key) > 0).all()


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(0, 9000), min_size=1, max_size=60), st.sampled_from([SLOTS, 148 * 6, 37, 3, 1]))
def test_tile_owners_never_wait_for_a_later_cta(q_counts, n_ctas):
    tab = bb.make_problems(q_counts, [100] * len(q_counts))
    tiles, tile_cta = _ffi.plan_preview_tiles(tab, n_ctas)
    _check_tiles(tab, tiles, tile_cta, n_ctas)


def test_tile_owners_are_the_lowest_quarter_of_the_grid():
    """Owners are dealt from CTA 0 upwards, one tile each while they last and never more than a quarter of the grid: the
    CTAs that own no tile exit when the queue is empty, so a grid that is only partly resident (a shared GPU) always gets
    its remaining CTAs started."""
    tab = bb.make_problems([2000] * 32, [2000] * 32)                 # a rank's share at 8 GPUs: 128 tiles
    tiles, tile_cta = _ffi.plan_preview_tiles(tab, SLOTS)
    _check_tiles(tab, tiles, tile_cta, SLOTS)
    assert len(tiles) == 128 and tile_cta.tolist() == list(range(128))
    tab = bb.make_problems([300] * 1000, [300] * 1000)               # many small problems: more tiles than a quarter of the grid
    tiles, tile_cta = _ffi.plan_preview_tiles(tab, SLOTS)
    _check_tiles(tab, tiles, tile_cta, SLOTS)
    assert tile_cta.max() == SLOTS // 4 - 1 and np.bincount(tile_cta).max() <= 4
    one = bb.make_problems([70000], [70000])
    tiles, tile_cta = _ffi.plan_preview_tiles(one, 64)               # more tiles than CTAs: consecutive tiles share a CTA, in order
    _check_tiles(one, tiles, tile_cta, 64)
    assert len(tiles) == 137 and tile_cta.max() == 15 and np.bincount(tile_cta).max() <= 9


# ---- the tensor form's planner (plan_items_tensor) and the copy chunks of its host path ---------------------------------

def _check_tensor_plan(tab, n_sms=148):
    items, seg_rows, planes = _ffi.plan_preview_tensor(tab, n_sms)
    assert seg_rows % 128 == 0
    xq = xt = 0
    by_problem = {}
    for it in items:
        by_problem.setdefault(int(it[7]), []).append(it)
    for p, (qb, qc, tb, tc, ob, _) in enumerate(np.asarray(tab)):
        if qc == 0 or tc == 0:
            assert p not in by_problem            # empty problems have no work item
        else:
            its = by_problem[p]
            blocks = {}
            for q_row0, q_valid, q_local0, out_row0, t_row0, t_count, t_local0, _ in its:
                assert q_local0 % 256 == 0 and q_valid == min(256, qc - q_local0) and q_valid > 0
                assert q_row0 == xq + q_local0 and out_row0 == ob + q_local0
                assert t_row0 == xt + t_local0 and t_row0 % 8 == 0 and q_row0 % 8 == 0      # 1024-byte swizzle atoms
                assert 0 < t_count <= seg_rows
                assert t_local0 % 128 == 0 and (t_count % 128 == 0 or t_local0 + t_count == tc)
                blocks.setdefault(int(q_local0), []).append((int(t_local0), int(t_count)))
            assert sorted(blocks) == list(range(0, qc, 256))                                  # every query block, once
            for ranges in blocks.values():                                                    # ... against the whole train set, once
                pos = 0
                for t0, n in sorted(ranges):
                    assert t0 == pos
                    pos += n
                assert pos == tc
        xq += (qc + 7) // 8 * 8
        xt += (tc + 7) // 8 * 8
    assert planes == (xq, xt)
    return items


def test_tensor_plan_tiles_every_problem_exactly_once():
    rng = np.random.default_rng(5)
    _check_tensor_plan(bb.make_problems([2000] * 256, [2000] * 256))
    _check_tensor_plan(bb.make_problems([1], [1]))
    _check_tensor_plan(bb.make_problems([300, 0, 257, 2000, 5, 640, 1, 900, 0, 1300], [500, 40, 0, 2000, 2500, 129, 1, 7, 0, 1111]))
    for _ in range(20):
        P = int(rng.integers(1, 40))
        _check_tensor_plan(bb.make_problems(rng.integers(0, 3000, P), rng.integers(0, 5000, P)), n_sms=int(rng.choice([1, 8, 148])))
    _check_tensor_plan(bb.make_problems([2000] * 20, [2000] * 20, shared_query=True))


def test_tensor_plan_cuts_long_train_ranges_to_fill_the_sms():
    """One large problem: whole train ranges per item would leave most SMs idle, so the ranges are cut (multiples of 128
    rows); a batch that already fills the machine keeps whole ranges (one A-tile load and one commit per query block)."""
    items = _check_tensor_plan(bb.make_problems([8192], [8192]))
    assert len(items) >= 96 and len(items) <= 148                  # 32 query blocks x 3-4 train ranges: one round
    items = _check_tensor_plan(bb.make_problems([65536], [65536]))
    assert len(items) % 256 == 0 and len(items) >= 1024            # 256 query blocks x >= 4 ranges
    items = _check_tensor_plan(bb.make_problems([2000] * 256, [2000] * 256))
    assert len(items) == 2048 and set(items[:, 5]) == {2000}


def test_host_chunks_are_whole_rounds_of_the_scan_for_equal_shapes():
    tab = bb.make_problems([2000] * 256, [2000] * 256)
    assert _ffi.plan_preview_host_chunks(tab, 512000, 512000) == (4, 74)          # 74 pairs = 592 items = 4 x 148
    assert _ffi.plan_preview_host_chunks(tab[:128], 256000, 256000) == (2, 74)    # a rank's share at 2 GPUs
    assert _ffi.plan_preview_host_chunks(tab[:32], 64000, 64000) == (2, 16)       # ... at 8 GPUs: shorter than a round
    assert _ffi.plan_preview_host_chunks(tab, 512000, 512000, forced=7) == (7, 37)
    for forced in range(1, 9):
        c, per = _ffi.plan_preview_host_chunks(tab, 512000, 512000, forced=forced)
        assert 1 <= c <= 8 and (c - 1) * per < 256 <= c * per
    rng = np.random.default_rng(6)
    rag = bb.make_problems(rng.integers(1200, 2100, 56), rng.integers(1500, 2300, 56))
    c, per = _ffi.plan_preview_host_chunks(rag, int(rag[:, 1].sum()), int(rag[:, 3].sum()))
    assert (c, per) == (2, 28)                                                     # ragged: equal numbers of problems
    with pytest.raises(_ffi.BfmError):
        _ffi.plan_preview_host_chunks(tab, 512000, 512000, forced=9)
