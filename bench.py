#!/usr/bin/env python
"""bench.py - Hamming pairs/s (+ matched frames/s) of the B200 matcher vs cv2.BFMatcher on host.

Contract: ``python bench.py --gpus N --steps K --warmup W`` (under torchrun for N > 1) prints ONE
JSON line from rank 0.  ``--impl reference`` times the reference's own implementation of the path
(cv2.BFMatcher on the box's host cores, reference slam/tracking.py:45,56) on the same config.

Workload (config.workload = "loop_closing"): BASELINE.json configs[3], the configuration the
metric's "at 1/2/4/8 B200" is quoted on: ONE list of 256 BoW-candidate keyframe pairs x (2000 x 2000) ORB
descriptors per step, knn k=2 + ratio 0.8.  STRONG scaling: rank r of N takes pairs[r*256/N : (r+1)*256/N]
(SURVEY 8(e)), ONE kernel launch per rank and step, and every rank ends up with every rank's match tables -
written by the kernel epilogue into all ranks' symmetric buffers (NVSwitch multicast stores where available)
plus a symmetric-memory barrier per step, all inside the timed region; BFM_GATHER=nccl selects a plain NCCL
all_gather instead.  A "step" is one pass of the hot path over the whole list: 1.024 G descriptor pairs.
Round 1's weak-scaling form (256 pairs per rank) is the extra key "weak".  The tracking (configs[1]),
frame-to-frame (configs[0], with and without the CPU solvePnPRansac), local-mapping (configs[2], both variants)
shapes and the device-resident local-map step are extra keys of the same line ("frames"); configs[4] (the size
sweep) is tools/size_sweep.py -> profiles/.

  value      pairs/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e        same metric through the public host API (Engine.plan_batch(...).run: numpy in -> numpy out,
             pinned host buffers; the step's inputs are read from host memory and its results written to
             host memory inside the timed region, by the kernel itself); at N > 1 the same call also performs
             the exchange (host copies + match + gather + barrier in one number)
  roofline   scan kernel vs the measured POPC issue peak (integer pipe; 8 POPC per pair is the
             algorithmic count) - plus the HBM view for context
  cpu_baseline  cv2.BFMatcher.knnMatch(k=2) + Python ratio test on the host cores (N = 1 only); the same leg
             checks the engine's full knn tables and match lists of the batch against cv2's
  verify     N > 1: bit-equality of every rank's gathered table with an NCCL all_gather of the locally computed
             tables (gather_verified_full), ShardedMatcher through the CUDA engine; tables_crc32 of the whole
             result of input set 0 is the same number for every N
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_PAIRS, N_DESC, RATIO = int(os.environ.get("BFM_BENCH_PAIRS", "256")), 2000, 0.8   # (BFM_BENCH_PAIRS: diagnostics only; the bench line is the 256-pair list)


def _env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock + throttle reasons during the timed region (pynvml; nvidia-smi fallback)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _reason_names(self, mask):
        nv = self._nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80)}
        return [k for k, v in names.items() if mask & v]

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.reasons.update(self._reason_names(mask))
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(2)
        if self._nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
def cv2_step(matcher, q, t, tab, ratio):
    """One reference pass over a batch: what a boslam-style call site would run per pair."""
    n = 0
    for p in range(tab.shape[0]):
        qb, qc, tb, tc = int(tab[p, 0]), int(tab[p, 1]), int(tab[p, 2]), int(tab[p, 3])
        rows = matcher.knnMatch(q[qb:qb + qc], t[tb:tb + tc], 2)
        good = [r[0] for r in rows if len(r) == 2 and r[0].distance < ratio * r[1].distance]
        n += len(good)
    return n


def bench_config():
    """The ONE workload both arms run (identical dict in both JSON lines): BASELINE.json configs[3]."""
    return {"workload": "loop_closing", "pairs": N_PAIRS, "desc_per_keyframe": N_DESC, "k": 2, "ratio": RATIO,
            "pairs_per_step": N_PAIRS * N_DESC * N_DESC,
            "split": "ONE list of 256 candidate pairs per step; rank r of n_gpus takes pairs[r*256/n : (r+1)*256/n]",
            "l2": "rotating input sets, > 126 MB in total per rank (more than L2)"}


def cv2_tables(matcher, q, t, tab):
    """cv2.knnMatch(k=2) of every pair as dense int32 tables [rows, 2] (idx, dist), -1 padded (rule R3)."""
    n = int((tab[:, 4] + tab[:, 1]).max()) if len(tab) else 0
    idx = np.full((n, 2), -1, np.int32)
    dist = np.full((n, 2), -1, np.int32)
    for p in range(tab.shape[0]):
        qb, qc, tb, tc, ob = (int(tab[p, c]) for c in range(5))
        for i, row in enumerate(matcher.knnMatch(q[qb:qb + qc], t[tb:tb + tc], 2)):
            for c, dm in enumerate(row):
                idx[ob + i, c] = dm.trainIdx
                dist[ob + i, c] = int(dm.distance)
    return idx, dist


def reference_arm(args, emit):
    """--impl reference: cv2.BFMatcher on the host cores; same config / metric / unit, the full 256-pair batch per step."""
    rank = _env_int("RANK", 0)
    if rank != 0:
        return
    import boslam_b200.synth as synth
    from boslam_b200.engine import make_problems
    from oracle import cv2_reference as ref
    n_sets = 2   # rotating input sets (L2 is irrelevant to a host run; kept so both arms read fresh data every step)
    sets = [synth.keyframe_pair_batch(N_PAIRS, N_DESC, seed=s) for s in range(n_sets)]
    tab = make_problems([N_DESC] * N_PAIRS, [N_DESC] * N_PAIRS)
    pairs = N_PAIRS * N_DESC * N_DESC
    if ref.HAVE_CV2:
        m = ref.matcher(False)
        kind, cores = "reference", ref.threads()
        step = lambda i: cv2_step(m, sets[i % n_sets][0], sets[i % n_sets][1], tab, RATIO)
        sample = (f"the full {N_PAIRS}-pair batch per step, cv2 {ref.version()} knnMatch(k=2) + Python ratio test incl. DMatch "
                  f"construction, os.cpu_count()={os.cpu_count()}")
    else:  # plain-C port of the same algorithm, single thread: a bounded sample of the batch
        from oracle import c_oracle
        kind, cores, sp = "port", 1, 8
        pairs = sp * N_DESC * N_DESC
        step = lambda i: sum(len(c_oracle.knn(sets[0][0][p * N_DESC:(p + 1) * N_DESC], sets[0][1][p * N_DESC:(p + 1) * N_DESC], 2)[0])
                             for p in range(sp))
        sample = f"{sp} of {N_PAIRS} pairs per step, plain-C oracle, 1 thread"
    for i in range(max(args.warmup, 1)):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    value = pairs * args.steps / dt
    line = {
        "impl": "reference", "metric": "hamming_pairs_per_s", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": bench_config(),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
def time_device_loop(torch, fn, steps, barrier, finish=None):
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    if finish is not None:
        finish()   # e.g. order the timed stream after the last step's (side-stream) barrier
    e1.record()
    torch.cuda.synchronize()
    barrier()
    return e0.elapsed_time(e1)


def extras(eng, torch, steps):
    """Tracking / frame-to-frame / local-mapping shapes: device-resident and end-to-end frames/s."""
    import boslam_b200 as bb
    import boslam_b200.synth as synth
    from boslam_b200.engine import make_problems
    out = {}

    trace_buf = torch.zeros((16384, 8), dtype=torch.int64, device="cuda")

    def gpu_span_us(plan):
        """First CTA entry -> last CTA exit of ONE call on the GPU's own clock (bfm_debug_timeline: every CTA stamps
        %globaltimer): the kernel time proper, free of launch and event latency.  Median of 5."""
        spans = []
        for _ in range(5):
            trace_buf.zero_()
            eng._lib.bfm_debug_timeline(eng._h, ctypes.c_void_p(trace_buf.data_ptr()), trace_buf.shape[0])
            plan.run()
            torch.cuda.synchronize()
            eng._lib.bfm_debug_timeline(eng._h, None, 0)
            a = trace_buf.cpu().numpy()[:, :7]
            a = a[a[:, 0] > 0]
            if len(a):
                spans.append((a.max() - a[:, 0].min()) / 1e3)
        return float(np.median(spans)) if spans else None

    def run(name, host_fn, plan, frames_per_call, pairs_per_call, reps):
        """e2e: the public host call (numpy in -> numpy out), wall clock.  device: the same problem resident in HBM
        through a bound DevicePlan (the bare library call, ~4 us of host time): `ms_device` = back-to-back calls on one
        stream between two CUDA events (what a loop of such calls costs per call, launch gaps included),
        `ms_device_idle_stream` = one call between two events on an idle stream (adds the event / launch latency
        of an empty stream, ~9 us on this pool), `us_gpu_span` = the kernel itself on the GPU clock."""
        for _ in range(3):
            host_fn()
            plan.run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            host_fn()
        e2e = (time.perf_counter() - t0) / reps
        n_dev = max(reps, 50)
        st = torch.cuda.current_stream().cuda_stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_dev):
            plan.run(st)
        e1.record()
        torch.cuda.synchronize()
        dev = e0.elapsed_time(e1) / n_dev * 1e-3
        idle = []
        for _ in range(7):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            plan.run(st)
            b.record()
            torch.cuda.synchronize()
            idle.append(a.elapsed_time(b))
        li = eng.launch_info()
        out[name] = {"frames_per_s_e2e": frames_per_call / e2e, "frames_per_s_device": frames_per_call / dev,
                     "pairs_per_s_e2e": pairs_per_call / e2e, "pairs_per_s_device": pairs_per_call / dev,
                     "ms_e2e": e2e * 1e3, "ms_device": dev * 1e3, "ms_device_idle_stream": float(np.median(idle)),
                     "us_gpu_span": gpu_span_us(plan), "kernels_per_call": li["kernels_launched"], "grid": li["scan_grid"],
                     "work_items": li["segments"], "queries_per_thread": li["queries_per_thread"]}

    reps = max(steps, 10)
    # configs[0]: two frames, 1000 descriptors each, crossCheck + gate < 30 (slam/tracking.py:56-57)
    q, t, _ = synth.correlated(1000, 1000, 11)
    qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    tab1 = make_problems([1000], [1000])
    run("frame_to_frame_1000x1000_crosscheck",
        lambda: eng.match(q, t, cross_check=True, max_distance=30, strict=True),
        eng.plan_device(qd, td, tab1, cross_check=True, max_distance=30, strict=True),
        1, 1000 * 1000, reps)
    # configs[1]: tracking, 2000 frame descriptors vs 20k local-map points, window + ratio 0.8
    q, t, qxy, txy, _ = synth.window_scene(2000, 20000, 12)
    qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    qxyd, txyd = torch.from_numpy(qxy).cuda(), torch.from_numpy(txy).cuda()
    tab2 = make_problems([2000], [20000])
    run("tracking_2000x20000_window_ratio",
        lambda: eng.match(q, t, k=2, ratio=RATIO, window=(qxy, txy, 15.0)),
        eng.plan_device(qd, td, tab2, k=2, ratio=RATIO, window=(qxyd, txyd, 15.0)),
        1, 2000 * 20000, reps)
    run("tracking_2000x20000_crosscheck_gate30",  # the reference-faithful variant (slam/tracking.py:121)
        lambda: eng.match(q, t, cross_check=True, max_distance=30),
        eng.plan_device(qd, td, tab2, cross_check=True, max_distance=30),
        1, 2000 * 20000, reps)
    # configs[2]: local mapping, new keyframe vs 20 covisible keyframes, 2000 descriptors each
    qb, tb = synth.keyframe_pair_batch(20, 2000, 13)
    tab3 = make_problems([2000] * 20, [2000] * 20)
    qbd, tbd = torch.from_numpy(qb).cuda(), torch.from_numpy(tb).cuda()
    run("local_mapping_20x2000x2000_k2_ratio",
        lambda: eng.match_batched(qb, tb, tab3, k=2, ratio=RATIO),
        eng.plan_device(qbd, tbd, tab3, k=2, ratio=RATIO),
        1, 20 * 2000 * 2000, reps)
    run("local_mapping_20x2000x2000_crosscheck_gate30",   # the reference's local-mapping matcher is crossCheck=True (slam/local_mapping.py:21)
        lambda: eng.match_batched(qb, tb, tab3, cross_check=True, max_distance=30),
        eng.plan_device(qbd, tbd, tab3, cross_check=True, max_distance=30),
        1, 20 * 2000 * 2000, reps)
    # the headline batch from ordinary (pageable) numpy arrays into ordinary numpy arrays: what a caller who knows
    # nothing about pinned memory gets (worker threads stage the arrays for the kernel's feeder CTAs)
    qpg, tpg = synth.keyframe_pair_batch(N_PAIRS, N_DESC, seed=16)
    tabp = make_problems([N_DESC] * N_PAIRS, [N_DESC] * N_PAIRS)
    for _ in range(3):
        eng.match_batched(qpg, tpg, tabp, k=2, ratio=RATIO)
    t0 = time.perf_counter()
    for _ in range(10):
        eng.match_batched(qpg, tpg, tabp, k=2, ratio=RATIO)
    dtp = (time.perf_counter() - t0) / 10
    out["loop_closing_256x2000x2000_pageable_numpy"] = {"ms_e2e": dtp * 1e3, "pairs_per_s_e2e": N_PAIRS * N_DESC * N_DESC / dtp,
                                                        "note": "pageable numpy in -> freshly allocated numpy out"}

    # configs[0] in full (experiments/pnp_one_way_tracking.py:30-41): crossCheck match of two RGB-D frames, then
    # cv2.solvePnPRansac on the matched (3-D, 2-D) pairs.  PnP stays on the host CPU in both arms (north star).
    try:
        import cv2
        from oracle import cv2_reference as ref
        prev, cur, _, _ = synth.rgbd_frame_pair(1000, seed=15)
        cam_mat = np.array([[384.23901367, 0., 322.43237305], [0., 384.23901367, 239.65332031], [0., 0., 1.]])
        dist_coefs = np.zeros((8, 1), np.float32)

        def pnp(inds_prev, inds):
            a = np.ascontiguousarray(prev["cloud_kp"][inds_prev, :])
            b = np.ascontiguousarray(cur["kp_arr"][inds, :].astype(np.float64))
            return cv2.solvePnPRansac(a, b, cam_mat, dist_coefs)

        def ours():
            qi, ti, _ = eng.match(prev["des"], cur["des"], cross_check=True)
            return pnp(qi, ti), len(qi)

        def theirs():
            ms = ref.matcher(True).match(prev["des"], cur["des"])
            inds_prev, inds = zip(*((m.queryIdx, m.trainIdx) for m in ms))
            return pnp(np.asarray(inds_prev), np.asarray(inds)), len(ms)

        (ok_a, _, tv_a, inl_a), n_a = ours()
        (ok_b, _, tv_b, inl_b), n_b = theirs()

        def clock(f, n):
            t0 = time.perf_counter()
            for _ in range(n):
                f()
            return (time.perf_counter() - t0) / n
        t_ours, t_theirs = clock(ours, reps), clock(theirs, 5)
        t_match = clock(lambda: eng.match(prev["des"], cur["des"], cross_check=True), reps)
        out["frame_to_frame_1000x1000_match_plus_pnp"] = {
            "frames_per_s_e2e": 1.0 / t_ours, "ms_e2e": t_ours * 1e3, "ms_match_only": t_match * 1e3,
            "cv2_frames_per_s": 1.0 / t_theirs, "cv2_ms": t_theirs * 1e3, "matches": int(n_a), "cv2_matches": int(n_b),
            "pnp_ok": bool(ok_a) and bool(ok_b), "pnp_inliers": int(len(inl_a)) if inl_a is not None else 0,
            "same_matches_as_cv2": bool(n_a == n_b), "note": "solvePnPRansac runs on the host CPU in both arms"}
    except Exception as e:  # cv2 missing: the line is context, not the metric
        out["frame_to_frame_1000x1000_match_plus_pnp"] = {"error": repr(e)}

    # the whole slam/tracking.py:96-128 step against a device-resident local map
    # (20k (keyframe, map point) edges, 2000 frame descriptors): projection + visibility + compaction +
    # cross-check match + gate + gather, one call.  CPU figure: the numpy restatement + cv2 match.
    sc = synth.local_map_scene(20000, 20000, 2000, seed=14)
    store = bb.MapStore(20000, engine=eng)
    store.update(np.arange(20000), sc["desc"], sc["pt3d"], sc["normal"])
    targs = (sc["des"], sc["kp"], sc["R"], sc["t"], sc["see_vector"], sc["edges"])
    for _ in range(3):
        r = store.track(*targs)
    t0 = time.perf_counter()
    for _ in range(reps):
        r = store.track(*targs)
    dt = (time.perf_counter() - t0) / reps
    out["local_map_track_20000edges_2000desc_fused"] = {
        "frames_per_s_e2e": 1.0 / dt, "ms_e2e": dt * 1e3, "visible": int(len(r.visible_edges)), "matches": int(len(r.inds)),
        "pairs_per_s_e2e": 2000 * len(r.visible_edges) / dt, "kernels_per_call": eng.launch_info()["kernels_launched"]}
    try:
        from oracle import localmap_oracle as lmo, cv2_reference as ref
        import math
        if ref.HAVE_CV2:
            def cpu():
                ok, pix = lmo.visible(sc["R"], sc["t"], sc["see_vector"], sc["pt3d"][sc["edges"]], sc["normal"][sc["edges"]],
                                      sc["fx"], sc["fy"], sc["cx"], sc["cy"], 640, 480, math.cos(math.pi / 3))
                vis = np.nonzero(ok)[0]
                feats = sc["desc"][sc["edges"][vis]]
                ms = [m for m in ref.matcher(True).match(sc["des"], feats) if m.distance <= 30]
                return len(ms)
            n_cpu = cpu()
            t0 = time.perf_counter()
            for _ in range(3):
                cpu()
            dtc = (time.perf_counter() - t0) / 3
            out["local_map_track_20000edges_2000desc_fused"].update(
                {"cpu_ms": dtc * 1e3, "cpu_kind": "numpy restatement of the projection loop (vectorised; the reference loops in "
                                                  "Python over g2o calls) + cv2 crossCheck match + gate", "cpu_matches": int(n_cpu)})
    except Exception as e:  # the CPU figure is optional context
        out["local_map_track_20000edges_2000desc_fused"]["cpu_error"] = repr(e)
    # MapPoint.add_observation's descriptor choice (slam/nodes.py:146-153) for the 2000 map points a new
    # keyframe touches, 10 stored observations each; CPU figure: the reference's literal double loop on 50 of them
    try:
        rngr = np.random.default_rng(17)
        base = rngr.integers(0, 256, (2000, 1, 32), dtype=np.uint8)
        obs = base ^ np.packbits(rngr.random((2000, 10, 256)) < 0.06, axis=2)
        cnts = np.full(2000, 10, np.int32)
        for _ in range(3):
            sel = bb.select_representative(obs, cnts, engine=eng)
        t0 = time.perf_counter()
        for _ in range(reps):
            sel = bb.select_representative(obs, cnts, engine=eng)
        dts = (time.perf_counter() - t0) / reps
        from oracle import representative_oracle as ro
        t0 = time.perf_counter()
        want = ro.select_batch(obs[:50], cnts[:50])
        dtc = (time.perf_counter() - t0) / 50 * 2000
        out["representative_descriptor_2000pts_10obs"] = {"ms_e2e": dts * 1e3, "points_per_s_e2e": 2000 / dts,
                                                          "cpu_ms_literal_loop_scaled": dtc * 1e3,
                                                          "agrees_with_reference_loop": bool(np.array_equal(sel[:50], want))}
    except Exception as e:
        out["representative_descriptor_2000pts_10obs"] = {"error": repr(e)}
    return out


def _masked_lists(torch, m, count, n_desc):
    """Match lists [.., 3, P*n_desc] with everything past each problem's count zeroed (only the filled prefix of a
    problem's slice is defined), so two tables can be compared bit for bit."""
    lead = m.shape[:-2]
    P = count.shape[-1]
    v = m.reshape(*lead, 3, P, n_desc)
    keep = torch.arange(n_desc, device=m.device).view(*([1] * len(lead)), 1, 1, n_desc) < count.reshape(*lead, 1, P, 1)
    return torch.where(keep, v, torch.zeros_like(v)).reshape(m.shape)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true",
                    help="skip the host-path leg (for ncu runs: its kernel waits on uploads a profiler replay cannot reproduce)")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the weak-scaling extra (256 pairs per rank)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    # stdout carries exactly ONE JSON line: anything a library prints there while we run (NCCL writes its
    # version banner to stdout) goes to stderr instead; the real stdout is restored for the final print
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    if args.impl == "reference":
        reference_arm(args, emit)
        return

    import zlib
    import torch
    import boslam_b200 as bb
    import boslam_b200.synth as synth
    from boslam_b200 import _ffi
    from boslam_b200.distributed import partition_pairs
    from boslam_b200.engine import PinnedBuffer, make_problems

    world = _env_int("WORLD_SIZE", 1)
    rank = _env_int("RANK", 0)
    local_rank = _env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    barrier = (lambda: dist.barrier()) if dist else (lambda: None)

    def max_over_ranks(x):
        if not dist:
            return float(x)
        tms = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        return float(tms.item())

    eng = bb.Engine(local_rank)
    dev = torch.device("cuda", local_rank)
    tune = {}
    if os.environ.get("BFM_TUNE"):   # diagnostics: e.g. BFM_TUNE=feeders=-1,feed_rows=16384 (A/B of the upload modes)
        tune = {k: int(v) for k, v in (kv.split("=") for kv in os.environ["BFM_TUNE"].split(","))}
        eng.set_tuning(**tune)
    # -- the workload: ONE list of 256 pairs per step; this rank's contiguous block of it ------
    blk0, blk1 = partition_pairs([N_DESC * N_DESC] * N_PAIRS, world)[rank]
    P_loc = blk1 - blk0
    tab = make_problems([N_DESC] * P_loc, [N_DESC] * P_loc)
    n_out = P_loc * N_DESC
    pairs_per_step = N_PAIRS * N_DESC * N_DESC             # the whole list, all ranks together
    set_bytes = 2 * n_out * 32
    n_sets = max(6, -(-130_000_000 // max(set_bytes, 1)))   # rotating input sets: > 126 MB L2 in total per rank

    # -- inputs.  Set 0 is this rank's block of the SAME global list for every world size (seed 1000), so the
    #    gathered tables can be compared across N (tables_crc32); the other sets are per-(set, rank) draws.  Pinned on
    #    the host + resident in HBM; the resident copies are uploaded FROM the pinned buffers, so every pinned page
    #    has been read by the device once before any timing --
    pinned, dev_sets = [], []
    for s_ in range(n_sets):
        if s_ == 0:
            gq, gt = synth.keyframe_pair_batch(N_PAIRS, N_DESC, seed=1000)
            q, t = gq[blk0 * N_DESC:blk1 * N_DESC], gt[blk0 * N_DESC:blk1 * N_DESC]
        else:
            q, t = synth.keyframe_pair_batch(P_loc, N_DESC, seed=1000 + 97 * s_ + 7919 * rank)
        pq, pt = PinnedBuffer(q.shape), PinnedBuffer(t.shape)
        pq.array[...] = q
        pt.array[...] = t
        pinned.append((pq, pt))
        dev_sets.append((torch.from_numpy(pq.array).to(dev), torch.from_numpy(pt.array).to(dev)))
    out = {"m": torch.empty((3, n_out), dtype=torch.int32, device=dev),
           "count": torch.zeros(P_loc, dtype=torch.int32, device=dev)}
    # the path's one exchange: every rank ends up with every rank's match tables.
    # Default: fused into the matching kernel's epilogue (stores to NVLink peer memory + one barrier);
    # BFM_GATHER=nccl selects the plain NCCL all_gather for comparison.
    fused = None
    gather_mode = os.environ.get("BFM_GATHER", "fused") if dist else "none"
    gathered_m = gathered_c = None
    if dist and gather_mode == "fused":
        try:
            from boslam_b200.distributed import FusedGather
            fused = FusedGather(n_out, P_loc, k=2)
        except Exception as e:  # symmetric memory unavailable on this box: say so and use NCCL
            print(f"[bench] fused gather unavailable ({type(e).__name__}: {e}); using NCCL all_gather", file=sys.stderr)
            gather_mode = "nccl"
    if dist and gather_mode == "nccl":
        gathered_m = torch.empty((world * 3, n_out), dtype=torch.int32, device=dev)
        gathered_c = torch.empty(world * P_loc, dtype=torch.int32, device=dev)

    def device_step(i):
        q, t = dev_sets[i % n_sets]
        if fused is not None:
            fused.run(eng, q, t, tab, k=2, ratio=RATIO)
            fused.barrier()
            return
        eng.match_batched_device(q, t, tab, k=2, ratio=RATIO, out=out)
        if dist and gather_mode != "none":   # "none": diagnostic only (no exchange: not a valid multi-GPU number)
            dist.all_gather_into_tensor(gathered_m, out["m"])
            dist.all_gather_into_tensor(gathered_c, out["count"])

    # measured POPC issue peak: the roofline denominator (not in MEASURED_PEAKS.json)
    popc = _ffi.microbench(local_rank, 4000, tests=("popc",))["popc"]

    sampler = ClockSampler(local_rank)
    for i in range(args.warmup):
        device_step(i)
    if fused is not None:
        fused.wait()
    launches0 = eng.kernel_launch_count()
    sampler.start()
    ms = time_device_loop(torch, device_step, args.steps, barrier, finish=(fused.wait if fused is not None else None))
    clocks = sampler.stop()
    launches = eng.kernel_launch_count() - launches0
    info = eng.launch_info()
    ms = max_over_ranks(ms)
    value = pairs_per_step * args.steps / (ms * 1e-3)

    # -- the same loop with every step's exchange completed before the next step starts (no overlap of barrier n with
    #    kernel n+1): the latency form of the step --
    sync_ms = None
    if fused is not None:
        def sync_step(i):
            device_step(i)
            fused.wait()
        sync_ms = max_over_ranks(time_device_loop(torch, sync_step, args.steps, barrier)) / args.steps

    # -- verification of the exchange, outside the timed loops: bit-equality of every rank's symmetric table (match
    #    lists, counts AND the dense knn tables) with an NCCL all_gather of the tables each rank computes locally --
    verify = {}
    q0, t0_ = dev_sets[0]
    loc = eng.match_batched_device(q0, t0_, tab, k=2, ratio=RATIO, want_knn=True)
    torch.cuda.synchronize()
    loc_m = _masked_lists(torch, loc["m"][:, :n_out], loc["count"][:P_loc], N_DESC)
    if dist:
        all_m = torch.empty((world, 3, n_out), dtype=torch.int32, device=dev)
        all_c = torch.empty((world, P_loc), dtype=torch.int32, device=dev)
        all_ki = torch.empty((world, n_out, 2), dtype=torch.int32, device=dev)
        all_kd = torch.empty((world, n_out, 2), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(all_m, loc_m.contiguous())
        dist.all_gather_into_tensor(all_c, loc["count"][:P_loc].contiguous())
        dist.all_gather_into_tensor(all_ki, loc["knn_idx"][:n_out].contiguous())
        dist.all_gather_into_tensor(all_kd, loc["knn_dist"][:n_out].contiguous())
    else:
        all_m, all_c = loc_m[None], loc["count"][None, :P_loc]
        all_ki, all_kd = loc["knn_idx"][None, :n_out], loc["knn_dist"][None, :n_out]
    if fused is not None:
        from boslam_b200.distributed import FusedGather
        fv = FusedGather(n_out, P_loc, k=2, want_knn=True)
        ok = True
        for rep in range(4):                     # four steps: every slot is reused at least once
            fv.run(eng, q0, t0_, tab, k=2, ratio=RATIO)
            fv.barrier()
            fv.wait()
            torch.cuda.synchronize()
            tb = fv.tables()
            ok = ok and torch.equal(tb["count"], all_c) and torch.equal(tb["knn_idx"], all_ki) and \
                torch.equal(tb["knn_dist"], all_kd) and torch.equal(_masked_lists(torch, tb["m"], tb["count"], N_DESC), all_m)
        # and the timed configuration's own table (match lists only), as left by one more step
        device_step(0)
        fused.wait()
        torch.cuda.synchronize()
        tb = fused.tables()
        ok = ok and torch.equal(tb["count"], all_c) and torch.equal(_masked_lists(torch, tb["m"], tb["count"], N_DESC), all_m)
        ok = ok and bool((all_c.sum(dim=1) > 0).all())
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        verify["gather_verified_full"] = bool(flag.item())
        verify["gather_verified"] = verify["gather_verified_full"]
        if not verify["gather_verified_full"]:
            raise SystemExit("fused gather verification failed: a rank's symmetric table differs from the NCCL gather of the local tables")
    # one checksum of the WHOLE gathered result of input set 0: identical for every world size (W = 1, 2, 4, 8
    # must be byte-identical)
    # (the tables are laid out [rank][...]; the match lists [rank][3][rows] are turned into [3][all rows], the order a single
    # GPU produces, so that the bytes - and the checksum - do not depend on how many ranks there are)
    crc = 0
    for a in (all_ki, all_kd, all_c, all_m.permute(1, 0, 2).contiguous()):
        crc = zlib.crc32(a.cpu().numpy().tobytes(), crc)
    verify["tables_crc32"] = int(crc)
    verify["matches_set0"] = int(all_c.sum().item())
    if dist:
        # ShardedMatcher (the host-level binding of 8(e)) through the real CUDA engine: sharded == unsharded
        from boslam_b200.distributed import ShardedMatcher
        rng = np.random.default_rng(77)
        sq, st_ = [], []
        for p in range(2 * world + 3):
            a, b, _ = synth.correlated(int(rng.integers(40, 400)), int(rng.integers(40, 500)), 500 + p)
            sq.append(a)
            st_.append(b)
        gi, gd = ShardedMatcher(engine=eng).knn_pairs(sq, st_, k=2)
        ok = True
        for p in range(len(sq)):
            ii, dd = eng.knn(sq[p], st_[p], 2)
            ok = ok and np.array_equal(gi[p], ii) and np.array_equal(gd[p], dd)
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        verify["sharded_matcher_verified"] = bool(flag.item())
        if not verify["sharded_matcher_verified"]:
            raise SystemExit("ShardedMatcher.knn_pairs (CUDA engine) differs from the unsharded engine result")

    # -- roofline: the step IS one launch of the scan kernel, so its average duration over the timed region is
    #    ms_per_step (CUDA events on the launching stream, launch gaps included: an upper bound on the kernel time);
    #    a second loop with the library's own per-launch events gives the isolated figure --
    eng.set_tuning(timing=1)
    scan_ms = []
    for i in range(args.steps):
        q, t = dev_sets[i % n_sets]
        eng.match_batched_device(q, t, tab, k=2, ratio=RATIO, out=out)
        scan_ms.append(eng.launch_info()["scan_ms"])
    eng.set_tuning(timing=0)
    scan_iso = max_over_ranks(float(np.mean(scan_ms)))
    kernel_ms = ms / args.steps if world == 1 else scan_iso
    pairs_per_launch = P_loc * N_DESC * N_DESC
    achieved_popc = pairs_per_launch * 8 / (kernel_ms * 1e-3)
    hbm_bytes = 32 * 2 * n_out + 8 * n_out  # descriptors in + packed row state out
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    from_file = {}
    try:   # figures that need a profiler (ncu --set full of this command) are READ FROM A FILE and labelled as such
        from_file = json.load(open(os.path.join(ROOT, "profiles", "ncu_scan_summary.json")))
    except Exception:
        pass
    popc_issued = 4 if info["popc_mode"] in (4, 40) else (5 if info["popc_mode"] in (5, 50) else info["popc_mode"])
    roofline = {
        "bound": "int_popc", "kernel": "bfm_scan_static_kernel" if info["scan_grid"] == info["segments"] else "bfm_scan_persistent_kernel", "achieved": achieved_popc / 1e9, "peak": popc["ops_per_s"] / 1e9,
        "unit": "GPOPC/s", "frac": achieved_popc / popc["ops_per_s"],
        "traffic": from_file.get("dram_bytes_per_launch") if world == 1 else None,
        "traffic_source": from_file.get("source") if world == 1 else None,
        "algorithmic": f"8 POPC per descriptor pair x {pairs_per_launch:.4g} pairs per launch",
        "popc_issued_per_pair": popc_issued,
        # the same launch against what the kernel really issues (carry-save tree: 4 POPC per pair, not 8):
        # how close the XU pipe is to saturation, measured live
        "frac_issued": achieved_popc * popc_issued / 8 / popc["ops_per_s"],
        "popc_mode": info["popc_mode"],
        "kernel_ms": kernel_ms, "kernel_ms_how": ("ms_per_step of the timed region (one launch per step)" if world == 1 else
                                                    "library events around each launch, separate loop (the step also holds the exchange)"),
        "kernel_ms_isolated": scan_iso,
        "from_file": from_file or None,
        "peak_source": "measured in this run: bfm_microbench POPC probe (16 POPC/clk/SM x 148 SMs x SM clock)",
        "popc_per_clk_per_sm": popc["ops_per_clk_per_sm"],
        "hbm": {"achieved": hbm_bytes / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": hbm_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": hbm_bytes,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
    }
    tensor_form = info["popc_mode"] == 0
    if tensor_form:
        # Tensor form (boslam_b200/csrc/bfm_tensor.cuh): a step is three launches - expansion to s8, the tcgen05 scan, the
        # tile-parallel finalize.  The dominant kernel is the scan; its duration comes from the library's events around
        # that launch alone (CUDA events on the launching stream, every step of a separate loop), its share of the step
        # is reported next to it.  Algorithmic work: 256 multiply-adds per pair = 512 ops.
        # (the 20-step loop lasts a few milliseconds: the burst figure is the honest denominator; `sustained` carries
        # the long-loop throughput next to the sustained peak)
        int8_peak = 2.0 * peaks.get("bf16_tflops", 2250.0 * 0.62)
        int8_peak_sustained = 2.0 * peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 2250.0 * 0.62))
        tops = pairs_per_launch * 512 / (scan_iso * 1e-3) / 1e12
        alu_pairs = 64 * 148 * clocks.get("sm_mhz", 1965.0) * 1e6 / 1.25 if isinstance(clocks, dict) and clocks.get("sm_mhz") else None
        roofline = {
            "bound": "tensor", "kernel": "bfm_tc::scan_kernel (tcgen05.mma kind::i8, TMEM accumulators)",
            "achieved": tops, "peak": int8_peak, "unit": "TOP/s", "frac": tops / int8_peak,
            "traffic": from_file.get("tensor_dram_bytes_per_launch") if world == 1 else None,
            "traffic_source": from_file.get("tensor_source") if world == 1 else None,
            "algorithmic": f"512 s8 ops (256 multiply-adds) per descriptor pair x {pairs_per_launch:.4g} pairs per launch",
            "peak_source": ("2 x the dense bf16 BURST figure of MEASURED_PEAKS.json (cuBLAS, best of 10; the kernel is timed in a loop of a few "
                            "milliseconds); B200's s8 tensor rate is twice its bf16 rate, the file holds no s8 measurement" if "bf16_tflops" in peaks else "fallback"),
            "peak_sustained": int8_peak_sustained,
            "kernel_ms": scan_iso, "kernel_ms_how": "library events around the scan launch alone, every step of a separate loop",
            "step_ms": ms / args.steps, "kernel_share_of_step": scan_iso / (ms / args.steps) if world == 1 else None,
            # the other pipe the kernel leans on: the epilogue turns every distance into a 16-bit key (two to a register) and
            # keeps the two smallest per row, 1 IMAD + 1.25 VIMNMX.U16x2 per pair and thread; the ALU pipe issues 64 lanes
            # per clock and SM (profiles/r02_tensor_scan.md: MMA alone 205 us, epilogue alone 168 us, both 228 us)
            "epilogue_alu": ({"pairs_per_s_bound": alu_pairs, "frac": pairs_per_launch / (scan_iso * 1e-3) / alu_pairs,
                              "how": "1.25 min/max per pair on the 64-lane ALU pipe x 148 SMs x SM clock under load"} if alu_pairs else None),
            "popc_form": {"peak_pairs_per_s": popc["ops_per_s"] / 4.0, "note": "bound of the POPC kernel this form replaces (4 POPC per pair on the XU pipe)"},
            "from_file": from_file or None,
            "hbm": {"achieved": (hbm_bytes + 2 * 256 * 2 * n_out) / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": (hbm_bytes + 2 * 256 * 2 * n_out) / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                    "algorithmic_bytes_per_launch": hbm_bytes + 2 * 256 * 2 * n_out,
                    "note": "descriptors in + expanded planes written and read once + row state",
                    "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
        }

    # -- e2e: numpy in -> numpy out through the public API, pinned host inputs.  At N > 1 the same call also delivers
    #    the exchange: its epilogue writes this rank's lists into every rank's table, then the barrier --
    host_out = bb.HostBatchBuffers(n_out, P_loc, k=2)  # pinned result arrays, reused every step
    plan = eng.plan_batch(tab, k=2, ratio=RATIO)       # the table and options are validated once, outside the loop
    fe = None
    if fused is not None:
        from boslam_b200.distributed import FusedGather
        fe = FusedGather(n_out, P_loc, k=2)

    def host_step(i):
        pq, pt = pinned[i % n_sets]
        if fe is not None:
            r = fe.run_host(plan, pq.array, pt.array, host_out)
            fe.barrier()
            return r
        return plan.run(pq.array, pt.array, host_out)

    e2e_steps = 0 if args.no_e2e else args.steps
    res = None
    for i in range(args.warmup if e2e_steps else 0):
        res = host_step(i)
    if fe is not None and e2e_steps:
        fe.wait()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        res = host_step(i)
    if fe is not None and e2e_steps:
        fe.wait()
        torch.cuda.synchronize()
    e2e_s = max(time.perf_counter() - t0, 1e-9)
    barrier()
    e2e_s = max_over_ranks(e2e_s)
    e2e = None
    if e2e_steps:
        n_match = int(res.counts.sum())
        e2e_ok = None
        if fe is not None:   # the e2e leg's exchange delivered the same tables (last step used input set (steps-1) % n_sets)
            fe.wait()
            torch.cuda.synchronize()
            q_, t_ = dev_sets[(e2e_steps - 1) % n_sets]
            l2 = eng.match_batched_device(q_, t_, tab, k=2, ratio=RATIO)
            lm = _masked_lists(torch, l2["m"][:, :n_out], l2["count"][:P_loc], N_DESC).contiguous()
            am = torch.empty((world, 3, n_out), dtype=torch.int32, device=dev)
            ac = torch.empty((world, P_loc), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(am, lm)
            dist.all_gather_into_tensor(ac, l2["count"][:P_loc].contiguous())
            tb = fe.tables()
            ok = torch.equal(tb["count"], ac) and torch.equal(_masked_lists(torch, tb["m"], tb["count"], N_DESC), am)
            ok = ok and np.array_equal(host_out.count[:P_loc], ac[rank].cpu().numpy())
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            e2e_ok = bool(flag.item())
            if not e2e_ok:
                raise SystemExit("e2e leg: the tables gathered by the host-path kernel differ from the NCCL gather")
        e2e = {"value": pairs_per_step * e2e_steps / e2e_s, "unit": "pairs/s",
               "h2d_bytes_per_step": int(2 * n_out * 32 + tab.nbytes), "d2h_bytes_per_step": n_match * 12 + P_loc * 4,
               "bytes_are": "per rank", "ms_per_step": e2e_s / e2e_steps * 1e3, "matches_last_step_this_rank": n_match,
               "h2d_gbs_per_gpu": 2 * n_out * 32 / (e2e_s / e2e_steps) / 1e9, "h2d_gbs_all_gpus": world * 2 * n_out * 32 / (e2e_s / e2e_steps) / 1e9,
               "tuning": tune or None,
               "copy_chunks": eng.launch_info().get("copy_chunks"), "exchange_in_timed_region": fe is not None,
               "exchange_verified": e2e_ok,
               "kernels_per_step": eng.launch_info().get("kernels_launched"),
               "how": ("numpy (pinned) in -> numpy (pinned) out through Engine.plan_batch(...).run: the copy engine uploads the step's "
                       "descriptors in copy_chunks chunks of whole problems and every chunk is matched by the tensor form (expansion, "
                       "tcgen05 scan, finalize: three launches) as soon as it has landed; results written by the finalize kernel "
                       "into pinned host memory" if eng.launch_info().get("popc_mode") == 0 else
                       "numpy (pinned) in -> numpy (pinned) out through Engine.plan_batch(...).run: ONE kernel launch per step whose "
                       "first CTAs stream the step's inputs from pinned host memory into HBM (copy_chunks = feed rounds) while the "
                       "others match; results written by the kernel into pinned host memory") +
                      (" AND, by the same epilogue, into every rank's symmetric table (NVSwitch multicast / peer stores), then "
                       "the symmetric-memory barrier (overlapping the next step's kernel; the last one is waited for inside the "
                       "timed region)" if fe is not None else "") + "; one stream sync per step"}

    # -- e2e, resident form: the descriptors live in a KeyframeBank since keyframe creation (what boslam's map does:
    #    KeyFrame.des is created once, slam/covisibility_graph.py:119-138); a step sends only this rank's block of the
    #    pair list (keyframe ids) and gets the match lists back in pinned host memory - plus, at N > 1, the exchange --
    e2e_res = None
    if e2e_steps:
        bank = bb.KeyframeBank(capacity_rows=n_sets * 2 * n_out + 64, engine=eng)
        pair_sets = []
        for s_ in range(n_sets):
            pq, pt = pinned[s_]
            ids = np.zeros((P_loc, 2), np.int64)
            for p_ in range(P_loc):
                kq, kt = (s_ * 2 * P_loc) + p_, (s_ * 2 * P_loc) + P_loc + p_
                bank.add(kq, pq.array[p_ * N_DESC:(p_ + 1) * N_DESC])
                bank.add(kt, pt.array[p_ * N_DESC:(p_ + 1) * N_DESC])
                ids[p_] = (kq, kt)
            pair_sets.append(ids)
        fb = None
        if fused is not None:
            from boslam_b200.distributed import FusedGather
            fb = FusedGather(n_out, P_loc, k=2)

        def bank_step(i):
            if fb is not None:
                r = fb.run_bank(bank, pair_sets[i % n_sets], k=2, ratio=RATIO)
                fb.barrier()
                return r
            return bank.match_pairs(pair_sets[i % n_sets], k=2, ratio=RATIO, copy=False)

        for i in range(args.warmup):
            rb = bank_step(i)
        if fb is not None:
            fb.wait()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            rb = bank_step(i)
        if fb is not None:
            fb.wait()
            torch.cuda.synchronize()
        res_s = max_over_ranks(max(time.perf_counter() - t0, 1e-9))
        barrier()
        # same bits as the host-array leg's last step (same input set)
        same = bool(np.array_equal(rb.counts, res.counts[:P_loc]) and int(rb.counts.sum()) == n_match)
        if same:
            for p_ in (0, P_loc // 2, P_loc - 1):
                same = same and all(np.array_equal(a, b) for a, b in zip(rb[p_], res[p_]))
        if not same:
            raise SystemExit("e2e resident leg: the bank's match lists differ from the host-array leg's")
        e2e_res = {"value": pairs_per_step * e2e_steps / res_s, "unit": "pairs/s", "ms_per_step": res_s / e2e_steps * 1e3,
                   "h2d_bytes_per_step": int(P_loc * 24), "d2h_bytes_per_step": n_match * 12 + P_loc * 4, "bytes_are": "per rank",
                   "exchange_in_timed_region": fb is not None, "same_lists_as_e2e": same,
                   "how": "KeyframeBank.match_pairs: descriptors resident in HBM since keyframe creation, a step uploads this "
                          "rank's block of the pair list as a problem table, ONE kernel launch matches it and writes the match "
                          "lists into pinned host memory" + (" and into every rank's symmetric table, then the barrier" if fb is not None else "") +
                          "; one stream sync per step"}

    line = {
        "metric": "hamming_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "s8" if info["popc_mode"] == 0 else "u32", "data": "synthetic",   # s8 operands, s32 accumulators
        "config": bench_config(),
        "run": {"pairs_this_rank": P_loc, "input_sets": n_sets, "input_mb_per_rank": n_sets * set_bytes / 1e6,
                "parallelism": (f"pair list split x{world}, match tables gathered by the kernel epilogue over NVLink "
                                f"({'NVSwitch multicast stores' if fused.multicast_ptr else 'peer stores'}) + symmetric-memory barrier "
                                "per step (side stream, overlapping the next step's kernel; all inside the timed region)"
                                if fused is not None else f"pair list split x{world}, NCCL all_gather of match tables"
                                if gather_mode == "nccl" else f"DIAGNOSTIC: pair list split x{world} with NO exchange")
                if world > 1 else "single GPU"},
        "clocks": clocks, "e2e": e2e, "e2e_resident": e2e_res, "gpu_launches": int(launches), "roofline": roofline,
        "launch": dict({k: info[k] for k in ("scan_grid", "scan_block", "queries_per_thread", "popc_mode",
                                             "train_rows_per_segment", "segments", "kernels_launched")},
                       form="tensor: s8 expansion + tcgen05 scan (one persistent CTA per SM) + tile-parallel finalize" if info["popc_mode"] == 0 else
                            "static: one work item per CTA" if info["scan_grid"] == info["segments"] else
                            "persistent: one wave of CTAs drawing work items from a ticket counter"),
        "frames_per_s": N_PAIRS * args.steps / (ms * 1e-3),
        "verify": verify,
    }
    if sync_ms is not None:
        line["step_latency_ms_exchange_completed"] = sync_ms
    line.update({k: v for k, v in verify.items() if k.startswith("gather_verified")})

    # -- weak-scaling extra (round 1's headline form): every rank its own 256 pairs, exchange fused ----------------
    if dist and fused is not None and not args.no_weak:
        from boslam_b200.distributed import FusedGather
        wq, wt = synth.keyframe_pair_batch(N_PAIRS, N_DESC, seed=5000 + rank)
        wsets = [(torch.from_numpy(wq).to(dev), torch.from_numpy(wt).to(dev))]
        for s_ in range(1, 4):   # 4 x 32.8 MB = 131 MB > L2
            a, b = synth.keyframe_pair_batch(N_PAIRS, N_DESC, seed=5000 + 97 * s_ + rank)
            wsets.append((torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)))
        wtab = make_problems([N_DESC] * N_PAIRS, [N_DESC] * N_PAIRS)
        fw = FusedGather(N_PAIRS * N_DESC, N_PAIRS, k=2)

        def weak_step(i):
            a, b = wsets[i % len(wsets)]
            fw.run(eng, a, b, wtab, k=2, ratio=RATIO)
            fw.barrier()
        for i in range(args.warmup):
            weak_step(i)
        fw.wait()
        wms = max_over_ranks(time_device_loop(torch, weak_step, args.steps, barrier, finish=fw.wait))
        line["weak"] = {"value": pairs_per_step * world * args.steps / (wms * 1e-3), "unit": "pairs/s", "ms_per_step": wms / args.steps,
                        "pairs_per_gpu": N_PAIRS, "note": "every rank matches its own 256 pairs; exchange fused, barrier in the timed region"}

    if rank == 0 and world == 1 and not args.no_sustained:
        # >= 2 s of back-to-back steps with the clock sampler running, beside the short burst above
        n_sus = max(200, int(2200.0 / max(ms / args.steps, 1e-3)))
        s2 = ClockSampler(local_rank)
        s2.start()
        sms = time_device_loop(torch, device_step, n_sus, barrier)
        line["sustained"] = {"value": pairs_per_step * n_sus / (sms * 1e-3), "unit": "pairs/s", "ms_per_step": sms / n_sus,
                             "steps": n_sus, "seconds": sms * 1e-3, "clocks": s2.stop()}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cv2_reference as ref
        q, t = pinned[0][0].array, pinned[0][1].array
        if ref.HAVE_CV2:
            m = ref.matcher(False)
            cv2_step(m, q, t, tab[:8], RATIO)
            reps, t0 = 0, time.perf_counter()
            while reps < 3 or time.perf_counter() - t0 < 10.0:
                n_good = cv2_step(m, q, t, tab, RATIO)
                reps += 1
                if reps >= 8:
                    break
            dt = (time.perf_counter() - t0) / reps
            line["cpu_baseline"] = {"value": pairs_per_step / dt, "unit": "pairs/s", "cores": ref.threads(), "kind": "reference",
                                    "sample": f"the full {N_PAIRS}-pair batch x {reps} reps, cv2 {ref.version()} knnMatch(k=2) + Python "
                                              f"ratio test incl. DMatch construction, os.cpu_count()={os.cpu_count()}",
                                    "ms_per_step": dt * 1e3, "matches": int(n_good)}
            # the checker (outside every timed region): the engine's FULL tables of this batch against cv2's, entry by entry
            ci, cd = cv2_tables(m, q, t, tab)
            keep = (ci[:, 1] >= 0) & (cd[:, 0].astype(np.float64) < RATIO * cd[:, 1].astype(np.float64))
            gi, gd, gres = eng.match_batched(q, t, tab, k=2, ratio=RATIO, want_knn=True)
            same = np.array_equal(gi, ci) and np.array_equal(gd, cd) and int(gres.counts.sum()) == int(keep.sum()) == int(n_good)
            for p in range(N_PAIRS):
                rows = np.nonzero(keep[p * N_DESC:(p + 1) * N_DESC])[0]
                a, b, c = gres[p]
                same = same and np.array_equal(a, rows) and np.array_equal(b, ci[p * N_DESC + rows, 0]) and \
                    np.array_equal(c.astype(np.int32), cd[p * N_DESC + rows, 0])
            same = same and np.array_equal(all_ki[0].cpu().numpy(), ci) and np.array_equal(all_kd[0].cpu().numpy(), cd)
            line["cpu_baseline"]["tables_equal_cv2"] = bool(same)
            assert same, "cv2 and engine disagree on the headline batch (full knn tables + match lists)"
            try:  # also a 1-thread figure (8 of the 256 pairs)
                import cv2
                nthreads = cv2.getNumThreads()
                cv2.setNumThreads(1)
                cv2_step(m, q, t, tab[:2], RATIO)
                t0 = time.perf_counter()
                cv2_step(m, q, t, tab[:8], RATIO)
                dt1 = time.perf_counter() - t0
                cv2.setNumThreads(nthreads)
                line["cpu_baseline"]["value_1_thread"] = 8 * N_DESC * N_DESC / dt1
            except Exception:
                pass
        else:
            from oracle import c_oracle
            t0 = time.perf_counter()
            for p in range(8):
                c_oracle.knn(q[p * N_DESC:(p + 1) * N_DESC], t[p * N_DESC:(p + 1) * N_DESC], 2)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 8 * N_DESC * N_DESC / dt, "unit": "pairs/s", "cores": 1, "kind": "port",
                                    "sample": "8 of 256 pairs, plain-C oracle, 1 thread"}
    if rank == 0 and world == 1 and not args.no_extras:
        line["frames"] = extras(eng, torch, args.steps)

    if rank == 0:
        emit(line)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
