"""Keyframe-pair batches sharded over the GPUs of one box.

Every (query keyframe, train keyframe) problem is independent, so rank r takes a contiguous
block of the pair list, runs its own engine on its own GPU and the only exchange is one
all-gather of the fixed-shape result tables (NCCL over NVLink).  The reductions are integer mins
over packed keys, so the gathered result is byte-identical for any world size.

The compute callable is injected so the host logic (partition, gather layout) can be exercised
on CPU with the gloo backend in tests; the product binding is :class:`ShardedMatcher`, which
always runs the CUDA engine.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np


def partition_pairs(costs: Sequence[int], world_size: int):
    """Contiguous block partition of the pair list into ``world_size`` shards of near-equal total
    cost (cost = Q*T of a pair).  Returns [(begin, end)] per rank; equal-cost pairs split evenly."""
    n = len(costs)
    c = np.asarray(costs, dtype=np.float64)
    if n == 0:
        return [(0, 0)] * world_size
    if np.all(c == c[0]):
        base, rem = divmod(n, world_size)
        out, b = [], 0
        for r in range(world_size):
            e = b + base + (1 if r < rem else 0)
            out.append((b, e))
            b = e
        return out
    cum = np.concatenate([[0.0], np.cumsum(c)])
    total = cum[-1]
    out, b = [], 0
    for r in range(world_size):
        if r == world_size - 1:
            e = n
        else:
            # the cut whose cumulative cost is closest to this rank's share of the total
            e = max(int(np.argmin(np.abs(cum - total * (r + 1) / world_size))), b)
        out.append((b, e))
        b = e
    return out


def gather_fixed(local, group=None):
    """All-gather a rank-local torch tensor whose shape is the same on every rank; returns the
    concatenation along dim 0 (rank order).  One ``all_gather_into_tensor`` call."""
    import torch
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    out = torch.empty((ws * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def gather_ragged(local, group=None):
    """All-gather along dim 0 of tensors whose dim-0 length differs per rank (ragged pair blocks):
    lengths are exchanged first, shards are padded to the maximum, gathered once and trimmed."""
    import torch
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    lens = torch.empty(ws, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(lens, n, group=group)
    lens = lens.cpu().tolist()
    m = max(lens) if lens else 0
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = gather_fixed(pad, group)
    parts = [out[r * m:r * m + lens[r]] for r in range(ws)]
    return torch.cat(parts, dim=0) if parts else out, lens


def _symmetric_memory():
    """``torch.distributed._symmetric_memory`` is a private torch module (present in torch 2.5 ... 2.11, the
    version this package was written against).  Import it here, check the three entry points the fused gather
    needs, and fail with a message that names the torch version instead of an AttributeError deep inside a step."""
    import torch
    try:
        import torch.distributed._symmetric_memory as symm
    except Exception as e:  # pragma: no cover - depends on the torch build
        raise RuntimeError(f"torch {torch.__version__} has no torch.distributed._symmetric_memory ({e}); the fused gather "
                           "needs it - use gather_fixed() / BFM_GATHER=nccl (plain NCCL all_gather) instead") from e
    missing = [n for n in ("empty", "rendezvous") if not hasattr(symm, n)]
    if missing:
        raise RuntimeError(f"torch {torch.__version__}: torch.distributed._symmetric_memory lacks {missing}; the fused "
                           "gather was written against torch 2.11 - use gather_fixed() / BFM_GATHER=nccl instead")
    return symm


def _multicast_ptr(hdl) -> int:
    """NVSwitch multicast address of a symmetric-memory handle, 0 when this torch build / box does not provide one
    (the epilogue then stores to every peer separately: measured 96.8 % instead of 99.9 % weak-scaling efficiency)."""
    import os
    import sys
    if not hasattr(hdl, "multicast_ptr"):
        print("[boslam_b200] symmetric-memory handle has no multicast_ptr in this torch build: per-peer stores", file=sys.stderr)
        return 0
    mc = int(hdl.multicast_ptr or 0)
    return mc if (mc and os.environ.get("BFM_MULTICAST", "1") != "0") else 0


class FusedGather:
    """Result tables of a sharded batch, gathered by the matching kernel itself.

    Every rank owns a symmetric buffer (``torch.distributed._symmetric_memory``: CUDA IPC / fabric
    handles exchanged once, peers mapped over NVLink) laid out as ``[slot][rank][table]``.  A step's
    kernel (:meth:`Engine.match_batched_device` with ``replicas``) writes this rank's match lists
    into slice ``rank`` of EVERY rank's buffer from its epilogue, so when all kernels have finished
    each rank holds the full table - there is no separate collective, only :meth:`barrier`.
    Three slots rotate between steps and the barrier runs on a side stream, so barrier n overlaps the
    kernel of step n + 1: consumers read step n between barrier n and barrier n + 1 (on the side stream or
    after ``wait``), and the kernel of step n + 3, which reuses the slot, first waits for barrier n + 1.

    Per-rank table (int32): ``m[3][n_out]`` (queryIdx, trainIdx, distance), ``count[P]`` and,
    with ``want_knn``, ``knn_idx[n_out][k]`` / ``knn_dist[n_out][k]``.
    """

    SLOTS = 3

    def __init__(self, n_out: int, n_problems: int, k: int = 1, want_knn: bool = False, group=None, device=None):
        import torch
        import torch.distributed as dist
        symm = _symmetric_memory()
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 8:
            raise ValueError("the fused gather writes to at most 8 destinations (one NVLink box)")
        self.n_out, self.P, self.k, self.want_knn = int(n_out), int(n_problems), int(k), bool(want_knn)
        a2 = lambda n: (n + 3) & ~3  # keep every sub-table 16-byte aligned
        self._o_m, self._o_c = 0, a2(3 * self.n_out)
        self._o_ki = self._o_c + a2(self.P)
        self._o_kd = self._o_ki + (a2(self.n_out * self.k) if want_knn else 0)
        self.table = self._o_kd + (a2(self.n_out * self.k) if want_knn else 0)
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.buf = symm.empty(self.SLOTS * self.world * self.table, dtype=torch.int32, device=dev)
        self.buf.fill_(-1)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self._ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        # NVSwitch multicast (NVLS): one multimem.st from the kernel epilogue lands in every rank's buffer
        self.multicast_ptr = _multicast_ptr(self.hdl)
        self.step = 0
        self._torch = torch
        self._side = torch.cuda.Stream(device=dev)
        self._barrier_done = {}          # step -> event of that step's barrier (the last few steps only)
        self._events = [(torch.cuda.Event(), torch.cuda.Event()) for _ in range(4 * self.SLOTS)]
        self._host_calls = {}            # (plan, host_out, slot) -> bound arguments of the host-path step
        self._dev_calls = {}             # (engine, problem table, slot, options) -> bound arguments of the device step
        self._bank_reps = {}             # slot -> destination list of the KeyframeBank step
        torch.cuda.synchronize(dev)
        self.hdl.barrier()
        torch.cuda.synchronize(dev)

    def _dest(self, base_ptr: int, slot: int, multicast: bool = False) -> dict:
        b = base_ptr + 4 * (slot * self.world + self.rank) * self.table
        d = {"m_query": b + 4 * self._o_m, "m_train": b + 4 * (self._o_m + self.n_out),
             "m_dist": b + 4 * (self._o_m + 2 * self.n_out), "count": b + 4 * self._o_c, "multicast": multicast}
        if self.want_knn:
            d["knn_idx"], d["knn_dist"] = b + 4 * self._o_ki, b + 4 * self._o_kd
        return d

    def destinations(self):
        """(own, peers) destination dicts of raw device pointers for the current step's slot."""
        slot = self.step % self.SLOTS
        if self.multicast_ptr:
            return self._dest(self.multicast_ptr, slot, multicast=True), []
        own = self._dest(self._ptrs[self.rank], slot)
        peers = [self._dest(self._ptrs[r], slot) for r in range(self.world) if r != self.rank]
        return own, peers

    def run(self, engine, q, t, problems, **kw):
        """One sharded step: match this rank's block, results land in every rank's table.  Queued on the
        current stream; before overwriting a slot it waits for the barrier that released its last readers."""
        n = self.step
        if n >= self.SLOTS:
            self._torch.cuda.current_stream().wait_event(self._barrier_done[n - self.SLOTS + 1])
        # steady state of a sharded loop: same engine, problem table and options every step, the slot's destinations
        # never change - everything but the two descriptor pointers is marshalled once per slot
        key = (id(engine), id(problems), n % self.SLOTS, tuple(sorted(kw.items())))
        hit = self._dev_calls.get(key)
        if hit is None:
            own, peers = self.destinations()
            hit = self._dev_calls[key] = (engine, problems, engine.bind_device_multi(problems, [own] + list(peers), self.want_knn, **kw))
            if len(self._dev_calls) > 64:
                self._dev_calls = {key: hit}
        engine.run_device_multi(q, t, hit[2])

    def run_host(self, plan, q, t, host_out):
        """The same step from HOST arrays (``plan`` = :meth:`Engine.plan_batch` of this rank's block; ``q`` / ``t``
        numpy, ideally pinned; ``host_out`` a :class:`HostBatchBuffers`): one kernel launch reads the inputs from host
        memory, matches, writes this rank's match lists into ``host_out`` AND into slice ``rank`` of every rank's
        table.  Synchronous (the host path always is); follow with :meth:`barrier`."""
        n = self.step
        if n >= self.SLOTS:   # the slot's last readers: the barrier that released them must have completed
            self._barrier_done[n - self.SLOTS + 1].synchronize()
        # the destination structs of a slot never change: built once per (plan, output buffers, slot)
        key = (id(plan), id(host_out), n % self.SLOTS)
        hit = self._host_calls.get(key)
        if hit is None:
            own, peers = self.destinations()
            hit = self._host_calls[key] = (plan, plan.bind_replicas(host_out, [own] + list(peers)))   # (keeps both alive: ids stay unique)
            if len(self._host_calls) > 64:
                self._host_calls = {key: hit}
        return plan.run_bound(q, t, hit[1])

    def run_bank(self, bank, pairs, **kw):
        """The same step over a :class:`boslam_b200.KeyframeBank` (descriptors resident since keyframe creation,
        ``pairs`` = this rank's block of (query id, train id)): one launch matches the block, writes the match lists
        into the bank's pinned host buffers AND into slice ``rank`` of every rank's table.  Synchronous; follow with
        :meth:`barrier`.  Returns the bank's :class:`BatchResult` (views of its pinned buffers)."""
        n = self.step
        if n >= self.SLOTS:
            self._barrier_done[n - self.SLOTS + 1].synchronize()
        slot = n % self.SLOTS
        reps = self._bank_reps.get(slot)
        if reps is None:   # one list object per slot, so the bank can recognise the configuration it has bound
            own, peers = self.destinations()
            reps = self._bank_reps[slot] = [own] + list(peers)
        return bank.match_pairs(pairs, replicas=reps, copy=False, **kw)

    def barrier(self):
        """All ranks' kernels of this step have finished (and their NVLink writes with them): the slot is
        complete on every rank once this barrier has run.  It is queued on a side stream behind this step's
        kernel, so the next step's kernel does not wait for it; :meth:`wait` orders the current stream
        after it.  Advances to the next slot."""
        torch = self._torch
        # (events come from a ring - creating two per step and entering a stream context cost ~20 us of the ~45 us
        # this call took; a ring entry is reused 4 * SLOTS steps later, long after its last waiter)
        ev, done = self._events[self.step % len(self._events)]
        cur = torch.cuda.current_stream()
        ev.record(cur)
        self._side.wait_event(ev)
        torch.cuda.set_stream(self._side)
        try:
            self.hdl.barrier()
        finally:
            torch.cuda.set_stream(cur)
        done.record(self._side)
        self._barrier_done[self.step] = done
        self._barrier_done.pop(self.step - 2 * self.SLOTS, None)
        self.step += 1

    def wait(self, step=None):
        """Order the current stream after the barrier of ``step`` (default: the last one): after this,
        :meth:`tables` of that step may be read on the current stream."""
        step = self.step - 1 if step is None else step
        self._torch.cuda.current_stream().wait_event(self._barrier_done[step])

    def tables(self, step=None):
        """Views of the gathered tables of ``step`` (default: the last completed one):
        m int32[world, 3, n_out], count int32[world, P] (+ knn_idx / knn_dist [world, n_out, k])."""
        step = self.step - 1 if step is None else step
        slot = step % self.SLOTS
        v = self.buf[slot * self.world * self.table:(slot + 1) * self.world * self.table].view(self.world, self.table)
        out = {"m": v[:, self._o_m:self._o_m + 3 * self.n_out].view(self.world, 3, self.n_out),
               "count": v[:, self._o_c:self._o_c + self.P]}
        if self.want_knn:
            out["knn_idx"] = v[:, self._o_ki:self._o_ki + self.n_out * self.k].view(self.world, self.n_out, self.k)
            out["knn_dist"] = v[:, self._o_kd:self._o_kd + self.n_out * self.k].view(self.world, self.n_out, self.k)
        return out


class ShardedMatcher:
    """Batched k-NN over a global list of keyframe pairs, sharded across the process group.

    ``knn_pairs(queries, trains, k)``: every rank passes the same global pair list (host arrays);
    each rank uploads and matches only its block, then the dense ``[rows, k]`` index / distance
    tables are all-gathered so every rank holds the full result.
    """

    def __init__(self, engine=None, device=None, group=None, compute=None):
        import torch.distributed as dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._compute = compute
        if compute is None:
            from .engine import Engine
            import torch
            self.device = torch.cuda.current_device() if device is None else device
            self.engine = engine if engine is not None else Engine(self.device)
            self._compute = self._engine_compute
            self._tdev = torch.device("cuda", self.device)
        else:
            import torch
            self._tdev = torch.device("cpu")

    def _engine_compute(self, qs, ts, k):
        """Local block through the CUDA engine -> (idx, dist) int32 [rows, k] torch tensors on the GPU."""
        import torch
        from .engine import make_problems
        if not qs:
            z = torch.zeros((0, k), dtype=torch.int32, device=self._tdev)
            return z, z.clone()
        qp = torch.from_numpy(np.concatenate(qs)).to(self._tdev)
        tp = torch.from_numpy(np.concatenate(ts)).to(self._tdev)
        tab = make_problems([len(a) for a in qs], [len(a) for a in ts])
        out = self.engine.match_batched_device(qp, tp, tab, k=k, want_knn=True)
        n = int(tab[:, 1].sum())
        return out["knn_idx"][:n], out["knn_dist"][:n]

    def knn_pairs(self, queries, trains, k: int = 2) -> Tuple[list, list]:
        import torch
        costs = [len(q) * len(t) for q, t in zip(queries, trains)]
        blocks = partition_pairs(costs, self.world)
        b, e = blocks[self.rank]
        idx, dist_ = self._compute(list(queries[b:e]), list(trains[b:e]), k)
        idx = torch.as_tensor(idx, device=self._tdev)
        dist_ = torch.as_tensor(dist_, device=self._tdev)
        both = torch.stack([idx, dist_], dim=1)  # [rows, 2, k]
        rows = [sum(len(q) for q in queries[bb:ee]) for bb, ee in blocks]
        if len(set(rows)) == 1:
            full = gather_fixed(both, self.group)
        else:
            full, _ = gather_ragged(both, self.group)
        full = full.cpu().numpy()
        out_i, out_d, o = [], [], 0
        for q in queries:
            out_i.append(full[o:o + len(q), 0])
            out_d.append(full[o:o + len(q), 1])
            o += len(q)
        return out_i, out_d
