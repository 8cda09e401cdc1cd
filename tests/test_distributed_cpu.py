"""Host-side sharding logic on CPU: world_size-2 gloo process group, oracle injected as the
compute callable (the product binding always uses the CUDA engine; this only tests partition +
gather layout)."""
import os
import socket

import numpy as np
import pytest

from boslam_b200 import synth
from boslam_b200.distributed import partition_pairs


def test_partition_equal_and_ragged():
    assert partition_pairs([4] * 8, 2) == [(0, 4), (4, 8)]
    assert partition_pairs([4] * 7, 4) == [(0, 2), (2, 4), (4, 6), (6, 7)]
    assert partition_pairs([], 2) == [(0, 0), (0, 0)]
    blocks = partition_pairs([100, 1, 1, 1, 1, 100], 2)
    assert blocks[0][0] == 0 and blocks[-1][1] == 6 and blocks[0][1] == blocks[1][0]
    costs = np.random.default_rng(0).integers(1, 1000, 50).tolist()
    for w in (2, 4, 8):
        bl = partition_pairs(costs, w)
        assert bl[0][0] == 0 and bl[-1][1] == 50
        assert all(bl[i][1] == bl[i + 1][0] for i in range(w - 1))
        loads = [sum(costs[b:e]) for b, e in bl]
        assert max(loads) <= sum(costs) / w + max(costs)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ragged, ret):
    import torch
    import torch.distributed as dist
    from oracle import hamming_oracle as orc
    from boslam_b200.distributed import ShardedMatcher
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        qs, ts = [], []
        for p in range(6):
            nq = int(rng.integers(5, 40)) if ragged else 24
            nt = int(rng.integers(5, 60)) if ragged else 30
            q, t, _ = synth.correlated(nq, nt, 200 + p)
            qs.append(q)
            ts.append(t)

        def compute(lq, lt, k):
            if not lq:
                z = np.zeros((0, k), np.int32)
                return torch.from_numpy(z), torch.from_numpy(z.copy())
            parts = [orc.knn(a, b, k) for a, b in zip(lq, lt)]
            return (torch.from_numpy(np.concatenate([p[0] for p in parts])),
                    torch.from_numpy(np.concatenate([p[1] for p in parts])))

        sm = ShardedMatcher(compute=compute)
        gi, gd = sm.knn_pairs(qs, ts, k=2)
        ok = True
        for p in range(6):
            oi, od = orc.knn(qs[p], ts[p], 2)
            ok = ok and np.array_equal(gi[p], oi) and np.array_equal(gd[p], od)
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ragged", [False, True])
def test_sharded_gather_world2_gloo(ragged):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ragged, ret)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret.get(0) is True and ret.get(1) is True


def test_symmetric_memory_guard_and_multicast_fallback(monkeypatch, capsys):
    """The fused gather rests on the private torch.distributed._symmetric_memory: a torch build without the entry
    points it needs must fail with a message that says so, and a handle without multicast_ptr must select the
    per-peer-store fallback (multicast_ptr == 0) instead of raising."""
    import sys
    import types
    from boslam_b200 import distributed as D
    symm = D._symmetric_memory()                       # this image's torch has it
    assert hasattr(symm, "empty") and hasattr(symm, "rendezvous")
    fake = types.ModuleType("torch.distributed._symmetric_memory")
    fake.empty = lambda *a, **k: None                  # rendezvous is gone
    monkeypatch.setitem(sys.modules, "torch.distributed._symmetric_memory", fake)
    import torch.distributed as td
    monkeypatch.setattr(td, "_symmetric_memory", fake, raising=False)
    with pytest.raises(RuntimeError, match="rendezvous"):
        D._symmetric_memory()

    class NoMulticast:
        pass

    class WithMulticast:
        multicast_ptr = 0x7F0000000000
    assert D._multicast_ptr(NoMulticast()) == 0 and "per-peer stores" in capsys.readouterr().err
    monkeypatch.delenv("BFM_MULTICAST", raising=False)
    assert D._multicast_ptr(WithMulticast()) == 0x7F0000000000
    monkeypatch.setenv("BFM_MULTICAST", "0")
    assert D._multicast_ptr(WithMulticast()) == 0
