"""Multi-GPU check of the fused gather (kernel epilogue writes into NVLink peer memory).  Needs a box
with >= 2 GPUs; skipped otherwise (the host-side sharding logic is covered on CPU with gloo)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret, multicast="1"):
    os.environ["BFM_MULTICAST"] = multicast
    import torch
    import torch.distributed as dist
    import boslam_b200 as bb
    from boslam_b200 import synth
    from boslam_b200.distributed import FusedGather
    from oracle import c_oracle, hamming_oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        P, N = 6, 500
        eng = bb.Engine(rank)
        tab = bb.make_problems([N] * P, [N] * P)
        fg = FusedGather(P * N, P, k=2, want_knn=True)
        ok = True
        for step in range(7):                      # seven steps: every slot gets reused twice
            data = [synth.keyframe_pair_batch(P, N, seed=100 * step + r) for r in range(world)]
            q, t = data[rank]
            fg.run(eng, torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), tab, k=2, ratio=0.8)
            fg.barrier()
            fg.wait()
            torch.cuda.synchronize()
            tb = {k: v.cpu().numpy() for k, v in fg.tables().items()}
            for r in range(world):
                qr, tr = data[r]
                for p in range(P):
                    a, b = qr[p * N:(p + 1) * N], tr[p * N:(p + 1) * N]
                    oi, od = c_oracle.knn(a, b, 2)
                    want = orc.match(a, b, k=2, ratio=0.8)
                    c = int(tb["count"][r, p])
                    ok = ok and c == len(want[0])
                    ok = ok and np.array_equal(tb["m"][r, 0, p * N:p * N + c], want[0])
                    ok = ok and np.array_equal(tb["m"][r, 1, p * N:p * N + c], want[1])
                    ok = ok and np.array_equal(tb["knn_idx"][r, p * N:(p + 1) * N], oi)
                    ok = ok and np.array_equal(tb["knn_dist"][r, p * N:(p + 1) * N], od)
        # one problem on its own: its rows are finalized tile by tile by several CTAs of the one launch, which
        # write the same peer / multicast destinations
        N1 = 2100
        fg1 = FusedGather(N1, 1, k=2, want_knn=True)
        tab1 = bb.make_problems([N1], [N1])
        data = [synth.correlated(N1, N1, 900 + r)[:2] for r in range(world)]
        q, t = data[rank]
        for step in range(2):
            fg1.run(eng, torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), tab1, k=2, ratio=0.8)
            ok = ok and eng.launch_info()["kernels_launched"] == 1
            fg1.barrier()
            fg1.wait()
            torch.cuda.synchronize()
            tb = {k: v.cpu().numpy() for k, v in fg1.tables().items()}
            for r in range(world):
                a, b = data[r]
                oi, od = c_oracle.knn(a, b, 2)
                want = orc.match(a, b, k=2, ratio=0.8)
                c = int(tb["count"][r, 0])
                ok = ok and c == len(want[0])
                ok = ok and np.array_equal(tb["m"][r, 0, :c], want[0]) and np.array_equal(tb["m"][r, 1, :c], want[1])
                ok = ok and np.array_equal(tb["knn_idx"][r], oi) and np.array_equal(tb["knn_dist"][r], od)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("multicast", ["1", "0"])     # NVSwitch multicast stores (where available) / per-peer stores
def test_fused_gather_all_gpus(multicast):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    port = _free_port()
    world = min(_n_gpus(), 8)                          # every GPU of the box: 2, 4 or 8 ranks
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret, multicast)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(600) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert all(ret.get(r) is True for r in range(world)), dict(ret)
