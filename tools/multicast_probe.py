import os, torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
buf = symm.empty(1 << 20, dtype=torch.int32, device=torch.device("cuda", rank))
hdl = symm.rendezvous(buf, dist.group.WORLD)
if rank == 0:
    print("backend", symm.get_backend(torch.device("cuda", rank)) if hasattr(symm, "get_backend") else "?")
    print("has_multicast_support", getattr(hdl, "has_multicast_support", None))
    print("multicast_ptr", hex(hdl.multicast_ptr) if getattr(hdl, "multicast_ptr", 0) else hdl.multicast_ptr)
    print("buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs])
dist.barrier(); dist.destroy_process_group()
