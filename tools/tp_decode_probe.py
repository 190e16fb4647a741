#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/tp_decode_probe.py : decode row-parallel linears, one-kernel peer all-reduce vs NCCL."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexq_b200 import capi, tp  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
capi.load()
M, N = 2048, 8192


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# decode: one-kernel all-reduce vs NCCL on the row-parallel layers
for (Nn, Kf) in ((8192, 28672), (8192, 8192)):
    Kd = Kf // world
    w6, wsc = capi.quant_pack_w6((0.02 * torch.randn(Nn, Kd, device=dev)).half())
    for Md in (1, 16, 64):
        xd = torch.randn(Md, Kd, device=dev).half()
        lin_n = tp.TPLinearW6Ax.from_packed(w6, wsc, Nn, Kd, "row", 6, rank, world)
        lin_o = tp.TPLinearW6Ax.from_packed(w6, wsc, Nn, Kd, "row", 6, rank, world).enable_oneshot_allreduce(64)
        y0 = lin_n.forward(xd).clone()
        ys = [lin_o.forward(xd).clone() for _ in range(5)]
        torch.cuda.synchronize()
        err = max((y.float() - y0.float()).abs().max().item() for y in ys)
        t_n = timeit(lambda: lin_n.forward(xd), iters=50)
        t_o = timeit(lambda: lin_o.forward(xd), iters=50)
        if rank == 0:
            print(f"decode row-parallel {Nn}x{Kd} M={Md}: NCCL {t_n:.1f} us, one-shot peer {t_o:.1f} us, max diff {err:.4g}", flush=True)
dist.destroy_process_group()
