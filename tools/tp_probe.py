#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/tp_probe.py : symmetric-memory all-reduce check + timing vs NCCL."""
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


for use_mc in (True, False):
    ar = tp.PeerAllReduce(M * N, dev, None, use_mc)
    if rank == 0:
        print(f"world {world} multicast_ptr {ar.multicast_ptr:#x} (requested {use_mc}) peers {[hex(p) for p in ar.peer_ptrs]}", flush=True)
    torch.manual_seed(rank)
    part = torch.randn(M, N, device=dev).half()
    ref = part.float().clone()
    dist.all_reduce(ref)
    y = ar.view(M, N)
    y.copy_(part)
    ar.reduce_(0, M * N)
    torch.cuda.synchronize()
    err = (y.float() - ref).abs().max().item()
    if rank == 0:
        print(f"  max |peer - fp32 allreduce| = {err:.4g} (fp16 rounding of the sum expected ~ {ref.abs().max().item() * 2 ** -11:.3g})", flush=True)
    us = timeit(lambda: ar.reduce_(0, M * N))
    if rank == 0:
        print(f"  peer all-reduce 32 MB ({'multimem' if ar.multicast_ptr else 'p2p'}): {us:.1f} us", flush=True)
    for rows in (512,):
        us = timeit(lambda: ar.reduce_(0, rows * N))
        if rank == 0:
            print(f"  peer all-reduce {rows} x {N}: {us:.1f} us", flush=True)
    del ar
t = torch.randn(M, N, device=dev).half()
us = timeit(lambda: dist.all_reduce(t))
if rank == 0:
    print(f"NCCL all-reduce 32 MB: {us:.1f} us", flush=True)
for rows in (512,):
    tt = t[:rows]
    us = timeit(lambda: dist.all_reduce(tt))
    if rank == 0:
        print(f"NCCL all-reduce {rows} x {N}: {us:.1f} us", flush=True)

# row-parallel linear: NCCL path vs overlapped peer path
K = 28672 // world
w6, wsc = capi.quant_pack_w6((0.02 * torch.randn(N, K, device=dev)).half())
x = torch.randn(M, K, device=dev).half()
lin = tp.TPLinearW6Ax.from_packed(w6, wsc, N, K, "row", 6, rank, world)
y_nccl = lin.forward(x).clone()
us_nccl = timeit(lambda: lin.forward(x))
for chunks, reserve, mc, ff in ((1, 0, True, 0.0), (2, 8, True, 0.0), (2, 8, True, 0.36), (2, 8, True, 0.27), (2, 16, True, 0.36), (3, 8, True, 0.0)):
    lin.enable_peer_allreduce(M, chunks=chunks, use_multicast=mc, sm_reserve=reserve, first_frac=ff)
    y_peer = lin.forward(x).clone()
    torch.cuda.synchronize()
    err = (y_peer.float() - y_nccl.float()).abs().max().item()
    us_peer = timeit(lambda: lin.forward(x))
    if rank == 0:
        print(f"row-parallel down 8192x{K} M={M}: NCCL {us_nccl:.1f} us, peer-overlapped chunks={chunks} reserve={reserve} mc={mc} first={ff} {us_peer:.1f} us, max diff {err:.4g}", flush=True)
dist.destroy_process_group()
