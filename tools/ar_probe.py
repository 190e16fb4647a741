#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/ar_probe.py : time of the multicast two-shot all-reduce against the number of blocks it
may use (0 = free-running grid) for the sizes the bench reduces; FLEXQ_AR_UNROLL selects the vectors in flight per thread."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexq_b200 import capi, tp  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = capi.load()
M, N = 2048, 8192
ar = tp.PeerAllReduce(M * N, dev, None, True)
y = ar.view(M, N)
y.copy_(torch.randn(M, N, device=dev).half() * 0.01)


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


for blocks in (0, 8, 12, 16, 24, 32):
    lib.flexq_set_allreduce_blocks(blocks)
    res = []
    for rows in (2048, 1152, 896, 512):
        y.mul_(0.1)
        res.append((rows, timeit(lambda: ar.reduce_(0, rows * N))))
    if rank == 0:
        print(f"unroll {os.environ.get('FLEXQ_AR_UNROLL', 'default')} multicast {bool(ar.multicast_ptr)} blocks {blocks}: " +
              "  ".join(f"{r} rows {us:.1f} us ({r * N * 2 / us / 1e3:.0f} GB/s alg)" for r, us in res), flush=True)
lib.flexq_set_allreduce_blocks(0)
dist.destroy_process_group()
