#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/tp_decode_stack.py [--layers L] : tensor-parallel decode of the LLaMA-2-70B
linear stack (qkv column, o row, gate_up column, down row-parallel; no attention / norms) under one CUDA graph,
row-parallel reductions by NCCL vs the one-kernel peer all-reduce."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexq_b200 import capi, tp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=20)
ap.add_argument("--batches", default="1,16")
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
capi.load()
HID, INTER, QKV = 8192, 28672, 10240


def mk(N, K, mode, xb):
    w6, wsc = capi.quant_pack_w6((0.02 * torch.randn(N, K, device=dev)).half())
    return tp.TPLinearW6Ax.from_packed(w6, wsc, N, K, mode, xb, rank, world)


layers = []
for _ in range(a.layers):
    layers.append((mk(QKV // world, HID, "column", 6), mk(HID, HID // world, "row", 6),
                   mk(2 * INTER // world, HID, "column", 6), mk(HID, INTER // world, "row", 8)))


def run_stack(x, attn_in, mlp_in):
    for qkv, o, gu, down in layers:
        qkv.forward(x)
        x = o.forward(attn_in)
        gu.forward(x)
        x = down.forward(mlp_in)
    return x


def graph_us(fn, reps=10):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for B in [int(b) for b in a.batches.split(",")]:
    x = torch.randn(B, HID, device=dev).half()
    attn_in = torch.randn(B, HID // world, device=dev).half()
    mlp_in = torch.randn(B, INTER // world, device=dev).half()
    rec = {"model": "llama2-70b", "tp": world, "layers": a.layers, "batch": B}
    ref = run_stack(x, attn_in, mlp_in).clone()
    try:
        rec["nccl_us"] = graph_us(lambda: run_stack(x, attn_in, mlp_in))
    except Exception as e:       # noqa: BLE001
        rec["nccl_err"] = str(e)[:120]
    for _, o, _, down in layers:
        o.enable_oneshot_allreduce(16)
        down.enable_oneshot_allreduce(16)
    got = run_stack(x, attn_in, mlp_in).clone()
    torch.cuda.synchronize()
    rec["max_diff_vs_nccl"] = (got.float() - ref.float()).abs().max().item()
    rec["oneshot_us"] = graph_us(lambda: run_stack(x, attn_in, mlp_in))
    for _, o, _, down in layers:
        o._oneshot = None
        down._oneshot = None
    if "nccl_us" in rec:
        rec["speedup"] = rec["nccl_us"] / rec["oneshot_us"]
    rec["tok_s_oneshot"] = B / (rec["oneshot_us"] * 1e-6) * a.layers / 80 if False else B / (rec["oneshot_us"] * 1e-6 * 80 / a.layers)
    if rank == 0:
        print(json.dumps(rec), flush=True)
dist.destroy_process_group()
