#!/usr/bin/env python
"""Run the W6Ax GEMM on one shape a few times (for ncu / compute-sanitizer captures)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexq_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=16)
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--k", type=int, default=8192)
ap.add_argument("--xb", type=int, default=6)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--fused", action="store_true")
a = ap.parse_args()
capi.load()
dev = torch.device("cuda")
w6, wsc = capi.quant_pack_w6((0.02 * torch.randn(a.n, a.k, device=dev)).half())
x = torch.randn(a.m, a.k, device=dev).half()
xq, sx = capi.quant_act(x, a.xb)
out = torch.empty(a.m, a.n, dtype=torch.float16, device=dev)
ws = capi.new_workspace(a.m, a.k)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(a.iters):
    flush.zero_()                       # evict L2
    e0.record()
    if a.fused:
        capi.linear_w6ax(x, w6, wsc, a.n, a.xb, ws, capi.ROUND_CUDA, out)
    else:
        capi.gemm_w6ax(xq, sx, w6, wsc, a.n, ws, out)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"M={a.m} N={a.n} K={a.k} xb={a.xb} us/launch: " + " ".join(f"{t:.1f}" for t in ts))
