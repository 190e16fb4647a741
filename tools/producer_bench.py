#!/usr/bin/env python
"""HBM throughput of the activation-side kernels: plain quantise, RMSNorm(+residual)+quantise, SiLU*up+quantise."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flexq_b200 import capi  # noqa: E402

capi.load()
dev = torch.device("cuda")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = peaks.get("hbm_gbs", 6650.0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def graph_us(fn, inner=16, reps=5):
    """Decode-size launches: CUDA-graph replays (a single eager launch through the Python wrapper reads ~8 us whatever the
    kernel does).  The inputs stay L2-resident, as they are right after the producing kernel in a decode step."""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
        with torch.cuda.graph(g, stream=s):
            for _ in range(inner):
                fn()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / inner)
    return best


def time_us(fn, reps=20, M=None):
    if M is not None and M <= 64:
        return graph_us(fn)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for M in (16, 2048, 8192):
    for K in (8192, 28672):
        x = torch.randn(M, K, device=dev).half()
        t = time_us(lambda: capi.quant_act(x, 6), M=M)
        b = M * K * 3
        print(json.dumps({"kernel": "quant_act", "M": M, "K": K, "us": round(t, 2), "alg_bytes": b, "gbs": round(b / t / 1e3, 1), "hbm_frac": round(b / t / 1e3 / hbm, 3)}), flush=True)
    K = 8192
    x = torch.randn(M, K, device=dev).half()
    g = torch.ones(K, device=dev).half()
    r = torch.randn(M, K, device=dev).half()
    t = time_us(lambda: capi.rmsnorm_quant(x, g, 1e-5, 6), M=M)
    b = M * K * 3
    print(json.dumps({"kernel": "rmsnorm_quant", "M": M, "K": K, "us": round(t, 2), "alg_bytes": b, "gbs": round(b / t / 1e3, 1), "hbm_frac": round(b / t / 1e3 / hbm, 3)}), flush=True)
    t = time_us(lambda: capi.rmsnorm_quant(x, g, 1e-5, 6, r), M=M)
    b = M * K * 7
    print(json.dumps({"kernel": "add_residual_rmsnorm_quant", "M": M, "K": K, "us": round(t, 2), "alg_bytes": b, "gbs": round(b / t / 1e3, 1), "hbm_frac": round(b / t / 1e3 / hbm, 3)}), flush=True)
    K = 28672
    gu = torch.randn(M, 2 * K, device=dev).half()
    t = time_us(lambda: capi.silu_mul_quant(gu[:, :K], gu[:, K:], 8), M=M)
    b = M * K * 5
    print(json.dumps({"kernel": "silu_mul_quant", "M": M, "K": K, "us": round(t, 2), "alg_bytes": b, "gbs": round(b / t / 1e3, 1), "hbm_frac": round(b / t / 1e3 / hbm, 3)}), flush=True)
