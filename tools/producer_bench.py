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


def time_us(fn, reps=20):
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
        t = time_us(lambda: capi.quant_act(x, 6))
        b = M * K * 3
        print(json.dumps({"kernel": "quant_act", "M": M, "K": K, "us": round(t, 2), "alg_bytes": b, "gbs": round(b / t / 1e3, 1), "hbm_frac": round(b / t / 1e3 / hbm, 3)}), flush=True)
    K = 8192
    x = torch.randn(M, K, device=dev).half()
    g = torch.ones(K, device=dev).half()
    r = torch.randn(M, K, device=dev).half()
    t = time_us(lambda: capi.rmsnorm_quant(x, g, 1e-5, 6))
    b = M * K * 3
    print(json.dumps({"kernel": "rmsnorm_quant", "M": M, "K": K, "us": round(t, 2), "alg_bytes": b, "gbs": round(b / t / 1e3, 1), "hbm_frac": round(b / t / 1e3 / hbm, 3)}), flush=True)
    t = time_us(lambda: capi.rmsnorm_quant(x, g, 1e-5, 6, r))
    b = M * K * 7
    print(json.dumps({"kernel": "add_residual_rmsnorm_quant", "M": M, "K": K, "us": round(t, 2), "alg_bytes": b, "gbs": round(b / t / 1e3, 1), "hbm_frac": round(b / t / 1e3 / hbm, 3)}), flush=True)
    K = 28672
    gu = torch.randn(M, 2 * K, device=dev).half()
    t = time_us(lambda: capi.silu_mul_quant(gu[:, :K], gu[:, K:], 8))
    b = M * K * 5
    print(json.dumps({"kernel": "silu_mul_quant", "M": M, "K": K, "us": round(t, 2), "alg_bytes": b, "gbs": round(b / t / 1e3, 1), "hbm_frac": round(b / t / 1e3 / hbm, 3)}), flush=True)
