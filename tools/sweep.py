#!/usr/bin/env python
"""M-sweep of the W6Ax GEMM on LLaMA layer shapes beside cuBLAS FP16 / INT8 (BASELINE.json
configs[1..3]): latency from CUDA-graph replays (so 2-8 us decode kernels are not hidden behind
launch overhead), weights rotated through > L2-size copies so every launch streams from HBM.

Writes JSON lines (one per shape x M x kernel) to --out and prints a table.
  TOPS  = 2*M*N*K / t                                   (engine/test/test_w6a6_kernel.cu:36-37)
  bytes = N*K*6/8 + N*(K/128)*2 + M*K*xb/8.. (see below)
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flexq_b200 import capi  # noqa: E402

SHAPES = {
    "7b": [("qkvo_4096x4096", 4096, 4096, 8), ("qkv_12288x4096", 12288, 4096, 8), ("gateup_11008x4096", 11008, 4096, 8),
           ("down_4096x11008", 4096, 11008, 8)],
    "70b": [("qo_8192x8192", 8192, 8192, 6), ("gateup_28672x8192", 28672, 8192, 6), ("down_8192x28672", 8192, 28672, 8)],
    "l3-8b": [("qkv_6144x4096", 6144, 4096, 6), ("o_4096x4096", 4096, 4096, 6), ("gateup_14336x4096", 14336, 4096, 6),
              ("down_4096x14336", 4096, 14336, 8)],
}
L2_BYTES = 126 << 20


def graph_time_us(fn_list, iters_per_copy=4, reps=5):
    """fn_list: callables (one per rotated buffer set).  Returns mean us per call."""
    torch.cuda.synchronize()
    for f in fn_list:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(iters_per_copy):
                for f in fn_list:
                    f()
    n = iters_per_copy * len(fn_list)
    g.replay()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", default="70b")
    ap.add_argument("--ms", default="1,2,4,8,16,32,64,128,256,512,1024,2048,4096")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--no-cublas", action="store_true")
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    i8_peak = 2 * peaks.get("bf16_tflops", 1590.0)
    capi.load()
    dev = torch.device("cuda")
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    rows = []
    for model in args.models.split(","):
        for name, N, K, xb in SHAPES[model]:
            wbytes = N * K * 6 // 8
            ncopy = max(1, min(8, (2 * L2_BYTES + wbytes - 1) // wbytes))
            w6s, wscs = [], []
            for _ in range(ncopy):
                w = (0.02 * torch.randn(N, K, device=dev)).half()
                w6, wsc = capi.quant_pack_w6(w)
                w6s.append(w6)
                wscs.append(wsc)
            wh = w                     # last fp16 weight for cuBLAS fp16 (L2 effects noted)
            wi8 = torch.randint(-32, 32, (K, N), device=dev, dtype=torch.int8)
            gws = capi.new_workspace()
            for M in [int(m) for m in args.ms.split(",")]:
                x = torch.randn(M, K, device=dev).half()
                xq, sx = capi.quant_act(x, xb, capi.ROUND_CUDA)
                out = torch.empty(M, N, dtype=torch.float16, device=dev)
                lws = capi.new_workspace(M, K)
                ops = 2.0 * M * N * K
                gemm_bytes = wbytes + N * (K // 128) * 2 + M * K + M * (K // 128) * 4 + M * N * 2
                fused_bytes = wbytes + N * (K // 128) * 2 + M * K * 2 + M * N * 2

                t_gemm = graph_time_us([lambda i=i: capi.gemm_w6ax(xq, sx, w6s[i], wscs[i], N, gws, out) for i in range(ncopy)])
                t_fused = graph_time_us([lambda i=i: capi.linear_w6ax(x, w6s[i], wscs[i], N, xb, lws, capi.ROUND_CUDA, out) for i in range(ncopy)])
                rec = {"model": model, "layer": name, "N": N, "K": K, "M": M, "x_bits": xb,
                       "gemm_us": t_gemm, "fused_us": t_fused, "gemm_tops": ops / t_gemm / 1e6, "fused_tops": ops / t_fused / 1e6,
                       "gemm_gbs": gemm_bytes / t_gemm / 1e3, "fused_gbs": fused_bytes / t_fused / 1e3,
                       "hbm_frac_gemm": gemm_bytes / t_gemm / 1e3 / hbm, "hbm_frac_fused": fused_bytes / t_fused / 1e3 / hbm,
                       "i8_frac_gemm": ops / t_gemm / 1e6 / i8_peak, "weight_copies": ncopy}
                if not args.no_cublas:
                    t_f16 = graph_time_us([lambda: torch.matmul(x, wh.t(), out=out)])
                    rec.update({"cublas_f16_us": t_f16, "cublas_f16_tops": ops / t_f16 / 1e6})
                    if M > 16 and M % 8 == 0:
                        xi8 = torch.randint(-32, 32, (M, K), device=dev, dtype=torch.int8)
                        try:
                            t_i8 = graph_time_us([lambda: torch._int_mm(xi8, wi8)])
                            rec.update({"cublas_i8_us": t_i8, "cublas_i8_tops": ops / t_i8 / 1e6})
                        except Exception as e:       # noqa: BLE001
                            rec["cublas_i8_err"] = str(e)[:80]
                rows.append(rec)
                print(json.dumps(rec), flush=True)
            del w6s, wscs, wh, wi8
            torch.cuda.empty_cache()
    with open(args.out, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
