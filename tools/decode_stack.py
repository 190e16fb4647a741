#!/usr/bin/env python
"""Synthetic-weight decode of a LLaMA linear stack (BASELINE.json configs[3]: LLaMA-3-8B, mixed
W6A6 / W6A8 per the reference's --flex_linear_quant policy: down_proj gets 8-bit activations,
algorithm/models/int_llama_layer.py:31-43).

Every decoder layer's linears (qkv, o, gate, up, down) run through the fused flexq_b200 path
(activation quantise + W6Ax GEMM), all layers with their own weights, captured in one CUDA graph.
tok/s = batch / time of one pass over the stack.  Attention, KV cache, norms and sampling are NOT
included -- this measures the quantised-linear hot path only.  Printed beside the same stack in
cuBLAS FP16 (torch.matmul)."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flexq_b200 import capi  # noqa: E402

MODELS = {
    # hidden, intermediate, qkv_out, layers
    "llama3-8b": (4096, 14336, 6144, 32),
    "llama2-7b": (4096, 11008, 12288, 32),
    "llama2-70b": (8192, 28672, 10240, 80),
}


def graph_us(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            fn()
    g.replay()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--batches", default="1,2,4,8,16")
    ap.add_argument("--layers", type=int, default=0, help="override layer count (memory)")
    ap.add_argument("--fuse-gate-up", action="store_true", help="gate and up as one N=2*inter GEMM")
    ap.add_argument("--chain", action="store_true",
                    help="realistic producer chain: (residual+)RMSNorm+quant -> qkv, quant -> o, residual+RMSNorm+quant -> gate_up, "
                         "SiLU*up+quant -> down (needs --fuse-gate-up); fp16 side runs rms_norm / silu*mul / matmul")
    ap.add_argument("--no-fp16", action="store_true", help="skip the fp16 (cuBLAS) side: its weights take 2.7x the memory")
    ap.add_argument("--clone-layers", action="store_true", help="quantise + pack one layer and clone it (same timing, faster set-up)")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    results = run_stack(a.model, a.layers, [int(b) for b in a.batches.split(",")], a.fuse_gate_up, a.chain, not a.no_fp16, a.clone_layers)
    if a.out:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        with open(a.out, "w") as f:
            for r in results:
                f.write(json.dumps(r) + "\n")


def run_stack(model, layers=0, batches=(1,), fuse_gate_up=True, chain=True, with_fp16=True, clone_layers=False, verbose=True):
    """Time one decode step of `layers` decoder layers' linears (see the module docstring); returns one record per batch."""
    import types
    a = types.SimpleNamespace(model=model, fuse_gate_up=fuse_gate_up, chain=chain)
    hid, inter, qkv, nl = MODELS[model]
    layers = layers or nl
    dev = torch.device("cuda")
    capi.load()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    shapes = [("qkv", qkv, hid, 6), ("o", hid, hid, 6)]
    shapes += [("gate_up", 2 * inter, hid, 6)] if a.fuse_gate_up else [("gate", inter, hid, 6), ("up", inter, hid, 6)]
    shapes += [("down", hid, inter, 8)]
    packed, fp16 = [], []
    for li in range(layers):
        lw, lf = [], []
        for si, (name, N, K, xb) in enumerate(shapes):
            if clone_layers and li > 0:
                w6, ws = packed[0][si][0].clone(), packed[0][si][1].clone()
                w = fp16[0][si].clone() if with_fp16 else None
            else:
                w = (0.02 * torch.randn(N, K, device=dev)).half()
                w6, ws = capi.quant_pack_w6(w)
                if not with_fp16:
                    w = None
            lw.append((w6, ws, N, K, xb))
            lf.append(w)
        packed.append(lw)
        fp16.append(lf)
    wbytes = sum(N * K * 6 // 8 + N * (K // 128) * 2 for _, N, K, _ in shapes) * layers
    results = []
    for B in batches:
        xs = {K: torch.randn(B, K, device=dev).half() for _, _, K, _ in shapes}
        outs = {(N, K): torch.empty(B, N, dtype=torch.float16, device=dev) for _, N, K, _ in shapes}
        wss = {K: capi.new_workspace(B, K) for _, _, K, _ in shapes}

        def run_q():
            for lw in packed:
                for w6, ws, N, K, xb in lw:
                    capi.linear_w6ax(xs[K], w6, ws, N, xb, wss[K], capi.ROUND_CUDA, outs[(N, K)])

        def run_f():
            for lf in fp16:
                for w in lf:
                    torch.matmul(xs[w.shape[1]], w.t(), out=outs[(w.shape[0], w.shape[1])])

        if a.chain:
            assert a.fuse_gate_up
            gamma = torch.ones(hid, device=dev).half()
            h0 = torch.randn(B, hid, device=dev).half()
            attn = torch.randn(B, hid, device=dev).half()
            gws = capi.new_workspace()

            def run_q():                                     # noqa: F811
                h = h0.clone()
                resid = None
                for (wqkv, wo, wgu, wdn) in packed:
                    xq, sx, _ = capi.rmsnorm_quant(h, gamma, 1e-5, 6, resid)          # (residual +) norm + quant
                    capi.gemm_w6ax(xq, sx, wqkv[0], wqkv[1], wqkv[2], gws, outs[(wqkv[2], wqkv[3])])
                    resid = h if resid is None else resid                              # residual stream lives in `resid`
                    o = capi.linear_w6ax(attn, wo[0], wo[1], wo[2], 6, wss[wo[3]], capi.ROUND_CUDA, outs[(wo[2], wo[3])])
                    xq, sx, _ = capi.rmsnorm_quant(o, gamma, 1e-5, 6, resid)          # resid += o; norm; quant
                    gu = capi.gemm_w6ax(xq, sx, wgu[0], wgu[1], wgu[2], gws, outs[(wgu[2], wgu[3])])
                    xq, sx, _ = capi.silu_mul_quant(gu[:, :inter], gu[:, inter:], 8)
                    h = capi.gemm_w6ax(xq, sx, wdn[0], wdn[1], wdn[2], gws, outs[(wdn[2], wdn[3])])
                return h

            def run_f():                                     # noqa: F811
                h = h0.clone()
                resid = None
                F = torch.nn.functional
                for (wqkv, wo, wgu, wdn) in fp16:
                    resid = h if resid is None else resid + h
                    x = F.rms_norm(resid, (hid,), gamma, 1e-5)
                    torch.matmul(x, wqkv.t(), out=outs[(wqkv.shape[0], wqkv.shape[1])])
                    o = torch.matmul(attn, wo.t(), out=outs[(wo.shape[0], wo.shape[1])])
                    resid = resid + o
                    x = F.rms_norm(resid, (hid,), gamma, 1e-5)
                    gu = torch.matmul(x, wgu.t(), out=outs[(wgu.shape[0], wgu.shape[1])])
                    act = F.silu(gu[:, :inter]) * gu[:, inter:]
                    h = torch.matmul(act, wdn.t(), out=outs[(wdn.shape[0], wdn.shape[1])])
                return h

        tq = graph_us(run_q)
        tf = graph_us(run_f) if with_fp16 else None
        rec = {"model": a.model, "layers": layers, "batch": B, "fuse_gate_up": a.fuse_gate_up, "chain": a.chain, "linears_per_layer": len(shapes),
               "w6ax_us": tq, "w6ax_tok_s": B / tq * 1e6, "fp16_us": tf, "fp16_tok_s": B / tf * 1e6 if tf else None,
               "speedup_vs_fp16": tf / tq if tf else None,
               "weight_gbs": wbytes / tq / 1e3, "hbm_frac": wbytes / tq / 1e3 / hbm,
               "note": ("decoder layer without attention / KV: norms, residuals, SiLU*up and all linears; synthetic weights" if a.chain
                        else "linear stack only (no attention / KV / norms); synthetic weights")}
        results.append(rec)
        if verbose:
            print(json.dumps(rec), flush=True)
    return results


if __name__ == "__main__":
    main()
