#!/usr/bin/env python
"""QuantLlamaMLP (LLaMA-2-70B sizes) with SiLU*up in the gate_up GEMM's epilogue against the separate SiLU*up+quantise pass:
CUDA-graph replay time per forward for a few token counts."""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexq_b200 import QuantLlamaMLP, capi, model_pack  # noqa: E402

hid, inter = 8192, 28672
dev = torch.device("cuda")


class Org(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.gate_proj = torch.nn.Linear(hid, inter, bias=False)
        self.up_proj = torch.nn.Linear(hid, inter, bias=False)
        self.down_proj = torch.nn.Linear(inter, hid, bias=False)


torch.manual_seed(0)
org = Org().half().to(dev)
args = types.SimpleNamespace(weight_quant_params=model_pack.default_quant_params(6, True), act_quant_params=model_pack.default_quant_params(6, False),
                             act_down_proj_quant_params=model_pack.default_quant_params(8, False), flex_linear_quant=True)


def graph_us(fn, reps=5, inner=4):
    fn(); torch.cuda.synchronize()
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
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / inner)
    return best


for fuse in (False, True):
    mlp = QuantLlamaMLP(org, hid, inter, "silu", args)
    mlp.set_quant_state(True, True)
    mlp.fuse_silu_epilogue = fuse
    for M in (16, 256, 2048, 4096):
        x = torch.randn(M, hid, device=dev).half()
        with torch.no_grad():
            us = graph_us(lambda: mlp(x))
        print(f"fuse_silu_epilogue={fuse} M={M}: {us:.1f} us per MLP forward", flush=True)
    del mlp
    torch.cuda.empty_cache()
