#!/usr/bin/env python
"""Host-side cost of one eager flexq_gemm_w6ax call (no CUDA graph): wall clock per call over many back-to-back launches of a
decode-size GEMM, i.e. what the CPU spends on argument checks, the plan, tensor-map encoding and the launch."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexq_b200 import capi  # noqa: E402

lib = capi.load()
dev = torch.device("cuda")
M, N, K = 16, 4096, 4096
w6, wsc = capi.quant_pack_w6((0.02 * torch.randn(N, K, device=dev)).half())
xq, sx = capi.quant_act(torch.randn(M, K, device=dev).half(), 6)
out = torch.empty(M, N, dtype=torch.float16, device=dev)
ws = capi.new_workspace()
args = (capi._ptr(xq), capi._ptr(sx), capi._ptr(w6), capi._ptr(wsc), capi._ptr(out), M, N, K, capi._ptr(ws), ws.numel(), capi._stream())
for _ in range(50):
    lib.flexq_gemm_w6ax(*args)
torch.cuda.synchronize()
n = 2000
t0 = time.perf_counter()
for _ in range(n):
    lib.flexq_gemm_w6ax(*args)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"FLEXQ_TMAP_CACHE={os.environ.get('FLEXQ_TMAP_CACHE', '1')}: {1e6 * (t1 - t0) / n:.2f} us of host time per eager call "
      f"({1e6 * (t2 - t0) / n:.2f} us per call including the GPU drain), M={M} N={N} K={K}")
