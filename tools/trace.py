#!/usr/bin/env python
"""Pipeline timeline of CTA 0 of the W6Ax GEMM (debug build with clock64 stamps)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexq_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=16)
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--k", type=int, default=8192)
ap.add_argument("--units", type=int, default=40)
ap.add_argument("--noflush", action="store_true")
ap.add_argument("--cta", type=int, default=0)
a = ap.parse_args()
lib = capi.load()
dev = torch.device("cuda")
w6, wsc = capi.quant_pack_w6((0.02 * torch.randn(a.n, a.k, device=dev)).half())
x = torch.randn(a.m, a.k, device=dev).half()
xq, sx = capi.quant_act(x, 6)
out = torch.empty(a.m, a.n, dtype=torch.float16, device=dev)
ws = capi.new_workspace()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(3):
    tr = torch.zeros(a.units * 16 + 2 * 160, dtype=torch.int64, device=dev)
    if not a.noflush:
        flush.zero_()
    else:
        torch.cuda.synchronize()
    e0.record()
    capi.check(lib.flexq_debug_gemm_trace(capi._ptr(xq), capi._ptr(sx), capi._ptr(w6), capi._ptr(wsc), capi._ptr(out), a.m, a.n, a.k,
                                          capi._ptr(ws), capi._ptr(tr), a.units | (a.cta << 16), capi._stream()), "trace")
    e1.record()
    torch.cuda.synchronize()
    print("launch us", e0.elapsed_time(e1) * 1e3)
full = tr.cpu().numpy()
w = full[a.units * 16:].reshape(160, 2)
w = w[w[:, 0] > 0]
print("CTAs", len(w), "start spread us", (w[:, 0].max() - w[:, 0].min()) / 1e3, "durations us: min %.1f median %.1f max %.1f" % tuple(np.percentile((w[:, 1] - w[:, 0]) / 1e3, [0, 50, 100])), "span us", (w[:, 1].max() - w[:, 0].min()) / 1e3)
d_ = (w[:, 1] - w[:, 0]) / 1e3
print("slowest CTAs", np.argsort(-d_)[:8], np.sort(-d_)[:8])
t = full[:a.units * 16].reshape(a.units, 16)
t0 = t[t > 0].min()
names = ["Wissue", "Wfull", "Aempty", "Afull", "MMArdy", "MMAcmt", "ACCfull", "ACCfree", "MMAdone", "Xissue", "FIXbeg", "FIXend", "END", "START"]
print("unit " + " ".join(f"{n:>8}" for n in names))
for i in range(a.units):
    print(f"{i:4d} " + " ".join(f"{(t[i, e] - t0) if t[i, e] else -1:8d}" for e in range(14)))
d = np.diff(t[2:, 5])
print("MMA commit interval cycles: mean %.0f median %.0f" % (d.mean(), np.median(d)))
