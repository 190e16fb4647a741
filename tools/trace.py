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
def run(cta, reps=3):
  global tr
  for it in range(reps):
    tr = torch.zeros(a.units * 16 + 2 * 160, dtype=torch.int64, device=dev)
    if not a.noflush:
        flush.zero_()
        xq.add_(0); sx.add_(0)          # weights cold, activations L2-hot: what a layer sees right after its quantiser
    else:
        torch.cuda.synchronize()
    e0.record()
    capi.check(lib.flexq_debug_gemm_trace(capi._ptr(xq), capi._ptr(sx), capi._ptr(w6), capi._ptr(wsc), capi._ptr(out), a.m, a.n, a.k,
                                          capi._ptr(ws), capi._ptr(tr), a.units | (cta << 16), capi._stream()), "trace")
    e1.record()
    torch.cuda.synchronize()
    print("launch us", e0.elapsed_time(e1) * 1e3)
  return tr.cpu().numpy()


names = ["Wissue", "Wfull", "Aempty", "Afull", "MMArdy", "MMAcmt", "ACCfull", "ACCfree", "MMAdone", "Xissue", "FIXbeg", "FIXend", "END", "START",
         "FIXflag", "FIXld"]


def windows(full):
    w = full[a.units * 16:].reshape(160, 2)
    return w, np.flatnonzero(w[:, 0] > 0)


def table(full, cta):
    w, live = windows(full)
    t = full[:a.units * 16].reshape(a.units, 16)
    t0 = t[t > 0].min()            # clock64 stamps, relative to the traced CTA's first one
    print(f"CTA {cta}: wall window {(w[cta, 1] - w[cta, 0]) / 1e3:.2f} us, started {(w[cta, 0] - w[live, 0].min()) / 1e3:.2f} us after the first CTA")
    print("unit " + " ".join(f"{n:>8}" for n in names))
    for i in range(a.units):
        if (t[i] > 0).any():
            print(f"{i:4d} " + " ".join(f"{(t[i, e] - t0) if t[i, e] else -1:8d}" for e in range(16)))
    return t


full = run(max(a.cta, 0))
w, live = windows(full)
d_ = (w[live, 1] - w[live, 0]) / 1e3
print("CTAs", len(live), "start spread us", (w[live, 0].max() - w[live, 0].min()) / 1e3,
      "durations us: min %.1f median %.1f max %.1f" % tuple(np.percentile(d_, [0, 50, 100])), "span us", (w[live, 1].max() - w[live, 0].min()) / 1e3)
order = np.argsort(-d_)
print("slowest CTAs", live[order[:8]], np.round(d_[order[:8]], 2), "fastest", live[order[-4:]], np.round(d_[order[-4:]], 2))
print("end times (us after the first start), sorted:", np.round(np.sort((w[live, 1] - w[live, 0].min()) / 1e3)[::8], 1))
t = table(full, max(a.cta, 0))
if a.cta < 0:      # also the three slowest CTAs and the fastest one
    for c in [int(live[i]) for i in list(order[:3]) + [order[-1]]]:
        table(run(c, reps=2), c)
if (t[2:, 5] > 0).sum() > 2:
    d = np.diff(t[2:, 5][t[2:, 5] > 0])
    print("MMA commit interval cycles: mean %.0f median %.0f" % (d.mean(), np.median(d)))
