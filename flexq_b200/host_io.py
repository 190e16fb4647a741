"""Host-buffer front end of the W6Ax linears: the call a host-side caller makes when activations
live in (pinned) host memory, as in the C-ABI usage of INTEGRATION.md with host buffers.

Three CUDA streams form a pipeline -- copy-in (H2D), compute (fused activation quantise + W6Ax
GEMM, plus the NCCL all-reduce of row-parallel shards), copy-out (D2H) -- with one device
staging buffer pair per layer, so the H2D of layer i+1, the kernels of layer i and the D2H of
layer i-1 overlap and PCIe runs in both directions at once.  Events guard the staging buffers
against reuse by the next pass.  No CPU compute path: everything numerical runs in
libflexq_b200.so.
"""
from __future__ import annotations

import torch


class HostStagedLinears:
    def __init__(self, layers, max_tokens: int, device: torch.device, copy_out=None):
        """`copy_out[i]`: what this rank hands back to the host for layer i -- True (all rows), False (nothing) or a row
        range (r0, r1).  Tensor parallel: the all-reduced output of a row-parallel layer is identical on every rank, so
        each rank returns its own slice of the rows and the job's D2H traffic is spread over all PCIe links."""
        self.layers = list(layers)
        self.copy_out = list(copy_out) if copy_out is not None else [True] * len(self.layers)
        self.dev = device
        self.s_in, self.s_comp, self.s_out = (torch.cuda.Stream(device=device) for _ in range(3))
        self.x_dev = [torch.empty(max_tokens, l.K, dtype=torch.float16, device=device) for l in self.layers]
        self.y_dev = [torch.empty(max_tokens, l.N, dtype=torch.float16, device=device) for l in self.layers]
        mk = lambda: [torch.cuda.Event() for _ in self.layers]          # noqa: E731
        self.ev_in, self.ev_comp, self.ev_out = mk(), mk(), mk()
        self.passes = 0
        for l in self.layers:
            l.workspace(max_tokens)

    def run(self, xs_host, ys_host):
        """Enqueue one pass over all layers: ys_host[i] <- layer_i(xs_host[i]).  Asynchronous; call
        synchronize() (or wait on `done_event()`) before reading ys_host."""
        first = self.passes == 0
        for i, (lin, xh, yh) in enumerate(zip(self.layers, xs_host, ys_host)):
            M = xh.shape[0]
            xd, yd = self.x_dev[i][:M], self.y_dev[i][:M]
            with torch.cuda.stream(self.s_in):
                if not first:
                    self.s_in.wait_event(self.ev_comp[i])        # previous pass has consumed x_dev[i]
                xd.copy_(xh, non_blocking=True)
                self.ev_in[i].record(self.s_in)
            with torch.cuda.stream(self.s_comp):
                self.s_comp.wait_event(self.ev_in[i])
                if not first:
                    self.s_comp.wait_event(self.ev_out[i])       # previous pass has drained y_dev[i]
                lin.forward(xd, yd)
                self.ev_comp[i].record(self.s_comp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_comp[i])
                co = self.copy_out[i]
                if isinstance(co, tuple):
                    yh.copy_(yd[co[0]:co[1]], non_blocking=True)
                elif co:
                    yh.copy_(yd, non_blocking=True)
                self.ev_out[i].record(self.s_out)
        self.passes += 1

    def join(self, stream: torch.cuda.Stream):
        """Make `stream` wait for everything enqueued so far (for event timing on that stream)."""
        for s in (self.s_in, self.s_comp, self.s_out):
            ev = torch.cuda.Event()
            ev.record(s)
            stream.wait_event(ev)

    def fork(self, stream: torch.cuda.Stream):
        """Make the pipeline streams wait for work already enqueued on `stream`."""
        ev = torch.cuda.Event()
        ev.record(stream)
        for s in (self.s_in, self.s_comp, self.s_out):
            s.wait_event(ev)

    def synchronize(self):
        for s in (self.s_in, self.s_comp, self.s_out):
            s.synchronize()
