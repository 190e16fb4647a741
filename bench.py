#!/usr/bin/env python
"""Benchmark of the W6Ax quantized-linear hot path on B200 (contract: see DESIGN.md "Measurement").

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): the LLaMA-2-70B
linear layers 8192x8192, 28672x8192, 8192x28672 (N x K), W6A6, prefill M = 2048, synthetic
random-init fp16 weights and activations.  One step = one pass of the fused path
(dynamic activation quantise -> W6A6 tcgen05 GEMM -> fp16) over the three layers.
With --gpus N > 1 the same layers are tensor-parallel over N ranks (o_proj-like 8192x8192 and
down 8192x28672 row-parallel with an NCCL all-reduce, gate/up 28672x8192 column-parallel), i.e.
strong scaling.  Metric: TOPS = 2*M*N*K summed over the layers / time (the reference harness's
definition, engine/test/test_w6a6_kernel.cu:36-37).

`--impl reference` times the reference's own CPU implementation of the path (the fake-quant
QuantLinear, restated in oracle/fakequant_torch.py because /root/reference does not travel to
the GPU box) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_TOKENS = 2048
# (name, N, K, tensor-parallel mode, activation bits)
LAYERS = [("attn_o_8192x8192", 8192, 8192, "row", 6),
          ("mlp_gate_28672x8192", 28672, 8192, "column", 6),
          ("mlp_down_8192x28672", 8192, 28672, "row", 6)]
# dram__bytes_read.sum + dram__bytes_write.sum of the captured launch (ncu --set full, profiles/ncu_prefill_r1.txt)
NCU_TRAFFIC_BYTES = 223.15e6 + 110.06e6
NCU_TRAFFIC_NOTE = ("ncu capture of the 28672x8192 M=2048 launch: 223.2 MB read + 110.1 MB written; algorithmic bytes of that "
                    "launch 314 MB (packed W 176 + X 17 + D 117 + scales 4)")
WORKLOAD = "llama2-70b linears {8192x8192, 28672x8192, 8192x28672} W6A6 g128 prefill M=2048"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region.  The sampler process is started early (its
    start-up takes longer than a short timed region); only samples stamped between begin() and end() are reported.  If the
    region was too short to catch two samples, `stop(extend)` keeps the same kernels running for a fraction of a second
    while sampling, so that the reported clocks are always clocks under this load."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.thread = [], None, None
        self.t0, self.t1 = None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def begin(self):
        self.t0 = time.monotonic()

    def end(self):
        self.t1 = time.monotonic()

    def _window(self):
        t0 = self.t0 if self.t0 is not None else 0.0
        t1 = self.t1 if self.t1 is not None else time.monotonic()
        return [r for t, r in self.rows if t0 <= t <= t1 + 0.02 and r and r[0].isdigit()]

    def stop(self, extend=None):
        if self.proc is None:
            return None
        rows = self._window()
        if len(rows) < 2 and extend is not None:
            self.t0 = time.monotonic()
            while time.monotonic() - self.t0 < 0.6:
                extend()
            torch.cuda.synchronize()
            self.t1 = time.monotonic()
            rows = self._window()
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in rows)
        if not sm:
            return None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in rows)]
        mx = max(int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit())
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU fake-quant path on the host cores
# --------------------------------------------------------------------------------------------
def cpu_reference_run(steps: int, warmup: int, sample_rows: int):
    from oracle.fakequant_torch import FakeQuantLinearCPU      # baseline leg: the only oracle use here
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    mods, xs = [], []
    for _, N, K, _, ab in LAYERS:
        mods.append(FakeQuantLinearCPU(0.02 * torch.randn(N, K, generator=g), ab, faithful=True))
        xs.append(torch.randn(sample_rows, K, generator=g))
    ops = sum(2.0 * sample_rows * N * K for _, N, K, _, _ in LAYERS)
    for _ in range(warmup):
        for m, x in zip(mods, xs):
            m(x)
    t0 = time.perf_counter()
    for _ in range(steps):
        for m, x in zip(mods, xs):
            m(x)
    dt = (time.perf_counter() - t0) / steps
    sample = (f"{sample_rows} of {M_TOKENS} token rows through all three layers, fp32, weights re-fake-quantised every "
              f"forward as the reference does (int_linear.py:60-62); {steps} steps")
    return ops / dt / 1e12, dt * 1e3, cores, sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = min(args.steps, 3), min(args.warmup, 1)
    tops, ms, cores, sample = cpu_reference_run(steps, warmup, sample_rows=32)
    line = {"impl": "reference", "metric": "W6A6 GEMM TOPS (2*M*N*K/t), llama2-70b linear layers", "value": tops,
            "unit": "TOPS", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "arm": "reference CPU fake-quant path (oracle port)"},
            "cpu_baseline": {"value": tops, "unit": "TOPS", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": tops, "unit": "TOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# --config c1: BASELINE.json configs[0], the reference's own CPU-runnable case, as a side record
# --------------------------------------------------------------------------------------------
def run_c1():
    """nn.Linear(4096, 4096, bias=False) default init under torch.manual_seed(0), x = randn(16, 4096), W6A6 g128:
    the reference's CPU fake-quant forward (faithful = weights re-quantised per call, and pre-quantised) on all host
    cores beside the fused CUDA path; outputs compared."""
    from oracle.fakequant_torch import FakeQuantLinearCPU      # baseline leg
    from flexq_b200 import QuantLinear, model_pack
    torch.manual_seed(0)
    lin = torch.nn.Linear(4096, 4096, bias=False)
    x = torch.randn(16, 4096)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)

    def cpu_ms(mod, n):
        mod(x)
        t0 = time.perf_counter()
        for _ in range(n):
            y = mod(x)
        return (time.perf_counter() - t0) / n * 1e3, y

    ms_f, y_cpu = cpu_ms(FakeQuantLinearCPU(lin.weight.detach(), 6, faithful=True), 5)
    ms_p, _ = cpu_ms(FakeQuantLinearCPU(lin.weight.detach(), 6, faithful=False), 20)
    q = QuantLinear(lin.half().cuda(), model_pack.default_quant_params(6, True), model_pack.default_quant_params(6, False))
    q.set_quant_state(True, True)
    xg = x.half().cuda()
    y = q(xg)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        q(xg)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        q(xg)
    for _ in range(5):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 200 * 1e3
    ref = y_cpu.double()
    err = (y.double().cpu() - ref)
    line = {"config": "C1: QuantLinear W6A6 g128, nn.Linear(4096,4096) seed 0, x = randn(16,4096)",
            "cpu_faithful_ms": ms_f, "cpu_prequantized_ms": ms_p, "cpu_cores": cores, "cpu_kind": "port (oracle/fakequant_torch.py, fp32)",
            "gpu_fused_us_graph": us, "speedup_vs_faithful": ms_f * 1e3 / us, "speedup_vs_prequantized": ms_p * 1e3 / us,
            "rms_rel_vs_cpu_fp32": float(err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()),
            "note": "GPU path quantises in fp16 (module dtype) and rounds activations half-away; CPU port is the fp32 fake-quant forward"}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="flexq_b200", choices=["flexq_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="c3", choices=["c3", "c1"], help="c3 = the bench workload (default); c1 = side record")
    args = ap.parse_args()
    if args.config == "c1":
        run_c1()
        return
    if args.impl == "reference":
        run_reference_arm(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    from flexq_b200 import capi, tp

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    capi.load()
    sampler = ClockSampler(local) if rank == 0 else None      # started early: nvidia-smi needs a moment before its first sample

    # ---- build the (sharded) layers: synthetic random-init weights, packed offline on the GPU
    torch.manual_seed(1234)
    layers, x_dev, x_host, y_host = [], [], [], []
    for name, N, K, mode, ab in LAYERS:
        Nl, Kl = (N // world, K) if mode == "column" else (N, K // world)
        w = (0.02 * torch.randn(Nl, Kl, device=dev)).half()
        w6, wsc = capi.quant_pack_w6(w)
        del w
        lin = tp.TPLinearW6Ax.from_packed(w6, wsc, Nl, Kl, mode, ab, rank, world)
        layers.append(lin)
        x = torch.randn(M_TOKENS, Kl, device=dev).half()
        x_dev.append(x)
        x_host.append(x.cpu().pin_memory())
        y_host.append(torch.empty(M_TOKENS, Nl, dtype=torch.float16).pin_memory())
    outs = [torch.empty(M_TOKENS, l.N, dtype=torch.float16, device=dev) for l in layers]
    total_ops = sum(2.0 * M_TOKENS * N * K for _, N, K, _, _ in LAYERS)
    stream = torch.cuda.current_stream()

    # Row-parallel reduction: our peer-memory all-reduce (flexq_allreduce_sum_synced_f16 over NVLink / NVSwitch symmetric
    # memory).  tp = 8: NVSwitch multicast path, two token chunks so the reduction of chunk 0 overlaps the GEMM of chunk 1
    # on 8 SMs the GEMM leaves free (measured: 196 us vs 288 us with NCCL for down_proj).  tp = 2, 4: peer-pointer path
    # after the whole GEMM (the 32 MB reduction alone: 69 / 96 us vs NCCL 83 / 119 us).
    # FLEXQ_BENCH_AR=nccl|peer overrides.  Any failure to set up symmetric memory falls back to NCCL.
    ar_mode = "nccl"
    want_peer = os.environ.get("FLEXQ_BENCH_AR", "peer") == "peer"
    ar_chunks = int(os.environ.get("FLEXQ_BENCH_AR_CHUNKS", "2" if world >= 8 else "1"))
    ar_reserve = int(os.environ.get("FLEXQ_BENCH_AR_RESERVE", "8"))
    if world > 1 and want_peer:
        try:
            for lin in layers:
                lin.enable_peer_allreduce(M_TOKENS, chunks=ar_chunks, use_multicast=(world == 8), sm_reserve=ar_reserve)
            ar_mode = "peer"
        except Exception as e:                       # noqa: BLE001
            for lin in layers:
                lin._ar = None
            if rank == 0:
                print(f"[bench] peer all-reduce unavailable ({str(e)[:100]}); using NCCL", file=sys.stderr)
        flag = torch.tensor([1.0 if ar_mode == "peer" else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)           # all ranks must agree
        if flag.item() < 1.0:
            ar_mode = "nccl"
            for lin in layers:
                lin._ar = None

    def step_device():
        for lin, x, o in zip(layers, x_dev, outs):
            if getattr(lin, "_ar", None) is not None:
                lin.forward(x)                                # result stays in the symmetric buffer
            else:
                lin.forward(x, o)

    # end to end: pinned host activations in, pinned host outputs back, through the host-buffer front end
    # (H2D / kernels / D2H on three streams, per-layer staging buffers -- flexq_b200/host_io.py)
    from flexq_b200.host_io import HostStagedLinears
    # replicated (all-reduced) outputs go back to the host from rank 0 only; sharded outputs from every rank
    copy_out = [not (mode == "row" and world > 1 and rank != 0) for _, _, _, mode, _ in LAYERS]
    pipe = HostStagedLinears(layers, M_TOKENS, dev, copy_out)

    def timed_e2e(steps, warmup):
        for _ in range(warmup):
            pipe.run(x_host, y_host)
        pipe.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        pipe.fork(stream)
        for _ in range(steps):
            pipe.run(x_host, y_host)
        pipe.join(stream)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    if sampler:
        sampler.begin()
    ms_step = timed(step_device, args.steps, args.warmup)
    ms_e2e = timed_e2e(max(5, args.steps // 2), 3)

    # ---- dominant kernel (the W6A6 GEMM) timed alone on pre-quantised operands -> roofline
    pre = []
    if True:
        for lin, x in zip(layers, x_dev):
            xq, sx = capi.quant_act(x, lin.x_bits, capi.ROUND_CUDA)
            pre.append((xq, sx))
        gws = capi.new_workspace()

        def gemm_only():
            for lin, (xq, sx), o in zip(layers, pre, outs):
                capi.gemm_w6ax(xq, sx, lin.w6, lin.w_scale, lin.N, gws, o)

        ms_gemm = timed(gemm_only, args.steps, 3)
        if sampler:
            sampler.end()
        clocks = sampler.stop(extend=step_device if world == 1 else None) if sampler else None
        _, bf16_tf, src = peaks()
        peak = 2.0 * bf16_tf                       # kind::i8 issues at twice the bf16 rate on sm_100
        achieved = total_ops / world / (ms_gemm * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TOPS", "frac": achieved / peak,
                "traffic": NCU_TRAFFIC_BYTES if world == 1 else None, "traffic_note": NCU_TRAFFIC_NOTE,
                "kernel": "w6ax_gemm_kernel<M_TILE=192,GP=1>", "launches_per_step": len(LAYERS),
                "avg_launch_ms": ms_gemm / len(LAYERS),
                "peak_source": f"2 x bf16_tflops of MEASURED_PEAKS.json ({src}); int8 dense = 2x bf16 dense on sm_100"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        tops, _, cores, sample = cpu_reference_run(steps=2, warmup=1, sample_rows=32)
        cpu = {"value": tops, "unit": "TOPS", "cores": cores, "kind": "port", "sample": sample}

    h2d = sum(x.numel() * 2 for x in x_host)
    d2h = sum(y.numel() * 2 for y, c in zip(y_host, copy_out) if c)
    line = {
        "metric": "W6A6 GEMM TOPS (2*M*N*K/t), llama2-70b linear layers", "value": total_ops / (ms_step * 1e-3) / 1e12,
        "unit": "TOPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "s8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "tokens": M_TOKENS, "parallelism": f"tp{world}",
                   "l2": "per-step working set (384 MB packed weights + activations) exceeds the 126 MB L2",
                   "step": "fused activation quantise + W6A6 GEMM per layer" + (
                       "" if world == 1 else " + NCCL all-reduce on row-parallel layers" if ar_mode == "nccl" else
                       " + peer-memory all-reduce (own kernel, NVLink symmetric memory) on row-parallel layers"
                       + (f", {ar_chunks} token chunks overlapped with the GEMM" if ar_chunks > 1 else ""))},
        "e2e": {"value": total_ops / (ms_e2e * 1e-3) / 1e12, "unit": "TOPS", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "path": "pinned host x -> H2D -> fused quantise + GEMM -> D2H -> pinned host y, every layer every step; "
                        "three streams so copies in both directions overlap the kernels"},
        "gpu_launches": 2 * len(LAYERS) * args.steps,
        "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
