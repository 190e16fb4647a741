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
# (name, N, K, tensor-parallel mode)
LAYERS = [("attn_o_8192x8192", 8192, 8192, "row"),
          ("mlp_gate_28672x8192", 28672, 8192, "column"),
          ("mlp_down_8192x28672", 8192, 28672, "row")]
# dram__bytes_read.sum + dram__bytes_write.sum of one captured launch (28672x8192, M=2048): a CONSTANT taken from the ncu
# capture committed as profiles/ncu_prefill_r2.txt (ncu is not run inside the bench), stated as such in the JSON line
NCU_TRAFFIC_BYTES = 224.54e6 + 114.07e6
NCU_TRAFFIC_NOTE = ("constant from profiles/ncu_prefill_r2.txt (ncu --set full of the 28672x8192 M=2048 launch: 224.5 MB read + "
                    "114.1 MB written), not measured in this run; algorithmic bytes of that launch 314 MB (packed W 176 + X 17 + D 117 + scales 4)")


def workload(xb):
    return (f"llama2-70b linears {{8192x8192, 28672x8192, 8192x28672}} W6A{xb} g128 prefill M=2048 "
            "(A6 and A8 activations both travel in int8 containers: same bytes, same MMAs)")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


def bind_to_gpu_numa_node(gpu_index: int):
    """Multi-rank runs: pin this process to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the
    end-to-end path are allocated on the NUMA node the GPU's PCIe root hangs off (one process per GPU, eight processes
    otherwise allocate wherever the launcher happened to run).  Returns the CPU count bound to, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:                                # noqa: BLE001 -- best effort: no NVML, containerised CPU sets, ...
        pass
    return None


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
def cpu_reference_run(steps: int, warmup: int, sample_rows: int, ab: int = 6):
    from oracle.fakequant_torch import FakeQuantLinearCPU      # baseline leg: the only oracle use here
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    mods, xs = [], []
    for _, N, K, _ in LAYERS:
        mods.append(FakeQuantLinearCPU(0.02 * torch.randn(N, K, generator=g), ab, faithful=True))
        xs.append(torch.randn(sample_rows, K, generator=g))
    ops = sum(2.0 * sample_rows * N * K for _, N, K, _ in LAYERS)
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
    tops, ms, cores, sample = cpu_reference_run(steps, warmup, sample_rows=32, ab=args.xbits)
    line = {"impl": "reference", "metric": f"W6A{args.xbits} GEMM TOPS (2*M*N*K/t), llama2-70b linear layers", "value": tops,
            "unit": "TOPS", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload(args.xbits), "arm": "reference CPU fake-quant path (oracle port)"},
            "cpu_baseline": {"value": tops, "unit": "TOPS", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": tops, "unit": "TOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# --config c1: BASELINE.json configs[0], the reference's own CPU-runnable case, as a side record
# --------------------------------------------------------------------------------------------
def run_c1():
    """nn.Linear(4096, 4096, bias=False) default init under torch.manual_seed(0), x = randn(16, 4096), W6A6 g128, fp32:
    the reference's CPU fake-quant forward (faithful = weights re-quantised per call, and pre-quantised) on all host
    cores beside the CUDA path on the same fp32 module (fp32-arithmetic quantisers, flexq_quant_act_f32 /
    flexq_quant_pack_w6_f32); outputs compared."""
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
    q = QuantLinear(lin.cuda(), model_pack.default_quant_params(6, True), model_pack.default_quant_params(6, False))
    q.set_quant_state(True, True)
    xg = x.cuda()
    y = q(xg)
    us = graph_time_us([lambda: q(xg)], iters_per_copy=20)
    ref = y_cpu.double()
    err = (y.double().cpu() - ref)
    line = {"config": "C1: QuantLinear W6A6 g128, nn.Linear(4096,4096) seed 0, x = randn(16,4096), fp32 module",
            "cpu_faithful_ms": ms_f, "cpu_prequantized_ms": ms_p, "cpu_cores": cores, "cpu_kind": "port (oracle/fakequant_torch.py, fp32)",
            "gpu_us_graph": us, "speedup_vs_faithful": ms_f * 1e3 / us, "speedup_vs_prequantized": ms_p * 1e3 / us,
            "rms_rel_vs_cpu_fp32": float(err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()),
            "max_abs_over_mean_abs": float(err.abs().max() / ref.abs().mean()),
            "tolerance": "rms-rel <= 1e-3, max-abs <= 1e-2 * mean|ref| (tests/test_gpu_parity.py)",
            "note": "same integers and fp32 activation scales as the CPU path; weight scales rounded to fp16, fp16 GEMM output"}
    print(json.dumps(line), flush=True)


def graph_time_us(fn_list, iters_per_copy=4, reps=5):
    """Mean us per call from CUDA-graph replays of the calls in fn_list (one per rotated weight copy)."""
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


# --------------------------------------------------------------------------------------------
# side records the metric asks for ("... vs M ...; decode tok/s"), measured in the same run at N = 1 and attached
# to the JSON line under "extra" (the headline stays the M = 2048 step)
# --------------------------------------------------------------------------------------------
def extra_records(capi, layers, fp16_w, dev, xb, hbm_gbs):
    extra = {}
    L2 = 126 << 20
    # decode: fused quantise + GEMM at M = 1 and 16 under a CUDA graph, weights rotated through > 2 x L2 of copies
    dec = []
    for (name, N, K, _), lin in zip(LAYERS, layers):
        wbytes = lin.w6.numel()
        ncopy = max(1, min(8, (2 * L2 + wbytes - 1) // wbytes))
        copies = [(lin.w6, lin.w_scale)] + [(lin.w6.clone(), lin.w_scale.clone()) for _ in range(ncopy - 1)]
        for M in (1, 16):
            x = torch.randn(M, K, device=dev).half()
            out = torch.empty(M, N, dtype=torch.float16, device=dev)
            ws = capi.new_workspace(M, K)
            us = graph_time_us([lambda c=c: capi.linear_w6ax(x, c[0], c[1], N, xb, ws, capi.ROUND_CUDA, out) for c in copies])
            nbytes = N * K * 6 // 8 + N * (K // 128) * 2 + M * K * 2 + M * N * 2
            dec.append({"shape": f"{N}x{K}", "M": M, "us": us, "hbm_gbs": nbytes / us / 1e3, "hbm_frac": nbytes / us / 1e3 / hbm_gbs})
        del copies
    extra["decode"] = dec
    extra["decode_note"] = ("fused activation quantise + W6Ax GEMM, CUDA-graph replays, weight copies rotated (> 2 x L2); bytes = packed W + "
                            "w-scales + fp16 X + fp16 D; hbm_frac against hbm_gbs of MEASURED_PEAKS.json")
    # M = 2048 GEMM beside cuBLAS FP16 (torch.matmul) and cuBLAS INT8 (torch._int_mm) on the same shapes
    vs = []
    for (name, N, K, _), lin, w in zip(LAYERS, layers, fp16_w):
        M = M_TOKENS
        x = torch.randn(M, K, device=dev).half()
        xq, sx = capi.quant_act(x, xb)
        out = torch.empty(M, N, dtype=torch.float16, device=dev)
        gws = capi.new_workspace()
        rec = {"shape": f"{N}x{K}", "M": M,
               "w6ax_gemm_us": graph_time_us([lambda: capi.gemm_w6ax(xq, sx, lin.w6, lin.w_scale, N, gws, out)]),
               "cublas_f16_us": graph_time_us([lambda: torch.matmul(x, w.t(), out=out)])}
        try:
            xi8 = torch.randint(-32, 32, (M, K), device=dev, dtype=torch.int8)
            wi8 = torch.randint(-32, 32, (K, N), device=dev, dtype=torch.int8)
            rec["cublas_i8_us"] = graph_time_us([lambda: torch._int_mm(xi8, wi8)])
            del xi8, wi8
        except Exception as e:                       # noqa: BLE001
            rec["cublas_i8_err"] = str(e)[:80]
        rec["vs_cublas_f16"] = rec["cublas_f16_us"] / rec["w6ax_gemm_us"]
        vs.append(rec)
    extra["vs_cublas"] = vs
    # BASELINE.json configs[1]: LLaMA-2-7B linear shapes, W6A8, a few M, fused path beside cuBLAS FP16 (the full sweep with
    # cuBLAS INT8 is tools/sweep.py -> profiles/sweep_r2.md)
    sw = []
    for name, N, K in (("qkvo_4096x4096", 4096, 4096), ("gateup_11008x4096", 11008, 4096), ("down_4096x11008", 4096, 11008)):
        w = (0.02 * torch.randn(N, K, device=dev)).half()
        ncopy = max(1, min(8, (2 * L2 + N * K * 6 // 8 - 1) // (N * K * 6 // 8)))
        copies = [capi.quant_pack_w6(w)] + [None] * (ncopy - 1)
        for i in range(1, ncopy):
            copies[i] = (copies[0][0].clone(), copies[0][1].clone())
        for M in (1, 16, 256, 2048):
            x = torch.randn(M, K, device=dev).half()
            out = torch.empty(M, N, dtype=torch.float16, device=dev)
            ws = capi.new_workspace(M, K)
            t = graph_time_us([lambda c=c: capi.linear_w6ax(x, c[0], c[1], N, 8, ws, capi.ROUND_CUDA, out) for c in copies])
            tf = graph_time_us([lambda: torch.matmul(x, w.t(), out=out)])
            sw.append({"layer": name, "M": M, "w6a8_fused_us": t, "cublas_f16_us": tf, "vs_cublas_f16": tf / t,
                       "tops": 2.0 * M * N * K / t / 1e6})
        del copies, w
    extra["llama2_7b_w6a8"] = sw
    return extra


def decode_l3_8b_record():
    """BASELINE.json configs[3]: LLaMA-3-8B, mixed W6A6 / W6A8 (down_proj A8), synthetic weights, all 32 layers' norms,
    residuals, SiLU*up and linears under one CUDA graph per token step (attention / KV cache excluded), beside the same
    chain in fp16 (rms_norm / silu / cuBLAS)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import decode_stack
    recs = decode_stack.run_stack("llama3-8b", 32, [1, 16], True, True, with_fp16=True, clone_layers=True, verbose=False)
    return [{"batch": r["batch"], "layers": r["layers"], "us_per_step": r["w6ax_us"], "tok_s": r["w6ax_tok_s"], "hbm_frac": r["hbm_frac"],
             "fp16_us_per_step": r["fp16_us"], "speedup_vs_fp16": r["speedup_vs_fp16"]} for r in recs]


def decode_tok_s_record():
    """80-layer LLaMA-2-70B W6 chain (norms, residuals, SiLU*up through the fused producers, all linears; no attention /
    KV) under one CUDA graph: tools/decode_stack.py --chain --fuse-gate-up."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import decode_stack
    recs = decode_stack.run_stack("llama2-70b", 80, [1, 16], True, True, with_fp16=False, clone_layers=True, verbose=False)
    return [{"batch": r["batch"], "layers": r["layers"], "us_per_step": r["w6ax_us"], "tok_s": r["w6ax_tok_s"], "hbm_frac": r["hbm_frac"]} for r in recs]


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
    ap.add_argument("--no-extra", action="store_true", help="skip the side records (decode, cuBLAS comparators, 80-layer decode chain)")
    ap.add_argument("--xbits", type=int, default=6, choices=[6, 8],
                    help="activation bits: 6 = BASELINE.json configs[2] (default), 8 = configs[4] (the W6A8 tensor-parallel stack)")
    ap.add_argument("--config", default="c3", choices=["c3", "c1"], help="c3 = the bench workload (default); c1 = side record")
    args = ap.parse_args()
    if args.config == "c1":
        run_c1()
        return
    if args.impl == "reference":
        run_reference_arm(args)
        return
    args.warmup = max(args.warmup, 3)
    xb = args.xbits

    import torch.distributed as dist
    from flexq_b200 import capi, tp

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if world > 1 and os.environ.get("FLEXQ_BENCH_NUMA", "1") == "1":
        numa = bind_to_gpu_numa_node(local)      # before any pinned host allocation: first touch places the pages
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    capi.load()
    sampler = ClockSampler(local) if rank == 0 else None      # started early: nvidia-smi needs a moment before its first sample

    # ---- build the (sharded) layers: synthetic random-init weights (the same full matrices on every rank, each rank
    # packs its shard offline on the GPU); rank 0 keeps what the tp check below needs
    torch.manual_seed(1234)
    layers, x_dev, x_host, y_host, fp16_w, check = [], [], [], [], [], []
    for name, N, K, mode in LAYERS:
        w_full = (0.02 * torch.randn(N, K, device=dev)).half()
        x_full = torch.randn(M_TOKENS, K, device=dev).half()
        w = w_full if world == 1 else tp.shard_weight(w_full, mode, rank, world)
        x = x_full if (world == 1 or mode == "column") else tp.shard_activation(x_full, rank, world)
        Nl, Kl = w.shape
        w6, wsc = capi.quant_pack_w6(w)
        if world > 1 and rank == 0:                  # a 256 x 256 output block recomputed at tp = 1 after the timed region
            check.append((x_full[:256].clone(), w_full[:256].clone()))
        fp16_w.append(w if world == 1 and not args.no_extra else None)
        del w_full, x_full
        lin = tp.TPLinearW6Ax.from_packed(w6, wsc, Nl, Kl, mode, xb, rank, world)
        layers.append(lin)
        x_dev.append(x)
        x_host.append(x.cpu().pin_memory())
    outs = [torch.empty(M_TOKENS, l.N, dtype=torch.float16, device=dev) for l in layers]
    total_ops = sum(2.0 * M_TOKENS * N * K for _, N, K, _ in LAYERS)
    stream = torch.cuda.current_stream()

    # Row-parallel reduction: our peer-memory all-reduce (flexq_allreduce_sum_synced_f16 over NVLink / NVSwitch symmetric
    # memory).  tp = 8: NVSwitch multicast path, two token chunks so the reduction of chunk 0 overlaps the GEMM of chunk 1
    # on 8 SMs the GEMM leaves free.  tp = 2, 4: peer-pointer path after the whole GEMM.
    # FLEXQ_BENCH_AR=nccl|peer overrides.  Any failure to set up symmetric memory falls back to NCCL.
    ar_mode = "nccl"
    want_peer = os.environ.get("FLEXQ_BENCH_AR", "peer") == "peer"
    ar_chunks = int(os.environ.get("FLEXQ_BENCH_AR_CHUNKS", "2" if world >= 8 else "1"))
    ar_reserve = int(os.environ.get("FLEXQ_BENCH_AR_RESERVE", "8"))
    ar_mc = os.environ.get("FLEXQ_BENCH_AR_MC", "1" if world >= 8 else "0") == "1"      # NVSwitch multicast (in-switch sum) or peer pointers
    if world > 1 and want_peer:
        try:
            for lin in layers:
                lin.enable_peer_allreduce(M_TOKENS, chunks=ar_chunks, use_multicast=ar_mc, sm_reserve=ar_reserve)
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
        res = []
        for lin, x, o in zip(layers, x_dev, outs):
            if getattr(lin, "_ar", None) is not None:
                res.append(lin.forward(x))                    # result stays in the symmetric buffer
            else:
                res.append(lin.forward(x, o))
        return res

    # end to end: pinned host activations in, pinned host outputs back, through the host-buffer front end
    # (H2D / kernels / D2H on three streams, per-layer staging buffers -- flexq_b200/host_io.py).  Every rank holds the
    # all-reduced output of a row-parallel layer: rank r hands rows [r*M/tp, (r+1)*M/tp) to the host, so the D2H traffic
    # of the job is one copy of every output, spread over the ranks' PCIe links; sharded outputs go back from every rank.
    from flexq_b200.host_io import HostStagedLinears
    rows = []
    for (_, N, K, mode), lin in zip(LAYERS, layers):
        r0, r1 = (rank * M_TOKENS // world, (rank + 1) * M_TOKENS // world) if (mode == "row" and world > 1) else (0, M_TOKENS)
        rows.append((r0, r1))
        y_host.append(torch.empty(r1 - r0, lin.N, dtype=torch.float16).pin_memory())
    pipe = HostStagedLinears(layers, M_TOKENS, dev, rows)

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

    # ---- dominant kernel (the W6Ax GEMM) timed alone on pre-quantised operands -> roofline
    pre = []
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
    hbm_gbs, bf16_tf, src = peaks()
    peak = 2.0 * bf16_tf                       # kind::i8 issues at twice the bf16 rate on sm_100
    achieved = total_ops / world / (ms_gemm * 1e-3) / 1e12
    roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TOPS", "frac": achieved / peak,
            "traffic": NCU_TRAFFIC_BYTES if world == 1 else None, "traffic_note": NCU_TRAFFIC_NOTE,
            "kernel": "flexq::w6ax_gemm_kernel (the 192-token-tile instantiation at this M)", "launches_per_step": len(LAYERS),
            "avg_launch_ms": ms_gemm / len(LAYERS),
            "peak_source": f"2 x bf16_tflops of MEASURED_PEAKS.json ({src}); int8 dense = 2x bf16 dense on sm_100"}

    # ---- tensor-parallel numerics: a 256 x 256 block of every layer's output on rank 0 against the same block computed
    # at tp = 1 from the unsharded weight rows and activations (fp16 partial sums are rounded before the reduction, so
    # row-parallel outputs agree to fp16 rounding, column-parallel ones exactly)
    tp_check = None
    if world > 1:
        res = step_device()
        torch.cuda.synchronize()
        if rank == 0:
            tp_check = []
            for (name, N, K, mode), y, (xb256, wb256) in zip(LAYERS, res, check):
                w6b, wscb = capi.quant_pack_w6(wb256)
                ref = capi.linear_w6ax(xb256, w6b, wscb, 256, xb, capi.new_workspace(256, K)).float()
                got = y[:256, :256].float()
                err = (got - ref).abs()
                tp_check.append({"layer": name, "mode": mode, "max_abs_over_mean_abs": float(err.max() / ref.abs().mean()),
                                 "rms_rel": float(err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())})

    # ---- BASELINE configs[4] as written: the same tensor-parallel stack with A8 activations (same int8 containers, same
    # MMAs, only the quantiser's range differs), timed like the main step; every rank takes part
    c5 = None
    if world > 1 and xb != 8:
        for lin in layers:
            lin.x_bits = 8
        ms_a8 = timed(step_device, args.steps, 3)
        for lin in layers:
            lin.x_bits = xb
        c5 = {"workload": workload(8), "parallelism": f"tp{world}", "value": total_ops / (ms_a8 * 1e-3) / 1e12, "unit": "TOPS",
              "ms_per_step": ms_a8, "steps": args.steps}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        tops, _, cores, sample = cpu_reference_run(steps=2, warmup=1, sample_rows=32, ab=xb)
        cpu = {"value": tops, "unit": "TOPS", "cores": cores, "kind": "port", "sample": sample}

    extra = {"c5_w6a8": c5} if c5 else None
    if world == 1 and not args.no_extra:
        extra = extra_records(capi, layers, fp16_w, dev, xb, hbm_gbs)
        del layers, fp16_w, x_dev, outs, pre, pipe
        torch.cuda.empty_cache()
        try:
            extra["decode_tok_s"] = decode_tok_s_record()
            extra["decode_tok_s_note"] = ("LLaMA-2-70B, all 80 layers, W6A6 (down_proj W6A8), synthetic weights, one CUDA graph per token step: "
                                          "norms, residuals, SiLU*up and every linear; attention / KV cache excluded")
        except Exception as e:                       # noqa: BLE001
            extra["decode_tok_s_err"] = str(e)[:200]
        torch.cuda.empty_cache()
        try:
            extra["decode_llama3_8b_mixed"] = decode_l3_8b_record()
        except Exception as e:                       # noqa: BLE001
            extra["decode_llama3_8b_err"] = str(e)[:200]

    h2d = sum(x.numel() * 2 for x in x_host)
    d2h = sum(y.numel() * 2 for y in y_host)
    line = {
        "metric": f"W6A{xb} GEMM TOPS (2*M*N*K/t), llama2-70b linear layers", "value": total_ops / (ms_step * 1e-3) / 1e12,
        "unit": "TOPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "s8", "data": "synthetic",
        "config": {"workload": workload(xb), "tokens": M_TOKENS, "x_bits": xb, "parallelism": f"tp{world}",
                   "l2": "per-step working set (384 MB packed weights + activations) exceeds the 126 MB L2",
                   "step": "fused activation quantise + W6Ax GEMM per layer" + (
                       "" if world == 1 else " + NCCL all-reduce on row-parallel layers" if ar_mode == "nccl" else
                       " + peer-memory all-reduce (own kernel, NVLink symmetric memory) on row-parallel layers"
                       + (f", {ar_chunks} token chunks overlapped with the GEMM" if ar_chunks > 1 else ""))},
        "e2e": {"value": total_ops / (ms_e2e * 1e-3) / 1e12, "unit": "TOPS", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "bytes_are": "per rank (every rank copies its own shard / row slice)",
                "path": "pinned host x -> H2D -> fused quantise + GEMM -> D2H -> pinned host y, every layer every step; "
                        "three streams so copies in both directions overlap the kernels",
                "numa_cpus_bound": numa},
        "gpu_launches": 2 * len(LAYERS) * args.steps,
        "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "tp_check": tp_check, "extra": extra,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
