"""Model-level quantise -> pack driver and packed checkpoint format.

The reference stops at fake quantisation: ``flexqllm`` (algorithm/flexq_quantize/flexqllm.py:48-122)
wraps every decoder layer, turns the quant state on, overwrites the weights with their fake-quantised
values (``weight_quant_inplace``, flexq_quantize/utils.py:60-63) and registers the scales
(``register_scales_and_zeros``, utils.py:116-123); the engine side then expects separately packed
weights (FasterTransformer on-disk shapes, e2e/.../models/llama/LlamaDecoderLayerWeight.cc:381-406,
492-515).  This module is that missing link for the sm_100a path:

* ``replace_linears``  -- walk a module tree and swap the LLaMA linears for real-quant ``QuantLinear``
  with the reference's bit policy (down_proj A8 under ``flex_linear_quant``, everything else A6:
  algorithm/models/int_llama_layer.py:31-43,75-94);
* ``pack_model`` / ``save_packed`` / ``load_packed`` -- one entry per linear:
  ``{"w6": uint8[flexq_w6_packed_bytes], "w_scale": f16[K/128, N], "N", "K", "x_bits", "bias"}``;
* ``shard_packed`` -- tensor-parallel shard of a packed entry without re-quantising: column mode slices
  whole 128-row tiles, row mode slices k-groups (per-group quantisation makes both exact);
* ``PackedLinear`` -- inference module over an entry (no fp16 weight kept).

Everything numerical runs in libflexq_b200.so; tensors must be on the GPU.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import capi
from .int_linear import QuantLinear

LLAMA_LINEARS = ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj")
TILE_N, GROUP, TILE_BYTES = 128, capi.GROUP, 12288


def default_quant_params(n_bits: int, weight: bool) -> dict:
    """The dicts algorithm/main.py:256-296 builds for --wbits/--abits N --{w,a}_group_size 128 --symmetric."""
    return dict(n_bits=n_bits, per_channel_axes=[0] if weight else [], symmetric=True, dynamic_method="per_group",
                group_size=GROUP, disable_zero_point=True)


def replace_linears(model: nn.Module, weight_quant_params: dict | None = None, act_quant_params: dict | None = None,
                    act_down_proj_quant_params: dict | None = None, flex_linear_quant: bool = True,
                    names=LLAMA_LINEARS, act_round: int = capi.ROUND_CUDA) -> nn.Module:
    """Swap every nn.Linear whose attribute name is in ``names`` for a QuantLinear with quant state on."""
    wq = weight_quant_params or default_quant_params(6, True)
    aq = act_quant_params or default_quant_params(6, False)
    aq_down = act_down_proj_quant_params or default_quant_params(8, False)
    for parent in list(model.modules()):
        for name, child in list(parent.named_children()):
            if isinstance(child, nn.Linear) and name in names:
                a = aq_down if (name == "down_proj" and flex_linear_quant) else aq
                q = QuantLinear(child, wq, a, act_round=act_round)
                q.set_quant_state(True, True)
                setattr(parent, name, q)
    return model


# ---- model-level helpers with the reference's names and semantics (algorithm/flexq_quantize/utils.py) --------------
def set_quant_state(model: nn.Module, weight_quant: bool = False, act_quant: bool = False):
    """utils.py:125-131 -- switch every QuantLinear of the model."""
    for m in model.modules():
        if isinstance(m, QuantLinear):
            m.set_quant_state(weight_quant, act_quant)


@torch.no_grad()
def weight_quant_inplace(model: nn.Module):
    """utils.py:116-123 -- overwrite every QuantLinear weight by its fake-quantised value (the weight quantiser also
    records the scales it used).  The real-quant path repacks lazily from the new weight."""
    for m in model.modules():
        if isinstance(m, QuantLinear):
            m.weight = m.weight_quantizer(m.weight)
            m.use_temporary_parameter = False


def register_scales_and_zeros(model: nn.Module):
    """utils.py:60-63 -- keep the weight quantisers' last scales / zero points as buffers."""
    for m in model.modules():
        if isinstance(m, QuantLinear):
            m.weight_quantizer.register_scales_and_zeros()


@torch.no_grad()
def pack_model(model: nn.Module) -> dict:
    """{qualified name: packed entry} for every kernel-backed QuantLinear of ``model`` (weights on the GPU)."""
    out = {}
    for name, m in model.named_modules():
        if isinstance(m, QuantLinear) and m.kernel_supported():
            w6, ws = m.pack_weights()
            out[name] = {"w6": w6, "w_scale": ws, "N": m.out_features, "K": m.in_features,
                         "x_bits": m.act_quantizer.n_bits, "bias": m.bias}
    return out


def save_packed(packed: dict, path: str):
    cpu = {k: {f: (v.cpu() if torch.is_tensor(v) else v) for f, v in e.items()} for k, e in packed.items()}
    torch.save({"format": "flexq_b200.w6g128.v1", "linears": cpu}, path)


def load_packed(path: str, device="cuda") -> dict:
    blob = torch.load(path, map_location="cpu")
    if blob.get("format") != "flexq_b200.w6g128.v1":
        raise capi.FlexQError(f"{path}: not a flexq_b200 packed checkpoint")
    return {k: {f: (v.to(device) if torch.is_tensor(v) else v) for f, v in e.items()} for k, e in blob["linears"].items()}


def shard_packed(entry: dict, mode: str, rank: int, world: int) -> dict:
    """Tensor-parallel shard of a packed linear (same result as packing the sharded fp16 weight)."""
    N, K = entry["N"], entry["K"]
    G, nt = K // GROUP, (N + TILE_N - 1) // TILE_N
    w6 = entry["w6"].view(nt, G, TILE_BYTES)
    ws = entry["w_scale"]
    if mode == "column":
        if N % (world * TILE_N):
            raise ValueError(f"column shard needs N={N} divisible by {world}*128")
        n = N // world
        t0, t1 = rank * n // TILE_N, (rank + 1) * n // TILE_N
        bias = entry.get("bias")
        return {**entry, "w6": w6[t0:t1].contiguous().view(-1), "w_scale": ws[:, rank * n:(rank + 1) * n].contiguous(), "N": n,
                "bias": None if bias is None else bias[rank * n:(rank + 1) * n].contiguous()}
    if mode == "row":
        if G % world:
            raise ValueError(f"row shard needs K/128={G} divisible by {world}")
        g = G // world
        bias = entry.get("bias")
        return {**entry, "w6": w6[:, rank * g:(rank + 1) * g].contiguous().view(-1), "w_scale": ws[rank * g:(rank + 1) * g].contiguous(),
                "K": K // world, "bias": bias if rank == 0 else None}
    raise ValueError(f"unknown tensor-parallel mode {mode!r}")


def interleave_gate_up(gate_w: torch.Tensor, up_w: torch.Tensor) -> torch.Tensor:
    """[2 * inter, K] weight for ``flexq_gemm_w6ax_silu_mul``: 8 rows of gate_proj, the same 8 rows of up_proj, the next 8 of
    gate_proj, ... so that one 128-row weight tile of the GEMM holds gate and up of 64 output columns and its epilogue can
    apply SiLU(gate) * up.  Per-(row, group) quantisation is row-wise, so the packed integers and scales are those of the
    two layers packed separately."""
    inter, K = gate_w.shape
    if up_w.shape != gate_w.shape or inter % 8:
        raise ValueError("gate_proj / up_proj must have the same [inter, K] shape with inter % 8 == 0")
    return torch.stack((gate_w.reshape(inter // 8, 8, K), up_w.reshape(inter // 8, 8, K)), dim=1).reshape(2 * inter, K)


def deinterleave_gate_up(y: torch.Tensor):
    """(gate, up) halves of a [..., 2 * inter] tensor laid out by ``interleave_gate_up`` (output of the plain GEMM)."""
    lead, n2 = y.shape[:-1], y.shape[-1]
    v = y.reshape(*lead, n2 // 16, 2, 8)
    return v[..., 0, :].reshape(*lead, n2 // 2), v[..., 1, :].reshape(*lead, n2 // 2)


class PackedLinear(nn.Module):
    """Inference-only linear over a packed entry: fused activation quantise + W6Ax GEMM (+ bias)."""

    def __init__(self, entry: dict, act_round: int = capi.ROUND_CUDA):
        super().__init__()
        self.register_buffer("w6", entry["w6"])
        self.register_buffer("w_scale", entry["w_scale"])
        self.bias = entry.get("bias")
        self.out_features, self.in_features, self.x_bits = entry["N"], entry["K"], entry["x_bits"]
        self.act_round = act_round
        self._ws = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise capi.FlexQError("PackedLinear needs CUDA tensors (no CPU fallback)")
        lead = x.shape[:-1]
        x2 = x.reshape(-1, self.in_features).half().contiguous()
        ws = capi.stream_workspace(x2.shape[0], self.in_features, x.device)
        y = capi.linear_w6ax(x2, self.w6, self.w_scale, self.out_features, self.x_bits, ws, self.act_round)
        if self.bias is not None:
            y = y + self.bias.to(y.dtype)
        return y.reshape(*lead, self.out_features).to(x.dtype)


def load_into(model: nn.Module, packed: dict) -> nn.Module:
    """Replace the modules named in ``packed`` by PackedLinear (the fp16 weights are dropped)."""
    for qual, entry in packed.items():
        parent = model
        parts = qual.split(".")
        for p in parts[:-1]:
            parent = parent[int(p)] if p.isdigit() else getattr(parent, p)
        if parts[-1].isdigit():
            parent[int(parts[-1])] = PackedLinear(entry)
        else:
            setattr(parent, parts[-1], PackedLinear(entry))
    return model
