"""``QuantLlamaMLP``: drop-in for the reference's quantized LLaMA MLP block
(/root/reference/algorithm/models/int_llama_layer.py:16-50), running the fused sm_100a chain

    x --quantise(A6)--> [gate;up] W6A6 GEMM (one launch, shared activations, N = 2*inter)
      --SiLU(gate)*up + quantise(A8) in one kernel--> down_proj W6A8 GEMM --> fp16

Same constructor signature (``org_module`` with gate_proj / up_proj / down_proj, sizes, ``hidden_act``, ``args`` carrying
``weight_quant_params`` / ``act_quant_params`` / ``act_down_proj_quant_params`` / ``flex_linear_quant``), same return value
``(down_proj(h), h)`` where ``h = act(gate_proj(x)) * up_proj(x)``, and the three ``QuantLinear`` sub-modules are kept
(``gate_proj`` / ``up_proj`` / ``down_proj``) so the reference's model-level helpers that iterate ``isinstance(m, QuantLinear)``
still see them.  The fused path is used when all three are kernel-backed (symmetric g128 W6, A6 for gate/up, A6 or A8 for
down, SiLU, CUDA fp16 input); otherwise the block falls back to composing the three modules like the reference.
No CPU compute path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import capi
from .int_linear import QuantLinear


class QuantLlamaMLP(nn.Module):
    def __init__(self, org_module: nn.Module, hidden_size: int, intermediate_size: int, hidden_act: str = "silu", args=None):
        super().__init__()
        wq = args.weight_quant_params
        aq = args.act_quant_params
        aq_down = args.act_down_proj_quant_params if getattr(args, "flex_linear_quant", False) else aq   # int_llama_layer.py:35-37
        self.gate_proj = QuantLinear(org_module.gate_proj, wq, aq)
        self.down_proj = QuantLinear(org_module.down_proj, wq, aq_down)
        self.up_proj = QuantLinear(org_module.up_proj, wq, aq)
        self.hidden_size, self.intermediate_size, self.hidden_act = hidden_size, intermediate_size, hidden_act
        self._fused = None          # (w6 [gate;up], w_scale, versions)
        self._ws = None

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        for m in (self.gate_proj, self.up_proj, self.down_proj):
            m.set_quant_state(weight_quant, act_quant)

    def _fusable(self, x: torch.Tensor) -> bool:
        g, u, d = self.gate_proj, self.up_proj, self.down_proj
        return (self.hidden_act == "silu" and x.is_cuda and g.kernel_supported() and u.kernel_supported() and d.kernel_supported()
                and g.bias is None and u.bias is None and g.act_quantizer.n_bits == u.act_quantizer.n_bits
                and self.intermediate_size % capi.GROUP == 0)

    @torch.no_grad()
    def _pack(self):
        g, u = self.gate_proj, self.up_proj
        ver = (g.weight.data_ptr(), g.weight._version, u.weight.data_ptr(), u.weight._version)
        if self._fused is None or self._fused[2] != ver:
            # per-row-group quantisation: packing the row-concatenated weight == concatenating the packed halves
            w = torch.cat([g.weight, u.weight], 0).contiguous()
            w = w if w.dtype in (torch.float16, torch.float32) else w.float()
            w6, ws = capi.quant_pack_w6(w)
            self._fused = (w6, ws, ver)
        return self._fused[0], self._fused[1]

    def forward(self, x: torch.Tensor):
        if not self._fusable(x):
            h = F.silu(self.gate_proj(x)) * self.up_proj(x) if self.hidden_act == "silu" else None
            if h is None:
                raise capi.FlexQError(f"QuantLlamaMLP: activation {self.hidden_act!r} is not implemented")
            return self.down_proj(h), h
        inter, hid = self.intermediate_size, self.hidden_size
        lead = x.shape[:-1]
        x2 = x.reshape(-1, hid)
        x2 = (x2 if x2.dtype == torch.float16 else x2.half()).contiguous()
        M = x2.shape[0]
        w6_gu, ws_gu = self._pack()
        w6_d, ws_d = self.down_proj.pack_weights()
        if self._ws is None or self._ws.device != x2.device:
            self._ws = capi.new_workspace()
        xq, sx = capi.quant_act(x2, self.gate_proj.act_quantizer.n_bits, self.gate_proj.act_round)
        gu = capi.gemm_w6ax(xq, sx, w6_gu, ws_gu, 2 * inter, self._ws)
        hq, sh, h = capi.silu_mul_quant(gu[:, :inter], gu[:, inter:], self.down_proj.act_quantizer.n_bits, want_out=True)
        y = capi.gemm_w6ax(hq, sh, w6_d, ws_d, hid, self._ws)
        if self.down_proj.bias is not None:
            y = y + self.down_proj.bias.to(y.dtype)
        return y.reshape(*lead, hid).to(x.dtype), h.reshape(*lead, inter).to(x.dtype)
