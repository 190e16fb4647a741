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
        # True: SiLU(gate) * up is applied by the gate_up GEMM's epilogue (flexq_gemm_w6ax_silu_mul, weights packed in the
        # interleaved row order) and only [M, inter] fp16 reaches HBM; False: plain gate_up GEMM, then one SiLU*up+quantise pass
        self.fuse_silu_epilogue = True

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
        ver = (g.weight.data_ptr(), g.weight._version, u.weight.data_ptr(), u.weight._version, self.fuse_silu_epilogue)
        if self._fused is None or self._fused[2] != ver:
            # per-row-group quantisation: packing the row-concatenated weight == concatenating the packed halves
            if self.fuse_silu_epilogue:
                from .model_pack import interleave_gate_up
                w = interleave_gate_up(g.weight, u.weight).contiguous()
            else:
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
        self._ws = capi.stream_workspace(device=x2.device)
        xq, sx = capi.quant_act(x2, self.gate_proj.act_quantizer.n_bits, self.gate_proj.act_round)
        if self.fuse_silu_epilogue:
            h = capi.gemm_w6ax_silu_mul(xq, sx, w6_gu, ws_gu, inter, self._ws)
            hq, sh = capi.quant_act(h, self.down_proj.act_quantizer.n_bits, capi.ROUND_CUDA)
        else:
            gu = capi.gemm_w6ax(xq, sx, w6_gu, ws_gu, 2 * inter, self._ws)
            hq, sh, h = capi.silu_mul_quant(gu[:, :inter], gu[:, inter:], self.down_proj.act_quantizer.n_bits, want_out=True)
        y = capi.gemm_w6ax(hq, sh, w6_d, ws_d, hid, self._ws)
        if self.down_proj.bias is not None:
            y = y + self.down_proj.bias.to(y.dtype)
        return y.reshape(*lead, hid).to(x.dtype), h.reshape(*lead, inter).to(x.dtype)


class QuantLlamaAttention(nn.Module):
    """Drop-in for the reference's ``QuantLlamaAttention`` (/root/reference/algorithm/models/int_llama_layer.py:53-200):
    same constructor ``(org_module, config, args)``, the four ``QuantLinear`` sub-modules under the same names and
    ``set_quant_state``.  In the FlexQ configuration the two attention matmuls are 16-bit pass-throughs (``QuantMatMul``
    with n_bits = 16, main.py:267-296), so only the projections are quantised:

        x --quantise(A6) once--> [q;k;v] W6A6 GEMM (one launch, N = (heads + 2 kv_heads) * head_dim)
          --rotary, KV cache, attention (the org module's own attention function)--> o_proj W6A6 --> fp16

    ``org_module`` is a ``transformers`` LlamaAttention (forward signature of transformers >= 4.54: rotary embeddings come in
    as ``position_embeddings``, the cache as ``past_key_values``); returns ``(attn_output, attn_weights)`` like it.
    """

    def __init__(self, org_module: nn.Module, config, args=None):
        super().__init__()
        self.config = config
        self.layer_idx = getattr(org_module, "layer_idx", None)
        self.hidden_size = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.head_dim = getattr(org_module, "head_dim", self.hidden_size // self.num_heads)
        self.num_key_value_heads = config.num_key_value_heads
        self.num_key_value_groups = self.num_heads // self.num_key_value_heads
        self.scaling = getattr(org_module, "scaling", self.head_dim ** -0.5)
        self.attention_dropout = getattr(org_module, "attention_dropout", 0.0)
        self.is_causal = True
        wq, aq = args.weight_quant_params, args.act_quant_params
        self.k_proj = QuantLinear(org_module.k_proj, wq, aq)
        self.v_proj = QuantLinear(org_module.v_proj, wq, aq)
        self.q_proj = QuantLinear(org_module.q_proj, wq, aq)
        self.o_proj = QuantLinear(org_module.o_proj, wq, aq)
        self.use_weight_quant = False
        self.use_act_quant = False
        self._fused = None

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant, self.use_act_quant = weight_quant, act_quant
        for m in (self.q_proj, self.k_proj, self.v_proj, self.o_proj):
            m.set_quant_state(weight_quant, act_quant)

    def _fusable(self, x: torch.Tensor) -> bool:
        ps = (self.q_proj, self.k_proj, self.v_proj)
        return (x.is_cuda and all(p.kernel_supported() and p.bias is None for p in ps) and x.dtype != torch.float32
                and len({p.act_quantizer.n_bits for p in ps}) == 1)

    @torch.no_grad()
    def _pack_qkv(self):
        ps = (self.q_proj, self.k_proj, self.v_proj)
        ver = tuple(v for p in ps for v in (p.weight.data_ptr(), p.weight._version))
        if self._fused is None or self._fused[2] != ver:
            w = torch.cat([p.weight for p in ps], 0).contiguous()        # per-(row, group) quantisation: same as packing each
            w = w if w.dtype in (torch.float16, torch.float32) else w.float()
            w6, ws = capi.quant_pack_w6(w)
            self._fused = (w6, ws, ver)
        return self._fused[0], self._fused[1]

    def _qkv(self, hidden_states: torch.Tensor):
        if not self._fusable(hidden_states):
            return self.q_proj(hidden_states), self.k_proj(hidden_states), self.v_proj(hidden_states)
        lead = hidden_states.shape[:-1]
        x2 = hidden_states.reshape(-1, self.hidden_size)
        x2 = (x2 if x2.dtype == torch.float16 else x2.half()).contiguous()
        nq, nkv = self.num_heads * self.head_dim, self.num_key_value_heads * self.head_dim
        w6, ws = self._pack_qkv()
        xq, sx = capi.quant_act(x2, self.q_proj.act_quantizer.n_bits, self.q_proj.act_round)
        qkv = capi.gemm_w6ax(xq, sx, w6, ws, nq + 2 * nkv, capi.stream_workspace(device=x2.device)).to(hidden_states.dtype)
        q, k, v = qkv[:, :nq], qkv[:, nq:nq + nkv], qkv[:, nq + nkv:]
        return q.reshape(*lead, nq), k.reshape(*lead, nkv), v.reshape(*lead, nkv)

    def forward(self, hidden_states: torch.Tensor, position_embeddings=None, attention_mask=None, past_key_values=None, **kwargs):
        from transformers.models.llama.modeling_llama import ALL_ATTENTION_FUNCTIONS, apply_rotary_pos_emb, eager_attention_forward
        input_shape = hidden_states.shape[:-1]
        hidden_shape = (*input_shape, -1, self.head_dim)
        q, k, v = self._qkv(hidden_states)
        query_states = q.reshape(hidden_shape).transpose(1, 2)
        key_states = k.reshape(hidden_shape).transpose(1, 2)
        value_states = v.reshape(hidden_shape).transpose(1, 2)
        cos, sin = position_embeddings
        query_states, key_states = apply_rotary_pos_emb(query_states, key_states, cos, sin)
        if past_key_values is not None:
            key_states, value_states = past_key_values.update(key_states, value_states, self.layer_idx)
        attn_fn = ALL_ATTENTION_FUNCTIONS.get_interface(self.config._attn_implementation, eager_attention_forward)
        attn_output, attn_weights = attn_fn(self, query_states, key_states, value_states, attention_mask,
                                            dropout=0.0 if not self.training else self.attention_dropout, scaling=self.scaling, **kwargs)
        attn_output = attn_output.reshape(*input_shape, -1).contiguous()
        return self.o_proj(attn_output), attn_weights


class QuantLlamaDecoderLayer(nn.Module):
    """Drop-in for the reference's ``QuantLlamaDecoderLayer`` (int_llama_layer.py:203-330): same constructor
    ``(config, ori_layer, args)``, sub-modules ``self_attn`` / ``mlp`` / ``input_layernorm`` / ``post_attention_layernorm``,
    ``set_quant_state`` and ``weight_quant_inplace``; forward signature and return value of the ``transformers`` decoder layer
    it replaces (>= 4.54: returns the hidden states)."""

    def __init__(self, config, ori_layer: nn.Module, args):
        super().__init__()
        self.hidden_size = config.hidden_size
        self.self_attn = QuantLlamaAttention(org_module=ori_layer.self_attn, config=config, args=args)
        self.mlp = QuantLlamaMLP(org_module=ori_layer.mlp, hidden_size=self.hidden_size, intermediate_size=config.intermediate_size,
                                 hidden_act=config.hidden_act, args=args)
        self.input_layernorm = ori_layer.input_layernorm
        self.post_attention_layernorm = ori_layer.post_attention_layernorm

    def forward(self, hidden_states: torch.Tensor, attention_mask=None, position_ids=None, past_key_values=None, use_cache=False,
                position_embeddings=None, **kwargs):
        residual = hidden_states
        hidden_states = self.input_layernorm(hidden_states)
        hidden_states, _ = self.self_attn(hidden_states=hidden_states, attention_mask=attention_mask, position_ids=position_ids,
                                          past_key_values=past_key_values, use_cache=use_cache, position_embeddings=position_embeddings,
                                          **kwargs)
        hidden_states = residual + hidden_states
        residual = hidden_states
        hidden_states = self.post_attention_layernorm(hidden_states)
        hidden_states, _ = self.mlp(hidden_states)
        return residual + hidden_states

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant, self.use_act_quant = weight_quant, act_quant
        for m in self.modules():
            if isinstance(m, QuantLinear):
                m.set_quant_state(weight_quant, act_quant)

    @torch.no_grad()
    def weight_quant_inplace(self):
        for m in self.modules():
            if isinstance(m, QuantLinear):
                m.weight = m.weight_quantizer(m.weight)
                m.use_temporary_parameter = False


def quantize_llama(model: nn.Module, args=None) -> nn.Module:
    """The layer walk of the reference's ``flexqllm`` (algorithm/flexq_quantize/flexqllm.py:48-122) without calibration
    data (FlexQ's W6Ax configuration is calibration free: dynamic per-group scales): every ``model.model.layers[i]`` becomes a
    ``QuantLlamaDecoderLayer`` with weight and activation quantisation on.  ``args`` carries ``weight_quant_params`` /
    ``act_quant_params`` / ``act_down_proj_quant_params`` / ``flex_linear_quant`` as main.py:256-296 builds them (defaults:
    W6, A6, down_proj A8, symmetric, group 128)."""
    import types
    from . import model_pack
    if args is None:
        args = types.SimpleNamespace(weight_quant_params=model_pack.default_quant_params(6, True),
                                     act_quant_params=model_pack.default_quant_params(6, False),
                                     act_down_proj_quant_params=model_pack.default_quant_params(8, False), flex_linear_quant=True)
    layers = model.model.layers
    for i in range(len(layers)):
        q = QuantLlamaDecoderLayer(model.config, layers[i], args)
        q.set_quant_state(weight_quant=True, act_quant=True)
        layers[i] = q
    return model
