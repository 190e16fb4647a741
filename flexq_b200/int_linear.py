"""``QuantLinear``: drop-in for the reference's quantized linear module, backed by the
sm_100a W6Ax kernels.

Interface parity with /root/reference/algorithm/flexq_quantize/int_linear.py:20-76: same
constructor, buffers (``weight``, ``bias``), attributes and ``set_quant_state``.  Behaviour:

* quant state (False, False): plain ``F.linear`` like the reference.
* quant state (True, True) with a kernel-supported configuration (symmetric, group 128, W6 and
  A6/A8 -- what ``--wbits 6 --abits 6 --w_group_size 128 --a_group_size 128 --symmetric
  [--flex_linear_quant]`` builds, algorithm/main.py:223-296): the weight is quantised + packed
  ONCE into the W6 tile layout (the reference re-fake-quantises it on every forward,
  int_linear.py:60-62) and ``forward`` runs the fused CUDA path
  activation quantise -> tcgen05 int8 GEMM with per-group scales -> fp16.
  There is no CPU fallback: CPU tensors or a missing extension raise.
* any other state/configuration with ``fake_quant_fallback=True`` reproduces the reference's
  fake-quant arithmetic with torch ops (accuracy-evaluation mode), otherwise raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import capi
from .quantizer import UniformAffineQuantizer


class QuantLinear(nn.Module):
    def __init__(self, org_module: nn.Linear, weight_quant_params: dict = {}, act_quant_params: dict = {},
                 disable_input_quant: bool = False, fake_quant_fallback: bool = False, act_round: int = capi.ROUND_CUDA):
        super().__init__()
        self.fwd_kwargs = dict()
        self.fwd_func = F.linear
        self.register_buffer("weight", org_module.weight)
        if org_module.bias is not None:
            self.register_buffer("bias", org_module.bias)
        else:
            self.bias = None
        self.in_features = org_module.in_features
        self.out_features = org_module.out_features
        self.use_weight_quant = False
        self.use_act_quant = False
        self.weight_quantizer = UniformAffineQuantizer(**weight_quant_params, shape=org_module.weight.shape)
        self.act_quantizer = None if disable_input_quant else UniformAffineQuantizer(**act_quant_params)
        self.disable_input_quant = disable_input_quant
        self.use_temporary_parameter = False
        self.fake_quant_fallback = fake_quant_fallback
        self.act_round = act_round
        # packed state (built lazily by pack_weights())
        self.w6 = None
        self.w_scale = None
        self._workspace = None
        self._packed_version = None

    # ---- reference API --------------------------------------------------------------------
    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant

    # ---- offline packing ------------------------------------------------------------------
    def kernel_supported(self) -> bool:
        return (self.use_weight_quant and self.use_act_quant and not self.disable_input_quant
                and not self.use_temporary_parameter
                and self.weight_quantizer.is_flexq_kernel_config() and self.weight_quantizer.n_bits == 6
                and self.act_quantizer is not None and self.act_quantizer.is_flexq_kernel_config()
                and self.in_features % capi.GROUP == 0)

    @torch.no_grad()
    def pack_weights(self):
        """Quantise (UniformAffineQuantizer semantics, input dtype arithmetic) and pack the
        weight into W6 tiles + fp16 group scales on the GPU.  Idempotent per weight version."""
        w = self.weight
        if not w.is_cuda:
            raise capi.FlexQError("QuantLinear real-quant path needs CUDA tensors (no CPU fallback)")
        ver = (w.data_ptr(), w._version, w.dtype)
        if self.w6 is None or self._packed_version != ver:
            wsrc = w.contiguous() if w.dtype in (torch.float16, torch.float32) else w.float().contiguous()
            self.w6, self.w_scale = capi.quant_pack_w6(wsrc)
            self._packed_version = ver
        return self.w6, self.w_scale

    def _get_workspace(self, M: int, device) -> torch.Tensor:
        # one workspace per (device, stream), shared by all modules: the partial-sum scratch alone is ~47 MB
        return capi.stream_workspace(M, self.in_features, device)

    # ---- forward ----------------------------------------------------------------------------
    def forward(self, input: torch.Tensor):
        if self.kernel_supported():
            return self._forward_kernel(input)
        quantised = self.use_weight_quant or (self.use_act_quant and not self.disable_input_quant)
        if quantised and not self.fake_quant_fallback and not self.use_temporary_parameter:
            raise capi.FlexQError(
                "this QuantLinear configuration is not implemented by the sm_100a kernels "
                "(need symmetric, group_size 128, W6 with A6/A8, both quant states on); pass "
                "fake_quant_fallback=True for the reference's torch fake-quant evaluation mode")
        return self._forward_fake(input)

    def _forward_kernel(self, x: torch.Tensor):
        if not x.is_cuda:
            raise capi.FlexQError("QuantLinear real-quant path needs CUDA tensors (no CPU fallback)")
        w6, w_scale = self.pack_weights()
        lead = x.shape[:-1]
        x2 = x.reshape(-1, self.in_features)
        M = x2.shape[0]
        if x2.dtype == torch.float32:
            # fp32 module (the reference's CPU-runnable configuration): the quantiser runs in fp32 arithmetic like
            # torch does for float tensors, the integers and the fp32 activation scales are the reference's
            xq, sx = capi.quant_act(x2.contiguous(), self.act_quantizer.n_bits)
            y = capi.gemm_w6ax(xq, sx, w6, w_scale, self.out_features, self._get_workspace(M, x2.device))
        else:
            if x2.dtype != torch.float16:
                x2 = x2.half()
            x2 = x2.contiguous()
            ws = self._get_workspace(M, x2.device)
            y = capi.linear_w6ax(x2, w6, w_scale, self.out_features, self.act_quantizer.n_bits, ws, self.act_round)
        if self.bias is not None:
            y = y + self.bias.to(y.dtype)
        return y.reshape(*lead, self.out_features).to(x.dtype)

    def _forward_fake(self, input: torch.Tensor):
        # reference arithmetic, int_linear.py:56-72
        if self.use_temporary_parameter:
            weight, bias = self.temp_weight, self.temp_bias
        elif self.use_weight_quant:
            weight, bias = self.weight_quantizer(self.weight), self.bias
        else:
            weight, bias = self.weight, self.bias
        if self.use_act_quant and not self.disable_input_quant:
            input = self.act_quantizer(input)
        return self.fwd_func(input, weight, bias, **self.fwd_kwargs)
