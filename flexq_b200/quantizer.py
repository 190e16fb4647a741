"""Host-side mirror of the reference's ``UniformAffineQuantizer``.

Interface parity with /root/reference/algorithm/flexq_quantize/quantizer.py:37-176 (same
constructor keywords, ``forward``, ``change_n_bits``, ``register_scales_and_zeros``, the
``scale`` / ``round_zero_point`` attributes callers read).  It describes *how* a tensor is
quantised; ``QuantLinear`` hands the symmetric group-128 6/8-bit configurations to the CUDA
kernels and uses ``forward`` (plain torch, any device) only for the fake-quant evaluation
mode the reference's accuracy scripts rely on.
"""
from __future__ import annotations

import torch
import torch.nn as nn

CLIPMIN = 1e-5   # quantizer.py:24


class UniformAffineQuantizer(nn.Module):
    def __init__(self, n_bits: int = 8, symmetric: bool = False, per_channel_axes=[], metric: str = "minmax",
                 dynamic: bool = False, dynamic_method: str = "per_group", group_size=None, shape=None,
                 disable_zero_point: bool = False, flex_quant: bool = False):
        super().__init__()
        self.symmetric = symmetric
        self.disable_zero_point = disable_zero_point
        self.flex_quant = flex_quant
        self.per_channel_axes = per_channel_axes
        self.metric = metric
        self.dynamic = dynamic
        self.dynamic_method = dynamic_method
        self.group_size = group_size
        self.deficiency = 0
        self.enable = True
        self.scale = None
        self.zero_point = None
        self.round_zero_point = None
        self.change_n_bits(n_bits)

    # ---- configuration --------------------------------------------------------------------
    def change_n_bits(self, n_bits: int):
        self.n_bits = n_bits
        if self.disable_zero_point:          # signed grid (quantizer.py:58-60)
            self.qmin, self.qmax = -(2 ** (n_bits - 1)), 2 ** (n_bits - 1) - 1
        else:                                # unsigned grid with zero point (:61-63)
            self.qmin, self.qmax = 0, 2 ** n_bits - 1

    def is_flexq_kernel_config(self) -> bool:
        """True when this quantiser is one the sm_100a kernels implement: symmetric, no zero
        point, group 128, 6 or 8 bit (the reference's --symmetric --*_group_size 128 setup,
        algorithm/main.py:223-296)."""
        return (self.enable and self.symmetric and self.disable_zero_point and self.group_size == 128
                and self.n_bits in (6, 8) and self.dynamic_method in ("per_group", "per_token", "per_channel")
                and self.metric != "fix0to1")

    # ---- calibration (quantizer.py:144-171) -------------------------------------------------
    def _grouped(self, x: torch.Tensor) -> torch.Tensor:
        if not self.group_size:
            return x
        if self.deficiency > 0:
            x = torch.nn.functional.pad(x, (0, self.deficiency))
        return x.reshape(-1, self.group_size)

    def per_token_dynamic_calibration(self, x: torch.Tensor):
        xg = self._grouped(x)
        lo = xg.amin(dim=-1, keepdim=True)
        hi = xg.amax(dim=-1, keepdim=True)
        if self.symmetric:
            scale = torch.maximum(hi.abs(), lo.abs()) / (2 ** (self.n_bits - 1) - 1)
            self.scale = scale.clamp(min=CLIPMIN, max=1e4)
            zero_point = (2 ** (self.n_bits - 1) - 1) * torch.ones_like(self.scale)
        else:
            levels = 2 ** self.n_bits if self.n_bits in (1, 2) else 2 ** self.n_bits - 1
            self.scale = ((hi - lo) / levels).clamp(min=CLIPMIN, max=1e4)
            zero_point = -lo / self.scale
        self.round_zero_point = None if self.disable_zero_point else zero_point.clamp(min=-1e4, max=1e4).round()

    # ---- fake quantisation (quantizer.py:93-126) --------------------------------------------
    def fake_quant(self, x: torch.Tensor, scale: torch.Tensor, round_zero_point):
        squeezed = False
        shape = None
        if self.deficiency > 0:
            x = torch.nn.functional.pad(x, (0, self.deficiency))
        if self.group_size:
            if x.dim() == 3 and x.shape[0] == 1:
                x, squeezed = x.squeeze(0), True
            assert x.dim() == 2, "only support linear layer now"
            shape = x.shape
            x = x.reshape(-1, self.group_size)
        q = torch.round(x / scale)
        if round_zero_point is not None:
            q = q + round_zero_point
        q = q.clamp(self.qmin, self.qmax)
        if round_zero_point is not None:
            q = q - round_zero_point
        out = q * scale
        if shape is not None:
            out = out.reshape(shape)
        if self.deficiency > 0:
            out = out[:, :-self.deficiency]
        return out.unsqueeze(0) if squeezed else out

    def forward(self, x: torch.Tensor):
        if self.n_bits >= 16 or not self.enable:
            return x
        if self.metric == "fix0to1":
            levels = 2 ** self.n_bits - 1
            x = x.mul_(levels).round_().div_(levels)
            if not self.flex_quant:
                return x
        if self.dynamic_method not in ("per_token", "per_channel", "per_group"):
            raise NotImplementedError(self.dynamic_method)
        self.per_token_dynamic_calibration(x)
        return self.fake_quant(x, self.scale, self.round_zero_point)

    def register_scales_and_zeros(self):
        self.register_buffer("scales", self.scale)
        self.register_buffer("zeros", self.round_zero_point)
        del self.scale
        del self.round_zero_point
