"""CPU baseline: torch restatement of the reference's fake-quant linear (multi-threaded).

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/flexq_oracle.py header): bench.py's
``cpu_baseline`` leg and ``--impl reference`` arm time this on the GPU box's host cores, because
/root/reference itself does not travel to the box.  It is the same arithmetic as
oracle.flexq_oracle.quant_sym_python / fakequant_linear (pinned to the reference by
tests/golden), expressed with torch CPU ops so that it uses every host thread exactly like the
reference's own module would:
  QuantLinear.forward              algorithm/flexq_quantize/int_linear.py:56-72
  UniformAffineQuantizer.forward   algorithm/flexq_quantize/quantizer.py:128-171, 93-126
``faithful=True`` re-fake-quantises the weight on every call, which is what the reference does
under flexqllm (flexq_quantize/flexqllm.py:106 leaves use_weight_quant on).
"""
from __future__ import annotations

import torch

CLIPMIN = 1e-5


def fake_quant_sym(x: torch.Tensor, n_bits: int, group: int = 128) -> torch.Tensor:
    shape = x.shape
    xg = x.reshape(-1, group)
    amax = torch.maximum(xg.amax(-1, keepdim=True).abs(), xg.amin(-1, keepdim=True).abs())
    scale = (amax / (2 ** (n_bits - 1) - 1)).clamp(min=CLIPMIN, max=1e4)
    q = torch.round(xg / scale).clamp(-(2 ** (n_bits - 1)), 2 ** (n_bits - 1) - 1)
    return (q * scale).reshape(shape)


class FakeQuantLinearCPU:
    def __init__(self, weight: torch.Tensor, a_bits: int, w_bits: int = 6, faithful: bool = True):
        self.weight, self.a_bits, self.w_bits, self.faithful = weight, a_bits, w_bits, faithful
        self.wq = None if faithful else fake_quant_sym(weight, w_bits)

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        w = fake_quant_sym(self.weight, self.w_bits) if self.faithful else self.wq
        return torch.nn.functional.linear(fake_quant_sym(x, self.a_bits), w)
