"""Generate golden vectors by running the REFERENCE's own python code.

Run in the build container only (needs /root/reference, read-only):
    python tests/golden/gen_golden.py
Writes small .npz fixtures next to this script.  The fixtures pin oracle/flexq_oracle.py
(and through it the CUDA path) to the reference's algorithm/flexq_quantize package:
  * UniformAffineQuantizer (quantizer.py:37-176) scales / ints / dequantised values,
  * QuantLinear.forward (int_linear.py:56-72) outputs,
with the parameter dicts main.py:256-296 builds for
``--wbits 6 --abits {6,8} --w_group_size 128 --a_group_size 128 --symmetric``.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/algorithm")
from flexq_quantize.int_linear import QuantLinear            # noqa: E402
from flexq_quantize.quantizer import UniformAffineQuantizer  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def params(n_bits, axes):
    # main.py:256-296 with --symmetric (=> disable_zero_point, main.py:223-224), group 128
    return dict(n_bits=n_bits, per_channel_axes=axes, symmetric=True, dynamic_method="per_group",
                group_size=128, disable_zero_point=True)


def quantizer_case(seed, rows, K, bits, dtype, kind):
    g = torch.Generator().manual_seed(seed)
    if kind == "randn":
        x = torch.randn(rows, K, generator=g)
    elif kind == "weight":
        x = 0.02 * torch.randn(rows, K, generator=g)
    elif kind == "edge":
        x = torch.randn(rows, K, generator=g)
        x[0, :128] = 0.0                      # all-zero group -> scale clamps to CLIPMIN
        x[1, 128:256] = 1e-7                  # tiny group
        x[2, :128] = torch.arange(128) - 63.5  # exact .5 ties after scaling
        x[3, 5] = 300.0                       # outlier
    x = x.to(dtype)
    q = UniformAffineQuantizer(**params(bits, []))
    deq = q(x.clone())
    scale = q.scale.reshape(rows, K // 128)
    xi = torch.clamp(torch.round(x.reshape(-1, 128) / q.scale), q.qmin, q.qmax).reshape(rows, K)
    return dict(x=x.numpy(), deq=deq.numpy(), scale=scale.numpy(), xint=xi.to(torch.int32).numpy(),
                bits=np.int32(bits))


def linear_case(seed, M, N, K, abits, dtype):
    torch.manual_seed(seed)
    lin = torch.nn.Linear(K, N, bias=False)
    x = torch.randn(M, K)
    lin = lin.to(dtype)
    x = x.to(dtype)
    ql = QuantLinear(lin, params(6, [0]), params(abits, []))
    ql.set_quant_state(True, True)
    with torch.no_grad():
        y = ql(x)
        wdeq = ql.weight_quantizer(ql.weight)
    return dict(x=x.numpy(), w=lin.weight.detach().numpy(), y=y.numpy(), wdeq=wdeq.numpy(),
                wscale=ql.weight_quantizer.scale.reshape(N, K // 128).numpy(), abits=np.int32(abits))


def main():
    torch.set_num_threads(1)
    out = {}
    for name, args in {
        "q_a6_f32_randn": (1, 7, 384, 6, torch.float32, "randn"),
        "q_a8_f32_randn": (2, 7, 384, 8, torch.float32, "randn"),
        "q_w6_f32_weight": (3, 24, 256, 6, torch.float32, "weight"),
        "q_a6_f32_edge": (4, 4, 256, 6, torch.float32, "edge"),
        "q_a8_f32_edge": (5, 4, 256, 8, torch.float32, "edge"),
        "q_a6_f16_randn": (6, 7, 384, 6, torch.float16, "randn"),
        "q_a8_f16_randn": (7, 7, 384, 8, torch.float16, "randn"),
        "q_w6_f16_weight": (8, 24, 256, 6, torch.float16, "weight"),
        "q_a6_f16_edge": (9, 4, 256, 6, torch.float16, "edge"),
    }.items():
        for k, v in quantizer_case(*args).items():
            out[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "quantizer_golden.npz"), **out)

    out = {}
    for name, args in {
        "lin_w6a6_f32": (11, 5, 40, 256, 6, torch.float32),
        "lin_w6a8_f32": (12, 5, 40, 256, 8, torch.float32),
        "lin_w6a6_f16": (13, 16, 64, 384, 6, torch.float16),
        "lin_w6a8_f16": (14, 16, 64, 384, 8, torch.float16),
    }.items():
        for k, v in linear_case(*args).items():
            out[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "linear_golden.npz"), **out)

    # BASELINE.json configs[0]: nn.Linear(4096, 4096, bias=False) default init under
    # torch.manual_seed(0), x = randn(16, 4096), fp32 on the CPU through the reference's QuantLinear.  Only the output
    # is stored (fp16-rounded to keep the fixture small would lose the point: fp32, 256 KB); the test regenerates w and x
    # from the same seed and checks their checksums.
    torch.manual_seed(0)
    lin = torch.nn.Linear(4096, 4096, bias=False)
    x = torch.randn(16, 4096)
    ql = QuantLinear(lin, params(6, [0]), params(6, []))
    ql.set_quant_state(True, True)
    with torch.no_grad():
        y = ql(x)
    np.savez_compressed(os.path.join(HERE, "c1_golden.npz"), y=y.numpy(), w_sum=np.float64(lin.weight.double().sum().item()),
                        w_abs_sum=np.float64(lin.weight.double().abs().sum().item()), x_sum=np.float64(x.double().sum().item()))
    print("wrote", [f for f in os.listdir(HERE) if f.endswith(".npz")])


if __name__ == "__main__":
    main()
