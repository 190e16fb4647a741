#!/usr/bin/env python
"""Where the real-LLaMA path (quantize_llama) and the torch fake-quant model part ways: end-to-end and per module on
identical inputs (2-layer random-init transformers LLaMA)."""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformers import LlamaConfig, LlamaForCausalLM  # noqa: E402

from flexq_b200 import QuantLinear, capi, model_pack, quantize_llama  # noqa: E402


def rel(a, b):
    return ((a.float() - b.float()).pow(2).mean().sqrt() / b.float().pow(2).mean().sqrt()).item()


cfg = LlamaConfig(hidden_size=512, intermediate_size=1536, num_hidden_layers=2, num_attention_heads=8, num_key_value_heads=4,
                  vocab_size=1000, max_position_embeddings=256)
torch.manual_seed(0)
base = LlamaForCausalLM(cfg).half().cuda().eval()
ids = torch.randint(0, 1000, (1, 24), device="cuda")
with torch.no_grad():
    fake = copy.deepcopy(base)
    for parent in list(fake.modules()):
        for name, child in list(parent.named_children()):
            if isinstance(child, torch.nn.Linear) and name in model_pack.LLAMA_LINEARS:
                a = model_pack.default_quant_params(8 if name == "down_proj" else 6, False)
                q = QuantLinear(child, model_pack.default_quant_params(6, True), a, fake_quant_fallback=True)
                q.set_quant_state(True, True)
                q.kernel_supported = lambda: False
                setattr(parent, name, q)
    real = quantize_llama(copy.deepcopy(base))
    for m in real.modules():
        if isinstance(m, QuantLinear):
            m.act_round = capi.ROUND_PYTHON
    ref, out, fp = fake(ids).logits, real(ids).logits, base(ids).logits
    print(f"end to end: real vs fake {rel(out, ref):.4f}   fp16 vs fake {rel(fp, ref):.4f}   real vs fp16 {rel(out, fp):.4f}")
    # per module on identical inputs: capture the inputs the fake model's modules see
    grabbed = {}

    def hook(name):
        def f(mod, args, kwargs, output):
            grabbed[name] = (args, kwargs, output)
        return f
    hs = []
    for i, layer in enumerate(fake.model.layers):
        for n in ("self_attn", "mlp"):
            hs.append(getattr(layer, n).register_forward_hook(hook(f"{i}.{n}"), with_kwargs=True))
        for n, mod in layer.named_modules():
            if isinstance(mod, QuantLinear):
                hs.append(mod.register_forward_hook(hook(f"{i}.{n}"), with_kwargs=True))
    fake(ids)
    for h in hs:
        h.remove()
    for i, layer in enumerate(real.model.layers):
        a, kw, o = grabbed[f"{i}.self_attn"]
        r = layer.self_attn(*a, **kw)[0]
        print(f"layer {i} self_attn same input: {rel(r, o[0]):.5f}")
        a, kw, o = grabbed[f"{i}.mlp"]
        r = layer.mlp(*a, **kw)
        r = r[0] if isinstance(r, tuple) else r
        print(f"layer {i} mlp       same input: {rel(r, o):.5f}")
        for n, mod in layer.named_modules():
            if isinstance(mod, QuantLinear):
                a, kw, o = grabbed[f"{i}.{n}"]
                print(f"layer {i} {n:20s} same input: {rel(mod(*a, **kw), o):.5f}   out rms {o.float().pow(2).mean().sqrt().item():.4f}")
