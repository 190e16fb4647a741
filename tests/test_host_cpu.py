"""CPU tests of the host side: C ABI surface, python mirrors of the reference modules, TP logic
over gloo (world_size 2).  No compute calls into the CUDA library here."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_capi_library_exports_every_header_symbol():
    from flexq_b200 import capi
    lib = capi.load()
    hdr = open(os.path.join(ROOT, "include", "flexq_b200.h")).read()
    declared = set(re.findall(r"^(?:int|size_t|const char\*)\s+(flexq_[a-z0-9_]+)\s*\(", hdr, flags=re.M))
    assert declared == set(capi.EXPORTED_SYMBOLS), declared ^ set(capi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    # size helpers are host-only arithmetic
    assert lib.flexq_w6_packed_bytes(8192, 8192) == 8192 * 8192 * 6 // 8
    assert lib.flexq_w6_packed_bytes(200, 256) == 2 * 2 * 12288
    assert lib.flexq_planes_bytes(16, 4096, 6) == 16 * 4096 * 6 // 8
    assert lib.flexq_sx_ld(5) == 8 and lib.flexq_xscale_ref_halves(5, 256) == 2 * 16
    assert lib.flexq_status_string(-1).decode().startswith("bad shape")
    assert lib.flexq_version() >= 100


def test_capi_argument_validation_without_gpu():
    """error behaviour: status codes, not printf (reference: flexq_gemm_wrapper.cu:43-46)"""
    from flexq_b200 import capi
    lib = capi.load()
    assert lib.flexq_gemm_w6ax(None, None, None, None, None, 1, 128, 128, None, 0, None) == -3
    one = ctypes.c_void_p(16)
    assert lib.flexq_gemm_w6ax(one, one, one, one, one, 1, 128, 100, one, 1 << 30, None) == -1    # K % 128
    assert lib.flexq_gemm_w6ax(one, one, one, one, one, 1, 128, 128, one, 16, None) == -4         # workspace
    assert lib.flexq_quant_act(one, one, one, 4, 256, 7, 0, None) == -2                           # bits
    assert lib.flexq_bit_packing_i32(one, one, 12, 128, 6, None) == -1                            # ragged planes
    # fused producers and the peer all-reduce
    assert lib.flexq_rmsnorm_quant_f16(None, None, one, 1e-5, None, one, one, 4, 256, 6, None) == -3
    assert lib.flexq_rmsnorm_quant_f16(one, None, one, 1e-5, None, one, one, 4, 200, 6, None) == -1       # K % 128
    assert lib.flexq_rmsnorm_quant_f16(one, None, one, 1e-5, None, one, one, 4, 32768, 6, None) == -1     # K > 16384
    assert lib.flexq_rmsnorm_quant_f16(one, None, one, 1e-5, None, one, one, 4, 256, 5, None) == -2
    assert lib.flexq_silu_mul_quant_f16(one, None, 256, None, one, one, 4, 256, 8, None) == -3
    assert lib.flexq_silu_mul_quant_f16(one, one, 128, None, one, one, 4, 256, 8, None) == -1             # ld_in < K
    assert lib.flexq_allreduce_sum_f16(None, None, 0, 1024, 0, 2, None) == -3
    assert lib.flexq_allreduce_sum_f16(one, None, 0, 1023, 0, 2, None) == -1                               # 16-byte vectors
    assert lib.flexq_allreduce_sum_f16(one, None, 0, 1024, 2, 2, None) == -1                               # rank >= world
    assert lib.flexq_allreduce_sum_f16(one, None, 0, 1024, 0, 1, None) == 0                                # world 1: nothing to do
    assert lib.flexq_set_sm_limit(0) == 0 and lib.flexq_set_allreduce_blocks(0) == 0


def _params(bits, axes):
    return dict(n_bits=bits, per_channel_axes=axes, symmetric=True, dynamic_method="per_group",
                group_size=128, disable_zero_point=True)


def test_quantizer_mirror_matches_reference_golden():
    from flexq_b200 import UniformAffineQuantizer
    g = np.load(os.path.join(ROOT, "tests", "golden", "quantizer_golden.npz"))
    for n in sorted({k.split("/")[0] for k in g.files}):
        x = torch.from_numpy(g[n + "/x"])
        q = UniformAffineQuantizer(**_params(int(g[n + "/bits"]), []))
        d = q(x.clone())
        assert torch.equal(d, torch.from_numpy(g[n + "/deq"])), n
        assert torch.equal(q.scale.reshape(x.shape[0], -1), torch.from_numpy(g[n + "/scale"])), n
        assert q.is_flexq_kernel_config()


def test_quantlinear_mirror_api_and_fake_mode():
    from flexq_b200 import QuantLinear, capi
    g = np.load(os.path.join(ROOT, "tests", "golden", "linear_golden.npz"))
    lin = nn.Linear(256, 40, bias=False)
    with torch.no_grad():
        lin.weight.copy_(torch.from_numpy(g["lin_w6a6_f32/w"]))
    ql = QuantLinear(lin, _params(6, [0]), _params(6, []), fake_quant_fallback=True)
    for attr in ("weight", "bias", "in_features", "out_features", "use_weight_quant", "use_act_quant",
                 "weight_quantizer", "act_quantizer", "use_temporary_parameter", "set_quant_state"):
        assert hasattr(ql, attr)
    x = torch.from_numpy(g["lin_w6a6_f32/x"])
    assert torch.equal(ql(x), lin(x))                         # quant state off == plain linear
    ql.set_quant_state(True, True)
    assert ql.kernel_supported()
    # on CPU the real-quant path refuses to run (no CPU fallback) ...
    with pytest.raises(capi.FlexQError):
        ql(x)
    # ... while the explicit fake-quant evaluation mode reproduces the reference bit for bit
    ql.set_quant_state(True, False)
    ql2 = QuantLinear(lin, _params(6, [0]), _params(6, []), fake_quant_fallback=True)
    ql2.set_quant_state(True, True)
    y = ql2._forward_fake(x)
    assert torch.allclose(y, torch.from_numpy(g["lin_w6a6_f32/y"]), rtol=0, atol=2e-6 * float(y.abs().mean()) + 1e-7)


def test_quantlinear_unsupported_config_raises():
    from flexq_b200 import QuantLinear, capi
    lin = nn.Linear(256, 16, bias=False)
    p = _params(4, [0])
    ql = QuantLinear(lin, p, _params(6, []))
    ql.set_quant_state(True, True)
    assert not ql.kernel_supported()
    with pytest.raises(capi.FlexQError):
        ql(torch.randn(2, 256))


# ---- tensor parallel host logic over gloo ----------------------------------------------------------
def _tp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from flexq_b200 import tp
    torch.manual_seed(0)
    N, K, M = 512, 1024, 4
    w = torch.randn(N, K)
    x = torch.randn(M, K)
    full = x @ w.t()
    # column parallel: shard rows of W; outputs concatenate, no communication
    wc = tp.shard_weight(w, "column", rank, world)
    assert wc.shape == (N // world, K)
    yc = x @ wc.t()
    gathered = [torch.zeros_like(yc) for _ in range(world)]
    dist.all_gather(gathered, yc)
    ok_col = torch.allclose(torch.cat(gathered, dim=1), full, atol=1e-4)
    # row parallel: shard K on 128-group boundaries; partial sums all-reduced
    wr = tp.shard_weight(w, "row", rank, world)
    xr = tp.shard_activation(x, rank, world)
    assert wr.shape == (N, K // world) and xr.shape == (M, K // world) and (K // world) % 128 == 0
    yr = tp.all_reduce_sum(xr @ wr.t())
    ok_row = torch.allclose(yr, full, atol=1e-3)
    q.put((rank, ok_col, ok_row))
    dist.destroy_process_group()


def test_tp_sharding_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_tp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok_c and ok_r for _, ok_c, ok_r in res), res


def test_tp_shard_validation():
    from flexq_b200 import tp
    w = torch.zeros(512, 1024)
    with pytest.raises(ValueError):
        tp.shard_weight(w, "row", 0, 3)          # K/128 = 8 groups not divisible by 3
    with pytest.raises(ValueError):
        tp.shard_weight(w, "diag", 0, 2)


def test_model_pack_driver_structure_and_packed_shards_cpu():
    """f1 host logic that needs no GPU: the linear swap follows the reference's bit policy
    (int_llama_layer.py:31-43) and packed entries shard on tile / group boundaries."""
    from flexq_b200 import QuantLinear, model_pack

    class MLP(nn.Module):
        def __init__(self):
            super().__init__()
            self.gate_proj, self.up_proj = nn.Linear(256, 512, bias=False), nn.Linear(256, 512, bias=False)
            self.down_proj, self.lm_head = nn.Linear(512, 256, bias=False), nn.Linear(256, 32, bias=False)

    m = model_pack.replace_linears(nn.ModuleList([MLP(), MLP()]))
    for blk in m:
        assert isinstance(blk.gate_proj, QuantLinear) and blk.gate_proj.act_quantizer.n_bits == 6
        assert blk.down_proj.act_quantizer.n_bits == 8 and blk.down_proj.weight_quantizer.n_bits == 6
        assert blk.gate_proj.use_weight_quant and blk.gate_proj.use_act_quant and blk.gate_proj.kernel_supported()
        assert isinstance(blk.lm_head, nn.Linear) and not isinstance(blk.lm_head, QuantLinear)
    m2 = model_pack.replace_linears(nn.ModuleList([MLP()]), flex_linear_quant=False)
    assert m2[0].down_proj.act_quantizer.n_bits == 6
    # reference-named model helpers (flexq_quantize/utils.py:60-63,116-131)
    model_pack.set_quant_state(m2, False, False)
    assert not m2[0].gate_proj.use_weight_quant and not m2[0].gate_proj.use_act_quant
    w_before = m2[0].gate_proj.weight.clone()
    model_pack.weight_quant_inplace(m2)
    wq = m2[0].gate_proj.weight
    assert not torch.equal(wq, w_before) and torch.equal(m2[0].gate_proj.weight_quantizer(wq.clone()), wq)   # idempotent
    model_pack.register_scales_and_zeros(m2)
    assert m2[0].gate_proj.weight_quantizer.scales.shape == (512 * 256 // 128, 1)

    N, K = 512, 384
    nt, G = N // 128, K // 128
    w6 = torch.arange(nt * G, dtype=torch.int32).repeat_interleave(12288).to(torch.uint8)       # byte = tile*G + group
    e = {"w6": w6, "w_scale": torch.arange(G * N, dtype=torch.float16).view(G, N), "N": N, "K": K, "x_bits": 6, "bias": None}
    c = model_pack.shard_packed(e, "column", 1, 2)
    assert c["N"] == 256 and c["K"] == K and c["w6"].numel() == 2 * G * 12288
    assert c["w6"].view(2, G, 12288)[:, :, 0].tolist() == [[6, 7, 8], [9, 10, 11]]
    assert torch.equal(c["w_scale"], e["w_scale"][:, 256:])
    r = model_pack.shard_packed(e, "row", 2, 3)
    assert r["K"] == 128 and r["w6"].view(nt, 1, 12288)[:, 0, 0].tolist() == [2, 5, 8, 11]
    assert torch.equal(r["w_scale"], e["w_scale"][2:3])
    with pytest.raises(ValueError):
        model_pack.shard_packed(e, "column", 0, 8)          # 512 / 8 is not a whole tile
    with pytest.raises(ValueError):
        model_pack.shard_packed(e, "row", 0, 2)             # 3 groups do not split in 2


def test_quant_llama_mlp_mirror_constructor_cpu():
    """QuantLlamaMLP keeps the reference's constructor, sub-module names and bit policy (int_llama_layer.py:16-50)."""
    import types
    from flexq_b200 import QuantLinear, QuantLlamaMLP, model_pack

    class Org(nn.Module):
        def __init__(self):
            super().__init__()
            self.gate_proj, self.up_proj, self.down_proj = nn.Linear(256, 512, bias=False), nn.Linear(256, 512, bias=False), nn.Linear(512, 256, bias=False)

    def mk(flex):
        return types.SimpleNamespace(weight_quant_params=model_pack.default_quant_params(6, True),
                                     act_quant_params=model_pack.default_quant_params(6, False),
                                     act_down_proj_quant_params=model_pack.default_quant_params(8, False), flex_linear_quant=flex)

    mlp = QuantLlamaMLP(Org(), 256, 512, "silu", mk(True))
    assert all(isinstance(m, QuantLinear) for m in (mlp.gate_proj, mlp.up_proj, mlp.down_proj))
    assert mlp.down_proj.act_quantizer.n_bits == 8 and mlp.up_proj.act_quantizer.n_bits == 6
    assert QuantLlamaMLP(Org(), 256, 512, "silu", mk(False)).down_proj.act_quantizer.n_bits == 6
    mlp.set_quant_state(True, True)
    assert mlp.gate_proj.use_weight_quant and mlp.down_proj.use_act_quant
    from flexq_b200 import capi
    with pytest.raises(capi.FlexQError):                  # real-quant path has no CPU fallback
        mlp(torch.randn(4, 256))
    mlp.set_quant_state(False, False)                     # quantisation off: plain modules, runs anywhere
    y, h = mlp(torch.randn(4, 256))
    assert y.shape == (4, 256) and h.shape == (4, 512)


@pytest.mark.parametrize("m_tiles,n_tiles,G,max_ctas", [
    (1, 64, 64, 148), (1, 224, 64, 148), (1, 2, 3, 148), (2, 32, 32, 148), (3, 7, 5, 148), (6, 64, 64, 148), (11, 224, 64, 148),
    (11, 64, 64, 148), (11, 64, 224, 148), (22, 224, 64, 148), (22, 64, 8, 140), (75, 9, 4, 148), (148, 3, 2, 148),
    (149, 3, 2, 148), (400, 5, 3, 148), (5, 1, 1, 148), (13, 28, 64, 147), (1, 32, 32, 148), (3, 32, 32, 148), (1, 8, 8, 148),
    (2, 86, 32, 148), (5, 64, 64, 140), (21, 224, 64, 148), (1, 1, 224, 148), (7, 3, 100, 148)])
def test_gemm_work_decomposition_covers_every_unit_once(m_tiles, n_tiles, G, max_ctas):
    """Host logic of the W6Ax GEMM's schedule (csrc/gemm_w6ax.cu Sched / plan_ctas), through the C ABI without a GPU:
    every (token tile, n-tile, k-group) unit is owned by exactly one CTA, CTA loads are balanced to within a couple of units (when
    token tiles do not outnumber CTAs), tiles cut by a range boundary use one fp32 slot that no other cut tile shares,
    and each CTA owns at most one slot."""
    import ctypes
    from flexq_b200 import capi
    lib = capi.load()
    n_ctas = ctypes.c_int(0)
    cap = 4096
    buf = (ctypes.c_int * (5 * cap))()
    lib.flexq_debug_schedule(m_tiles, n_tiles, G, max_ctas, 0, buf, cap, ctypes.byref(n_ctas))
    P = n_ctas.value
    assert 1 <= P <= max_ctas
    owned = np.zeros((m_tiles, n_tiles, G), dtype=np.int32)
    slot_of_tile, loads = {}, []
    for cta in range(P):
        n = lib.flexq_debug_schedule(m_tiles, n_tiles, G, max_ctas, cta, buf, cap, ctypes.byref(n_ctas))
        assert 0 < n <= cap
        seg = np.ctypeslib.as_array(buf)[:5 * n].reshape(n, 5).copy()
        loads.append(int((seg[:, 3] - seg[:, 2]).sum()))
        for mt, nt, g0, g1, slot in seg:
            assert 0 <= g0 < g1 <= G
            owned[mt, nt, g0:g1] += 1
            if slot >= 0:
                assert slot < P
                assert slot_of_tile.setdefault((mt, nt), slot) == slot, "contributors of a cut tile disagree on its slot"
            else:
                assert g0 == 0 and g1 == G
    assert (owned == 1).all()
    slots = list(slot_of_tile.values())
    assert len(slots) == len(set(slots)), "two cut tiles share an fp32 slot"
    if m_tiles <= max_ctas:
        # Ureg is rounded to a whole unit: the spare CTAs share that rounding error times the number of token tiles
        assert max(loads) - min(loads) <= max(2, 0.02 * np.mean(loads) + m_tiles / 2), (min(loads), max(loads))


def test_interleave_gate_up_layout():
    """model_pack.interleave_gate_up: 8 gate rows, their 8 up rows, ... (the row order flexq_gemm_w6ax_silu_mul expects),
    and deinterleave_gate_up is its inverse on the GEMM's output columns."""
    from flexq_b200 import model_pack
    inter, K = 40, 6
    gate = torch.arange(inter * K, dtype=torch.float32).reshape(inter, K)
    up = -gate - 1
    w = model_pack.interleave_gate_up(gate, up)
    assert w.shape == (2 * inter, K)
    for r in range(2 * inter):
        blk, j = divmod(r, 16)
        src = gate if j < 8 else up
        assert torch.equal(w[r], src[8 * blk + (j % 8)])
    y = torch.randn(3, 2 * inter)
    g, u = model_pack.deinterleave_gate_up(y)
    assert g.shape == (3, inter) and torch.equal(g[:, 8:16], y[:, 16:24]) and torch.equal(u[:, 8:16], y[:, 24:32])
    with pytest.raises(ValueError):
        model_pack.interleave_gate_up(gate[:36], up[:36])


@pytest.mark.parametrize("m_tiles,n_tiles,G,expect_ctas,expect_cut", [
    (3, 32, 32, 96, False),      # 4096x4096, M=512: whole tiles on 96 CTAs beat stream-K on 148 (cut-tile sums cost ~16 steps)
    (4, 32, 32, 128, False),     # ... M=768 (128-token tile at M=512): 128 whole-tile CTAs
    (1, 32, 32, 128, True),      # one token tile: every weight tile cut into exactly 4 runs
    (1, 64, 64, 128, True),      # 8192x8192: 2 runs per tile
    (11, 224, 64, 148, True),    # the bench shape: stream-K over all SMs
    (2, 64, 64, 128, False),     # 8192x8192, M=256
])
def test_gemm_plan_choices(m_tiles, n_tiles, G, expect_ctas, expect_cut):
    """plan_ctas (csrc/gemm_w6ax.cu) prices the aligned plan against stream-K: CTA count and whether any tile is cut, for
    the shapes the choice was measured on."""
    import ctypes
    from flexq_b200 import capi
    lib = capi.load()
    cap = 4096
    buf = (ctypes.c_int * (5 * cap))()
    n_ctas = ctypes.c_int(0)
    lib.flexq_debug_schedule(m_tiles, n_tiles, G, 148, 0, buf, cap, ctypes.byref(n_ctas))
    assert n_ctas.value == expect_ctas
    cut = False
    runs_per_cta = []
    for cta in range(n_ctas.value):
        n = lib.flexq_debug_schedule(m_tiles, n_tiles, G, 148, cta, buf, cap, ctypes.byref(n_ctas))
        segs = np.ctypeslib.as_array(buf)[:5 * n].reshape(n, 5)
        cut |= bool((segs[:, 4] >= 0).any())
        runs_per_cta.append(int((segs[:, 4] >= 0).sum()))
    assert cut == expect_cut
    if expect_cut and expect_ctas < 148:
        assert max(runs_per_cta) == 1          # aligned plan: one cut run per CTA
