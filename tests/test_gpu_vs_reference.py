"""GPU tests against the reference's OWN kernels, compiled from their sources under
/root/reference into oracle/_ref (oracle/Makefile `make ref`; the .so files travel to the GPU
box).  On sm_100a ptxas emulates the reference's binary MMAs with IMMA, so they run -- slowly --
and serve as a GPU-side oracle: identical inputs -> identical packed words / scales / integer
sums, fp16 outputs within the stated tolerance."""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def _load(name):
    path = os.path.join(REF_DIR, name)
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (run `make -C oracle ref` where /root/reference exists)")
    return ctypes.CDLL(path)


@pytest.fixture(scope="module")
def capi():
    from flexq_b200 import capi as c
    c.load()
    return c


@pytest.fixture(scope="module")
def ref_engine():
    return _load("libflexq_ref_engine.so")


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


@pytest.mark.parametrize("R,K,bits", [(1, 128, 6), (4, 512, 8), (8, 4096, 6), (64, 1024, 6), (256, 512, 6)])
def test_packer_word_exact_vs_reference_kernel(capi, ref_engine, oracle, R, K, bits):
    rng = np.random.default_rng(R + K)
    raw = torch.from_numpy(rng.integers(0, 1 << bits, size=(R, K)).astype(np.int32)).cuda()
    ref = torch.zeros(R * K * bits // 32, dtype=torch.int32, device="cuda")
    assert ref_engine.ref_flexq_bit_packing_i32(_p(raw), _p(ref), R, K, bits, None) == 0
    torch.cuda.synchronize()
    ours = capi.bit_packing_i32(raw, bits)
    assert torch.equal(ours, ref)
    # and the oracle's restatement is what the reference kernel really does
    assert np.array_equal(ref.cpu().numpy().view(np.uint32), oracle.pack_planes(raw.cpu().numpy(), bits))


@pytest.mark.parametrize("M,K,bits", [(1, 4096, 6), (4, 1024, 8), (8, 4096, 6), (16, 8192, 8), (64, 1024, 6)])
def test_fused_act_quant_bit_exact_vs_reference_kernel(capi, oracle, M, K, bits):
    ref_e2e = _load("libflexq_ref_e2e.so")            # IEEE-division build of the reference source
    rng = np.random.default_rng(M * K + bits)
    xn = rng.standard_normal((M, K)).astype(np.float16)
    xn[0, :128] = 0
    x = torch.from_numpy(xn).cuda()
    planes = torch.zeros(M * K * bits // 32, dtype=torch.int32, device="cuda")
    xs = torch.zeros(K // 128, 2 * capi.ceil4(M), dtype=torch.float16, device="cuda")
    assert ref_e2e.ref_e2e_quant_pack_f16(_p(x), _p(planes), _p(xs), M, K, bits, None) == 0
    torch.cuda.synchronize()
    ours_planes, ours_xs = capi.bit_packing_f16(x, bits)
    assert torch.equal(ours_planes, planes)
    assert torch.equal(ours_xs[:, :2 * M], xs[:, :2 * M])
    # native int8 path carries the same integers and scales
    xq, sx = capi.quant_act(x, bits, capi.ROUND_CUDA)
    ints = oracle.unpack_planes(planes.cpu().numpy(), M, K, bits)
    assert np.array_equal(xq.cpu().numpy().astype(np.int32), ints)
    assert np.array_equal(sx.cpu().numpy()[:, :M], xs.cpu().numpy()[:, 0:2 * M:2].astype(np.float32))
    # the numpy/C oracle is bit-identical to the reference kernel
    q_ref, s_ref = oracle.quant_act_cuda(xn, bits)
    assert np.array_equal(ints, q_ref)


def test_fused_act_quant_vs_reference_fastmath_build(capi, oracle):
    """The reference's Release flags use --use_fast_math (approximate division): ints may differ
    from the IEEE build only at near-ties and by one step."""
    ref_fm = _load("libflexq_ref_e2e_fastmath.so")
    M, K, bits = 16, 4096, 6
    x = torch.from_numpy(np.random.default_rng(0).standard_normal((M, K)).astype(np.float16)).cuda()
    planes = torch.zeros(M * K * bits // 32, dtype=torch.int32, device="cuda")
    xs = torch.zeros(K // 128, 2 * capi.ceil4(M), dtype=torch.float16, device="cuda")
    assert ref_fm.ref_e2e_quant_pack_f16(_p(x), _p(planes), _p(xs), M, K, bits, None) == 0
    torch.cuda.synchronize()
    xq, sx = capi.quant_act(x, bits, capi.ROUND_CUDA)
    ints = oracle.unpack_planes(planes.cpu().numpy(), M, K, bits)
    d = np.abs(xq.cpu().numpy().astype(np.int32) - ints)
    assert d.max() <= 1 and (d != 0).mean() < 1e-3
    assert np.array_equal(sx.cpu().numpy()[:, :M], xs.cpu().numpy()[:, 0:2 * M:2].astype(np.float32))


def _run_ref_gemm(ref_engine, capi, xq, wq, xs, sw, M, N, K, xb):
    xp = capi.bit_packing_i32(torch.from_numpy(xq.astype(np.int32)).cuda(), xb)
    wp = capi.bit_packing_i32(torch.from_numpy(wq.astype(np.int32)).cuda(), 6)
    d = torch.zeros(M, N, dtype=torch.float16, device="cuda")
    rc = ref_engine.ref_fqbmma_gemm(_p(xp), _p(wp), _p(xs), _p(sw), M, N, K, _p(d), xb, None)
    assert rc == 0, rc
    torch.cuda.synchronize()
    return xp, wp, d


@pytest.mark.parametrize("M,xb", [(1, 6), (2, 8), (4, 6), (8, 6), (8, 8)])
def test_int32_group_sums_vs_reference_gemm_kernel(capi, ref_engine, oracle, M, xb):
    """Reference BMMA kernel with all scales 1.0, one k-group and |S| <= 2048: its fp16 output IS
    the INT32 group sum (fp16 holds integers up to 2048 exactly) -> bit-exact comparison."""
    N, K = 256, 128
    rng = np.random.default_rng(M * 10 + xb)
    xq = rng.integers(-4, 5, size=(M, K)).astype(np.int8)
    wq = rng.integers(-4, 5, size=(N, K)).astype(np.int8)
    S_ref = oracle.group_sums(xq.astype(np.int32), wq.astype(np.int32))
    assert np.abs(S_ref).max() <= 2048
    xs = torch.ones(1, 2 * capi.ceil4(M), dtype=torch.float16, device="cuda")
    sw = torch.ones(1, N, dtype=torch.float16, device="cuda")
    _, _, d = _run_ref_gemm(ref_engine, capi, xq, wq, xs, sw, M, N, K, xb)
    w6 = capi.pack_w6(torch.from_numpy(wq).cuda())
    S = capi.gemm_w6ax_groupsums(torch.from_numpy(xq).cuda(), w6, N).cpu().numpy()
    assert np.array_equal(S[:, :, 0], d.cpu().numpy().astype(np.int32))
    assert np.array_equal(S, S_ref)


@pytest.mark.parametrize("M,N,K,xb", [(1, 4096, 4096, 6), (4, 512, 1024, 8), (8, 4096, 4096, 6), (8, 1024, 11008 // 128 * 128, 8)])
def test_fp16_output_vs_reference_gemm_kernel(capi, ref_engine, oracle, M, N, K, xb):
    """Same packed operands and scales through the reference kernel and through ours (drop-in
    entry on the reference layouts): fp16 outputs agree within the stated tolerance; the
    reference's own check is |d| <= 1e-4*65504 (engine/test/test_kernel.h:59-69)."""
    rng = np.random.default_rng(M + N + K)
    xq = rng.integers(-(1 << (xb - 1)), 1 << (xb - 1), size=(M, K)).astype(np.int8)
    wq = rng.integers(-32, 32, size=(N, K)).astype(np.int8)
    sx = (rng.random((M, K // 128)) * 0.1).astype(np.float16)          # harness-style scales
    sw_np = (rng.random((K // 128, N)) * 0.1).astype(np.float16)
    xs = torch.from_numpy(oracle.x_scale_layout(sx)).cuda()
    sw = torch.from_numpy(sw_np).cuda()
    xp, wp, d_ref = _run_ref_gemm(ref_engine, capi, xq, wq, xs, sw, M, N, K, xb)
    w6 = capi.planes_to_w6(wp, N, K)
    ws = capi.new_workspace(M, K)
    d = capi.gemm_ref_layout(xp, xs, w6, sw, M, N, K, xb, ws)
    a, b = d.float().cpu().numpy().astype(np.float64), d_ref.float().cpu().numpy().astype(np.float64)
    assert np.abs(a - b).max() <= 1e-4 * 65504                         # the reference's own bar
    rms = np.sqrt(np.mean((a - b) ** 2)) / np.sqrt(np.mean(b ** 2))
    assert rms <= 1e-3, rms
    exact = oracle.gemm_exact(oracle.group_sums(xq.astype(np.int32), wq.astype(np.int32)), sx, sw_np)
    assert np.sqrt(np.mean((a - exact) ** 2)) <= np.sqrt(np.mean((b - exact) ** 2)) * 1.5 + 1e-6
