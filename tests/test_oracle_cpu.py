"""CPU tests: the oracle against the reference's golden vectors / validator formulas, the C
restatement against the numpy one, and host-side logic.  No GPU needed."""
import ctypes
import os

import numpy as np
import pytest


def _cases(npz):
    return sorted({k.split("/")[0] for k in npz.files})


# ---- a1/a2: python quantiser pinned to the reference's own outputs --------------------------
def test_quantizer_matches_reference_golden(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "quantizer_golden.npz"))
    names = _cases(g)
    assert len(names) >= 9
    for n in names:
        x, bits = g[n + "/x"], int(g[n + "/bits"])
        q, s, d = oracle.quant_sym_python(x, bits)
        assert np.array_equal(q, g[n + "/xint"]), n
        assert np.array_equal(s, g[n + "/scale"]), n
        assert np.array_equal(d, g[n + "/deq"]), n
        assert np.abs(q).max() <= 2 ** (bits - 1) - 1 or "edge" in n


def test_fakequant_linear_matches_reference_golden(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "linear_golden.npz"))
    for n in _cases(g):
        x, w, ab = g[n + "/x"], g[n + "/w"], int(g[n + "/abits"])
        _, ws, wd = oracle.quant_sym_python(w, 6)
        assert np.array_equal(wd, g[n + "/wdeq"]) and np.array_equal(ws, g[n + "/wscale"])
        y = oracle.fakequant_linear(x, w, 6, ab).astype(np.float64)
        ref = g[n + "/y"].astype(np.float64)
        tol = 2e-6 if x.dtype == np.float32 else 2e-3      # accumulation order only
        assert np.abs(y - ref).max() <= tol * np.abs(ref).mean(), n


def test_c_quantizer_matches_golden(oracle_c, golden_dir):
    g = np.load(os.path.join(golden_dir, "quantizer_golden.npz"))
    for n in _cases(g):
        x = g[n + "/x"]
        if x.dtype != np.float32:
            continue
        rows, K = x.shape
        bits = int(g[n + "/bits"])
        q = np.zeros((rows, K), np.int32)
        s = np.zeros((rows, K // 128), np.float32)
        d = np.zeros((rows, K), np.float32)
        rc = oracle_c.fq_quant_python_f32(x.ctypes.data_as(ctypes.c_void_p), q.ctypes.data_as(ctypes.c_void_p),
                                          s.ctypes.data_as(ctypes.c_void_p), d.ctypes.data_as(ctypes.c_void_p),
                                          rows, K, bits, 128)
        assert rc == 0
        assert np.array_equal(q, g[n + "/xint"]) and np.array_equal(s, g[n + "/scale"]) and np.array_equal(d, g[n + "/deq"]), n


# ---- a5: plane layout pinned by the reference validator formula -------------------------------
@pytest.mark.parametrize("R,K,bits", [(1, 128, 6), (2, 256, 8), (4, 256, 6), (8, 384, 8), (24, 256, 6)])
def test_planes_layout_vs_abq_validator(oracle, oracle_c, R, K, bits):
    rng = np.random.default_rng(R * K + bits)
    raw = rng.integers(0, 1 << bits, size=(R, K)).astype(np.int32)
    planes = oracle.pack_planes(raw, bits)
    abq = oracle.abq_pack(raw, bits)
    # literal restatement of engine/test_packing_kernel.cu:133-146
    chunk = min(R, 8)
    for b in range(bits):
        for m in range(R):
            for k32 in range(K // 32):
                idx = (k32 // 4) * (R * bits * 4) + (m // chunk) * (bits * chunk * 4) + b * (chunk * 4) + (m % chunk) * 4 + k32 % 4
                assert abq[b, m, k32] == planes[idx]
    # bit order: element k%32==0 is bit 31 (bit_packing.cu:75)
    assert (abq[0, 0, 0] >> 31) & 1 == raw[0, 0] & 1
    # C restatement agrees word for word
    out = np.zeros_like(planes)
    assert oracle_c.fq_pack_planes(raw.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), R, K, bits) == 0
    assert np.array_equal(out, planes)
    # round trip to two's complement
    assert np.array_equal(oracle.unpack_planes(planes, R, K, bits), oracle.to_twos_complement_range(raw, bits))


def test_planes_reject_ragged_rows(oracle_c):
    x = np.zeros((12, 128), np.int32)
    o = np.zeros(12 * 128 * 6 // 32, np.uint32)
    assert oracle_c.fq_pack_planes(x.ctypes.data_as(ctypes.c_void_p), o.ctypes.data_as(ctypes.c_void_p), 12, 128, 6) == -1


# ---- a6: fused CUDA-style activation quantiser, numpy vs C ------------------------------------
@pytest.mark.parametrize("bits", [6, 8])
def test_quant_act_cuda_numpy_vs_c(oracle, oracle_c, bits):
    rng = np.random.default_rng(bits)
    M, K = 6, 384
    x = rng.standard_normal((M, K)).astype(np.float16)
    x[0, :128] = 0
    x[1, :128] = np.float16(1e-7)
    x[2, :128] = (np.arange(128) - 63.5).astype(np.float16)
    q, s = oracle.quant_act_cuda(x, bits)
    qc = np.zeros((M, K), np.int32)
    sc = np.zeros((M, K // 128), np.uint16)
    assert oracle_c.fq_quant_act(x.view(np.uint16).ctypes.data_as(ctypes.c_void_p), qc.ctypes.data_as(ctypes.c_void_p),
                                 sc.ctypes.data_as(ctypes.c_void_p), M, K, bits) == 0
    assert np.array_equal(q, qc)
    assert np.array_equal(s.view(np.uint16), sc)
    assert np.all(q[0, :128] == 0) and s[0, 0] == 0            # 0/0 -> NaN -> (int) 0
    hi = (1 << (bits - 1)) - 1
    assert q.max() <= hi and q.min() >= -hi - 1
    # exact ties round away from zero (C round()), unlike the python path
    t = x[2, :128].astype(np.float32) / s[2, 0].astype(np.float32)
    ties = np.abs(t - np.trunc(t)) == 0.5
    if ties.any():
        assert np.array_equal(q[2, :128][ties], (np.trunc(t[ties]) + np.sign(t[ties])).astype(np.int32))


# ---- a8/a11: GEMM semantics ----------------------------------------------------------------------
def _half_bits(a):
    return np.ascontiguousarray(a.astype(np.float16)).view(np.uint16)


@pytest.mark.parametrize("M,N,K,xb", [(1, 8, 128, 6), (2, 16, 256, 8), (4, 8, 256, 6), (8, 16, 128, 8)])
def test_compute_ref_literal_vs_group_sum_formula(oracle, oracle_c, M, N, K, xb):
    """The reference's CPU golden (bit-serial over packed planes) equals sum_g sx*sw*S[m,n,g] with
    S the two's-complement integer dot product: this pins the plane weights / sign rule /
    scale indexing that the INT32-sum restatement relies on."""
    rng = np.random.default_rng(M * 100 + N + K + xb)
    xr = rng.integers(0, 1 << xb, size=(M, K)).astype(np.int32)
    wr = rng.integers(0, 64, size=(N, K)).astype(np.int32)
    xp, wp = oracle.pack_planes(xr, xb), oracle.pack_planes(wr, 6)
    sx = (rng.random((M, K // 128)) * 0.1).astype(np.float16)
    sw = (rng.random((K // 128, N)) * 0.1).astype(np.float16)
    xs = oracle.x_scale_layout(sx)
    ref = np.zeros((M, N), np.uint16)
    oracle_c.fq_compute_ref(wp.ctypes.data_as(ctypes.c_void_p), _half_bits(sw).ctypes.data_as(ctypes.c_void_p),
                            xp.ctypes.data_as(ctypes.c_void_p), _half_bits(xs).ctypes.data_as(ctypes.c_void_p),
                            ref.ctypes.data_as(ctypes.c_void_p), M, N, K, 6, xb, 1, 128)
    ref = ref.view(np.float16).astype(np.float64)
    exact = oracle.compute_ref_vectorised(wp, sw, xp, xs, M, N, K, 6, xb)
    # compute_ref accumulates ~K*xb*6 terms in fp32 then rounds to half
    assert np.abs(ref - exact).max() <= 2e-3 * max(np.abs(exact).max(), 1e-6) + 1e-3
    # and the two integer restatements agree exactly
    xq, wq = oracle.to_twos_complement_range(xr, xb), oracle.to_twos_complement_range(wr, 6)
    S = oracle.group_sums(xq, wq)
    Sc = np.zeros_like(S)
    oracle_c.fq_group_sums(xq.ctypes.data_as(ctypes.c_void_p), wq.ctypes.data_as(ctypes.c_void_p),
                           Sc.ctypes.data_as(ctypes.c_void_p), M, N, K)
    assert np.array_equal(S, Sc)


def test_group_sums_extremes_exact(oracle):
    xq = np.full((2, 256), -128, np.int32)
    wq = np.full((3, 256), -32, np.int32)
    S = oracle.group_sums(xq, wq)
    assert S.shape == (2, 3, 2) and np.all(S == 128 * 128 * 32)


def test_kernel_numerics_tolerance_statement(oracle):
    """The stated FP16 tolerance (rms-rel 1e-3, max-abs 1e-2*mean|ref|) holds between the exact
    value, the reference kernel's fp16-product numerics and the fake-quant path on C1-like data."""
    rng = np.random.default_rng(0)
    M, N, K = 16, 256, 1024
    x = rng.standard_normal((M, K)).astype(np.float16)
    w = (0.02 * rng.standard_normal((N, K))).astype(np.float16)
    xq, sx, _ = oracle.quant_sym_python(x, 6)          # fp16 arithmetic, as the reference's fp16 models
    wq, sw, _ = oracle.quant_sym_python(w, 6)
    S = oracle.group_sums(xq, wq)
    exact = oracle.gemm_exact(S, sx, sw.T)
    emu = oracle.gemm_refkernel_numerics(S, sx, sw.T).astype(np.float64)
    fq = oracle.fakequant_linear(x, w, 6, 6).astype(np.float64)
    for a, b in [(emu, exact), (emu, fq), (exact, fq)]:
        rms = np.sqrt(np.mean((a - b) ** 2)) / np.sqrt(np.mean(b ** 2))
        mx = np.abs(a - b).max() / np.abs(b).mean()
        assert rms <= 1e-3 and mx <= 1e-2, (rms, mx)
    # the two activation rounding behaviours differ by at most one step
    xq2, sx2 = oracle.quant_act_cuda(x, 6)
    assert np.array_equal(sx2, sx) and np.abs(xq2 - xq).max() <= 1


# ---- native W6 layout ----------------------------------------------------------------------------
@pytest.mark.parametrize("N,K", [(128, 128), (200, 384), (8, 256)])
def test_w6_native_roundtrip_and_size(oracle, N, K):
    rng = np.random.default_rng(N + K)
    w = rng.integers(-32, 32, size=(N, K)).astype(np.int32)
    pk = oracle.pack_w6_native(w)
    assert pk.size == (N + 127) // 128 * (K // 128) * 12288
    assert np.array_equal(oracle.unpack_w6_native(pk, N, K), w)
