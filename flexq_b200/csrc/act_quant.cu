// Fused dynamic activation quantisation (north-star subsystem 2).
//
// One pass over X: per-token, per-128-group absmax (warp-shuffle reduction over the 16 lanes
// that own a group) -> scale -> 6- or 8-bit quantise -> pack.  Each thread moves one 128-bit
// vector of 8 halves; a warp covers two groups, so global reads and the int8 writes are fully
// coalesced.  Two output flavours share the arithmetic:
//   * native  : int8 containers Xq[M][K] + fp32 scales sx[G][ldsx]   (feeds the tcgen05 GEMM)
//   * planes  : the reference's bit-plane tensor + duplicated half scales (API parity with
//               flexq_bit_packing(const half*,...), /root/reference/e2e/src/fastertransformer/
//               kernels/flexqgemm/src/pack/bit_packing.cu:80-199)
// Arithmetic follows the reference line by line (mode FLEXQ_ROUND_CUDA: bit_packing.cu:119-166)
// or the python quantiser (mode FLEXQ_ROUND_PYTHON: algorithm/flexq_quantize/quantizer.py:
// 112-116,153-155 evaluated in fp16 like torch does for half tensors).
#include "common.cuh"

namespace flexq {

// quantise the 8 halves held by this thread; returns the scale actually divided by
template <int MODE>
__device__ __forceinline__ float quant8(const uint4& raw, int bits, int* q) {
    const __half* h = reinterpret_cast<const __half*>(&raw);
    float xf[8];
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        xf[i] = __half2float(h[i]);
        amax = fmaxf(amax, fabsf(xf[i]));
    }
    // group = 16 consecutive lanes
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    const int hi = (1 << (bits - 1)) - 1, lo = -(1 << (bits - 1));
    float r;
    if (MODE == FLEXQ_ROUND_CUDA) {
        const float s = __fdiv_rn(amax, (float)hi);                     // bit_packing.cu:151
        r = __half2float(__float2half_rn(s));                           // :155,:158
        // round(x / r) with IEEE division is the reference's arithmetic (:160).  x * (1/r) is within
        // 2 ulp of x / r, so it rounds to the same integer unless it lands within 1e-4 of a .5 tie;
        // only then (and for r == 0 / non-finite values) is the exact division evaluated.
        const float rcp = __frcp_rn(r);
        const bool plain = r > 0.f && rcp < 3.0e38f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float t = xf[i] * rcp;
            const float fr = fabsf(t - truncf(t));
            int v;
            if (!plain || fabsf(fr - 0.5f) < 1e-4f || !(fabsf(t) < 1e6f))
                v = __float2int_rz(roundf(__fdiv_rn(xf[i], r)));        // half away from zero; NaN -> 0, saturating
            else
                v = __float2int_rz(t + copysignf(0.5f, t));             // same integer: t is >= 1e-4 away from a tie
            q[i] = max(lo, min(hi, v));
        }
    } else {
        // torch half arithmetic = fp32 op, rounded to half after every op
        __half sh = __float2half_rn(__fdiv_rn(amax, (float)hi));        // quantizer.py:154
        const __half cmin = __float2half_rn(1e-5f), cmax = __float2half_rn(1e4f);
        if (__hlt(sh, cmin)) sh = cmin;                                  // :155
        if (__hgt(sh, cmax)) sh = cmax;
        r = __half2float(sh);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float t = __half2float(__float2half_rn(__fdiv_rn(xf[i], r)));   // :112 (x/scale in half)
            const float rr = rintf(t);                                            // half to even
            q[i] = (int)fminf(fmaxf(rr, (float)lo), (float)hi);                   // :116
        }
    }
    return r;
}

// Fast path of the CUDA-mode quantiser for one thread's 8 halves -> 8 int8 containers in two words.
// Same integers as quant8<FLEXQ_ROUND_CUDA> (the reference's round-half-away of the IEEE quotient,
// bit_packing.cu:151-160) or `fallback` is set and the caller recomputes with the exact routine:
//   * t = |x| * rcp(r) is within 2 ulp of |x| / r (<= 1 unit of 2^-16 for t < 128);
//   * a = t + 128.5 + 4*2^-16, rounded down, lies in [128, 256) where one ulp is 2^-16: byte 2 of its
//     bit pattern is floor(t + 0.5) and the low 16 bits are the fraction, shifted by 4 units;
//   * the integer can only differ from the exact one when that fraction is within [0, 16) -- then fall back;
//   * signs come back with one PRMT (sign replication) and a per-byte negate without cross-byte carries.
// Anything unusual (r == 0, denormal/inf/NaN scale, a NaN input, t that could reach hi + 0.5) also falls back.
__device__ __forceinline__ float quant8_fast(const uint4& raw, int bits, uint2& packed, bool& fallback) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    uint32_t aw[4];
#pragma unroll
    for (int i = 0; i < 4; i++) aw[i] = w[i] & 0x7FFF7FFFu;
    // |x| max; the NaN-propagating forms make a NaN input poison amax (-> fallback)
    const __half2 m2 = __hmax2_nan(__hmax2_nan(*reinterpret_cast<__half2*>(&aw[0]), *reinterpret_cast<__half2*>(&aw[1])),
                                   __hmax2_nan(*reinterpret_cast<__half2*>(&aw[2]), *reinterpret_cast<__half2*>(&aw[3])));
    uint32_t ab = __float_as_uint(__half2float(__hmax_nan(__low2half(m2), __high2half(m2))));
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) ab = max(ab, __shfl_xor_sync(0xffffffffu, ab, o));   // non-negative floats order like integers, NaN on top
    const float amax = __uint_as_float(ab);
    const float hi = (float)((1 << (bits - 1)) - 1);
    // r = half(amax / hi).  amax is a half (11-bit significand n) and hi is odd (31 / 127), so n / hi is either
    // exactly representable or at least 2^-19 (relative) away from every half rounding boundary; the product
    // with the rounded reciprocal is within 2^-23 of it and therefore rounds to the same half.
    const float r = __half2float(__float2half_rn(amax * (bits == 6 ? 1.f / 31.f : 1.f / 127.f)));
    float rcp;                                                            // <= 1 ulp off 1/r: part of the 2-ulp budget above
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(r));
    fallback = !(r > 0.f && rcp < 3.0e38f && amax * rcp < hi + 0.49f);    // false for NaN amax as well
    const float2 rcp2 = make_float2(rcp, rcp);
    constexpr float kBias = 128.5f + 4.f / 65536.f;
    uint32_t b[8];
    uint32_t tie = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float2 f = __half22float2(*reinterpret_cast<__half2*>(&aw[i]));
        const float2 t = __fmul2_rn(f, rcp2);
        b[2 * i] = __float_as_uint(__fadd_rd(t.x, kBias));
        b[2 * i + 1] = __float_as_uint(__fadd_rd(t.y, kBias));
        tie = min(tie, min(b[2 * i] & 0xFFF0u, b[2 * i + 1] & 0xFFF0u));
    }
    fallback |= tie == 0u;
    constexpr uint32_t H = 0x80808080u;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const uint32_t m01 = __byte_perm(b[4 * j], b[4 * j + 1], 0x0062);
        const uint32_t m23 = __byte_perm(b[4 * j + 2], b[4 * j + 3], 0x0062);
        const uint32_t mag = __byte_perm(m01, m23, 0x5410);                    // 4 magnitudes, each <= 127
        uint32_t sg;                                                           // 0xFF where the half is negative:
        asm("prmt.b32 %0, %1, %2, 0xFDB9;" : "=r"(sg) : "r"(w[2 * j]), "r"(w[2 * j + 1]));   // selector msb = replicate the byte's sign
        const uint32_t q = (((mag ^ sg) | H) - (sg & ~H)) ^ H;                 // per byte: sg ? -mag : mag
        if (j == 0) packed.x = q; else packed.y = q;
    }
    return r;
}

// grid = (rows of sx incl. padding, column blocks).  A warp owns U consecutive pairs of groups of one
// token row (32 lanes x 8 halves = 2 groups per pass); all U loads are issued before the arithmetic.
template <int MODE, int U>
__global__ void __launch_bounds__(256) quant_act_native_kernel(const uint4* __restrict__ x, int8_t* __restrict__ xq,
                                                               float* __restrict__ sx, int M, int K, int ldsx, int bits) {
    // programmatic dependent launch: the GEMM that consumes Xq/sx may start streaming its weights now;
    // this kernel itself waits for the producer of X (no-ops when launched without the attribute)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int m = blockIdx.x;
    const int lane16 = threadIdx.x & 15;
    const int nvec = K >> 3;                                                // 8-half vectors per row
    const int v0 = ((blockIdx.y * 8 + (threadIdx.x >> 5)) * U) * 32 + (threadIdx.x & 31);
    if (m >= M) {                                  // padding rows of sx
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int v = v0 + u * 32;
            if (v < nvec && lane16 == 0) sx[(size_t)(v >> 4) * ldsx + m] = 0.f;
        }
        return;
    }
    const uint4* xrow = x + (size_t)m * nvec;
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int v = v0 + u * 32;
        raw[u] = v < nvec ? __ldg(xrow + v) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int v = v0 + u * 32;
        if (v >= nvec) break;                      // whole 16-lane groups leave together (nvec % 16 == 0)
        uint2 o;
        float r;
        bool slow = true;
        if (MODE == FLEXQ_ROUND_CUDA) {
            r = quant8_fast(raw[u], bits, o, slow);
            slow = __any_sync(__activemask(), slow);       // the exact routine shuffles across the group's 16 lanes
        }
        if (slow) {
            int q[8];
            r = quant8<MODE>(raw[u], bits, q);
            o.x = (uint32_t)(q[0] & 0xFF) | ((uint32_t)(q[1] & 0xFF) << 8) | ((uint32_t)(q[2] & 0xFF) << 16) | ((uint32_t)(q[3] & 0xFF) << 24);
            o.y = (uint32_t)(q[4] & 0xFF) | ((uint32_t)(q[5] & 0xFF) << 8) | ((uint32_t)(q[6] & 0xFF) << 16) | ((uint32_t)(q[7] & 0xFF) << 24);
        }
        *reinterpret_cast<uint2*>(xq + (size_t)m * K + (size_t)v * 8) = o;
        if (lane16 == 0) sx[(size_t)(v >> 4) * ldsx + m] = r;
    }
}

// ------------------------------------------------------------------------------------------------
// Producer-side fusions: the op that produces a GEMM's activations quantises them in
// the same pass, so the fp16 tensor makes no HBM round trip between the producer and the quantiser.
// Both kernels end in the same per-thread quantiser as quant_act_native_kernel (identical integers and
// scales for identical fp16 values).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float clamp_half_range(float v) {      // clamp_inf_for_half<half>, reduce_kernel_utils.cuh:356-361
    return v > 0.f ? fminf(v, 65504.f - 1000.f) : fmaxf(v, -65504.f + 1000.f);
}

template <int MODE>
__device__ __forceinline__ void quant_store(const uint4& raw, int m, int v, int K, int ldsx, int bits, int8_t* __restrict__ xq,
                                            float* __restrict__ sx, int lane16) {
    uint2 o;
    float r;
    bool slow = true;
    if (MODE == FLEXQ_ROUND_CUDA) {
        r = quant8_fast(raw, bits, o, slow);
        slow = __any_sync(__activemask(), slow);
    }
    if (slow) {
        int q[8];
        r = quant8<MODE>(raw, bits, q);
        o.x = (uint32_t)(q[0] & 0xFF) | ((uint32_t)(q[1] & 0xFF) << 8) | ((uint32_t)(q[2] & 0xFF) << 16) | ((uint32_t)(q[3] & 0xFF) << 24);
        o.y = (uint32_t)(q[4] & 0xFF) | ((uint32_t)(q[5] & 0xFF) << 8) | ((uint32_t)(q[6] & 0xFF) << 16) | ((uint32_t)(q[7] & 0xFF) << 24);
    }
    *reinterpret_cast<uint2*>(xq + (size_t)m * K + (size_t)v * 8) = o;
    if (lane16 == 0) sx[(size_t)(v >> 4) * ldsx + m] = r;
}

// RMSNorm (T5 style: no mean, no bias) with optional residual add, then quantise.  One CTA per token row,
// the row stays in registers between the two passes.  Arithmetic of the reference kernels
// generalT5LayerNormFlexQFusion / generalAddResidualT5LayerNormFlexQFusion (layernorm_kernels.cu:2494-2517,
// 1852-1905): residual' = half(clamp(x + residual)); var = sum(h^2) / K in fp32; rstd = rsqrtf(var + eps);
// y = half(clamp((float(h) * rstd) * float(gamma))); then the a6 quantiser on y.
template <int U, bool RESID, int MAXT>
__global__ void __launch_bounds__(MAXT) rmsnorm_quant_kernel(const uint4* __restrict__ x, uint4* __restrict__ resid,
                                                            const uint4* __restrict__ gamma, uint4* __restrict__ normed,
                                                            int8_t* __restrict__ xq, float* __restrict__ sx, float eps, int M, int K,
                                                            int ldsx, int bits) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int m = blockIdx.x;
    const int lane16 = threadIdx.x & 15;
    const int nvec = K >> 3;
    if (m >= M) {
        for (int v = threadIdx.x; v < nvec; v += blockDim.x)
            if ((v & 15) == 0) sx[(size_t)(v >> 4) * ldsx + m] = 0.f;
        return;
    }
    uint4 raw[U];
    float ss = 0.f;
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int v = u * blockDim.x + threadIdx.x;
        raw[u] = make_uint4(0, 0, 0, 0);
        if (v < nvec) {
            raw[u] = __ldg(x + (size_t)m * nvec + v);
            if (RESID) {
                const uint4 rr = resid[(size_t)m * nvec + v];
                __half2* a = reinterpret_cast<__half2*>(&raw[u]);
                const __half2* b = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float2 fa = __half22float2(a[i]), fb = __half22float2(b[i]);
                    a[i] = __floats2half2_rn(clamp_half_range(fa.x + fb.x), clamp_half_range(fa.y + fb.y));
                }
                resid[(size_t)m * nvec + v] = raw[u];
            }
            const __half2* h = reinterpret_cast<const __half2*>(&raw[u]);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float2 f = __half22float2(h[i]);
                ss = fmaf(f.x, f.x, ss);
                ss = fmaf(f.y, f.y, ss);
            }
        }
    }
    __shared__ float warp_sum[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = ss;
    __syncthreads();
    // every warp adds the (<= 32) warp sums itself: one block-wide sync and five shuffles instead of a serial loop on
    // thread 0 between two syncs (~1000 cycles of a decode-size launch that is all latency)
    float t = (int)(threadIdx.x & 31) < (int)(blockDim.x >> 5) ? warp_sum[threadIdx.x & 31] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    const float rstd = rsqrtf(t / (float)K + eps);
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int v = u * blockDim.x + threadIdx.x;
        if (v >= nvec) break;
        const uint4 gg = __ldg(gamma + v);
        __half2* h = reinterpret_cast<__half2*>(&raw[u]);
        const __half2* g2 = reinterpret_cast<const __half2*>(&gg);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float2 f = __half22float2(h[i]), g = __half22float2(g2[i]);
            h[i] = __floats2half2_rn(clamp_half_range((f.x * rstd) * g.x), clamp_half_range((f.y * rstd) * g.y));
        }
        if (normed) normed[(size_t)m * nvec + v] = raw[u];
        quant_store<FLEXQ_ROUND_CUDA>(raw[u], m, v, K, ldsx, bits, xq, sx, lane16);
    }
}

// SiLU(gate) * up, then quantise (the activations of down_proj).  Arithmetic of flexq_generic_activation with
// SiluActivation<half2> (activation_kernels.cu:129-144,246-298): silu in fp32 with __expf, product in fp32,
// rounded once to half.  `ld_in` = row stride of gate/up in halves (2*K when they are the two halves of a fused
// gate_up GEMM output).
template <int U>
__global__ void __launch_bounds__(256) silu_mul_quant_kernel(const __half* __restrict__ gate, const __half* __restrict__ up, long long ld_in,
                                                             uint4* __restrict__ out, int8_t* __restrict__ xq, float* __restrict__ sx,
                                                             int M, int K, int ldsx, int bits) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int m = blockIdx.x;
    const int lane16 = threadIdx.x & 15;
    const int nvec = K >> 3;
    const int v0 = ((blockIdx.y * 8 + (threadIdx.x >> 5)) * U) * 32 + (threadIdx.x & 31);
    if (m >= M) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int v = v0 + u * 32;
            if (v < nvec && lane16 == 0) sx[(size_t)(v >> 4) * ldsx + m] = 0.f;
        }
        return;
    }
    const uint4* grow = reinterpret_cast<const uint4*>(gate + (size_t)m * ld_in);
    const uint4* urow = reinterpret_cast<const uint4*>(up + (size_t)m * ld_in);
    uint4 g[U], w[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int v = v0 + u * 32;
        g[u] = v < nvec ? __ldg(grow + v) : make_uint4(0, 0, 0, 0);
        w[u] = v < nvec ? __ldg(urow + v) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int v = v0 + u * 32;
        if (v >= nvec) break;
        __half2* a = reinterpret_cast<__half2*>(&g[u]);
        const __half2* b = reinterpret_cast<const __half2*>(&w[u]);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float2 fa = __half22float2(a[i]), fb = __half22float2(b[i]);
            const float sx_ = __fdividef(fa.x, 1.0f + __expf(-fa.x)), sy_ = __fdividef(fa.y, 1.0f + __expf(-fa.y));
            a[i] = __floats2half2_rn(sx_ * fb.x, sy_ * fb.y);
        }
        if (out) out[(size_t)m * nvec + v] = g[u];
        quant_store<FLEXQ_ROUND_CUDA>(g[u], m, v, K, ldsx, bits, xq, sx, lane16);
    }
}

template <typename Kern, typename... Args>
static int launch_pdl(Kern kern, dim3 grid, int threads, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(threads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, kern, args...);
}

int rmsnorm_quant(const __half* x, __half* residual, const __half* gamma, float eps, __half* normed, int8_t* xq, float* sx, int M, int K,
                  int bits, cudaStream_t stream) {
    if (!x || !gamma || !xq || !sx) return FLEXQ_ERR_NULL;
    if (M <= 0 || K < kGroup || K % kGroup || K > 256 * 8 * 8) return FLEXQ_ERR_BAD_SHAPE;
    if (bits != 6 && bits != 8) return FLEXQ_ERR_BAD_BITS;
    const int ldsx = ceil4(M);
    // few rows (decode): one vector per thread, up to 1024 threads per row -- the shortest dependent chain, since only M
    // rows are in flight; many rows (prefill): 256 threads x up to 8 vectors, several rows resident per SM
    const int nvec = K / 8;
    const int tmax = ldsx >= 128 ? 256 : 1024;
    const int threads = nvec >= tmax ? tmax : ((nvec + 31) / 32) * 32;
    const int need = (nvec + threads - 1) / threads;
    const dim3 grid(ldsx);
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    uint4* rv = reinterpret_cast<uint4*>(residual);
    const uint4* gv = reinterpret_cast<const uint4*>(gamma);
    uint4* nv = reinterpret_cast<uint4*>(normed);
#define FQ_RMS(U_)                                                                                                              \
    return residual ? launch_pdl(rmsnorm_quant_kernel<U_, true, 256>, grid, threads, stream, xv, rv, gv, nv, xq, sx, eps, M, K, ldsx, bits) \
                    : launch_pdl(rmsnorm_quant_kernel<U_, false, 256>, grid, threads, stream, xv, rv, gv, nv, xq, sx, eps, M, K, ldsx, bits)
    if (tmax == 1024) {       // decode-sized: need is 1 (K <= 8192) or 2
        if (need <= 1)
            return residual ? launch_pdl(rmsnorm_quant_kernel<1, true, 1024>, grid, threads, stream, xv, rv, gv, nv, xq, sx, eps, M, K, ldsx, bits)
                            : launch_pdl(rmsnorm_quant_kernel<1, false, 1024>, grid, threads, stream, xv, rv, gv, nv, xq, sx, eps, M, K, ldsx, bits);
        return residual ? launch_pdl(rmsnorm_quant_kernel<2, true, 1024>, grid, threads, stream, xv, rv, gv, nv, xq, sx, eps, M, K, ldsx, bits)
                        : launch_pdl(rmsnorm_quant_kernel<2, false, 1024>, grid, threads, stream, xv, rv, gv, nv, xq, sx, eps, M, K, ldsx, bits);
    }
    if (need <= 1) { FQ_RMS(1); }
    if (need <= 2) { FQ_RMS(2); }
    if (need <= 4) { FQ_RMS(4); }
    FQ_RMS(8);
#undef FQ_RMS
}

int silu_mul_quant(const __half* gate, const __half* up, long long ld_in, __half* out, int8_t* xq, float* sx, int M, int K, int bits,
                   cudaStream_t stream) {
    if (!gate || !up || !xq || !sx) return FLEXQ_ERR_NULL;
    if (M <= 0 || K < kGroup || K % kGroup || ld_in < K || ld_in % 8) return FLEXQ_ERR_BAD_SHAPE;
    if (bits != 6 && bits != 8) return FLEXQ_ERR_BAD_BITS;
    const int ldsx = ceil4(M);
    const int nvec = K / 8;
    const bool big = (long long)ldsx * nvec >= 4LL * 148 * 2048;
    const int per_block = 256 * (big ? 2 : 1);
    const dim3 grid(ldsx, (nvec + per_block - 1) / per_block);
    uint4* ov = reinterpret_cast<uint4*>(out);
    if (big) return launch_pdl(silu_mul_quant_kernel<2>, grid, 256, stream, gate, up, ld_in, ov, xq, sx, M, K, ldsx, bits);
    return launch_pdl(silu_mul_quant_kernel<1>, grid, 256, stream, gate, up, ld_in, ov, xq, sx, M, K, ldsx, bits);
}

// Reference plane layout.  Lane l (0..15) of a group owns k = 8l..8l+7; plane word k32 = l/4 is
// assembled from the four lanes l = 4*k32 .. 4*k32+3; element k%32==0 sits in bit 31
// (engine/src/pack/bit_packing.cu:75 "__brev(__ballot_sync)").
__global__ void __launch_bounds__(256) quant_act_planes_kernel(const uint4* __restrict__ x, uint32_t* __restrict__ planes,
                                                               __half* __restrict__ xs, int M, int K, int bits) {
    const int G = K / kGroup;
    const long long vec = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long grp = vec >> 4;
    const int lane16 = threadIdx.x & 15;
    if (grp >= (long long)M * G) return;
    const int m = (int)(grp / G), g = (int)(grp - (long long)m * G);
    const uint4 raw = __ldg(x + ((size_t)m * K + (size_t)g * kGroup) / 8 + lane16);
    int q[8];
    const float r = quant8<FLEXQ_ROUND_CUDA>(raw, bits, q);
    const int chunk = M < 8 ? M : 8;
    const int ld = 2 * ceil4(M);                                       // SCALE_PACKING_A(SCALE_SIZE_X(M))
    if (lane16 == 0) {
        const __half sh = __float2half_rn(r);
        xs[(size_t)g * ld + 2 * m] = sh;                               // bit_packing.cu:152-156
        xs[(size_t)g * ld + 2 * m + 1] = sh;
    }
    const int sub = lane16 & 3, k32 = lane16 >> 2;
    const size_t base = (size_t)g * ((size_t)M * bits * 4) + (size_t)(m / chunk) * (bits * chunk * 4) + (size_t)(m % chunk) * 4 + k32;
    for (int b = 0; b < bits; b++) {
        uint32_t byte = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) byte |= (uint32_t)((q[i] >> b) & 1) << (7 - i);
        uint32_t word = byte << (8 * (3 - sub));
        word |= __shfl_xor_sync(0xffffffffu, word, 1);
        word |= __shfl_xor_sync(0xffffffffu, word, 2);
        if (sub == 0) planes[base + (size_t)b * (chunk * 4)] = word;
    }
}

int quant_act_native(const __half* x, int8_t* xq, float* sx, int M, int K, int bits, int mode, cudaStream_t stream) {
    if (!x || !xq || !sx) return FLEXQ_ERR_NULL;
    if (M <= 0 || K < kGroup || K % kGroup) return FLEXQ_ERR_BAD_SHAPE;
    if (bits != 6 && bits != 8) return FLEXQ_ERR_BAD_BITS;
    const int ldsx = ceil4(M);
    const int nvec = K / 8;
    // 4 vectors per thread once there is enough work to fill the machine; 1 for decode-sized inputs (latency)
    const bool big = (long long)ldsx * nvec >= 4LL * 148 * 2048;
    const int per_block = 256 * (big ? 4 : 1);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(ldsx, (nvec + per_block - 1) / per_block);
    cfg.blockDim = dim3(256);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    if (mode == FLEXQ_ROUND_PYTHON) {
        if (big) return (int)cudaLaunchKernelEx(&cfg, quant_act_native_kernel<FLEXQ_ROUND_PYTHON, 4>, xv, xq, sx, M, K, ldsx, bits);
        return (int)cudaLaunchKernelEx(&cfg, quant_act_native_kernel<FLEXQ_ROUND_PYTHON, 1>, xv, xq, sx, M, K, ldsx, bits);
    }
    if (big) return (int)cudaLaunchKernelEx(&cfg, quant_act_native_kernel<FLEXQ_ROUND_CUDA, 4>, xv, xq, sx, M, K, ldsx, bits);
    return (int)cudaLaunchKernelEx(&cfg, quant_act_native_kernel<FLEXQ_ROUND_CUDA, 1>, xv, xq, sx, M, K, ldsx, bits);
}

// fp32 activations, UniformAffineQuantizer arithmetic evaluated in fp32 exactly as torch does for float tensors
// (/root/reference/algorithm/flexq_quantize/quantizer.py:144-155 scale = absmax / qmax clamped to [1e-5, 1e4];
// :112-116 x_int = clamp(round_half_even(x / scale), qmin, qmax)).  This is the reference's CPU-runnable path
// (BASELINE.json configs[0]); sx receives the fp32 scale itself, which the GEMM applies unrounded.
__global__ void __launch_bounds__(256) quant_act_f32_kernel(const float4* __restrict__ x, int8_t* __restrict__ xq, float* __restrict__ sx,
                                                            int M, int K, int ldsx, int bits) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int m = blockIdx.x;
    const int lane16 = threadIdx.x & 15;
    const int nvec = K >> 3;                                                // 8-float vectors per row
    const int v = blockIdx.y * 256 + threadIdx.x;
    if (v >= nvec) return;                                                  // whole 16-lane groups leave together
    if (m >= M) {
        if (lane16 == 0) sx[(size_t)(v >> 4) * ldsx + m] = 0.f;
        return;
    }
    const float4* src = x + ((size_t)m * nvec + v) * 2;
    const float4 a = __ldg(src), b = __ldg(src + 1);
    const float xf[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) amax = fmaxf(amax, fabsf(xf[i]));
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    const int hi = (1 << (bits - 1)) - 1, lo = -(1 << (bits - 1));
    const float s = fminf(fmaxf(__fdiv_rn(amax, (float)hi), 1e-5f), 1e4f);
    uint32_t w[2] = {0u, 0u};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float q = fminf(fmaxf(rintf(__fdiv_rn(xf[i], s)), (float)lo), (float)hi);
        w[i >> 2] |= (uint32_t)((int)q & 0xFF) << (8 * (i & 3));
    }
    *reinterpret_cast<uint2*>(xq + (size_t)m * K + (size_t)v * 8) = make_uint2(w[0], w[1]);
    if (lane16 == 0) sx[(size_t)(v >> 4) * ldsx + m] = s;
}

int quant_act_f32(const float* x, int8_t* xq, float* sx, int M, int K, int bits, cudaStream_t stream) {
    if (!x || !xq || !sx) return FLEXQ_ERR_NULL;
    if (M <= 0 || K < kGroup || K % kGroup) return FLEXQ_ERR_BAD_SHAPE;
    if (bits != 6 && bits != 8) return FLEXQ_ERR_BAD_BITS;
    const int ldsx = ceil4(M);
    const int nvec = K / 8;
    return launch_pdl(quant_act_f32_kernel, dim3(ldsx, (nvec + 255) / 256), 256, stream, reinterpret_cast<const float4*>(x), xq, sx, M, K,
                      ldsx, bits);
}

int quant_act_planes(const __half* x, uint32_t* planes, __half* xs, int M, int K, int bits, cudaStream_t stream) {
    if (!x || !planes || !xs) return FLEXQ_ERR_NULL;
    if (M <= 0 || K < kGroup || K % kGroup || (M > 8 && M % 8)) return FLEXQ_ERR_BAD_SHAPE;
    if (bits != 6 && bits != 8) return FLEXQ_ERR_BAD_BITS;
    // padding entries of X_SCALE (rows M..ceil4(M)) are zero in the reference harness
    FLEXQ_CUDA_TRY(cudaMemsetAsync(xs, 0, (size_t)(K / kGroup) * 2 * ceil4(M) * sizeof(__half), stream));
    const long long threads = (long long)M * (K / kGroup) * 16;
    const int blocks = (int)((threads + 255) / 256);
    quant_act_planes_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), planes, xs, M, K, bits);
    return (int)cudaGetLastError();
}

}  // namespace flexq
