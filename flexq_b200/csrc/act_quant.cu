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

template <int MODE>
__global__ void __launch_bounds__(256) quant_act_native_kernel(const uint4* __restrict__ x, int8_t* __restrict__ xq,
                                                               float* __restrict__ sx, int M, int K, int ldsx, int bits) {
    // programmatic dependent launch: the GEMM that consumes Xq/sx may start streaming its weights now;
    // this kernel itself waits for the producer of X (no-ops when launched without the attribute)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int G = K / kGroup;
    const long long vec = (long long)blockIdx.x * blockDim.x + threadIdx.x;    // 8-half vector index
    const long long grp = vec >> 4;                                           // (row, group) index
    const int lane16 = threadIdx.x & 15;
    const long long total = (long long)ldsx * G;
    if (grp >= total) return;                      // whole 16-lane groups exit together
    const int m = (int)(grp / G), g = (int)(grp - (long long)m * G);
    if (m >= M) {                                  // padding rows of sx
        if (lane16 == 0) sx[(size_t)g * ldsx + m] = 0.f;
        return;
    }
    const uint4 raw = __ldg(x + ((size_t)m * K + (size_t)g * kGroup) / 8 + lane16);
    int q[8];
    const float r = quant8<MODE>(raw, bits, q);
    uint2 o;
    o.x = (uint32_t)(q[0] & 0xFF) | ((uint32_t)(q[1] & 0xFF) << 8) | ((uint32_t)(q[2] & 0xFF) << 16) | ((uint32_t)(q[3] & 0xFF) << 24);
    o.y = (uint32_t)(q[4] & 0xFF) | ((uint32_t)(q[5] & 0xFF) << 8) | ((uint32_t)(q[6] & 0xFF) << 16) | ((uint32_t)(q[7] & 0xFF) << 24);
    *reinterpret_cast<uint2*>(xq + (size_t)m * K + (size_t)g * kGroup + lane16 * 8) = o;
    if (lane16 == 0) sx[(size_t)g * ldsx + m] = r;
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
    const long long threads = (long long)ldsx * (K / kGroup) * 16;
    const int blocks = (int)((threads + 255) / 256);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(256);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    if (mode == FLEXQ_ROUND_PYTHON) return (int)cudaLaunchKernelEx(&cfg, quant_act_native_kernel<FLEXQ_ROUND_PYTHON>, xv, xq, sx, M, K, ldsx, bits);
    return (int)cudaLaunchKernelEx(&cfg, quant_act_native_kernel<FLEXQ_ROUND_CUDA>, xv, xq, sx, M, K, ldsx, bits);
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
