// Reference bit-plane layout: packer and converters (API parity with the reference engine).
//
// Layout (normative: /root/reference/engine/test_packing_kernel.cu:139-141 and
// engine/src/pack/bit_packing.cu:84-98): u32[K/128][R/chunk][bits][chunk][4], chunk = min(R,8);
// word (b, r, k32) holds bit b of elements k = 32*k32 .. 32*k32+31 of row r with element
// k%32 == 0 in bit 31 (bit_packing.cu:75).  Rows are two's complement in `bits` bits.
#include "common.cuh"

namespace flexq {

// 16 lanes per (row, 128-group); lane l owns k = 8l .. 8l+7 (two 128-bit loads of int32).
__global__ void __launch_bounds__(256) pack_planes_i32_kernel(const int4* __restrict__ in, uint32_t* __restrict__ planes,
                                                              int R, int K, int bits) {
    const int G = K / kGroup;
    const long long vec = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long grp = vec >> 4;
    const int lane16 = threadIdx.x & 15;
    if (grp >= (long long)R * G) return;
    const int r = (int)(grp / G), g = (int)(grp - (long long)r * G);
    const int4* src = in + ((size_t)r * K + (size_t)g * kGroup + lane16 * 8) / 4;
    const int4 a = __ldg(src), b4 = __ldg(src + 1);
    const int q[8] = {a.x, a.y, a.z, a.w, b4.x, b4.y, b4.z, b4.w};
    const int chunk = R < 8 ? R : 8;
    const int sub = lane16 & 3, k32 = lane16 >> 2;
    const size_t base = (size_t)g * ((size_t)R * bits * 4) + (size_t)(r / chunk) * (bits * chunk * 4) + (size_t)(r % chunk) * 4 + k32;
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

// planes -> int8 [R][K]; one thread per (row, 32-element word): reads `bits` words, writes 32 B
__global__ void __launch_bounds__(256) planes_to_i8_kernel(const uint32_t* __restrict__ planes, int8_t* __restrict__ out,
                                                           int R, int K, int bits) {
    const int W = K / 32;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)R * W) return;
    const int r = (int)(idx / W), k32 = (int)(idx - (long long)r * W);
    const int chunk = R < 8 ? R : 8;
    const size_t base = (size_t)(k32 / 4) * ((size_t)R * bits * 4) + (size_t)(r / chunk) * (bits * chunk * 4) + (size_t)(r % chunk) * 4 + (k32 % 4);
    uint32_t pl[8];
    for (int b = 0; b < bits; b++) pl[b] = planes[base + (size_t)b * (chunk * 4)];
    uint32_t o[8];
#pragma unroll
    for (int wv = 0; wv < 8; wv++) {
        uint32_t word = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int l = 4 * wv + j;
            int v = 0;
            for (int b = 0; b < bits; b++) v |= (int)((pl[b] >> (31 - l)) & 1u) << b;
            v = (v << (32 - bits)) >> (32 - bits);                     // sign-extend: MSB plane weighs -2^(bits-1)
            word |= (uint32_t)(v & 0xFF) << (8 * j);
        }
        o[wv] = word;
    }
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)r * K + (size_t)k32 * 32);
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// X_SCALE half[G][2*ceil4(M)] (pairs) -> sx f32[G][ceil4(M)]
__global__ void xscale_ref_to_sx_kernel(const __half* __restrict__ xs, float* __restrict__ sx, int M, int G) {
    const int ld = ceil4(M);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= G * ld) return;
    const int g = idx / ld, m = idx - g * ld;
    sx[idx] = (m < M) ? __half2float(xs[(size_t)g * 2 * ld + 2 * m]) : 0.f;
}

int pack_planes_i32(const int32_t* in, uint32_t* planes, int R, int K, int bits, cudaStream_t stream) {
    if (!in || !planes) return FLEXQ_ERR_NULL;
    if (R <= 0 || K < kGroup || K % kGroup || (R > 8 && R % 8)) return FLEXQ_ERR_BAD_SHAPE;
    if (bits < 1 || bits > 8) return FLEXQ_ERR_BAD_BITS;
    const long long threads = (long long)R * (K / kGroup) * 16;
    pack_planes_i32_kernel<<<(int)((threads + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const int4*>(in), planes, R, K, bits);
    return (int)cudaGetLastError();
}

int planes_to_i8(const uint32_t* planes, int8_t* out, int R, int K, int bits, cudaStream_t stream) {
    if (!planes || !out) return FLEXQ_ERR_NULL;
    if (R <= 0 || K < kGroup || K % kGroup || (R > 8 && R % 8)) return FLEXQ_ERR_BAD_SHAPE;
    if (bits < 1 || bits > 8) return FLEXQ_ERR_BAD_BITS;
    const long long threads = (long long)R * (K / 32);
    planes_to_i8_kernel<<<(int)((threads + 255) / 256), 256, 0, stream>>>(planes, out, R, K, bits);
    return (int)cudaGetLastError();
}

int xscale_ref_to_sx(const __half* xs, float* sx, int M, int K, cudaStream_t stream) {
    if (!xs || !sx) return FLEXQ_ERR_NULL;
    if (M <= 0 || K < kGroup || K % kGroup) return FLEXQ_ERR_BAD_SHAPE;
    const int total = (K / kGroup) * ceil4(M);
    xscale_ref_to_sx_kernel<<<(total + 255) / 256, 256, 0, stream>>>(xs, sx, M, K / kGroup);
    return (int)cudaGetLastError();
}

}  // namespace flexq
