// Offline weight packer (north-star subsystem 1): [N][K] weights -> W6 tiles + fp16 group scales.
//
// W6 tile layout (DESIGN.md "Data layout in HBM"): tile (nt, g) = weight rows 128nt..128nt+127 x
// k 128g..128g+127, stored as one contiguous 12288-byte block at ((nt*G + g) * 12288) so that a
// single cp.async.bulk moves it into shared memory.  Inside a tile: 256 units of 48 B, unit
// (q, r) = k-half q of row r at byte 48*(128q + r); a unit is four 12-byte sub-chunks of 16
// k-values (three u32 words, see w6_encode16).  The order is chosen so that the GEMM's expander
// threads read their unit with conflict-free 128-bit shared loads and emit the swizzle-128B
// UMMA operand with conflict-free 128-bit stores.
//
// The quantising variants restate UniformAffineQuantizer for weights
// (/root/reference/algorithm/flexq_quantize/quantizer.py:144-171 calibration, :93-126
// fake_quant) with arithmetic in the input dtype, and emit the integers instead of the
// fake-quantised floats (the exporter the reference lacks: flexq_quantize/utils.py:116-123).
#include "common.cuh"

namespace flexq {

template <typename T>
__device__ __forceinline__ int load_int(const T* p);
template <>
__device__ __forceinline__ int load_int<int32_t>(const int32_t* p) { return *p; }
template <>
__device__ __forceinline__ int load_int<int8_t>(const int8_t* p) { return (int)*p; }

// one thread per unit (64 k-values -> 48 bytes)
template <typename T>
__global__ void __launch_bounds__(128) pack_w6_kernel(const T* __restrict__ w, uint8_t* __restrict__ w6, int N, int K) {
    const int G = K / kGroup;
    const int r = threadIdx.x;                 // row in tile
    const int g = blockIdx.x, nt = blockIdx.y, q = blockIdx.z;
    const int n = nt * kTileN + r;
    uint32_t words[12];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        int e[16];
#pragma unroll
        for (int i = 0; i < 16; i++) e[i] = (n < N) ? load_int<T>(w + (size_t)n * K + (size_t)g * kGroup + 64 * q + 16 * s + i) : 0;
        w6_encode16(e, words + 3 * s);
    }
    uint4* dst = reinterpret_cast<uint4*>(w6 + ((size_t)nt * G + g) * kTileBytes + 48 * (128 * q + r));
    dst[0] = make_uint4(words[0], words[1], words[2], words[3]);
    dst[1] = make_uint4(words[4], words[5], words[6], words[7]);
    dst[2] = make_uint4(words[8], words[9], words[10], words[11]);
}

template <typename T> struct Arith;
template <> struct Arith<float> {
    static __device__ __forceinline__ float load(const float* p) { return *p; }
    static __device__ __forceinline__ float rnd(float v) { return v; }
};
template <> struct Arith<__half> {
    static __device__ __forceinline__ float load(const __half* p) { return __half2float(*p); }
    static __device__ __forceinline__ float rnd(float v) { return __half2float(__float2half_rn(v)); }   // torch half op
};

// one thread per (row, group): absmax -> scale -> quantise -> both units of the row
template <typename T>
__global__ void __launch_bounds__(128) quant_pack_w6_kernel(const T* __restrict__ w, uint8_t* __restrict__ w6,
                                                            __half* __restrict__ w_scale, int N, int K) {
    const int G = K / kGroup;
    const int r = threadIdx.x;
    const int g = blockIdx.x, nt = blockIdx.y;
    const int n = nt * kTileN + r;
    const T* src = w + (size_t)n * K + (size_t)g * kGroup;
    float amax = 0.f;
    if (n < N)
        for (int i = 0; i < kGroup; i++) amax = fmaxf(amax, fabsf(Arith<T>::load(src + i)));
    float s = Arith<T>::rnd(__fdiv_rn(amax, 31.f));                               // quantizer.py:154
    s = fminf(fmaxf(s, Arith<T>::rnd(1e-5f)), Arith<T>::rnd(1e4f));               // :155
    if (n < N) w_scale[(size_t)g * N + n] = __float2half_rn(s);
#pragma unroll 1
    for (int q = 0; q < 2; q++) {
        uint32_t words[12];
#pragma unroll
        for (int sgm = 0; sgm < 4; sgm++) {
            int e[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                float v = 0.f;
                if (n < N) {
                    const float t = Arith<T>::rnd(__fdiv_rn(Arith<T>::load(src + 64 * q + 16 * sgm + i), s));   // :112
                    v = fminf(fmaxf(rintf(t), -32.f), 31.f);                                                  // :112,:116
                }
                e[i] = (int)v;
            }
            w6_encode16(e, words + 3 * sgm);
        }
        uint4* dst = reinterpret_cast<uint4*>(w6 + ((size_t)nt * G + g) * kTileBytes + 48 * (128 * q + r));
        dst[0] = make_uint4(words[0], words[1], words[2], words[3]);
        dst[1] = make_uint4(words[4], words[5], words[6], words[7]);
        dst[2] = make_uint4(words[8], words[9], words[10], words[11]);
    }
}

// W6 tiles -> int8 [N][K] (value w, not 4w); tests and converters only
__global__ void __launch_bounds__(128) unpack_w6_kernel(const uint8_t* __restrict__ w6, int8_t* __restrict__ out, int N, int K) {
    const int G = K / kGroup;
    const int r = threadIdx.x;
    const int g = blockIdx.x, nt = blockIdx.y, q = blockIdx.z;
    const int n = nt * kTileN + r;
    if (n >= N) return;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(w6 + ((size_t)nt * G + g) * kTileBytes + 48 * (128 * q + r));
#pragma unroll
    for (int s = 0; s < 4; s++) {
        uint32_t o[4];
        w6_expand16(src[3 * s], src[3 * s + 1], src[3 * s + 2], o);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            // arithmetic >> 2 per byte: containers hold 4*w
            uint32_t word = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int v = (int)(int8_t)((o[i] >> (8 * j)) & 0xFF) >> 2;
                word |= (uint32_t)(v & 0xFF) << (8 * j);
            }
            *reinterpret_cast<uint32_t*>(out + (size_t)n * K + (size_t)g * kGroup + 64 * q + 16 * s + 4 * i) = word;
        }
    }
}

static int check_nk(int N, int K) { return (N <= 0 || K < kGroup || K % kGroup) ? FLEXQ_ERR_BAD_SHAPE : 0; }

template <typename T>
int pack_w6(const T* w, uint8_t* w6, int N, int K, cudaStream_t stream) {
    if (!w || !w6) return FLEXQ_ERR_NULL;
    if (int e = check_nk(N, K)) return e;
    dim3 grid(K / kGroup, ceil_div(N, kTileN), 2);
    pack_w6_kernel<T><<<grid, 128, 0, stream>>>(w, w6, N, K);
    return (int)cudaGetLastError();
}
template int pack_w6<int32_t>(const int32_t*, uint8_t*, int, int, cudaStream_t);
template int pack_w6<int8_t>(const int8_t*, uint8_t*, int, int, cudaStream_t);

template <typename T>
int quant_pack_w6(const T* w, uint8_t* w6, __half* w_scale, int N, int K, cudaStream_t stream) {
    if (!w || !w6 || !w_scale) return FLEXQ_ERR_NULL;
    if (int e = check_nk(N, K)) return e;
    dim3 grid(K / kGroup, ceil_div(N, kTileN));
    quant_pack_w6_kernel<T><<<grid, 128, 0, stream>>>(w, w6, w_scale, N, K);
    return (int)cudaGetLastError();
}
template int quant_pack_w6<float>(const float*, uint8_t*, __half*, int, int, cudaStream_t);
template int quant_pack_w6<__half>(const __half*, uint8_t*, __half*, int, int, cudaStream_t);

int unpack_w6(const uint8_t* w6, int8_t* out, int N, int K, cudaStream_t stream) {
    if (!w6 || !out) return FLEXQ_ERR_NULL;
    if (int e = check_nk(N, K)) return e;
    dim3 grid(K / kGroup, ceil_div(N, kTileN), 2);
    unpack_w6_kernel<<<grid, 128, 0, stream>>>(w6, out, N, K);
    return (int)cudaGetLastError();
}

}  // namespace flexq
