// Micro-benchmark: cycles per tcgen05.mma kind::i8 (M=128, A from TMEM, B from shared memory with 128-byte swizzle,
// K = 32) as issued by the W6Ax GEMM, alone and with competing shared-memory traffic:
//   mode 0: MMAs only            mode 1: + 8 warps streaming broadcast LDS.128      mode 2: + 8 warps streaming distinct LDS.128
//   mode 3: + one thread issuing 24 KB bulk copies global -> shared back to back
// Prints cycles per MMA for N = 64, 128, 192, 256 (one CTA per SM, 148 CTAs).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../flexq_b200/csrc/common.cuh"
using namespace flexq;

constexpr int kIters = 400;

template <int N, int MODE>
__global__ void __launch_bounds__(512, 1) k(long long* cyc, const uint8_t* gsrc, float* sink) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* smem = raw + (base - smem_u32(raw));
    __shared__ uint32_t tb;
    __shared__ __align__(8) unsigned long long bars[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4 * 256 * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (i & 3);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1); fence_barrier_init(); }
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc<512>(smem_u32(&tb));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tb;
    if (warp == 0) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_i8(128, N);
            long long t0 = clock64();
            for (int it = 0; it < kIters; it++) {
                const uint64_t bdesc = umma_desc_sw128(base + (it & 3) * (256 * 128));
#pragma unroll
                for (int kk = 0; kk < 4; kk++)
                    umma_i8_ts(tmem + (N <= 192 ? (it & 1) * 224 : 0), tmem + 480 + 8 * (kk & 3), bdesc + (uint64_t)((32 * kk) >> 4), idesc, kk > 0 ? 1u : 0u);
            }
            umma_commit(smem_u32(&bars[0]));
            mbar_wait(smem_u32(&bars[0]), 0);
            long long t1 = clock64();
            cyc[blockIdx.x] = t1 - t0;
            reinterpret_cast<volatile uint32_t*>(&tb)[0] = 0xFFFFFFFFu;      // stop flag for the traffic generators
        }
        __syncwarp();
    } else if (MODE == 3 && warp == 1) {
        if (lane == 0) {
            uint32_t ph = 0;
            while (reinterpret_cast<volatile uint32_t*>(&tb)[0] != 0xFFFFFFFFu) {
                mbar_expect_tx(smem_u32(&bars[1]), 24576);
                bulk_g2s(base + 4 * 256 * 128, gsrc + (blockIdx.x & 63) * 24576, 24576, smem_u32(&bars[1]));
                mbar_wait(smem_u32(&bars[1]), ph);
                ph ^= 1;
            }
        }
        __syncwarp();
    } else if ((MODE == 1 || MODE == 2) && warp >= 8) {
        float acc = 0.f;
        const uint32_t a0 = base + (MODE == 2 ? lane * 16 : 0);
        while (reinterpret_cast<volatile uint32_t*>(&tb)[0] != 0xFFFFFFFFu) {
#pragma unroll
            for (int j = 0; j < 16; j++) { const float4 v = lds_f4(a0 + ((j * 512 + warp * 64) & 0x7FFF)); acc += v.x + v.w; }
        }
        sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    }
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int N, int MODE>
void run(const uint8_t* gsrc, float* sink, long long* cyc) {
    const int smem = 4 * 256 * 128 + 24576 + 2048;
    cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<N, MODE><<<148, 512, smem>>>(cyc, gsrc, sink);
    k<N, MODE><<<148, 512, smem>>>(cyc, gsrc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e) { printf("N=%d mode %d: %s\n", N, MODE, cudaGetErrorString(e)); return; }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; i++) c += h[i];
    c /= 148;
    printf("N=%3d mode %d: %.1f cycles per MMA (128 x %d x 32 int8), ideal %d\n", N, MODE, c / (kIters * 4), N, 128 * N / 256);
}

int main() {
    uint8_t* gsrc; float* sink; long long* cyc;
    cudaMalloc(&gsrc, 64 * 24576); cudaMemset(gsrc, 1, 64 * 24576);
    cudaMalloc(&sink, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    run<64, 0>(gsrc, sink, cyc); run<128, 0>(gsrc, sink, cyc); run<192, 0>(gsrc, sink, cyc); run<256, 0>(gsrc, sink, cyc);
    run<192, 1>(gsrc, sink, cyc); run<192, 2>(gsrc, sink, cyc); run<192, 3>(gsrc, sink, cyc);
    run<256, 1>(gsrc, sink, cyc); run<256, 3>(gsrc, sink, cyc);
    return 0;
}
