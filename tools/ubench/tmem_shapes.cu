// Micro-test: register <-> (lane, column) mapping of tcgen05.ld shapes 16x256b / 16x128b / 16x64b on sm_100a.
// TMEM is filled through the 32x32b shape (thread = lane, register = column) with the value (lane << 16 | column) and
// read back through the other shapes; the host prints, for every thread and register of warp 0, which (lane, column)
// it received.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k(uint32_t* out) {
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t0 = tbase + ((uint32_t)(warp * 32) << 16);
    uint32_t v[16];
    for (int c = 0; c < 16; c++) v[c] = ((uint32_t)(warp * 32 + lane) << 16) | c;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(t0),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                 "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[8];
    // 16x256b.x2: 8 registers per thread, lanes [base, base+16), 16 columns
    for (int half = 0; half < 2; half++) {
        const uint32_t ta = t0 + ((uint32_t)(16 * half) << 16);
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(ta));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; i++) out[((0 * 2 + half) * 128 + threadIdx.x) * 8 + i] = r[i];
    }
    // 16x128b.x4: 8 registers per thread
    for (int half = 0; half < 2; half++) {
        const uint32_t ta = t0 + ((uint32_t)(16 * half) << 16);
        asm volatile("tcgen05.ld.sync.aligned.16x128b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(ta));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; i++) out[((1 * 2 + half) * 128 + threadIdx.x) * 8 + i] = r[i];
    }
    // 16x64b.x8: 8 registers per thread
    for (int half = 0; half < 2; half++) {
        const uint32_t ta = t0 + ((uint32_t)(16 * half) << 16);
        asm volatile("tcgen05.ld.sync.aligned.16x64b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(ta));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; i++) out[((2 * 2 + half) * 128 + threadIdx.x) * 8 + i] = r[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tbase) : "memory");
}

int main() {
    uint32_t* d; const int n = 3 * 2 * 128 * 8;
    cudaMalloc(&d, n * 4); cudaMemset(d, 0xFF, n * 4);
    k<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    uint32_t* h = new uint32_t[n];
    cudaMemcpy(h, d, n * 4, cudaMemcpyDeviceToHost);
    const char* names[3] = {"16x256b.x2", "16x128b.x4", "16x64b.x8"};
    for (int s = 0; s < 3; s++)
        for (int half = 0; half < 2; half++) {
            printf("shape %s, lane base +%d, warp 0 (thread: reg -> lane.col) and warp 2 thread 5\n", names[s], 16 * half);
            for (int t = 0; t < 32; t += (t < 8 ? 1 : 8)) {
                printf("  t%02d:", t);
                for (int i = 0; i < 8; i++) { uint32_t v = h[((s * 2 + half) * 128 + t) * 8 + i]; printf(" %u.%u", v >> 16, v & 0xFFFF); }
                printf("\n");
            }
            printf("  w2t05:");
            for (int i = 0; i < 8; i++) { uint32_t v = h[((s * 2 + half) * 128 + 64 + 5) * 8 + i]; printf(" %u.%u", v >> 16, v & 0xFFFF); }
            printf("\n");
        }
    return 0;
}
