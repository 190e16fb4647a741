// Micro-benchmark: issue/pipe throughput of the dequant epilogue's instruction forms on sm_100a.
// Prints warp-instructions per clock per SM sub-partition for each form, at 1/2/4 warps per sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N 32
#define ITERS 2000

template <int V>
__global__ void k(float* out, long long* cyc, float sw, float bias, const float* sx) {
    float t[N], acc[N], s[N];
    unsigned r[N];
#pragma unroll
    for (int i = 0; i < N; i++) { acc[i] = threadIdx.x * 0.5f + i; s[i] = sx[i] ; r[i] = threadIdx.x + i; t[i] = i; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
        if (V == 0) {          // scalar FFMA, 3 distinct registers: acc = t*s+acc
#pragma unroll
            for (int i = 0; i < N; i++) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(t[i]), "f"(s[i]));
        } else if (V == 1) {   // FFMA2 3 distinct pairs
#pragma unroll
            for (int i = 0; i < N; i += 2) {
                unsigned long long a, b, c;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(acc[i]), "f"(acc[i + 1]));
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(t[i]), "f"(t[i + 1]));
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(c) : "f"(s[i]), "f"(s[i + 1]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(b), "l"(c));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(acc[i]), "=f"(acc[i + 1]) : "l"(a));
            }
        } else if (V == 2) {   // scalar FFMA with two shared operands: t = t*sw+bias
#pragma unroll
            for (int i = 0; i < N; i++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(t[i]) : "f"(sw), "f"(bias));
        } else if (V == 3) {   // FFMA2 with broadcast scalars
#pragma unroll
            for (int i = 0; i < N; i += 2) {
                unsigned long long a, b, c;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(t[i]), "f"(t[i + 1]));
                asm volatile("mov.b64 %0, {%1,%1};" : "=l"(b) : "f"(sw));
                asm volatile("mov.b64 %0, {%1,%1};" : "=l"(c) : "f"(bias));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(b), "l"(c));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(t[i]), "=f"(t[i + 1]) : "l"(a));
            }
        } else if (V == 4) {   // integer add with immediate (VIADD)
#pragma unroll
            for (int i = 0; i < N; i++) { unsigned x; asm volatile("add.u32 %0, %1, 0x4b400000;" : "=r"(x) : "r"(r[i])); asm volatile("" :: "r"(x)); }
        } else if (V == 5) {   // full current pattern per 2 elements: 2 VIADD + FFMA2(bcast) + FFMA2(acc)
#pragma unroll
            for (int i = 0; i < N; i += 2) {
                unsigned x0, x1; asm volatile("add.u32 %0, %1, 0x4b400000;" : "=r"(x0) : "r"(r[i])); asm volatile("add.u32 %0, %1, 0x4b400000;" : "=r"(x1) : "r"(r[i + 1]));
                unsigned long long a, b, c, d, e;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(a) : "r"(x0), "r"(x1));
                asm volatile("mov.b64 %0, {%1,%1};" : "=l"(b) : "f"(sw));
                asm volatile("mov.b64 %0, {%1,%1};" : "=l"(c) : "f"(bias));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(b), "l"(c));
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(d) : "f"(acc[i]), "f"(acc[i + 1]));
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(e) : "f"(s[i]), "f"(s[i + 1]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(e));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(acc[i]), "=f"(acc[i + 1]) : "l"(d));
                
            }
        } else if (V == 6) {   // scalar pattern per element: VIADD + FFMA(shared) + FFMA(acc)
#pragma unroll
            for (int i = 0; i < N; i++) {
                unsigned x0; asm volatile("add.u32 %0, %1, 0x4b400000;" : "=r"(x0) : "r"(r[i]));
                float tt;
                asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(tt) : "f"(__uint_as_float(x0)), "f"(sw), "f"(bias));
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(tt), "f"(s[i]));
                
            }
        } else if (V == 7) {   // mixed: FFMA2 for t, scalar for acc
#pragma unroll
            for (int i = 0; i < N; i += 2) {
                unsigned x0, x1; asm volatile("add.u32 %0, %1, 0x4b400000;" : "=r"(x0) : "r"(r[i])); asm volatile("add.u32 %0, %1, 0x4b400000;" : "=r"(x1) : "r"(r[i + 1]));
                unsigned long long a, b, c;
                float t0, t1;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(a) : "r"(x0), "r"(x1));
                asm volatile("mov.b64 %0, {%1,%1};" : "=l"(b) : "f"(sw));
                asm volatile("mov.b64 %0, {%1,%1};" : "=l"(c) : "f"(bias));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(b), "l"(c));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(t0), "=f"(t1) : "l"(a));
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(t0), "f"(s[i]));
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i + 1]) : "f"(t1), "f"(s[i + 1]));
                
            }
        } else if (V == 8) {   // I2F conversion
#pragma unroll
            for (int i = 0; i < N; i++) { float f; asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(r[i])); r[i] = __float_as_uint(f) ^ it; }
        } else if (V == 9) {   // FMUL scalar with shared operand then FFMA acc (2 fma-pipe ops, 2+3 regs)
#pragma unroll
            for (int i = 0; i < N; i++) {
                float tt;
                asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(tt) : "f"(__uint_as_float(r[i])), "f"(sw));
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(tt), "f"(s[i]));
            }
        } else if (V == 10) {  // half2 HFMA2: 3 distinct regs
#pragma unroll
            for (int i = 0; i < N; i++) asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(__float_as_uint(t[i])), "r"(__float_as_uint(s[i])));
        }
    }
    long long t1 = clock64();
    float sum = 0;
#pragma unroll
    for (int i = 0; i < N; i++) sum += acc[i] + t[i] + __uint_as_float(r[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const char* name, double inst_per_iter) {
    float* out; long long* cyc; float* sx;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sx, N * 4);
    cudaMemset(sx, 0, N * 4);
    printf("%-46s", name);
    for (int wps : {1, 2, 4}) {
        int threads = wps * 4 * 32;
        k<V><<<148, threads>>>(out, cyc, 1.0f, 0.0f, sx);
        k<V><<<148, threads>>>(out, cyc, 1.0f, 0.0f, sx);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
        double ipc = inst_per_iter * ITERS * wps / c;     // warp-instr per clk per SMSP
        printf("  w/smsp=%d: %.3f inst/clk (%.2f clk/iter)", wps, ipc, c / ITERS);
    }
    printf("\n");
    cudaError_t e = cudaGetLastError(); if (e) printf("err %s\n", cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc); cudaFree(sx);
}

int main() {
    run<0>("FFMA   acc=t*s+acc (3 distinct regs)", N);
    run<1>("FFMA2  acc=t*s+acc (3 distinct pairs)", N / 2);
    run<2>("FFMA   t=t*sw+bias (2 shared regs)", N);
    run<3>("FFMA2  t=t*sw+bias (broadcast scalars)", N / 2);
    run<4>("VIADD  r+=imm", N);
    run<5>("pattern now: 2 VIADD+FFMA2+FFMA2 /2el", N / 2 * 4);
    run<6>("pattern scalar: VIADD+FFMA+FFMA /el", N * 3);
    run<7>("pattern mix: FFMA2 t + 2 FFMA acc /2el", N / 2 * 5);
    run<8>("I2F", N * 2);
    run<9>("FMUL(shared)+FFMA(acc) /el", N * 2);
    run<10>("HFMA2 3 distinct regs", N);
    return 0;
}
