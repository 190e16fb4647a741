// Micro-benchmark: does FFMA2 take fp32 subnormal inputs at full rate and without flushing on sm_100a?
// The prefill epilogue feeds the tensor core's int32 accumulator (a non-negative integer < 2^23) to an FMA as the
// subnormal it is; this checks (a) numerics against a double-precision evaluation and (b) warp-instructions per
// clock per SM sub-partition for subnormal vs normal first operands.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

#define N 32
#define ITERS 2000

__global__ void numerics(const uint32_t* acc, float sw, float sx, float* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float c1 = sw * 0x1p100f, c2 = -2080800.f * (sw * 0x1p-49f);
    unsigned long long a, b, c, d, e;
    asm volatile("mov.b64 %0, {%1,%1};" : "=l"(a) : "r"(acc[i]));
    asm volatile("mov.b64 %0, {%1,%1};" : "=l"(b) : "f"(c1));
    asm volatile("mov.b64 %0, {%1,%1};" : "=l"(c) : "f"(c2));
    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(b), "l"(c));
    asm volatile("mov.b64 %0, {%1,%1};" : "=l"(d) : "f"(sx));
    asm volatile("mov.b64 %0, {%1,%1};" : "=l"(e) : "f"(0.f));
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(e) : "l"(a), "l"(d));
    float lo, hi;
    asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(e));
    out[i] = lo * 0x1p47f;
}

template <bool DENORM>
__global__ void k(float* out, long long* cyc, float sw, float bias, const float* sx, const uint32_t* seed) {
    float acc[N], s[N];
    unsigned r[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        acc[i] = 0.f; s[i] = sx[i];
        r[i] = DENORM ? (seed[i] & 0x3FFFFF) + 32u : (0x3F800000u | (seed[i] & 0x3FFFFF));
    }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            unsigned long long a, b, c, d, e;
            asm volatile("mov.b64 %0, {%1,%2};" : "=l"(a) : "r"(r[i]), "r"(r[i + 1]));
            asm volatile("mov.b64 %0, {%1,%1};" : "=l"(b) : "f"(sw));
            asm volatile("mov.b64 %0, {%1,%1};" : "=l"(c) : "f"(bias));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(b), "l"(c));
            asm volatile("mov.b64 %0, {%1,%2};" : "=l"(d) : "f"(acc[i]), "f"(acc[i + 1]));
            asm volatile("mov.b64 %0, {%1,%2};" : "=l"(e) : "f"(s[i]), "f"(s[i + 1]));
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(e));
            asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(acc[i]), "=f"(acc[i + 1]) : "l"(d));
        }
    }
    long long t1 = clock64();
    float sum = 0;
#pragma unroll
    for (int i = 0; i < N; i++) sum += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <bool DENORM>
void run(const char* name) {
    float* out; long long* cyc; float* sx; uint32_t* seed;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sx, N * 4); cudaMalloc(&seed, N * 4);
    float hs[N]; uint32_t hseed[N];
    for (int i = 0; i < N; i++) { hs[i] = 1e-3f * (i + 1); hseed[i] = 2654435761u * (i + 7); }
    cudaMemcpy(sx, hs, sizeof(hs), cudaMemcpyHostToDevice); cudaMemcpy(seed, hseed, sizeof(hseed), cudaMemcpyHostToDevice);
    printf("%-40s", name);
    for (int wps : {1, 2, 4}) {
        const int threads = wps * 4 * 32;
        const float sw = DENORM ? 0.01f * 0x1p100f : 0.01f, bias = DENORM ? -2080800.f * (0.01f * 0x1p-49f) : -0.01f;
        k<DENORM><<<148, threads>>>(out, cyc, sw, bias, sx, seed);
        k<DENORM><<<148, threads>>>(out, cyc, sw, bias, sx, seed);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
        printf("  w/smsp=%d: %.3f FFMA2/clk", wps, (double)N * ITERS * wps / c);
    }
    printf("\n");
    cudaError_t e = cudaGetLastError(); if (e) printf("err %s\n", cudaGetErrorString(e));
}

int main() {
    // numerics: accumulator values B + 4S for S in a spread of magnitudes and both signs
    const int n = 4096;
    uint32_t* hacc = new uint32_t[n]; float* hout = new float[n];
    for (int i = 0; i < n; i++) {
        long long S = (long long)((i * 2654435761u) % 1040385u) - 520192;     // [-520192, 520192]
        if (i < 8) S = (i & 1) ? 520192 : -520192;
        if (i >= 8 && i < 40) S = i - 24;
        hacc[i] = (uint32_t)(2080800ll + 4 * S);
    }
    uint32_t* dacc; float* dout;
    cudaMalloc(&dacc, n * 4); cudaMalloc(&dout, n * 4);
    cudaMemcpy(dacc, hacc, n * 4, cudaMemcpyHostToDevice);
    const float sw = 0.0123f, sx = 0.0371f;
    numerics<<<n / 256, 256>>>(dacc, sw, sx, dout, n);
    cudaMemcpy(hout, dout, n * 4, cudaMemcpyDeviceToHost);
    double worst = 0; int bad = 0;
    for (int i = 0; i < n; i++) {
        const double S4 = (double)hacc[i] - 2080800.0;
        const double ref = 0.25 * S4 * (double)sw * (double)sx;       // the operands hold 4*w
        const double err = fabs(hout[i] - ref), tol = fabs(ref) * 3e-7 + 0.25 * 2080800.0 * sw * sx * 1.2e-7;   // c2 is rounded once
        if (err > tol) bad++;
        if (ref != 0 && err / fabs(ref) > worst && fabs(S4) > 4000) worst = err / fabs(ref);
    }
    printf("numerics: %d of %d outside tolerance; worst rel err for |S| > 1000: %.3g; S=0 -> %g (abs)\n", bad, n, worst, hout[24]);
    run<false>("FFMA2 pair, normal operands");
    run<true>("FFMA2 pair, subnormal first operand");
    return 0;
}
