// CLI drop-in for the reference's GEMM harness (engine/test_bgemm_kernel.cu + test/test_kernel.h):
// `test_bgemm_kernel M N K X_BITS W_BITS [debug]`.  Random ints in [0,2^bits) (interpreted as two's
// complement), random half scales in the reference layouts, pack W and X with flexq_bit_packing,
// CPU golden with the semantics of compute_ref (test_bgemm_kernel.cu:113-146), run the kernel
// through the FQBMMA Init/Exec pair, tolerance 1e-4*65504 (test_kernel.h:59-69), and the
// reference's stdout lines.  One line per sm_100a kernel family member.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <vector>

#include "flexq_compat.h"

static inline int ceil4(int m) { return (m + 3) / 4 * 4; }

// CPU golden: same value as compute_ref -- sum over plane pairs of (+-2^(i+j)) popc(x_i & w_j) sx sw --
// evaluated from the *packed* words: planes -> two's-complement ints -> per-group integer dot
// products -> fp32 accumulation of S * sw * sx -> half.
static void unpack_planes(const std::vector<int>& words, std::vector<int>& out, int R, int K, int bits) {
    const int chunk = std::min(R, 8);
    out.assign((size_t)R * K, 0);
    for (int b = 0; b < bits; b++)
        for (int r = 0; r < R; r++)
            for (int w = 0; w < K / 32; w++) {
                const size_t idx = (size_t)(w / 4) * ((size_t)R * bits * 4) + (size_t)(r / chunk) * (bits * chunk * 4) + (size_t)b * (chunk * 4) +
                                   (size_t)(r % chunk) * 4 + w % 4;
                const unsigned word = (unsigned)words[idx];
                const int weight = (b == bits - 1) ? -(1 << b) : (1 << b);
                for (int l = 0; l < 32; l++)
                    if ((word >> (31 - l)) & 1u) out[(size_t)r * K + w * 32 + l] += weight;
            }
}

static void compute_ref(const std::vector<int>& wp, const std::vector<half>& ws, const std::vector<int>& xp, const std::vector<half>& xs,
                        std::vector<half>& ref, int M, int N, int K, int w_bits, int x_bits) {
    std::vector<int> xi, wi;
    unpack_planes(xp, xi, M, K, x_bits);
    unpack_planes(wp, wi, N, K, w_bits);
    const int G = K / 128, ld = 2 * ceil4(M);
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            float acc = 0.f;
            for (int g = 0; g < G; g++) {
                int s = 0;
                const int* a = &xi[(size_t)m * K + g * 128];
                const int* b = &wi[(size_t)n * K + g * 128];
                for (int k = 0; k < 128; k++) s += a[k] * b[k];
                acc += (float)s * __half2float(ws[(size_t)g * N + n]) * __half2float(xs[(size_t)g * ld + 2 * m]);
            }
            ref[(size_t)m * N + n] = __float2half(acc);
        }
}

template <typename F>
static float time_ms(F&& f, cudaStream_t s, int warmup, int repeat) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < warmup; i++) f();
    cudaEventRecord(a, s);
    for (int i = 0; i < repeat; i++) f();
    cudaEventRecord(b, s);
    cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    return ms / repeat;
}

int main(int argc, char** argv) {
    if (argc < 6) { printf("Usage: ./test_bgemm_kernel M N K X_BITS W_BITS\n"); return -1; }
    const int m = atoi(argv[1]), n = atoi(argv[2]), k = atoi(argv[3]), x_bits = atoi(argv[4]), w_bits = atoi(argv[5]);
    const int group_size = 128, repeat = 1000, warmup = 10;
    if (k < 128 || k % 128 != 0) { printf("Unsupported computational layout! k must >= 128 and k %% 128 == 0!\n"); return -1; }
    if (w_bits != 6 || (x_bits != 6 && x_bits != 8)) { printf("unsupport w%da%d!\n", w_bits, x_bits); return 0; }
    const char* seed = getenv("FLEXQ_SEED");
    srand(seed ? atoi(seed) : (unsigned)time(0));
    cudaStream_t stream; cudaStreamCreate(&stream);

    const int G = k / group_size, ldx = 2 * ceil4(m);
    std::vector<int> h_x((size_t)m * k), h_w((size_t)n * k);
    for (auto& v : h_x) v = rand() % (1 << x_bits);
    for (auto& v : h_w) v = rand() % (1 << w_bits);
    std::vector<half> h_xs((size_t)G * ldx, __float2half(0.f)), h_ws((size_t)G * n);
    for (int g = 0; g < G; g++)
        for (int r = 0; r < m; r++) h_xs[(size_t)g * ldx + 2 * r] = h_xs[(size_t)g * ldx + 2 * r + 1] = __float2half(0.1f * rand() / RAND_MAX);
    for (auto& v : h_ws) v = __float2half(0.1f * rand() / RAND_MAX);

    int *d_x, *d_w, *d_xp, *d_wp; half *d_xs, *d_ws, *d_out;
    const size_t xw = (size_t)x_bits * m * (k / 32), ww = (size_t)w_bits * n * (k / 32);
    if (cudaMalloc(&d_x, h_x.size() * 4) || cudaMalloc(&d_w, h_w.size() * 4) || cudaMalloc(&d_xp, xw * 4) || cudaMalloc(&d_wp, ww * 4) ||
        cudaMalloc(&d_xs, h_xs.size() * 2) || cudaMalloc(&d_ws, h_ws.size() * 2) || cudaMalloc(&d_out, (size_t)m * n * 2)) { printf("alloc failed\n"); return -1; }
    cudaMemcpy(d_x, h_x.data(), h_x.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_w, h_w.data(), h_w.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_xs, h_xs.data(), h_xs.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(d_ws, h_ws.data(), h_ws.size() * 2, cudaMemcpyHostToDevice);
    if (flexq_bit_packing(d_w, d_wp, n, k, w_bits, stream) != cudaSuccess) { printf("Line %d: 'weight bit_pack' failed\n", __LINE__); return -1; }
    if (flexq_bit_packing(d_x, d_xp, m, k, x_bits, stream) != cudaSuccess) { printf("Line %d: 'activation bit_pack' failed\n", __LINE__); return -1; }
    std::vector<int> h_xp(xw), h_wp(ww);
    cudaMemcpy(h_xp.data(), d_xp, xw * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(h_wp.data(), d_wp, ww * 4, cudaMemcpyDeviceToHost);
    std::vector<half> h_ref((size_t)m * n), h_out((size_t)m * n);
    compute_ref(h_wp, h_ws, h_xp, h_xs, h_ref, m, n, k, w_bits, x_bits);

    printf(x_bits == 6 ? "test_w6a6_kernel\n" : "test_w6a8_kernel\n");
    // advisory tile fields of the sm_100a kernel that serves this M (weights tile 128 x 128k, tokens tile M_TILE)
    const int mt = m <= 16 ? 16 : m <= 32 ? 32 : m <= 64 ? 64 : m <= 128 ? 128 : 192;
    const int gp = m <= 32 ? 4 : m <= 64 ? 2 : 1;
    printf("100 %d %d %d %d %d sign %d %d %d %d %d %d %d %d %d %d %d ", x_bits, w_bits, m, n, k, mt, 128, 128 * gp, mt, 128, 128, mt, 128, 32, gp, 1);
    FQBMMAInitFn_t init_fn = x_bits == 8 ? FQBMMA_W6A8_InitFn : FQBMMA_W6A6_InitFn;
    FQBMMAOpState st = (*init_fn)(d_xp, d_wp, d_xs, d_ws, m, n, k, d_out, group_size, false);
    int ret = 0;
    float exec_ms = 0, pack_ms = 0;
    if (!st.initSuccess) { ret = -1; }
    else {
        FQBMMA_ExecFn(st, stream);
        if (cudaStreamSynchronize(stream) != cudaSuccess) ret = -1;
    }
    if (ret == 0) {
        exec_ms = time_ms([&] { FQBMMA_ExecFn(st, stream); }, stream, warmup, repeat);
        pack_ms = time_ms([&] { flexq_bit_packing(d_x, d_xp, m, k, x_bits, stream); }, stream, warmup, repeat);
        cudaMemcpy(h_out.data(), d_out, h_out.size() * 2, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < h_out.size(); i++)
            if (fabsf(__half2float(h_ref[i]) - __half2float(h_out[i])) > 0.0001f * 65504.f) { ret = -2; break; }
    }
    const float gop = (float)m / 1e9f * n * k * 2, bgop = gop * x_bits * w_bits;
    printf("packing %f (us) exec %f (us) %f TOPS | %f B-TOPS | %s\n", pack_ms * 1e3, exec_ms * 1e3, exec_ms > 0 ? gop / exec_ms : 0.f,
           exec_ms > 0 ? bgop / exec_ms : 0.f, ret == 0 ? "PASSED" : ret == -1 ? "ERROR" : "FAILED");
    printf("The best kernel config is %d, %d, %d, %d, %d, %d, %d, %d, %d, %d, %d with %f TOPS\n", mt, 128, 128 * gp, mt, 128, 128, mt, 128, 32, gp, 1,
           ret == 0 ? gop / exec_ms : 0.f);
    printf(ret == 0 ? "SUCCESS! consistent results!\n" : "ERROR! Inconsistent results!\n");
    flexq_compat_release();
    cudaFree(d_x); cudaFree(d_w); cudaFree(d_xp); cudaFree(d_wp); cudaFree(d_xs); cudaFree(d_ws); cudaFree(d_out);
    cudaStreamDestroy(stream);
    return ret == 0 ? 0 : 1;
}
