// CLI drop-in for the reference's packer harness (engine/test_packing_kernel.cu): same argv
// (`test_packing_kernel M K X_BITS`), same checks, same stdout lines.
// Word-exact validation of the FlexQ plane layout against an ABQ-layout pack of the same ints,
// with the index relation of engine/test_packing_kernel.cu:139-141, then timing of both packers.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <vector>

#include "flexq_compat.h"

// ABQ-LLM layout u32[bit][M][K/32], element k%32==0 in bit 31 (baseline for the cross-check)
__global__ void abq_layout_pack_kernel(const int* __restrict__ in, unsigned* __restrict__ out, int m, int k) {
    const int bit = blockIdx.y;
    const int words = m * (k / 32);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < words; idx += gridDim.x * blockDim.x) {
        unsigned w = 0;
        for (int l = 0; l < 32; l++) w |= (unsigned)((in[(size_t)idx * 32 + l] >> bit) & 1) << (31 - l);
        out[(size_t)bit * words + idx] = w;
    }
}

static void abq_pack(const int* in, int* out, int m, int k, int bits, cudaStream_t s) {
    const int words = m * (k / 32);
    dim3 grid(std::min((words + 255) / 256, 4096), bits);
    abq_layout_pack_kernel<<<grid, 256, 0, s>>>(in, reinterpret_cast<unsigned*>(out), m, k);
}

template <typename F>
static float time_us(F&& f, cudaStream_t s, int warmup, int repeat) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < warmup; i++) f();
    cudaEventRecord(a, s);
    for (int i = 0; i < repeat; i++) f();
    cudaEventRecord(b, s);
    cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    return ms * 1e3f / repeat;
}

int main(int argc, char** argv) {
    if (argc < 4) { printf("Usage: ./test_packing_kernel M K X_BITS\n"); return -1; }
    const int m = atoi(argv[1]), k = atoi(argv[2]), bits = atoi(argv[3]);
    if (k < 128 || k % 128 != 0) { printf("Unsupported computational layout! k must >= 128 and k %% 128 == 0!\n"); return -1; }
    if (bits < 1 || bits > 8) { printf("unsupport x_bits %d: for bit_pack func \n", bits); return -1; }
    const char* seed = getenv("FLEXQ_SEED");
    srand(seed ? atoi(seed) : (unsigned)time(0));
    cudaStream_t stream; cudaStreamCreate(&stream);
    const size_t n = (size_t)m * k, words = n / 32 * bits;
    std::vector<int> h(n), h_abq(words), h_fq(words);
    for (auto& v : h) v = rand() % (1 << bits);
    int *d_x, *d_abq, *d_fq;
    if (cudaMalloc(&d_x, n * 4) || cudaMalloc(&d_abq, words * 4) || cudaMalloc(&d_fq, words * 4)) { printf("alloc failed\n"); return -1; }
    cudaMemcpy(d_x, h.data(), n * 4, cudaMemcpyHostToDevice);
    abq_pack(d_x, d_abq, m, k, bits, stream);
    cudaError_t err = flexq_bit_packing(d_x, d_fq, m, k, bits, stream);
    if (err != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) {
        printf("Line %d: 'activation bit_pack' failed: %s\n", __LINE__, cudaGetErrorString(err != cudaSuccess ? err : cudaGetLastError()));
        return -1;
    }
    const float t_abq = time_us([&] { abq_pack(d_x, d_abq, m, k, bits, stream); }, stream, 10, 1000);
    const float t_fq = time_us([&] { flexq_bit_packing(d_x, d_fq, m, k, bits, stream); }, stream, 10, 1000);
    cudaMemcpy(h_abq.data(), d_abq, words * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(h_fq.data(), d_fq, words * 4, cudaMemcpyDeviceToHost);
    bool ok = true;
    const int chunk = std::min(m, 8);
    for (int b = 0; b < bits && ok; b++)
        for (int r = 0; r < m && ok; r++)
            for (int w = 0; w < k / 32; w++) {
                const size_t fq = (size_t)(w / 4) * ((size_t)m * bits * 4) + (size_t)(r / chunk) * (bits * chunk * 4) + (size_t)b * (chunk * 4) +
                                  (size_t)(r % chunk) * 4 + w % 4;
                if (h_abq[(size_t)b * (n / 32) + (size_t)r * (k / 32) + w] != h_fq[fq]) { ok = false; break; }
            }
    printf(ok ? "FlexQ bit packing kernel SUCCESS! consistent results!\n" : "FlexQ bit packing kernel ERROR! Inconsistent results!\n");
    printf("\nKernel performance:\n");
    printf("ABQ packing %f (us) exec\n", t_abq);
    printf("FlexQ bit packing %f (us) exec\n", t_fq);
    cudaFree(d_x); cudaFree(d_abq); cudaFree(d_fq);
    cudaStreamDestroy(stream);
    return ok ? 0 : 1;
}
