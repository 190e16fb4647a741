// CLI drop-in for the reference's cuBLAS INT8 baseline harness (engine/test_cublas_kernel.cu):
// `test_cublas_kernel M N K`, cublasGemmEx 8I x 8I -> 32I (CUBLAS_COMPUTE_32I, tensor op), 1000
// timed iterations, one stdout line in the reference's format (:154-155).
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { if ((x) != 0) { printf("error at %s:%d\n", __FILE__, __LINE__); exit(EXIT_FAILURE); } } while (0)

int main(int argc, char** argv) {
    if (argc != 4) { printf("Usage: %s M N K\n", argv[0]); return -1; }
    const int m = atoi(argv[1]), n = atoi(argv[2]), k = atoi(argv[3]);
    const int iters = 1000;
    std::vector<int8_t> ha((size_t)m * k), hb((size_t)k * n);
    srand(0x2019);
    for (auto& v : ha) v = (int8_t)(rand() % 256 - 128);
    for (auto& v : hb) v = (int8_t)(rand() % 256 - 128);
    int8_t *da, *db; int32_t* dc;
    CK(cudaMalloc(&da, ha.size())); CK(cudaMalloc(&db, hb.size())); CK(cudaMalloc(&dc, (size_t)m * n * 4));
    CK(cudaMemcpy(da, ha.data(), ha.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), hb.size(), cudaMemcpyHostToDevice));
    cublasHandle_t h; CK(cublasCreate(&h));
    const int32_t alpha = 1, beta = 0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    for (int i = 0; i < iters; i++)      // row-major C[m][n] = A[m][k] B[k][n] expressed column-major, as the reference does
        CK(cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_N, n, m, k, &alpha, db, CUDA_R_8I, n, da, CUDA_R_8I, k, &beta, dc, CUDA_R_32I, n,
                        CUBLAS_COMPUTE_32I, CUBLAS_GEMM_DEFAULT_TENSOR_OP));
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double avg = ms / iters, tflops = 2.0 * m * n * k * 1e-12 / (avg / 1000.0);
    printf("cuBLAS-W8A8-GEMM. m: %6d, n: %6d, k: %6d,\t Time: %.4f ms, TFLOPS: %4.4f\n", m, n, k, avg, tflops);
    cublasDestroy(h); cudaFree(da); cudaFree(db); cudaFree(dc);
    return 0;
}
