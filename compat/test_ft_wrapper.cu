// Exercises the FasterTransformer adapter surface (e2e/.../flexqgemm/flexq_gemm_wrapper.h:6-48):
// FLEXQGEMMWrapper::pack + gemm(int* A ...) must agree with the fused gemm(half* A ...) overload,
// exactly as in the reference where the half overload is pack-then-gemm (flexq_gemm_wrapper.cu:99-122).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "flexq_compat.h"

int main(int argc, char** argv) {
    const int M = argc > 1 ? atoi(argv[1]) : 8, N = argc > 2 ? atoi(argv[2]) : 4096, K = argc > 3 ? atoi(argv[3]) : 4096;
    const int xb = argc > 4 ? atoi(argv[4]) : 6;
    srand(1);
    std::vector<half> hx((size_t)M * K), hws((size_t)(K / 128) * N);
    std::vector<int> hw((size_t)N * K);
    for (auto& v : hx) v = __float2half((rand() / (float)RAND_MAX - 0.5f) * 4.f);
    for (auto& v : hw) v = rand() % 64 - 32;
    for (auto& v : hws) v = __float2half(0.01f + 0.02f * rand() / RAND_MAX);
    half *dx, *dws, *d1, *d2, *dxs; int *dw, *dxp; uint8_t* w6; char* ws;
    const size_t ws_bytes = flexq_linear_workspace_bytes(M, K);
    cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&dws, hws.size() * 2); cudaMalloc(&d1, (size_t)M * N * 2); cudaMalloc(&d2, (size_t)M * N * 2);
    cudaMalloc(&dxs, flexq_xscale_ref_halves(M, K) * 2); cudaMalloc(&dw, hw.size() * 4); cudaMalloc(&dxp, flexq_planes_bytes(M, K, xb));
    cudaMalloc(&w6, flexq_w6_packed_bytes(N, K)); cudaMalloc(&ws, ws_bytes);
    cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dws, hws.data(), hws.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
    flexq_workspace_init(ws, ws_bytes, nullptr);
    if (flexq_pack_w6_i32(dw, w6, N, K, nullptr)) { printf("pack failed\n"); return 1; }
    FLEXQGEMMWrapper wrap(xb, 6, true);
    wrap.gemm(M, N, K, dx, reinterpret_cast<const int*>(w6), nullptr, d1, nullptr, reinterpret_cast<const float*>(dws), nullptr, nullptr, false,
              ws, ws_bytes, nullptr);
    wrap.pack(dx, dxp, dxs, M, K, xb, nullptr);
    wrap.gemm(M, N, K, dxp, reinterpret_cast<const int*>(w6), nullptr, d2, reinterpret_cast<float*>(dxs), reinterpret_cast<const float*>(dws), nullptr,
              nullptr, false, ws, ws_bytes, nullptr);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("FT wrapper ERROR: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    std::vector<half> h1((size_t)M * N), h2((size_t)M * N);
    cudaMemcpy(h1.data(), d1, h1.size() * 2, cudaMemcpyDeviceToHost);
    cudaMemcpy(h2.data(), d2, h2.size() * 2, cudaMemcpyDeviceToHost);
    double maxd = 0, maxv = 0;
    for (size_t i = 0; i < h1.size(); i++) {
        maxd = fmax(maxd, fabs(__half2float(h1[i]) - __half2float(h2[i])));
        maxv = fmax(maxv, fabs(__half2float(h1[i])));
    }
    const bool ok = maxv > 0 && maxd <= maxv * 2e-3;     // same integers and scales; only fp32 atomic order may differ
    printf("FT wrapper %s max|d|=%g max|v|=%g\n", ok ? "SUCCESS" : "ERROR", maxd, maxv);
    return ok ? 0 : 1;
}
