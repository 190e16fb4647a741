// Implementation of the source-compatible reference interface on top of the flexq_b200 C ABI.
#include "flexq_compat.h"

#include <cstdio>
#include <map>
#include <mutex>

namespace {
struct CachedWeight {
    uint8_t* w6 = nullptr;
    int N = 0, K = 0;
};
struct StreamWorkspace {
    void* ws = nullptr;
    size_t bytes = 0;
};
std::map<const int*, CachedWeight> g_cache;          // converted W6 tiles by reference weight pointer (flexq_compat_invalidate)
std::map<cudaStream_t, StreamWorkspace> g_ws;        // one workspace per stream: Execs on one stream run one after the other
std::mutex g_mu;

cudaError_t to_cuda(int st) { return st > 0 ? (cudaError_t)st : (st ? cudaErrorInvalidValue : cudaSuccess); }

// workspace of the stream, grown to the problem at hand (no fixed token limit); zeroed when (re)allocated
void* stream_workspace(cudaStream_t stream, size_t need, size_t* bytes) {
    StreamWorkspace& w = g_ws[stream];
    if (w.bytes < need) {
        if (w.ws) { cudaStreamSynchronize(stream); cudaFree(w.ws); w.ws = nullptr; w.bytes = 0; }
        if (cudaMalloc(&w.ws, need) != cudaSuccess) return nullptr;
        if (flexq_workspace_init(w.ws, need, stream)) { cudaFree(w.ws); w.ws = nullptr; return nullptr; }
        w.bytes = need;
    }
    *bytes = w.bytes;
    return w.ws;
}

FQBMMAOpState init_common(int xb, int* X, int* W, half* XS, half* WS, int M, int N, int K, half* D, int group, bool bias) {
    FQBMMAOpState st{};
    st.args.M = M; st.args.N = N; st.args.K = K; st.args.X = X; st.args.W = W;
    st.args.X_SCALE = XS; st.args.W_SCALE = WS; st.args.D = D; st.args.group_size = group; st.args.bias = bias;
    st.x_bits = xb;
    st.shared_mem_size = 0; st.gridDim = dim3(1); st.blockDim = dim3(512);
    if (bias || group != FLEXQ_GROUP || K < 128 || K % 128 || N % 8 || M <= 0) {
        fprintf(stderr, "flexq_compat: unsupported problem (bias=%d group=%d M=%d N=%d K=%d)\n", (int)bias, group, M, N, K);
        return st;                       // initSuccess stays false, like the reference on a bad config
    }
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(W);
    if (it == g_cache.end() || it->second.N != N || it->second.K != K) {
        CachedWeight c; c.N = N; c.K = K;
        int8_t* scratch = nullptr;
        if (cudaMalloc(&c.w6, flexq_w6_packed_bytes(N, K)) != cudaSuccess || cudaMalloc(&scratch, (size_t)N * K) != cudaSuccess) return st;
        const int rc = flexq_planes_to_w6(W, c.w6, scratch, N, K, nullptr);
        cudaDeviceSynchronize();
        cudaFree(scratch);
        if (rc) { fprintf(stderr, "flexq_compat: %s\n", flexq_status_string(rc)); return st; }
        if (it != g_cache.end()) cudaFree(it->second.w6);
        g_cache[W] = c;
    }
    st.initSuccess = true;
    return st;
}
}  // namespace

cudaError_t flexq_bit_packing(const int* in, int* packed, const int M, const int K, const int BIT, cudaStream_t stream) {
    return to_cuda(flexq_bit_packing_i32(in, packed, M, K, BIT, stream));
}

void flexq_bit_packing(const half* in, int* packed, half* scale, const int M, const int K, const int BIT, cudaStream_t stream) {
    int st = flexq_bit_packing_f16(in, packed, scale, M, K, BIT, stream);
    if (st) printf("[FlexQ][Error] %s\n", flexq_status_string(st));
}

FQBMMAOpState FQBMMA_W6A6_InitFn(int* X, int* W, half* XS, half* WS, int M, int N, int K, half* D, int group, bool bias) {
    return init_common(6, X, W, XS, WS, M, N, K, D, group, bias);
}
FQBMMAOpState FQBMMA_W6A8_InitFn(int* X, int* W, half* XS, half* WS, int M, int N, int K, half* D, int group, bool bias) {
    return init_common(8, X, W, XS, WS, M, N, K, D, group, bias);
}

void FQBMMA_ExecFn(FQBMMAOpState& st, cudaStream_t stream) {
    CachedWeight c;
    void* ws = nullptr;
    size_t ws_bytes = 0;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_cache.find(st.args.W);
        if (!st.initSuccess || it == g_cache.end()) { fprintf(stderr, "flexq_compat: Exec on an uninitialised state\n"); return; }
        c = it->second;
        ws = stream_workspace(stream, flexq_linear_workspace_bytes(st.args.M, st.args.K), &ws_bytes);
        if (!ws) { fprintf(stderr, "flexq_compat: cannot allocate the workspace\n"); return; }
    }
    int rc = flexq_gemm_ref_layout(st.args.X, st.args.X_SCALE, c.w6, st.args.W_SCALE, st.args.D, st.args.M, st.args.N, st.args.K,
                                   st.x_bits, ws, ws_bytes, stream);
    if (rc) fprintf(stderr, "flexq_compat: %s\n", flexq_status_string(rc));
}

void flexq_compat_invalidate(const int* W) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(W);
    if (it != g_cache.end()) { cudaDeviceSynchronize(); cudaFree(it->second.w6); g_cache.erase(it); }
}

void flexq_compat_release() {
    std::lock_guard<std::mutex> lk(g_mu);
    cudaDeviceSynchronize();
    for (auto& kv : g_cache) cudaFree(kv.second.w6);
    for (auto& kv : g_ws) cudaFree(kv.second.ws);
    g_cache.clear();
    g_ws.clear();
}

FLEXQGEMMWrapper::FLEXQGEMMWrapper(int X_BITS, int W_BITS, bool SIGNED) : x_bits_(X_BITS), w_bits_(W_BITS), signed_(SIGNED) {}
FLEXQGEMMWrapper::~FLEXQGEMMWrapper() {}

void FLEXQGEMMWrapper::pack(const half* in, int* packed, half* x_scale, int M, int K, int BIT, cudaStream_t stream) {
    flexq_bit_packing(in, packed, x_scale, M, K, BIT, stream);
}

void FLEXQGEMMWrapper::gemm(const int M, const int N, const int K, const int* A, const int* B, const half*, half* D, float* x_scale,
                            const float* w_scale, const float*, const float*, bool bias, char* ws, size_t ws_bytes, cudaStream_t stream) {
    if (w_bits_ != 6 || (x_bits_ != 6 && x_bits_ != 8) || !signed_ || bias) { printf("[FlexQ][Error] unsupport w%da%d\n", w_bits_, x_bits_); return; }
    int st = flexq_gemm_ref_layout(A, x_scale, reinterpret_cast<const uint8_t*>(B), w_scale, D, M, N, K, x_bits_, ws, ws_bytes, stream);
    if (st) printf("[FlexQ][Error] %s\n", flexq_status_string(st));
}

void FLEXQGEMMWrapper::gemm(const int M, const int N, const int K, const half* A, const int* B, const half*, half* D, float*,
                            const float* w_scale, const float*, const float*, bool bias, char* ws, size_t ws_bytes, cudaStream_t stream) {
    if (w_bits_ != 6 || (x_bits_ != 6 && x_bits_ != 8) || !signed_ || bias) { printf("[FlexQ][Error] unsupport w%da%d\n", w_bits_, x_bits_); return; }
    int st = flexq_linear_w6ax_f16(A, reinterpret_cast<const uint8_t*>(B), w_scale, D, M, N, K, x_bits_, FLEXQ_ROUND_CUDA, ws, ws_bytes, stream);
    if (st) printf("[FlexQ][Error] %s\n", flexq_status_string(st));
}
