// extern "C" boundary of libflexq_b200.so -- see include/flexq_b200.h for the contract and the
// reference interfaces each entry point replaces.  No torch types, no allocation, no sync.
#include "common.cuh"

namespace flexq {
int quant_act_native(const __half*, int8_t*, float*, int, int, int, int, cudaStream_t);
int quant_act_planes(const __half*, uint32_t*, __half*, int, int, int, cudaStream_t);
int quant_act_f32(const float*, int8_t*, float*, int, int, int, cudaStream_t);
int debug_schedule(int, int, int, int, int, int*, int, int*);
template <typename T> int pack_w6(const T*, uint8_t*, int, int, cudaStream_t);
template <typename T> int quant_pack_w6(const T*, uint8_t*, __half*, int, int, cudaStream_t);
int unpack_w6(const uint8_t*, int8_t*, int, int, cudaStream_t);
int pack_planes_i32(const int32_t*, uint32_t*, int, int, int, cudaStream_t);
int planes_to_i8(const uint32_t*, int8_t*, int, int, int, cudaStream_t);
int xscale_ref_to_sx(const __half*, float*, int, int, cudaStream_t);
int gemm_w6ax(const int8_t*, const float*, const uint8_t*, const __half*, __half*, int, int, int, void*, size_t, cudaStream_t);
int gemm_w6ax_silu_mul(const int8_t*, const float*, const uint8_t*, const __half*, __half*, int, int, int, void*, size_t, cudaStream_t);
int gemm_w6ax_groupsums(const int8_t*, const uint8_t*, int32_t*, int, int, int, cudaStream_t);
int rmsnorm_quant(const __half*, __half*, const __half*, float, __half*, int8_t*, float*, int, int, int, cudaStream_t);
int silu_mul_quant(const __half*, const __half*, long long, __half*, int8_t*, float*, int, int, int, cudaStream_t);
int allreduce_oneshot_f16(void* const*, void* const*, size_t, int, int, void*, cudaStream_t);
void set_sm_limit(int);
void set_allreduce_blocks(int);
int allreduce_sum_f16(void*, void* const*, void* const*, size_t, size_t, int, int, cudaStream_t);
int gemm_w6ax_trace(const int8_t*, const float*, const uint8_t*, const __half*, __half*, int, int, int, void*, long long*, int, cudaStream_t);
}  // namespace flexq

using namespace flexq;

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" {

int flexq_version(void) { return 100; }

const char* flexq_status_string(int s) {
    switch (s) {
        case FLEXQ_OK: return "ok";
        case FLEXQ_ERR_BAD_SHAPE: return "bad shape (need K % 128 == 0, K >= 128, positive M/N; plane layouts need R % min(R,8) == 0)";
        case FLEXQ_ERR_BAD_BITS: return "unsupported bit width";
        case FLEXQ_ERR_NULL: return "null pointer";
        case FLEXQ_ERR_WORKSPACE: return "workspace too small or misaligned";
        case FLEXQ_ERR_NO_DEVICE: return "no usable sm_100 device / driver entry point";
        case FLEXQ_ERR_TENSORMAP: return "cuTensorMapEncodeTiled failed";
        default: return s > 0 ? cudaGetErrorString((cudaError_t)s) : "unknown flexq status";
    }
}

size_t flexq_w6_packed_bytes(int N, int K) { return (size_t)ceil_div(N, kTileN) * (size_t)(K / kGroup) * kTileBytes; }
size_t flexq_planes_bytes(int R, int K, int bits) { return (size_t)R * K / 8 * bits; }
int flexq_sx_ld(int M) { return ceil4(M); }
size_t flexq_xscale_ref_halves(int M, int K) { return (size_t)(K / kGroup) * 2 * ceil4(M); }
size_t flexq_gemm_workspace_bytes(void) { return kGemmWorkspaceBytes; }
size_t flexq_linear_workspace_bytes(int M, int K) {
    return flexq_gemm_workspace_bytes() + align_up((size_t)M * K, 256) + align_up((size_t)(K / kGroup) * ceil4(M) * sizeof(float), 256);
}

int flexq_workspace_init(void* ws, size_t bytes, void* stream) {
    if (!ws) return FLEXQ_ERR_NULL;
    return (int)cudaMemsetAsync(ws, 0, bytes, (cudaStream_t)stream);
}

int flexq_bit_packing_i32(const int32_t* in, int32_t* planes, int R, int K, int bits, void* stream) {
    return pack_planes_i32(in, reinterpret_cast<uint32_t*>(planes), R, K, bits, (cudaStream_t)stream);
}

int flexq_bit_packing_f16(const void* x, int32_t* planes, void* xs, int M, int K, int bits, void* stream) {
    return quant_act_planes((const __half*)x, reinterpret_cast<uint32_t*>(planes), (__half*)xs, M, K, bits, (cudaStream_t)stream);
}

int flexq_quant_act(const void* x, int8_t* xq, float* sx, int M, int K, int bits, int mode, void* stream) {
    return quant_act_native((const __half*)x, xq, sx, M, K, bits, mode, (cudaStream_t)stream);
}

int flexq_debug_schedule(int m_tiles, int n_tiles, int groups, int max_ctas, int cta, int* segments, int cap, int* n_ctas) {
    return debug_schedule(m_tiles, n_tiles, groups, max_ctas, cta, segments, cap, n_ctas);
}

int flexq_quant_act_f32(const float* x, int8_t* xq, float* sx, int M, int K, int bits, void* stream) {
    return quant_act_f32(x, xq, sx, M, K, bits, (cudaStream_t)stream);
}

int flexq_pack_w6_i32(const int32_t* w, uint8_t* w6, int N, int K, void* stream) { return pack_w6<int32_t>(w, w6, N, K, (cudaStream_t)stream); }
int flexq_pack_w6_i8(const int8_t* w, uint8_t* w6, int N, int K, void* stream) { return pack_w6<int8_t>(w, w6, N, K, (cudaStream_t)stream); }
int flexq_quant_pack_w6_f16(const void* w, uint8_t* w6, void* ws, int N, int K, void* stream) {
    return quant_pack_w6<__half>((const __half*)w, w6, (__half*)ws, N, K, (cudaStream_t)stream);
}
int flexq_quant_pack_w6_f32(const float* w, uint8_t* w6, void* ws, int N, int K, void* stream) {
    return quant_pack_w6<float>(w, w6, (__half*)ws, N, K, (cudaStream_t)stream);
}

int flexq_planes_to_i8(const int32_t* planes, int8_t* out, int R, int K, int bits, void* stream) {
    return planes_to_i8(reinterpret_cast<const uint32_t*>(planes), out, R, K, bits, (cudaStream_t)stream);
}
int flexq_planes_to_w6(const int32_t* planes, uint8_t* w6, int8_t* scratch, int N, int K, void* stream) {
    if (!scratch) return FLEXQ_ERR_NULL;
    int e = planes_to_i8(reinterpret_cast<const uint32_t*>(planes), scratch, N, K, 6, (cudaStream_t)stream);
    if (e) return e;
    return pack_w6<int8_t>(scratch, w6, N, K, (cudaStream_t)stream);
}
int flexq_xscale_ref_to_sx(const void* xs, float* sx, int M, int K, void* stream) {
    return xscale_ref_to_sx((const __half*)xs, sx, M, K, (cudaStream_t)stream);
}
int flexq_w6_to_i8(const uint8_t* w6, int8_t* out, int N, int K, void* stream) { return unpack_w6(w6, out, N, K, (cudaStream_t)stream); }

int flexq_gemm_w6ax(const int8_t* xq, const float* sx, const uint8_t* w6, const void* w_scale, void* d, int M, int N, int K,
                    void* ws, size_t ws_bytes, void* stream) {
    return gemm_w6ax(xq, sx, w6, (const __half*)w_scale, (__half*)d, M, N, K, ws, ws_bytes, (cudaStream_t)stream);
}

int flexq_gemm_w6ax_silu_mul(const int8_t* xq, const float* sx, const uint8_t* w6_gate_up, const void* w_scale_gate_up, void* h, int M,
                             int inter, int K, void* ws, size_t ws_bytes, void* stream) {
    return gemm_w6ax_silu_mul(xq, sx, w6_gate_up, (const __half*)w_scale_gate_up, (__half*)h, M, inter, K, ws, ws_bytes, (cudaStream_t)stream);
}

int flexq_debug_gemm_trace(const int8_t* xq, const float* sx, const uint8_t* w6, const void* w_scale, void* d, int M, int N, int K,
                           void* ws, long long* trace, int trace_units, void* stream) {
    return gemm_w6ax_trace(xq, sx, w6, (const __half*)w_scale, (__half*)d, M, N, K, ws, trace, trace_units, (cudaStream_t)stream);
}

int flexq_gemm_w6ax_groupsums(const int8_t* xq, const uint8_t* w6, int32_t* S, int M, int N, int K, void* stream) {
    return gemm_w6ax_groupsums(xq, w6, S, M, N, K, (cudaStream_t)stream);
}

// workspace layout for the fused / reference-layout entries: [gemm workspace][Xq][sx]
static int carve(void* ws, size_t ws_bytes, int M, int K, int8_t** xq, float** sx) {
    if (!ws) return FLEXQ_ERR_NULL;
    if (M <= 0 || K < kGroup || K % kGroup) return FLEXQ_ERR_BAD_SHAPE;
    if (ws_bytes < flexq_linear_workspace_bytes(M, K) || ((uintptr_t)ws & 255)) return FLEXQ_ERR_WORKSPACE;
    uint8_t* base = reinterpret_cast<uint8_t*>(ws) + flexq_gemm_workspace_bytes();
    *xq = reinterpret_cast<int8_t*>(base);
    *sx = reinterpret_cast<float*>(base + align_up((size_t)M * K, 256));
    return 0;
}

int flexq_linear_w6ax_f16(const void* x, const uint8_t* w6, const void* w_scale, void* d, int M, int N, int K, int x_bits,
                          int mode, void* ws, size_t ws_bytes, void* stream) {
    int8_t* xq; float* sx;
    if (int e = carve(ws, ws_bytes, M, K, &xq, &sx)) return e;
    if (int e = quant_act_native((const __half*)x, xq, sx, M, K, x_bits, mode, (cudaStream_t)stream)) return e;
    return gemm_w6ax(xq, sx, w6, (const __half*)w_scale, (__half*)d, M, N, K, ws, flexq_gemm_workspace_bytes(), (cudaStream_t)stream);
}

int flexq_gemm_ref_layout(const int32_t* x_planes, const void* x_scale, const uint8_t* w6, const void* w_scale, void* d,
                          int M, int N, int K, int x_bits, void* ws, size_t ws_bytes, void* stream) {
    int8_t* xq; float* sx;
    if (x_bits != 6 && x_bits != 8) return FLEXQ_ERR_BAD_BITS;
    if (int e = carve(ws, ws_bytes, M, K, &xq, &sx)) return e;
    if (int e = planes_to_i8(reinterpret_cast<const uint32_t*>(x_planes), xq, M, K, x_bits, (cudaStream_t)stream)) return e;
    if (int e = xscale_ref_to_sx((const __half*)x_scale, sx, M, K, (cudaStream_t)stream)) return e;
    return gemm_w6ax(xq, sx, w6, (const __half*)w_scale, (__half*)d, M, N, K, ws, flexq_gemm_workspace_bytes(), (cudaStream_t)stream);
}

int flexq_rmsnorm_quant_f16(const void* x, void* residual, const void* gamma, float eps, void* normed, int8_t* xq, float* sx, int M,
                            int K, int bits, void* stream) {
    return rmsnorm_quant((const __half*)x, (__half*)residual, (const __half*)gamma, eps, (__half*)normed, xq, sx, M, K, bits,
                         (cudaStream_t)stream);
}
int flexq_silu_mul_quant_f16(const void* gate, const void* up, long long ld_in, void* out, int8_t* xq, float* sx, int M, int K,
                             int bits, void* stream) {
    return silu_mul_quant((const __half*)gate, (const __half*)up, ld_in, (__half*)out, xq, sx, M, K, bits, (cudaStream_t)stream);
}

int flexq_allreduce_oneshot_f16(void* const* data_ptrs, void* const* flag_ptrs, size_t elems, int rank, int world, void* out,
                                void* stream) {
    return allreduce_oneshot_f16(data_ptrs, flag_ptrs, elems, rank, world, out, (cudaStream_t)stream);
}
int flexq_set_sm_limit(int n_ctas) { set_sm_limit(n_ctas); return 0; }
int flexq_set_allreduce_blocks(int n_blocks) { set_allreduce_blocks(n_blocks); return 0; }

int flexq_allreduce_sum_f16(void* multicast_ptr, void* const* peer_ptrs, size_t offset_elems, size_t elems, int rank, int world,
                            void* stream) {
    return allreduce_sum_f16(multicast_ptr, peer_ptrs, nullptr, offset_elems, elems, rank, world, (cudaStream_t)stream);
}
int flexq_allreduce_sum_synced_f16(void* multicast_ptr, void* const* peer_ptrs, void* const* flag_ptrs, size_t offset_elems,
                                   size_t elems, int rank, int world, void* stream) {
    if (!flag_ptrs) return FLEXQ_ERR_NULL;
    return allreduce_sum_f16(multicast_ptr, peer_ptrs, flag_ptrs, offset_elems, elems, rank, world, (cudaStream_t)stream);
}

}  // extern "C"
