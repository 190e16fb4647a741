// Source-compatible host interface of the reference engine / FasterTransformer adapter, backed
// by libflexq_b200 (C ABI in include/flexq_b200.h).  A caller of the reference headers
//   engine/src/pack/bit_packing.h:34                      flexq_bit_packing(const int*, ...)
//   e2e/.../flexqgemm/src/pack/bit_packing.h:32-34        flexq_bit_packing(const half*, ...)
//   engine/src/bgemm/flexq_bmma_op.h:19-34,187-188        FQBMMAOpState, FQBMMAInitFn_t, FQBMMAExecFn_t
//   e2e/.../flexqgemm/flexq_gemm_wrapper.h:6-48           class FLEXQGEMMWrapper
// compiles against this header unchanged.  Tile-shape template arguments of the reference's 325
// instantiated configurations are advisory here: every (x_bits, 6, signed) configuration maps to
// the one sm_100a kernel family, selected by M at run time.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../include/flexq_b200.h"

// ---- packers -------------------------------------------------------------------------------
cudaError_t flexq_bit_packing(const int* in_data, int* packed_data, const int M, const int K, const int BIT, cudaStream_t stream);
void flexq_bit_packing(const half* in_data, int* packed_data, half* T_out_scale, const int M, const int K, const int BIT, cudaStream_t stream);

// ---- GEMM op state / function-pointer pairs --------------------------------------------------
struct FQBMMAOpState {
    size_t shared_mem_size;
    dim3 gridDim;
    dim3 blockDim;
    bool initSuccess = false;
    struct Argument_t {
        int M, N, K;
        int* X;
        int* W;
        half* X_SCALE;
        half* W_SCALE;
        half* D;
        int group_size;
        bool bias = false;
    } args;
    int x_bits = 6;          // extension: which activation width the planes in X hold
};
typedef FQBMMAOpState (*FQBMMAInitFn_t)(int*, int*, half*, half*, int, int, int, half*, int, bool);
typedef void (*FQBMMAExecFn_t)(FQBMMAOpState&, cudaStream_t);

// One pair per activation width.  Init converts the weight planes to W6 tiles once per weight
// pointer (cached, freed by flexq_compat_release()); Exec is asynchronous and takes the workspace of its stream
// (one per stream, grown to the problem: no token limit, Execs on different streams never share scratch).
FQBMMAOpState FQBMMA_W6A6_InitFn(int* X, int* W, half* X_SCALE, half* W_SCALE, int M, int N, int K, half* D, int group_size, bool bias);
FQBMMAOpState FQBMMA_W6A8_InitFn(int* X, int* W, half* X_SCALE, half* W_SCALE, int M, int N, int K, half* D, int group_size, bool bias);
void FQBMMA_ExecFn(FQBMMAOpState& state, cudaStream_t stream);
// The converted tiles are cached by weight pointer (+ N, K).  A caller that repacks a weight in place, or frees it so
// that the allocator may hand the address to another weight of the same shape, calls flexq_compat_invalidate(W) first;
// flexq_compat_release() frees everything (converted tiles and the per-stream workspaces).
void flexq_compat_invalidate(const int* W);
void flexq_compat_release();

// legacy symbol names (common/base.h:286-308 naming) resolve to the pairs above
#define FQ_COMPAT_ALIAS(XB, BM, BN, BK, WM, WN, WK, NSTAGE, STRIDE)                                                       \
    static const FQBMMAInitFn_t FQBMMA_##XB##x6xtrue_##BM##x##BN##x##BK##_##WM##x##WN##x##WK##_8x8x128_##NSTAGE##_##STRIDE##_InitFn = \
        (XB == 8) ? FQBMMA_W6A8_InitFn : FQBMMA_W6A6_InitFn;                                                              \
    static const FQBMMAExecFn_t FQBMMA_##XB##x6xtrue_##BM##x##BN##x##BK##_##WM##x##WN##x##WK##_8x8x128_##NSTAGE##_##STRIDE##_ExecFn = FQBMMA_ExecFn;
// the eight configurations FLEXQGEMMWrapper dispatches to (flexq_gemm_wrapper.cu:53-84)
FQ_COMPAT_ALIAS(6, 1, 32, 256, 8, 48, 128, 2, 1)
FQ_COMPAT_ALIAS(6, 2, 32, 512, 16, 48, 128, 2, 1)
FQ_COMPAT_ALIAS(6, 4, 32, 512, 24, 48, 128, 2, 1)
FQ_COMPAT_ALIAS(6, 8, 16, 256, 48, 48, 128, 4, 1)
FQ_COMPAT_ALIAS(8, 1, 32, 256, 8, 48, 128, 4, 1)
FQ_COMPAT_ALIAS(8, 2, 32, 256, 16, 48, 128, 4, 1)
FQ_COMPAT_ALIAS(8, 4, 64, 256, 32, 48, 128, 4, 1)
FQ_COMPAT_ALIAS(8, 8, 64, 384, 64, 48, 128, 2, 1)

// ---- FasterTransformer adapter -----------------------------------------------------------------
// B holds W6 tiles (same byte size as the reference's packed planes for N % 128 == 0); the workspace
// must be flexq_linear_workspace_bytes(M, K) bytes, zeroed once with flexq_workspace_init.
class FLEXQGEMMWrapper {
private:
    int x_bits_;
    int w_bits_;
    bool signed_;

public:
    FLEXQGEMMWrapper(int X_BITS, int W_BITS, bool SIGNED);
    ~FLEXQGEMMWrapper();
    void pack(const half* in_data, int* packed_data, half* x_scale, int M, int K, int BIT, cudaStream_t stream);
    void gemm(const int M, const int N, const int K, const int* A, const int* B, const half* C, half* D, float* x_scale,
              const float* w_scale, const float* scale_inter, const float* scale_out, bool bias, char* flexq_gemm_workspace,
              size_t flexq_gemm_ws_bytes, cudaStream_t stream);
    void gemm(const int M, const int N, const int K, const half* A, const int* B, const half* C, half* D, float* x_scale,
              const float* w_scale, const float* scale_inter, const float* scale_out, bool bias, char* flexq_gemm_workspace,
              size_t flexq_gemm_ws_bytes, cudaStream_t stream);
};
