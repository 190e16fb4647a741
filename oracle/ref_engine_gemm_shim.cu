// oracle/_ref builder shim (TEST INFRASTRUCTURE).  Compiles the reference's own engine
// sources *where they lie* under /root/reference/engine (passed as -I by oracle/Makefile;
// nothing is copied into this repo) and gives them a C ABI so that tests on the GPU box
// can call the reference packer and the reference BMMA GEMM as a GPU-side oracle.
//   reference packer : engine/src/pack/bit_packing.cu:42-122
//   reference GEMM   : engine/src/bgemm/flexq_bmma_op.h:163-188, flexq_bmma_kernel.h:119-447
//   M-bucket -> tile config table: e2e/.../flexqgemm/flexq_gemm_wrapper.cu:53-84
#include "src/bgemm/flexq_bmma_op.h"

// the eight tile configs the reference's own FT wrapper dispatches to
FQ_INSTANTIATE_FUN(FQBMMA, 6, 6, true, 1, 32, 256, 8, 48, 128, 8, 8, 128, 2, 1);
FQ_INSTANTIATE_FUN(FQBMMA, 6, 6, true, 2, 32, 512, 16, 48, 128, 8, 8, 128, 2, 1);
FQ_INSTANTIATE_FUN(FQBMMA, 6, 6, true, 4, 32, 512, 24, 48, 128, 8, 8, 128, 2, 1);
FQ_INSTANTIATE_FUN(FQBMMA, 6, 6, true, 8, 16, 256, 48, 48, 128, 8, 8, 128, 4, 1);
FQ_INSTANTIATE_FUN(FQBMMA, 8, 6, true, 1, 32, 256, 8, 48, 128, 8, 8, 128, 4, 1);
FQ_INSTANTIATE_FUN(FQBMMA, 8, 6, true, 2, 32, 256, 16, 48, 128, 8, 8, 128, 4, 1);
FQ_INSTANTIATE_FUN(FQBMMA, 8, 6, true, 4, 64, 256, 32, 48, 128, 8, 8, 128, 4, 1);
FQ_INSTANTIATE_FUN(FQBMMA, 8, 6, true, 8, 64, 384, 64, 48, 128, 8, 8, 128, 2, 1);

extern "C" {

// Same dispatch as FLEXQGEMMWrapper::gemm (flexq_gemm_wrapper.cu:53-96).  Returns 0 on
// success, -1 unsupported bits / K, -2 init failure, else the CUDA error code.
int ref_fqbmma_gemm(int* X, int* W, void* X_SCALE, void* W_SCALE, int M, int N, int K, void* D,
                    int x_bits, void* stream)
{
    FQBMMAInitFn_t init_fn; FQBMMAExecFn_t exec_fn;
    if (K < 128 || K % 128 != 0) return -1;
    if (x_bits == 6) {
        if (M == 1)      { init_fn = FQBMMA_6x6xtrue_1x32x256_8x48x128_8x8x128_2_1_InitFn;  exec_fn = FQBMMA_6x6xtrue_1x32x256_8x48x128_8x8x128_2_1_ExecFn; }
        else if (M == 2) { init_fn = FQBMMA_6x6xtrue_2x32x512_16x48x128_8x8x128_2_1_InitFn; exec_fn = FQBMMA_6x6xtrue_2x32x512_16x48x128_8x8x128_2_1_ExecFn; }
        else if (M == 4) { init_fn = FQBMMA_6x6xtrue_4x32x512_24x48x128_8x8x128_2_1_InitFn; exec_fn = FQBMMA_6x6xtrue_4x32x512_24x48x128_8x8x128_2_1_ExecFn; }
        else             { init_fn = FQBMMA_6x6xtrue_8x16x256_48x48x128_8x8x128_4_1_InitFn; exec_fn = FQBMMA_6x6xtrue_8x16x256_48x48x128_8x8x128_4_1_ExecFn; }
    } else if (x_bits == 8) {
        if (M == 1)      { init_fn = FQBMMA_8x6xtrue_1x32x256_8x48x128_8x8x128_4_1_InitFn;  exec_fn = FQBMMA_8x6xtrue_1x32x256_8x48x128_8x8x128_4_1_ExecFn; }
        else if (M == 2) { init_fn = FQBMMA_8x6xtrue_2x32x256_16x48x128_8x8x128_4_1_InitFn; exec_fn = FQBMMA_8x6xtrue_2x32x256_16x48x128_8x8x128_4_1_ExecFn; }
        else if (M == 4) { init_fn = FQBMMA_8x6xtrue_4x64x256_32x48x128_8x8x128_4_1_InitFn; exec_fn = FQBMMA_8x6xtrue_4x64x256_32x48x128_8x8x128_4_1_ExecFn; }
        else             { init_fn = FQBMMA_8x6xtrue_8x64x384_64x48x128_8x8x128_2_1_InitFn; exec_fn = FQBMMA_8x6xtrue_8x64x384_64x48x128_8x8x128_2_1_ExecFn; }
    } else return -1;
    FQBMMAOpState st = (*init_fn)(X, W, (half*)X_SCALE, (half*)W_SCALE, M, N, K, (half*)D, 128, false);
    if (!st.initSuccess) return -2;
    (*exec_fn)(st, (cudaStream_t)stream);
    return (int)cudaGetLastError();
}

}  // extern "C"
