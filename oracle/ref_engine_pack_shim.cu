// oracle/_ref builder shim (TEST INFRASTRUCTURE).  Compiles the reference's engine packer
// where it lies (/root/reference/engine/src/pack/bit_packing.cu:20-122; include root passed
// as -I by oracle/Makefile; nothing is copied) and exposes it with a C ABI.  Kept in its own
// translation unit because the packer's BLOCK_M/BLOCK_K macros collide with the GEMM headers.
#include "src/pack/bit_packing.cu"

extern "C" {

int ref_flexq_bit_packing_i32(const int* in, int* packed, int M, int K, int bits, void* stream)
{
    return (int)flexq_bit_packing(in, packed, M, K, bits, (cudaStream_t)stream);
}

int ref_abq_bit_packing_i32(const int* in, int* packed, int M, int K, int bits, void* stream)
{
    abq_bit_packing(in, packed, M, K, bits, (cudaStream_t)stream);
    return (int)cudaGetLastError();
}

}  // extern "C"
