// oracle/_ref builder shim (TEST INFRASTRUCTURE).  Compiles the reference's fused
// activation quantise+pack kernel where it lies
// (/root/reference/e2e/src/fastertransformer/kernels/flexqgemm/src/pack/bit_packing.cu:80-221,
// include root passed as -I by oracle/Makefile) and exposes it with a C ABI.
#include "src/fastertransformer/kernels/flexqgemm/src/pack/bit_packing.cu"

extern "C" int ref_e2e_quant_pack_f16(const void* in, int* packed, void* x_scale, int M, int K, int bits, void* stream)
{
    flexq_bit_packing((const half*)in, packed, (half*)x_scale, M, K, bits, (cudaStream_t)stream);
    return (int)cudaGetLastError();
}
