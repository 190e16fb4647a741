/*
 * flexq_b200 -- C ABI of the B200-native (sm_100a) W6Ax quantized-linear hot path.
 *
 * This header is the drop-in boundary.  Every entry point takes plain device pointers,
 * sizes and an opaque CUDA stream (a cudaStream_t passed as void*), returns an int status
 * (0 = OK, <0 = FLEXQ_ERR_*, >0 = cudaError_t of the failing CUDA call) and never
 * allocates, synchronises or falls back to the CPU.  Each declaration cites the
 * reference (hoffmann-muki/FlexQ, paths under /root/reference) interface it replaces.
 *
 * Data formats (details: DESIGN.md "Data layout in HBM")
 *   Xq      int8  [M][K]            quantised activations in int8 containers (A6 or A8)
 *   sx      f32   [K/128][ldsx]     activation scales, ldsx = flexq_sx_ld(M); value is the
 *                                   fp16-rounded scale the quantiser divided by
 *   W6      u8    [ceil(N/128)][K/128][12288]  6-bit weights, 128x128 tiles, TMA-bulk friendly
 *   w_scale f16   [K/128][N]        same as the reference's W_SCALE (test_bgemm_kernel.cu:57-63)
 *   D       f16   [M][N]            row major, as the reference
 *   planes  u32   [K/128][R/chunk][bits][chunk][4]   reference bit-plane layout
 *                                   (engine/src/pack/bit_packing.cu:42-99), chunk = min(R,8)
 *   X_SCALE f16   [K/128][2*ceil4(M)]  reference activation-scale layout, entries duplicated
 *                                   in pairs (engine/test_bgemm_kernel.cu:41-54)
 */
#ifndef FLEXQ_B200_H_
#define FLEXQ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLEXQ_OK                 0
#define FLEXQ_ERR_BAD_SHAPE     -1   /* K % 128 != 0, K < 128, M/N <= 0, plane layout needs R % min(R,8) == 0 */
#define FLEXQ_ERR_BAD_BITS      -2   /* x_bits not in {6,8} (weights are always 6 bit)   */
#define FLEXQ_ERR_NULL          -3
#define FLEXQ_ERR_WORKSPACE     -4   /* workspace too small / misaligned                  */
#define FLEXQ_ERR_NO_DEVICE     -5   /* no sm_100 device, or driver lacks cuTensorMapEncodeTiled */
#define FLEXQ_ERR_TENSORMAP     -6

#define FLEXQ_GROUP              128 /* engine/src/bgemm/flexq_bmma_kernel.h:54 */

/* activation rounding behaviour -- the reference has two */
#define FLEXQ_ROUND_CUDA         0   /* e2e/.../flexqgemm/src/pack/bit_packing.cu:150-163: fp16-rounded
                                        scale, no min clamp, round half away from zero             */
#define FLEXQ_ROUND_PYTHON       1   /* algorithm/flexq_quantize/quantizer.py:112-116,153-155: scale
                                        clamped to [1e-5,1e4], round half to even, fp16 arithmetic  */

int         flexq_version(void);
const char* flexq_status_string(int status);

/* ---- sizes ---------------------------------------------------------------------------- */
size_t flexq_w6_packed_bytes(int N, int K);            /* 12288 * ceil(N/128) * K/128           */
size_t flexq_planes_bytes(int R, int K, int bits);     /* R*K*bits/8, as the reference allocates  */
int    flexq_sx_ld(int M);                             /* leading dim of sx: ceil4(M)             */
size_t flexq_xscale_ref_halves(int M, int K);          /* K/128 * 2*ceil4(M)                      */
size_t flexq_gemm_workspace_bytes(void);               /* split-K scratch, shape independent      */
size_t flexq_linear_workspace_bytes(int M, int K);     /* gemm workspace + Xq + sx                */

/* Zero a workspace once after allocation (the GEMM leaves its cut-tile records zeroed again after every call).
 * One workspace must not be shared by GEMMs running concurrently on different streams. */
int flexq_workspace_init(void* workspace, size_t bytes, void* stream);

/* ---- reference-layout packers (API parity) ------------------------------------------------
 * replaces: cudaError_t flexq_bit_packing(const int*, int*, int M, int K, int BIT, cudaStream_t)
 *           engine/src/pack/bit_packing.h:34 (kernel bit_packing.cu:42-99, launch :113-122)     */
int flexq_bit_packing_i32(const int32_t* in, int32_t* planes, int R, int K, int bits, void* stream);

/* replaces: void flexq_bit_packing(const half*, int*, half* T_out_scale, int M, int K, int BIT,
 *           cudaStream_t)  e2e/src/fastertransformer/kernels/flexqgemm/src/pack/bit_packing.h:34
 *           (kernel bit_packing.cu:80-199): per-(row,group) absmax -> scale -> quantise -> planes,
 *           scales written twice as half at [g][2m], [g][2m+1].                                 */
int flexq_bit_packing_f16(const void* x_half, int32_t* planes, void* x_scale_half,
                          int M, int K, int bits, void* stream);

/* ---- native activation path (north-star subsystem 2) -------------------------------------
 * Fused per-token per-group absmax -> 6/8-bit quantise -> int8 containers + fp32 scales.
 * Same arithmetic as flexq_bit_packing_f16 (mode FLEXQ_ROUND_CUDA) or as
 * UniformAffineQuantizer (mode FLEXQ_ROUND_PYTHON).                                            */
int flexq_quant_act(const void* x_half, int8_t* xq, float* sx, int M, int K, int bits, int mode,
                    void* stream);

/* fp32 activations: the same quantiser evaluated in fp32 arithmetic, as UniformAffineQuantizer does for float
 * tensors (quantizer.py:144-155 scale = absmax/qmax clamped to [1e-5,1e4]; :112-116 round half to even, clamp) --
 * the reference's CPU-runnable path (QuantLinear on an fp32 module, int_linear.py:56-72).  sx receives the fp32
 * scale unrounded.                                                                                              */
int flexq_quant_act_f32(const float* x, int8_t* xq, float* sx, int M, int K, int bits, void* stream);

/* ---- offline weight packer (north-star subsystem 1) --------------------------------------
 * ints [N][K] (two's complement in 6 bits, i.e. [-32,31]; int32 or int8 input) -> W6 tiles.
 * Caller of the int32 form today: engine/test_bgemm_kernel.cu:222-224 (flexq_bit_packing on W). */
int flexq_pack_w6_i32(const int32_t* w_int, uint8_t* w6, int N, int K, void* stream);
int flexq_pack_w6_i8(const int8_t* w_int, uint8_t* w6, int N, int K, void* stream);

/* fp16 / fp32 weights [N][K] -> per-group symmetric 6-bit quantisation exactly as
 * UniformAffineQuantizer does for weights (quantizer.py:144-171 + 93-126, arithmetic in the
 * input dtype) -> W6 tiles + w_scale f16 [K/128][N].  This is the exporter the reference lacks
 * (flexq_quantize/utils.py:116-123 only overwrites W with its fake-quantised value).            */
int flexq_quant_pack_w6_f16(const void* w_half, uint8_t* w6, void* w_scale_half, int N, int K, void* stream);
int flexq_quant_pack_w6_f32(const float* w, uint8_t* w6, void* w_scale_half, int N, int K, void* stream);

/* ---- converters from the reference layouts -------------------------------------------------- */
int flexq_planes_to_i8(const int32_t* planes, int8_t* out, int R, int K, int bits, void* stream);
int flexq_planes_to_w6(const int32_t* w_planes, uint8_t* w6, int8_t* scratch_NK, int N, int K, void* stream);
int flexq_xscale_ref_to_sx(const void* x_scale_half, float* sx, int M, int K, void* stream);
int flexq_w6_to_i8(const uint8_t* w6, int8_t* out, int N, int K, void* stream);   /* unpack (debug/tests) */

/* ---- W6A6 / W6A8 GEMM (north-star subsystem 3) -------------------------------------------
 * D[m][n] = half( sum_g sx[g][m] * float(w_scale[g][n]) * S[m][n][g] ),
 * S[m][n][g] = sum_{k in group g} Xq[m][k] * W[n][k]   (INT32, exact; tcgen05.mma kind::i8).
 * replaces: FQBMMAInitFn_t / FQBMMAExecFn_t pairs  engine/src/bgemm/flexq_bmma_op.h:163-188
 *           (kernel flexq_bmma_kernel.h:119-447) and FLEXQGEMMWrapper::gemm(int* A ...)
 *           e2e/.../flexqgemm/flexq_gemm_wrapper.cu:21-97.                                       */
int flexq_gemm_w6ax(const int8_t* xq, const float* sx, const uint8_t* w6, const void* w_scale_half,
                    void* d_half, int M, int N, int K, void* workspace, size_t workspace_bytes,
                    void* stream);

/* SURVEY 8(f3), first half -- the gate_up GEMM of the MLP with SiLU(gate) * up applied by the GEMM's epilogue, so the
 * [M][2*inter] fp16 intermediate makes no HBM round trip:
 *   H[m][j] = half( silu(half(gate[m][j])) * half(up[m][j]) ),  gate / up = the W6Ax GEMM outputs of rows j of gate_proj / up_proj
 * (fp32 SiLU with __expf and fast division on the fp16-rounded GEMM outputs: the arithmetic of flexq_silu_mul_quant_f16).
 * w6_gate_up / w_scale_gate_up: the 2*inter rows packed as usual, ordered 8 gate rows, their 8 up rows, 8 gate rows, ...
 * (flexq_b200.model_pack.interleave_gate_up); inter % 8 == 0.  H is [M][inter] fp16; quantise it with flexq_quant_act.
 * replaces: the gate / up GEMMs + the activation kernel of  e2e/src/fastertransformer/layers/FfnLayer.cc:371-401,
 *           kernels/activation_kernels.cu:245-440 (their A8 quantisation stays a separate pass here).             */
int flexq_gemm_w6ax_silu_mul(const int8_t* xq, const float* sx, const uint8_t* w6_gate_up, const void* w_scale_gate_up_half,
                             void* h_half, int M, int inter, int K, void* workspace, size_t workspace_bytes, void* stream);

/* Debug/parity entry: the INT32 per-K-group partial sums S[M][N][K/128] of the same kernel. */
int flexq_gemm_w6ax_groupsums(const int8_t* xq, const uint8_t* w6, int32_t* S, int M, int N, int K,
                              void* stream);

/* Debug / tests (host only, no device work): the work decomposition of the GEMM.  For a problem of m_tiles token
 * tiles x n_tiles weight tiles x `groups` k-groups on at most max_ctas CTAs, writes the segments CTA `cta` walks as
 * int[5] = {token tile, n-tile, first group, end group, fp32 slot or -1 when the CTA covers the whole tile} (up to
 * `cap` of them), stores the number of CTAs launched in *n_ctas and returns the CTA's segment count.        */
int flexq_debug_schedule(int m_tiles, int n_tiles, int groups, int max_ctas, int cta, int* segments, int cap, int* n_ctas);

/* Debug only: the GEMM with clock64 stamps of one CTA's pipeline events, trace[unit][16],
 * plus per-CTA wall-clock windows (token tile by M as in flexq_gemm_w6ax: 16 / 32 / 64 / 128, above that always
 * the 192-token tile); trace_units = steps to stamp | traced CTA << 16; used by tools/trace.py. */
int flexq_debug_gemm_trace(const int8_t* xq, const float* sx, const uint8_t* w6, const void* w_scale_half,
                           void* d_half, int M, int N, int K, void* workspace, long long* trace,
                           int trace_units, void* stream);

/* Fused linear: fp16 activations in, fp16 out (activation quantise + GEMM on one stream).
 * replaces: FLEXQGEMMWrapper::gemm(half* A ...) flexq_gemm_wrapper.cu:99-122 and is what
 * QuantLinear.forward (algorithm/flexq_quantize/int_linear.py:56-72) maps to.                   */
int flexq_linear_w6ax_f16(const void* x_half, const uint8_t* w6, const void* w_scale_half, void* d_half,
                          int M, int N, int K, int x_bits, int mode, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Engine-level drop-in on the reference's own operand layouts: X as bit planes + duplicated
 * half scales (what flexq_bit_packing / the FT layers produce), W already converted once with
 * flexq_planes_to_w6.  Workspace as flexq_linear_workspace_bytes(M, K).                          */
int flexq_gemm_ref_layout(const int32_t* x_planes, const void* x_scale_half, const uint8_t* w6,
                          const void* w_scale_half, void* d_half, int M, int N, int K, int x_bits,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Producer-side fusions: the kernel that produces a GEMM's activations also quantises
 * them (native int8 containers + fp32 scales, a6 rounding), feeding flexq_gemm_w6ax directly.
 *
 * flexq_rmsnorm_quant_f16 replaces invokeGeneralT5LayerNorm / invokeGeneralAddBiasResidualT5PreLayerNorm with
 * norm_output_scale != NULL (e2e/src/fastertransformer/kernels/layernorm_kernels.cu:2494-2690, 1852-2110):
 * if `residual` != NULL it is updated in place to half(x + residual) and normalised instead of x;
 * y = half((h * rsqrt(mean(h^2) + eps)) * gamma); `normed` (optional, may be NULL) receives y.  K <= 16384.
 *
 * flexq_silu_mul_quant_f16 replaces invokeGenericActivation<SiluActivation> with output_scale != NULL
 * (kernels/activation_kernels.cu:246-440, called from layers/FfnLayer.cc:442-452): y = half(silu(gate) * up),
 * gate/up rows `ld_in` halves apart (2*K for the halves of a fused gate_up output); `out` optional.       */
int flexq_rmsnorm_quant_f16(const void* x_half, void* residual_half_inout, const void* gamma_half, float eps,
                            void* normed_half_out, int8_t* xq, float* sx, int M, int K, int bits, void* stream);
int flexq_silu_mul_quant_f16(const void* gate_half, const void* up_half, long long ld_in, void* out_half,
                             int8_t* xq, float* sx, int M, int K, int bits, void* stream);

/* Tensor-parallel reduction of row-parallel partial outputs (replaces ftNcclAllReduceSum,
 * e2e/src/fastertransformer/utils/nccl_utils.cc:56-68, called after the down / o_proj GEMMs in
 * layers/TensorParallelSiluFfnLayer.cc:53-56; peer-memory precedent kernels/custom_ar_kernels.cu:139-260).
 * In-place sum over `world` ranks of `elems` halves starting `offset_elems` into a symmetric allocation:
 * `multicast_ptr` = the NVSwitch multicast mapping of the allocation (in-switch reduction, multimem PTX),
 * or NULL to use `peer_ptrs[0..world)` = every rank's mapping of it (host array of device pointers).
 * elems and offset_elems must be multiples of 8.  This rank reduces and publishes its 1/world slice;
 * the caller orders ranks with a symmetric-memory barrier on `stream` before and after the call.       */
/* Same reduction with the cross-rank ordering inside the kernel: `flag_ptrs[r]` = rank r's mapping of a
 * zero-initialised symmetric array of 32 uint32; the kernel first waits until every rank has entered it (all
 * partials written) and returns only after every rank has published its slice -- no barriers on the stream,
 * CUDA-graph capturable.  Bounded spins: a missing rank makes the others trap instead of hanging.        */
int flexq_allreduce_sum_synced_f16(void* multicast_ptr, void* const* peer_ptrs, void* const* flag_ptrs,
                                   size_t offset_elems, size_t elems, int rank, int world, void* stream);

/* One-kernel variant for decode-sized reductions (no stream barriers, CUDA-graph capturable;
 * oneShotAllReduceKernel of kernels/custom_ar_kernels.cu:139-190 is the reference's counterpart).
 * `data_ptrs[r]` / `flag_ptrs[r]` = rank r's mapping of a symmetric buffer holding that rank's partial (elems
 * halves) and of a zero-initialised symmetric array of 32 uint32 (arrive/done slots, epoch, ticket).
 * `out` (private, elems halves) receives the sum; the call returns after every rank has finished reading,
 * so the partial may be overwritten by the next kernel of the stream.  A rank that never arrives makes the
 * others trap after a bounded spin instead of hanging.  elems % 8 == 0, at most 64 CTAs.                */
int flexq_allreduce_oneshot_f16(void* const* data_ptrs, void* const* flag_ptrs, size_t elems, int rank, int world,
                                void* out_half, void* stream);

/* Overlap knobs (process-wide): cap the persistent GEMM at n_ctas CTAs (0 = every SM) so that a kernel on
 * another stream finds free SMs, and cap the all-reduce grid at n_blocks (0 = default).                  */
int flexq_set_sm_limit(int n_ctas);
int flexq_set_allreduce_blocks(int n_blocks);
int flexq_allreduce_sum_f16(void* multicast_ptr, void* const* peer_ptrs, size_t offset_elems, size_t elems,
                            int rank, int world, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLEXQ_B200_H_ */
