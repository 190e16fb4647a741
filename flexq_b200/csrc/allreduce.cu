// Sum all-reduce of the fp16 partial outputs of a row-parallel W6Ax linear over NVLink 5 / NVSwitch peer
// memory (replaces ftNcclAllReduceSum after the down / o_proj GEMMs,
// /root/reference/e2e/src/fastertransformer/layers/TensorParallelSiluFfnLayer.cc:53-56 and
// utils/nccl_utils.cc:56-68; the reference's own peer-memory variant is kernels/custom_ar_kernels.cu:139-260).
//
// The buffer lives in symmetric memory: the same allocation on every rank, mapped into every rank's address
// space (peer pointers) and, on NVSwitch, behind one multicast address.  Two-shot, in place: rank r owns the
// r-th 1/world slice of the elements,
//   * multicast path: one multimem.ld_reduce pulls the slice already summed inside the switch (fp32
//     accumulation), one multimem.st broadcasts the result to every rank -- each byte crosses this GPU's links
//     once in each direction;
//   * peer path (no multicast): loads the slice from every rank's copy, sums in fp32 in rank order (so every
//     rank would compute identical bits), stores to every rank's copy.
// Cross-rank ordering (all partials written before, all slices stored after) is the caller's: a symmetric-memory
// barrier on the same stream on both sides (flexq_b200/tp.py).
#include "common.cuh"

namespace flexq {

struct PeerPtrs {
    __half* p[8];
};

__device__ __forceinline__ uint4 multimem_ld_reduce_f16x8(const void* mc) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st_f16x8(void* mc, const uint4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f16x2 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

constexpr int kArUnroll = 4;
static int g_ar_blocks = 4 * 148;           // grid cap (set small to stay on SMs the GEMM leaves free)
void set_allreduce_blocks(int n) { g_ar_blocks = n > 0 ? n : 4 * 148; }

// vec0 .. vec1: this rank's range of 16-byte vectors
__global__ void __launch_bounds__(256) allreduce_multimem_kernel(__half* mc, long long vec0, long long vec1) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = vec0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    uint4* base = reinterpret_cast<uint4*>(mc);
    for (; i + (kArUnroll - 1) * stride < vec1; i += kArUnroll * stride) {
        uint4 v[kArUnroll];
#pragma unroll
        for (int u = 0; u < kArUnroll; u++) v[u] = multimem_ld_reduce_f16x8(base + i + u * stride);
#pragma unroll
        for (int u = 0; u < kArUnroll; u++) multimem_st_f16x8(base + i + u * stride, v[u]);
    }
    for (; i < vec1; i += stride) multimem_st_f16x8(base + i, multimem_ld_reduce_f16x8(base + i));
}

template <int WORLD>
__global__ void __launch_bounds__(256) allreduce_peer_kernel(PeerPtrs peers, long long vec0, long long vec1) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = vec0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < vec1; i += stride) {
        uint4 in[WORLD];
#pragma unroll
        for (int r = 0; r < WORLD; r++) in[r] = __ldcv(reinterpret_cast<const uint4*>(peers.p[r]) + i);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = 0.f;
#pragma unroll
        for (int r = 0; r < WORLD; r++) {
            const __half2* h = reinterpret_cast<const __half2*>(&in[r]);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float2 f = __half22float2(h[j]);
                acc[2 * j] += f.x;
                acc[2 * j + 1] += f.y;
            }
        }
        uint4 out;
        __half2* o = reinterpret_cast<__half2*>(&out);
#pragma unroll
        for (int j = 0; j < 4; j++) o[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
#pragma unroll
        for (int r = 0; r < WORLD; r++) *(reinterpret_cast<uint4*>(peers.p[r]) + i) = out;
    }
}

int allreduce_sum_f16(void* multicast_ptr, void* const* peer_ptrs, size_t offset_elems, size_t elems, int rank, int world,
                      cudaStream_t stream) {
    if (world < 1 || world > 8 || rank < 0 || rank >= world) return FLEXQ_ERR_BAD_SHAPE;
    if (!multicast_ptr && !peer_ptrs) return FLEXQ_ERR_NULL;
    if (elems % 8 || offset_elems % 8) return FLEXQ_ERR_BAD_SHAPE;             // 16-byte vectors
    if (elems == 0 || world == 1) return 0;
    const long long nvec = (long long)(elems / 8), v_off = (long long)(offset_elems / 8);
    const long long per = (nvec + world - 1) / world;
    const long long lo = per * rank, hi = per * (rank + 1);
    const long long vec0 = v_off + (lo < nvec ? lo : nvec), vec1 = v_off + (hi < nvec ? hi : nvec);
    if (vec1 <= vec0) return 0;
    const long long work = vec1 - vec0;
    long long nb = (work + 256 * kArUnroll - 1) / (256 * kArUnroll);
    const int blocks = (int)(nb > g_ar_blocks ? g_ar_blocks : (nb < 1 ? 1 : nb));
    if (multicast_ptr) {
        allreduce_multimem_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<__half*>(multicast_ptr), vec0, vec1);
        return (int)cudaGetLastError();
    }
    PeerPtrs pp{};
    for (int r = 0; r < world; r++) {
        if (!peer_ptrs[r]) return FLEXQ_ERR_NULL;
        pp.p[r] = reinterpret_cast<__half*>(peer_ptrs[r]);
    }
    switch (world) {
        case 2: allreduce_peer_kernel<2><<<blocks, 256, 0, stream>>>(pp, vec0, vec1); break;
        case 4: allreduce_peer_kernel<4><<<blocks, 256, 0, stream>>>(pp, vec0, vec1); break;
        case 8: allreduce_peer_kernel<8><<<blocks, 256, 0, stream>>>(pp, vec0, vec1); break;
        default: return FLEXQ_ERR_BAD_SHAPE;
    }
    return (int)cudaGetLastError();
}

}  // namespace flexq
