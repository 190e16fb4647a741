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

// ---- cross-rank hand-shake through flag words in symmetric memory (used by the synced two-shot kernels and the
// one-shot kernel).  Layout per rank (uint32): [0, 8) arrive[src], [8, 16) done[src], [16] epoch of the last finished
// call, [17] CTA ticket.  The epoch is a device-side counter, so calls can sit in a CUDA graph.
struct FlagPtrs {
    uint32_t* p[8];
};
constexpr int kFlagDone = 8, kFlagEpoch = 16, kFlagTicket = 17;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Bounded by wall clock, not by polls: a peer stalled on its host (GC, logging, checkpoint I/O) for seconds is ordinary
// rank skew and must not kill the context; only a rank that stays away for kArTimeoutNs (2 minutes) is taken as missing
// and the kernel fails instead of hanging the GPU.
constexpr unsigned long long kArTimeoutNs = 120ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ void spin_until(const uint32_t* f, uint32_t epoch) {
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
        if ((++spins & 0xFFFFu) == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > kArTimeoutNs) __trap();
        }
    }
}
// Opening: tell every peer that this rank's data is in place, wait until every peer said the same.  Returns the epoch.
__device__ __forceinline__ uint32_t ar_open(const FlagPtrs& flags, int rank, int world) {
    uint32_t* mine = flags.p[rank];
    const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(mine + kFlagEpoch) + 1u;   // bumped by the last CTA, after every CTA read it
    if (blockIdx.x == 0 && (int)threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(flags.p[threadIdx.x] + rank, epoch);
    }
    if ((int)threadIdx.x < world) spin_until(mine + threadIdx.x, epoch);
    __syncthreads();
    return epoch;
}
// Closing: the last CTA of this rank tells the peers that this rank is done (reading / publishing) and waits for all.
__device__ __forceinline__ void ar_close(const FlagPtrs& flags, int rank, int world, uint32_t epoch) {
    uint32_t* mine = flags.p[rank];
    __threadfence_system();
    __syncthreads();
    __shared__ uint32_t s_last;
    if (threadIdx.x == 0) s_last = (atomicAdd(mine + kFlagTicket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        __threadfence_system();        // acquire side of the ticket: the other CTAs' stores are ordered before the done flags below
        if ((int)threadIdx.x < world) {
            st_release_sys(flags.p[threadIdx.x] + kFlagDone + rank, epoch);
            spin_until(mine + kFlagDone + threadIdx.x, epoch);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mine[kFlagTicket] = 0u;
            *reinterpret_cast<volatile uint32_t*>(mine + kFlagEpoch) = epoch;
        }
    }
}

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

// 16-byte vectors a thread of the multicast kernel keeps in flight (all loads of an iteration are issued before the first
// store); FLEXQ_AR_UNROLL (1, 2, 4, 8) overrides it at run time.  Measured on 8 x B200 (tools/ar_probe.py,
// profiles/ar_probe_tp8_r2.txt): the time does not depend on it, nor on the number of blocks beyond 8 -- 22 us fixed
// (two symmetric-memory barriers) + 2.2 us per MB, i.e. the in-switch reduction runs at ~460 GB/s whatever is in flight,
// and a free-running grid is slower (119 us for 32 MB) than 8 blocks of 1024 threads (92 us).
#ifndef FLEXQ_AR_UNROLL
#define FLEXQ_AR_UNROLL 4
#endif
static int ar_unroll() {
    static int v = 0;
    if (v == 0) {
        const char* e = getenv("FLEXQ_AR_UNROLL");
        const int n = e ? atoi(e) : FLEXQ_AR_UNROLL;
        v = (n == 1 || n == 2 || n == 4 || n == 8) ? n : FLEXQ_AR_UNROLL;
    }
    return v;
}
// 0: free-running grid (up to 4 x 148 blocks of 256 threads).  n > 0: at most n blocks of 1024 threads -- used while a
// persistent GEMM (one CTA per SM, all of its shared memory) runs on the other SMs: every SM that hosts even one small
// block of ours is lost to the GEMM, so the reduction is packed onto as few SMs as the GEMM leaves free.
static thread_local int g_ar_blocks = 0;      // per host thread (set around the launches of one layer by the thread issuing them)
void set_allreduce_blocks(int n) { g_ar_blocks = n > 0 ? n : 0; }

// vec0 .. vec1: this rank's range of 16-byte vectors
template <int UNROLL>
__global__ void __launch_bounds__(1024) allreduce_multimem_kernel(__half* mc, long long vec0, long long vec1, FlagPtrs flags, int rank,
                                                                 int world) {
    const bool synced = flags.p[0] != nullptr;
    uint32_t epoch = 0;
    if (synced) epoch = ar_open(flags, rank, world);
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = vec0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    uint4* base = reinterpret_cast<uint4*>(mc);
    for (; i + (UNROLL - 1) * stride < vec1; i += UNROLL * stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) v[u] = multimem_ld_reduce_f16x8(base + i + u * stride);
#pragma unroll
        for (int u = 0; u < UNROLL; u++) multimem_st_f16x8(base + i + u * stride, v[u]);
    }
    for (; i < vec1; i += stride) multimem_st_f16x8(base + i, multimem_ld_reduce_f16x8(base + i));
    if (synced) ar_close(flags, rank, world, epoch);
}

// U vectors per thread and iteration: all WORLD x U loads are issued before the first sum, so a thread pays one NVLink
// round trip per U vectors instead of one per vector (measured on 2 x B200, 32 MB: U = 1 -> 69 us)
template <int WORLD, int U>
__global__ void __launch_bounds__(1024) allreduce_peer_kernel(PeerPtrs peers, long long vec0, long long vec1, FlagPtrs flags, int rank) {
    const bool synced = flags.p[0] != nullptr;
    uint32_t epoch = 0;
    if (synced) epoch = ar_open(flags, rank, WORLD);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = vec0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < vec1; i += stride * U) {
        uint4 in[U][WORLD];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long long idx = i + u * stride;
            if (idx < vec1) {
#pragma unroll
                for (int r = 0; r < WORLD; r++) in[u][r] = __ldcv(reinterpret_cast<const uint4*>(peers.p[r]) + idx);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long long idx = i + u * stride;
            if (idx >= vec1) break;
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] = 0.f;
#pragma unroll
            for (int r = 0; r < WORLD; r++) {
                const __half2* h = reinterpret_cast<const __half2*>(&in[u][r]);
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
            for (int r = 0; r < WORLD; r++) *(reinterpret_cast<uint4*>(peers.p[r]) + idx) = out;
        }
    }
    if (synced) ar_close(flags, rank, WORLD, epoch);
}

// ------------------------------------------------------------------------------------------------
// One-shot variant for decode-sized reductions (M x N fp16 <= ~1 MB): ONE kernel, no stream barriers.
// Every rank keeps its partial in a symmetric allocation and owns two flag words per peer.  Arrive:
// store-release the call's epoch into my slot on every peer, then spin (acquire, bounded) until every
// peer's slot in my array has reached it.  Reduce: read all ranks' partials (peer loads, fp32 sum in rank
// order -> every rank gets identical bits) into a private output.  Done: the same exchange once more, so
// that no rank overwrites its partial while a peer still reads it.  The epoch is a device-side counter, so
// the call is CUDA-graph capturable.  (The reference's FasterTransformer fork has the same idea in
// kernels/custom_ar_kernels.cu:139-260, oneShotAllReduceKernel.)
// ------------------------------------------------------------------------------------------------
template <int WORLD>
__global__ void __launch_bounds__(256) allreduce_oneshot_kernel(PeerPtrs data, FlagPtrs flags, long long nvec, int rank,
                                                                uint4* __restrict__ out) {
    // the partial was written by earlier work of this stream (PDL: wait for it before publishing)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t epoch = ar_open(flags, rank, WORLD);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        uint4 in[WORLD];
#pragma unroll
        for (int r = 0; r < WORLD; r++) in[r] = __ldcv(reinterpret_cast<const uint4*>(data.p[r]) + i);
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
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int j = 0; j < 4; j++) oh[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
        out[i] = o;
    }
    // closing phase: nobody may overwrite its partial (the next GEMM of this stream) while a peer still reads it
    ar_close(flags, rank, WORLD, epoch);
}

int allreduce_oneshot_f16(void* const* data_ptrs, void* const* flag_ptrs, size_t elems, int rank, int world, void* out,
                          cudaStream_t stream) {
    if (!data_ptrs || !flag_ptrs || !out) return FLEXQ_ERR_NULL;
    if ((world != 2 && world != 4 && world != 8) || rank < 0 || rank >= world) return FLEXQ_ERR_BAD_SHAPE;
    if (elems == 0 || elems % 8) return FLEXQ_ERR_BAD_SHAPE;
    PeerPtrs dp{};
    FlagPtrs fp{};
    for (int r = 0; r < world; r++) {
        if (!data_ptrs[r] || !flag_ptrs[r]) return FLEXQ_ERR_NULL;
        dp.p[r] = reinterpret_cast<__half*>(data_ptrs[r]);
        fp.p[r] = reinterpret_cast<uint32_t*>(flag_ptrs[r]);
    }
    const long long nvec = (long long)(elems / 8);
    long long nb = (nvec + 255) / 256;
    const int blocks = (int)(nb > 64 ? 64 : nb);           // all CTAs co-resident: the ticket protocol relies on it
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(256);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    uint4* o = reinterpret_cast<uint4*>(out);
    switch (world) {
        case 2: return (int)cudaLaunchKernelEx(&cfg, allreduce_oneshot_kernel<2>, dp, fp, nvec, rank, o);
        case 4: return (int)cudaLaunchKernelEx(&cfg, allreduce_oneshot_kernel<4>, dp, fp, nvec, rank, o);
        default: return (int)cudaLaunchKernelEx(&cfg, allreduce_oneshot_kernel<8>, dp, fp, nvec, rank, o);
    }
}

int allreduce_sum_f16(void* multicast_ptr, void* const* peer_ptrs, void* const* flag_ptrs, size_t offset_elems, size_t elems, int rank,
                      int world, cudaStream_t stream) {
    if (world < 1 || world > 8 || rank < 0 || rank >= world) return FLEXQ_ERR_BAD_SHAPE;
    if (!multicast_ptr && !peer_ptrs) return FLEXQ_ERR_NULL;
    if (elems % 8 || offset_elems % 8) return FLEXQ_ERR_BAD_SHAPE;             // 16-byte vectors
    if (elems == 0 || world == 1) return 0;
    FlagPtrs fp{};
    if (flag_ptrs)
        for (int r = 0; r < world; r++) {
            if (!flag_ptrs[r]) return FLEXQ_ERR_NULL;
            fp.p[r] = reinterpret_cast<uint32_t*>(flag_ptrs[r]);
        }
    const long long nvec = (long long)(elems / 8), v_off = (long long)(offset_elems / 8);
    const long long per = (nvec + world - 1) / world;
    const long long lo = per * rank, hi = per * (rank + 1);
    const long long vec0 = v_off + (lo < nvec ? lo : nvec), vec1 = v_off + (hi < nvec ? hi : nvec);
    if (vec1 <= vec0 && !flag_ptrs) return 0;              // synced calls always launch: the peers wait for this rank
    const long long work = vec1 > vec0 ? vec1 - vec0 : 0;
    const int threads = g_ar_blocks > 0 ? 1024 : 256;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    const long long cap = g_ar_blocks > 0 ? g_ar_blocks : 4 * sms;
    const int unroll = multicast_ptr ? ar_unroll() : 4;
    long long nb = (work + threads * unroll - 1) / (threads * unroll);
    const int blocks = (int)(nb > cap ? cap : (nb < 1 ? 1 : nb));
    if (multicast_ptr) {
        __half* mc = reinterpret_cast<__half*>(multicast_ptr);
        switch (unroll) {
            case 1: allreduce_multimem_kernel<1><<<blocks, threads, 0, stream>>>(mc, vec0, vec1, fp, rank, world); break;
            case 2: allreduce_multimem_kernel<2><<<blocks, threads, 0, stream>>>(mc, vec0, vec1, fp, rank, world); break;
            case 4: allreduce_multimem_kernel<4><<<blocks, threads, 0, stream>>>(mc, vec0, vec1, fp, rank, world); break;
            default: allreduce_multimem_kernel<8><<<blocks, threads, 0, stream>>>(mc, vec0, vec1, fp, rank, world); break;
        }
        return (int)cudaGetLastError();
    }
    PeerPtrs pp{};
    for (int r = 0; r < world; r++) {
        if (!peer_ptrs[r]) return FLEXQ_ERR_NULL;
        pp.p[r] = reinterpret_cast<__half*>(peer_ptrs[r]);
    }
    switch (world) {
        case 2: allreduce_peer_kernel<2, 4><<<blocks, threads, 0, stream>>>(pp, vec0, vec1, fp, rank); break;
        case 4: allreduce_peer_kernel<4, 2><<<blocks, threads, 0, stream>>>(pp, vec0, vec1, fp, rank); break;
        case 8: allreduce_peer_kernel<8, 1><<<blocks, threads, 0, stream>>>(pp, vec0, vec1, fp, rank); break;
        default: return FLEXQ_ERR_BAD_SHAPE;
    }
    return (int)cudaGetLastError();
}

}  // namespace flexq
