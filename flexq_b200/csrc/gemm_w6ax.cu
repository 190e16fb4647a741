// W6A6 / W6A8 GEMM for sm_100a: TMA-streamed packed 6-bit weights -> in-SM expansion to int8
// -> tcgen05.mma kind::i8 (accumulators in TMEM) -> per-128-k-group scale epilogue -> fp16.
//
// Replaces the reference's bit-serial BMMA kernel FQBMMAKernel::mainLoop
// (/root/reference/engine/src/bgemm/flexq_bmma_kernel.h:119-447) and its launcher
// (flexq_bmma_op.h:64-70,163-188).  Same mathematics:
//     D[m][n] = half( sum_g sx[m,g] * sw[n,g] * S[m,n,g] ),  S = INT32 dot over one 128-k group,
// but S comes from one int8 MMA per 32 k-values instead of x_bits*6 popcount MMAs.
//
// Orientation: the weight tile is the UMMA "A" operand (128 weight rows = UMMA M = TMEM lanes)
// and the activations are the "B" operand (M_TILE tokens = UMMA N = TMEM columns), so decode
// (1..16 tokens) and prefill (256-token tiles) run the same kernel with a different M_TILE.
//
// Work decomposition ("stream-K"): a unit is (n-tile, m-tile, k-group); the U units are split
// evenly over P persistent CTAs (one per SM).  A CTA walks its units as segments of consecutive
// groups of one tile.  A segment covering all groups of its tile stores fp16 directly; partial
// segments are summed in an fp32 slot (red.global.add) and the CTA that completes the tile
// converts, stores and re-zeroes the slot.
//
// Warp roles (512 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc),
// warps 4-7 = weight expanders (6 bit -> int8, swizzled UMMA layout), warps 8-15 = epilogue.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace flexq {

struct GemmParams {
    const uint8_t* w6;
    const __half* w_scale;   // [G][N]
    const float* sx;         // [G][ldsx]
    __half* D;               // [M][N]
    int32_t* S;              // [M][N][G] (DUMP only)
    float* slots;            // [kMaxCtas][M_TILE][128] fp32 partial tiles
    int* cnt;                // [kMaxCtas] group counters
    int M, N, K, G, ldsx;
    int m_tiles;             // tile index = nt * m_tiles + mt
    int U;                   // total units
    long long* trace;        // TRACE builds: [unit][16] clock64 stamps of CTA 0
    int trace_units;
};

#define FQ_TRACE(unit, ev)                                                                   \
    do {                                                                                     \
        if constexpr (TRACE) {                                                               \
            if (blockIdx.x == 0 && (unit) < p.trace_units) p.trace[(unit) * 16 + (ev)] = clock64(); \
        }                                                                                    \
    } while (0)

template <int M_TILE>
struct Cfg {
    static constexpr int SMEM_BUDGET = 224 * 1024;                 // of the 227 KB a CTA may use
    static constexpr int NACC = (512 / M_TILE) < 8 ? (512 / M_TILE) : 8;   // TMEM accumulator buffers
    static constexpr int NA = 3;                                   // expanded-weight stages (16 KB)
    static constexpr int NX = (M_TILE >= 256) ? 3 : (M_TILE >= 128 ? 4 : 6);   // activation stages
    static constexpr int NS = 8;                                   // scale stages (sx row + sw row)
    static constexpr int S_BYTES = M_TILE * 4 + kTileN * 2;        // f32 sx[M_TILE] | f16 sw[128]
    static constexpr int X_BYTES = M_TILE * 128;
    static constexpr int A_BYTES = kTileN * 128;
    static constexpr int NW_FIT = (SMEM_BUDGET - 2048 - NA * A_BYTES - NX * X_BYTES - NS * S_BYTES) / kTileBytes;
    static constexpr int NW = NW_FIT > 12 ? 12 : NW_FIT;           // packed-weight stages (12 KB each)
    static constexpr int OFF_A = 0;
    static constexpr int OFF_X = OFF_A + NA * A_BYTES;
    static constexpr int OFF_W = OFF_X + NX * X_BYTES;
    static constexpr int OFF_S = OFF_W + NW * kTileBytes;
    static constexpr int OFF_BAR = OFF_S + NS * S_BYTES;
    static constexpr int NDONE = 16;                               // "MMAs of unit u retired" ring (> NACC, NX, NA)
    static constexpr int NBAR = 2 * (NW + NS) + NA + NX + NACC + NDONE;
    static constexpr int OFF_MISC = OFF_BAR + NBAR * 8;
    static constexpr int SMEM_BYTES = OFF_MISC + 16 + 1024;        // + alignment slack
    static constexpr int TMEM_COLS = (NACC * M_TILE < 32) ? 32 : NACC * M_TILE;
    static constexpr int CPT = M_TILE / 2;                         // columns per epilogue thread
    static_assert(NW >= 4, "too few weight stages");
    static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS <= 512, "TMEM columns must be a power of two <= 512");
};

// owner(u) = the CTA whose unit range [floor(c*U/P), floor((c+1)*U/P)) contains u
__device__ __forceinline__ int unit_owner(int u, int U, int P) {
    return (int)((((long long)u + 1) * P - 1) / U);
}

template <int M_TILE, bool DUMP, bool TRACE = false>
__global__ void __launch_bounds__(512, 1) w6ax_gemm_kernel(const __grid_constant__ CUtensorMap tmap_x, const GemmParams p) {
    using C = Cfg<M_TILE>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int P = gridDim.x;
    const int u_begin = (int)(((long long)blockIdx.x * p.U) / P);
    const int u_end = (int)(((long long)(blockIdx.x + 1) * p.U) / P);
    const int G = p.G;

    // barrier addresses
    const uint32_t bar0 = smem_base + C::OFF_BAR;
    // full/empty rings.  One tcgen05.commit per unit arrives on bar_done(unit % NDONE); it frees the
    // activation stage and the expanded-weight stage of that unit and publishes its accumulator.
    auto bar_w_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_w_empty = [&](int s) { return bar0 + 8u * (C::NW + s); };
    auto bar_s_full = [&](int s) { return bar0 + 8u * (2 * C::NW + s); };
    auto bar_s_empty = [&](int s) { return bar0 + 8u * (2 * C::NW + C::NS + s); };
    auto bar_a_full = [&](int s) { return bar0 + 8u * (2 * C::NW + 2 * C::NS + s); };
    auto bar_x_full = [&](int s) { return bar0 + 8u * (2 * C::NW + 2 * C::NS + C::NA + s); };
    auto bar_acc_empty = [&](int b) { return bar0 + 8u * (2 * C::NW + 2 * C::NS + C::NA + C::NX + b); };
    auto bar_done = [&](int u) { return bar0 + 8u * (2 * C::NW + 2 * C::NS + C::NA + C::NX + C::NACC + (u % C::NDONE)); };
    auto done_parity = [&](int u) { return (uint32_t)((u / C::NDONE) & 1); };
    uint32_t* misc = reinterpret_cast<uint32_t*>(smem + C::OFF_MISC);   // [0] tmem base, [1] finisher flag

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::NA; s++) mbar_init(bar_a_full(s), 128);
        for (int s = 0; s < C::NX; s++) mbar_init(bar_x_full(s), 1);
        for (int s = 0; s < C::NW; s++) { mbar_init(bar_w_full(s), 1); mbar_init(bar_w_empty(s), 128); }
        for (int s = 0; s < C::NS; s++) { mbar_init(bar_s_full(s), 1); mbar_init(bar_s_empty(s), 256); }
        for (int b = 0; b < C::NACC; b++) mbar_init(bar_acc_empty(b), 256);
        for (int u = 0; u < C::NDONE; u++) mbar_init(bar_done(u), 1);
        fence_barrier_init();
        prefetch_tensormap(&tmap_x);
    }
    if (warp == 1) tmem_alloc<C::TMEM_COLS>(smem_u32(&misc[0]));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = misc[0];

    // Register budget (64K regs, 512 threads): producer/MMA warpgroup 40, expanders 72, the two
    // epilogue warpgroups 200 each (fp32 tile accumulators live in registers).
    // (setmaxnreg sits inside each role branch so that ptxas allocates per role.)
    if (warp == 0) {
        // ===================== TMA producer: packed weight tiles =====================
        // (three independent producer threads -- weights, activations, scales -- so that the HBM
        //  weight stream runs NW stages ahead instead of being gated by the activation ring)
        reg_dealloc<32>();
        if (lane == 0) {
            int iw = 0;
            for (int u = u_begin; u < u_end;) {
                const int tile = u / G, g0 = u - tile * G;
                const int g1 = min(G, g0 + (u_end - u));
                const int nt = tile / p.m_tiles;
                const uint8_t* wsrc = p.w6 + ((size_t)nt * G) * kTileBytes;
                for (int g = g0; g < g1; g++, iw++) {
                    const int s = iw % C::NW;
                    mbar_wait(bar_w_empty(s), ((iw / C::NW) & 1) ^ 1);
                    FQ_TRACE(iw, 0);
                    mbar_expect_tx(bar_w_full(s), kTileBytes);
                    bulk_g2s(smem_base + C::OFF_W + s * kTileBytes, wsrc + (size_t)g * kTileBytes, kTileBytes, bar_w_full(s));
                }
                u += g1 - g0;
            }
        }
        __syncwarp();
    } else if (warp == 2) {
        // ===================== TMA producer: activation tiles =====================
        reg_dealloc<32>();
        if (lane == 0) {
            int ix = 0;
            for (int u = u_begin; u < u_end;) {
                const int tile = u / G, g0 = u - tile * G;
                const int g1 = min(G, g0 + (u_end - u));
                const int nt = tile / p.m_tiles, mt = tile - nt * p.m_tiles;
                for (int g = g0; g < g1; g++, ix++) {   // M_TILE rows x 128 B, swizzle-128B, rows >= M zero-filled
                    const int s = ix % C::NX;
                    if (ix >= C::NX) mbar_wait(bar_done(ix - C::NX), done_parity(ix - C::NX));
                    FQ_TRACE(ix, 9);
                    mbar_expect_tx(bar_x_full(s), C::X_BYTES);
                    tma_load_2d(smem_base + C::OFF_X + s * C::X_BYTES, &tmap_x, g * kGroup, mt * M_TILE, bar_x_full(s));
                }
                u += g1 - g0;
            }
        }
        __syncwarp();
    } else if (warp == 3) {
        // ===================== TMA producer: scales =====================
        reg_dealloc<32>();
        if (lane == 0 && !DUMP) {
            int is = 0;
            for (int u = u_begin; u < u_end;) {
                const int tile = u / G, g0 = u - tile * G;
                const int g1 = min(G, g0 + (u_end - u));
                const int nt = tile / p.m_tiles, mt = tile - nt * p.m_tiles;
                const int m0 = mt * M_TILE;
                const int cols = min(M_TILE, p.ldsx - m0);
                const int rows = min(kTileN, p.N - nt * kTileN);
                for (int g = g0; g < g1; g++, is++) {   // sx[g][m0..] (f32) and w_scale[g][n0..] (f16)
                    const int s = is % C::NS;
                    const uint32_t dst = smem_base + C::OFF_S + s * C::S_BYTES;
                    mbar_wait(bar_s_empty(s), ((is / C::NS) & 1) ^ 1);
                    mbar_expect_tx(bar_s_full(s), cols * 4 + rows * 2);
                    bulk_g2s(dst, p.sx + (size_t)g * p.ldsx + m0, cols * 4, bar_s_full(s));
                    bulk_g2s(dst + M_TILE * 4, p.w_scale + (size_t)g * p.N + nt * kTileN, rows * 2, bar_s_full(s));
                }
                u += g1 - g0;
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        reg_dealloc<32>();
        {
            constexpr uint32_t idesc = umma_idesc_i8(kTileN, M_TILE);
            const int n_units = u_end - u_begin;
            for (int i = 0; i < n_units; i++) {
                const int buf = i % C::NACC;
                const int sa = i % C::NA, sx_ = i % C::NX;
                // (divergent per-lane waits were measured slower than one lane waiting in turn)
                if (lane == 0) {
                    mbar_wait(bar_acc_empty(buf), (i / C::NACC) & 1);   // armed (biased) by the epilogue
                    mbar_wait(bar_x_full(sx_), (i / C::NX) & 1);
                    mbar_wait(bar_a_full(sa), (i / C::NA) & 1);
                }
                if (lane == 0) {
                    FQ_TRACE(i, 4);
                    tc_fence_after();
                    const uint32_t a_addr = smem_base + C::OFF_A + sa * C::A_BYTES;
                    const uint32_t b_addr = smem_base + C::OFF_X + sx_ * C::X_BYTES;
                    const uint32_t d_tmem = tmem_base + buf * M_TILE;
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        umma_i8(d_tmem, umma_desc_sw128(a_addr + 32 * k), umma_desc_sw128(b_addr + 32 * k), idesc, 1u);
                    umma_commit(bar_done(i));
                    FQ_TRACE(i, 5);
                }
            }
            __syncwarp();
        }
    } else if (warp < 8) {
        // ===================== weight expanders =====================
        reg_dealloc<64>();
        const int r = threadIdx.x - 128;                 // weight row within the tile
        const int n_units = u_end - u_begin;
        for (int i = 0; i < n_units; i++) {
            const int sw = i % C::NW, sa = i % C::NA;
            mbar_wait(bar_w_full(sw), (i / C::NW) & 1);
            if (r == 0) FQ_TRACE(i, 1);
            const uint8_t* wp = smem + C::OFF_W + sw * kTileBytes;
            uint4 in[2][3];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const uint4* src = reinterpret_cast<const uint4*>(wp + 48 * (128 * q + r));
                in[q][0] = src[0]; in[q][1] = src[1]; in[q][2] = src[2];
            }
            mbar_arrive(bar_w_empty(sw));                // packed tile fully in registers
            if (i >= C::NA) mbar_wait(bar_done(i - C::NA), done_parity(i - C::NA));
            if (r == 0) FQ_TRACE(i, 2);
            uint8_t* arow = smem + C::OFF_A + sa * C::A_BYTES + r * 128;
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const uint32_t w[12] = {in[q][0].x, in[q][0].y, in[q][0].z, in[q][0].w, in[q][1].x, in[q][1].y,
                                        in[q][1].z, in[q][1].w, in[q][2].x, in[q][2].y, in[q][2].z, in[q][2].w};
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    uint32_t o[4];
                    w6_expand16(w[3 * s], w[3 * s + 1], w[3 * s + 2], o);
                    const int chunk = (4 * q + s) ^ (r & 7);     // 128-byte swizzle
                    *reinterpret_cast<uint4*>(arow + 16 * chunk) = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(bar_a_full(sa));
            if (r == 0) FQ_TRACE(i, 3);
        }
    } else {
        // ===================== epilogue =====================
        reg_alloc<208>();
        // Accumulators are re-armed with the bit pattern of 1.5*2^23 before every group, so the
        // int32 sum 4*S read back from TMEM *is* the float (kMagicF + 4*S): one FMA with the
        // per-row weight scale removes the bias exactly (kMagicF*sw is exact in fp32 for an
        // fp16-valued sw) and a second FMA applies the per-token scale and accumulates.
        constexpr int CPT = C::CPT;
        constexpr int CH = CPT >= 128 ? 16 : (CPT < 32 ? CPT : 32);   // columns per tcgen05.ld (register budget)
        constexpr uint32_t kMagicI = 0x4B400000u;
        constexpr float kMagicF = 12582912.f;
        const int e = threadIdx.x - 256;
        const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
        const int half_id = e >> 7;
        const int r = quad * 32 + lane;
        const int col0 = half_id * CPT;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + col0;
        // arm every accumulator buffer once
#pragma unroll
        for (int b = 0; b < C::NACC; b++) {
#pragma unroll
            for (int c = 0; c < CPT; c += (CPT < 16 ? 8 : 16)) {
                if constexpr (CPT < 16) tmem_st8_same(t_lane + b * M_TILE + c, kMagicI);
                else tmem_st16_same(t_lane + b * M_TILE + c, kMagicI);
            }
        }
        tmem_wait_st();
        tc_fence_before();
#pragma unroll
        for (int b = 0; b < C::NACC; b++) mbar_arrive(bar_acc_empty(b));

        float2 acc[CPT / 2];
        int i = 0, is = 0;
        for (int u = u_begin; u < u_end;) {
            const int tile = u / G, g0 = u - tile * G;
            const int g1 = min(G, g0 + (u_end - u));
            const int nt = tile / p.m_tiles, mt = tile - nt * p.m_tiles;
            const int n = nt * kTileN + r;
            const int mbase = mt * M_TILE + col0;
            const bool n_ok = n < p.N;
#pragma unroll
            for (int j = 0; j < CPT / 2; j++) acc[j] = make_float2(0.f, 0.f);
            for (int g = g0; g < g1; g++, i++) {
                const int buf = i % C::NACC;
                float2 sw2 = make_float2(0.f, 0.f), bias2 = make_float2(0.f, 0.f);
                const float* sxs = nullptr;
                int ss = 0;
                if (!DUMP) {
                    ss = is % C::NS;
                    mbar_wait(bar_s_full(ss), (is / C::NS) & 1);
                    is++;
                    const uint8_t* st = smem + C::OFF_S + ss * C::S_BYTES;
                    sxs = reinterpret_cast<const float*>(st) + col0;
                    const float swv = n_ok ? 0.25f * __half2float(reinterpret_cast<const __half*>(st + M_TILE * 4)[r]) : 0.f;   // operands hold 4*w
                    sw2 = make_float2(swv, swv);
                    bias2 = make_float2(-kMagicF * swv, -kMagicF * swv);
                }
                mbar_wait(bar_done(i), done_parity(i));
                if (e == 0) FQ_TRACE(i, 6);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < CPT; c += CH) {
                    uint32_t v[CH];
                    const uint32_t ta = t_lane + buf * M_TILE + c;
                    if constexpr (CH == 8) tmem_ld8(ta, v);
                    else if constexpr (CH == 16) tmem_ld16(ta, v);
                    else tmem_ld32(ta, v);
                    tmem_wait_ld();
                    // re-arm this chunk of the accumulator for its next group
                    if constexpr (CH == 8) tmem_st8_same(ta, kMagicI);
                    else {
#pragma unroll
                        for (int cc = 0; cc < CH; cc += 16) tmem_st16_same(ta + cc, kMagicI);
                    }
                    if (c + CH >= CPT) {                 // accumulator fully read and re-armed: hand it back
                        tmem_wait_st();
                        tc_fence_before();
                        mbar_arrive(bar_acc_empty(buf));
                        if (e == 0) FQ_TRACE(i, 7);
                    }
                    if constexpr (DUMP) {
                        if (n_ok) {
#pragma unroll
                            for (int j = 0; j < CH; j++) {
                                const int m = mbase + c + j;
                                if (m < p.M) p.S[((size_t)m * p.N + n) * G + g] = ((int32_t)(v[j] - kMagicI)) >> 2;
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < CH; j += 4) {
                            const float4 s4 = *reinterpret_cast<const float4*>(sxs + c + j);
                            const float2 t0 = __ffma2_rn(make_float2(__uint_as_float(v[j + 0]), __uint_as_float(v[j + 1])), sw2, bias2);
                            const float2 t1 = __ffma2_rn(make_float2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), sw2, bias2);
                            acc[(c + j) / 2] = __ffma2_rn(t0, make_float2(s4.x, s4.y), acc[(c + j) / 2]);
                            acc[(c + j) / 2 + 1] = __ffma2_rn(t1, make_float2(s4.z, s4.w), acc[(c + j) / 2 + 1]);
                        }
                    }
                }
                if (!DUMP) mbar_arrive(bar_s_empty(ss));
                if (e == 0) FQ_TRACE(i, 8);
            }
            if constexpr (!DUMP) {
                if (g0 == 0 && g1 == G) {
                    // whole tile reduced by this CTA: store fp16 directly
                    if (n_ok) {
#pragma unroll
                        for (int j = 0; j < CPT; j++) {
                            const int m = mbase + j;
                            const float a = (j & 1) ? acc[j / 2].y : acc[j / 2].x;
                            if (m < p.M) p.D[(size_t)m * p.N + n] = __float2half_rn(a);
                        }
                    }
                } else {
                    // partial tile: accumulate in the slot owned by the tile's first CTA
                    const int slot = unit_owner(tile * G, p.U, P);
                    float* sl = p.slots + (size_t)slot * kSlotFloats + (size_t)col0 * kTileN + r;
#pragma unroll
                    for (int j = 0; j < CPT; j++) atomicAdd(sl + j * kTileN, (j & 1) ? acc[j / 2].y : acc[j / 2].x);
                    named_bar_sync(1, 256);              // every thread's red.adds are issued ...
                    if (e == 0) {
                        __threadfence();                 // ... and released (cumulatively) by one fence
                        const int old = atomicAdd(p.cnt + slot, g1 - g0);
                        __threadfence();
                        misc[1] = (old + (g1 - g0) == G) ? 1u : 0u;
                    }
                    named_bar_sync(1, 256);
                    const bool last = misc[1] != 0;
                    if (last) {
#pragma unroll
                        for (int j = 0; j < CPT; j++) {
                            const float vsum = __ldcg(sl + j * kTileN);
                            __stcg(sl + j * kTileN, 0.f);
                            const int m = mbase + j;
                            if (n_ok && m < p.M) p.D[(size_t)m * p.N + n] = __float2half_rn(vsum);
                        }
                        if (e == 0) p.cnt[slot] = 0;
                    }
                    named_bar_sync(1, 256);              // flag word is reused by the next partial segment
                }
            }
            u += g1 - g0;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<C::TMEM_COLS>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode_fn() {
    static PFN_tmapEncodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmapEncodeTiled>(ptr);
    }
    return fn;
}

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 0;
    }
    return n;
}

template <int M_TILE, bool DUMP, bool TRACE = false>
static int launch(const int8_t* xq, const GemmParams& p_in, cudaStream_t stream) {
    using C = Cfg<M_TILE>;
    static_assert(C::SMEM_BYTES <= 232448, "shared memory budget exceeded");
    GemmParams p = p_in;
    PFN_tmapEncodeTiled enc = get_encode_fn();
    const int sms = num_sms();
    if (!enc || sms <= 0) return FLEXQ_ERR_NO_DEVICE;

    CUtensorMap tmap;
    const cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.M};
    const cuuint64_t strides[1] = {(cuuint64_t)p.K};
    const cuuint32_t box[2] = {128u, (cuuint32_t)M_TILE};
    const cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t*>(xq), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return FLEXQ_ERR_TENSORMAP;

    const int n_tiles = ceil_div(p.N, kTileN);
    p.m_tiles = ceil_div(p.M, M_TILE);
    p.U = n_tiles * p.m_tiles * p.G;
    const int P = p.U < sms ? p.U : (sms < kMaxCtas ? sms : kMaxCtas);

    static bool attr_set = false;
    if (!attr_set) {
        FLEXQ_CUDA_TRY(cudaFuncSetAttribute(w6ax_gemm_kernel<M_TILE, DUMP, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    w6ax_gemm_kernel<M_TILE, DUMP, TRACE><<<P, 512, C::SMEM_BYTES, stream>>>(tmap, p);
    return (int)cudaGetLastError();
}

// experiment knob: FLEXQ_MTILE_CAP=128 keeps prefill on the 128-token tile
static int mtile_cap() {
    static int cap = 0;
    if (cap == 0) {
        const char* e = getenv("FLEXQ_MTILE_CAP");
        cap = e ? atoi(e) : 256;
    }
    return cap;
}

template <bool DUMP>
static int dispatch(const int8_t* xq, const GemmParams& p, cudaStream_t stream) {
    if (p.M <= 16) return launch<16, DUMP>(xq, p, stream);
    if (p.M <= 32) return launch<32, DUMP>(xq, p, stream);
    if (p.M <= 64) return launch<64, DUMP>(xq, p, stream);
    if (p.M <= 128 || mtile_cap() < 256) return launch<128, DUMP>(xq, p, stream);
    return launch<256, DUMP>(xq, p, stream);
}

int gemm_w6ax(const int8_t* xq, const float* sx, const uint8_t* w6, const __half* w_scale, __half* D, int M, int N, int K,
              void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!xq || !sx || !w6 || !w_scale || !D || !workspace) return FLEXQ_ERR_NULL;
    if (M <= 0 || N <= 0 || K < kGroup || K % kGroup) return FLEXQ_ERR_BAD_SHAPE;
    if (workspace_bytes < flexq_gemm_workspace_bytes() || ((uintptr_t)workspace & 15)) return FLEXQ_ERR_WORKSPACE;
    GemmParams p{};
    p.w6 = w6; p.w_scale = w_scale; p.sx = sx; p.D = D; p.S = nullptr;
    p.cnt = reinterpret_cast<int*>(workspace);
    p.slots = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + kCntBytes);
    p.M = M; p.N = N; p.K = K; p.G = K / kGroup; p.ldsx = ceil4(M);
    return dispatch<false>(xq, p, stream);
}

// debug: same GEMM with clock64 stamps of CTA 0's pipeline events (tools/trace.py)
int gemm_w6ax_trace(const int8_t* xq, const float* sx, const uint8_t* w6, const __half* w_scale, __half* D, int M, int N, int K,
                    void* workspace, long long* trace, int trace_units, cudaStream_t stream) {
    GemmParams p{};
    p.w6 = w6; p.w_scale = w_scale; p.sx = sx; p.D = D; p.S = nullptr;
    p.cnt = reinterpret_cast<int*>(workspace);
    p.slots = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + kCntBytes);
    p.M = M; p.N = N; p.K = K; p.G = K / kGroup; p.ldsx = ceil4(M);
    p.trace = trace; p.trace_units = trace_units;
    if (M <= 16) return launch<16, false, true>(xq, p, stream);
    return launch<256, false, true>(xq, p, stream);
}

int gemm_w6ax_groupsums(const int8_t* xq, const uint8_t* w6, int32_t* S, int M, int N, int K, cudaStream_t stream) {
    if (!xq || !w6 || !S) return FLEXQ_ERR_NULL;
    if (M <= 0 || N <= 0 || K < kGroup || K % kGroup) return FLEXQ_ERR_BAD_SHAPE;
    GemmParams p{};
    p.w6 = w6; p.S = S;
    p.M = M; p.N = N; p.K = K; p.G = K / kGroup; p.ldsx = ceil4(M);
    return dispatch<true>(xq, p, stream);
}

}  // namespace flexq
