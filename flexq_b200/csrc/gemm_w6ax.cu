// W6A6 / W6A8 GEMM for sm_100a: TMA-streamed packed 6-bit weights -> in-SM expansion to int8
// -> tcgen05.mma kind::i8 (weights read from TMEM, activations from shared memory, accumulators
// in TMEM) -> per-128-k-group scale epilogue -> fp16.
//
// Replaces the reference's bit-serial BMMA kernel FQBMMAKernel::mainLoop
// (/root/reference/engine/src/bgemm/flexq_bmma_kernel.h:119-447) and its launcher
// (flexq_bmma_op.h:64-70,163-188).  Same mathematics:
//     D[m][n] = half( sum_g sx[m,g] * sw[n,g] * S[m,n,g] ),  S = INT32 dot over one 128-k group,
// but S comes from one int8 MMA per 32 k-values instead of x_bits*6 popcount MMAs.
//
// Orientation: the weight tile is the UMMA "A" operand (128 weight rows = UMMA M = TMEM lanes)
// and the activations are the "B" operand (M_TILE tokens = UMMA N = TMEM columns), so decode
// (1..16 tokens) and prefill (192-token tiles) run the same kernel with different tile constants.
// The expanded weights never touch shared memory: each expander thread owns one weight row,
// turns its 96 packed bytes into 128 int8 in registers and writes them to TMEM with one
// tcgen05.st, where the MMA reads them as its A operand.
//
// Work decomposition: a unit is (token tile, n-tile, k-group); P persistent CTAs (one per SM) each walk a contiguous
// range of units as segments of consecutive groups of one tile, GP groups per pipeline step.  plan_ctas() prices two
// plans -- stream-K over all CTAs, or (small problems) an aligned plan that cuts every tile into a whole number of runs
// -- and the host picks the token tile (128 / 192) the same way.  A segment covering all groups of its tile stores fp16
// directly (fragment layout -> swizzled staging tile -> TMA store); tiles cut by a range boundary are summed by parked
// hand-off (128- / 192-token tiles), vector reductions into a shared fp32 slot, or a thread-block-cluster exchange
// through distributed shared memory (16-token tile, aligned plan) -- see the epilogue.
//
// Warp roles (CTRL_HIGH layout; 512 threads, 640 for the 192-token tile): epilogue warpgroups first (2 or 3), then the
// four weight expanders, then W producer (TMA, weights), two MMA issuers (alternate steps; the first owns the TMEM
// allocation) and X producer (TMA, activations + scales) on the highest warp ids.
//
// Experiment switches (-D...) that stayed in the source are off where they measured slower; profiles/r2_experiments/
// holds the A/B records and DESIGN.md 3.3 the list.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace flexq {

struct GemmParams {
    const uint8_t* w6;
    __half* D;               // [M][N]
    int32_t* S;              // [M][N][G] (DUMP only)
    float* slots;            // [kMaxCtas + kSlotPool][kSlotFloats] fp32 tiles: accumulation slots, then the parking pool
    int* cnt;                // [kMaxCtas][kRecInts] cut-tile records + the pool's bump counter (all zero between launches
                             // except the bump counter)
    int M, N, K, G;
    int m_tiles, n_tiles;    // tile index = mt * n_tiles + nt: CTAs that run concurrently stream the same weight rows
                             // (one token tile each), so a weight row is fetched from HBM once and hit in L2 by the rest
    int P, Pn, R;            // P CTAs: R rows x Pn columns of regular CTAs + (P - R*Pn) spare ones (see Sched)
    int Ureg;                // units [0, Ureg) of every token tile belong to the regular CTAs
    int whole_rows;          // 1: more token tiles than CTAs, every CTA owns whole token tiles
    int handoff;             // 192-token tile: 1 = cut tiles are summed by parked hand-off (stream-K plan: contributors arrive
                             // far apart; aligned plan with two runs per tile), 0 = by reductions into a shared slot
                             // (aligned plan with more runs per tile: they all finish together)
    int ldd;                 // row pitch of D in elements (N; the number of gate/up pairs with silu)
    int silu;                // 1: W6 rows alternate 8 gate rows / 8 up rows (flexq_gemm_w6ax_silu_mul); the epilogue writes
                             // half(silu(half(gate)) * half(up)) for every pair: D is [M][N / 2]
    unsigned nonce;          // cluster exchange: a value no earlier launch left in shared memory (the "armed" flag)
    int cluster;             // > 1: launched as thread-block clusters of this many CTAs, the runs of one weight tile; its
                             // partial sums meet in the first CTA's shared memory (decode tiles, aligned plan)
    long long* trace;        // TRACE builds: [step][16] clock64 stamps of CTA 0
    int trace_units;
};

#define FQ_TRACE(unit, ev)                                                                   \
    do {                                                                                     \
        if constexpr (TRACE) {                                                               \
            if (blockIdx.x == (p.trace_units >> 16) && (unit) < (p.trace_units & 0xFFFF)) p.trace[(unit) * 16 + (ev)] = clock64(); \
        }                                                                                    \
    } while (0)

// Accumulator seeding by the tensor core itself ("bias MMA"): the first MMA of every k-group is an
// unsigned 255 x 255, K = 32 product of two constant operands, i.e. it overwrites the accumulator with
// kBiasB = 32*255*255 = 2080800 >= max(-4S) = 128*127*128, and the four weight MMAs accumulate on top.  The
// accumulator then holds the non-negative integer B + 4S < 2^23, whose bit pattern *is* the fp32 subnormal
// (B + 4S) * 2^-149: the epilogue feeds it to an FMA unconverted (no per-element integer op at all) and
// carries the power-of-two factors in its scale constants.  One more MMA per group on a tensor pipe that
// has headroom, against one issue slot per output element per group saved.
#ifndef FLEXQ_BIASMMA
#define FLEXQ_BIASMMA 1
#endif
constexpr uint32_t kBiasB = 32u * 255u * 255u;

// experiment switches of the epilogue (defaults = the measured best)
#ifndef FLEXQ_EPI_WG
#define FLEXQ_EPI_WG 3
#endif
#ifndef FLEXQ_EPI_FRAG
#define FLEXQ_EPI_FRAG 1
#endif
#ifndef FLEXQ_EPI_TSTORE
#define FLEXQ_EPI_TSTORE 1
#endif
#ifndef FLEXQ_CLUSTER
#define FLEXQ_CLUSTER 1
#endif
#ifndef FLEXQ_CLUSTER_ASYNC
#define FLEXQ_CLUSTER_ASYNC 1     // cluster exchange by st.async + an mbarrier of the first CTA instead of two cluster barriers
#endif
#ifndef FLEXQ_CLUSTER_FLAG
#define FLEXQ_CLUSTER_FLAG 1      // "receive barrier armed" is published by a flag word instead of a cluster barrier
#endif
#ifndef FLEXQ_EXP_CONSTS
#define FLEXQ_EXP_CONSTS 0       // measured 2 % slower at M >= 512, 10-40 % on small layers: the expanders wait for the scale block
#endif

#ifndef FLEXQ_HANDOFF_MIN_TILE
#define FLEXQ_HANDOFF_MIN_TILE 128
#endif
constexpr int kHandoffMinTile = FLEXQ_HANDOFF_MIN_TILE;   // token tiles from this size up sum cut tiles by parked hand-off (run-time rule in launch())
constexpr int kClusterMaxTile = 16;      // largest token tile that sums cut tiles through a thread-block cluster

template <int M_TILE, int GP>
struct Cfg {
    static constexpr bool BIAS = (FLEXQ_BIASMMA != 0) && (M_TILE >= 128);
    // constant 0xFF operand region of the bias MMA: K-major, no swizzle, 8-row x 16-byte core matrices, two per
    // 8-row group (K = 32 bytes); shared by both operands (every byte is the same)
    static constexpr int ONES_BYTES = BIAS ? ((M_TILE > 128 ? M_TILE : 128) / 8) * 256 : 0;
    // fp16 staging tile of the output ([token][128 weight rows], two swizzle-128B boxes of 64 rows per epilogue
    // warpgroup): the finished tile leaves through TMA stores instead of 2-byte scattered global stores
    static constexpr bool TSTORE = (FLEXQ_EPI_TSTORE != 0) && (FLEXQ_EPI_FRAG != 0) && (FLEXQ_EPI_WG == 3) && BIAS && GP == 1;
    static constexpr int STAGE_BYTES = TSTORE ? M_TILE * kTileN * 2 : 0;
    static constexpr int SMEM_BUDGET = 222 * 1024 - ONES_BYTES - STAGE_BYTES;    // of the 227 KB a CTA may use
    // ---- TMEM columns: NAB accumulator step-buffers (GP groups x M_TILE) + NAT weight stages (GP x 32)
    static constexpr int ACC_COLS = GP * M_TILE;
    static constexpr int A_COLS = GP * 32;
    static constexpr int NAB = (M_TILE * GP <= 64) ? 4 : 2;
    static constexpr int NAT_FIT = (512 - NAB * ACC_COLS) / A_COLS;
    static constexpr int NAT = NAT_FIT > 4 ? 4 : NAT_FIT;
    static constexpr int A_COL0 = NAB * ACC_COLS;
    static_assert(NAT >= 2 && NAB * ACC_COLS + NAT * A_COLS <= 512, "TMEM budget");
    // ---- shared memory rings
    // activation / scale stages.  Decode tiles: the stages are small (8 KB + 1.3 KB at 16 tokens) but their loads queue
    // behind the saturated weight stream (measured 3000-5000 cycles from issue to arrival, two to three steps): with two
    // stages every MMA waited for its activations, the TMEM weight stages and then the weight ring backed up behind it
#ifndef FLEXQ_NX_DECODE
#define FLEXQ_NX_DECODE 6
#endif
    static constexpr int NX = (M_TILE >= 128) ? 4 : (M_TILE <= 16 ? FLEXQ_NX_DECODE : 4);
    static constexpr int NS = (M_TILE >= 128) ? 4 : (M_TILE <= 16 ? FLEXQ_NX_DECODE : 4);
    static constexpr int X_BYTES = GP * M_TILE * 128;              // [GP][M_TILE][128 B], swizzle-128B
    static constexpr int SX_BYTES = GP * M_TILE * 4;               // f32 [GP][M_TILE]
    static constexpr int SW_BYTES = GP * kTileN * 2;               // f16 [GP][128]
    // experiment (off): per-row scale constants (c1, c2) of the bias-MMA epilogue written once per step by the expander
    // thread that owns the row instead of being recomputed from the fp16 scale by every epilogue warpgroup (LDS.U16 +
    // convert + 2 FMUL per row and warpgroup -> one LDS.64).  12 fewer instructions per epilogue warp-step, but the
    // expanders then depend on the scale block: profiles/r2_experiments/sweep_b24_*
    static constexpr bool XCONST = (FLEXQ_EXP_CONSTS != 0) && (FLEXQ_EPI_FRAG != 0) && (FLEXQ_EPI_WG == 3) && BIAS && GP == 1;
    static constexpr int CONST_BYTES = XCONST ? kTileN * 8 : 0;
    static constexpr int S_BYTES = SX_BYTES + SW_BYTES + CONST_BYTES;
    static constexpr int W_BYTES = GP * kTileBytes;                // GP consecutive packed tiles
    static constexpr int NW_FIT = (SMEM_BUDGET - NX * X_BYTES - NS * S_BYTES) / W_BYTES;
    static constexpr int NW = NW_FIT > 10 ? 10 : NW_FIT;
    static_assert(NW >= 3, "too few weight stages");
    // Two MMA issuer warps take alternate steps.  A ring consumed by the issuers must have an even
    // number of stages so that each stage is always consumed by the same issuer: an mbarrier parity
    // wait cannot tell phase p from phase p+2, so an issuer must never skip a phase of a barrier.
    static_assert(NX % 2 == 0 && NAT % 2 == 0 && NAB % 2 == 0, "issuer-consumed rings need even depth");
    static constexpr int OFF_STAGE = 0;
    static constexpr int OFF_X = STAGE_BYTES;
    static constexpr int OFF_W = OFF_X + NX * X_BYTES;
    static constexpr int OFF_S = OFF_W + NW * W_BYTES;
    static constexpr int OFF_BAR = OFF_S + NS * S_BYTES;
    static constexpr int NDONE = 16;                               // "MMAs of step i retired" ring (> NAB, NAT, NX)
    static constexpr int NBAR = 2 * (NW + NS) + NAT + NX + NAB + NDONE + 1;        // + the cluster exchange's receive barrier
    static constexpr int OFF_MISC = OFF_BAR + NBAR * 8;
    static constexpr int OFF_ONES = (OFF_MISC + 32 + 4 * kMaxParked + 127) / 128 * 128;   // misc: tmem base, flags, parked slot ids
    // cluster exchange (decode tiles, FLEXQ_CLUSTER_ASYNC): room for the partial tiles of up to three other CTAs
    static constexpr int RECV_BYTES = (FLEXQ_CLUSTER != 0 && FLEXQ_CLUSTER_ASYNC != 0 && M_TILE <= kClusterMaxTile) ? 3 * M_TILE * kTileN * 4 : 0;
    static constexpr int OFF_RECV = OFF_ONES + ONES_BYTES;
    static constexpr int SMEM_BYTES = OFF_RECV + RECV_BYTES + 1024; // + alignment slack
    // epilogue warpgroups.  Measured on B200 (70B shapes, M >= 2048): 3 warpgroups of 64 columns (12 warps, 128 regs)
    // beat 2 x 96 columns (8 warps, 200 regs) by 3-8 % on the 192-token tile -- one more warp per scheduler to
    // cover FFMA2 dependencies; 4 x 48 columns are 4 % slower again (per-warp step overhead).  The 128-token
    // tile and the decode tiles keep 2.
    static constexpr int EPI_WG = (FLEXQ_EPI_WG == 4 && M_TILE >= 128) ? 4 : (FLEXQ_EPI_WG == 3 && M_TILE == 192) ? 3 : 2;
    static constexpr int EPI_THREADS = 128 * EPI_WG;
    // Warp roles.  The warp scheduler favours the highest warp id among eligible warps (measured model in
    // B300_MICROARCH.md "Multi-warp arbiter"), so the short latency-critical roles -- MMA issue, TMA issue -- get the
    // highest ids and the bulk math the lowest: a woken issuer must not queue behind twelve epilogue warps.
#ifndef FLEXQ_CTRL_HIGH
#define FLEXQ_CTRL_HIGH 1
#endif
    static constexpr bool CTRL_HIGH = FLEXQ_CTRL_HIGH != 0;
    // Expander warps: one per TMEM lane quadrant, or two (each takes every other k-group of a step) for the decode tiles
    // with four groups per step: there a step's expansion (1000 cycles with four warps) is the longest part of the tail
    // that follows the last weight bytes, and of the time a ring stage stays occupied.
#ifndef FLEXQ_EXP8
#define FLEXQ_EXP8 0        // measured: no gain (+-1 %, profiles/r2_experiments/sweep_b25_*): the tail is skew between CTAs, not expansion
#endif
    static constexpr int EXP_WARPS = (FLEXQ_EXP8 != 0 && CTRL_HIGH && M_TILE <= 32 && GP == 4) ? 8 : 4;
    static constexpr int THREADS = 128 + 32 * EXP_WARPS + EPI_THREADS;
    static constexpr int EPI_WARP0 = CTRL_HIGH ? 0 : 8;                    // EPI_WG warpgroups
    static constexpr int EXP_WARP0 = CTRL_HIGH ? 4 * EPI_WG : 4;           // EXP_WARPS expander warps
    static constexpr int CTRL_WARP0 = CTRL_HIGH ? 4 * EPI_WG + EXP_WARPS : 0;      // W producer, issuer, issuer, X producer
    // setmaxnreg pool = registers the CTA is launched with (regs/thread x THREADS: 80 x 768, 96 x 640 or 128 x 512):
    //   768 threads: 128*32 + 128*64 + 512*96 = 61440;  640: 128*32 + 128*64 + 384*128 = 61440;  512: 128*32 + 128*72 + 256*200 = 64512
    //   640 threads with eight expander warps: 128*32 + 256*72 + 256*128 = 55296
    static constexpr int EPI_REGS = (EPI_WG == 4) ? 96 : (EPI_WG == 3 || EXP_WARPS == 8) ? 128 : 200;
    static constexpr int EXP_REGS = (EPI_WG >= 3) ? 64 : 72;
    static constexpr int CPT = M_TILE / EPI_WG;                    // columns per epilogue thread
#ifndef FLEXQ_LOPS_BIG
#define FLEXQ_LOPS_BIG 2
#endif
#ifndef FLEXQ_LOPS_SMALL
#define FLEXQ_LOPS_SMALL 2
#endif
    // of every 4 accumulator elements, how many get their float bias by LOP3 (ALU pipe) instead of an
    // integer add (FMA pipe): balances the two pipes against the expander's ALU work (measured per tile)
    static constexpr int MAGIC_LOPS = (M_TILE >= 192) ? FLEXQ_LOPS_BIG : FLEXQ_LOPS_SMALL;
    // Fragment layout of the TMEM drain (tcgen05.ld.16x256b): a thread holds 4 rows x 16 columns of its warpgroup's
    // 128 x 64 slice instead of 1 row x 64 columns, so it needs 16 token scales per k-group instead of 64 (8 LDS.64
    // touching 32 contiguous bytes per warp instead of 16 broadcast LDS.128) -- the shared-memory data pipe, which
    // also feeds the MMA's activation operand and takes the TMA writes, was the saturated resource.
    static constexpr bool FRAG = (FLEXQ_EPI_FRAG != 0) && BIAS && CPT == 64 && GP == 1;
    static constexpr int CH = FRAG ? 16 : (EPI_WG >= 3) ? 16 : (CPT < 32 ? CPT : 32);   // accumulator values per tcgen05.ld
    static_assert(CPT % CH == 0 && (CH == 8 || CH == 16 || CH == 32), "epilogue chunking");
};

// Accumulator biasing strategy.  true: the epilogue re-arms TMEM with the bit pattern of 1.5*2^23 after
// every read (tcgen05.st) and all MMAs accumulate; false: the first MMA of a group overwrites
// (accumulate = 0) and the epilogue adds the bias with one integer add per element.
#ifndef FLEXQ_REARM
#define FLEXQ_REARM 0
#endif
constexpr bool kRearm = FLEXQ_REARM != 0;
#ifndef FLEXQ_ISSUER_WAITS_SCALES
#define FLEXQ_ISSUER_WAITS_SCALES 1
#endif
constexpr bool kIssuerWaitsScales = FLEXQ_ISSUER_WAITS_SCALES != 0;
#ifndef FLEXQ_FIXUP_HANDOFF
#define FLEXQ_FIXUP_HANDOFF 1
#endif
#ifndef FLEXQ_EPI_PREFETCH
#define FLEXQ_EPI_PREFETCH 0
#endif

// owner(u) = the CTA whose unit range [floor(c*U/P), floor((c+1)*U/P)) contains u
__host__ __device__ __forceinline__ int unit_owner(int u, int U, int P) {
    return (int)((((long long)u + 1) * P - 1) / U);
}

// Work decomposition.  A unit is (token tile mt, n-tile nt, k-group g); within a token tile the units are ordered
// u = nt * G + g (Umt = n_tiles * G of them).
//   * R = m_tiles rows x Pn columns of "regular" CTAs: row r owns token tile r, the Pn columns split units [0, Ureg)
//     of every token tile into the same contiguous ranges, so the CTAs of a column stream identical weight tiles at
//     the same time (one HBM fetch, L2 hits for the other rows).
//   * the P - R*Pn CTAs that do not fill another column ("spare") share units [Ureg, Umt) of all token tiles,
//     linearised as (mt, u); Ureg is chosen so that every CTA gets the same number of units (+-1).
//   * more token tiles than CTAs: every CTA owns whole token tiles mt = cta, cta + P, ...
// Splits are at k-group granularity (stream-K): a tile cut by a range boundary is summed through the fp32 slot of
// the CTA that owns the tile's first unit (each CTA owns the first unit of at most one cut tile).
struct Sched {
    int a, b;          // index range of this CTA
    int L;             // indices per row of the index space
    int mt0, mts;      // token tile of row i = mt0 + i * mts
    int ub;            // unit of index 0 within a row
};

__host__ __device__ __forceinline__ Sched make_sched(const GemmParams& p, int cta) {
    const int Umt = p.n_tiles * p.G;
    Sched s;
    if (p.whole_rows) {                      // m_tiles > P: whole token tiles, several passes
        const int passes = (p.m_tiles - cta + p.P - 1) / p.P;
        s.a = 0; s.b = passes * Umt; s.L = Umt; s.mt0 = cta; s.mts = p.P; s.ub = 0;
    } else if (cta < p.R * p.Pn) {
        const int j = cta % p.Pn;
        s.a = (int)(((long long)j * p.Ureg) / p.Pn);
        s.b = (int)(((long long)(j + 1) * p.Ureg) / p.Pn);
        s.L = 0x40000000; s.mt0 = cta / p.Pn; s.mts = 0; s.ub = 0;
    } else {
        const int sp = cta - p.R * p.Pn, nsp = p.P - p.R * p.Pn;
        const long long tot = (long long)p.R * (Umt - p.Ureg);
        s.a = (int)((sp * tot) / nsp);
        s.b = (int)(((sp + 1) * tot) / nsp);
        s.L = Umt - p.Ureg; s.mt0 = 0; s.mts = 1; s.ub = p.Ureg;
    }
    return s;
}

// f(mt, nt, g0, g1) for every maximal run of k-groups [g0, g1) of one tile in this CTA's range, in order
template <typename F>
__host__ __device__ __forceinline__ void walk_segments(const Sched& s, int G, F&& f) {
    for (int idx = s.a; idx < s.b;) {
        const int row = idx / s.L;
        const int rend = (long long)(row + 1) * s.L < (long long)s.b ? (row + 1) * s.L : s.b;
        const int mt = s.mt0 + row * s.mts;
        for (int u = s.ub + (idx - row * s.L), ue = u + (rend - idx); u < ue;) {
            const int nt = u / G, g0 = u - nt * G;
            const int g1 = g0 + (ue - u) < G ? g0 + (ue - u) : G;
            f(mt, nt, g0, g1);
            u += g1 - g0;
        }
        idx = rend;
    }
}

// slot of the fp32 partial sums of tile (mt, nt): the CTA that owns the tile's first unit
__host__ __device__ __forceinline__ int tile_slot(const GemmParams& p, int mt, int nt) {
    const int u0 = nt * p.G;
    if (u0 < p.Ureg) return mt * p.Pn + unit_owner(u0, p.Ureg, p.Pn);
    const int Ls = p.n_tiles * p.G - p.Ureg;
    const long long tot = (long long)p.R * Ls, i0 = (long long)mt * Ls + (u0 - p.Ureg);
    const int nsp = p.P - p.R * p.Pn;
    return p.R * p.Pn + (int)(((i0 + 1) * nsp - 1) / tot);
}

// host: fill the decomposition fields of p (n_tiles, m_tiles, G set) for at most max_ctas CTAs
// what summing a cut tile costs a CTA, in steps of that tile configuration (measured from pipeline traces)
template <int M_TILE> constexpr int kFixSteps = M_TILE >= 128 ? 16 : M_TILE >= 64 ? 6 : M_TILE >= 32 ? 4 : 3;

// Returns the plan's estimated length in steps (gp = k-groups per step of the tile configuration; fix_steps = what summing
// a cut tile costs, in steps: the tail of the kernel that nothing overlaps).
static int plan_ctas(GemmParams& p, int max_ctas, int gp = 1, int fix_steps = 0) {
    const long long Umt = (long long)p.n_tiles * p.G;
    auto steps_of = [&](long long groups) { return (int)((groups + gp - 1) / gp); };
    if (p.m_tiles > max_ctas) {
        p.whole_rows = 1; p.P = max_ctas; p.R = max_ctas; p.Pn = 1; p.Ureg = (int)Umt;
        return ceil_div(p.m_tiles, p.P) * p.n_tiles * steps_of(p.G);
    }
    p.whole_rows = 0;
    p.R = p.m_tiles;
    p.Pn = max_ctas / p.R;
    if (p.Pn > Umt) p.Pn = (int)Umt;                                      // at least one unit per CTA
    // a cut tile collects at most kMaxParked parked runs: ranges of >= G/32 units keep a tile within 34 CTAs (only
    // problems far smaller than the machine are affected)
    const int min_units = (p.G + 31) / 32;
    if (p.Pn > Umt / min_units) {
        p.Pn = (int)(Umt / min_units) > 0 ? (int)(Umt / min_units) : 1;
        p.P = p.R * p.Pn; p.Ureg = (int)Umt;
        return steps_of(ceil_div((int)Umt, p.Pn)) + fix_steps;
    }
    // stream-K over all CTAs: the regular columns plus the spare CTAs that do not fill another column
    int spare = p.Pn < Umt ? max_ctas - p.R * p.Pn : 0;
    int P_sk = p.R * p.Pn + spare;
    int Ureg = spare ? (int)((Umt * p.R * p.Pn + P_sk / 2) / P_sk) : (int)Umt;
    // every regular CTA and every spare CTA must own at least one unit
    if (spare && (Ureg < p.Pn || (long long)p.R * (Umt - Ureg) < spare)) { P_sk = p.R * p.Pn; Ureg = (int)Umt; spare = 0; }
    // A spare CTA walks the tail units of ceil(R / spare) token tiles, and every one of those runs is cut at its head (it
    // shares a tile with the last regular CTA of that row): extra parked hand-offs that make the spare CTAs the stragglers
    // of a short kernel (4096x4096, M=2048: 64 us against a median CTA of 55).  Use them only when the steps they take off
    // everybody else outweigh that.
    static const bool spare_model = [] { const char* e = getenv("FLEXQ_SPARE_MODEL"); return !(e && e[0] == '0'); }();
    if (spare && spare_model) {
        const long long steps_with = steps_of((Umt * p.R + P_sk - 1) / P_sk), steps_without = steps_of((Umt + p.Pn - 1) / p.Pn);
        const long long extra_cuts = (p.R + spare - 1) / spare;                  // row-head runs per spare CTA
        if (steps_with + extra_cuts * ((fix_steps + 2) / 3) >= steps_without) { P_sk = p.R * p.Pn; Ureg = (int)Umt; spare = 0; }
    }
    const bool sk_cuts = (Umt * p.R) % P_sk != 0 || (Umt * p.R / P_sk) % p.G != 0;
    const int cost_sk = steps_of((Umt * p.R + P_sk - 1) / P_sk) + (sk_cuts ? fix_steps : 0);
    // Fewer n-tiles than columns: a whole number C of columns per n-tile cuts every tile into exactly C runs, so every
    // CTA walks one run of one tile -- one cut-tile sum per CTA (none when C = 1) instead of two (a range that straddles
    // a tile boundary ends one tile and starts another).  Taken when its estimated length, with fewer CTAs at work, is
    // not longer; memory-bound decode tiles (gp > 1) additionally keep >= 55 % of the SMs streaming (measured: 86 and 96
    // whole-tile CTAs beat 148 stream-K CTAs by 6 and 11 % on the 7B gate_up / qkv layers at M <= 32).
    static const bool allow_aligned = [] { const char* e = getenv("FLEXQ_ALIGN"); return !(e && e[0] == '0'); }();
    if (allow_aligned && p.Pn >= p.n_tiles && p.Pn < Umt) {
        const int aligned = p.Pn / p.n_tiles * p.n_tiles, C = aligned / p.n_tiles;
        const int cost_al = steps_of(ceil_div(p.G, C)) + (C > 1 ? (fix_steps + 1) / 2 : 0);
        static const int min_sm_pct = [] { const char* e = getenv("FLEXQ_ALIGN_SM_PCT"); return e ? atoi(e) : 55; }();
        const bool enough_sms = gp == 1 || (long long)aligned * p.R * 100 >= (long long)max_ctas * min_sm_pct;
        if (cost_al <= cost_sk && enough_sms) {
            p.Pn = aligned; p.P = p.R * p.Pn; p.Ureg = (int)Umt;
            return cost_al;
        }
    }
    p.P = P_sk; p.Ureg = Ureg;
    return cost_sk;
}

// debug / test: the segments (mt, nt, g0, g1, slot) CTA `cta` walks for a problem of m_tiles x n_tiles x G units on at
// most max_ctas CTAs; returns the number of segments (written up to cap), *n_ctas = CTAs launched
int debug_schedule(int m_tiles, int n_tiles, int G, int max_ctas, int cta, int* out, int cap, int* n_ctas) {
    GemmParams p{};
    p.m_tiles = m_tiles; p.n_tiles = n_tiles; p.G = G;
    plan_ctas(p, max_ctas, 1, kFixSteps<192>);          // the plan of the prefill tiles (1 group per step)
    if (n_ctas) *n_ctas = p.P;
    if (cta < 0 || cta >= p.P) return 0;
    int n = 0;
    walk_segments(make_sched(p, cta), G, [&](int mt, int nt, int g0, int g1) {
        if (n < cap) {
            out[5 * n] = mt; out[5 * n + 1] = nt; out[5 * n + 2] = g0; out[5 * n + 3] = g1;
            out[5 * n + 4] = (g0 == 0 && g1 == G) ? -1 : tile_slot(p, mt, nt);
        }
        n++;
    });
    return n;
}

// int32 group sum 4S (|4S| <= 2^21) -> the float 12582912 + 4S, bit-wise: the low 23 bits of 4S with bit 22
// flipped are 4S + 2^22, i.e. the mantissa of 2^23 + 2^22 + 4S.  One LOP3 on the ALU pipe; an integer add with an
// immediate (VIADD) would share the FMA pipe with the two FFMA2 of every element pair.
template <bool REARMED, bool LOP>
__device__ __forceinline__ float magic_f32(uint32_t s) {
    if (REARMED) return __uint_as_float(s);
    if (!LOP) return __uint_as_float(s + 0x4B400000u);
    uint32_t r;
    asm("lop3.b32 %0, %1, 0x007fffff, 0x4b400000, 0x6a;" : "=r"(r) : "r"(s));   // (s & 0x7fffff) ^ 0x4b400000
    return __uint_as_float(r);
}

template <int M_TILE, int GP, bool DUMP, bool TRACE>
__global__ void __launch_bounds__(Cfg<M_TILE, GP>::THREADS, 1)
w6ax_gemm_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_sx,
                 const __grid_constant__ CUtensorMap tmap_sw, const __grid_constant__ CUtensorMap tmap_w,
                 const __grid_constant__ CUtensorMap tmap_d, const GemmParams p) {
    using C = Cfg<M_TILE, GP>;
    // how tiles cut by a range boundary are summed (see the epilogue): parked partial tiles for the 192-token tile
    // (contributors arrive far apart: measured 2-11 % faster at M >= 512), shared-slot reductions for the smaller
    // tiles, whose contributors finish together (hand-off 5-65 % slower there: the completing CTA waits for the rest)
    constexpr bool kHandoff = (FLEXQ_FIXUP_HANDOFF != 0) && M_TILE >= kHandoffMinTile;
    // Decode tiles under the aligned plan run as clusters: the C runs of a weight tile are C consecutive CTAs, which finish
    // together; instead of meeting in global memory (reductions, a fence, an atomic and a load back: 2400-3800 cycles
    // of a 6-12 us kernel) the partial tiles are written into the first CTA's shared memory (the weight ring, idle by
    // then) between two cluster barriers and summed there in rank order.
    // Measured (profiles/r2_experiments/sweep_b20_*): 1.5-4.5 % at M <= 16 (most of what looks like cut-tile cost in a trace
    // is skew between the contributors, which no protocol removes); the 32- and 64-token tiles move 16-32 KB per CTA through
    // distributed shared memory and come out 3-9 % slower, so they keep the reductions.
    constexpr bool kClusterOk = (FLEXQ_CLUSTER != 0) && M_TILE <= kClusterMaxTile && !DUMP;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int P = gridDim.x;
    if constexpr (TRACE) {     // per-CTA wall-clock window after the step stamps: [trace_units*16 + 2*cta + {0,1}]
        if (threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            p.trace[(p.trace_units & 0xFFFF) * 16 + 2 * blockIdx.x] = (long long)t;
        }
    }
    const int G = p.G;
    const Sched sch = make_sched(p, blockIdx.x);
    const int Umt = p.n_tiles * G;

    // full/empty rings.  One tcgen05.commit per step arrives on bar_done(step % NDONE); it frees the
    // activation stage and the TMEM weight stage of that step and publishes its accumulators.
    const uint32_t bar0 = smem_base + C::OFF_BAR;
    auto bar_w_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_w_empty = [&](int s) { return bar0 + 8u * (C::NW + s); };
    auto bar_s_full = [&](int s) { return bar0 + 8u * (2 * C::NW + s); };
    auto bar_s_empty = [&](int s) { return bar0 + 8u * (2 * C::NW + C::NS + s); };
    auto bar_a_full = [&](int s) { return bar0 + 8u * (2 * C::NW + 2 * C::NS + s); };
    auto bar_x_full = [&](int s) { return bar0 + 8u * (2 * C::NW + 2 * C::NS + C::NAT + s); };
    auto bar_acc_empty = [&](int b) { return bar0 + 8u * (2 * C::NW + 2 * C::NS + C::NAT + C::NX + b); };
    auto bar_done = [&](int i) { return bar0 + 8u * (2 * C::NW + 2 * C::NS + C::NAT + C::NX + C::NAB + (i % C::NDONE)); };
    auto done_parity = [&](int i) { return (uint32_t)((i / C::NDONE) & 1); };
    uint32_t* misc = reinterpret_cast<uint32_t*>(smem + C::OFF_MISC);   // [0] tmem base, [1] finisher flag

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::NAT; s++) mbar_init(bar_a_full(s), 32 * C::EXP_WARPS);
        for (int s = 0; s < C::NX; s++) mbar_init(bar_x_full(s), 1);
        for (int s = 0; s < C::NS; s++) { mbar_init(bar_s_full(s), 1); mbar_init(bar_s_empty(s), C::EPI_THREADS); }
        for (int b = 0; b < C::NAB; b++) mbar_init(bar_acc_empty(b), C::EPI_THREADS);
        for (int i = 0; i < C::NDONE; i++) mbar_init(bar_done(i), 1);
        fence_barrier_init();
        prefetch_tensormap(&tmap_x);
        prefetch_tensormap(&tmap_sx);
        prefetch_tensormap(&tmap_sw);
    }
    const int cw = warp - C::CTRL_WARP0;        // 0 = W producer, 1 and 2 = MMA issuers, 3 = X producer
#ifndef FLEXQ_PRODUCER_SLEEP_NS
#define FLEXQ_PRODUCER_SLEEP_NS 0
#endif
    auto producer_wait = [&](const uint32_t bar, const uint32_t parity) {
        if constexpr (M_TILE >= 128 && FLEXQ_PRODUCER_SLEEP_NS > 0) mbar_wait_sleepy<FLEXQ_PRODUCER_SLEEP_NS>(bar, parity);
        else mbar_wait_parked(bar, parity);
    };
    constexpr int kEarlyW = C::NW < 3 ? C::NW : 3;
    // weight loads of steps [lo, hi) of this CTA (one thread)
    auto w_produce = [&](const int lo, const int hi) {
        // weights are read once when there is a single token tile (decode): keep them from
        // displacing activations/scales in L2; with several token tiles the same weight row is
        // re-streamed per token tile and should stay resident
        const uint64_t pol = (p.m_tiles == 1) ? l2_policy_evict_first() : l2_policy_evict_last();
        int it = 0;
        walk_segments(sch, G, [&](const int mt, const int nt, const int g0, const int g1) {
            const uint8_t* wsrc = p.w6 + ((size_t)nt * G) * kTileBytes;
            for (int g = g0; g < g1; g += GP, it++) {
                if (it < lo) continue;
                if (it >= hi) return;
                const int ng = min(GP, g1 - g);
                const int s = it % C::NW;
                producer_wait(bar_w_empty(s), ((it / C::NW) & 1) ^ 1);
                FQ_TRACE(it, 0);
                mbar_expect_tx(bar_w_full(s), ng * kTileBytes);
                if (p.m_tiles == 1) {
                    bulk_g2s_hint(smem_base + C::OFF_W + s * C::W_BYTES, wsrc + (size_t)g * kTileBytes, ng * kTileBytes, bar_w_full(s), pol);
                } else {
                    // re-streamed weights go through tensor-map TMA loads (W6 seen as [bytes/128][128]); measured:
                    // plain bulk copies never hit in L2 on the second pass, tiled loads do
                    for (int j = 0; j < ng; j++)
                        tma_load_2d_hint(smem_base + C::OFF_W + s * C::W_BYTES + j * kTileBytes, &tmap_w, 0,
                                         (nt * G + g + j) * (kTileBytes / 128), bar_w_full(s), pol);
                }
            }
        });
    };
    if (warp == C::CTRL_WARP0 && lane == 0) {
        // The weight stream starts before the rest of the prologue (TMEM allocation, the other barriers, the CTA-wide
        // sync): weights are static, the ring is empty and its barriers are this thread's own -- at decode sizes the
        // prologue is ~0.5 us of a 6-12 us kernel.
        // (three stages: each issue costs this thread ~300 cycles, and the CTA-wide sync below waits for it)
        for (int s = 0; s < C::NW; s++) { mbar_init(bar_w_full(s), 1); mbar_init(bar_w_empty(s), 32 * C::EXP_WARPS); }
        fence_barrier_init();
        w_produce(0, kEarlyW);
    }
    if (cw == 1) tmem_alloc<512>(smem_u32(&misc[0]));
    if constexpr (C::BIAS) {   // constant operand of the bias MMA, read through the async proxy
        for (int i = threadIdx.x; i < C::ONES_BYTES / 16; i += C::THREADS)
            reinterpret_cast<uint4*>(smem + C::OFF_ONES)[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
        fence_proxy_async_smem();
    }
    const uint32_t bar_recv = bar0 + 8u * (C::NBAR - 1);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (kClusterOk && C::RECV_BYTES > 0 && p.cluster > 1) {
        // The first CTA of a cluster will receive (cluster - 1) partial tiles on this barrier; it must be armed before anyone
        // sends.  One thread of an epilogue warp (idle until the first MMA retires) arms it after the CTA-wide sync and
        // publishes "armed" through a flag word holding this launch's nonce; the same thread of every other CTA polls that
        // word over distributed shared memory.  Measured alternatives: a cluster barrier split into release arrives here and
        // waits at the exchange costs ~1300 cycles of start-up (a release arrive is a cluster-scope fence, executed by every
        // thread), 800 more when the arrives are issued before the sync; relaxed arrives for the threads that publish
        // nothing lose partial tiles.
        if (threadIdx.x == 32) {
            mbar_init(bar_recv, 1);
            fence_barrier_init();
            if (blockIdx.x % (uint32_t)p.cluster == 0) {
                mbar_expect_tx(bar_recv, (uint32_t)(p.cluster - 1) * (M_TILE * kTileN * 4));
                if (FLEXQ_CLUSTER_FLAG) st_release_cluster_u32(cluster_map_shared(smem_u32(&misc[4 + kMaxParked]), blockIdx.x % (uint32_t)p.cluster), p.nonce);
            }
        }
    }
    if (kClusterOk && C::RECV_BYTES > 0 && p.cluster > 1) {
        if (!FLEXQ_CLUSTER_FLAG) cluster_arrive();
        else if (threadIdx.x == 32 && blockIdx.x % (uint32_t)p.cluster != 0) {
            const uint32_t flag = cluster_map_shared(smem_u32(&misc[4 + kMaxParked]), 0);
            uint32_t spins = 0;
            while (ld_acquire_cluster_u32(flag) != p.nonce)
                if (++spins > (1u << 22)) __trap();
        }
    }
    const uint32_t tmem_base = misc[0];
    if (threadIdx.x == 0) FQ_TRACE(0, 13);
    // Programmatic dependent launch: let the next kernel of the stream start its prologue / weight
    // prefetch as our CTAs retire.  Weights are static, so only the roles that touch data produced
    // by earlier kernels (activations, scales, output, split-K scratch) wait for them below.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // Every role walks the same sequence of steps:
    //   for each segment (tile, [g0,g1)) of this CTA's unit range, for g = g0; g < g1; g += GP
    // Register budget (64K regs): producer/MMA warpgroup 32, expanders 72, epilogue 200
    // (setmaxnreg sits inside each role branch so that ptxas allocates per role).
    if (cw == 0) {
        // ===================== TMA producer: packed weight tiles =====================
        reg_dealloc<32>();
        if (lane == 0) {
            w_produce(kEarlyW, 0x7fffffff);
        } else if (TRACE && lane == 1 && blockIdx.x == (p.trace_units >> 16)) {
            // trace builds: an otherwise idle lane watches the "MMAs of step i retired" barriers (event 8)
            const int n = min(sch.b - sch.a, p.trace_units & 0xFFFF);
            for (int i = 0; i < n && GP == 1; i++) {
                mbar_wait(bar_done(i), done_parity(i));
                FQ_TRACE(i, 8);
            }
        }
        __syncwarp();
    } else if (cw == 3) {
        // ===================== TMA producer: activation tiles + scale blocks =====================
        reg_dealloc<32>();
        if (lane == 0) {
            asm volatile("griddepcontrol.wait;" ::: "memory");     // activations / scales come from earlier kernels
            const uint64_t pol_x = l2_policy_evict_last();          // every n-tile re-reads the activations: keep them in L2
            int it = 0;
            walk_segments(sch, G, [&](const int mt, const int nt, const int g0, const int g1) {
                for (int g = g0; g < g1; g += GP, it++) {
                    {   // [ng][M_TILE][128 B] swizzle-128B tiles, one TMA per k-group; rows >= M are zero-filled
                        const int ng = min(GP, g1 - g);
                        const int s = it % C::NX;
                        if (it >= C::NX) producer_wait(bar_done(it - C::NX), done_parity(it - C::NX));
                        FQ_TRACE(it, 9);
                        mbar_expect_tx(bar_x_full(s), ng * (M_TILE * 128));
                        for (int j = 0; j < ng; j++)
                            tma_load_2d_hint(smem_base + C::OFF_X + s * C::X_BYTES + j * (M_TILE * 128), &tmap_x, (g + j) * kGroup, mt * M_TILE,
                                             bar_x_full(s), pol_x);
                    }
                    if (!DUMP) {   // sx[g..g+GP][m0..] (f32) and w_scale[g..g+GP][n0..] (f16)
                        const int s = it % C::NS;
                        const uint32_t dst = smem_base + C::OFF_S + s * C::S_BYTES;
                        producer_wait(bar_s_empty(s), ((it / C::NS) & 1) ^ 1);
                        mbar_expect_tx(bar_s_full(s), C::SX_BYTES + C::SW_BYTES);
                        tma_load_2d(dst, &tmap_sx, mt * M_TILE, g, bar_s_full(s));
                        tma_load_2d(dst + C::SX_BYTES, &tmap_sw, nt * kTileN, g, bar_s_full(s));
                    }
                }
            });
        }
        __syncwarp();
    } else if (cw == 1 || cw == 2) {
        // ===================== MMA issuers (steps alternate between the two warps) =====================
        reg_dealloc<32>();
        // The whole warp walks the loop converged and one elected lane issues: every MMA operand then derives from
        // warp-uniform values (uniform registers), where a lane-0 branch made ptxas wrap each tcgen05.mma in an
        // elect / broadcast / retry loop (~12 instructions and ~80 cycles per MMA).
        const int my_parity = __shfl_sync(0xffffffffu, cw, 0) - 1;
        const uint32_t tmem_base = __shfl_sync(0xffffffffu, misc[0], 0);
        const uint32_t smem_base_u = __shfl_sync(0xffffffffu, smem_base, 0);
        const bool leader = elect_one();
        {
            constexpr uint32_t idesc = umma_idesc_i8(kTileN, M_TILE);
            int it = 0;
            walk_segments(sch, G, [&](const int mt, const int nt, const int g0, const int g1) {
                for (int g = g0; g < g1; g += GP, it++) {
                    if ((it & 1) != my_parity) continue;
                    const int ng = min(GP, g1 - g);
                    const int ab = it % C::NAB, st = it % C::NAT, sx_ = it % C::NX;
                    // the barrier expected to complete last is waited on last (an already-complete wait
                    // still costs ~200 cycles): accumulator hand-back when tensor-bound, weights when HBM-bound
                    mbar_wait(bar_x_full(sx_), (it / C::NX) & 1);
                    // the scale block of this step: waited for here by one warp so that the epilogue warps need only the
                    // "MMAs retired" barrier (which then implies it); it lands long before the accumulators come back
                    if (kIssuerWaitsScales && !DUMP) mbar_wait(bar_s_full(it % C::NS), (it / C::NS) & 1);
                    if (GP == 1) {
                        mbar_wait(bar_a_full(st), (it / C::NAT) & 1);
                        mbar_wait(bar_acc_empty(ab), (it / C::NAB) & 1);   // handed back by the epilogue
                    } else {
                        mbar_wait(bar_acc_empty(ab), (it / C::NAB) & 1);
                        mbar_wait(bar_a_full(st), (it / C::NAT) & 1);
                    }
                    if (leader) FQ_TRACE(it, 4);
                    tc_fence_after();
                    // descriptor of the stage base once; per MMA only the 16-byte-unit address field advances
                    const uint64_t b_desc0 = umma_desc_sw128(smem_base_u + C::OFF_X + sx_ * C::X_BYTES);
                    const uint32_t d_tmem = tmem_base + ab * C::ACC_COLS;
                    const uint32_t a_tmem = tmem_base + C::A_COL0 + st * C::A_COLS;
                    if (leader) {
#pragma unroll
                        for (int j = 0; j < GP; j++) {
                            if (j >= ng) break;
#ifdef FLEXQ_EXP_SKIPBIASMMA
                            if (false)                 // experiment: cost of the fifth MMA (results invalid)
#endif
                            if constexpr (C::BIAS) {   // accumulator := 32 * 255 * 255 (unsigned x unsigned, constant operands)
                                const uint64_t ones = umma_desc_nosw(smem_base_u + C::OFF_ONES, 128, 256);
                                umma_i8(d_tmem + j * M_TILE, ones, ones, umma_idesc_u8(kTileN, M_TILE), 0u);
                            }
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                umma_i8_ts(d_tmem + j * M_TILE, a_tmem + j * 32 + 8 * k,
                                           b_desc0 + (uint64_t)((j * (M_TILE * 128) + 32 * k) >> 4), idesc, (kRearm || C::BIAS || k > 0) ? 1u : 0u);
                        }
                        umma_commit(bar0 + 8u * (2 * C::NW + 2 * C::NS + C::NAT + C::NX + C::NAB + (it % C::NDONE)));
                        FQ_TRACE(it, 5);
                    }
                    __syncwarp();
                }
            });
        }
        __syncwarp();
    } else if (warp >= C::EXP_WARP0 && warp < C::EXP_WARP0 + C::EXP_WARPS) {
        // ===================== weight expanders: smem (packed) -> registers -> TMEM (int8) =====================
        reg_dealloc<C::EXP_REGS>();
        const int r = (threadIdx.x - 32 * C::EXP_WARP0) & 127;   // weight row within the tile == TMEM lane
        const int xh = (threadIdx.x - 32 * C::EXP_WARP0) >> 7;   // which of the EXP_WARPS / 4 warps of that row's quadrant
        // warp-uniform TMEM address (see the epilogue); the thread part of the shared-memory source address is kept
        // opaque in one register and the packed-weight ring position is carried (NW is not a power of two)
        const uint32_t a_lane = __shfl_sync(0xffffffffu, tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + C::A_COL0, 0);
        uint32_t w_thread = smem_base + C::OFF_W + 48u * (uint32_t)r;
        asm volatile("" : "+r"(w_thread));
        int it = 0, sw = 0;
        uint32_t w_par = 0;
        int ss_e = 0;
        uint32_t s_par_e = 0;
        walk_segments(sch, G, [&](const int mt, const int nt, const int g0, const int g1) {
            for (int g = g0; g < g1; g += GP, it++) {
                const int ng = min(GP, g1 - g);
                const int st = it % C::NAT;
                mbar_wait(bar0 + 8u * sw, w_par);        // bar_w_full(sw)
                if (r == 0 && xh == 0) FQ_TRACE(it, 1);
                if (it >= C::NAT) {                      // MMAs that read this TMEM stage have retired
                    mbar_wait(bar_done(it - C::NAT), done_parity(it - C::NAT));
                    tc_fence_after();
                }
                if (r == 0 && xh == 0) FQ_TRACE(it, 2);
                const uint32_t wp = w_thread + (uint32_t)sw * C::W_BYTES;
                for (int j = xh; j < ng; j += C::EXP_WARPS / 4) {
                    uint32_t out[32];
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const uint32_t src = wp + (uint32_t)(j * kTileBytes + 48 * 128 * q);
                        const uint4 i0 = lds_u4(src), i1 = lds_u4(src + 16), i2 = lds_u4(src + 32);
                        const uint32_t w[12] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w, i2.x, i2.y, i2.z, i2.w};
#pragma unroll
                        for (int s = 0; s < 4; s++) w6_expand16(w[3 * s], w[3 * s + 1], w[3 * s + 2], out + 16 * q + 4 * s);
                    }
                    tmem_st32(a_lane + st * C::A_COLS + j * 32, out);   // this row's 128 int8 (value 4*w)
                }
                mbar_arrive(bar0 + 8u * (C::NW + sw));   // bar_w_empty(sw): packed tiles consumed
                if constexpr (C::XCONST && !DUMP) {
                    // this row's scale constants for the epilogue (see there): after the expansion, so that waiting for the
                    // scale block (issued with the step's activations) costs nothing the MMA would not wait for anyway; the
                    // a_full arrive below publishes them (release) along the chain issuer -> commit -> epilogue
                    mbar_wait(bar_s_full(ss_e), s_par_e);
                    const uint32_t sblk_e = smem_base + C::OFF_S + (uint32_t)ss_e * C::S_BYTES;
                    const float swh = __half2float(__ushort_as_half(lds_u16(sblk_e + C::SX_BYTES + (uint32_t)r * 2u)));
                    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(sblk_e + C::SX_BYTES + C::SW_BYTES + (uint32_t)r * 8u),
                                 "f"(swh * 0x1p100f), "f"(swh * (-(float)kBiasB * 0x1p-49f)) : "memory");
                    if (++ss_e == C::NS) { ss_e = 0; s_par_e ^= 1u; }
                }
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(bar_a_full(st));
                if (r == 0 && xh == 0) FQ_TRACE(it, 3);
                if (++sw == C::NW) { sw = 0; w_par ^= 1u; }
            }
        });
    } else {
        // ===================== epilogue =====================
        // The int32 group sum 4*S read back from TMEM is turned into the float (kMagicF + 4*S) bit-wise (magic_f32:
        // LOP3 or integer add; with FLEXQ_REARM the accumulators are pre-biased in TMEM instead).  One FMA with the
        // per-row weight scale removes the bias exactly (kMagicF*sw is exact in fp32 for an fp16-valued sw) and a
        // second FMA applies the per-token scale and accumulates.
        reg_alloc<C::EPI_REGS>();
        constexpr int CPT = C::CPT, CH = C::CH;
        constexpr uint32_t kMagicI = 0x4B400000u;
        constexpr float kMagicF = 12582912.f;
        constexpr float kOutScale = C::BIAS ? 0x1p47f : 1.f;   // see the scale constants below
        const int e = threadIdx.x - 32 * C::EPI_WARP0;
        const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
        const int wg_id = e >> 7;                        // which CPT-column slice of the token tile
        const int r = quad * 32 + lane;
        const int col0 = wg_id * CPT;
        // fragment layout: rows fr0 + 8k (k = 0..3), columns col0 + 8i + fc0 + {0,1} (i = 0..7)
        const int fr0 = quad * 32 + (lane >> 2), fc0 = 2 * (lane & 3);
        // the same for every lane of the warp: taken from lane 0 so that it lives in a uniform register (tcgen05.ld takes
        // a uniform address; otherwise it is rebuilt from the thread index and moved with R2UR every step)
        const uint32_t t_lane = __shfl_sync(0xffffffffu, tmem_base + ((uint32_t)(quad * 32) << 16) + col0, 0);
        // arm every accumulator buffer once
        if constexpr (kRearm) {
#pragma unroll
            for (int b = 0; b < C::NAB * GP; b++) {
#pragma unroll
                for (int c = 0; c < CPT; c += (CPT < 16 ? 8 : 16)) {
                    if constexpr (CPT < 16) tmem_st8_same(t_lane + b * M_TILE + c, kMagicI);
                    else tmem_st16_same(t_lane + b * M_TILE + c, kMagicI);
                }
            }
            tmem_wait_st();
            tc_fence_before();
        }
#pragma unroll
        for (int b = 0; b < C::NAB; b++) mbar_arrive(bar_acc_empty(b));

        asm volatile("griddepcontrol.wait;" ::: "memory");         // D and the split-K scratch belong to earlier kernels until now
        float2 acc[CPT / 2];
        int it = 0;
        int ss = 0, dn = 0, ab = 0;
        uint32_t s_par = 0, dn_par = 0;
        uint32_t s_base = smem_base + C::OFF_S;
        asm volatile("" : "+r"(s_base));              // opaque: keep it in a register instead of re-deriving it every step
        // fragment layout: this thread's weight-scale and token-scale addresses within stage 0 of the scale ring
        // thread part of the scale addresses, kept opaque so that it stays in two registers instead of being rebuilt from
        // the thread index every step; the stage part (sblk) is uniform
        uint32_t sw_off = C::SX_BYTES + (uint32_t)fr0 * 2u, sx_off = (uint32_t)(col0 + fc0) * 4u;
        uint32_t sc_off = C::SX_BYTES + C::SW_BYTES + (uint32_t)fr0 * 8u;
        asm volatile("" : "+r"(sw_off), "+r"(sx_off), "+r"(sc_off));
        const uint32_t bar_s_full0 = s_base + (C::OFF_BAR - C::OFF_S) + 8u * (2 * C::NW);
        const uint32_t bar_s_empty0 = bar_s_full0 + 8u * C::NS;
        const uint32_t bar_acc_empty0 = bar_s_full0 + 8u * (2 * C::NS + C::NAT + C::NX);
        const uint32_t bar_done0 = bar_acc_empty0 + 8u * C::NAB;
        // Cross-step prefetch (1 group per step, even chunk count): all epilogue warps reach the end of a step together,
        // so the barrier wait + first TMEM load of the next step would be a bubble nothing fills.  Instead the last
        // chunk of a step probes the next step's "MMAs retired" barrier (normally long complete) and issues that
        // step's first load before doing its own math; `pre` says the load is already in flight.
        constexpr bool PREFETCH = (FLEXQ_EPI_PREFETCH != 0) && GP == 1 && ((CPT / CH) % 2 == 0) && !kRearm;
        bool pre = false;
        const int n_steps = sch.b - sch.a;
        // Finished tile -> D.  Fragment layout with staging (TSTORE): every thread converts its values to fp16 and stores them
        // to the warpgroup's staging tile ([token][64 rows] boxes, swizzle-128B); one thread per warpgroup then issues two
        // TMA stores (rows / tokens beyond N / M are clipped by the tensor map).  2-byte global stores straight from the fragment cost the LSU a
        // sector per 16 bytes (measured: 8 % of the epilogue's stall samples at K = 8192, proportionally more at smaller K).
        // SiLU(gate) * up on the fp16-rounded GEMM outputs, arithmetic of silu_mul_quant_kernel (act_quant.cu; reference
        // activation_kernels.cu:129-144): fp32, __expf, fast division; the caller rounds the product to fp16
        auto silu_mul = [&](const float g_acc, const float u_acc) {
            const float g = __half2float(__float2half_rn(kOutScale * g_acc)), u = __half2float(__float2half_rn(kOutScale * u_acc));
            return __fdividef(g, 1.0f + __expf(-g)) * u;
        };
        auto store_tile = [&](const int mt, const int nt, const int n, const bool n_ok, const int mbase) {
            if (p.silu) {
                // weight rows come in blocks of 8 gate rows followed by their 8 up rows, so a thread of the fragment layout
                // holds gate (k = 0, 2) and up (k = 1, 3) of the same output columns, a thread of the row layout finds its
                // partner 8 lanes away; the tile yields 64 output columns: nt * 64 + 16 * (row / 32) + 8 * kk + row % 8
                if constexpr (C::TSTORE && !DUMP) {
                    if ((e & 127) == 0) bulk_wait_read_all();
                    named_bar_sync(2 + wg_id, 128);
                    const uint32_t rowbase = smem_base + C::OFF_STAGE + (uint32_t)wg_id * 16384u + (uint32_t)fc0 * 128u + 2u * (uint32_t)(lane >> 2);
#pragma unroll
                    for (int kk = 0; kk < 2; kk++) {
                        const uint32_t ch = (uint32_t)(2 * quad + kk);       // 16-byte column of output columns 16 * quad + 8 * kk + 0..7
                        const uint32_t off0 = ((ch ^ (uint32_t)fc0) & 7u) << 4, off1 = 128u + (((ch ^ (uint32_t)(fc0 + 1)) & 7u) << 4);
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const float2 g2 = acc[(2 * kk) * 8 + i], u2 = acc[(2 * kk + 1) * 8 + i];
                            sts_u16(rowbase + (uint32_t)i * 1024u + off0, __half_as_ushort(__float2half_rn(silu_mul(g2.x, u2.x))));
                            sts_u16(rowbase + (uint32_t)i * 1024u + off1, __half_as_ushort(__float2half_rn(silu_mul(g2.y, u2.y))));
                        }
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(2 + wg_id, 128);
                    if ((e & 127) == 0) {
                        tma_store_2d(&tmap_d, smem_base + C::OFF_STAGE + (uint32_t)wg_id * 16384u, nt * 64, mt * M_TILE + col0);
                        bulk_commit_group();
                    }
                } else if constexpr (C::FRAG) {
#pragma unroll
                    for (int kk = 0; kk < 2; kk++) {
                        const int hn = nt * 64 + 16 * quad + 8 * kk + (lane >> 2);
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const int m = mbase + 8 * j + fc0;
                            const float2 g2 = acc[(2 * kk) * 8 + j], u2 = acc[(2 * kk + 1) * 8 + j];
                            unsigned short* dst = reinterpret_cast<unsigned short*>(p.D) + (size_t)m * p.ldd + hn;
                            if (hn < p.ldd && m < p.M) __stcs(dst, __half_as_ushort(__float2half_rn(silu_mul(g2.x, u2.x))));
                            if (hn < p.ldd && m + 1 < p.M) __stcs(dst + p.ldd, __half_as_ushort(__float2half_rn(silu_mul(g2.y, u2.y))));
                        }
                    }
                } else {
                    const bool gate_lane = ((lane >> 3) & 1) == 0;
                    const int hn = nt * 64 + 16 * quad + 8 * (lane >> 4) + (lane & 7);
#pragma unroll
                    for (int j = 0; j < CPT; j++) {
                        const int m = mbase + j;
                        const float a = (j & 1) ? acc[j / 2].y : acc[j / 2].x;
                        const float other = __shfl_xor_sync(0xffffffffu, a, 8);
                        if (gate_lane && hn < p.ldd && m < p.M)
                            __stcs(reinterpret_cast<unsigned short*>(p.D) + (size_t)m * p.ldd + hn, __half_as_ushort(__float2half_rn(silu_mul(a, other))));
                    }
                }
                return;
            }
            if constexpr (C::TSTORE && !DUMP) {
                if ((e & 127) == 0) bulk_wait_read_all();            // the previous tile's stores have read the staging tile
                named_bar_sync(2 + wg_id, 128);
                // every thread stores its own halves: 2 conversions + 2 st.shared.u16 per value pair (an exchange with the
                // neighbouring row's lane for one packed st.shared.u32 -- shuffle + 4 selects -- measured no faster); the 8 lanes
                // of a row octet fill one 16-byte column, the 4 token rows of a warp land in 4 different columns (swizzle):
                // conflict free
                {
                    const uint32_t rowbase = smem_base + C::OFF_STAGE + (uint32_t)wg_id * 16384u + (uint32_t)(quad >> 1) * 8192u +
                                             (uint32_t)fc0 * 128u + 2u * (uint32_t)(lane >> 2);
                    const uint32_t c0 = (uint32_t)(4 * (quad & 1));
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t off0 = (((c0 | (uint32_t)k) ^ (uint32_t)fc0) & 7u) << 4, off1 = 128u + ((((c0 | (uint32_t)k) ^ (uint32_t)(fc0 + 1)) & 7u) << 4);
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const float2 a = __fmul2_rn(acc[k * 8 + i], make_float2(kOutScale, kOutScale));   // exact: a power of two
                            sts_u16(rowbase + (uint32_t)i * 1024u + off0, __half_as_ushort(__float2half_rn(a.x)));
                            sts_u16(rowbase + (uint32_t)i * 1024u + off1, __half_as_ushort(__float2half_rn(a.y)));
                        }
                    }
                }
                fence_proxy_async_smem();
                named_bar_sync(2 + wg_id, 128);
                if ((e & 127) == 0) {
                    const uint32_t st = smem_base + C::OFF_STAGE + (uint32_t)wg_id * 16384u;
                    tma_store_2d(&tmap_d, st, nt * kTileN, mt * M_TILE + col0);
                    if (nt * kTileN + 64 < p.N) tma_store_2d(&tmap_d, st + 8192u, nt * kTileN + 64, mt * M_TILE + col0);
                    bulk_commit_group();
                }
            } else if constexpr (C::FRAG) {
                // acc[k * 8 + j].{x,y}: row fr0 + 8k, column col0 + 8j + fc0 + {0,1}
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int n2 = nt * kTileN + fr0 + 8 * k;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int m = mbase + 8 * j + fc0;
                        unsigned short* dst = reinterpret_cast<unsigned short*>(p.D) + (size_t)m * p.ldd + n2;
                        if (n2 < p.N && m < p.M) __stcs(dst, __half_as_ushort(__float2half_rn(kOutScale * acc[k * 8 + j].x)));
                        if (n2 < p.N && m + 1 < p.M) __stcs(dst + p.ldd, __half_as_ushort(__float2half_rn(kOutScale * acc[k * 8 + j].y)));
                    }
                }
            } else {
                if (n_ok) {
#pragma unroll
                    for (int j = 0; j < CPT; j++) {
                        const int m = mbase + j;
                        const float a = kOutScale * ((j & 1) ? acc[j / 2].y : acc[j / 2].x);
                        if (m < p.M) __stcs(reinterpret_cast<unsigned short*>(p.D) + (size_t)m * p.ldd + n, __half_as_ushort(__float2half_rn(a)));   // streaming: do not displace X/W in L2
                    }
                }
            }
        };
        int c_mt = 0, c_nt = 0;
        walk_segments(sch, G, [&](const int mt, const int nt, const int g0, const int g1) {
            const int n = nt * kTileN + r;
            const int mbase = mt * M_TILE + col0;
            const bool n_ok = n < p.N;
#pragma unroll
            for (int j = 0; j < CPT / 2; j++) acc[j] = make_float2(0.f, 0.f);
            for (int g = g0; g < g1; g += GP, it++) {
                const int ng = min(GP, g1 - g);
                // ring positions are carried, not derived from `it`: no div/mod or address rebuild per step
                const uint32_t sblk = s_base + (uint32_t)ss * C::S_BYTES;
                if (!DUMP && !kIssuerWaitsScales) mbar_wait(bar_s_full0 + 8u * ss, s_par);
                if (!PREFETCH || !pre) {
                    mbar_wait(bar_done0 + 8u * dn, dn_par);
                    tc_fence_after();
                }
                if (e == 0) FQ_TRACE(it, 6);
                // TMEM drain, software pipelined: the load of chunk c+1 is in flight while chunk c is
                // dequantised, and the buffer goes back to the MMA issuers as soon as its last chunk has
                // been read and re-armed -- before that chunk's math.
                constexpr int NCH = CPT / CH;
                uint32_t v[2][CH];
                auto ld_chunk_of = [&](int buf, int j, int c, uint32_t* dst) {
#ifdef FLEXQ_EXP_NOLD
                    for (int q = 0; q < CH; q++) dst[q] = (uint32_t)(j + c + q + it);   // experiment: no TMEM traffic
                    return;
#endif
                    if constexpr (C::FRAG) {     // chunk c = (column half c >> 1, lane half c & 1): 16 lanes x 32 columns
                        tmem_ld_16x256b_x4(t_lane + ((uint32_t)(16 * (c & 1)) << 16) + buf * C::ACC_COLS + 32 * (c >> 1), dst);
                        return;
                    }
                    const uint32_t ta = t_lane + buf * C::ACC_COLS + j * M_TILE + c * CH;
                    if constexpr (CH == 8) tmem_ld8(ta, dst);
                    else if constexpr (CH == 16) tmem_ld16(ta, dst);
                    else tmem_ld32(ta, dst);
                };
                auto ld_chunk = [&](int j, int c, uint32_t* dst) { ld_chunk_of(ab, j, c, dst); };
                auto rearm_chunk = [&](int j, int c) {
                    const uint32_t ta = t_lane + ab * C::ACC_COLS + j * M_TILE + c * CH;
                    if constexpr (CH == 8) tmem_st8_same(ta, kMagicI);
                    else {
#pragma unroll
                        for (int cc = 0; cc < CH; cc += 16) tmem_st16_same(ta + cc, kMagicI);
                    }
                };
#ifdef FLEXQ_EXP_NOEPI
                // experiment: upper bound of the producer/expander/MMA side (no TMEM drain, results invalid)
                tc_fence_before();
                mbar_arrive(bar_acc_empty0 + 8u * ab);
                if (false)
#endif
                if (!PREFETCH || !pre) ld_chunk(0, 0, v[0]);
                pre = false;
#pragma unroll
                for (int j = 0; j < GP; j++) {           // unrolled: register double-buffer indices stay static
                    if (j >= ng) break;
#ifdef FLEXQ_EXP_NOEPI
                    break;
#endif
                    float2 sw2 = make_float2(0.f, 0.f), bias2 = make_float2(0.f, 0.f);
                    const uint32_t sxs = sblk + (uint32_t)(j * M_TILE + col0) * 4u;
                    float2 fc[4];                        // fragment layout: scale constants (c1, c2) of the thread's four rows
                    float2 sxp[4];                       // token scales of the current column half
                    if constexpr (C::FRAG && !DUMP) {
                        // rows beyond N: the scale block is zero-filled by TMA and the row is never stored -- no predicate
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            if constexpr (C::XCONST) {
                                fc[k] = lds_f2(sblk + sc_off + 64u * k);     // (sw * 2^100, -B * sw * 2^-49) from the row's expander thread
                            } else {
                                const float swh = __half2float(__ushort_as_half(lds_u16(sblk + sw_off + 16u * k)));
                                fc[k] = make_float2(swh * 0x1p100f, swh * (-(float)kBiasB * 0x1p-49f));   // immediates: nothing to keep in registers
                            }
                        }
                    }
                    if (!DUMP && !C::FRAG) {
                        const float swh = n_ok ? __half2float(__ushort_as_half(lds_u16(sblk + C::SX_BYTES + (uint32_t)(j * kTileN + r) * 2u))) : 0.f;
                        if constexpr (C::BIAS) {
                            // accumulator bits = the subnormal (B + 4S) * 2^-149:  t = fma(v, sw * 2^100, -B * sw * 2^-49) = 4S * sw * 2^-49;
                            // the tile's fp32 sums therefore carry a factor 2^-47 (operands hold 4*w) that the store removes
                            const float c1 = swh * 0x1p100f;
                            const float c2 = -(float)kBiasB * (swh * 0x1p-49f);
                            sw2 = make_float2(c1, c1);
                            bias2 = make_float2(c2, c2);
                        } else {
                            const float swv = 0.25f * swh;   // operands hold 4*w
                            sw2 = make_float2(swv, swv);
                            bias2 = make_float2(-kMagicF * swv, -kMagicF * swv);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < NCH; c++) {
                        uint32_t* cur = v[(j * NCH + c) & 1];
                        uint32_t* nxt = v[(j * NCH + c + 1) & 1];
                        const bool last_of_step = (c == NCH - 1) && (j == ng - 1);
#ifndef FLEXQ_EXP_NOLD
                        tmem_wait_ld();                              // chunk (j, c) is in registers
#endif
                        if constexpr (kRearm) rearm_chunk(j, c);
                        if (!last_of_step) {
                            if (c + 1 < NCH) ld_chunk(j, c + 1, nxt);
                            else ld_chunk(j + 1, 0, nxt);
                        } else {
                            if constexpr (kRearm) tmem_wait_st();    // every chunk read (and re-armed):
                            tc_fence_before();                       // hand the buffer back before the last math
                            mbar_arrive(bar_acc_empty0 + 8u * ab);
                            if constexpr (PREFETCH) {
                                if (it + 1 < n_steps) {
                                    const int dn1 = (dn + 1 == C::NDONE) ? 0 : dn + 1;
                                    pre = mbar_try_wait(bar_done0 + 8u * dn1, dn1 == 0 ? dn_par ^ 1u : dn_par);
                                    if (pre) {
                                        tc_fence_after();
                                        ld_chunk_of((ab + 1 == C::NAB) ? 0 : ab + 1, 0, 0, nxt);
                                    }
                                }
                            }
                        }
                        if constexpr (C::FRAG) {
                            const int h = c & 1, chh = c >> 1;
                            if constexpr (DUMP) {
#pragma unroll
                                for (int q = 0; q < 16; q++) {
                                    const int n2 = nt * kTileN + fr0 + 16 * h + 8 * ((q >> 1) & 1);
                                    const int m = mt * M_TILE + col0 + 32 * chh + 8 * (q >> 2) + fc0 + (q & 1);
                                    if (n2 < p.N && m < p.M) p.S[((size_t)m * p.N + n2) * G + g] = ((int32_t)(cur[q] - kBiasB)) >> 2;
                                }
                            } else {
                                if (h == 0) {
#pragma unroll
                                    for (int i = 0; i < 4; i++) sxp[i] = lds_f2(sblk + sx_off + (uint32_t)(32 * chh + 8 * i) * 4u);
                                }
#pragma unroll
                                for (int i = 0; i < 4; i++) {
#pragma unroll
                                    for (int rr = 0; rr < 2; rr++) {
                                        const int k = 2 * h + rr;
                                        const float2 t = __ffma2_rn(make_float2(__uint_as_float(cur[4 * i + 2 * rr]), __uint_as_float(cur[4 * i + 2 * rr + 1])),
                                                                    make_float2(fc[k].x, fc[k].x), make_float2(fc[k].y, fc[k].y));
                                        acc[k * 8 + 4 * chh + i] = __ffma2_rn(t, sxp[i], acc[k * 8 + 4 * chh + i]);
                                    }
                                }
                            }
                        } else if constexpr (DUMP) {
                            if (n_ok) {
#pragma unroll
                                for (int q = 0; q < CH; q++) {
                                    const int m = mbase + c * CH + q;
                                    if (m < p.M) p.S[((size_t)m * p.N + n) * G + g + j] = ((int32_t)(C::BIAS ? cur[q] - kBiasB : kRearm ? cur[q] - kMagicI : cur[q])) >> 2;
                                }
                            }
                        } else {
#ifdef FLEXQ_EXP_NOMATH
                            acc[0].x += __uint_as_float(cur[0] ^ cur[CH - 1]);   // experiment: drain only
                            if (false)
#endif
#pragma unroll
                            for (int q = 0; q < CH; q += 4) {
#ifdef FLEXQ_EXP_NOSX
                                const float4 s4 = make_float4(sw2.x, bias2.x, sw2.x, bias2.x);      // experiment: no shared-memory scale reads
#else
                                const float4 s4 = lds_f4(sxs + (uint32_t)(c * CH + q) * 4u);
#endif
                                constexpr uint32_t kAdd = kRearm ? 0u : kMagicI;   // int32 4S -> bits of the float (kMagicF + 4S)
#ifdef FLEXQ_EXP_SCALAR
                                float2& a0 = acc[(c * CH + q) / 2];
                                float2& a1 = acc[(c * CH + q) / 2 + 1];
                                a0.x = fmaf(fmaf(__uint_as_float(cur[q + 0] + kAdd), sw2.x, bias2.x), s4.x, a0.x);
                                a0.y = fmaf(fmaf(__uint_as_float(cur[q + 1] + kAdd), sw2.x, bias2.x), s4.y, a0.y);
                                a1.x = fmaf(fmaf(__uint_as_float(cur[q + 2] + kAdd), sw2.x, bias2.x), s4.z, a1.x);
                                a1.y = fmaf(fmaf(__uint_as_float(cur[q + 3] + kAdd), sw2.x, bias2.x), s4.w, a1.y);
#else
                                constexpr int L = C::MAGIC_LOPS;
                                constexpr bool RAW = kRearm || C::BIAS;      // accumulator bits are already the float to scale
                                const float2 t0 = __ffma2_rn(make_float2(magic_f32<RAW, (L > 0)>(cur[q + 0]), magic_f32<RAW, (L > 2)>(cur[q + 1])), sw2, bias2);
                                const float2 t1 = __ffma2_rn(make_float2(magic_f32<RAW, (L > 1)>(cur[q + 2]), magic_f32<RAW, (L > 3)>(cur[q + 3])), sw2, bias2);
                                acc[(c * CH + q) / 2] = __ffma2_rn(t0, make_float2(s4.x, s4.y), acc[(c * CH + q) / 2]);
                                acc[(c * CH + q) / 2 + 1] = __ffma2_rn(t1, make_float2(s4.z, s4.w), acc[(c * CH + q) / 2 + 1]);
#endif
                            }
                        }
                    }
                }
                if (e == 0) FQ_TRACE(it, 7);
                if (!DUMP) mbar_arrive(bar_s_empty0 + 8u * ss);
                if (++ss == C::NS) { ss = 0; s_par ^= 1u; }
                if (++dn == C::NDONE) { dn = 0; dn_par ^= 1u; }
                ab = (ab + 1 == C::NAB) ? 0 : ab + 1;
            }
            if (e == 0) FQ_TRACE(it - 1, 10);
            if constexpr (!DUMP) {
                if (g0 == 0 && g1 == G) {
                    store_tile(mt, nt, n, n_ok, mbase);      // whole tile reduced by this CTA
                } else if (kClusterOk && p.cluster > 1) {
                    // the CTA's only run (aligned plan): its partial sums stay in registers for the cluster exchange below
                    c_mt = mt; c_nt = nt;
                } else {
                    if (!kHandoff || !p.handoff) {
                    // Tile cut by a range boundary, reduction variant: every contributor adds its fp32 partial tile into
                    // the slot of the CTA that owns the tile's first unit with 16-byte vector reductions (thread-linear
                    // layout, a warp covers 512 contiguous bytes) and counts its groups; the one that completes the
                    // count loads the sum, re-zeroes the slot and writes D.
                    const int slot = tile_slot(p, mt, nt);
                    int* rec = p.cnt + (size_t)slot * kRecInts;
                    float4* sl = reinterpret_cast<float4*>(p.slots + (size_t)slot * kSlotFloats) + e;
#pragma unroll
                    for (int j = 0; j < CPT / 4; j++)
                        asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(sl + j * C::EPI_THREADS), "f"(acc[2 * j].x),
                                     "f"(acc[2 * j].y), "f"(acc[2 * j + 1].x), "f"(acc[2 * j + 1].y)
                                     : "memory");
                    named_bar_sync(1, C::EPI_THREADS);              // every thread's reductions are issued ...
                    if (e == 0) {                        // ... and released (cumulatively) by one acq_rel atomic
                        int old;
                        asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], %2;" : "=r"(old) : "l"(rec), "r"(g1 - g0) : "memory");
                        misc[1] = (old + (g1 - g0) == G) ? 1u : 0u;
                    }
                    named_bar_sync(1, C::EPI_THREADS);
                    const bool last = misc[1] != 0;
                    if (e == 0) FQ_TRACE(it - 1, 14);
                    if (last) {
                        // all loads first (independent, in flight together), then zero + store
#pragma unroll
                        for (int j = 0; j < CPT / 4; j++) {
                            const float4 v = __ldcg(sl + j * C::EPI_THREADS);
                            acc[2 * j] = make_float2(v.x, v.y);
                            acc[2 * j + 1] = make_float2(v.z, v.w);
                        }
                        if (TRACE && e == 0 && acc[0].x != 12345.f) FQ_TRACE(it - 1, 15);      // depends on the first load
#pragma unroll
                        for (int j = 0; j < CPT / 4; j++) __stcg(sl + j * C::EPI_THREADS, make_float4(0.f, 0.f, 0.f, 0.f));
                        store_tile(mt, nt, n, n_ok, mbase);
                        if (e == 0) rec[0] = 0;
                    }
                    named_bar_sync(1, C::EPI_THREADS);              // flag word is reused by the next partial segment
                    } else {
                    // Tile cut by a range boundary ("hand-off"): every contributor announces its groups on the tile's
                    // record; all but the one that completes the count park their fp32 partial tile in a pool slot (plain
                    // coalesced 16-byte stores, thread-linear [value quad][thread]: every contributor runs the same
                    // configuration, so a thread meets its own elements again) and publish the slot; the completing CTA
                    // keeps its partial in registers, waits until everything announced is parked (normally long done:
                    // the other contributors met this tile at the start of their ranges), adds the parked tiles and
                    // writes D.  Against fp32 reductions into a shared slot (red.add, then load + zero by the finisher)
                    // this moves a third of the bytes through L2 and needs no zeroed scratch; with two contributors --
                    // every cut tile of a prefill-sized problem -- the sum is the same bits whoever finishes.
                    int* rec = p.cnt + (size_t)tile_slot(p, mt, nt) * kRecInts;
                    const int ngr = g1 - g0;
                    if (e == 0) {
                        // one round trip: groups arrived (low half) and arrival index (high half) in one word, and a pool
                        // slot drawn at the same time (the completing CTA's draw is simply not used)
                        int old, draw;
                        asm volatile("atom.add.relaxed.gpu.global.s32 %0, [%2], %4;\n\tatom.add.relaxed.gpu.global.s32 %1, [%3], 1;"
                                     : "=r"(old), "=r"(draw) : "l"(rec), "l"(p.cnt + kMaxCtas * kRecInts), "r"(ngr + (1 << 16)) : "memory");
                        misc[1] = ((old & 0xFFFF) + ngr == G) ? 1u : 0u;
                        misc[2] = (uint32_t)kMaxCtas + (uint32_t)draw % (uint32_t)kSlotPool;    // parking slots follow the accumulation slots
                        misc[3] = (uint32_t)(old >> 16);            // runs that arrived before this one
                    }
                    named_bar_sync(1, C::EPI_THREADS);
                    const bool last = misc[1] != 0;
                    constexpr int V4 = CPT / 4;
                    if (e == 0) FQ_TRACE(it - 1, 14);
                    if (!last) {
                        const uint32_t slot = misc[2];
                        float4* sl = reinterpret_cast<float4*>(p.slots + (size_t)slot * kSlotFloats) + e;
#pragma unroll
                        for (int j = 0; j < V4; j++)
                            __stcg(sl + j * C::EPI_THREADS, make_float4(acc[2 * j].x, acc[2 * j].y, acc[2 * j + 1].x, acc[2 * j + 1].y));
                        named_bar_sync(1, C::EPI_THREADS);          // every thread's stores are issued ...
                        if (e == 0) {                               // ... and released, with the slot id, by one thread
                            const int k = (int)misc[3];
                            if (k >= kMaxParked) __trap();          // plan_ctas bounds the contributors of a tile
                            *reinterpret_cast<volatile int*>(rec + 4 + k) = (int)slot;
                            asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(rec + 1), "r"(ngr) : "memory");
                        }
                    } else {
                        if (e == 0) {
                            int parked, spins = 0;
                            do {
                                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(parked) : "l"(rec + 1) : "memory");
                                if (++spins > (1 << 26)) __trap();
                            } while (parked != G - ngr);
                            const int np = (int)misc[3];             // everyone else arrived before the completing run
                            for (int k = 0; k < np; k++) {
                                misc[4 + k] = (uint32_t) * reinterpret_cast<volatile int*>(rec + 4 + k);
                                rec[4 + k] = 0;
                            }
                            misc[3] = (uint32_t)np;
                            rec[0] = 0; rec[1] = 0;                 // the record is this tile's alone: leave it zeroed
                        }
                        named_bar_sync(1, C::EPI_THREADS);
                        if (e == 0) FQ_TRACE(it - 1, 15);
                        const int npark = (int)misc[3];
                        constexpr int CHK = V4 > 8 ? 8 : V4;        // 16-byte loads in flight per thread and parked tile
                        for (int k = 0; k < npark; k++) {
                            const float4* src = reinterpret_cast<const float4*>(p.slots + (size_t)misc[4 + k] * kSlotFloats) + e;
#pragma unroll
                            for (int j0 = 0; j0 < V4; j0 += CHK) {
                                float4 v[CHK];
#pragma unroll
                                for (int j = 0; j < CHK; j++) v[j] = __ldcg(src + (j0 + j) * C::EPI_THREADS);
#pragma unroll
                                for (int j = 0; j < CHK; j++) {
                                    acc[2 * (j0 + j)].x += v[j].x; acc[2 * (j0 + j)].y += v[j].y;
                                    acc[2 * (j0 + j) + 1].x += v[j].z; acc[2 * (j0 + j) + 1].y += v[j].w;
                                }
                            }
                        }
                        store_tile(mt, nt, n, n_ok, mbase);
                    }
                    named_bar_sync(1, C::EPI_THREADS);              // the flag words are reused by the next cut run
                    }
                }
            }
            if (e == 0) FQ_TRACE(it - 1, 11);
                    });
        if (kClusterOk && p.cluster > 1) {
            const uint32_t crank = blockIdx.x % (uint32_t)p.cluster;
            constexpr int V4 = CPT / 4;
            constexpr uint32_t kPart = M_TILE * kTileN * 4;          // bytes of one partial tile, thread-linear [quad][thread]
            const uint32_t mine = smem_base + (C::RECV_BYTES > 0 ? C::OFF_RECV : C::OFF_W) + (uint32_t)e * 16u;
            if constexpr (C::RECV_BYTES > 0) {
                // the other CTAs send their partial tiles with asynchronous stores that count their bytes on the first CTA's
                // barrier: no cluster-wide barrier on the tail (two of them cost ~3000 cycles of a 9000-cycle kernel)
                if (FLEXQ_CLUSTER_FLAG) named_bar_sync(1, C::EPI_THREADS);      // thread 32 has seen the first CTA's "armed" flag
                else cluster_wait();
                if (crank != 0) {
                    const uint32_t dst = cluster_map_shared(mine + (crank - 1) * kPart, 0), rb = cluster_map_shared(bar_recv, 0);
#pragma unroll
                    for (int j = 0; j < V4; j++)
                        st_async_cluster_f4(dst + (uint32_t)j * C::EPI_THREADS * 16u,
                                            make_float4(acc[2 * j].x, acc[2 * j].y, acc[2 * j + 1].x, acc[2 * j + 1].y), rb);
                } else {
                    mbar_wait(bar_recv, 0);
                    // un-publish "armed": a CUDA-graph replay launches this kernel again with the same nonce, and shared memory
                    // keeps its contents between launches
                    if (FLEXQ_CLUSTER_FLAG && e == 0) misc[4 + kMaxParked] = 0u;
                }
            } else {
            cluster_sync_all();                                      // every CTA of the cluster has drained its weight ring
            if (crank != 0) {
                const uint32_t dst = cluster_map_shared(mine + (crank - 1) * kPart, 0);
#pragma unroll
                for (int j = 0; j < V4; j++)
                    st_cluster_f4(dst + (uint32_t)j * C::EPI_THREADS * 16u, make_float4(acc[2 * j].x, acc[2 * j].y, acc[2 * j + 1].x, acc[2 * j + 1].y));
            }
            cluster_sync_all();                                      // the partial tiles have landed in rank 0
            }
            if (crank == 0) {
                for (uint32_t rk = 1; rk < (uint32_t)p.cluster; rk++) {      // rank order: the same bits every run
#pragma unroll
                    for (int j = 0; j < V4; j++) {
                        const float4 v = lds_f4(mine + (rk - 1) * kPart + (uint32_t)j * C::EPI_THREADS * 16u);
                        acc[2 * j].x += v.x; acc[2 * j].y += v.y; acc[2 * j + 1].x += v.z; acc[2 * j + 1].y += v.w;
                    }
                }
                const int n = c_nt * kTileN + r;
                store_tile(c_mt, c_nt, n, n < p.N, c_mt * M_TILE + col0);
            }
            if (e == 0) FQ_TRACE(it - 1, 11);
        }
    }

    if (kClusterOk && C::RECV_BYTES > 0 && !FLEXQ_CLUSTER_FLAG && p.cluster > 1 && !(warp >= C::EPI_WARP0 && warp < C::EPI_WARP0 + 4 * C::EPI_WG)) cluster_wait();
    if (kClusterOk && C::RECV_BYTES == 0 && p.cluster > 1 && !(warp >= C::EPI_WARP0 && warp < C::EPI_WARP0 + 4 * C::EPI_WG)) {
        cluster_sync_all();      // the two barriers of the epilogue's cluster exchange: every thread of the cluster takes part
        cluster_sync_all();
    }
    if constexpr (C::TSTORE && !DUMP) {      // the last tile's TMA stores must have left shared memory before the CTA exits
        if (warp >= C::EPI_WARP0 && warp < C::EPI_WARP0 + 4 * C::EPI_WG && ((threadIdx.x - 32 * C::EPI_WARP0) & 127) == 0) {
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) FQ_TRACE(0, 12);
    if constexpr (TRACE) {
        if (threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            p.trace[(p.trace_units & 0xFFFF) * 16 + 2 * blockIdx.x + 1] = (long long)t;
        }
    }
    if (cw == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
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

// Encoded 2-D tensor maps, cached per host thread: the descriptor is a pure function of (pointer, type, shape, pitch, box,
// swizzle, L2 promotion), so a hit can never be stale, and a decode step that launches the same layers again (eagerly, outside
// a CUDA graph) skips five driver calls per GEMM.  64 entries, round robin.
struct TmapKey {
    const void* ptr; int dtype, swizzle, l2; unsigned long long d0, d1, pitch; unsigned b0, b1;
};
static CUresult encode_2d_cached(PFN_tmapEncodeTiled enc, CUtensorMap* out, CUtensorMapDataType dt, void* ptr, const cuuint64_t* dims,
                                 const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* ones, CUtensorMapSwizzle sw,
                                 CUtensorMapL2promotion l2) {
    struct Entry { TmapKey k; CUtensorMap m; bool used; };
    static thread_local Entry cache[64];
    static thread_local int next = 0;
    static const bool enabled = [] { const char* e = getenv("FLEXQ_TMAP_CACHE"); return !(e && e[0] == '0'); }();
    if (!enabled) return enc(out, dt, 2, ptr, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TmapKey k;
    memset(&k, 0, sizeof(k));            // padding bytes take part in the comparison
    k.ptr = ptr; k.dtype = (int)dt; k.swizzle = (int)sw; k.l2 = (int)l2; k.d0 = dims[0]; k.d1 = dims[1]; k.pitch = strides[0]; k.b0 = box[0]; k.b1 = box[1];
    for (int i = 0; i < 64; i++)
        if (cache[i].used && memcmp(&cache[i].k, &k, sizeof(k)) == 0) { *out = cache[i].m; return CUDA_SUCCESS; }
    const CUresult r = enc(out, dt, 2, ptr, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS) {
        Entry& e = cache[next];
        next = (next + 1) & 63;
        e.k = k; e.m = *out; e.used = true;
    }
    return r;
}

// Optional cap on the number of CTAs (= SMs) the persistent GEMM occupies, so that a concurrent kernel on
// another stream (the peer-memory all-reduce of the previous token chunk) finds free SMs.  0 = all SMs.
static thread_local int g_sm_limit = 0;       // per host thread: a launch-time setting of the thread that issues the GEMMs
void set_sm_limit(int n) { g_sm_limit = n > 0 ? n : 0; }

// per-device caches (a process may drive several GPUs, from several threads): SM count and, per kernel instantiation,
// whether the dynamic shared memory attribute has been raised on that device
constexpr int kMaxDevices = 64;

static int current_device() {
    int dev = 0;
    return cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < kMaxDevices ? dev : -1;
}

static int num_sms() {
    static std::atomic<int> cache[kMaxDevices];
    const int dev = current_device();
    if (dev < 0) return 0;
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 0;
        cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

// FLEXQ_PDL=0 disables programmatic dependent launch (A/B experiments)
static bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FLEXQ_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// FLEXQ_CLUSTER_RUNTIME=0 disables the cluster launch of decode tiles (A/B experiments)
static bool cluster_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FLEXQ_CLUSTER_RUNTIME");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

struct GemmArgs {
    const int8_t* xq;
    const float* sx;
    const __half* w_scale;
    GemmParams p;
};

template <int M_TILE, int GP, bool DUMP, bool TRACE = false>
static int launch(const GemmArgs& a, cudaStream_t stream) {
    using C = Cfg<M_TILE, GP>;
    static_assert(C::SMEM_BYTES <= 232448, "shared memory budget exceeded");
    GemmParams p = a.p;
    PFN_tmapEncodeTiled enc = get_encode_fn();
    const int sms = num_sms();
    if (!enc || sms <= 0) return FLEXQ_ERR_NO_DEVICE;
    const cuuint32_t ones[3] = {1u, 1u, 1u};

    CUtensorMap tmap_x, tmap_sx, tmap_sw, tmap_w;
    {   // packed weights as a [total_bytes/128][128 B] byte matrix: one tile = 96 consecutive rows
        const cuuint64_t rows = (cuuint64_t)ceil_div(p.N, kTileN) * p.G * (kTileBytes / 128);
        const cuuint64_t dims[2] = {128u, rows};
        const cuuint64_t strides[1] = {128u};
        const cuuint32_t box[2] = {128u, (cuuint32_t)(kTileBytes / 128)};
        if (encode_2d_cached(enc, &tmap_w, CU_TENSOR_MAP_DATA_TYPE_UINT8, const_cast<uint8_t*>(p.w6), dims, strides, box, ones,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B) != CUDA_SUCCESS)
            return FLEXQ_ERR_TENSORMAP;
    }
    {   // activations [M][K] int8, box = 128 bytes of k x M_TILE tokens, 128-byte swizzle
        const cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.M};
        const cuuint64_t strides[1] = {(cuuint64_t)p.K};
        const cuuint32_t box[2] = {128u, (cuuint32_t)M_TILE};
        if (encode_2d_cached(enc, &tmap_x, CU_TENSOR_MAP_DATA_TYPE_UINT8, const_cast<int8_t*>(a.xq), dims, strides, box, ones,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B) != CUDA_SUCCESS)
            return FLEXQ_ERR_TENSORMAP;
    }
    if (!DUMP) {
        const int ldsx = ceil4(p.M);
        const cuuint64_t dims_sx[2] = {(cuuint64_t)ldsx, (cuuint64_t)p.G};
        const cuuint64_t str_sx[1] = {(cuuint64_t)ldsx * 4};
        const cuuint32_t box_sx[2] = {(cuuint32_t)M_TILE, (cuuint32_t)GP};
        if (encode_2d_cached(enc, &tmap_sx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, const_cast<float*>(a.sx), dims_sx, str_sx, box_sx, ones,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE) != CUDA_SUCCESS)
            return FLEXQ_ERR_TENSORMAP;
        const cuuint64_t dims_sw[2] = {(cuuint64_t)p.N, (cuuint64_t)p.G};
        const cuuint64_t str_sw[1] = {(cuuint64_t)p.N * 2};
        const cuuint32_t box_sw[2] = {(cuuint32_t)kTileN, (cuuint32_t)GP};
        if (encode_2d_cached(enc, &tmap_sw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, const_cast<__half*>(a.w_scale), dims_sw, str_sw, box_sw, ones,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE) != CUDA_SUCCESS)
            return FLEXQ_ERR_TENSORMAP;
    } else {
        tmap_sx = tmap_x;
        tmap_sw = tmap_x;
    }

    CUtensorMap tmap_d = tmap_x;
    if (C::TSTORE && !DUMP) {   // D [M][N] fp16 seen as boxes of 64 weight rows (128 bytes, swizzle-128B) x 64 tokens
        const cuuint64_t dims_d[2] = {(cuuint64_t)p.ldd, (cuuint64_t)p.M};
        const cuuint64_t str_d[1] = {(cuuint64_t)p.ldd * 2};
        const cuuint32_t box_d[2] = {64u, 64u};
        if (encode_2d_cached(enc, &tmap_d, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, p.D, dims_d, str_d, box_d, ones, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_NONE) != CUDA_SUCCESS)
            return FLEXQ_ERR_TENSORMAP;
    }
    p.n_tiles = ceil_div(p.N, kTileN);
    p.m_tiles = ceil_div(p.M, M_TILE);
    {
        int max_ctas = sms < kMaxCtas ? sms : kMaxCtas;
        if (g_sm_limit > 0 && g_sm_limit < max_ctas) max_ctas = g_sm_limit;
        plan_ctas(p, max_ctas, GP, kFixSteps<M_TILE>);
    }
    const int P = p.P;

    static std::atomic<bool> attr_set[kMaxDevices];      // per instantiation and device; setting it twice is harmless
    const int dev = current_device();
    if (dev < 0) return FLEXQ_ERR_NO_DEVICE;
    if (!attr_set[dev].load(std::memory_order_acquire)) {
        FLEXQ_CUDA_TRY(cudaFuncSetAttribute(w6ax_gemm_kernel<M_TILE, GP, DUMP, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set[dev].store(true, std::memory_order_release);
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(P);
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        na++;
    }
    cfg.attrs = attr;
    // Decode tiles, aligned plan with C runs per weight tile: launch the C CTAs of a tile as one cluster (see kClusterOk in
    // the kernel) when C partial tiles fit the weight ring of the first CTA and the device can hold all clusters at once
    {   // aligned plan = every CTA walks one run of one tile (plan_ctas): Pn a multiple of n_tiles, no spare CTAs
        static const int force_handoff = [] { const char* e = getenv("FLEXQ_HANDOFF"); return e ? atoi(e) : -1; }();
        const bool aligned = !p.whole_rows && p.P == p.R * p.Pn && p.Pn % p.n_tiles == 0;
        // measured (profiles/r2_experiments/sweep_b27_*): with four or more runs per tile finishing together the completing
        // CTA of a hand-off waits for too many parked runs (reductions 6-12 % faster); with two, hand-off stays ahead (4-20 %)
        p.handoff = force_handoff >= 0 ? force_handoff : (aligned && p.Pn / p.n_tiles > 2 ? 0 : 1);
    }
    p.cluster = 0;
    if (FLEXQ_CLUSTER && !DUMP && M_TILE <= kClusterMaxTile && !p.whole_rows && p.m_tiles == 1 && p.P == p.Pn && p.Pn % p.n_tiles == 0 && cluster_enabled()) {
        const int Cn = p.Pn / p.n_tiles;
        constexpr int kFit = C::RECV_BYTES > 0 ? 4 : 1 + (C::NW * C::W_BYTES) / (M_TILE * kTileN * 4);
        if (Cn >= 2 && Cn <= 8 && Cn <= kFit) {
            static std::atomic<int> resident[kMaxDevices][9];         // 1 + clusters of this size the device holds at once (0 = unknown)
            int f = resident[dev][Cn].load(std::memory_order_acquire);
            attr[na].id = cudaLaunchAttributeClusterDimension;
            attr[na].val.clusterDim.x = Cn; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
            cfg.numAttrs = na + 1;
            if (f == 0) {
                int nclusters = 0;
                if (cudaOccupancyMaxActiveClusters(&nclusters, w6ax_gemm_kernel<M_TILE, GP, DUMP, TRACE>, &cfg) != cudaSuccess) {
                    (void)cudaGetLastError();
                    nclusters = 0;
                }
                f = 1 + nclusters;
                resident[dev][Cn].store(f, std::memory_order_release);
            }
            if (f - 1 >= p.P / Cn) {
                static std::atomic<unsigned> nonce{0x9E3779B9u};
                p.cluster = Cn; na++;
                p.nonce = nonce.fetch_add(0x9E3779B9u, std::memory_order_relaxed) | 1u;
            }       // a second wave of clusters would double the kernel
        }
    }
    cfg.numAttrs = na;
    return (int)cudaLaunchKernelEx(&cfg, w6ax_gemm_kernel<M_TILE, GP, DUMP, TRACE>, tmap_x, tmap_sx, tmap_sw, tmap_w, tmap_d, p);
}

template <bool DUMP>
static int dispatch(const GemmArgs& a, cudaStream_t stream) {
    const int M = a.p.M;
    if (M <= 16) return launch<16, 4, DUMP>(a, stream);
    if (M <= 32) return launch<32, 4, DUMP>(a, stream);
    if (M <= 64) return launch<64, 2, DUMP>(a, stream);
    if (M <= 128) return launch<128, 1, DUMP>(a, stream);
    // Token tile for large M, by estimated length of the two plans: a 192-token step runs 1.2 x as long as a 128-token
    // step (measured, equal padded rows: 192 is 20-30 % faster per row), summing a cut tile costs ~16 steps either way.
    static int force = -1;
    if (force < 0) { const char* e = getenv("FLEXQ_MTILE"); force = e ? atoi(e) : 0; }
    bool use128;
    if (force) use128 = force == 128;
    else {
        const int sms = num_sms();
        int max_ctas = sms < kMaxCtas ? sms : kMaxCtas;
        if (g_sm_limit > 0 && g_sm_limit < max_ctas) max_ctas = g_sm_limit;
        GemmParams q = a.p;
        q.n_tiles = ceil_div(q.N, kTileN);
        q.m_tiles = ceil_div(M, 192);
        const int len192 = plan_ctas(q, max_ctas, 1, kFixSteps<192>);
        q.m_tiles = ceil_div(M, 128);
        const int len128 = plan_ctas(q, max_ctas, 1, kFixSteps<128>);
        use128 = len128 * 5 < len192 * 6;
    }
    if (use128) return launch<128, 1, DUMP>(a, stream);
    return launch<192, 1, DUMP>(a, stream);
}

static int check_shape(int M, int N, int K) {
    // N % 8: w_scale rows are fetched by TMA (16-byte row pitch), as the reference's 16-byte stores need
    return (M <= 0 || N <= 0 || N % 8 || K < kGroup || K % kGroup) ? FLEXQ_ERR_BAD_SHAPE : 0;
}

static GemmArgs make_args(const int8_t* xq, const float* sx, const uint8_t* w6, const __half* w_scale, __half* D, int32_t* S,
                          int M, int N, int K, void* workspace) {
    GemmArgs a{};
    a.xq = xq; a.sx = sx; a.w_scale = w_scale;
    a.p.w6 = w6; a.p.D = D; a.p.S = S;
    if (workspace) {
        a.p.cnt = reinterpret_cast<int*>(workspace);
        a.p.slots = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + kCntBytes);
    }
    a.p.M = M; a.p.N = N; a.p.K = K; a.p.G = K / kGroup;
    a.p.ldd = N; a.p.silu = 0;
    return a;
}

int gemm_w6ax(const int8_t* xq, const float* sx, const uint8_t* w6, const __half* w_scale, __half* D, int M, int N, int K,
              void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!xq || !sx || !w6 || !w_scale || !D || !workspace) return FLEXQ_ERR_NULL;
    if (int e = check_shape(M, N, K)) return e;
    if (workspace_bytes < flexq_gemm_workspace_bytes() || ((uintptr_t)workspace & 15)) return FLEXQ_ERR_WORKSPACE;
    return dispatch<false>(make_args(xq, sx, w6, w_scale, D, nullptr, M, N, K, workspace), stream);
}

// gate_up GEMM with SiLU(gate) * up applied by the epilogue (first half): w6 / w_scale hold 2 * inter rows,
// 8 gate rows followed by their 8 up rows (model_pack.interleave_gate_up); H is [M][inter] fp16
int gemm_w6ax_silu_mul(const int8_t* xq, const float* sx, const uint8_t* w6, const __half* w_scale, __half* H, int M, int inter, int K,
                       void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!xq || !sx || !w6 || !w_scale || !H || !workspace) return FLEXQ_ERR_NULL;
    if (inter <= 0 || inter % 8) return FLEXQ_ERR_BAD_SHAPE;
    if (int e = check_shape(M, 2 * inter, K)) return e;
    if (workspace_bytes < flexq_gemm_workspace_bytes() || ((uintptr_t)workspace & 15)) return FLEXQ_ERR_WORKSPACE;
    GemmArgs a = make_args(xq, sx, w6, w_scale, H, nullptr, M, 2 * inter, K, workspace);
    a.p.ldd = inter; a.p.silu = 1;
    return dispatch<false>(a, stream);
}

// debug: same GEMM with clock64 stamps of CTA 0's pipeline events (tools/trace.py)
int gemm_w6ax_trace(const int8_t* xq, const float* sx, const uint8_t* w6, const __half* w_scale, __half* D, int M, int N, int K,
                    void* workspace, long long* trace, int trace_units, cudaStream_t stream) {
    GemmArgs a = make_args(xq, sx, w6, w_scale, D, nullptr, M, N, K, workspace);
    a.p.trace = trace; a.p.trace_units = trace_units;
    if (M <= 16) return launch<16, 4, false, true>(a, stream);
    if (M <= 32) return launch<32, 4, false, true>(a, stream);
    if (M <= 64) return launch<64, 2, false, true>(a, stream);
    if (M <= 128) return launch<128, 1, false, true>(a, stream);
    return launch<192, 1, false, true>(a, stream);
}

int gemm_w6ax_groupsums(const int8_t* xq, const uint8_t* w6, int32_t* S, int M, int N, int K, cudaStream_t stream) {
    if (!xq || !w6 || !S) return FLEXQ_ERR_NULL;
    if (int e = check_shape(M, N, K)) return e;
    return dispatch<true>(make_args(xq, nullptr, w6, nullptr, nullptr, S, M, N, K, nullptr), stream);
}

}  // namespace flexq
