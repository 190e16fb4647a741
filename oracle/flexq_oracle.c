/*
 * CPU oracle (plain C) for the FlexQ W6Ax hot path.  TEST INFRASTRUCTURE ONLY: loaded by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; never by the product.
 *
 * Each function restates one piece of the reference (citations are path:line under
 * /root/reference).  The restatement is literal where the reference's own tests pin
 * behaviour: fq_compute_ref keeps the exact loop nest and float/double promotion of the
 * reference CPU golden so that its fp16 output can be compared bit-for-bit.
 *
 * Build: make -C oracle   (-> oracle/_build/libflexq_oracle.so)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FQ_GROUP 128

/* ---- IEEE binary16 <-> binary32 (round-to-nearest-even), no compiler extensions ---- */
static float h2f(uint16_t h)
{
    uint32_t s = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1Fu, m = h & 0x3FFu, u;
    if (e == 0) {
        if (m == 0) { u = s; }
        else {
            int sh = 0;
            while (!(m & 0x400u)) { m <<= 1; sh++; }
            m &= 0x3FFu;
            u = s | ((uint32_t)(127 - 15 - sh + 1) << 23) | (m << 13);
        }
    } else if (e == 31) { u = s | 0x7F800000u | (m << 13); }
    else { u = s | ((e + 112u) << 23) | (m << 13); }
    float f; memcpy(&f, &u, 4); return f;
}

static uint16_t f2h(float f)
{
    uint32_t x; memcpy(&x, &f, 4);
    uint32_t s = (x >> 16) & 0x8000u, e = (x >> 23) & 0xFFu, m = x & 0x7FFFFFu;
    if (e == 255) return (uint16_t)(s | 0x7C00u | (m ? 0x200u : 0));
    int ne = (int)e - 127 + 15;
    if (ne >= 31) return (uint16_t)(s | 0x7C00u);
    if (ne <= 0) {
        if (ne < -10) return (uint16_t)s;
        m |= 0x800000u;
        int shift = 14 - ne;
        uint32_t r = m >> shift, rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (r & 1))) r++;
        return (uint16_t)(s | r);
    }
    uint32_t r = ((uint32_t)ne << 10) | (m >> 13), rem = m & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1))) r++;
    return (uint16_t)(s | r);
}

float fq_half_to_float(uint16_t h) { return h2f(h); }
uint16_t fq_float_to_half(float f) { return f2h(f); }

/* ---- a5: engine/src/pack/bit_packing.cu:42-99 -------------------------------------- */
/* in: int32 [R][K] two's complement; out: u32 [K/128][R/chunk][bits][chunk][4];
 * word bit 31 = element k%32==0 (":75 __brev(__ballot_sync)").  Index formula as in the
 * reference's validator engine/test_packing_kernel.cu:139-141. */
int fq_pack_planes(const int32_t *in, uint32_t *out, int R, int K, int bits)
{
    int chunk = R < 8 ? R : 8;
    if (K % FQ_GROUP || R % chunk) return -1;
    for (int b = 0; b < bits; b++)
        for (int r = 0; r < R; r++)
            for (int k32 = 0; k32 < K / 32; k32++) {
                uint32_t w = 0;
                for (int l = 0; l < 32; l++)
                    w |= (uint32_t)((in[(size_t)r * K + k32 * 32 + l] >> b) & 1) << (31 - l);
                size_t idx = (size_t)(k32 / 4) * ((size_t)R * bits * 4) + (size_t)(r / chunk) * (bits * chunk * 4)
                           + (size_t)b * (chunk * 4) + (size_t)(r % chunk) * 4 + (k32 % 4);
                out[idx] = w;
            }
    return 0;
}

/* ---- a6: e2e/.../flexqgemm/src/pack/bit_packing.cu:119-166 (IEEE division) ---------- */
/* x: half bits [M][K]; q: int32 [M][K]; scale: half bits [M][G] */
int fq_quant_act(const uint16_t *x, int32_t *q, uint16_t *scale, int M, int K, int bits)
{
    if (K % FQ_GROUP) return -1;
    int G = K / FQ_GROUP, lo = -(1 << (bits - 1)), hi = (1 << (bits - 1)) - 1;
    for (int m = 0; m < M; m++)
        for (int g = 0; g < G; g++) {
            const uint16_t *p = x + (size_t)m * K + g * FQ_GROUP;
            float maxv = -1.0f;                                   /* :125 half maxv_h = -1 */
            for (int i = 0; i < FQ_GROUP; i++) { float a = fabsf(h2f(p[i])); if (a > maxv) maxv = a; }
            maxv /= (float)hi;                                    /* :151 */
            uint16_t sh = f2h(maxv);                              /* :155 */
            scale[(size_t)m * G + g] = sh;
            float r = h2f(sh);                                    /* :158 */
            for (int i = 0; i < FQ_GROUP; i++) {
                float t = roundf(h2f(p[i]) / r);                  /* :160 */
                int v;                                            /* (int): NaN->0, saturating */
                if (t != t) v = 0; else if (t >= 2147483648.0f) v = INT32_MAX;
                else if (t <= -2147483648.0f) v = INT32_MIN; else v = (int)t;
                q[(size_t)m * K + g * FQ_GROUP + i] = v < lo ? lo : (v > hi ? hi : v);
            }
        }
    return 0;
}

/* ---- per-group INT32 sums (mathematical content of flexq_bmma_kernel.h:351-406) ----- */
/* xq [M][K], wq [N][K] int32 -> S [M][N][G] */
void fq_group_sums(const int32_t *xq, const int32_t *wq, int32_t *S, int M, int N, int K)
{
    int G = K / FQ_GROUP;
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++)
            for (int g = 0; g < G; g++) {
                int32_t acc = 0;
                const int32_t *a = xq + (size_t)m * K + g * FQ_GROUP, *b = wq + (size_t)n * K + g * FQ_GROUP;
                for (int k = 0; k < FQ_GROUP; k++) acc += a[k] * b[k];
                S[((size_t)m * N + n) * G + g] = acc;
            }
}

/* ---- a11: engine/test_bgemm_kernel.cu:113-146, literal ------------------------------ */
static int int_pow(int base, int e) { int r = 1; while (e) { if (e % 2) r *= base; e /= 2; base *= base; } return r; }

void fq_compute_ref(const int32_t *w, const uint16_t *w_scale, const int32_t *x, const uint16_t *x_scale,
                    uint16_t *ref_c, int M, int N, int K, int W_BIT, int X_BIT, int SIGNED, int group_size)
{
    int chunk_m = M < 8 ? M : 8, chunk_n = N < 8 ? N : 8;
    int x_scale_ld = 2 * ((M + 3) / 4 * 4);                       /* SCALE_PACKING_A(SCALE_SIZE_X(M)) */
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            float tmp = 0;
            for (int xb = 0; xb < X_BIT; xb++) {
                int XM = SIGNED && (xb == X_BIT - 1) ? -1 * int_pow(2, xb) : int_pow(2, xb);
                for (int wb = 0; wb < W_BIT; wb++) {
                    int WM = SIGNED && (wb == W_BIT - 1) ? -1 * int_pow(2, wb) : int_pow(2, wb);
                    for (int kt = 0; kt < K / 32; kt++) {
                        int w_int = w[(size_t)(kt / 4) * (N * W_BIT * 4) + (n / chunk_n) * (W_BIT * chunk_n * 4) + wb * (chunk_n * 4) + (n % chunk_n) * 4 + (kt % 4)];
                        int x_int = x[(size_t)(kt / 4) * (M * X_BIT * 4) + (m / chunk_m) * (X_BIT * chunk_m * 4) + xb * (chunk_m * 4) + (m % chunk_m) * 4 + (kt % 4)];
                        float ws = h2f(w_scale[(size_t)(kt / (group_size / 32)) * N + n]);
                        float xs = h2f(x_scale[(size_t)(kt / (group_size / 32)) * x_scale_ld + 2 * m]);
                        for (int k = 0; k < 32; k++) {
                            uint32_t mask = 1u << k;
                            int xv = (int)(((uint32_t)x_int & mask) >> k);
                            int wv = (int)(((uint32_t)w_int & mask) >> k);
                            /* reference: ((mask<<k)&v)>>k on signed int gives -1 for bit 31; the product
                             * xv*wv is 1 iff both bits are set in either formulation. */
                            tmp += 1.0 * (XM * WM * xv * wv) * ws * xs;   /* double product, float accumulate */
                        }
                    }
                }
            }
            ref_c[(size_t)m * N + n] = f2h(tmp);
        }
}

/* ---- a1/a2 python path (algorithm/flexq_quantize/quantizer.py:93-171), fp32 ---------- */
/* x [rows][K] float -> q int32, scale [rows][G] float, deq float.  torch.round = half-even. */
int fq_quant_python_f32(const float *x, int32_t *q, float *scale, float *deq, int rows, int K, int bits, int group)
{
    if (K % group) return -1;
    int G = K / group; float qmin = -(float)(1 << (bits - 1)), qmax = (float)((1 << (bits - 1)) - 1);
    for (int r = 0; r < rows; r++)
        for (int g = 0; g < G; g++) {
            const float *p = x + (size_t)r * K + g * group;
            float mn = p[0], mx = p[0];
            for (int i = 1; i < group; i++) { if (p[i] < mn) mn = p[i]; if (p[i] > mx) mx = p[i]; }
            float am = fabsf(mx) > fabsf(mn) ? fabsf(mx) : fabsf(mn);   /* :153 */
            float s = am / qmax;                                        /* :154 */
            if (s < 1e-5f) s = 1e-5f; if (s > 1e4f) s = 1e4f;           /* :155 */
            scale[(size_t)r * G + g] = s;
            for (int i = 0; i < group; i++) {
                float t = nearbyintf(p[i] / s);                         /* :112 */
                t = t < qmin ? qmin : (t > qmax ? qmax : t);            /* :116 */
                q[(size_t)r * K + g * group + i] = (int32_t)t;
                deq[(size_t)r * K + g * group + i] = t * s;             /* :122 */
            }
        }
    return 0;
}
