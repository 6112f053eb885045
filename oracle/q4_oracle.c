/*
 * q4_oracle.c -- CPU restatement of the fastLLaMa q4_0/q4_1 matmul hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under fastllama_b200/ may include, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * Every function restates, in plain scalar C, what the reference computes on an
 * x86-64 AVX2 host (the build the reference's own CMake produces with
 * -march=native on any recent Xeon/EPYC).  Citations are file:line relative to
 * the reference tree (lib/ggml.c unless stated otherwise).
 *
 * Parity status: the reference ships no tests or golden vectors for this path
 * (SURVEY.md section 4), so this restatement is pinned against the reference
 * itself: oracle/Makefile compiles lib/ggml.c in place into
 * oracle/_ref/libggml_ref.so and tests/test_oracle_vs_ref.py checks every
 * function below bit-for-bit against ggml_internal_get_quantize_fn() of that
 * library, and against the committed fixtures in tests/golden/ that
 * oracle/gen_golden.py produced from it.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_QK 32

/* Block formats: lib/ggml.c:589-626.  All three are packed, no padding. */
typedef struct { float d; uint8_t qs[ORC_QK / 2]; } orc_block_q4_0;            /* 20 B */
typedef struct { float d; float m; uint8_t qs[ORC_QK / 2]; } orc_block_q4_1;   /* 24 B */
typedef struct { float d; float s; int8_t qs[ORC_QK]; } orc_block_q8_0;        /* 40 B */

_Static_assert(sizeof(orc_block_q4_0) == 20, "q4_0 block must be 20 bytes");
_Static_assert(sizeof(orc_block_q4_1) == 24, "q4_1 block must be 24 bytes");
_Static_assert(sizeof(orc_block_q8_0) == 40, "q8_0 block must be 40 bytes");

int orc_block_bytes(int ggml_type) {
    /* GGML_TYPE_SIZE table, lib/ggml.c:3294-3307; enum values include/ggml.h:201-214 */
    switch (ggml_type) {
        case 2: return 20;  /* Q4_0 */
        case 3: return 24;  /* Q4_1 */
        case 6: return 40;  /* Q8_0 */
        default: return -1;
    }
}

/* ------------------------------------------------------------------------- */
/* Activation quantizer: quantize_row_q8_0, AVX2 branch lib/ggml.c:1342-1403  */
/* plus the trailing "s" recomputation at :1433-1440.                         */
/*   d  = amax / 127                (division, :1362)                         */
/*   id = 127 / amax  (0 if amax==0)  -- NOT 1/d (:1364)                      */
/*   q  = round-to-nearest-EVEN(x * id)   (_mm256_round_ps NEAREST, :1374)    */
/*   s  = d * (float)(sum of q)     (int sum converted once, :1436-1439)      */
/* ------------------------------------------------------------------------- */
void orc_quantize_row_q8_0(const float *x, void *vy, int k) {
    orc_block_q8_0 *y = (orc_block_q8_0 *)vy;
    const int nb = k / ORC_QK;
    for (int i = 0; i < nb; i++) {
        const float *xb = x + i * ORC_QK;
        float amax = 0.0f;
        for (int l = 0; l < ORC_QK; l++) {
            const float a = fabsf(xb[l]);
            if (a > amax) amax = a;
        }
        const float d = amax / 127.f;
        const float id = (amax != 0.0f) ? 127.f / amax : 0.0f;
        int sum = 0;
        for (int l = 0; l < ORC_QK; l++) {
            /* nearbyintf honours the current rounding mode = round-half-even */
            const float r = nearbyintf(xb[l] * id);
            int q = (int)r;
            if (q > 127) q = 127;      /* packs_epi32/epi16 saturation, :1388-1392 */
            if (q < -128) q = -128;
            y[i].qs[l] = (int8_t)q;
            sum += q;
        }
        y[i].d = d;
        y[i].s = d * (float)sum;
    }
}

/* Scalar variant the reference keeps for non-SIMD hosts: lib/ggml.c:1249-1274
 * (1/d multiplier, roundf = half away from zero).  Kept to document the
 * difference; the x86 oracle is the AVX2 variant above. */
void orc_quantize_row_q8_0_scalar(const float *x, void *vy, int k) {
    orc_block_q8_0 *y = (orc_block_q8_0 *)vy;
    const int nb = k / ORC_QK;
    for (int i = 0; i < nb; i++) {
        float amax = 0.0f;
        for (int l = 0; l < ORC_QK; l++) {
            const float a = fabsf(x[i * ORC_QK + l]);
            if (a > amax) amax = a;
        }
        const float d = amax / 127.f;
        const float id = d ? 1.0f / d : 0.0f;
        int sum = 0;
        for (int l = 0; l < ORC_QK; l++) {
            const float v = x[i * ORC_QK + l] * id;
            y[i].qs[l] = (int8_t)roundf(v);
            sum += y[i].qs[l];
        }
        y[i].d = d;
        y[i].s = d * sum;
    }
}

/* ------------------------------------------------------------------------- */
/* Weight quantizers that define model-file contents:                        */
/*   quantize_row_q4_0_reference lib/ggml.c:630-664                           */
/*   quantize_row_q4_1_reference lib/ggml.c:917-956                           */
/* nibble packing: qs[j] = q[2j] | q[2j+1] << 4                               */
/* ------------------------------------------------------------------------- */
void orc_quantize_row_q4_0(const float *x, void *vy, int k) {
    orc_block_q4_0 *y = (orc_block_q4_0 *)vy;
    const int nb = k / ORC_QK;
    for (int i = 0; i < nb; i++) {
        const float *xb = x + i * ORC_QK;
        float amax = 0.0f;
        for (int l = 0; l < ORC_QK; l++) {
            const float a = fabsf(xb[l]);
            if (a > amax) amax = a;
        }
        const float d = amax / 7.0f;             /* (1<<3)-1 */
        const float id = d ? 1.0f / d : 0.0f;
        y[i].d = d;
        for (int l = 0; l < ORC_QK; l += 2) {
            const uint8_t q0 = (uint8_t)((int8_t)roundf(xb[l + 0] * id) + 8);
            const uint8_t q1 = (uint8_t)((int8_t)roundf(xb[l + 1] * id) + 8);
            y[i].qs[l / 2] = (uint8_t)(q0 | (q1 << 4));
        }
    }
}

void orc_quantize_row_q4_1(const float *x, void *vy, int k) {
    orc_block_q4_1 *y = (orc_block_q4_1 *)vy;
    const int nb = k / ORC_QK;
    for (int i = 0; i < nb; i++) {
        const float *xb = x + i * ORC_QK;
        float mn = FLT_MAX, mx = -FLT_MAX;
        for (int l = 0; l < ORC_QK; l++) {
            if (xb[l] < mn) mn = xb[l];
            if (xb[l] > mx) mx = xb[l];
        }
        const float d = (mx - mn) / 15.0f;       /* (1<<4)-1 */
        const float id = d ? 1.0f / d : 0.0f;
        y[i].d = d;
        y[i].m = mn;
        for (int l = 0; l < ORC_QK; l += 2) {
            const uint8_t q0 = (uint8_t)roundf((xb[l + 0] - mn) * id);
            const uint8_t q1 = (uint8_t)roundf((xb[l + 1] - mn) * id);
            y[i].qs[l / 2] = (uint8_t)(q0 | (q1 << 4));
        }
    }
}

/* ------------------------------------------------------------------------- */
/* The SIMD weight quantizers = quantize_fns[type].quantize_row_q, AVX2 branches */
/*   quantize_row_q4_0 lib/ggml.c:739-803, quantize_row_q4_1 lib/ggml.c:965-1038  */
/* (what ggml_compute_forward_add_q_f32 :6516-6518 re-quantises a LoRA-merged row */
/* with).  Differences from the _reference variants above:                       */
/*   q4_0: id = 7 / amax (not 1 / d), round-half-EVEN (_mm256_round_ps NEAREST), */
/*         saturating packs to int8, then + 8                                    */
/*   q4_1: id = d ? 1 / d : 0 as in the reference, but round-half-EVEN and the   */
/*         saturating int8 packs                                                 */
/* packNibbles (:473-487) keeps the low 4 bits of every byte pair.               */
/* ------------------------------------------------------------------------- */
static inline int orc_sat_i8(int v) { return v > 127 ? 127 : (v < -128 ? -128 : v); }

void orc_quantize_row_q4_0_simd(const float *x, void *vy, int k) {
    orc_block_q4_0 *y = (orc_block_q4_0 *)vy;
    const int nb = k / ORC_QK;
    for (int i = 0; i < nb; i++) {
        const float *xb = x + i * ORC_QK;
        float amax = 0.0f;
        for (int l = 0; l < ORC_QK; l++) {
            const float a = fabsf(xb[l]);
            if (a > amax) amax = a;
        }
        const float d = amax / 7.0f;
        const float id = (amax != 0.0f) ? 7.0f / amax : 0.0f;
        y[i].d = d;
        uint8_t q[ORC_QK];
        for (int l = 0; l < ORC_QK; l++) {
            const int v = orc_sat_i8((int)nearbyintf(xb[l] * id));   /* |x * id| <= 7 (+ rounding), never saturates in practice */
            q[l] = (uint8_t)((int8_t)(v + 8));                       /* _mm256_add_epi8: wraps modulo 256 */
        }
        for (int l = 0; l < ORC_QK; l += 2) y[i].qs[l / 2] = (uint8_t)((q[l] & 0x0F) | ((q[l + 1] & 0x0F) << 4));
    }
}

void orc_quantize_row_q4_1_simd(const float *x, void *vy, int k) {
    orc_block_q4_1 *y = (orc_block_q4_1 *)vy;
    const int nb = k / ORC_QK;
    for (int i = 0; i < nb; i++) {
        const float *xb = x + i * ORC_QK;
        /* _mm256_max_ps / _mm256_min_ps over finite inputs = plain max / min */
        float mn = xb[0], mx = xb[0];
        for (int l = 1; l < ORC_QK; l++) {
            if (xb[l] < mn) mn = xb[l];
            if (xb[l] > mx) mx = xb[l];
        }
        const float d = (mx - mn) / 15.0f;
        const float id = d ? 1.0f / d : 0.0f;
        y[i].m = mn;
        y[i].d = d;
        uint8_t q[ORC_QK];
        for (int l = 0; l < ORC_QK; l++) q[l] = (uint8_t)orc_sat_i8((int)nearbyintf((xb[l] - mn) * id));
        for (int l = 0; l < ORC_QK; l += 2) y[i].qs[l / 2] = (uint8_t)((q[l] & 0x0F) | ((q[l + 1] & 0x0F) << 4));
    }
}

/* ggml_vec_dot_f32 (lib/ggml.c:2295-2325) as the AVX2 + FMA build computes it: 4 accumulators of 8 lanes over steps of 32    */
/* (GGML_F32_STEP 32, GGML_F32_EPR 8), the GGML_F32x8_REDUCE tree (:1943-1958), then the leftover loop                        */
/* "sumf += x[i]*y[i]" as gcc -O3 compiles it for this target: the loop is vectorised (products 8 at a time, then one group   */
/* of 4, each product ROUNDED and added in order -- no fma), and the scalar epilogue of the last <= 3 elements is contracted   */
/* to an fma.  Pinned by tests/golden/lora_ops.npz (leftovers 8, 16) and tests/golden/f32_dot.npz (every leftover count).      */
float orc_vec_dot_f32(int n, const float *x, const float *y) {
    float sum[4][8];
    for (int j = 0; j < 4; j++) for (int l = 0; l < 8; l++) sum[j][l] = 0.0f;
    const int np = n & ~31;
    for (int i = 0; i < np; i += 32)
        for (int j = 0; j < 4; j++)
            for (int l = 0; l < 8; l++) sum[j][l] = fmaf(x[i + 8 * j + l], y[i + 8 * j + l], sum[j][l]);
    float t[8];
    for (int l = 0; l < 8; l++) t[l] = (sum[0][l] + sum[1][l]) + (sum[2][l] + sum[3][l]);
    float sumf = ((t[0] + t[4]) + (t[1] + t[5])) + ((t[2] + t[6]) + (t[3] + t[7]));
    const int rem = n - np;
    const int nma = (rem & ~7) + ((rem & 4) ? 4 : 0);
    for (int i = np; i < np + nma; i++) { const float p = x[i] * y[i]; sumf = sumf + p; }
    for (int i = np + nma; i < n; i++) sumf = fmaf(x[i], y[i], sumf);
    return sumf;
}

/* ggml_compute_forward_add_q_f32 (lib/ggml.c:6414-6520): dst row = quantize_row_q(dequantize_row_q(src0 row) + src1 row) */
void orc_dequantize_row_q4_0(const void *vx, float *y, int k);
void orc_dequantize_row_q4_1(const void *vx, float *y, int k);
int orc_add_q_f32(int ggml_type, int rows, int k, const void *src0, const float *src1, void *dst) {
    const int bb = orc_block_bytes(ggml_type);
    if (bb < 0 || ggml_type == 6 || k % ORC_QK) return -1;
    float *w = (float *)malloc(sizeof(float) * (size_t)k);
    for (int r = 0; r < rows; r++) {
        const uint8_t *s0 = (const uint8_t *)src0 + (size_t)r * (k / ORC_QK) * bb;
        uint8_t *d = (uint8_t *)dst + (size_t)r * (k / ORC_QK) * bb;
        if (ggml_type == 2) orc_dequantize_row_q4_0(s0, w, k); else orc_dequantize_row_q4_1(s0, w, k);
        for (int i = 0; i < k; i++) w[i] += src1[(size_t)r * k + i];      /* ggml_vec_acc_f32, :2286 */
        if (ggml_type == 2) orc_quantize_row_q4_0_simd(w, d, k); else orc_quantize_row_q4_1_simd(w, d, k);
    }
    free(w);
    return 0;
}

/* ggml_quantize_q4_0 / _q4_1 (lib/ggml.c:12122-12166): n floats, rows of k,
 * returns bytes written.  The histogram side effect is not reproduced. */
size_t orc_quantize_q4(int ggml_type, const float *src, void *dst, int n, int k) {
    const int bb = orc_block_bytes(ggml_type);
    const int nb = k / ORC_QK;
    for (int j = 0; j < n; j += k) {
        char *row = (char *)dst + (size_t)(j / ORC_QK) * bb;
        if (ggml_type == 2) orc_quantize_row_q4_0(src + j, row, k);
        else                orc_quantize_row_q4_1(src + j, row, k);
    }
    (void)nb;
    return (size_t)(n / ORC_QK) * bb;
}

/* ------------------------------------------------------------------------- */
/* De-quantizers (on the path through get_rows, lib/ggml.c:8333-8360):        */
/*   dequantize_row_q4_0 :1443-1559   y = (q - 8) * d                         */
/*   dequantize_row_q4_1 :1561-1665   y = q * d + m  (AVX2: fmadd, :1585)     */
/* The AVX2 q4_1 kernel uses a fused multiply-add, the scalar tail uses       */
/* mul+add; restated with fmaf to follow the x86 build.                       */
/* ------------------------------------------------------------------------- */
void orc_dequantize_row_q4_0(const void *vx, float *y, int k) {
    const orc_block_q4_0 *x = (const orc_block_q4_0 *)vx;
    const int nb = k / ORC_QK;
    for (int i = 0; i < nb; i++) {
        const float d = x[i].d;
        for (int j = 0; j < ORC_QK / 2; j++) {
            const int lo = x[i].qs[j] & 0x0F;
            const int hi = x[i].qs[j] >> 4;
            y[i * ORC_QK + 2 * j + 0] = (float)(lo - 8) * d;
            y[i * ORC_QK + 2 * j + 1] = (float)(hi - 8) * d;
        }
    }
}

void orc_dequantize_row_q4_1(const void *vx, float *y, int k) {
    const orc_block_q4_1 *x = (const orc_block_q4_1 *)vx;
    const int nb = k / ORC_QK;
    for (int i = 0; i < nb; i++) {
        const float d = x[i].d, m = x[i].m;
        for (int j = 0; j < ORC_QK / 2; j++) {
            const int lo = x[i].qs[j] & 0x0F;
            const int hi = x[i].qs[j] >> 4;
            y[i * ORC_QK + 2 * j + 0] = fmaf((float)lo, d, m);
            y[i * ORC_QK + 2 * j + 1] = fmaf((float)hi, d, m);
        }
    }
}

/* ------------------------------------------------------------------------- */
/* Block dot products.                                                        */
/*   ggml_vec_dot_q4_0_q8_0, AVX2 branch lib/ggml.c:2445-2487                 */
/*   ggml_vec_dot_q4_1_q8_0, AVX2 branch lib/ggml.c:2639-2687                 */
/* The AVX2 code keeps 8 fp32 lane accumulators; lane l sums the 4 products   */
/* of elements 4l..4l+3 of each block as an exact integer (maddubs + madd),   */
/* converts it to float and does acc[l] = fma(dx*dy, q, acc[l]) block after   */
/* block, then reduces the lanes as                                           */
/*   ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7))            (:2482-2487)            */
/* Restated lane for lane so the result is bit-identical to the x86 build.    */
/* ------------------------------------------------------------------------- */
static inline void orc_lane_sums_q4(const uint8_t *qs, const int8_t *q8, int off, int lanes[8]) {
    for (int l = 0; l < 8; l++) {
        int s = 0;
        for (int e = 0; e < 4; e++) {
            const int idx = 4 * l + e;                 /* element index 0..31 */
            const uint8_t byte = qs[idx >> 1];
            const int q4 = (idx & 1) ? (byte >> 4) : (byte & 0x0F);
            s += (q4 - off) * (int)q8[idx];
        }
        lanes[l] = s;
    }
}

static inline float orc_hsum8(const float a[8]) {
    const float r0 = a[0] + a[4], r1 = a[1] + a[5], r2 = a[2] + a[6], r3 = a[3] + a[7];
    const float s0 = r0 + r2, s1 = r1 + r3;
    return s0 + s1;
}

void orc_vec_dot_q4_0_q8_0(int n, float *s, const void *vx, const void *vy) {
    const orc_block_q4_0 *x = (const orc_block_q4_0 *)vx;
    const orc_block_q8_0 *y = (const orc_block_q8_0 *)vy;
    const int nb = n / ORC_QK;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < nb; i++) {
        const float d = x[i].d * y[i].d;
        int lanes[8];
        orc_lane_sums_q4(x[i].qs, y[i].qs, 8, lanes);
        for (int l = 0; l < 8; l++) acc[l] = fmaf(d, (float)lanes[l], acc[l]);
    }
    *s = orc_hsum8(acc);
}

void orc_vec_dot_q4_1_q8_0(int n, float *s, const void *vx, const void *vy) {
    const orc_block_q4_1 *x = (const orc_block_q4_1 *)vx;
    const orc_block_q8_0 *y = (const orc_block_q8_0 *)vy;
    const int nb = n / ORC_QK;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float summs = 0.0f;
    for (int i = 0; i < nb; i++) {
        /* gcc -O3 contracts "summs += m*s" into an fma on FMA hosts (:2651) */
        summs = fmaf(x[i].m, y[i].s, summs);
        const float d = x[i].d * y[i].d;
        int lanes[8];
        orc_lane_sums_q4(x[i].qs, y[i].qs, 0, lanes);
        for (int l = 0; l < 8; l++) acc[l] = fmaf(d, (float)lanes[l], acc[l]);
    }
    *s = orc_hsum8(acc) + summs;
}

/* Order-free "true value" of the same block arithmetic: exact integer block
 * sums, every fp32 scale product formed as the reference forms it, then
 * accumulated in double.  Tests use it to bound how far ANY fp32 summation
 * order (the reference's 8 lanes, a GPU warp tree) may sit from the exact sum.
 * mag returns sum_i |d_i * q_i| (+ |m*s| terms), the natural error scale. */
void orc_vec_dot_q4_q8_exact(int ggml_type, int n, double *s, double *mag, const void *vx, const void *vy) {
    const orc_block_q8_0 *y = (const orc_block_q8_0 *)vy;
    const int nb = n / ORC_QK;
    double acc = 0.0, m_acc = 0.0;
    for (int i = 0; i < nb; i++) {
        int lanes[8], tot = 0;
        if (ggml_type == 2) {
            const orc_block_q4_0 *x = (const orc_block_q4_0 *)vx;
            orc_lane_sums_q4(x[i].qs, y[i].qs, 8, lanes);
            for (int l = 0; l < 8; l++) tot += lanes[l];
            const float d = x[i].d * y[i].d;
            acc += (double)d * (double)tot;
            m_acc += fabs((double)d * (double)tot);
        } else {
            const orc_block_q4_1 *x = (const orc_block_q4_1 *)vx;
            orc_lane_sums_q4(x[i].qs, y[i].qs, 0, lanes);
            for (int l = 0; l < 8; l++) tot += lanes[l];
            const float d = x[i].d * y[i].d;
            acc += (double)d * (double)tot + (double)x[i].m * (double)y[i].s;
            m_acc += fabs((double)d * (double)tot) + fabs((double)x[i].m * (double)y[i].s);
        }
    }
    *s = acc;
    if (mag) *mag = m_acc;
}

/* ------------------------------------------------------------------------- */
/* ggml_compute_forward_mul_mat_q_f32, lib/ggml.c:7928-8176 (non-BLAS path):  */
/*   INIT    :8105-8119  every src1 row (K floats) -> q8_0 row in wdata       */
/*   COMPUTE :8125-8163  dst[n*M + m] = vec_dot_q(K, W row m, q8 row n)       */
/* W: M rows of K/32 blocks (row stride = K/32*block bytes), X: N rows of K   */
/* floats, dst: N rows of M floats.  2-D case only (ne02 = ne03 = 1), which   */
/* is all Model::eval emits for quantized weights (lib/llama.cpp:328-465).    */
/* ------------------------------------------------------------------------- */
int orc_mul_mat_q_f32(int ggml_type, int M, int K, int N, const void *W, const float *X, float *dst) {
    const int bb = orc_block_bytes(ggml_type);
    if ((ggml_type != 2 && ggml_type != 3) || K % ORC_QK != 0) return -1;
    const size_t wrow = (size_t)(K / ORC_QK) * bb;
    const size_t qrow = (size_t)(K / ORC_QK) * sizeof(orc_block_q8_0);
    char *wdata = (char *)malloc(qrow * (size_t)N);
    if (!wdata) return -2;
    for (int n = 0; n < N; n++) orc_quantize_row_q8_0(X + (size_t)n * K, wdata + qrow * n, K);
    for (int m = 0; m < M; m++) {
        const char *wr = (const char *)W + wrow * m;
        for (int n = 0; n < N; n++) {
            float *out = dst + (size_t)n * M + m;
            if (ggml_type == 2) orc_vec_dot_q4_0_q8_0(K, out, wr, wdata + qrow * n);
            else                orc_vec_dot_q4_1_q8_0(K, out, wr, wdata + qrow * n);
        }
    }
    free(wdata);
    return 0;
}

/* Same product with the order-free accumulation (double). */
int orc_mul_mat_q_exact(int ggml_type, int M, int K, int N, const void *W, const float *X, double *dst, double *mag) {
    const int bb = orc_block_bytes(ggml_type);
    if ((ggml_type != 2 && ggml_type != 3) || K % ORC_QK != 0) return -1;
    const size_t wrow = (size_t)(K / ORC_QK) * bb;
    const size_t qrow = (size_t)(K / ORC_QK) * sizeof(orc_block_q8_0);
    char *wdata = (char *)malloc(qrow * (size_t)N);
    if (!wdata) return -2;
    for (int n = 0; n < N; n++) orc_quantize_row_q8_0(X + (size_t)n * K, wdata + qrow * n, K);
    for (int m = 0; m < M; m++) {
        const char *wr = (const char *)W + wrow * m;
        for (int n = 0; n < N; n++)
            orc_vec_dot_q4_q8_exact(ggml_type, K, dst + (size_t)n * M + m, mag ? mag + (size_t)n * M + m : NULL,
                                    wr, wdata + qrow * n);
    }
    free(wdata);
    return 0;
}

/* get_rows on a quantized matrix: lib/ggml.c:8333-8360. */
int orc_get_rows_q(int ggml_type, int K, int n_ids, const void *W, const int32_t *ids, float *dst) {
    const int bb = orc_block_bytes(ggml_type);
    if (ggml_type != 2 && ggml_type != 3) return -1;
    const size_t wrow = (size_t)(K / ORC_QK) * bb;
    for (int i = 0; i < n_ids; i++) {
        const char *wr = (const char *)W + wrow * (size_t)ids[i];
        if (ggml_type == 2) orc_dequantize_row_q4_0(wr, dst + (size_t)i * K, K);
        else                orc_dequantize_row_q4_1(wr, dst + (size_t)i * K, K);
    }
    return 0;
}
