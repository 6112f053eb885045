// fl_decode_dev.cuh -- device helpers shared by the fused decode kernels (fl_decode_kernels.cu) and the
// persistent token kernel (fl_token_kernel.cu): prepared q8_0 activations, the q4 block dot, and the
// closed-form tiling of a CTA's row slice.
#pragma once
#include "fl_common.cuh"
#include "fl_cuda.h"

// ---- shared with fl_quant_kernels.cu (duplicated small device helpers) ----------------------------
struct fd_yprep {
    uint32_t ye[4], yo[4];
    float d, s;
    int c;
};
template <int TYPE>
__device__ __forceinline__ void fd_prep_y(const fl_block_q8_0 *yb, fd_yprep &p) {
    const uint32_t *q = (const uint32_t *)yb->qs;
    int sum = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t a = q[2 * j], b = q[2 * j + 1];
        p.ye[j] = __byte_perm(a, b, 0x6420);
        p.yo[j] = __byte_perm(a, b, 0x7531);
        sum = fl_dp4a_ss(0x01010101u, a, sum);
        sum = fl_dp4a_ss(0x01010101u, b, sum);
    }
    p.d = yb->d;
    p.s = yb->s;
    p.c = (TYPE == FL_TYPE_Q4_0) ? -8 * sum : 0;
}
__device__ __forceinline__ int fd_block_isum(const uint32_t w[4], const fd_yprep &p) {
    int lo = p.c, hi = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        lo = fl_dp4a_us(w[j] & 0x0F0F0F0Fu, p.ye[j], lo);
        hi = fl_dp4a_us(w[j] & 0xF0F0F0F0u, p.yo[j], hi);
    }
    return lo + (hi >> 4);
}
template <int TYPE>
__device__ __forceinline__ void fd_block(const uint8_t *blk, const fd_yprep &yp, float &acc, float &accm) {
    uint32_t w[4];
    float dx;
    if (TYPE == FL_TYPE_Q4_0) {
        const uint32_t *bw = (const uint32_t *)blk;
        dx = __uint_as_float(bw[0]);
        w[0] = bw[1]; w[1] = bw[2]; w[2] = bw[3]; w[3] = bw[4];
    } else {
        const uint2 *bw = (const uint2 *)blk;
        const uint2 dm = bw[0], q01 = bw[1], q23 = bw[2];
        dx = __uint_as_float(dm.x);
        accm = __fmaf_rn(__uint_as_float(dm.y), yp.s, accm);
        w[0] = q01.x; w[1] = q01.y; w[2] = q23.x; w[3] = q23.y;
    }
    const int isum = fd_block_isum(w, yp);
    acc = __fmaf_rn(__fmul_rn(dx, yp.d), (float)isum, acc);
}

#define FD_NBL 4

// The CTA's slice [r0, r1) of the concatenated row space, cut per matrix ("segment") into tiles of
// at most R rows that never straddle a matrix.  Everything is closed-form, so the producer can start
// issuing copies a few cycles after launch and nobody builds a table.
struct fd_slice {
    int f0, f1, f2;          // per segment: first owned row, relative to the segment
    int n0, n1, n2;          // rows owned
    int t0, t1;              // cumulative tile counts after segment 0 and 1
    int ntiles;
};
__device__ __forceinline__ void fd_seg_span(int r0, int r1, int sbase, int rows_sg, int &first, int &n) {
    const int lo = max(r0, sbase), hi = min(r1, sbase + rows_sg);
    first = lo - sbase;
    n = max(0, hi - lo);
}
__device__ __forceinline__ fd_slice fd_make_slice(const fl_mv_args &A, int mtot, int R) {
    fd_slice sl;
    const int half = mtot / 2;             // even split points keep rope pairs in one CTA
    const int r0 = 2 * (int)(((long)half * blockIdx.x) / gridDim.x);
    const int r1 = 2 * (int)(((long)half * (blockIdx.x + 1)) / gridDim.x);
    const int m0 = A.seg_rows[0], m1 = A.nseg > 1 ? A.seg_rows[1] : 0, m2 = A.nseg > 2 ? A.seg_rows[2] : 0;
    fd_seg_span(r0, r1, 0, m0, sl.f0, sl.n0);
    fd_seg_span(r0, r1, m0, m1, sl.f1, sl.n1);
    fd_seg_span(r0, r1, m0 + m1, m2, sl.f2, sl.n2);
    sl.t0 = (sl.n0 + R - 1) / R;
    sl.t1 = sl.t0 + (sl.n1 + R - 1) / R;
    sl.ntiles = sl.t1 + (sl.n2 + R - 1) / R;
    return sl;
}
__device__ __forceinline__ void fd_tile_of(const fd_slice &sl, int R, int t, int &seg, int &row0, int &rows) {
    seg = (t < sl.t0) ? 0 : (t < sl.t1) ? 1 : 2;
    const int j = t - (seg == 0 ? 0 : seg == 1 ? sl.t0 : sl.t1);
    const int first = seg == 0 ? sl.f0 : seg == 1 ? sl.f1 : sl.f2;
    const int n = seg == 0 ? sl.n0 : seg == 1 ? sl.n1 : sl.n2;
    row0 = first + j * R;
    rows = min(R, n - j * R);
}


// ---- P*V of the decode attention: out[d] = sum_j p_j * V[d][j] -------------------------------------
// Canonical summation order, shared by k_attn_decode and the token kernel so that both give the same bits
// however many threads they put on one output dimension: the float4 groups of the V row are dealt round-robin
// to FD_PV_SUBS subsequences; each subsequence keeps four lane accumulators (by position mod 4) and folds them
// as (a0+a1)+(a2+a3); subsequence 0 appends the n_pos % 4 tail; the partials are then added in index order.
#define FD_PV_SUBS 16
template <int NS>
__device__ __forceinline__ void fd_pv_partials(const float *v, const float *sc, int n_pos, int sub0, float *out /* [NS] */) {
    const int n4 = n_pos >> 2;
    float a[NS][4];
#pragma unroll
    for (int u = 0; u < NS; u++) a[u][0] = a[u][1] = a[u][2] = a[u][3] = 0.f;
    for (int i0 = 0; i0 < n4; i0 += FD_PV_SUBS) {
        float4 vv[NS];
#pragma unroll
        for (int u = 0; u < NS; u++) {
            const int i = i0 + sub0 + u;
            vv[u] = (i < n4) ? __ldcg((const float4 *)(v + 4 * i)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < NS; u++) {
            const int i = i0 + sub0 + u;
            if (i < n4) {
                a[u][0] = __fmaf_rn(vv[u].x, sc[4 * i + 0], a[u][0]);
                a[u][1] = __fmaf_rn(vv[u].y, sc[4 * i + 1], a[u][1]);
                a[u][2] = __fmaf_rn(vv[u].z, sc[4 * i + 2], a[u][2]);
                a[u][3] = __fmaf_rn(vv[u].w, sc[4 * i + 3], a[u][3]);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < NS; u++) {
        float acc = __fadd_rn(__fadd_rn(a[u][0], a[u][1]), __fadd_rn(a[u][2], a[u][3]));
        if (sub0 + u == 0)
            for (int j = 4 * n4; j < n_pos; j++) acc = __fmaf_rn(__ldcg(v + j), sc[j], acc);
        out[u] = acc;
    }
}
// thread handles subsequences [sub0, sub0 + NS) of output dimension `col` (column of the [FD_PV_SUBS][ncols] partial table)
template <int NS>
__device__ __forceinline__ void fd_pv_do(const float *v, const float *sc, int n_pos, int sub0, float *part, int ncols, int col) {
    float o[NS];
    fd_pv_partials<NS>(v, sc, n_pos, sub0, o);
#pragma unroll
    for (int u = 0; u < NS; u++) part[(sub0 + u) * ncols + col] = o[u];
}
template <int MAXNS>
__device__ __forceinline__ void fd_pv_store_partials(const float *v, const float *sc, int n_pos, int sub0, int ns, float *part, int ncols, int col) {
    if (ns == 1) fd_pv_do<1>(v, sc, n_pos, sub0, part, ncols, col);
    else if (ns == 2) fd_pv_do<2>(v, sc, n_pos, sub0, part, ncols, col);
    else if (ns == 4 || MAXNS == 4) fd_pv_do<4>(v, sc, n_pos, sub0, part, ncols, col);
    else if (ns == 8) fd_pv_do<(MAXNS >= 8 ? 8 : 4)>(v, sc, n_pos, sub0, part, ncols, col);
    else fd_pv_do<(MAXNS >= 16 ? 16 : 4)>(v, sc, n_pos, sub0, part, ncols, col);
}
__device__ __forceinline__ float fd_pv_combine(const float *part, int ncols, int col) {
    float acc = part[col];
#pragma unroll
    for (int u = 1; u < FD_PV_SUBS; u++) acc = __fadd_rn(acc, part[u * ncols + col]);
    return acc;
}
