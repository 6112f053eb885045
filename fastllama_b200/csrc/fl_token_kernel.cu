// fl_token_kernel.cu -- the whole decode step of a LLaMA token as ONE persistent kernel.
//
// Why: with one kernel per matrix group the decode step is 160 launches of 3-18 us whose fixed costs
// (launch, barrier init, prologue, first-tile latency, drain) leave HBM idle ~70 % of the time
// (DESIGN.md section 4).  Here 148 CTAs (one per SM, co-resident by cooperative launch) walk a
// "program" of phases -- per layer: wq|wk|wv, attention, wo, w1|w3, w2; then the LM head -- separated
// by grid-wide barriers, and each CTA's producer lane streams the weight tiles of ALL phases through one
// mbarrier ring, running ahead of the consumers across phase boundaries: while the grid synchronises
// and the next activation vector is quantised, the next matrices are already landing in shared memory.
//
// Each phase does exactly what the corresponding k_mv_fused / k_attn_decode launch does (same prologue,
// block arithmetic, combine order, epilogue), so results are bit-identical to the multi-kernel path.
// Activations move between phases through L2: they are read with ld.global.cg (L1 is not coherent
// across SMs inside a kernel) and published with __threadfence() before the barrier arrive.
#include <cuda_fp16.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "fl_common.cuh"
#include "fl_decode.h"
#include "fl_decode_dev.cuh"
#include "fl_kernels.h"

#define TK_CW 16                 // consumer warps
#define TK_NT (TK_CW * 32)
#define TK_TG 4                  // tile groups; ring slot s always belongs to group s % 4 (S is a multiple of 4)
#define TK_WPG 4                 // warps per tile group = kparts * G
#define TK_RMAX 4
#define TK_THREADS (TK_NT + 32)

enum { TK_PH_MATVEC = 0, TK_PH_ATTN = 1 };

struct tk_phase {
    int kind;
    int nb, kparts, G, R, P, nfull, mtot;
    uint32_t row_bytes;
    fl_mv_args a;
    // attention
    const float *q, *kcache, *vcache;
    float *out;
    int k_row_stride, n_head, head_dim, n_ctx;
    float scale;
};

struct tk_params {
    const tk_phase *phases;
    int n_phases;
    unsigned *grid_bar;
    const uint16_t *exp_tab;
    unsigned long long *prof;      // optional: [n_phases][gridDim.x][4] globaltimer stamps of thread 0
    int S;
    uint32_t slot_bytes;
    uint32_t off_y, off_red, off_rowbuf, off_cnt, off_sc, off_stage0;
};

__device__ __forceinline__ unsigned long long tk_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void tk_bar_consumers(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(TK_NT) : "memory"); }

__device__ __forceinline__ void tk_grid_sync(unsigned *bar, unsigned target) {
    tk_bar_consumers(13);                               // every consumer warp of this CTA has finished the phase
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while (v < target);
        __threadfence();
    }
    tk_bar_consumers(13);
}

// ---- prologue: q8_0 activations of the phase into shared memory (same arithmetic as k_mv_fused) ----
__device__ __forceinline__ void tk_prologue(const fl_mv_args &A, int K, fl_block_q8_0 *ysm, double *red, int warp, int lane, int tid) {
    const int nvec = K >> 2;
    const float4 *x4 = (const float4 *)A.x;
    const float4 *xa4 = (const float4 *)A.xadd;
    auto load_x = [&](int idx) -> float4 {
        float4 v = __ldcg(x4 + idx);
        if (xa4) { const float4 w = __ldcg(xa4 + idx); v.x = __fadd_rn(v.x, w.x); v.y = __fadd_rn(v.y, w.y); v.z = __fadd_rn(v.z, w.z); v.w = __fadd_rn(v.w, w.w); }
        return v;
    };
    float scale = 1.0f;
    if (A.pro == FL_PRO_RMSNORM) {
        double acc = 0.0;
#pragma unroll 4
        for (int i = tid; i < nvec; i += TK_NT) {
            const float4 v = load_x(i);
            acc += (double)__fmul_rn(v.x, v.x);
            acc += (double)__fmul_rn(v.y, v.y);
            acc += (double)__fmul_rn(v.z, v.z);
            acc += (double)__fmul_rn(v.w, v.w);
        }
        acc = fl_warp_sum_d(acc);
        if (lane == 0) red[warp] = acc;
        tk_bar_consumers(15);
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < TK_CW; w++) t += red[w];
            const float mean = (float)(t / (double)K);
            ((float *)(red + 16))[0] = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, 1e-6f)));
        }
        tk_bar_consumers(15);
        scale = ((float *)(red + 16))[0];
    }
    const float4 *g4 = (const float4 *)A.gamma, *b4 = (const float4 *)A.b;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    int i = warp * 32 + lane;
    float4 xv = (i < nvec) ? load_x(i) : zero4;
    float4 ov = zero4;
    if (A.pro == FL_PRO_RMSNORM) ov = (i < nvec) ? __ldg(g4 + i) : zero4;
    else if (A.pro == FL_PRO_SILUMUL) ov = (i < nvec) ? __ldcg(b4 + i) : zero4;
    for (int base = warp * 32; base < nvec; base += TK_NT) {
        const int inext = i + TK_NT;
        float4 xn = zero4, on = zero4;
        if (base + TK_NT < nvec) {
            xn = (inext < nvec) ? load_x(inext) : zero4;
            if (A.pro == FL_PRO_RMSNORM) on = (inext < nvec) ? __ldg(g4 + inext) : zero4;
            else if (A.pro == FL_PRO_SILUMUL) on = (inext < nvec) ? __ldcg(b4 + inext) : zero4;
        }
        float v[4] = {xv.x, xv.y, xv.z, xv.w};
        const float o[4] = {ov.x, ov.y, ov.z, ov.w};
        if (A.sum_out && blockIdx.x == 0 && i < nvec) ((float4 *)A.sum_out)[i] = xv;
        if (A.pro == FL_PRO_RMSNORM) {
#pragma unroll
            for (int c = 0; c < 4; c++) v[c] = __fmul_rn(o[c], __fmul_rn(v[c], scale));
            if (A.normed_out && blockIdx.x == 0 && i < nvec) ((float4 *)A.normed_out)[i] = make_float4(v[0], v[1], v[2], v[3]);
        } else if (A.pro == FL_PRO_SILUMUL) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const uint16_t h = __half_as_ushort(__float2half_rn(v[c]));
                v[c] = __fmul_rn(__half2float(__ushort_as_half(__ldg(A.silu_tab + h))), o[c]);
            }
        }
        float amax = fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3])));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
        const float d = __fdiv_rn(amax, 127.f);
        const float id = (amax != 0.0f) ? __fdiv_rn(127.f, amax) : 0.0f;
        int q[4], sum = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            q[c] = max(-128, min(127, __float2int_rn(__fmul_rn(v[c], id))));
            sum += q[c];
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        sum += __shfl_xor_sync(0xffffffffu, sum, 4);
        if (i < nvec) {
            fl_block_q8_0 *yb = ysm + (i >> 3);
            const uint32_t packed = (uint32_t)(q[0] & 0xFF) | ((uint32_t)(q[1] & 0xFF) << 8) | ((uint32_t)(q[2] & 0xFF) << 16) | ((uint32_t)(q[3] & 0xFF) << 24);
            ((uint32_t *)yb->qs)[i & 7] = packed;
            if ((i & 7) == 0) {
                yb->d = d;
                yb->s = __fmul_rn(d, (float)sum);
            }
        }
        xv = xn; ov = on; i = inext;
    }
    tk_bar_consumers(15);
}

// ---- main loop of a matvec phase for one consumer warp ------------------------------------------------
template <int TYPE, int NFULL>
__device__ __forceinline__ void tk_consume(const tk_phase &ph, const tk_params &prm, const fd_slice &sl, int T0, const fl_block_q8_0 *ysm,
                                           float *rowbuf, int *cnt, uint8_t *stage0, uint32_t bar0, int warp, int lane) {
    constexpr int BB = (TYPE == FL_TYPE_Q4_0) ? 20 : 24;
    const fl_mv_args &A = ph.a;
    const int S = prm.S, R = ph.R, kparts = ph.kparts, G = ph.G;
    const int tg = warp / TK_WPG, wl = warp % TK_WPG;
    const int p = wl % kparts, g = wl / kparts;
    const int b0 = p * ph.P;
    const int b1 = min(ph.nb, b0 + ph.P);

    fd_yprep yp[FD_NBL];
    bool valid[FD_NBL];
#pragma unroll
    for (int j = 0; j < FD_NBL; j++) {
        const int ib = b0 + lane + 32 * j;
        valid[j] = (j < NFULL) || ib < b1;
        if (valid[j]) {
            fd_prep_y<TYPE>(ysm + ib, yp[j]);
        } else {
            yp[j].d = 0.f; yp[j].s = 0.f; yp[j].c = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) { yp[j].ye[q] = 0; yp[j].yo[q] = 0; }
        }
    }
    const bool pair = (A.epi == FL_EPI_QKV);
    const bool staged = (kparts > 1) || pair;
    const int target = pair ? 2 * kparts : kparts;
    const int n_past = pair ? *A.n_past : 0;
    const int ntiles = sl.ntiles;
    int t = ((tg - (T0 & 3)) + 4) & 3;                 // first tile of this phase owned by the warp's tile group
    int T = T0 + t;
    int s = T % S;
    uint32_t par = (uint32_t)(T / S) & 1u;
    for (; t < ntiles; t += TK_TG) {
        int seg, row0, rows;
        fd_tile_of(sl, R, t, seg, row0, rows);
        fl_mbar_wait(bar0 + 8u * s, par);
        const uint8_t *tile = stage0 + (size_t)s * prm.slot_bytes;
        float *dseg = A.seg_dst[seg];
        for (int rr = g; rr < rows; rr += G) {
            const uint8_t *wrow = tile + (size_t)rr * ph.row_bytes + (size_t)(b0 + lane) * BB;
            float acc = 0.0f, accm = 0.0f;
#pragma unroll
            for (int j = 0; j < FD_NBL; j++) {
                if (j < NFULL) fd_block<TYPE>(wrow + (size_t)(32 * j) * BB, yp[j], acc, accm);
                else if (valid[j]) fd_block<TYPE>(wrow + (size_t)(32 * j) * BB, yp[j], acc, accm);
            }
            float tot = fl_warp_sum(acc);
            if (TYPE == FL_TYPE_Q4_1) tot = __fadd_rn(tot, fl_warp_sum(accm));
            if (lane == 0) {
                const int row = row0 + rr;
                if (!staged) {
                    dseg[row] = (A.epi == FL_EPI_RESADD) ? __fadd_rn(tot, __ldcg(A.res + row)) : tot;
                } else {
                    volatile float *rb = rowbuf + (size_t)s * TK_RMAX * 4;
                    rb[rr * kparts + p] = tot;
                    __threadfence_block();
                    const int gid = pair ? (rr >> 1) : rr;
                    const int old = atomicAdd(&cnt[s * TK_RMAX + gid], 1);
                    if (old == target - 1) {
                        cnt[s * TK_RMAX + gid] = 0;
                        __threadfence_block();
                        if (pair) {
                            const int ra = gid << 1;
                            float x0 = rb[ra * kparts], x1 = rb[(ra + 1) * kparts];
                            for (int q = 1; q < kparts; q++) { x0 = __fadd_rn(x0, rb[ra * kparts + q]); x1 = __fadd_rn(x1, rb[(ra + 1) * kparts + q]); }
                            const int r2 = row0 + ra;
                            if (seg < 2) {
                                const int ip = (r2 % A.head_dim) >> 1;
                                const float2 cs = ((const float2 *)A.rope_cs)[(size_t)n_past * (A.head_dim >> 1) + ip];
                                const float y0 = __fmaf_rn(x0, cs.x, -__fmul_rn(x1, cs.y));
                                const float y1 = __fmaf_rn(x0, cs.y, __fmul_rn(x1, cs.x));
                                float *o = (seg == 0) ? (dseg + r2) : (A.kcache + (size_t)n_past * A.n_embd + r2);
                                o[0] = y0; o[1] = y1;
                            } else {
                                A.vcache[(size_t)r2 * A.n_ctx + n_past] = x0;
                                A.vcache[(size_t)(r2 + 1) * A.n_ctx + n_past] = x1;
                            }
                        } else {
                            float tsum = rb[rr * kparts];
                            for (int q = 1; q < kparts; q++) tsum = __fadd_rn(tsum, rb[rr * kparts + q]);
                            dseg[row] = (A.epi == FL_EPI_RESADD) ? __fadd_rn(tsum, __ldcg(A.res + row)) : tsum;
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) fl_mbar_arrive(bar0 + 8u * (S + s));
        s += TK_TG;
        if (s >= S) { s -= S; par ^= 1u; }
    }
}

// ---- attention phase: one head, all consumer threads of the CTA ------------------------------------
__device__ __forceinline__ void tk_attention(const tk_phase &ph, const tk_params &prm, float *sc, double *red, int head, int warp, int lane, int tid) {
    float *redf = (float *)(red + 20);                // [16] floats; red[0..16] are the double partials
    float *part = sc + ph.n_ctx;                      // [TK_NT]
    const int hd = ph.head_dim;
    const int n_pos = *ph.a.n_past + 1;
    const float *q = ph.q + (size_t)head * hd;
    for (int j0 = warp * 4; j0 < n_pos; j0 += TK_CW * 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int e = lane; e < hd; e += 32) {
            const float qe = __ldcg(q + e);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = min(j0 + u, n_pos - 1);
                acc[u] = __fmaf_rn(__ldcg(ph.kcache + (size_t)j * ph.k_row_stride + (size_t)head * hd + e), qe, acc[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const float a = fl_warp_sum(acc[u]);
            if (lane == 0 && j0 + u < n_pos) sc[j0 + u] = __fmul_rn(a, ph.scale);
        }
    }
    tk_bar_consumers(12);
    float mx = -INFINITY;
    for (int j = tid; j < n_pos; j += TK_NT) mx = fmaxf(mx, sc[j]);
    mx = fl_warp_max(mx);
    if (lane == 0) redf[warp] = mx;
    tk_bar_consumers(12);
    mx = redf[0];
    for (int w = 1; w < TK_CW; w++) mx = fmaxf(mx, redf[w]);
    double sum = 0.0;
    for (int j = tid; j < n_pos; j += TK_NT) {
        const uint16_t hh = __half_as_ushort(__float2half_rn(__fsub_rn(sc[j], mx)));
        const float e = __half2float(__ushort_as_half(__ldg(prm.exp_tab + hh)));
        sc[j] = e;
        sum += (double)e;
    }
    sum = fl_warp_sum_d(sum);
    if (lane == 0) red[warp] = sum;
    tk_bar_consumers(12);
    double tot = 0.0;
    for (int w = 0; w < TK_CW; w++) tot += red[w];
    const float inv = (float)(1.0 / tot);
    tk_bar_consumers(12);
    for (int j = tid; j < n_pos; j += TK_NT) sc[j] = __fmul_rn(sc[j], inv);
    tk_bar_consumers(12);
    const int npt = 256 / hd;                           // same split of the positions as k_attn_decode (256 threads): same summation order
    if (npt >= 1 && tid < npt * hd) {
        const int d = tid % hd, sub = tid / hd;
        const float *v = ph.vcache + ((size_t)head * hd + d) * ph.n_ctx;
        const int n4 = n_pos >> 2;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int i = sub; i < n4; i += npt) {
            const float4 vv = __ldcg((const float4 *)(v + 4 * i));
            a0 = __fmaf_rn(vv.x, sc[4 * i + 0], a0);
            a1 = __fmaf_rn(vv.y, sc[4 * i + 1], a1);
            a2 = __fmaf_rn(vv.z, sc[4 * i + 2], a2);
            a3 = __fmaf_rn(vv.w, sc[4 * i + 3], a3);
        }
        float acc = __fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2, a3));
        if (sub == 0)
            for (int j = 4 * n4; j < n_pos; j++) acc = __fmaf_rn(__ldcg(v + j), sc[j], acc);
        part[tid] = acc;
    }
    tk_bar_consumers(12);
    if (tid < hd) {
        float acc = part[tid];
        for (int u = 1; u < npt; u++) acc = __fadd_rn(acc, part[tid + u * hd]);
        ph.out[(size_t)head * hd + tid] = acc;
    }
}

template <int TYPE>
__device__ __forceinline__ void tk_consume_dispatch(const tk_phase &ph, const tk_params &prm, const fd_slice &sl, int T0, const fl_block_q8_0 *ysm,
                                                    float *rowbuf, int *cnt, uint8_t *stage0, uint32_t bar0, int warp, int lane) {
    switch (ph.nfull) {
        case 4: tk_consume<TYPE, 4>(ph, prm, sl, T0, ysm, rowbuf, cnt, stage0, bar0, warp, lane); break;
        case 3: tk_consume<TYPE, 3>(ph, prm, sl, T0, ysm, rowbuf, cnt, stage0, bar0, warp, lane); break;
        case 2: tk_consume<TYPE, 2>(ph, prm, sl, T0, ysm, rowbuf, cnt, stage0, bar0, warp, lane); break;
        default: tk_consume<TYPE, 0>(ph, prm, sl, T0, ysm, rowbuf, cnt, stage0, bar0, warp, lane); break;
    }
}

__global__ void __launch_bounds__(TK_THREADS, 1) k_decode_token(const tk_params prm) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = (uint64_t *)smem;
    fl_block_q8_0 *ysm = (fl_block_q8_0 *)(smem + prm.off_y);
    double *red = (double *)(smem + prm.off_red);            // 32 doubles
    float *rowbuf = (float *)(smem + prm.off_rowbuf);        // [S][TK_RMAX][4]
    int *cnt = (int *)(smem + prm.off_cnt);                  // [S][TK_RMAX]
    float *sc = (float *)(smem + prm.off_sc);                // attention: [n_ctx] + [TK_NT]
    uint8_t *stage0 = smem + prm.off_stage0;
    const int S = prm.S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = fl_smem_u32(bars);

    if (warp == TK_CW) {
        // ------------------------------ producer: all phases, as far ahead as the ring allows ------------------------------
        if (lane == 0) {
            for (int s = 0; s < S; s++) {
                fl_mbar_init(bar0 + 8u * s, 1);
                fl_mbar_init(bar0 + 8u * (S + s), TK_WPG);
            }
            fl_mbar_fence_init();
        }
        __syncwarp();
        asm volatile("bar.arrive 14, %0;" ::"r"(TK_NT + 32) : "memory");
        if (lane == 0) {
            const uint64_t pol = fl_policy_evict_first();
            int s = 0;
            uint32_t par = 1;
            for (int pi = 0; pi < prm.n_phases; pi++) {
                const tk_phase &ph = prm.phases[pi];
                if (ph.kind != TK_PH_MATVEC) continue;
                const fd_slice sl = fd_make_slice(ph.a, ph.mtot, ph.R);
                for (int t = 0; t < sl.ntiles; t++) {
                    int seg, row0, rows;
                    fd_tile_of(sl, ph.R, t, seg, row0, rows);
                    fl_mbar_wait(bar0 + 8u * (S + s), par);
                    const uint32_t bytes = (uint32_t)rows * ph.row_bytes;
                    const uint8_t *src = (const uint8_t *)ph.a.seg_w[seg] + (size_t)row0 * ph.row_bytes;
                    fl_mbar_expect_tx(bar0 + 8u * s, bytes);
                    fl_bulk_g2s_hint(fl_smem_u32(stage0 + (size_t)s * prm.slot_bytes), src, bytes, bar0 + 8u * s, pol);
                    if (++s == S) { s = 0; par ^= 1u; }
                }
            }
        }
        return;
    }

    // ------------------------------ consumers ------------------------------
    const int tid = threadIdx.x;
    for (int i = tid; i < S * TK_RMAX; i += TK_NT) cnt[i] = 0;
    asm volatile("bar.sync 14, %0;" ::"r"(TK_NT + 32) : "memory");     // mbarriers initialised
    int T0 = 0;
    unsigned epoch = 0;
    for (int pi = 0; pi < prm.n_phases; pi++) {
        const tk_phase &ph = prm.phases[pi];
        unsigned long long *pr = (prm.prof && tid == 0) ? prm.prof + ((size_t)pi * gridDim.x + blockIdx.x) * 4 : nullptr;
        if (pr) pr[0] = tk_now();
        if (pi > 0) {
            epoch++;
            tk_grid_sync(prm.grid_bar, epoch * gridDim.x);               // results of phase pi-1 are visible everywhere
        }
        if (pr) pr[1] = tk_now();
        if (ph.kind == TK_PH_ATTN) {
            if ((int)blockIdx.x < ph.n_head) tk_attention(ph, prm, sc, red, blockIdx.x, warp, lane, tid);
            if (pr) pr[2] = pr[3] = tk_now();
            continue;
        }
        tk_prologue(ph.a, ph.nb * 32, ysm, red, warp, lane, tid);
        if (pr) pr[2] = tk_now();
        const fd_slice sl = fd_make_slice(ph.a, ph.mtot, ph.R);
        if (ph.a.type == FL_TYPE_Q4_0) tk_consume_dispatch<FL_TYPE_Q4_0>(ph, prm, sl, T0, ysm, rowbuf, cnt, stage0, bar0, warp, lane);
        else                           tk_consume_dispatch<FL_TYPE_Q4_1>(ph, prm, sl, T0, ysm, rowbuf, cnt, stage0, bar0, warp, lane);
        T0 += sl.ntiles;
        if (pr) pr[3] = tk_now();
    }
}

// =================================================================================================
// host side
// =================================================================================================
struct fl_token_plan_impl {
    tk_phase *d_phases = nullptr;
    unsigned *d_bar = nullptr;
    unsigned long long *d_prof = nullptr;
    tk_params prm;
    size_t smem = 0;
    int n_kernels = 0;
};

static int tk_geometry(tk_phase &ph, size_t &tile_bytes) {
    const fl_mv_args &a = ph.a;
    const int bb = fl_block_bytes(a.type);
    const int nb = a.K / 32;
    const size_t row_bytes = a.row_stride_bytes ? a.row_stride_bytes : (size_t)nb * bb;
    FL_REQUIRE(bb > 0 && a.K > 0 && a.K % 32 == 0 && row_bytes % 16 == 0, "token kernel: unsupported matrix K=%d", a.K);
    int kparts = 1;
    while (kparts * 128 < nb) kparts *= 2;
    FL_REQUIRE(kparts <= TK_WPG, "token kernel: K=%d needs %d K-slices (max %d)", a.K, kparts, TK_WPG);
    const int P = (nb + kparts - 1) / kparts;
    const int last = nb - (kparts - 1) * P;
    FL_REQUIRE(last > 0, "token kernel: K=%d splits badly", a.K);
    int nfull = std::min(P, last) / 32;
    if (nfull > FD_NBL) nfull = FD_NBL;
    if (nfull == 1) nfull = 0;
    int G = TK_WPG / kparts;
    int R = std::max(2, G);                       // even tiles keep rope pairs together
    if (R > TK_RMAX) R = TK_RMAX;
    int mtot = 0;
    for (int i = 0; i < a.nseg; i++) {
        FL_REQUIRE(a.seg_rows[i] > 0 && a.seg_rows[i] % 2 == 0 && ((uintptr_t)a.seg_w[i] & 15) == 0, "token kernel: bad segment %d", i);
        mtot += a.seg_rows[i];
    }
    ph.kind = TK_PH_MATVEC;
    ph.nb = nb; ph.kparts = kparts; ph.G = G; ph.R = R; ph.P = P; ph.nfull = nfull; ph.mtot = mtot; ph.row_bytes = (uint32_t)row_bytes;
    tile_bytes = (size_t)R * row_bytes;
    return 0;
}

int flk_token_plan_create(const fl_token_step *steps, int n_steps, const uint16_t *silu_tab, const uint16_t *exp_tab, const void *rope_cs, void **out) {
    std::vector<tk_phase> phases((size_t)n_steps);
    size_t max_tile = 0, max_y = 0;
    int max_ctx = 0;
    for (int i = 0; i < n_steps; i++) {
        tk_phase &ph = phases[i];
        memset(&ph, 0, sizeof(ph));
        if (steps[i].kind == 1) {
            ph.kind = TK_PH_ATTN;
            ph.q = steps[i].q; ph.kcache = steps[i].kcache; ph.vcache = steps[i].vcache; ph.out = steps[i].out;
            ph.k_row_stride = steps[i].k_row_stride; ph.n_head = steps[i].n_head; ph.head_dim = steps[i].head_dim; ph.n_ctx = steps[i].n_ctx;
            ph.scale = steps[i].scale;
            ph.a.n_past = steps[i].n_past;
            FL_REQUIRE(ph.n_ctx % 4 == 0 && ph.head_dim <= 256 && 256 % ph.head_dim == 0 && ph.n_head <= 148, "token kernel: unsupported attention shape");
            max_ctx = std::max(max_ctx, ph.n_ctx);
        } else {
            ph.a = steps[i].mv;
            ph.a.silu_tab = silu_tab;
            ph.a.rope_cs = rope_cs;
            size_t tb = 0;
            if (tk_geometry(ph, tb)) return -1;
            max_tile = std::max(max_tile, tb);
            max_y = std::max(max_y, (size_t)ph.nb * 40);
        }
    }
    int dev = 0, sm = 0, optin = 0;
    FL_CUDA_OK(cudaGetDevice(&dev));
    FL_CUDA_OK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    FL_CUDA_OK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    fl_token_plan_impl *pl = new fl_token_plan_impl();
    tk_params &p = pl->prm;
    const size_t slot = (max_tile + 127) & ~(size_t)127;
    int S = 16;
    size_t off = 0;
    for (;; S -= 4) {
        if (S < 4) { delete pl; fl_set_error("token kernel: tiles of %zu bytes do not fit shared memory", slot); return -1; }
        p.off_y = ((size_t)(2 * S) * 8 + 127) & ~(size_t)127;
        p.off_red = (p.off_y + max_y + 127) & ~(size_t)127;
        p.off_rowbuf = (p.off_red + 32 * sizeof(double) + 127) & ~(size_t)127;
        p.off_cnt = (p.off_rowbuf + (size_t)S * TK_RMAX * 4 * sizeof(float) + 127) & ~(size_t)127;
        p.off_sc = (p.off_cnt + (size_t)S * TK_RMAX * sizeof(int) + 127) & ~(size_t)127;
        off = (p.off_sc + ((size_t)max_ctx + TK_NT) * sizeof(float) + 127) & ~(size_t)127;
        if (off + (size_t)S * slot <= (size_t)optin - 1024) break;
    }
    p.off_stage0 = (uint32_t)off;
    p.S = S;
    p.slot_bytes = (uint32_t)slot;
    p.n_phases = n_steps;
    p.exp_tab = exp_tab;
    pl->smem = off + (size_t)S * slot;
    FL_CUDA_OK(cudaMalloc((void **)&pl->d_phases, sizeof(tk_phase) * (size_t)n_steps));
    FL_CUDA_OK(cudaMemcpy(pl->d_phases, phases.data(), sizeof(tk_phase) * (size_t)n_steps, cudaMemcpyHostToDevice));
    FL_CUDA_OK(cudaMalloc((void **)&pl->d_bar, 256));
    p.phases = pl->d_phases;
    p.grid_bar = pl->d_bar;
    p.prof = nullptr;
    if (getenv("FASTLLAMA_B200_TOKEN_PROF")) {
        FL_CUDA_OK(cudaMalloc((void **)&pl->d_prof, sizeof(unsigned long long) * 4 * (size_t)n_steps * sm));
        FL_CUDA_OK(cudaMemset(pl->d_prof, 0, sizeof(unsigned long long) * 4 * (size_t)n_steps * sm));
        p.prof = pl->d_prof;
    }
    cudaFuncAttributes fa;
    FL_CUDA_OK(cudaFuncGetAttributes(&fa, k_decode_token));
    FL_CUDA_OK(cudaFuncSetAttribute(k_decode_token, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes));
    int per_sm = 0;
    FL_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_decode_token, TK_THREADS, pl->smem));
    if (per_sm < 1) { delete pl; fl_set_error("token kernel: one CTA per SM does not fit (smem %zu)", pl->smem); return -1; }
    pl->n_kernels = sm;
    *out = pl;
    return 0;
}

int flk_token_plan_launch(cudaStream_t st, void *plan) {
    fl_token_plan_impl *pl = (fl_token_plan_impl *)plan;
    FL_CUDA_OK(cudaMemsetAsync(pl->d_bar, 0, 4, st));
    void *args[] = {(void *)&pl->prm};
    // cooperative launch: all 148 CTAs are guaranteed co-resident, which the grid barrier needs
    FL_CUDA_OK(cudaLaunchCooperativeKernel((const void *)k_decode_token, dim3(pl->n_kernels), dim3(TK_THREADS), args, pl->smem, st));
    fl_count_launch();
    return 0;
}

// tooling: stamps of the last launch, [n_steps][n_ctas][4] nanoseconds (needs FASTLLAMA_B200_TOKEN_PROF at create time)
int flk_token_plan_profile(void *plan, unsigned long long *out, size_t max_words, int *n_ctas) {
    fl_token_plan_impl *pl = (fl_token_plan_impl *)plan;
    FL_REQUIRE(pl && pl->d_prof, "token plan was created without FASTLLAMA_B200_TOKEN_PROF");
    const size_t words = (size_t)4 * pl->prm.n_phases * pl->n_kernels;
    FL_REQUIRE(max_words >= words, "profile buffer too small (%zu words needed)", words);
    FL_CUDA_OK(cudaMemcpy(out, pl->d_prof, words * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    *n_ctas = pl->n_kernels;
    return 0;
}

int flk_token_plan_destroy(void *plan) {
    fl_token_plan_impl *pl = (fl_token_plan_impl *)plan;
    if (!pl) return 0;
    if (pl->d_phases) cudaFree(pl->d_phases);
    if (pl->d_bar) cudaFree(pl->d_bar);
    if (pl->d_prof) cudaFree(pl->d_prof);
    delete pl;
    return 0;
}
