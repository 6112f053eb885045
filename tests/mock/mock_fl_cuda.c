/*
 * mock_fl_cuda.c -- TEST-ONLY stand-in for libfl_cuda.so that runs on the CPU.
 *
 * Purpose: exercise the HOST logic of libggml_b200 / the drop-in pyfastllama.so (arena mirrors,
 * upload policy, graph walking, output copy-back) in the GPU-less build container.  "Device" memory
 * is malloc'ed host memory kept strictly separate from the caller's arenas, so a missing upload or
 * copy-back shows up exactly as it would on the GPU.  It is never shipped, never loaded by the
 * product (tests copy libggml_b200.so next to it so that $ORIGIN resolves here), and its arithmetic
 * comes from the oracle (oracle/q4_oracle.c) plus plain scalar loops.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fl_cuda.h"

/* oracle (linked in) */
void orc_quantize_row_q8_0(const float *x, void *vy, int k);
void orc_quantize_row_q4_0(const float *x, void *vy, int k);
void orc_quantize_row_q4_1(const float *x, void *vy, int k);
void orc_dequantize_row_q4_0(const void *vx, float *y, int k);
void orc_dequantize_row_q4_1(const void *vx, float *y, int k);
void orc_vec_dot_q4_0_q8_0(int n, float *s, const void *vx, const void *vy);
void orc_vec_dot_q4_1_q8_0(int n, float *s, const void *vx, const void *vy);
void orc_quantize_row_q4_0_simd(const float *x, void *vy, int k);
void orc_quantize_row_q4_1_simd(const float *x, void *vy, int k);
float orc_vec_dot_f32(int n, const float *x, const float *y);
int orc_add_q_f32(int ggml_type, int rows, int k, const void *src0, const float *src1, void *dst);

static int g_ready = 0;
static uint64_t g_launches = 0;
static char g_err[256] = "";
static uint16_t tab_silu[1 << 16], tab_exp[1 << 16];

static float h2f(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1F, m = h & 0x3FF, bits;
    if (e == 0) {
        if (!m) bits = sign;
        else { int k = -1; do { k++; m <<= 1; } while (!(m & 0x400)); bits = sign | ((uint32_t)(127 - 15 - k) << 23) | ((m & 0x3FF) << 13); }
    } else if (e == 31) bits = sign | 0x7F800000u | (m << 13);
    else bits = sign | ((e + 112) << 23) | (m << 13);
    float f; memcpy(&f, &bits, 4); return f;
}
#include <immintrin.h>
static uint16_t f2h(float f) { return (uint16_t)_cvtss_sh(f, 0); }

int fl_init(int device) {
    (void)device;
    if (!g_ready) {
        for (int i = 0; i < (1 << 16); i++) {
            float f = h2f((uint16_t)i);
            tab_silu[i] = f2h(f / (1.0f + expf(-f)));
            tab_exp[i] = f2h(expf(f));
        }
        g_ready = 1;
    }
    return 0;
}
void fl_shutdown(void) { g_ready = 0; }
int fl_is_initialized(void) { return g_ready; }
const char *fl_last_error(void) { return g_err; }
int fl_device_props(char *name, int n, int *sm, size_t *hbm, int *maj, int *mnr) {
    if (name && n > 0) snprintf(name, (size_t)n, "mock-cpu");
    if (sm) *sm = 1; if (hbm) *hbm = 0; if (maj) *maj = 0; if (mnr) *mnr = 0;
    return 0;
}
void *fl_stream(void) { return NULL; }
uint64_t fl_launch_count(void) { return g_launches; }
void *fl_dev_malloc(size_t b) { void *p = malloc(b ? b : 1); if (p) memset(p, 0xA5, b); return p; }   /* poison */
int fl_dev_free(void *p) { free(p); return 0; }
int fl_dev_memset(void *p, int v, size_t b) { memset(p, v, b); return 0; }
int fl_h2d(void *d, const void *s, size_t b) { memcpy(d, s, b); return 0; }
int fl_d2h(void *d, const void *s, size_t b) { memcpy(d, s, b); return 0; }
int fl_d2d(void *d, const void *s, size_t b) { memmove(d, s, b); return 0; }
int fl_d2d_2d(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h) { for (size_t r = 0; r < h; r++) memmove((char *)d + r * dp, (const char *)s + r * sp, w); return 0; }
int fl_sync(void) { return 0; }
void *fl_host_alloc_pinned(size_t b) { return malloc(b ? b : 1); }
int fl_host_free_pinned(void *p) { free(p); return 0; }
void *fl_event_create(void) { return malloc(8); }
int fl_event_destroy(void *e) { free(e); return 0; }
int fl_event_record(void *e) { (void)e; return 0; }
int fl_event_sync(void *e) { (void)e; return 0; }
int fl_event_elapsed_ms(void *a, void *b, float *ms) { (void)a; (void)b; *ms = 0.f; return 0; }
int fl_dev_fill_normal(float *p, size_t n, uint64_t seed, float std) { (void)seed; for (size_t i = 0; i < n; i++) p[i] = std * (float)((i * 2654435761u) % 1000) / 1000.f; return 0; }

int fl_dev_quantize_q8_0(const float *x, size_t stride, void *y, int k, int nrows) {
    g_launches++;
    for (int r = 0; r < nrows; r++) orc_quantize_row_q8_0((const float *)((const char *)x + r * stride), (char *)y + (size_t)r * (k / 32) * 40, k);
    return 0;
}
int fl_dev_quantize_q4(int type, const float *x, void *y, int k, int nrows) {
    g_launches++;
    const int bb = type == 2 ? 20 : 24;
    for (int r = 0; r < nrows; r++) {
        if (type == 2) orc_quantize_row_q4_0(x + (size_t)r * k, (char *)y + (size_t)r * (k / 32) * bb, k);
        else orc_quantize_row_q4_1(x + (size_t)r * k, (char *)y + (size_t)r * (k / 32) * bb, k);
    }
    return 0;
}
int fl_dev_mul_mat_q(int type, const void *W, size_t wrs, int M, int K, const void *Y, int N, float *dst, size_t drs, int impl) {
    (void)impl; g_launches++;
    for (int n = 0; n < N; n++)
        for (int m = 0; m < M; m++) {
            float *o = dst + (size_t)n * drs + m;
            const void *wr = (const char *)W + (size_t)m * wrs, *yr = (const char *)Y + (size_t)n * (K / 32) * 40;
            if (type == 2) orc_vec_dot_q4_0_q8_0(K, o, wr, yr); else orc_vec_dot_q4_1_q8_0(K, o, wr, yr);
        }
    return 0;
}
typedef struct mock_graph_fwd mock_graph_fwd;
static int mock_capturing(void);
static void mock_record_deq(int type, const void *W, size_t wrs, int K, const int32_t *ids, int n, float *dst, size_t drs);
int fl_dev_dequantize_rows(int type, const void *W, size_t wrs, int K, const int32_t *ids, int n, float *dst, size_t drs) {
    if (mock_capturing()) { mock_record_deq(type, W, wrs, K, ids, n, dst, drs); return 0; }
    g_launches++;
    for (int i = 0; i < n; i++) {
        const void *wr = (const char *)W + (size_t)(ids ? ids[i] : i) * wrs;
        if (type == 2) orc_dequantize_row_q4_0(wr, dst + (size_t)i * drs, K); else orc_dequantize_row_q4_1(wr, dst + (size_t)i * drs, K);
    }
    return 0;
}
int fl_dev_time_mul_mat_q(int t, const void *W, size_t a, int M, int K, const void *Y, int N, float *d, size_t b, int i, int it, size_t f, float *ms) {
    (void)t; (void)W; (void)a; (void)M; (void)K; (void)Y; (void)N; (void)d; (void)b; (void)i; (void)it; (void)f; *ms = 0; return 0;
}
int fl_dev_time_mul_mat_q_rot(int t, const void *W, size_t a, int M, int K, const void *Y, int N, float *d, size_t b, int i, int it, size_t f, size_t cs, int nc, float *ms) {
    (void)t; (void)W; (void)a; (void)M; (void)K; (void)Y; (void)N; (void)d; (void)b; (void)i; (void)it; (void)f; (void)cs; (void)nc; *ms = 0; return 0;
}

/* host-buffer entry points */
int fl_quantize_rows_q8_0(const float *x, void *y, int k, int n) { return fl_dev_quantize_q8_0(x, (size_t)k * 4, y, k, n); }
int fl_quantize_row_q8_0(const float *x, void *y, int k) { return fl_quantize_rows_q8_0(x, y, k, 1); }
int fl_quantize_rows_q4(int t, const float *x, void *y, int k, int n) { return fl_dev_quantize_q4(t, x, y, k, n); }
int fl_dev_quantize_q4_simd(int type, const float *x, void *y, int k, int nrows) {
    g_launches++;
    const int bb = type == 2 ? 20 : 24;
    for (int r = 0; r < nrows; r++) {
        if (type == 2) orc_quantize_row_q4_0_simd(x + (size_t)r * k, (char *)y + (size_t)r * (k / 32) * bb, k);
        else orc_quantize_row_q4_1_simd(x + (size_t)r * k, (char *)y + (size_t)r * (k / 32) * bb, k);
    }
    return 0;
}
int fl_quantize_rows_q4_simd(int t, const float *x, void *y, int k, int n) { return fl_dev_quantize_q4_simd(t, x, y, k, n); }
int fl_dev_add_q_f32(int type, const void *W, size_t wrs, int M, int K, const float *X, size_t xrs, void *dst, size_t drs) {
    g_launches++;
    for (int r = 0; r < M; r++)
        if (orc_add_q_f32(type, 1, K, (const char *)W + (size_t)r * wrs, X + (size_t)r * xrs, (char *)dst + (size_t)r * drs)) return -1;
    return 0;
}
int fl_dev_mul_mat_f32_ref(const float *A, size_t lda, int Ma, const float *B, size_t ldb, int Mb, int K, float *out, size_t ldo) {
    g_launches++;
    for (int j = 0; j < Mb; j++)
        for (int i = 0; i < Ma; i++) out[(size_t)j * ldo + i] = orc_vec_dot_f32(K, A + (size_t)i * lda, B + (size_t)j * ldb);
    return 0;
}
int fl_dequantize_rows_q4(int t, const void *x, float *y, int k, int n) { return fl_dev_dequantize_rows(t, x, (size_t)(k / 32) * (t == 2 ? 20 : 24), k, NULL, n, y, (size_t)k); }
int fl_vec_dot_q4_q8(int t, int n, float *s, const void *x, const void *y) { if (t == 2) orc_vec_dot_q4_0_q8_0(n, s, x, y); else orc_vec_dot_q4_1_q8_0(n, s, x, y); return 0; }
int fl_mul_mat_q_f32(int t, int M, int K, int N, const void *W, const float *X, float *dst) {
    void *q = malloc((size_t)(K / 32) * 40 * (size_t)(N ? N : 1));
    fl_dev_quantize_q8_0(X, (size_t)K * 4, q, K, N);
    fl_dev_mul_mat_q(t, W, (size_t)(K / 32) * (t == 2 ? 20 : 24), M, K, q, N, dst, (size_t)M, 0);
    free(q);
    return 0;
}
int fl_get_rows_q(int t, int K, int n, const void *W, int total, const int32_t *ids, float *dst) { (void)total; return fl_dev_dequantize_rows(t, W, (size_t)(K / 32) * (t == 2 ? 20 : 24), K, ids, n, dst, (size_t)K); }

/* ---- strided ops ------------------------------------------------------------------------------ */
static inline char *at(const fl_view *v, int64_t i0, int64_t i1, int64_t i2, int64_t i3) {
    return (char *)v->data + i0 * v->nb[0] + i1 * v->nb[1] + i2 * v->nb[2] + i3 * v->nb[3];
}
static inline char *lin(const fl_view *v, int64_t n) {
    int64_t i0 = n % v->ne[0]; n /= v->ne[0];
    int64_t i1 = n % v->ne[1]; n /= v->ne[1];
    int64_t i2 = n % v->ne[2]; int64_t i3 = n / v->ne[2];
    return at(v, i0, i1, i2, i3);
}
static inline int64_t nel(const fl_view *v) { return v->ne[0] * v->ne[1] * v->ne[2] * v->ne[3]; }

int fl_dev_rms_norm(const fl_view *s, const fl_view *d) {
    g_launches++;
    for (int64_t i3 = 0; i3 < s->ne[3]; i3++) for (int64_t i2 = 0; i2 < s->ne[2]; i2++) for (int64_t i1 = 0; i1 < s->ne[1]; i1++) {
        const float *x = (const float *)at(s, 0, i1, i2, i3); float *y = (float *)at(d, 0, i1, i2, i3);
        double sum = 0; for (int64_t i = 0; i < s->ne[0]; i++) sum += (double)(x[i] * x[i]);
        float mean = (float)(sum / (double)s->ne[0]); float sc = 1.0f / sqrtf(mean + 1e-6f);
        for (int64_t i = 0; i < s->ne[0]; i++) y[i] = x[i] * sc;
    }
    return 0;
}
int fl_dev_add(const fl_view *a, const fl_view *b, const fl_view *d) { g_launches++; for (int64_t i = 0; i < nel(d); i++) *(float *)lin(d, i) = *(float *)lin(a, i) + *(float *)lin(b, i); return 0; }
int fl_dev_mul(const fl_view *a, const fl_view *b, const fl_view *d) { g_launches++; for (int64_t i = 0; i < nel(d); i++) *(float *)lin(d, i) = *(float *)lin(a, i) * *(float *)lin(b, i); return 0; }
int fl_dev_repeat(const fl_view *s, const fl_view *d) {
    g_launches++;
    for (int64_t i3 = 0; i3 < d->ne[3]; i3++) for (int64_t i2 = 0; i2 < d->ne[2]; i2++) for (int64_t i1 = 0; i1 < d->ne[1]; i1++) for (int64_t i0 = 0; i0 < d->ne[0]; i0++)
        *(float *)at(d, i0, i1, i2, i3) = *(float *)at(s, i0 % s->ne[0], i1 % s->ne[1], i2 % s->ne[2], i3 % s->ne[3]);
    return 0;
}
int fl_dev_scale(const fl_view *t, float v) { g_launches++; for (int64_t i = 0; i < nel(t); i++) *(float *)lin(t, i) *= v; return 0; }
int fl_dev_silu(const fl_view *s, const fl_view *d) { g_launches++; for (int64_t i = 0; i < nel(d); i++) *(float *)lin(d, i) = h2f(tab_silu[f2h(*(float *)lin(s, i))]); return 0; }
int fl_dev_diag_mask_inf(const fl_view *t, int n_past) {
    g_launches++;
    for (int64_t k = 0; k < t->ne[2] * t->ne[3]; k++) for (int64_t j = 0; j < t->ne[1]; j++) for (int64_t i = n_past; i < t->ne[0]; i++)
        if (i > n_past + j) *(float *)((char *)t->data + k * t->nb[2] + j * t->nb[1] + i * t->nb[0]) = -INFINITY;
    return 0;
}
int fl_dev_soft_max(const fl_view *t) {
    g_launches++;
    for (int64_t r = 0; r < t->ne[1] * t->ne[2] * t->ne[3]; r++) {
        float *p = (float *)((char *)t->data + r * t->nb[1]); float mx = -INFINITY; double sum = 0;
        for (int64_t i = 0; i < t->ne[0]; i++) if (p[i] > mx) mx = p[i];
        for (int64_t i = 0; i < t->ne[0]; i++) { if (p[i] == -INFINITY) p[i] = 0; else { float v = h2f(tab_exp[f2h(p[i] - mx)]); sum += v; p[i] = v; } }
        float inv = (float)(1.0 / sum); for (int64_t i = 0; i < t->ne[0]; i++) p[i] *= inv;
    }
    return 0;
}
int fl_dev_rope(const fl_view *t, int n_past, int n_dims, int mode) {
    g_launches++;
    const float ts = powf(10000.0f, -2.0f / n_dims);
    for (int64_t i3 = 0; i3 < t->ne[3]; i3++) for (int64_t i2 = ((mode & 1) ? n_past : 0); i2 < t->ne[2]; i2++) for (int64_t i1 = 0; i1 < t->ne[1]; i1++) {
        float theta = (float)((mode & 1) ? i2 : n_past + i2);
        for (int i0 = 0; i0 < n_dims; i0 += 2) {
            float c = cosf(theta), s = sinf(theta); theta *= ts;
            float *p0 = (float *)at(t, (mode & 2) ? i0 / 2 : i0, i1, i2, i3), *p1 = (float *)at(t, (mode & 2) ? i0 / 2 + n_dims / 2 : i0 + 1, i1, i2, i3);
            float x0 = *p0, x1 = *p1; *p0 = fmaf(x0, c, -(x1 * s)); *p1 = fmaf(x0, s, x1 * c);   /* the reference build contracts these (GNU mode) */
        }
    }
    return 0;
}
int fl_dev_cpy_f32(const fl_view *s, const fl_view *d) { g_launches++; for (int64_t i = 0; i < nel(s); i++) *(float *)lin(d, i) = *(float *)lin(s, i); return 0; }
int fl_dev_mul_mat_f32(const fl_view *a, const fl_view *b, const fl_view *d) {
    g_launches++;
    const int64_t K = a->ne[0];
    float *xs = malloc(sizeof(float) * (K ? K : 1)), *ys = malloc(sizeof(float) * (K ? K : 1));
    for (int64_t i3 = 0; i3 < d->ne[3]; i3++) for (int64_t i2 = 0; i2 < d->ne[2]; i2++) for (int64_t i1 = 0; i1 < d->ne[1]; i1++) for (int64_t i0 = 0; i0 < d->ne[0]; i0++) {
        for (int64_t k = 0; k < K; k++) { xs[k] = *(const float *)at(a, k, i0, i2, i3); ys[k] = *(const float *)at(b, k, i1, i2, i3); }
        *(float *)at(d, i0, i1, i2, i3) = orc_vec_dot_f32((int)K, xs, ys);          /* ggml_vec_dot_f32's order */
    }
    free(xs); free(ys);
    return 0;
}

/* ---- fused decode step + graph capture (recorded and replayed, like a CUDA graph) ---------------- */
/* ---- tensor-parallel plumbing of the mock: collectives are delegated to a callback the test installs
 * (torch.distributed / gloo on the "device" buffers, which are host memory here) */
typedef void (*mock_coll_cb)(int kind, void *send, void *recv, size_t n);   /* kind 0 = allreduce in place, 1 = allgather */
static mock_coll_cb g_coll = NULL;
static int g_rank = 0, g_world = 1;
void fl_mock_set_collective(mock_coll_cb cb, int rank, int world) { g_coll = cb; g_rank = rank; g_world = world; }
int fl_comm_unique_id(void *out) { memset(out, 0, 128); return 0; }
int fl_comm_init(int rank, int world, const void *id) { (void)id; g_rank = rank; g_world = world; return 0; }
int fl_comm_rank(void) { return g_rank; }
int fl_comm_world(void) { return g_world; }

enum { OP_MV = 1, OP_ATTN = 2, OP_DEQ = 3, OP_ALLREDUCE = 10, OP_ALLGATHER = 11, OP_PLAN = 20 };
typedef struct {
    int kind;
    fl_mv_args mv;
    struct { const float *q, *k, *v; float *out; const int *n_past; int n_embd, n_head, hd, n_ctx; float scale; int out_ll, out_seq, n_out_peer; float *out_peer[7]; } at;
    int swiglu;                      /* w1|w3 step fused with the silu*mul of the next one (as the token kernel does) */
    struct { int type; const void *W; size_t wrs; int K; const int32_t *ids; int n; float *dst; size_t drs; } dq;
    struct { float *send, *recv; size_t n; } co;
    void *plan;
} mock_op;
typedef struct { mock_op *ops; int n, cap; volatile unsigned *ll_count; int n_ll; } mock_graph;
static mock_graph *g_capture = NULL;

/* peer-mapped buffers between the CPU processes of a gloo test: POSIX shared memory, one segment per rank, named by
 * $FL_MOCK_SESSION (unset: no peer memory, the caller keeps the collective path) */
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
static void mock_barrier(void) { float z = 0.f; if (g_coll && g_world > 1) g_coll(0, &z, &z, 1); }
int fl_comm_shared_alloc(size_t bytes, void **peers) {
    const char *sess = getenv("FL_MOCK_SESSION");
    if (!sess || g_world <= 1 || !g_coll) return -1;
    char name[128];
    snprintf(name, sizeof(name), "/flmock_%s_%d", sess, g_rank);
    int fd = shm_open(name, O_CREAT | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)bytes) != 0) return -1;
    peers[g_rank] = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    memset(peers[g_rank], 0, bytes);
    mock_barrier();                                              /* every segment exists and is zeroed */
    for (int r = 0; r < g_world; r++) {
        if (r == g_rank) continue;
        snprintf(name, sizeof(name), "/flmock_%s_%d", sess, r);
        fd = shm_open(name, O_RDWR, 0600);
        if (fd < 0) return -1;
        peers[r] = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
    }
    mock_barrier();
    snprintf(name, sizeof(name), "/flmock_%s_%d", sess, g_rank);
    shm_unlink(name);                                            /* everybody has it mapped; the name can go */
    return 0;
}
static void record(const mock_op *op) {
    if (g_capture->n == g_capture->cap) { g_capture->cap = g_capture->cap ? 2 * g_capture->cap : 64; g_capture->ops = realloc(g_capture->ops, sizeof(mock_op) * g_capture->cap); }
    g_capture->ops[g_capture->n++] = *op;
}
/* LL vectors: {value, epoch} words.  Polling is real: in the gloo tests the producer is another process writing through POSIX
 * shared memory. */
static float ll_wait(const float *slot, int i, unsigned e) {
    const volatile unsigned *w = (const volatile unsigned *)slot;
    long spins = 0;
    while (w[2 * i + 1] != e) { if (++spins > 2000000000L) { fprintf(stderr, "mock: LL word %d never reached epoch %u (has %u)\n", i, e, w[2 * i + 1]); abort(); } }
    __sync_synchronize();
    unsigned v = w[2 * i]; float f; memcpy(&f, &v, 4); return f;
}
static void ll_store(float *slot, int i, float v, unsigned e, float *const *peers, int n_peers) {
    unsigned u; memcpy(&u, &v, 4);
    volatile unsigned *w = (volatile unsigned *)slot;
    w[2 * i] = u; __sync_synchronize(); w[2 * i + 1] = e;
    for (int p = 0; p < n_peers; p++) { volatile unsigned *q = (volatile unsigned *)peers[p]; q[2 * i] = u; __sync_synchronize(); q[2 * i + 1] = e; }
}
static void run_mv(const fl_mv_args *a, int swiglu, unsigned e_in, unsigned e_out) {
    const int K = a->K, nb = K / 32, bb = a->type == 2 ? 20 : 24;
    const size_t rstride = a->row_stride_bytes ? a->row_stride_bytes : (size_t)nb * bb;
    float *v = malloc(sizeof(float) * K), *xin = malloc(sizeof(float) * K);
    for (int i = 0; i < K; i++) {
        const float t = a->x_ll ? ll_wait(a->x, i, e_in) : a->x[i];
        xin[i] = a->xadd ? t + a->xadd[i] : t;
    }
    if (a->sum_out) memcpy(a->sum_out, xin, sizeof(float) * K);
    if (a->pro == FL_PRO_RMSNORM) {
        double sum = 0; for (int i = 0; i < K; i++) sum += (double)(xin[i] * xin[i]);
        float mean = (float)(sum / (double)K), sc = 1.0f / sqrtf(mean + 1e-6f);
        for (int i = 0; i < K; i++) { v[i] = a->gamma[i] * (xin[i] * sc); if (a->normed_out) a->normed_out[i] = v[i]; }
    } else if (a->pro == FL_PRO_SILUMUL) {
        for (int i = 0; i < K; i++) v[i] = h2f(tab_silu[f2h(xin[i])]) * a->b[i];
    } else memcpy(v, xin, sizeof(float) * K);
    free(xin);
    void *q8 = malloc((size_t)nb * 40);
    orc_quantize_row_q8_0(v, q8, K);
    const int n_past = a->epi == FL_EPI_QKV ? *a->n_past : 0;
    float *seg_out[3] = {NULL, NULL, NULL};
    for (int sg = 0; sg < a->nseg; sg++) {
        float *tmp = seg_out[sg] = malloc(sizeof(float) * a->seg_rows[sg]);
        for (int r = 0; r < a->seg_rows[sg]; r++) {
            const void *wr = (const char *)a->seg_w[sg] + (size_t)r * rstride;
            if (a->type == 2) orc_vec_dot_q4_0_q8_0(K, tmp + r, wr, q8); else orc_vec_dot_q4_1_q8_0(K, tmp + r, wr, q8);
        }
    }
    if (swiglu) {
        for (int r = 0; r < a->seg_rows[0]; r++) {
            const float o = h2f(tab_silu[f2h(seg_out[0][r])]) * seg_out[1][r];
            if (a->out_ll) ll_store(a->seg_dst[0], r, o, e_out, a->dst_peer, a->n_dst_peer); else a->seg_dst[0][r] = o;
        }
    } else for (int sg = 0; sg < a->nseg; sg++) {
        float *tmp = seg_out[sg];
        if (a->epi == FL_EPI_QKV) {
            const int hd = a->head_dim; const float ts = powf(10000.0f, -2.0f / hd);
            for (int r = 0; r < a->seg_rows[sg]; r += 2) {
                float x0 = tmp[r], x1 = tmp[r + 1];
                if (sg < 2) {
                    float theta = (float)n_past; for (int i = 0; i < (r % hd) / 2; i++) theta *= ts;
                    float c = cosf(theta), s = sinf(theta), y0 = fmaf(x0, c, -(x1 * s)), y1 = fmaf(x0, s, x1 * c);
                    float *o = sg == 0 ? a->seg_dst[0] + r : a->kcache + (size_t)n_past * a->n_embd + r;
                    o[0] = y0; o[1] = y1;
                } else { a->vcache[(size_t)r * a->n_ctx + n_past] = x0; a->vcache[(size_t)(r + 1) * a->n_ctx + n_past] = x1; }
            }
        } else for (int r = 0; r < a->seg_rows[sg]; r++) {
            const float o = a->epi == FL_EPI_RESADD ? tmp[r] + (a->res_ll ? a->res[2 * r] : a->res[r]) : tmp[r];
            if (a->out_ll) ll_store(a->seg_dst[sg], r, o, e_out, a->dst_peer, a->n_dst_peer); else a->seg_dst[sg][r] = o;
        }
    }
    for (int sg = 0; sg < a->nseg; sg++) free(seg_out[sg]);
    free(q8); free(v);
}
static void run_attn(const mock_op *o, unsigned e_out) {
    const int hd = o->at.hd, n_pos = *o->at.n_past + 1;
    float *p = malloc(sizeof(float) * n_pos);
    for (int h = 0; h < o->at.n_head; h++) {
        float mx = -INFINITY; double sum = 0;
        for (int j = 0; j < n_pos; j++) { const float acc = orc_vec_dot_f32(hd, o->at.k + (size_t)j * o->at.n_embd + h * hd, o->at.q + h * hd); p[j] = acc * o->at.scale; if (p[j] > mx) mx = p[j]; }
        for (int j = 0; j < n_pos; j++) { float v = h2f(tab_exp[f2h(p[j] - mx)]); sum += v; p[j] = v; }
        float inv = (float)(1.0 / sum);
        for (int j = 0; j < n_pos; j++) p[j] *= inv;
        for (int d = 0; d < hd; d++) {
            const float r = orc_vec_dot_f32(n_pos, o->at.v + ((size_t)h * hd + d) * o->at.n_ctx, p);
            if (o->at.out_ll) ll_store(o->at.out, h * hd + d, r, e_out, o->at.out_peer, o->at.n_out_peer); else o->at.out[h * hd + d] = r;
        }
    }
    free(p);
}
static void run_op(const mock_op *o) {
    g_launches++;
    if (o->kind == OP_PLAN) {
        const mock_graph *pg = o->plan;
        const unsigned base = pg->ll_count ? *pg->ll_count : 0u;
        for (int i = 0; i < pg->n; i++) {
            const mock_op *q = &pg->ops[i];
            if (q->kind == OP_MV) run_mv(&q->mv, q->swiglu, base + (unsigned)q->mv.x_seq + 1u, base + (unsigned)q->mv.out_seq + 1u);
            else if (q->kind == OP_ATTN) run_attn(q, base + (unsigned)q->at.out_seq + 1u);
        }
        if (pg->ll_count) *pg->ll_count = base + (unsigned)pg->n_ll;
        return;
    }
    if (o->kind == OP_MV) run_mv(&o->mv, 0, 0, 0);
    else if (o->kind == OP_ATTN) run_attn(o, 0);
    else if (o->kind == OP_ALLREDUCE) g_coll(0, o->co.send, o->co.recv, o->co.n);
    else if (o->kind == OP_ALLGATHER) g_coll(1, o->co.send, o->co.recv, o->co.n);
    else for (int i = 0; i < o->dq.n; i++) {
        const void *wr = (const char *)o->dq.W + (size_t)(o->dq.ids ? o->dq.ids[i] : i) * o->dq.wrs;
        if (o->dq.type == 2) orc_dequantize_row_q4_0(wr, o->dq.dst + (size_t)i * o->dq.drs, o->dq.K); else orc_dequantize_row_q4_1(wr, o->dq.dst + (size_t)i * o->dq.drs, o->dq.K);
    }
}
int fl_dev_mv_fused_supported(int type, int K, int mtot) { return (type == 2 || type == 3) && K % 64 == 0 && mtot >= 2 && mtot % 2 == 0; }
int fl_dev_rope_table(int n_dims, int n_pos) { (void)n_dims; (void)n_pos; return 0; }
int fl_dev_mv_fused(const fl_mv_args *a) { mock_op o; memset(&o, 0, sizeof(o)); o.kind = OP_MV; o.mv = *a; if (g_capture) record(&o); else run_op(&o); return 0; }
int fl_dev_attn_decode(const float *q, const float *k, const float *v, float *out, const int *n_past, int n_embd, int n_head, int head_dim, int n_ctx, float scale) {
    mock_op o; memset(&o, 0, sizeof(o)); o.kind = OP_ATTN; o.at.q = q; o.at.k = k; o.at.v = v; o.at.out = out; o.at.n_past = n_past; o.at.n_embd = n_embd; o.at.n_head = n_head; o.at.hd = head_dim; o.at.n_ctx = n_ctx; o.at.scale = scale;
    if (g_capture) record(&o); else run_op(&o); return 0;
}
/* the persistent token kernel: the mock runs its steps one after another */
int fl_token_plan_create_ll(const fl_token_step *steps, int n, unsigned *epoch_counter, void **out) {
    mock_graph *pg = calloc(1, sizeof(mock_graph));
    pg->ops = calloc((size_t)n, sizeof(mock_op)); pg->n = pg->cap = n;
    pg->ll_count = epoch_counter;
    for (int i = 0; i < n; i++) {
        mock_op *o = &pg->ops[i];
        if (steps[i].kind == 0) {
            o->kind = OP_MV; o->mv = steps[i].mv; o->swiglu = steps[i].mv.swiglu;
            if (o->mv.out_ll && o->mv.out_seq + 1 > pg->n_ll) pg->n_ll = o->mv.out_seq + 1;
            if (o->mv.x_ll && o->mv.x_seq + 1 > pg->n_ll) pg->n_ll = o->mv.x_seq + 1;
        } else {
            o->kind = OP_ATTN; o->at.q = steps[i].q; o->at.k = steps[i].kcache; o->at.v = steps[i].vcache; o->at.out = steps[i].out; o->at.n_past = steps[i].n_past;
            o->at.n_embd = steps[i].k_row_stride; o->at.n_head = steps[i].n_head; o->at.hd = steps[i].head_dim; o->at.n_ctx = steps[i].n_ctx; o->at.scale = steps[i].scale;
            o->at.out_ll = steps[i].out_ll; o->at.out_seq = steps[i].out_seq; o->at.n_out_peer = steps[i].n_out_peer;
            for (int r = 0; r < 7; r++) o->at.out_peer[r] = steps[i].out_peer[r];
            if (o->at.out_ll && o->at.out_seq + 1 > pg->n_ll) pg->n_ll = o->at.out_seq + 1;
        }
    }
    /* the token kernel's SwiGLU fusion (fl_token_kernel.cu, plan creation): w1|w3 followed by silu(.)*(.) over exactly their outputs */
    for (int i = 0; i + 1 < n; i++) {
        mock_op *p0 = &pg->ops[i], *p1 = &pg->ops[i + 1];
        if (p0->kind != OP_MV || p1->kind != OP_MV || p0->swiglu) continue;
        const fl_mv_args *a = &p0->mv;
        if (a->nseg == 2 && a->epi == FL_EPI_STORE && a->seg_rows[0] == a->seg_rows[1] && p1->mv.pro == FL_PRO_SILUMUL && p1->mv.x == a->seg_dst[0] &&
            p1->mv.b == a->seg_dst[1] && p1->mv.K == a->seg_rows[0] && p1->mv.xadd == NULL) {
            p0->swiglu = 1; p1->mv.pro = FL_PRO_PLAIN; p1->mv.b = NULL;
        }
    }
    if (pg->n_ll > 0 && !epoch_counter) { free(pg->ops); free(pg); return -1; }
    *out = pg; return 0;
}
int fl_token_plan_create(const fl_token_step *steps, int n, void **out) { return fl_token_plan_create_ll(steps, n, NULL, out); }
int fl_token_plan_launch(void *plan) { mock_op o; memset(&o, 0, sizeof(o)); o.kind = OP_PLAN; o.plan = plan; if (g_capture) record(&o); else run_op(&o); return 0; }
int fl_token_plan_profile(void *plan, unsigned long long *out, size_t n, int *c) { (void)plan; (void)out; (void)n; *c = 0; return -1; }
int fl_token_plan_profile2(void *plan, unsigned *out, size_t n) { (void)plan; (void)out; (void)n; return -1; }
int fl_token_plan_error(void *plan) { (void)plan; return 0; }

int fl_token_plan_destroy(void *plan) { mock_graph *pg = plan; if (pg) { free(pg->ops); free(pg); } return 0; }
int fl_graph_begin_capture(void) { g_capture = calloc(1, sizeof(mock_graph)); return 0; }
int fl_graph_end_capture(void **out) { *out = g_capture; g_capture = NULL; return 0; }
int fl_graph_launch(void *ge) { mock_graph *g = ge; for (int i = 0; i < g->n; i++) run_op(&g->ops[i]); return 0; }
int fl_graph_destroy(void *ge) { mock_graph *g = ge; if (g) { free(g->ops); free(g); } return 0; }

static int mock_capturing(void) { return g_capture != NULL; }
static void mock_record_deq(int type, const void *W, size_t wrs, int K, const int32_t *ids, int n, float *dst, size_t drs) {
    mock_op o; memset(&o, 0, sizeof(o)); o.kind = OP_DEQ; o.dq.type = type; o.dq.W = W; o.dq.wrs = wrs; o.dq.K = K; o.dq.ids = ids; o.dq.n = n; o.dq.dst = dst; o.dq.drs = drs; record(&o);
}


int fl_comm_allreduce_f32(float *buf, size_t n) {
    if (g_world <= 1) return 0;
    mock_op o; memset(&o, 0, sizeof(o)); o.kind = OP_ALLREDUCE; o.co.send = buf; o.co.recv = buf; o.co.n = n;
    if (g_capture) record(&o); else run_op(&o); return 0;
}
int fl_comm_allgather_f32(const float *send, float *recv, size_t n) {
    if (g_world <= 1) { if (send != recv) memmove(recv, send, n * sizeof(float)); return 0; }
    mock_op o; memset(&o, 0, sizeof(o)); o.kind = OP_ALLGATHER; o.co.send = (float *)send; o.co.recv = recv; o.co.n = n;
    if (g_capture) record(&o); else run_op(&o); return 0;
}
int fl_dev_pack_cols(int type, const void *W, size_t wrs, int M, int blk0, int nblk, void *dst, size_t drs) {
    const int bb = type == 2 ? 20 : 24;
    for (int m = 0; m < M; m++) memcpy((char *)dst + (size_t)m * drs, (const char *)W + (size_t)m * wrs + (size_t)blk0 * bb, (size_t)nblk * bb);
    return 0;
}
