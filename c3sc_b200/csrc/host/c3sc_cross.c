/* c3sc_cross.c -- host driver that asks the Bellman operator for fibers in BATCHES.
 *
 * What the reference does here: valuef_interp (src/valuefunc.c:603-767) hands bellman_vi /
 * bellman_pi to C3's ftapprox_cross, which calls the operator one fiber at a time.  C3 is not
 * vendored (README.md:17-39), so this file restates the published algorithm it implements --
 * alternating TT-cross with QR + maxvol pivoting (Oseledets & Tyrtyshnikov 2010; the C3 paper,
 * Gorodetsky et al. 2018, Alg. 3) -- with the reference's set-up:
 *   - start index sets from uniform_stride (src/util.c:995-1006, src/valuefunc.c:676-690)
 *   - ft_cross_args maxiter 5 (src/valuefunc.c:632), fixed ranks (the adapt == 0 branch, :732)
 *   - the result in the valuef_precompute_cores layout (src/valuefunc.c:165-189)
 * The one difference that matters for the GPU: every core step requests ALL r_k * r_{k+1} fibers of
 * the core in one call (c3sc_fiber_batch_fn), which is what lets stage 1 of the backup share the
 * core tiles between fibers.
 *
 * Plain C, no arithmetic of the backup itself: the operator is a callback (the GPU path through
 * c3sc_vi_batch / c3sc_pi_batch in c3sc_cross_run_vi / _pi; the parity tests plug the CPU oracle
 * into the same driver).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../../../include/c3sc_cross.h"

struct c3sc_cross {
    uint32_t d;
    uint64_t n[C3SC_MAXD], r[C3SC_MAXD + 1], nmax;
    /* left index sets I[k]: r[k] multi-indices over dims 0..k-1 (row-major, stride d);
       right index sets J[k]: r[k] multi-indices over dims k..d-1 (stored at their dims) */
    int32_t *I[C3SC_MAXD + 1], *J[C3SC_MAXD + 1];
    /* what goes to the operator and comes back (fiber descriptors, fiber values), kept between runs; page-locked
       when the operator is the GPU (c3sc_cross_pin_buffers): pageable memory crosses PCIe through staging copies */
    void *xbuf; size_t xcap; int xpinned, want_pinned;
};
static void exchange_free(c3sc_cross *c)
{
    if (c->xbuf) { if (c->xpinned) c3sc_host_free(c->xbuf); else free(c->xbuf); }
    c->xbuf = NULL; c->xcap = 0; c->xpinned = 0;
}
static void *exchange_get(c3sc_cross *c, size_t bytes)
{
    if (c->xbuf && c->xcap >= bytes && c->xpinned == c->want_pinned) return c->xbuf;
    exchange_free(c);
    if (c->want_pinned && c3sc_host_alloc(bytes, &c->xbuf) == C3SC_OK) c->xpinned = 1;
    else c->xbuf = malloc(bytes);
    c->xcap = c->xbuf ? bytes : 0;
    return c->xbuf;
}
int c3sc_cross_pin_buffers(c3sc_cross *c, int on)
{
    if (!c) return C3SC_EINVAL;
    c->want_pinned = on != 0;
    return C3SC_OK;
}

/* src/util.c:995-1006 */
static uint64_t uniform_stride(uint64_t N, uint64_t M)
{
    uint64_t stride = 1;
    if (M < 2) return 0;
    while (stride * (M - 1) < N - 1) stride++;
    return stride - 1;
}

int c3sc_cross_create(uint32_t d, const uint64_t *n, const uint64_t *ranks, c3sc_cross **out)
{
    if (!n || !ranks || !out || d < 1 || d > C3SC_MAXD) return C3SC_EINVAL;
    if (ranks[0] != 1 || ranks[d] != 1) return C3SC_EINVAL;
    c3sc_cross *c = (c3sc_cross *)calloc(1, sizeof *c);
    if (!c) return C3SC_EINVAL;
    c->d = d;
    for (uint32_t k = 0; k < d; k++) {
        c->n[k] = n[k];
        if (n[k] > c->nmax) c->nmax = n[k];
    }
    for (uint32_t k = 0; k <= d; k++) c->r[k] = ranks[k];
    /* a rank cannot exceed the size of either unfolding it indexes */
    for (uint32_t k = 1; k < d; k++) {
        uint64_t cap = c->r[k - 1] * n[k - 1];
        if (c->r[k] > cap) c->r[k] = cap;
    }
    for (uint32_t k = d - 1; k >= 1; k--) {
        uint64_t cap = c->r[k + 1] * n[k];
        if (c->r[k] > cap) c->r[k] = cap;
    }
    for (uint32_t k = 0; k <= d; k++) {
        c->I[k] = (int32_t *)calloc(c->r[k] * d + 1, sizeof(int32_t));
        c->J[k] = (int32_t *)calloc(c->r[k] * d + 1, sizeof(int32_t));
        if (!c->I[k] || !c->J[k]) { c3sc_cross_destroy(c); return C3SC_EINVAL; }
    }
    /* start sets: element j of a set takes node (stride_i * j) mod N_i in every dimension i it spans,
       stride_i = uniform_stride(N_i, rank) -- the "diagonal" start of src/valuefunc.c:676-690 */
    for (uint32_t k = 0; k <= d; k++)
        for (uint64_t j = 0; j < c->r[k]; j++)
            for (uint32_t i = 0; i < d; i++) {
                const uint64_t st = uniform_stride(n[i], c->r[k] < n[i] ? c->r[k] : n[i]);
                const int32_t node = (int32_t)(((st ? st : 1) * j) % n[i]);
                if (i < k) c->I[k][j * d + i] = node;
                else c->J[k][j * d + i] = node;
            }
    *out = c;
    return C3SC_OK;
}

void c3sc_cross_destroy(c3sc_cross *c)
{
    if (!c) return;
    for (uint32_t k = 0; k <= c->d; k++) { free(c->I[k]); free(c->J[k]); }
    exchange_free(c);
    free(c);
}

/* ranks + index sets of `src` (ValueF keeps isl / isr and hands copies to the next step, src/valuefunc.c:706-712) */
int c3sc_cross_copy(const c3sc_cross *src, c3sc_cross **out)
{
    if (!src || !out) return C3SC_EINVAL;
    c3sc_cross *c = (c3sc_cross *)calloc(1, sizeof *c);
    if (!c) return C3SC_EINVAL;
    c->d = src->d; c->nmax = src->nmax;
    memcpy(c->n, src->n, sizeof c->n);
    memcpy(c->r, src->r, sizeof c->r);
    for (uint32_t k = 0; k <= c->d; k++) {
        const size_t len = c->r[k] * c->d + 1;
        c->I[k] = (int32_t *)malloc(len * sizeof(int32_t));
        c->J[k] = (int32_t *)malloc(len * sizeof(int32_t));
        if (!c->I[k] || !c->J[k]) { c3sc_cross_destroy(c); return C3SC_EINVAL; }
        memcpy(c->I[k], src->I[k], len * sizeof(int32_t));
        memcpy(c->J[k], src->J[k], len * sizeof(int32_t));
    }
    *out = c;
    return C3SC_OK;
}

int c3sc_cross_ranks(const c3sc_cross *c, uint64_t *ranks)
{
    if (!c || !ranks) return C3SC_EINVAL;
    for (uint32_t k = 0; k <= c->d; k++) ranks[k] = c->r[k];
    return C3SC_OK;
}

/* the driver's index sets at bond k (ValueF keeps them as isl / isr between solver steps,
 * src/valuefunc.c:706-712): left[r_k * d] uses dims 0..k-1, right[r_k * d] dims k..d-1, others 0 */
int c3sc_cross_index_sets(const c3sc_cross *c, uint32_t k, int32_t *left, int32_t *right)
{
    if (!c || k > c->d) return C3SC_EINVAL;
    for (uint64_t j = 0; j < c->r[k]; j++)
        for (uint32_t i = 0; i < c->d; i++) {
            if (left) left[j * c->d + i] = i < k ? c->I[k][j * c->d + i] : 0;
            if (right) right[j * c->d + i] = i >= k ? c->J[k][j * c->d + i] : 0;
        }
    return C3SC_OK;
}

/* ---- small dense linear algebra (column-major) --------------------------------------------- */
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define C3SC_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))   /* same arithmetic, wider registers (no FMA) */
#else
#define C3SC_CLONES
#endif
/* Stride-1 kernels with a fixed association order (eight partial sums), so the compiler can keep them in
 * vector registers without -ffast-math and the result does not depend on the vector width it picks. */
C3SC_CLONES static double dot8(const double *x, const double *y, size_t n)
{
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        s0 += x[i] * y[i];         s1 += x[i + 1] * y[i + 1]; s2 += x[i + 2] * y[i + 2]; s3 += x[i + 3] * y[i + 3];
        s4 += x[i + 4] * y[i + 4]; s5 += x[i + 5] * y[i + 5]; s6 += x[i + 6] * y[i + 6]; s7 += x[i + 7] * y[i + 7];
    }
    for (; i < n; i++) s0 += x[i] * y[i];
    return ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
}
C3SC_CLONES static void axpy(double a, const double *restrict x, double *restrict y, size_t n)
{
    for (size_t i = 0; i < n; i++) y[i] += a * x[i];
}
/* y += a x, returning max |y| of the updated vector (same four-lane maximum as absmax) */
C3SC_CLONES static double axpy_absmax(double a, const double *restrict x, double *restrict y, size_t n)
{
    double m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    size_t i = 0;
    for (; i + 4 <= n; i += 4) {
        const double y0 = y[i] + a * x[i], y1 = y[i + 1] + a * x[i + 1], y2 = y[i + 2] + a * x[i + 2], y3 = y[i + 3] + a * x[i + 3];
        y[i] = y0; y[i + 1] = y1; y[i + 2] = y2; y[i + 3] = y3;
        const double b0 = fabs(y0), b1 = fabs(y1), b2 = fabs(y2), b3 = fabs(y3);
        m0 = b0 > m0 ? b0 : m0; m1 = b1 > m1 ? b1 : m1; m2 = b2 > m2 ? b2 : m2; m3 = b3 > m3 ? b3 : m3;
    }
    for (; i < n; i++) { y[i] += a * x[i]; const double b = fabs(y[i]); m0 = b > m0 ? b : m0; }
    m0 = m1 > m0 ? m1 : m0; m2 = m3 > m2 ? m3 : m2;
    return m2 > m0 ? m2 : m0;
}
C3SC_CLONES static double absmax(const double *x, size_t n)
{
    double m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    size_t i = 0;
    for (; i + 4 <= n; i += 4) {
        const double a = fabs(x[i]), b = fabs(x[i + 1]), c = fabs(x[i + 2]), e = fabs(x[i + 3]);
        m0 = a > m0 ? a : m0; m1 = b > m1 ? b : m1; m2 = c > m2 ? c : m2; m3 = e > m3 ? e : m3;
    }
    for (; i < n; i++) { const double a = fabs(x[i]); m0 = a > m0 ? a : m0; }
    m0 = m1 > m0 ? m1 : m0; m2 = m3 > m2 ? m3 : m2;
    return m2 > m0 ? m2 : m0;
}

/* Householder QR of A (m x n, m >= n): Q (m x n, explicit, orthonormal columns) overwrites A. */
static void qr_explicit_q(double *A, size_t m, size_t n, double *work /* n + m */)
{
    double *tau = work, *v = work + n;
    for (size_t k = 0; k < n; k++) {
        double *ak = A + k * m;
        const double nrm = sqrt(dot8(ak + k, ak + k, m - k));
        if (nrm == 0.0) { tau[k] = 0.0; continue; }
        const double alpha = ak[k] >= 0.0 ? -nrm : nrm;
        const double v0 = ak[k] - alpha;
        for (size_t i = k + 1; i < m; i++) ak[i] /= v0;
        tau[k] = -v0 / alpha;
        ak[k] = alpha;
        for (size_t j = k + 1; j < n; j++) {
            double *aj = A + j * m;
            const double sc = (aj[k] + dot8(ak + k + 1, aj + k + 1, m - k - 1)) * tau[k];
            aj[k] -= sc;
            axpy(-sc, ak + k + 1, aj + k + 1, m - k - 1);
        }
    }
    /* accumulate Q = H_0 .. H_{n-1} [I; 0] in place, last reflector first */
    for (size_t kk = n; kk-- > 0;) {
        double *ak = A + kk * m;
        v[kk] = 1.0;
        for (size_t i = kk + 1; i < m; i++) v[i] = ak[i];
        for (size_t i = 0; i < m; i++) ak[i] = 0.0;
        ak[kk] = 1.0;
        if (tau[kk] == 0.0) continue;
        for (size_t j = kk; j < n; j++) {
            double *aj = A + j * m;
            const double sc = dot8(v + kk, aj + kk, m - kk) * tau[kk];
            axpy(-sc, v + kk, aj + kk, m - kk);
        }
    }
}

/* ---- the pivoting step of one core: twin rows -> QR basis -> maxvol, by a team of threads ------------------------
 * One VI sweep is 2(d-1) of these steps in sequence, each on an unfolding of r N x r numbers between two operator
 * calls: with the backup on the GPU they are what a sweep costs.  The phases share one OpenMP parallel region
 * (the reference is an OpenMP library itself, src/bellman.c:1390-1404).  Rows are cut into LA_NB fixed blocks; a
 * thread owns a contiguous run of blocks and is the only one that ever writes their rows, from the unfolding of
 * the operator's values to the store of the cross core, so the data stays in its core's cache.  Everything that
 * couples rows (column norms, reflector dots, pivot searches) is reduced per BLOCK, published, and combined by
 * every thread in block order after a barrier: the numbers depend on LA_NB, never on the number of threads
 * (C3SC_HOST_THREADS, default min(8, omp_get_max_threads(), CPUs of the affinity mask)), and every thread takes
 * the same branches.  Barriers are what the step costs on a team, so each phase is arranged to need one per
 * column: the QR takes the dots of the pivot column with ALL columns in one pass (earlier reflectors: the Gram
 * matrix of the compact WY form; itself: the norm; later columns: the reflection) and forms Q = (I - V T V^T) E
 * row block by row block with no further exchange; the pivot searches publish, per block, the maximum together
 * with the block's first near-maximal row.  Published records are double-buffered: a fast thread may already
 * write the next one while a slow one reads. */
#ifdef _OPENMP
#include <omp.h>
#include <sched.h>
#include <stdatomic.h>
#endif
#ifndef LA_NB
#define LA_NB 16
#endif
#define RANK_EPS 1e-11
#define TIE_EPS 1e-8     /* see team_maxvol */
#define LA_NONE ((size_t)-1)
#define LA_SWL 3

/* start rows of maxvol, one column: maxima over the block's unused rows (all of them / those not withheld), the first
 * row reaching each maximum exactly and the first within TIE_EPS of it */
typedef struct { double a_max, e_max, a_candval, e_candval; size_t a_arg, e_arg, a_cand, e_cand; } la_erec;
/* swaps of maxvol: the block's largest eligible |B| and the (at most LA_SWL, ascending) columns within TIE_EPS of it */
typedef struct { double best; size_t nl, j[LA_SWL], cand[LA_SWL], arg[LA_SWL]; double cm[LA_SWL], candval[LA_SWL]; double pad[7]; } la_srec;   /* 3 cache lines */

typedef struct {
    size_t mcap, ncap, bscap; /* capacity: rows, columns, rows per block */
    size_t pstride;           /* doubles between the partial sums of two blocks: whole cache lines */
    size_t m, n, bs;          /* this step: rows, columns, rows per block */
    int threads;
    double *Q, *B;            /* m x n column-major: unfolding in / orthonormal basis out; cross core Q inv(Q[P]) out */
    size_t *P;                /* n pivot rows out */
    char *skip, *used;        /* m */
    double *part;             /* [2][LA_NB][ncap] partial sums */
    la_erec *erec;            /* [2][LA_NB] */
    la_srec *srec;            /* [2][LA_NB] */
    double *priv;             /* [threads][privlen] */
    size_t privlen;
    double *V1;               /* n x n: the top of the reflector matrix, shared copy */
    double *k1; int64_t *qk; uint32_t *slot; size_t slotcap;   /* twin rows */
    double *twin_scale;       /* [LA_NB] largest |entry| of a block's rows, a cache line each */
    /* where the unfolding comes from and where the core goes (NULL: w->Q is filled by the caller, w->B read by it) */
    const double *vals; double *core; size_t ldo, rk, N, rk1; int right;
    int rc;
#ifdef _OPENMP
    char pad0[64]; atomic_int bar_count;                      /* sense-reversing barrier of the team, a cache line each */
    char pad1[60]; atomic_int bar_sense;
    char pad2[60];
#endif
} la_ws;
typedef struct { int tid, nt, b0, b1, sense; } la_thr;       /* a thread of the team and the blocks it owns */

/* A step has ~60 barriers with a few microseconds of work between them: the team spins (libgomp's barrier parks
 * threads in the kernel under some team sizes, 50 us a time); a thread that has lost its CPU is waited for with
 * sched_yield. */
#ifdef LA_PROFILE
static double g_bar_wait[64]; static long g_bar_n[64];
static double now_s(void);
#endif
static inline void la_barrier(la_ws *w, la_thr *t)
{
#ifdef _OPENMP
    if (t->nt == 1) return;
#ifdef LA_PROFILE
    const double t0_ = now_s();
#endif
    const int s = !t->sense;
    t->sense = s;
    if (atomic_fetch_add_explicit(&w->bar_count, 1, memory_order_acq_rel) == t->nt - 1) {
        atomic_store_explicit(&w->bar_count, 0, memory_order_relaxed);
        atomic_store_explicit(&w->bar_sense, s, memory_order_release);
    } else {
        for (unsigned spins = 0; atomic_load_explicit(&w->bar_sense, memory_order_acquire) != s; spins++) {
#if defined(__x86_64__)
            if (spins < 20000) __builtin_ia32_pause();
#else
            if (spins < 20000) continue;
#endif
            else sched_yield();
        }
    }
#ifdef LA_PROFILE
    g_bar_wait[t->tid] += now_s() - t0_; g_bar_n[t->tid]++;
#endif
#else
    (void)w; (void)t;
#endif
}
#define LA_BARRIER() la_barrier(w, th)

static void la_ws_free(la_ws *w)
{
    if (!w) return;
    free(w->Q); free(w->B); free(w->P); free(w->skip); free(w->used); free(w->part); free(w->erec); free(w->srec); free(w->priv);
    free(w->V1); free(w->k1); free(w->qk); free(w->slot); free(w->twin_scale); free(w);
}

static int la_threads(void)
{
    const char *e = getenv("C3SC_HOST_THREADS");
    int t = e ? atoi(e) : 0;
#ifdef _OPENMP
    if (t <= 0) { t = omp_get_max_threads(); if (t > 8) t = 8; }
    if (t > LA_NB) t = LA_NB;
#ifdef __linux__
    {                                                   /* a spinning team must not outnumber the CPUs it may run on */
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0 && t > CPU_COUNT(&set)) t = CPU_COUNT(&set);
    }
#endif
#else
    t = 1;
#endif
    return t < 1 ? 1 : t;
}

static size_t la_block_rows(size_t m) { return ((m + LA_NB - 1) / LA_NB + 7) & ~(size_t)7; }

/* everything two threads may write side by side starts on a cache line of its own */
static void *la_alloc(size_t bytes) { return aligned_alloc(64, (bytes + 63) & ~(size_t)63); }

static la_ws *la_ws_create(size_t mcap, size_t ncap)
{
    la_ws *w = (la_ws *)la_alloc(sizeof *w);
    if (!w) return NULL;
    memset(w, 0, sizeof *w);
    w->mcap = mcap; w->ncap = ncap; w->bscap = la_block_rows(mcap); w->threads = la_threads();
    w->pstride = (ncap + 7) & ~(size_t)7;
    w->slotcap = 16;
    while (w->slotcap < 4 * mcap) w->slotcap <<= 1;
    /* per thread: 8 vectors of n, the Gram / T / M / inverse matrices (3 n^2), [Q[P] | I] (2 n^2), the column maxima of its
       blocks, a row block of reflectors */
    w->privlen = (8 * ncap + 5 * ncap * ncap + LA_NB * ncap + w->bscap * ncap + 7) & ~(size_t)7;
    w->Q = (double *)la_alloc(mcap * ncap * sizeof(double));
    w->B = (double *)la_alloc(mcap * ncap * sizeof(double));
    w->P = (size_t *)la_alloc(ncap * sizeof(size_t));
    w->skip = (char *)la_alloc(mcap + 1); w->used = (char *)la_alloc(mcap + 1);
    w->part = (double *)la_alloc(2 * LA_NB * w->pstride * sizeof(double));
    w->erec = (la_erec *)la_alloc(2 * LA_NB * sizeof(la_erec));
    w->srec = (la_srec *)la_alloc(2 * LA_NB * sizeof(la_srec));
    w->priv = (double *)la_alloc((size_t)w->threads * w->privlen * sizeof(double));
    w->V1 = (double *)la_alloc(ncap * ncap * sizeof(double));
    w->k1 = (double *)la_alloc(2 * mcap * sizeof(double));
    w->twin_scale = (double *)la_alloc(LA_NB * 8 * sizeof(double));
    w->qk = (int64_t *)la_alloc(mcap * sizeof(int64_t));
    w->slot = (uint32_t *)la_alloc(w->slotcap * sizeof(uint32_t));
    if (!w->Q || !w->B || !w->P || !w->skip || !w->used || !w->part || !w->erec || !w->srec || !w->priv || !w->V1 || !w->k1 ||
        !w->qk || !w->slot || !w->twin_scale) { la_ws_free(w); return NULL; }
    return w;
}

/* rows [lo, hi) of block blk that lie at or below row `from` */
static inline void la_rows(const la_ws *w, int blk, size_t from, size_t *lo, size_t *hi)
{
    size_t a = (size_t)blk * w->bs, b = a + w->bs;
    if (b > w->m) b = w->m;
    if (a < from) a = from;
    if (a > b) a = b;
    *lo = a; *hi = b;
}
static inline int la_mine(const la_ws *w, size_t row, const la_thr *th)
{
    const int blk = (int)(row / w->bs);
    return blk >= th->b0 && blk < th->b1;
}
#define PART(buf, blk, j) part[((size_t)(buf) * LA_NB + (size_t)(blk)) * pstride + (j)]

/* own rows of the unfolding from the operator's values vals[(a + b*rk) * ldo + j]:
 * left  unfolding: rows (a,j) = a + j*rk, columns b;  right unfolding: rows (j,b) = j + b*N, columns a */
static void team_unfold(la_ws *w, la_thr *th)
{
    if (!w->vals) return;
    const size_t m = w->m, rk = w->rk, N = w->N, rk1 = w->rk1, ldo = w->ldo;
    const double *vals = w->vals;
    double *Q = w->Q;
    for (int blk = th->b0; blk < th->b1; blk++) {
        size_t lo, hi;
        la_rows(w, blk, 0, &lo, &hi);
        if (!w->right) {
            for (size_t b = 0; b < rk1; b++)
                for (size_t i = lo; i < hi; i++) Q[i + b * m] = vals[(i % rk + b * rk) * ldo + i / rk];
        } else {
            for (size_t a = 0; a < rk; a++)
                for (size_t i = lo; i < hi;) {          /* runs of j inside one b */
                    const size_t j = i % N, b = i / N;
                    size_t len = N - j;
                    if (len > hi - i) len = hi - i;
                    memcpy(Q + i + a * m, vals + (a + b * rk) * ldo + j, len * sizeof(double));
                    i += len;
                }
        }
    }
}
/* own rows of the cross core into the valuef_precompute_cores layout core[j*rk*rk1 + a + b*rk] */
static void team_store(la_ws *w, la_thr *th)
{
    if (!w->core) return;
    const size_t m = w->m, rk = w->rk, N = w->N, rk1 = w->rk1;
    const double *B = w->B;
    double *core = w->core;
    for (int blk = th->b0; blk < th->b1; blk++) {
        size_t lo, hi;
        la_rows(w, blk, 0, &lo, &hi);
        if (!w->right) {
            for (size_t b = 0; b < rk1; b++)
                for (size_t i = lo; i < hi; i++) core[(i / rk) * rk * rk1 + i % rk + b * rk] = B[i + b * m];
        } else {
            for (size_t i = lo; i < hi; i++) {
                double *t = core + (i % N) * rk * rk1 + (i / N) * rk;
                for (size_t a = 0; a < rk; a++) t[a] = B[i + a * m];
            }
        }
    }
}

/* Rows of the unfolding A (m x n, before the QR) that repeat an earlier row to round-off carry no
 * information for the pivoting (absorbing faces with a constant boundary cost produce whole families of
 * them); when the unfolding is rank-deficient the QR completes Q with arbitrary directions and maxvol would
 * happily pick such twins, which makes the NEXT unfolding rank-deficient as well.  skip[i] = 1 withholds
 * row i from the pivoting.  Twins are found through two fixed random projections of the rows (every thread
 * projects its own rows before the QR overwrites them) and a walk of thread 0 over the rows in order through a
 * hash of the quantised projection.  The walk is only needed when the QR met a column below TWIN_COND of the
 * largest: in a well-conditioned unfolding twin rows of Q agree to round-off, a second twin adds no volume, and
 * among equal candidates the first in scan order wins anyway. */
#define TWIN_COND 1e-4
static void team_twin_project(la_ws *w, la_thr *th)
{
    const size_t m = w->m, n = w->n;
    const double *A = w->Q;
    double *k1 = w->k1, *k2 = w->k1 + m;
    for (int blk = th->b0; blk < th->b1; blk++) {
        size_t lo, hi;
        la_rows(w, blk, 0, &lo, &hi);
        double scale = 0.0;
        uint64_t st = 0x7F1BE7ull;
        memset(w->skip + lo, 0, hi - lo);
        for (size_t i = lo; i < hi; i++) k1[i] = k2[i] = 0.0;
        for (size_t j = 0; j < n; j++) {
            const double mx = absmax(A + j * m + lo, hi - lo);
            if (mx > scale) scale = mx;
            st = st * 6364136223846793005ull + 1442695040888963407ull;
            axpy(0.5 + (double)(st >> 11) / 9007199254740992.0, A + j * m + lo, k1 + lo, hi - lo);
            st = st * 6364136223846793005ull + 1442695040888963407ull;
            axpy(0.5 + (double)(st >> 11) / 9007199254740992.0, A + j * m + lo, k2 + lo, hi - lo);
        }
        w->twin_scale[blk * 8] = scale;
    }
}
/* after a barrier that follows team_twin_project on every thread */
static void twin_walk(la_ws *w)
{
    const size_t m = w->m, n = w->n;
    double *k1 = w->k1, *k2 = w->k1 + m;
    double scale = 0.0;
    for (int blk = 0; blk < LA_NB; blk++) if (w->twin_scale[blk * 8] > scale) scale = w->twin_scale[blk * 8];
    const double tol = 1e-12 * scale * (double)n, quantum = 1024.0 * tol;
    if (!(tol > 0.0)) return;                           /* all-zero unfolding */
    size_t H = 16, eligible = m;
    while (H < 4 * m) H <<= 1;
    uint32_t *slot = w->slot;                           /* row + 1, keyed by the quantised k1 */
    int64_t *qk = w->qk;
    memset(slot, 0, H * sizeof(uint32_t));
    for (size_t i = 0; i < m; i++) {                    /* earlier rows stay eligible */
        qk[i] = (int64_t)floor(k1[i] / quantum);
        int twin = 0;
        for (int64_t dq = -1; dq <= 1 && !twin; dq++) {
            const int64_t q = qk[i] + dq;
            for (size_t h = (size_t)((uint64_t)q * 0x9E3779B97F4A7C15ull) & (H - 1); slot[h]; h = (h + 1) & (H - 1)) {
                const size_t r = slot[h] - 1;
                if (qk[r] == q && fabs(k1[i] - k1[r]) <= tol && fabs(k2[i] - k2[r]) <= tol) { twin = 1; break; }
            }
        }
        if (twin) { w->skip[i] = 1; eligible--; continue; }
        size_t h = (size_t)((uint64_t)qk[i] * 0x9E3779B97F4A7C15ull) & (H - 1);
        while (slot[h]) h = (h + 1) & (H - 1);
        slot[h] = (uint32_t)(i + 1);
    }
    if (eligible < n) memset(w->skip, 0, m);            /* not enough distinct rows: no restriction */
}

/* Orthonormal basis Q (m x n) of the column space of A, by Householder QR with column pivoting; Q
 * overwrites A (column order is immaterial: the cross core B = Q inv(Q[P,:]) only depends on span(Q)).
 * Columns whose remaining norm falls below RANK_EPS times the largest column norm are numerically dependent:
 * a reflector built from them would point wherever the round-off of the operator's values points, and the
 * pivoting after it would follow.  They get no reflector, so their Q columns are H_0..H_{rank-1} e_k -- a
 * completion that depends on the well-determined part only.
 * One barrier per reflector: the pass before it takes the dots of the pivot column a_k (rows below k) with every
 * column -- with itself for the norm, with the later columns for the reflection (the reflector is a_k / v0, so the
 * unscaled dots serve), with the earlier reflectors for G = V^T V.  From G and tau every thread builds the
 * triangular factor T of H_0 .. H_{r-1} = I - V T V^T (LAPACK dlarft, forward columnwise) and M = T V1^T, V1 the top
 * n x n of V, and then forms its own rows of Q = E - V M.  Row k of the matrix is frozen once reflector k's barrier
 * has passed (column swaps only touch the rows below), so it can be read by everyone afterwards.  Every thread
 * keeps its own copy of tau and of the downdated column norms (LAPACK dgeqp3 style, recomputed exactly once they
 * have lost six digits). */
static double team_qr_basis(la_ws *w, la_thr *th)
{
    const size_t m = w->m, n = w->n, ncap = w->ncap, pstride = w->pstride;
    const int b0 = th->b0, b1 = th->b1;
    double *A = w->Q, *part = w->part;
    double *pv = w->priv + (size_t)th->tid * w->privlen;
    double *tau = pv, *vn2 = pv + ncap, *vref = pv + 2 * ncap, *rowk = pv + 3 * ncap, *sc = pv + 4 * ncap;
    size_t *redo = (size_t *)(pv + 5 * ncap);
    double *G = pv + 8 * ncap, *T = G + ncap * ncap, *M = T + ncap * ncap, *Vb = pv + 8 * ncap + 5 * ncap * ncap + LA_NB * ncap;
    int pb = 0;
    for (int blk = b0; blk < b1; blk++) {
        size_t lo, hi;
        la_rows(w, blk, 0, &lo, &hi);
        for (size_t j = 0; j < n; j++) PART(pb, blk, j) = dot8(A + j * m + lo, A + j * m + lo, hi - lo);
    }
    LA_BARRIER();
    for (size_t j = 0; j < n; j++) {
        double s = 0.0;
        for (int blk = 0; blk < LA_NB; blk++) s += PART(pb, blk, j);
        vn2[j] = vref[j] = s;
    }
    pb ^= 1;
    size_t rank = n;
    double ref = 0.0, cond = 1.0;                         /* smallest pivot column norm over the largest */
    for (size_t k = 0; k < n; k++) {
        size_t p = k; double best = -1.0;
        for (size_t j = k; j < n; j++) if (vn2[j] > best) { best = vn2[j]; p = j; }
        for (size_t j = k; j < p; j++) if (vn2[j] >= best * (1.0 - TIE_EPS)) { p = j; break; }
        double *ak = A + k * m;
        if (p != k) {
            double *ap = A + p * m;
            for (int blk = b0; blk < b1; blk++) {
                size_t lo, hi;
                la_rows(w, blk, k, &lo, &hi);
                for (size_t i = lo; i < hi; i++) { const double t = ak[i]; ak[i] = ap[i]; ap[i] = t; }
            }
            double t = vn2[k]; vn2[k] = vn2[p]; vn2[p] = t;
            t = vref[k]; vref[k] = vref[p]; vref[p] = t;
        }
        for (int blk = b0; blk < b1; blk++) {
            size_t lo, hi;
            la_rows(w, blk, k + 1, &lo, &hi);
            for (size_t j = 0; j < n; j++) PART(pb, blk, j) = dot8(ak + lo, A + j * m + lo, hi - lo);
        }
        LA_BARRIER();
        for (size_t j = 0; j < n; j++) {
            double s = 0.0;
            for (int blk = 0; blk < LA_NB; blk++) s += PART(pb, blk, j);
            sc[j] = s;
            rowk[j] = A[k + j * m];
        }
        pb ^= 1;
        const double akk = rowk[k];
        const double nrm = sqrt(akk * akk + sc[k]);
        if (k == 0) ref = nrm;
        if (nrm <= RANK_EPS * ref || nrm == 0.0) { rank = k; cond = 0.0; break; }
        if (nrm < cond * ref) cond = nrm / ref;
        const double alpha = akk >= 0.0 ? -nrm : nrm;
        const double v0 = akk - alpha;
        tau[k] = -v0 / alpha;
        for (size_t j = 0; j < k; j++) G[j + k * n] = rowk[j] + sc[j] / v0;               /* v_j . v_k, v_k[k] = 1 */
        for (size_t j = k + 1; j < n; j++) sc[j] = (rowk[j] + sc[j] / v0) * tau[k];
        for (int blk = b0; blk < b1; blk++) {
            size_t lo, hi;
            la_rows(w, blk, k + 1, &lo, &hi);
            for (size_t i = lo; i < hi; i++) ak[i] /= v0;
            for (size_t j = k + 1; j < n; j++) axpy(-sc[j], ak + lo, A + j * m + lo, hi - lo);
        }
        size_t nre = 0;
        for (size_t j = k + 1; j < n; j++) {
            const double ajk = rowk[j] - sc[j];
            vn2[j] -= ajk * ajk;
            if (!(vn2[j] > 1e-6 * vref[j])) redo[nre++] = j;
        }
        if (nre) {                                       /* the same columns on every thread */
            for (int blk = b0; blk < b1; blk++) {
                size_t lo, hi;
                la_rows(w, blk, k + 1, &lo, &hi);
                for (size_t q = 0; q < nre; q++) PART(pb, blk, redo[q]) = dot8(A + redo[q] * m + lo, A + redo[q] * m + lo, hi - lo);
            }
            LA_BARRIER();
            for (size_t q = 0; q < nre; q++) {
                double s = 0.0;
                for (int blk = 0; blk < LA_NB; blk++) s += PART(pb, blk, redo[q]);
                vn2[redo[q]] = vref[redo[q]] = s;
            }
            pb ^= 1;
        }
    }
    /* the top n x n of V (unit lower triangular in its first `rank` columns), from the owners of those rows */
    double *V1 = w->V1;
    for (int blk = b0; blk < b1; blk++) {
        size_t lo, hi;
        la_rows(w, blk, 0, &lo, &hi);
        for (size_t i = lo; i < hi && i < n; i++)
            for (size_t k = 0; k < rank; k++) V1[i + k * n] = i < k ? 0.0 : (i == k ? 1.0 : A[i + k * m]);
    }
    LA_BARRIER();
    for (size_t k = 0; k < rank; k++) {                 /* T: upper triangular */
        for (size_t i = 0; i < k; i++) {
            double z = 0.0;
            for (size_t l = i; l < k; l++) z += T[i + l * n] * G[l + k * n];
            T[i + k * n] = -tau[k] * z;
        }
        T[k + k * n] = tau[k];
    }
    for (size_t c = 0; c < n; c++)                      /* M[k,c] = sum_j T[k,j] V1[c,j],  k <= j <= min(c, rank-1) */
        for (size_t k = 0; k < rank; k++) {
            double s = 0.0;
            for (size_t j = k; j < rank && j <= c; j++) s += T[k + j * n] * V1[c + j * n];
            M[k + c * n] = s;
        }
    for (int blk = b0; blk < b1; blk++) {               /* own rows of Q = E - V M */
        size_t lo, hi;
        la_rows(w, blk, 0, &lo, &hi);
        const size_t L = hi - lo;
        if (!L) continue;
        for (size_t k = 0; k < rank; k++) {
            double *vb = Vb + k * L;
            const double *ak = A + k * m;
            for (size_t i = lo; i < hi; i++) vb[i - lo] = i < k ? 0.0 : (i == k ? 1.0 : ak[i]);
        }
        for (size_t c = 0; c < n; c++) {
            double *qc = A + c * m + lo;
            for (size_t i = 0; i < L; i++) qc[i] = 0.0;
            if (c >= lo && c < hi) qc[c - lo] = 1.0;
            for (size_t k = 0; k < rank; k++) {
                const double mk = M[k + c * n];
                if (mk != 0.0) axpy(-mk, Vb + k * L, qc, L);
            }
        }
    }    return cond;
}

/* S = inv(Q[P,:]) by Gauss-Jordan with partial pivoting on [Q[P] | I] (every thread its own copy) */
static int invert_pivot_block(const la_ws *w, const size_t *P, double *aug /* n x 2n row-major */, double *S)
{
    const size_t m = w->m, n = w->n;
    const double *Q = w->Q;
    for (size_t a = 0; a < n; a++)
        for (size_t b = 0; b < n; b++) {
            aug[a * 2 * n + b] = Q[P[a] + b * m];
            aug[a * 2 * n + n + b] = a == b ? 1.0 : 0.0;
        }
    for (size_t k = 0; k < n; k++) {
        size_t piv = k; double best = fabs(aug[k * 2 * n + k]);
        for (size_t i = k + 1; i < n; i++)
            if (fabs(aug[i * 2 * n + k]) > best) { best = fabs(aug[i * 2 * n + k]); piv = i; }
        if (best == 0.0) return 2;
        if (piv != k)
            for (size_t jx = 0; jx < 2 * n; jx++) { double t = aug[k * 2 * n + jx]; aug[k * 2 * n + jx] = aug[piv * 2 * n + jx]; aug[piv * 2 * n + jx] = t; }
        const double pv = 1.0 / aug[k * 2 * n + k];
        for (size_t jx = 0; jx < 2 * n; jx++) aug[k * 2 * n + jx] *= pv;
        for (size_t i = 0; i < n; i++) {
            if (i == k) continue;
            const double fct = aug[i * 2 * n + k];
            if (fct == 0.0) continue;
            for (size_t jx = 0; jx < 2 * n; jx++) aug[i * 2 * n + jx] -= fct * aug[k * 2 * n + jx];
        }
    }
    for (size_t a = 0; a < n; a++)
        for (size_t b = 0; b < n; b++) S[a + b * n] = aug[a * 2 * n + n + b];
    return 0;
}

/* max |x[i]| over the rows that are not withheld, the held-back row `hold` (or LA_NONE) counted with value hv */
static double absmax_eligible(const double *x, const char *skip, int any_skip, size_t lo, size_t hi, size_t hold, double hv)
{
    double c;
    if (!any_skip) {
        if (hold >= lo && hold < hi) {
            c = fabs(hv);
            const double c0 = absmax(x + lo, hold - lo), c1 = absmax(x + hold + 1, hi - hold - 1);
            if (c0 > c) c = c0;
            if (c1 > c) c = c1;
        } else c = absmax(x + lo, hi - lo);
        return c;
    }
    c = 0.0;
    for (size_t i = lo; i < hi; i++) {
        if (skip[i]) continue;
        const double a = fabs(i == hold ? hv : x[i]);
        if (a > c) c = a;
    }
    return c;
}

/* the swap record of a block: its largest eligible |B| and, for the columns within TIE_EPS of it, the first
 * eligible row within TIE_EPS of the column's maximum and the first row that reaches it */
static void swap_record(const la_ws *w, int blk, const double *cme /* n */, size_t hold, la_srec *r)
{
    const size_t m = w->m, n = w->n;
    const double *B = w->B;
    const char *skip = w->skip;
    size_t lo, hi;
    la_rows(w, blk, 0, &lo, &hi);
    double best = 0.0;
    for (size_t j = 0; j < n; j++) if (cme[j] > best) best = cme[j];
    r->best = best; r->nl = 0;
    if (!(best > 1.0)) return;                          /* nothing here can ask for a swap */
    for (size_t j = 0; j < n && r->nl < LA_SWL; j++) {
        if (cme[j] < best * (1.0 - TIE_EPS)) continue;
        const double thr = cme[j] * (1.0 - TIE_EPS);
        size_t cand = LA_NONE, arg = LA_NONE;
        for (size_t i = lo; i < hi; i++) {
            if (skip[i] || i == hold) continue;
            const double a = fabs(B[i + j * m]);
            if (cand == LA_NONE && a >= thr) cand = i;
            if (a == cme[j]) { arg = i; break; }
        }
        const size_t l = r->nl++;
        r->j[l] = j; r->cm[l] = cme[j]; r->cand[l] = cand; r->arg[l] = arg;
        r->candval[l] = cand == LA_NONE ? -1.0 : fabs(B[cand + j * m]);
    }
}

/* maxvol: rows P of Q (m x n) with |det Q[P]| locally maximal, and B = Q inv(Q[P]).
 * Symmetric problems (V(x) = V(-x)) make mirrored rows tie exactly in exact arithmetic; which one wins would
 * then depend on the last bits of the operator's values.  Entries within TIE_EPS of the maximum count as
 * tied and the first in scan order wins (blocks in order; inside a block the first row within TIE_EPS of the
 * block's own maximum if that row is within TIE_EPS of the overall one, else the block's maximum), so two
 * operators that agree to round-off pick the same rows.
 * Start rows: Gaussian elimination with row pivoting on a copy, one barrier per column; the chosen row is read by
 * everyone and left alone by its owner from then on.  When the rows that are not withheld do not reach a direction
 * (their maximum below 1e-6 of the overall one) all rows compete.
 * Swaps while some eligible |B[i,j]| > 1 + delta: the first near-maximal entry in column-major order enters; one
 * barrier per swap.  The owner of the entering row holds its new values back until the others have read the old. */
static int team_maxvol(la_ws *w, la_thr *th)
{
    const size_t m = w->m, n = w->n, ncap = w->ncap;
    const int tid = th->tid, b0 = th->b0, b1 = th->b1;
    const double *Q = w->Q;
    double *B = w->B;
    const char *skip = w->skip;
    char *used = w->used;
    double *pv = w->priv + (size_t)tid * w->privlen;
    double *f = pv, *col = pv + ncap, *newrow = pv + 2 * ncap;
    size_t *P = (size_t *)(pv + 3 * ncap);
    double *S = pv + 8 * ncap, *aug = S + ncap * ncap, *cme_all = pv + 8 * ncap + 5 * ncap * ncap;
#define CME(blk) (cme_all + (size_t)((blk) - b0) * ncap)     /* eligible column maxima of an own block */
    int any_skip[LA_NB];
    int pb = 0;
    for (int blk = b0; blk < b1; blk++) {
        size_t lo, hi;
        la_rows(w, blk, 0, &lo, &hi);
        memset(used + lo, 0, hi - lo);
        any_skip[blk] = 0;
        for (size_t i = lo; i < hi; i++) any_skip[blk] |= skip[i];
        for (size_t j = 0; j < n; j++) memcpy(B + j * m + lo, Q + j * m + lo, (hi - lo) * sizeof(double));
    }
    for (size_t j = 0; j < n; j++) {
        const double *bj = B + j * m;
        for (int blk = b0; blk < b1; blk++) {
            size_t lo, hi;
            la_rows(w, blk, 0, &lo, &hi);
            la_erec r = { -1.0, -1.0, -1.0, -1.0, LA_NONE, LA_NONE, LA_NONE, LA_NONE };
            for (size_t i = lo; i < hi; i++) {
                if (used[i]) continue;
                const double a = fabs(bj[i]);
                if (a > r.a_max) { r.a_max = a; r.a_arg = i; }
                if (!skip[i] && a > r.e_max) { r.e_max = a; r.e_arg = i; }
            }
            for (size_t i = lo; i < hi && (r.a_cand == LA_NONE || (r.e_cand == LA_NONE && r.e_arg != LA_NONE)); i++) {
                if (used[i]) continue;
                const double a = fabs(bj[i]);
                if (r.a_cand == LA_NONE && a >= r.a_max * (1.0 - TIE_EPS)) { r.a_cand = i; r.a_candval = a; }
                if (r.e_cand == LA_NONE && !skip[i] && a >= r.e_max * (1.0 - TIE_EPS)) { r.e_cand = i; r.e_candval = a; }
            }
            w->erec[pb * LA_NB + blk] = r;
        }
        LA_BARRIER();
        const la_erec *er = w->erec + pb * LA_NB;
        pb ^= 1;
        double best_any = 0.0, best_e = -1.0;
        for (int blk = 0; blk < LA_NB; blk++) {
            if (er[blk].a_max > best_any) best_any = er[blk].a_max;
            if (er[blk].e_max > best_e) best_e = er[blk].e_max;
        }
        const int all_rows = best_e < 1e-6 * best_any;  /* the distinct rows do not reach this direction */
        const double thr = (all_rows ? best_any : best_e) * (1.0 - TIE_EPS);
        size_t piv = LA_NONE;
        for (int blk = 0; blk < LA_NB && piv == LA_NONE; blk++) {
            const double bmax = all_rows ? er[blk].a_max : er[blk].e_max;
            if (bmax < 0.0 || bmax < thr) continue;
            if (all_rows) piv = er[blk].a_candval >= thr ? er[blk].a_cand : er[blk].a_arg;
            else piv = er[blk].e_candval >= thr ? er[blk].e_cand : er[blk].e_arg;
        }
        if (piv == LA_NONE) return 2;                  /* no unused row left: n > m (the same on every thread) */
        P[j] = piv;
        if (la_mine(w, piv, th)) used[piv] = 1;
        const double pvv = bj[piv];
        if (pvv == 0.0) continue;
        for (size_t c = j + 1; c < n; c++) f[c] = B[piv + c * m] / pvv;
        for (int blk = b0; blk < b1; blk++) {
            size_t lo, hi;
            la_rows(w, blk, 0, &lo, &hi);
            for (int seg = 0; seg < 2; seg++) {         /* the pivot row stays as the others read it */
                size_t s0 = lo, s1 = hi;
                if (piv >= lo && piv < hi) { if (seg == 0) s1 = piv; else s0 = piv + 1; }
                else if (seg == 1) break;
                for (size_t c = j + 1; c < n; c++)
                    if (f[c] != 0.0) axpy(-f[c], bj + s0, B + c * m + s0, s1 - s0);
            }
        }
    }
    if (invert_pivot_block(w, P, aug, S)) return 2;
    for (int blk = b0; blk < b1; blk++) {               /* B[:,b] = sum_a S[a,b] Q[:,a] */
        size_t lo, hi;
        la_rows(w, blk, 0, &lo, &hi);
        for (size_t b = 0; b < n; b++) {
            double *bb = B + b * m;
            for (size_t i = lo; i < hi; i++) bb[i] = 0.0;
            for (size_t a = 0; a < n; a++) axpy(S[a + b * n], Q + a * m + lo, bb + lo, hi - lo);
            CME(blk)[b] = absmax_eligible(bb, skip, any_skip[blk], lo, hi, LA_NONE, 0.0);
        }
        swap_record(w, blk, CME(blk), LA_NONE, w->srec + pb * LA_NB + blk);
    }
    size_t held = LA_NONE;                               /* entering row whose new values this thread holds back */
    int it = 0;
    for (; it < 200; it++) {
        LA_BARRIER();
        if (held != LA_NONE) { for (size_t b = 0; b < n; b++) B[held + b * m] = newrow[b]; held = LA_NONE; }
        const la_srec *sr = w->srec + pb * LA_NB;
        pb ^= 1;
        double best = 0.0;
        for (int blk = 0; blk < LA_NB; blk++) if (sr[blk].best > best) best = sr[blk].best;
        if (best <= 1.0 + 1e-2) break;
        const double thr = best * (1.0 - TIE_EPS);
        size_t bi = LA_NONE, bj = LA_NONE;              /* first near-maximal entry, column-major order */
        for (int blk = 0; blk < LA_NB; blk++)
            for (size_t l = 0; l < sr[blk].nl; l++)
                if (sr[blk].cm[l] >= thr && (bj == LA_NONE || sr[blk].j[l] < bj)) {
                    bj = sr[blk].j[l];
                    bi = sr[blk].candval[l] >= thr ? sr[blk].cand[l] : sr[blk].arg[l];
                }
        if (bi == LA_NONE) break;
        /* row bi replaces P[bj]:  B <- B - B[:,bj] (B[bi,:] - e_bj) / B[bi,bj] */
        const double pvv = B[bi + bj * m];
        for (size_t b = 0; b < n; b++) col[b] = (B[bi + b * m] - (b == bj ? 1.0 : 0.0)) / pvv;
        const double fs = 1.0 - col[bj];
        if (la_mine(w, bi, th)) {
            for (size_t b = 0; b < n; b++)
                newrow[b] = b == bj ? pvv * fs : (col[b] == 0.0 ? B[bi + b * m] : B[bi + b * m] + (-col[b]) * pvv);
            held = bi;
        }
        for (int blk = b0; blk < b1; blk++) {
            size_t lo, hi;
            la_rows(w, blk, 0, &lo, &hi);
            const int has = bi >= lo && bi < hi;
            double *bb = B + bj * m;
            for (size_t b = 0; b < n; b++) {
                if (b == bj || col[b] == 0.0) continue;
                double *xb = B + b * m;
                if (!any_skip[blk]) {
                    double c = has ? fabs(newrow[b]) : 0.0;
                    for (int seg = 0; seg < 2; seg++) {
                        size_t s0 = lo, s1 = hi;
                        if (has) { if (seg == 0) s1 = bi; else s0 = bi + 1; }
                        else if (seg == 1) break;
                        const double cs = axpy_absmax(-col[b], bb + s0, xb + s0, s1 - s0);
                        if (cs > c) c = cs;
                    }
                    CME(blk)[b] = c;
                } else {
                    if (has) { axpy(-col[b], bb + lo, xb + lo, bi - lo); axpy(-col[b], bb + bi + 1, xb + bi + 1, hi - bi - 1); }
                    else axpy(-col[b], bb + lo, xb + lo, hi - lo);
                    CME(blk)[b] = absmax_eligible(xb, skip, 1, lo, hi, has ? bi : LA_NONE, has ? newrow[b] : 0.0);
                }
            }
            for (size_t i = lo; i < hi; i++) if (i != bi) bb[i] *= fs;
            CME(blk)[bj] = absmax_eligible(bb, skip, any_skip[blk], lo, hi, has ? bi : LA_NONE, has ? newrow[bj] : 0.0);
            swap_record(w, blk, CME(blk), has ? bi : LA_NONE, w->srec + pb * LA_NB + blk);
        }
        P[bj] = bi;
    }
    if (it == 200) {                                     /* swap limit: the last entering row is still held back */
        LA_BARRIER();
        if (held != LA_NONE) for (size_t b = 0; b < n; b++) B[held + b * m] = newrow[b];
    }
    if (tid == 0) for (size_t j = 0; j < n; j++) w->P[j] = P[j];
    return 0;
#undef CME
}

/* twin rows + QR basis + maxvol of an m x n unfolding: w->Q <- basis, w->B <- cross core, w->P <- rows */
static int pivot_step(la_ws *w, size_t m, size_t n)
{
    if (m > w->mcap || n > w->ncap || n > m) return 1;
    w->m = m; w->n = n; w->rc = 0;
#ifdef _OPENMP
    atomic_store(&w->bar_count, 0); atomic_store(&w->bar_sense, 0);
#endif
    w->bs = la_block_rows(m);
    int T = m * n < 8192 ? 1 : w->threads;
    (void)T;
#ifdef _OPENMP
#pragma omp parallel num_threads(T) if (T > 1)
#endif
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        const int tid = 0, nt = 1;
#endif
        la_thr th = { tid, nt, tid * LA_NB / nt, (tid + 1) * LA_NB / nt, 0 };
        team_unfold(w, &th);
        team_twin_project(w, &th);
        const double cond = team_qr_basis(w, &th);     /* the same on every thread */
        if (cond < TWIN_COND) {
            if (tid == 0) twin_walk(w);
            la_barrier(w, &th);
        }
        const int rc = team_maxvol(w, &th);             /* the same on every thread */
        if (tid == 0) w->rc = rc;
        if (!rc) team_store(w, &th);
    }
    return w->rc;
}

/* ---- one core step: all fibers of core k in ONE operator call --------------------------------- */
static double g_t_eval, g_t_piv, g_t_dot;
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

/* all r_k * r_{k+1} fibers of core k in ONE operator call: vals[(a + b*rk) * ldo + j] = T(I_k[a], j, J_{k+1}[b]) */
static int eval_core(const c3sc_cross *c, uint32_t k, c3sc_fiber_batch_fn f, void *arg, int32_t *dv, int32_t *fi,
                     double *vals, uint64_t *nfib)
{
    const uint32_t d = c->d;
    const size_t rk = c->r[k], rk1 = c->r[k + 1], ldo = c->nmax;
    const size_t F = rk * rk1;
    for (size_t b = 0; b < rk1; b++)
        for (size_t a = 0; a < rk; a++) {
            const size_t fidx = a + b * rk;
            dv[fidx] = (int32_t)k;
            for (uint32_t i = 0; i < d; i++)
                fi[fidx * d + i] = i < k ? c->I[k][a * d + i] : (i > k ? c->J[k + 1][b * d + i] : 0);
        }
    const double t0_ = now_s();
    int rc = f(F, dv, fi, ldo, vals, arg);
    g_t_eval += now_s() - t0_;
    if (rc) return rc;
    *nfib += F;
    return 0;
}
/* an edge core goes straight into the valuef_precompute_cores layout core[j*rk*rk1 + a + b*rk] */
static void store_vals(const double *vals, size_t ldo, size_t rk, size_t N, size_t rk1, double *core)
{
    for (size_t b = 0; b < rk1; b++)
        for (size_t a = 0; a < rk; a++) {
            const double *v = vals + (a + b * rk) * ldo;
            double *t = core + a + b * rk;
            for (size_t j = 0; j < N; j++) t[j * rk * rk1] = v[j];
        }
}

/* <A, B> of two trains in the ValueF layout (discrete inner product over the grid nodes):
 * M (rA_k x rB_k) <- sum_j A_k[j]^T M B_k[j], two r^3 products per node.  w1, w2, w3: rmax^2 each. */
static double tt_dot2(uint32_t d, const uint64_t *n, const uint64_t *ra, double *const *A, const uint64_t *rb, double *const *B,
                      double *w1, double *w2, double *w3)
{
    /* per core: At[j + e*N] = A_k[j][e] (nodes contiguous), likewise Bt; W[(j + x*N) + q*a0*N] = sum_y M[x,y] Bt[j + (y + q*b0)*N]
       (axpys of length N); new M[p,q] = <At[:, (x,p)], W[:, (x,q)]> over (j,x) (dots of length a0*N).  w3 unused. */
    (void)w3;
    size_t cap = 1;
    for (uint32_t k = 0; k < d; k++) {
        const size_t ea = n[k] * ra[k] * ra[k + 1], eb = n[k] * rb[k] * rb[k + 1], ew = n[k] * ra[k] * rb[k + 1];
        if (ea > cap) cap = ea;
        if (eb > cap) cap = eb;
        if (ew > cap) cap = ew;
    }
    double *At = (double *)malloc(3 * cap * sizeof(double));
    if (!At) return NAN;
    double *Bt = At + cap, *W = At + 2 * cap;
    w1[0] = 1.0;
    const int T = la_threads();
    (void)T;
    for (uint32_t k = 0; k < d; k++) {
        const size_t a0 = ra[k], a1 = ra[k + 1], b0 = rb[k], b1 = rb[k + 1], N = n[k];
        const int same = A[k] == B[k] && a0 == b0 && a1 == b1;
        const double *Bu = same ? At : Bt;
        /* every column q of the carried matrix is one thread's from the transposes to the dots: the same sums
           in the same order whatever the team size */
#ifdef _OPENMP
#pragma omp parallel num_threads(T) if (T > 1 && N * a0 * b0 * b1 >= 32768)
#endif
        {
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
            for (size_t e = 0; e < a0 * a1; e++)
                for (size_t j = 0; j < N; j++) At[j + e * N] = A[k][j * a0 * a1 + e];
            if (!same) {
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
                for (size_t e = 0; e < b0 * b1; e++)
                    for (size_t j = 0; j < N; j++) Bt[j + e * N] = B[k][j * b0 * b1 + e];
            }
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
            for (size_t q = 0; q < b1; q++) {
                double *Wq = W + q * a0 * N;
                for (size_t e = 0; e < a0 * N; e++) Wq[e] = 0.0;
                for (size_t y = 0; y < b0; y++)
                    for (size_t x = 0; x < a0; x++) axpy(w1[x + y * a0], Bu + (y + q * b0) * N, Wq + x * N, N);
                for (size_t p = 0; p < a1; p++) w2[p + q * a1] = dot8(At + p * a0 * N, Wq, a0 * N);
            }
        }
        memcpy(w1, w2, a1 * b1 * sizeof(double));
    }
    free(At);
    return w1[0];
}
static double tt_dot(uint32_t d, const uint64_t *n, const uint64_t *r, double *const *A, double *const *B, double *w1, double *w2,
                     double *w3)
{
    return tt_dot2(d, n, r, A, r, B, w1, w2, w3);
}

/* valuef_norm / valuef_norm2diff (src/valuefunc.c:315-335) on nodal cores: the discrete l2 norm over the
 * grid nodes (C3's function_train_norm2 integrates the piecewise-linear interpolant instead; the
 * solvers only use these numbers as a Cauchy criterion, src/bellman.c:2307-2338). */
double c3sc_cores_dot(uint32_t d, const uint64_t *n, const uint64_t *ra, const double *const *A, const uint64_t *rb,
                      const double *const *B)
{
    size_t rmax = 1;
    for (uint32_t k = 0; k <= d; k++) { if (ra[k] > rmax) rmax = ra[k]; if (rb[k] > rmax) rmax = rb[k]; }
    double *w = (double *)malloc(3 * rmax * rmax * sizeof(double));
    if (!w) return NAN;
    const double v = tt_dot2(d, n, ra, (double *const *)A, rb, (double *const *)B, w, w + rmax * rmax, w + 2 * rmax * rmax);
    free(w);
    return v;
}
double c3sc_cores_norm(uint32_t d, const uint64_t *n, const uint64_t *r, const double *const *A)
{
    return sqrt(fabs(c3sc_cores_dot(d, n, r, A, r, A)));
}
double c3sc_cores_norm2diff(uint32_t d, const uint64_t *n, const uint64_t *ra, const double *const *A, const uint64_t *rb,
                            const double *const *B)
{
    const double aa = c3sc_cores_dot(d, n, ra, A, ra, A), ab = c3sc_cores_dot(d, n, ra, A, rb, B), bb = c3sc_cores_dot(d, n, rb, B, rb, B);
    return sqrt(fabs(aa - 2.0 * ab + bb));
}

/* ---- continuous L2 inner product of two piecewise-linear (LINELM) function trains ---------------------------------
 * What the reference's valuef_norm / valuef_norm2diff compute (src/valuefunc.c:315-335 -> C3 function_train_norm2 /
 * norm2diff on linear elements): <f, g> = integral over the box of f g, f and g the multilinear interpolants of the
 * nodal cores.  Per dimension the nodal values meet through the mass matrix of the hat functions on the grid,
 *     M_ii = (h_{i-1} + h_i) / 3,   M_{i,i+1} = M_{i+1,i} = h_i / 6        (h_i = x_{i+1} - x_i, h_{-1} = h_{N-1} = 0),
 * so the train contraction of c3sc_cores_dot becomes
 *     Z_{k+1}[a2, b2] = sum_{i,j} M^k_ij sum_{a,b} A_k[i][a, a2] Z_k[a, b] B_k[j][b, b2]
 * evaluated as T_j = Z_k B_k[j], H_i = sum_j M_ij T_j (three terms), Z_{k+1} += A_k[i]^T H_i.  C3 itself is absent:
 * pinned against a dense evaluation of the same integral (tests/test_cross_driver.py).  The nodal form (the discrete
 * l2 of the node values) stays available as c3sc_cores_dot / _norm / _norm2diff. */
double c3sc_cores_dot_l2(uint32_t d, const uint64_t *n, const double *const *xgrid, const uint64_t *ra, const double *const *A,
                         const uint64_t *rb, const double *const *B)
{
    if (!n || !xgrid || !ra || !A || !rb || !B || d < 1 || d > C3SC_MAXD) return NAN;
    size_t rmax = 1, nmax = 1;
    for (uint32_t k = 0; k <= d; k++) { if (ra[k] > rmax) rmax = ra[k]; if (rb[k] > rmax) rmax = rb[k]; }
    for (uint32_t k = 0; k < d; k++) if (n[k] > nmax) nmax = n[k];
    const size_t r2 = rmax * rmax;
    double *Z = (double *)calloc(2 * r2 + (nmax + 1) * r2, sizeof(double));
    if (!Z) return NAN;
    double *Zn = Z + r2, *T = Zn + r2;                      /* T[j]: ra[k] x rb[k+1], column-major, all nodes of the dimension */
    Z[0] = 1.0;
    for (uint32_t k = 0; k < d; k++) {
        const size_t a1 = ra[k], a2 = ra[k + 1], b1 = rb[k], b2 = rb[k + 1], N = n[k];
        const double *x = xgrid[k];
        /* T_j[a, c] = sum_b Z[a, b] B_k[j][b, c]   (Z column-major a + b*a1, B block column-major b + c*b1) */
        for (size_t j = 0; j < N; j++) {
            const double *Bj = B[k] + j * b1 * b2;
            double *Tj = T + j * a1 * b2;
            for (size_t c = 0; c < b2; c++)
                for (size_t a = 0; a < a1; a++) {
                    double sum = 0.0;
                    for (size_t b = 0; b < b1; b++) sum += Z[a + b * a1] * Bj[b + c * b1];
                    Tj[a + c * a1] = sum;
                }
        }
        for (size_t e = 0; e < a2 * b2; e++) Zn[e] = 0.0;
        double *H = T + N * a1 * b2;                        /* one more block of scratch */
        for (size_t i = 0; i < N; i++) {
            const double hl = i > 0 ? x[i] - x[i - 1] : 0.0, hr = i + 1 < N ? x[i + 1] - x[i] : 0.0;
            const double mc = (hl + hr) / 3.0, ml = hl / 6.0, mr = hr / 6.0;
            const double *Tc = T + i * a1 * b2, *Tl = i > 0 ? Tc - a1 * b2 : Tc, *Tr = i + 1 < N ? Tc + a1 * b2 : Tc;
            for (size_t e = 0; e < a1 * b2; e++) H[e] = mc * Tc[e] + ml * Tl[e] + mr * Tr[e];
            const double *Ai = A[k] + i * a1 * a2;
            for (size_t c = 0; c < b2; c++)
                for (size_t q = 0; q < a2; q++) {
                    double sum = 0.0;
                    for (size_t a = 0; a < a1; a++) sum += Ai[a + q * a1] * H[a + c * a1];
                    Zn[q + c * a2] += sum;
                }
        }
        for (size_t e = 0; e < a2 * b2; e++) Z[e] = Zn[e];
    }
    const double v = Z[0];
    free(Z);
    return v;
}
double c3sc_cores_norm_l2(uint32_t d, const uint64_t *n, const double *const *xgrid, const uint64_t *r, const double *const *A)
{
    return sqrt(fabs(c3sc_cores_dot_l2(d, n, xgrid, r, A, r, A)));
}
double c3sc_cores_norm2diff_l2(uint32_t d, const uint64_t *n, const double *const *xgrid, const uint64_t *ra, const double *const *A,
                               const uint64_t *rb, const double *const *B)
{
    const double aa = c3sc_cores_dot_l2(d, n, xgrid, ra, A, ra, A), ab = c3sc_cores_dot_l2(d, n, xgrid, ra, A, rb, B),
                 bb = c3sc_cores_dot_l2(d, n, xgrid, rb, B, rb, B);
    return sqrt(fabs(aa - 2.0 * ab + bb));
}

int c3sc_cross_run(c3sc_cross *c, c3sc_fiber_batch_fn f, void *arg, const c3sc_cross_opts *opts, double *const *cores,
                   uint64_t *nfibers, double *rel_change)
{
    if (!c || !f || !cores) return C3SC_EINVAL;
    const uint32_t d = c->d;
    const uint32_t maxiter = opts && opts->maxiter ? opts->maxiter : 5;       /* src/valuefunc.c:632 */
    const double tol = opts ? opts->tol : 0.0;
    const int verbose = opts ? opts->verbose : 0;
    size_t rmax = 1, fmax = 1, tmax = 1;
    for (uint32_t k = 0; k < d; k++) {
        if (c->r[k + 1] > rmax) rmax = c->r[k + 1];
        if (c->r[k] * c->r[k + 1] > fmax) fmax = c->r[k] * c->r[k + 1];
        if (c->r[k] * c->n[k] * c->r[k + 1] > tmax) tmax = c->r[k] * c->n[k] * c->r[k + 1];
    }
    (void)tmax;
    const size_t ldo = c->nmax;
    double *vals = (double *)exchange_get(c, fmax * ldo * sizeof(double) + fmax * (d + 1) * sizeof(int32_t));
    int32_t *fi = vals ? (int32_t *)(vals + fmax * ldo) : NULL, *dv = fi ? fi + fmax * d : NULL;
    double *work = (double *)malloc((rmax * rmax * 3 + 16) * sizeof(double));
    double **prev = (double **)calloc(d, sizeof(double *));
    int32_t *tmpI = (int32_t *)malloc(rmax * d * sizeof(int32_t));
    size_t mcap = 1;
    for (uint32_t k = 0; k < d; k++) {
        if (k + 1 < d && c->r[k] * c->n[k] > mcap) mcap = c->r[k] * c->n[k];
        if (k >= 1 && c->n[k] * c->r[k + 1] > mcap) mcap = c->n[k] * c->r[k + 1];
    }
    la_ws *ws = la_ws_create(mcap, rmax);
    size_t *P = ws ? ws->P : NULL;
    int rc = C3SC_OK;
    uint64_t nfib = 0;
    double change = 1.0, prev_norm2 = 0.0;
    g_t_eval = g_t_piv = g_t_dot = 0.0;
    if (!vals || !ws || !work || !prev || !tmpI) { rc = C3SC_EINVAL; goto done; }
    for (uint32_t k = 0; k < d; k++) {
        prev[k] = (double *)calloc(c->r[k] * c->n[k] * c->r[k + 1], sizeof(double));
        if (!prev[k]) { rc = C3SC_EINVAL; goto done; }
    }
    for (uint32_t it = 0; it < maxiter; it++) {
        /* ---- left -> right: fix the left index sets I[1..d-1] ---------------------------------- */
        for (uint32_t k = 0; k + 1 < d; k++) {
            const size_t rk = c->r[k], rk1 = c->r[k + 1], N = c->n[k], m = rk * N;
            rc = eval_core(c, k, f, arg, dv, fi, vals, &nfib);
            if (rc) goto done;
            ws->vals = vals; ws->core = cores[k]; ws->ldo = ldo; ws->rk = rk; ws->N = N; ws->rk1 = rk1; ws->right = 0;
            { const double t_ = now_s(); const int mv = pivot_step(ws, m, rk1); g_t_piv += now_s() - t_; if (mv) { rc = C3SC_ENUMERIC; goto done; } }
            for (size_t b = 0; b < rk1; b++) {                      /* new left set: row (a,j) = a + j*rk */
                const size_t a = P[b] % rk, j = P[b] / rk;
                for (uint32_t i = 0; i < k; i++) tmpI[b * d + i] = c->I[k][a * d + i];
                tmpI[b * d + k] = (int32_t)j;
            }
            for (size_t b = 0; b < rk1; b++)
                for (uint32_t i = 0; i <= k; i++) c->I[k + 1][b * d + i] = tmpI[b * d + i];
        }
        rc = eval_core(c, d - 1, f, arg, dv, fi, vals, &nfib);
        if (rc) goto done;
        store_vals(vals, ldo, c->r[d - 1], c->n[d - 1], 1, cores[d - 1]);
        /* ---- right -> left: fix the right index sets J[1..d-1] ---------------------------------- */
        for (uint32_t k = d - 1; k >= 1; k--) {
            const size_t rk = c->r[k], rk1 = c->r[k + 1], N = c->n[k], m = N * rk1;
            rc = eval_core(c, k, f, arg, dv, fi, vals, &nfib);
            if (rc) goto done;
            ws->vals = vals; ws->core = cores[k]; ws->ldo = ldo; ws->rk = rk; ws->N = N; ws->rk1 = rk1; ws->right = 1;
            { const double t_ = now_s(); const int mv = pivot_step(ws, m, rk); g_t_piv += now_s() - t_; if (mv) { rc = C3SC_ENUMERIC; goto done; } }
            for (size_t a = 0; a < rk; a++) {
                const size_t j = P[a] % N, b = P[a] / N;
                tmpI[a * d + k] = (int32_t)j;
                for (uint32_t i = k + 1; i < d; i++) tmpI[a * d + i] = c->J[k + 1][b * d + i];
            }
            for (size_t a = 0; a < rk; a++)
                for (uint32_t i = k; i < d; i++) c->J[k][a * d + i] = tmpI[a * d + i];
        }
        rc = eval_core(c, 0, f, arg, dv, fi, vals, &nfib);
        if (rc) goto done;
        store_vals(vals, ldo, 1, c->n[0], c->r[1], cores[0]);
        /* ---- change of the train against the previous sweep pair (valuef_norm2diff idea); skipped when
                nobody asks for it (no tolerance, no output, not verbose): a fifth of a sweep's host time ---- */
        if (tol > 0.0 || rel_change || verbose) {
            double *w1 = work, *w2 = work + rmax * rmax, *w3 = work + 2 * rmax * rmax;
            const double t_ = now_s();
            const double aa = tt_dot(d, c->n, c->r, cores, cores, w1, w2, w3);
            const double ab = it == 0 ? 0.0 : tt_dot(d, c->n, c->r, cores, prev, w1, w2, w3);
            const double bb = prev_norm2;                            /* <prev, prev> is last sweep's <cores, cores> */
            prev_norm2 = aa;
            g_t_dot += now_s() - t_;
            const double diff2 = aa - 2.0 * ab + bb;
            change = aa > 0.0 ? sqrt(fabs(diff2) / aa) : 0.0;
            if (verbose) fprintf(stderr, "c3sc_cross: sweep %u, |T|=%g, rel change %g, fibers %llu; cumulative s: operator %.4f pivoting (unfold + twin rows + qr + maxvol + store, %d threads) %.4f norms %.4f\n",
                                 it, sqrt(aa), change, (unsigned long long)nfib, g_t_eval, ws->threads, g_t_piv, g_t_dot);
            for (uint32_t k = 0; k < d; k++) memcpy(prev[k], cores[k], c->r[k] * c->n[k] * c->r[k + 1] * sizeof(double));
            if (tol > 0.0 && change < tol) break;
        }
    }
done:
    if (nfibers) *nfibers = nfib;
    if (rel_change) *rel_change = change;
    if (prev) for (uint32_t k = 0; k < d; k++) free(prev[k]);
    free(prev); la_ws_free(ws); free(work); free(tmpI);
    return rc;
}

/* ---- rank adaptation: TT rounding + kicked ranks ------------------------------------------------------
 * The adapt == 1 branch of valuef_interp (src/valuefunc.c:637-648,:706-730) calls C3's
 * ftapprox_cross_rankadapt with cross_tol / round_tol / kickrank / maxrank.  C3 is not vendored; the
 * published algorithm (C3 paper, Alg. 4; Oseledets 2011 TT-rounding) is restated here on nodal cores:
 *   run the fixed-rank cross; round the result to round_tol; every interior rank the rounding did NOT
 *   reduce (and that is below maxrank) is kicked by kickrank and the cross is run again; the rounded
 *   train is the answer.
 * Inner products are the discrete l2 ones over the grid nodes (like c3sc_cores_norm).              */

/* one-sided Jacobi SVD of A (m x n, column-major, m >= n): A <- U (orthonormal columns where s > 0),
 * s[n] descending, V (n x n). */
static void svd_jacobi(double *A, size_t m, size_t n, double *s, double *V)
{
    for (size_t i = 0; i < n * n; i++) V[i] = 0.0;
    for (size_t i = 0; i < n; i++) V[i + i * n] = 1.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        int rotated = 0;
        for (size_t p = 0; p + 1 < n; p++)
            for (size_t q = p + 1; q < n; q++) {
                double al = 0.0, be = 0.0, ga = 0.0;
                for (size_t i = 0; i < m; i++) {
                    const double x = A[i + p * m], y = A[i + q * m];
                    al += x * x; be += y * y; ga += x * y;
                }
                if (ga == 0.0 || fabs(ga) <= 1e-15 * sqrt(al * be)) continue;
                rotated = 1;
                const double zeta = (be - al) / (2.0 * ga);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                for (size_t i = 0; i < m; i++) {
                    const double x = A[i + p * m], y = A[i + q * m];
                    A[i + p * m] = cs * x - sn * y; A[i + q * m] = sn * x + cs * y;
                }
                for (size_t i = 0; i < n; i++) {
                    const double x = V[i + p * n], y = V[i + q * n];
                    V[i + p * n] = cs * x - sn * y; V[i + q * n] = sn * x + cs * y;
                }
            }
        if (!rotated) break;
    }
    for (size_t j = 0; j < n; j++) {
        double nr = 0.0;
        for (size_t i = 0; i < m; i++) nr += A[i + j * m] * A[i + j * m];
        s[j] = sqrt(nr);
        if (s[j] > 0.0) for (size_t i = 0; i < m; i++) A[i + j * m] /= s[j];
    }
    for (size_t j = 0; j + 1 < n; j++) {                               /* selection sort, descending */
        size_t b = j;
        for (size_t q = j + 1; q < n; q++) if (s[q] > s[b]) b = q;
        if (b == j) continue;
        double t = s[j]; s[j] = s[b]; s[b] = t;
        for (size_t i = 0; i < m; i++) { t = A[i + j * m]; A[i + j * m] = A[i + b * m]; A[i + b * m] = t; }
        for (size_t i = 0; i < n; i++) { t = V[i + j * n]; V[i + j * n] = V[i + b * n]; V[i + b * n] = t; }
    }
}

/* function_train_round on nodal cores: right-to-left QR, then left-to-right truncated SVD with the
 * tail of every unfolding below eps * |T| / sqrt(d-1).  cores_out[k] needs the capacity of cores_in[k]. */
int c3sc_cores_round(uint32_t d, const uint64_t *n, const uint64_t *rin, const double *const *cin, double eps,
                     uint64_t *rout, double *const *cout)
{
    if (!n || !rin || !cin || !rout || !cout || d < 1 || d > C3SC_MAXD || rin[0] != 1 || rin[d] != 1 || eps < 0.0)
        return C3SC_EINVAL;
    for (uint32_t k = 1; k < d; k++)
        if (rin[k] > rin[k - 1] * n[k - 1] || rin[k] > rin[k + 1] * n[k]) return C3SC_EINVAL;   /* ranks beyond an unfolding */
    uint64_t r[C3SC_MAXD + 1];
    size_t tmax = 1, rmax = 1;
    for (uint32_t k = 0; k <= d; k++) { r[k] = rin[k]; if (rin[k] > rmax) rmax = rin[k]; }
    for (uint32_t k = 0; k < d; k++) if (rin[k] * n[k] * rin[k + 1] > tmax) tmax = rin[k] * n[k] * rin[k + 1];
    double *W[C3SC_MAXD] = {0};
    double *Q = (double *)malloc(tmax * sizeof(double)), *A0 = (double *)malloc(tmax * sizeof(double));
    double *R = (double *)malloc(rmax * rmax * sizeof(double)), *V = (double *)malloc(rmax * rmax * sizeof(double));
    double *sv = (double *)malloc(rmax * sizeof(double));
    double *work = (double *)malloc((tmax + rmax + 16) * sizeof(double));
    int rc = C3SC_OK;
    if (!Q || !A0 || !R || !V || !sv || !work) { rc = C3SC_EINVAL; goto done; }
    for (uint32_t k = 0; k < d; k++) {                                  /* ValueF layout -> [a + j*rk + b*rk*N] */
        const size_t rk = r[k], rk1 = r[k + 1], N = n[k];
        W[k] = (double *)malloc(rk * N * rk1 * sizeof(double));
        if (!W[k]) { rc = C3SC_EINVAL; goto done; }
        for (size_t j = 0; j < N; j++)
            for (size_t b = 0; b < rk1; b++)
                for (size_t a = 0; a < rk; a++) W[k][a + j * rk + b * rk * N] = cin[k][j * rk * rk1 + a + b * rk];
    }
    for (uint32_t k = d - 1; k >= 1; k--) {                             /* rows of core k orthonormal */
        const size_t rk = r[k], rk1 = r[k + 1], N = n[k], m = N * rk1;
        for (size_t a = 0; a < rk; a++)
            for (size_t j = 0; j < N; j++)
                for (size_t b = 0; b < rk1; b++) Q[(j + b * N) + a * m] = W[k][a + j * rk + b * rk * N];
        memcpy(A0, Q, m * rk * sizeof(double));
        qr_explicit_q(Q, m, rk, work);
        for (size_t a = 0; a < rk; a++)                                 /* R = Q^T A0 */
            for (size_t c2 = 0; c2 < rk; c2++) {
                double t = 0.0;
                for (size_t i = 0; i < m; i++) t += Q[i + a * m] * A0[i + c2 * m];
                R[a + c2 * rk] = t;
            }
        for (size_t a = 0; a < rk; a++)
            for (size_t j = 0; j < N; j++)
                for (size_t b = 0; b < rk1; b++) W[k][a + j * rk + b * rk * N] = Q[(j + b * N) + a * m];
        const size_t rp = r[k - 1], Np = n[k - 1], mp = rp * Np;         /* core k-1 <- core k-1 x R^T */
        for (size_t x = 0; x < mp; x++) {
            for (size_t a2 = 0; a2 < rk; a2++) {
                double t = 0.0;
                for (size_t a = 0; a < rk; a++) t += W[k - 1][x + a * mp] * R[a2 + a * rk];
                work[a2] = t;
            }
            for (size_t a2 = 0; a2 < rk; a2++) W[k - 1][x + a2 * mp] = work[a2];
        }
    }
    if (d > 1) {
        double nrm2 = 0.0;
        for (size_t e = 0; e < r[0] * n[0] * r[1]; e++) nrm2 += W[0][e] * W[0][e];
        const double delta2 = eps * eps * nrm2 / (double)(d - 1);
        for (uint32_t k = 0; k + 1 < d; k++) {
            const size_t rk = r[k], rk1 = r[k + 1], N = n[k], m = rk * N;
            size_t p, rnew;
            double *U, *G;                                                /* U: m x rnew, G = diag(s) Vt: rnew x rk1 */
            if (m >= rk1) {
                p = rk1;
                svd_jacobi(W[k], m, rk1, sv, V);                          /* W[k] <- U, V: rk1 x rk1 */
                U = W[k];
            } else {                                                      /* wide unfolding: SVD of the transpose */
                p = m;
                for (size_t i = 0; i < m; i++)
                    for (size_t b = 0; b < rk1; b++) Q[b + i * rk1] = W[k][i + b * m];
                svd_jacobi(Q, rk1, m, sv, V);                             /* Q <- V' (rk1 x m), V <- U' (m x m) */
                U = V;
            }
            double tail = 0.0;
            rnew = p;
            while (rnew > 1 && tail + sv[rnew - 1] * sv[rnew - 1] <= delta2) { tail += sv[rnew - 1] * sv[rnew - 1]; rnew--; }
            G = A0;
            for (size_t a2 = 0; a2 < rnew; a2++)
                for (size_t b = 0; b < rk1; b++)
                    G[a2 + b * rnew] = sv[a2] * (m >= rk1 ? V[b + a2 * rk1] : Q[b + a2 * rk1]);
            if (m < rk1)
                for (size_t a2 = 0; a2 < rnew; a2++)
                    for (size_t i = 0; i < m; i++) W[k][i + a2 * m] = U[i + a2 * m];
            /* core k keeps its first rnew columns (already in place); core k+1 <- G x core k+1 */
            const size_t N1 = n[k + 1], r2 = r[k + 2];
            double *Wn = (double *)malloc(rnew * N1 * r2 * sizeof(double));
            if (!Wn) { rc = C3SC_EINVAL; goto done; }
            for (size_t c2 = 0; c2 < N1 * r2; c2++)
                for (size_t a2 = 0; a2 < rnew; a2++) {
                    double t = 0.0;
                    for (size_t b = 0; b < rk1; b++) t += G[a2 + b * rnew] * W[k + 1][b + c2 * rk1];
                    Wn[a2 + c2 * rnew] = t;
                }
            free(W[k + 1]);
            W[k + 1] = Wn;
            r[k + 1] = rnew;
        }
    }
    for (uint32_t k = 0; k < d; k++) {
        const size_t rk = r[k], rk1 = r[k + 1], N = n[k];
        for (size_t j = 0; j < N; j++)
            for (size_t b = 0; b < rk1; b++)
                for (size_t a = 0; a < rk; a++) cout[k][j * rk * rk1 + a + b * rk] = W[k][a + j * rk + b * rk * N];
    }
    for (uint32_t k = 0; k <= d; k++) rout[k] = r[k];
done:
    for (uint32_t k = 0; k < d; k++) free(W[k]);
    free(Q); free(A0); free(R); free(V); free(sv); free(work);
    return rc;
}

/* SplitMix64, for the index nodes a kicked rank starts from */
static uint64_t mix64(uint64_t *st)
{
    uint64_t z = (*st += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* change interior rank k to rnew.  Shrinking keeps the first rnew multi-indices of both sets.  Growing
 * appends right multi-indices not yet in J[k] (C3's cross_index_copylast repeats the last one instead,
 * which makes the next maxvol matrix singular by construction; fresh nodes serve the same purpose);
 * the left set I[k] only needs the room -- the next left-to-right sweep rebuilds it. */
static int cross_resize(c3sc_cross *c, uint32_t k, uint64_t rnew)
{
    const uint32_t d = c->d;
    const uint64_t rold = c->r[k];
    if (rnew == rold) return 0;
    int32_t *I = (int32_t *)realloc(c->I[k], (rnew * d + 1) * sizeof(int32_t));
    if (!I) return 1;
    c->I[k] = I;
    int32_t *J = (int32_t *)realloc(c->J[k], (rnew * d + 1) * sizeof(int32_t));
    if (!J) return 1;
    c->J[k] = J;
    uint64_t st = 0xC35CA0DA00000000ull + ((uint64_t)k << 20) + rold;
    for (uint64_t j = rold; j < rnew; j++) {
        for (uint32_t i = 0; i < d; i++) { I[j * d + i] = 0; J[j * d + i] = 0; }
        for (int attempt = 0; attempt < 64; attempt++) {
            for (uint32_t i = k; i < d; i++)                              /* interior nodes: the end nodes are the */
                J[j * d + i] = c->n[i] > 2 ? 1 + (int32_t)(mix64(&st) % (c->n[i] - 2))   /* faces, where an absorbing */
                                           : (int32_t)(mix64(&st) % c->n[i]);            /* problem is degenerate    */
            int dup = 0;
            for (uint64_t q = 0; q < j && !dup; q++) {
                int same = 1;
                for (uint32_t i = k; i < d && same; i++) same = J[q * d + i] == J[j * d + i];
                dup = same;
            }
            if (!dup) break;
        }
        for (uint32_t i = 0; i < k; i++) I[j * d + i] = (int32_t)(mix64(&st) % c->n[i]);
    }
    c->r[k] = rnew;
    return 0;
}

/* start ranks of the next solver step from the ranks the last one found: min(r+1, maxrank), the
 * vref branch of valuef_interp (src/valuefunc.c:637-648); index sets are kept (:706-712). */
int c3sc_cross_set_ranks(c3sc_cross *c, const uint64_t *ranks)
{
    if (!c || !ranks || ranks[0] != 1 || ranks[c->d] != 1) return C3SC_EINVAL;
    uint64_t r[C3SC_MAXD + 1];
    for (uint32_t k = 0; k <= c->d; k++) r[k] = ranks[k] ? ranks[k] : 1;
    for (uint32_t k = 1; k < c->d; k++) if (r[k] > r[k - 1] * c->n[k - 1]) r[k] = r[k - 1] * c->n[k - 1];
    for (uint32_t k = c->d - 1; k >= 1; k--) if (r[k] > r[k + 1] * c->n[k]) r[k] = r[k + 1] * c->n[k];
    for (uint32_t k = 1; k < c->d; k++)
        if (cross_resize(c, k, r[k])) return C3SC_EINVAL;
    return C3SC_OK;
}

/* ftapprox_cross_rankadapt: see the section header.  cores[k] needs n[k]*cap[k]*cap[k+1] doubles with
 * cap = min(maxrank, unfolding bounds) -- c3sc_cross_adapt_capacity() gives cap[]. */
static uint64_t adapt_maxrank(const c3sc_cross *c, const c3sc_adapt_opts *a)
{
    uint64_t minN = c->n[0];
    for (uint32_t k = 1; k < c->d; k++) if (c->n[k] < minN) minN = c->n[k];
    uint64_t mr = a && a->maxrank ? a->maxrank : minN;
    return mr < minN ? mr : minN;                                        /* src/valuefunc.c:625-631 */
}

int c3sc_cross_adapt_capacity(const c3sc_cross *c, const c3sc_adapt_opts *a, uint64_t *cap)
{
    if (!c || !cap) return C3SC_EINVAL;
    const uint64_t mr = adapt_maxrank(c, a);
    for (uint32_t k = 0; k <= c->d; k++) cap[k] = (k == 0 || k == c->d) ? 1 : (c->r[k] > mr ? c->r[k] : mr);
    return C3SC_OK;
}

int c3sc_cross_run_adapt(c3sc_cross *c, c3sc_fiber_batch_fn f, void *arg, const c3sc_cross_opts *opts,
                         const c3sc_adapt_opts *aopts, uint64_t *ranks_out, double *const *cores, uint64_t *nfibers,
                         double *rel_change)
{
    if (!c || !f || !aopts || !ranks_out || !cores) return C3SC_EINVAL;
    const uint32_t d = c->d;
    const uint64_t maxrank = adapt_maxrank(c, aopts);
    const uint32_t rounds = aopts->maxiter_adapt ? aopts->maxiter_adapt : 5;
    const int verbose = opts ? opts->verbose : 0;
    uint64_t cap[C3SC_MAXD + 1], total = 0;
    c3sc_cross_adapt_capacity(c, aopts, cap);
    double *raw[C3SC_MAXD] = {0};
    int rc = C3SC_OK;
    for (uint32_t k = 0; k < d; k++) {
        raw[k] = (double *)malloc(c->n[k] * cap[k] * cap[k + 1] * sizeof(double));
        if (!raw[k]) { rc = C3SC_EINVAL; goto done; }
    }
    for (uint32_t round = 0;; round++) {
        uint64_t nf = 0;
        rc = c3sc_cross_run(c, f, arg, opts, raw, &nf, rel_change);
        total += nf;
        if (rc) goto done;
        rc = c3sc_cores_round(d, c->n, c->r, (const double *const *)raw, aopts->round_tol, ranks_out, cores);
        if (rc) goto done;
        if (verbose) {
            fprintf(stderr, "c3sc_cross_run_adapt: round %u, cross ranks", round);
            for (uint32_t k = 0; k <= d; k++) fprintf(stderr, " %llu", (unsigned long long)c->r[k]);
            fprintf(stderr, " -> rounded");
            for (uint32_t k = 0; k <= d; k++) fprintf(stderr, " %llu", (unsigned long long)ranks_out[k]);
            fprintf(stderr, "\n");
        }
        if (aopts->kickrank == 0 || round + 1 >= rounds) break;
        int adapt = 0;
        for (uint32_t k = 1; k < d; k++) {
            if (ranks_out[k] != c->r[k] || c->r[k] >= maxrank) continue;  /* rounding cut it, or at the cap */
            uint64_t rnew = c->r[k] + aopts->kickrank;
            if (rnew > maxrank) rnew = maxrank;
            if (rnew > c->r[k - 1] * c->n[k - 1]) rnew = c->r[k - 1] * c->n[k - 1];
            if (rnew > c->r[k + 1] * c->n[k]) rnew = c->r[k + 1] * c->n[k];
            if (rnew <= c->r[k]) continue;
            if (cross_resize(c, k, rnew)) { rc = C3SC_EINVAL; goto done; }
            adapt = 1;
        }
        if (!adapt) break;
    }
done:
    if (nfibers) *nfibers = total;
    for (uint32_t k = 0; k < d; k++) free(raw[k]);
    return rc;
}

uint32_t c3sc_cross_dim(const c3sc_cross *c) { return c ? c->d : 0; }

/* ---- fiber memo (include/c3sc_cross.h) ---------------------------------------------------------------- */
struct c3sc_fiber_memo {
    uint32_t d;
    c3sc_fiber_batch_fn f;
    void *arg;
    size_t ldo;                    /* of the stored rows (fixed by the first call) */
    /* open-addressing table of slots into the stores */
    size_t cap, count;             /* cap = 0 or a power of two */
    int64_t *slot;                 /* -1 = empty */
    int32_t *keys;                 /* [count][d + 1]: dim_vary, fixed indices with the varying one zeroed */
    double *vals;                  /* [count][ldo] */
    size_t store_cap;
    /* scratch of a call: the descriptors of the misses, compacted */
    int32_t *mdv, *mfi;
    int64_t *rslot;                /* per request: its slot */
    size_t rcap;
    uint64_t requested, computed;
};

static uint64_t memo_hash(const int32_t *key, uint32_t n)
{
    uint64_t h = 0x9e3779b97f4a7c15ull;
    for (uint32_t i = 0; i < n; i++) { h ^= (uint64_t)(uint32_t)key[i]; h *= 0xff51afd7ed558ccdull; h ^= h >> 32; }
    return h;
}

int c3sc_fiber_memo_create(uint32_t d, c3sc_fiber_batch_fn f, void *arg, c3sc_fiber_memo **out)
{
    if (!out || !f || d < 1) return C3SC_EINVAL;
    c3sc_fiber_memo *m = (c3sc_fiber_memo *)calloc(1, sizeof *m);
    if (!m) return C3SC_EINVAL;
    m->d = d; m->f = f; m->arg = arg;
    *out = m;
    return C3SC_OK;
}

void c3sc_fiber_memo_clear(c3sc_fiber_memo *m)
{
    if (!m) return;
    m->count = 0;
    for (size_t i = 0; i < m->cap; i++) m->slot[i] = -1;
}

void c3sc_fiber_memo_destroy(c3sc_fiber_memo *m)
{
    if (!m) return;
    free(m->slot); free(m->keys); free(m->vals); free(m->mdv); free(m->mfi); free(m->rslot);
    free(m);
}

void c3sc_fiber_memo_stats(const c3sc_fiber_memo *m, uint64_t *requested, uint64_t *computed)
{
    if (requested) *requested = m ? m->requested : 0;
    if (computed) *computed = m ? m->computed : 0;
}

static int memo_grow_table(c3sc_fiber_memo *m, size_t want)
{
    size_t cap = m->cap ? m->cap : 1024;
    while (cap < 2 * want) cap *= 2;
    if (cap == m->cap) return 0;
    int64_t *s = (int64_t *)malloc(cap * sizeof *s);
    if (!s) return 1;
    for (size_t i = 0; i < cap; i++) s[i] = -1;
    for (size_t e = 0; e < m->count; e++) {
        size_t h = (size_t)memo_hash(m->keys + e * (m->d + 1), m->d + 1) & (cap - 1);
        while (s[h] >= 0) h = (h + 1) & (cap - 1);
        s[h] = (int64_t)e;
    }
    free(m->slot);
    m->slot = s; m->cap = cap;
    return 0;
}

static int memo_grow_store(c3sc_fiber_memo *m, size_t want)
{
    if (want <= m->store_cap) return 0;
    size_t cap = m->store_cap ? m->store_cap : 1024;
    while (cap < want) cap *= 2;
    int32_t *k = (int32_t *)realloc(m->keys, cap * (m->d + 1) * sizeof *k);
    if (!k) return 1;
    m->keys = k;
    double *v = (double *)realloc(m->vals, cap * m->ldo * sizeof *v);
    if (!v) return 1;
    m->vals = v; m->store_cap = cap;
    return 0;
}

int c3sc_fiber_memo_call(size_t F, const int32_t *dim_vary, const int32_t *fixed_ind, size_t ldo, double *out, void *memo)
{
    c3sc_fiber_memo *m = (c3sc_fiber_memo *)memo;
    if (!m || (F && (!dim_vary || !fixed_ind || !out))) return C3SC_EINVAL;
    if (F == 0) return C3SC_OK;
    const uint32_t d = m->d, kw = d + 1;
    if (m->count == 0) m->ldo = ldo;
    if (ldo != m->ldo) return C3SC_EINVAL;                     /* one row length per memo */
    if (F > m->rcap) {
        free(m->rslot); free(m->mdv); free(m->mfi);
        m->rslot = (int64_t *)malloc(F * sizeof *m->rslot);
        m->mdv = (int32_t *)malloc(F * sizeof *m->mdv);
        m->mfi = (int32_t *)malloc(F * d * sizeof *m->mfi);
        m->rcap = (m->rslot && m->mdv && m->mfi) ? F : 0;
        if (!m->rcap) return C3SC_EINVAL;
    }
    if (memo_grow_table(m, m->count + F) || memo_grow_store(m, m->count + F)) return C3SC_EINVAL;
    /* look every request up; a miss takes the next slot at once, so a repeat inside the batch finds it */
    size_t nmiss = 0;
    const size_t first_new = m->count;
    int32_t key[C3SC_MAXD + 1];
    for (size_t i = 0; i < F; i++) {
        const int32_t k = dim_vary[i];
        key[0] = k;
        for (uint32_t q = 0; q < d; q++) key[1 + q] = ((int32_t)q == k) ? 0 : fixed_ind[i * d + q];
        size_t h = (size_t)memo_hash(key, kw) & (m->cap - 1);
        int64_t s;
        while ((s = m->slot[h]) >= 0 && memcmp(m->keys + (size_t)s * kw, key, kw * sizeof(int32_t)) != 0) h = (h + 1) & (m->cap - 1);
        if (s < 0) {
            s = (int64_t)m->count++;
            m->slot[h] = s;
            memcpy(m->keys + (size_t)s * kw, key, kw * sizeof(int32_t));
            m->mdv[nmiss] = k;
            memcpy(m->mfi + nmiss * d, fixed_ind + i * d, d * sizeof(int32_t));
            nmiss++;
        }
        m->rslot[i] = s;
    }
    m->requested += F;
    m->computed += nmiss;
    int rc = C3SC_OK;
    if (nmiss == F) {
        /* nothing known: the operator writes straight into the caller's buffer (the driver's page-locked one) */
        rc = m->f(F, dim_vary, fixed_ind, ldo, out, m->arg);
        if (rc == C3SC_OK) memcpy(m->vals + first_new * ldo, out, F * ldo * sizeof(double));
    } else if (nmiss > 0) {
        /* the misses, compacted, through the front of the caller's buffer (nmiss < F rows; the driver's buffer is page-locked,
           a buffer of the memo's own would have to be page-locked per call, which costs more than a core batch) */
        rc = m->f(nmiss, m->mdv, m->mfi, ldo, out, m->arg);
        if (rc == C3SC_OK) memcpy(m->vals + first_new * ldo, out, nmiss * ldo * sizeof(double));
    }
    if (rc != C3SC_OK) {                                        /* forget what was not computed */
        m->count = first_new;
        for (size_t i = 0; i < m->cap; i++) if (m->slot[i] >= (int64_t)first_new) m->slot[i] = -1;
        return rc;
    }
    if (nmiss < F)
        for (size_t i = 0; i < F; i++) memcpy(out + i * ldo, m->vals + (size_t)m->rslot[i] * ldo, ldo * sizeof(double));
    return C3SC_OK;
}

/* A memo copies every value it stores (out of page-locked memory the device has just written): about 0.1 ms per 400-fiber
 * core batch, more than the batch costs on the GPU.  It pays when sweeps repeat fibers, i.e. from the second sweep pair on
 * (index sets that have settled ask for the same fibers again); a single sweep pair goes without.  c3sc_cross_uses_memo is
 * the rule, for callers that wrap their own operators. */
int c3sc_cross_uses_memo(const c3sc_cross_opts *opts) { return !opts || opts->maxiter != 1; }
static int memo_pays(const c3sc_cross_opts *opts) { return c3sc_cross_uses_memo(opts); }

/* ---- the GPU operators behind the driver ---------------------------------------------------------- */
struct vi_ctx { c3sc_problem *p; const c3sc_valuef *vf; };
static int vi_cb(size_t F, const int32_t *dv, const int32_t *fi, size_t ldo, double *out, void *arg)
{
    struct vi_ctx *x = (struct vi_ctx *)arg;
    return c3sc_vi_batch(x->p, x->vf, F, dv, fi, ldo, out, NULL);
}

struct pi_ctx { c3sc_problem *p; const c3sc_valuef *pol, *iter; double *rows; size_t cap; };
static int pi_cb(size_t F, const int32_t *dv, const int32_t *fi, size_t ldo, double *out, void *arg)
{
    struct pi_ctx *x = (struct pi_ctx *)arg;
    /* the index sets move between sweeps, so the policy rows are rebuilt for every request
       (improvement against `pol`, evaluation against `iter`, src/bellman.c:1831-1871) */
    return c3sc_pi_batch(x->p, x->pol, x->iter, F, dv, fi, ldo, 0, x->rows, NULL, out);
}

/* one c3control_step_pi (src/bellman.c:2214-2262): next = cross(bellman_pi(.; policy of vf_policy, vf_iter)) */
int c3sc_cross_run_pi(c3sc_cross *c, c3sc_problem *p, const c3sc_valuef *vf_policy, const c3sc_valuef *vf_iter,
                      uint32_t dx, const c3sc_cross_opts *opts, double *const *cores, uint64_t *nfibers, double *rel_change)
{
    size_t fmax = 1;
    if (!c) return C3SC_EINVAL;
    c->want_pinned = 1;
    for (uint32_t k = 0; k < c->d; k++)
        if (c->r[k] * c->r[k + 1] > fmax) fmax = c->r[k] * c->r[k + 1];
    struct pi_ctx x = {p, vf_policy, vf_iter, NULL, 0};
    x.rows = (double *)malloc(fmax * c->nmax * (2 * (size_t)dx + 3) * sizeof(double));
    if (!x.rows) return C3SC_EINVAL;
    c3sc_fiber_memo *memo = NULL;
    int rc = C3SC_OK;
    if (!memo_pays(opts)) rc = c3sc_cross_run(c, pi_cb, &x, opts, cores, nfibers, rel_change);
    else {
        rc = c3sc_fiber_memo_create(c->d, pi_cb, &x, &memo);
        if (rc == C3SC_OK) rc = c3sc_cross_run(c, c3sc_fiber_memo_call, memo, opts, cores, nfibers, rel_change);
        c3sc_fiber_memo_destroy(memo);
    }
    free(x.rows);
    return rc;
}

/* one c3control_step_vi (src/bellman.c:2177-2211): next = cross(bellman_vi(.; vf)) */
int c3sc_cross_run_vi(c3sc_cross *c, c3sc_problem *p, const c3sc_valuef *vf, const c3sc_cross_opts *opts,
                      double *const *cores, uint64_t *nfibers, double *rel_change)
{
    struct vi_ctx x = {p, vf};
    if (!c) return C3SC_EINVAL;
    c->want_pinned = 1;
    /* the operator (a backup against the fixed vf) is the same in every sweep of this call: repeated fibers come from the memo */
    if (!memo_pays(opts)) return c3sc_cross_run(c, vi_cb, &x, opts, cores, nfibers, rel_change);
    c3sc_fiber_memo *memo = NULL;
    int rc = c3sc_fiber_memo_create(c->d, vi_cb, &x, &memo);
    if (rc == C3SC_OK) rc = c3sc_cross_run(c, c3sc_fiber_memo_call, memo, opts, cores, nfibers, rel_change);
    c3sc_fiber_memo_destroy(memo);
    return rc;
}

/* c3control_vi_solve (src/bellman.c:2282-2340): iterate next = cross(bellman_vi(.; current)) until the l2
 * difference between iterates drops below abs_conv_tol or maxiter steps were taken.  cores0 / ranks0: the
 * start value function (any ranks); cores_out: n[k]*r[k]*r[k+1] doubles each with the driver's ranks. */
int c3sc_vi_solve(c3sc_cross *c, c3sc_problem *p, const uint64_t *ranks0, const double *const *cores0, uint32_t maxiter,
                  double abs_conv_tol, const c3sc_cross_opts *opts, double *const *cores_out, uint32_t *iters_done,
                  double *last_diff, uint64_t *nfibers)
{
    if (!c || !p || !ranks0 || !cores0 || !cores_out) return C3SC_EINVAL;
    const uint32_t d = c->d;
    c3sc_valuef *vf = NULL;
    int rc = c3sc_valuef_create(d, c->n, ranks0, cores0, &vf);
    if (rc) return rc;
    double **cur = (double **)calloc(d, sizeof(double *));
    uint64_t rcur[C3SC_MAXD + 1], total = 0;
    if (!cur) { c3sc_valuef_destroy(vf); return C3SC_EINVAL; }
    for (uint32_t k = 0; k <= d; k++) rcur[k] = ranks0[k];
    for (uint32_t k = 0; k < d; k++) {
        size_t len0 = c->n[k] * ranks0[k] * ranks0[k + 1], len1 = c->n[k] * c->r[k] * c->r[k + 1];
        cur[k] = (double *)malloc((len0 > len1 ? len0 : len1) * sizeof(double));
        if (!cur[k]) { rc = C3SC_EINVAL; goto out; }
        memcpy(cur[k], cores0[k], len0 * sizeof(double));
    }
    uint32_t it = 0;
    double diff = 0.0;
    for (; it < maxiter; it++) {
        uint64_t nf = 0;
        rc = c3sc_cross_run_vi(c, p, vf, opts, cores_out, &nf, NULL);
        if (rc) goto out;
        total += nf;
        diff = c3sc_cores_norm2diff(d, c->n, rcur, (const double *const *)cur, c->r, (const double *const *)cores_out);
        if (opts && opts->verbose)
            fprintf(stderr, "c3sc_vi_solve: iteration %u, l2 difference %.5e, l2 norm %.5e\n", it + 1, diff,
                    c3sc_cores_norm(d, c->n, c->r, (const double *const *)cores_out));
        int same = 1;
        for (uint32_t k = 0; k <= d; k++) same = same && rcur[k] == c->r[k];
        for (uint32_t k = 0; k < d; k++) memcpy(cur[k], cores_out[k], c->n[k] * c->r[k] * c->r[k + 1] * sizeof(double));
        for (uint32_t k = 0; k <= d; k++) rcur[k] = c->r[k];
        if (same) rc = c3sc_valuef_update(vf, (const double *const *)cur);
        else { c3sc_valuef_destroy(vf); vf = NULL; rc = c3sc_valuef_create(d, c->n, c->r, (const double *const *)cur, &vf); }
        if (rc) goto out;
        if (diff < abs_conv_tol) { it++; break; }
    }
    if (iters_done) *iters_done = it;
    if (last_diff) *last_diff = diff;
    if (nfibers) *nfibers = total;
out:
    for (uint32_t k = 0; k < d; k++) free(cur[k]);
    free(cur);
    c3sc_valuef_destroy(vf);
    return rc;
}

/* c3control_step_vi with the adapt == 1 branch of valuef_interp */
int c3sc_cross_run_vi_adapt(c3sc_cross *c, c3sc_problem *p, const c3sc_valuef *vf, const c3sc_cross_opts *opts,
                            const c3sc_adapt_opts *aopts, uint64_t *ranks_out, double *const *cores, uint64_t *nfibers,
                            double *rel_change)
{
    struct vi_ctx x = {p, vf};
    c3sc_fiber_memo *memo = NULL;                            /* one operator for all the cross runs of the adaptive loop */
    int rc = c3sc_fiber_memo_create(c->d, vi_cb, &x, &memo);
    if (rc == C3SC_OK) rc = c3sc_cross_run_adapt(c, c3sc_fiber_memo_call, memo, opts, aopts, ranks_out, cores, nfibers, rel_change);
    c3sc_fiber_memo_destroy(memo);
    return rc;
}
