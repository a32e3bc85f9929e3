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
#define C3SC_CLONES __attribute__((target_clones("avx2", "default")))   /* same arithmetic, wider registers (no FMA) */
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
 * calls: with the backup on the GPU they are what a sweep costs.  The three phases share one OpenMP parallel region
 * (the reference is an OpenMP library itself, src/bellman.c:1390-1404).  Rows are cut into LA_NB fixed blocks; a
 * thread owns a contiguous run of blocks and is the only one that ever writes their rows, so the unfolding stays in
 * its core's cache from the first phase to the last.  Everything that couples rows (column norms, reflector dots,
 * pivot searches) is reduced per BLOCK, published, and summed by every thread in block order after a barrier: the
 * numbers depend on LA_NB, never on the number of threads (C3SC_HOST_THREADS, default min(8, omp_get_max_threads())),
 * and every thread takes the same branches.  Partial results are double-buffered where the next phase would
 * otherwise overwrite them before a slow thread has read them. */
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

typedef struct {
    size_t mcap, ncap;        /* capacity: rows, columns */
    size_t m, n, bs;          /* this step: rows, columns, rows per block */
    int threads;
    double *Q, *B;            /* m x n column-major: unfolding in / orthonormal basis out; cross core Q inv(Q[P]) out */
    size_t *P;                /* n pivot rows out */
    char *skip, *used;        /* m */
    double *part;             /* [2][LA_NB][ncap] partial sums */
    double *mp;               /* [2][LA_NB][2] pivot search: max over all rows, max over eligible rows */
    size_t *ip;               /* [3][LA_NB] pivot search: row / entry */
    double *priv;             /* [threads][8 ncap] */
    double *S, *aug;          /* n x n, n x 2n (thread 0) */
    double *k1; int64_t *qk; uint32_t *slot; size_t slotcap;   /* twin rows */
    size_t sh_bi, sh_bj; double sh_best;                      /* maxvol: masked scan result of thread 0 */
    int rc;
#ifdef _OPENMP
    char pad0[64]; atomic_int bar_count;                      /* sense-reversing barrier of the team, a cache line each */
    char pad1[60]; atomic_int bar_sense;
    char pad2[60];
#endif
} la_ws;
typedef struct { int tid, nt, b0, b1, sense; } la_thr;       /* a thread of the team and the blocks it owns */

/* A step has ~150 barriers with a few microseconds of work between them: the team spins (libgomp's barrier parks
 * threads in the kernel under some team sizes, 50 us a time); a thread that has lost its CPU is waited for with
 * sched_yield. */
static inline void la_barrier(la_ws *w, la_thr *t)
{
#ifdef _OPENMP
    if (t->nt == 1) return;
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
#else
    (void)w; (void)t;
#endif
}
#define LA_BARRIER() la_barrier(w, th)

static void la_ws_free(la_ws *w)
{
    if (!w) return;
    free(w->Q); free(w->B); free(w->P); free(w->skip); free(w->used); free(w->part); free(w->mp); free(w->ip); free(w->priv);
    free(w->S); free(w->aug); free(w->k1); free(w->qk); free(w->slot); free(w);
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

static la_ws *la_ws_create(size_t mcap, size_t ncap)
{
    la_ws *w = (la_ws *)calloc(1, sizeof *w);
    if (!w) return NULL;
    w->mcap = mcap; w->ncap = ncap; w->threads = la_threads();
    w->slotcap = 16;
    while (w->slotcap < 4 * mcap) w->slotcap <<= 1;
    w->Q = (double *)malloc(mcap * ncap * sizeof(double));
    w->B = (double *)malloc(mcap * ncap * sizeof(double));
    w->P = (size_t *)malloc(ncap * sizeof(size_t));
    w->skip = (char *)malloc(mcap + 1); w->used = (char *)malloc(mcap + 1);
    w->part = (double *)malloc(2 * LA_NB * ncap * sizeof(double));
    w->mp = (double *)malloc(2 * LA_NB * 2 * sizeof(double));
    w->ip = (size_t *)malloc(3 * LA_NB * sizeof(size_t));
    w->priv = (double *)malloc((size_t)w->threads * 8 * ncap * sizeof(double));
    w->S = (double *)malloc(ncap * ncap * sizeof(double));
    w->aug = (double *)malloc(2 * ncap * ncap * sizeof(double));
    w->k1 = (double *)malloc(2 * mcap * sizeof(double));
    w->qk = (int64_t *)malloc(mcap * sizeof(int64_t));
    w->slot = (uint32_t *)malloc(w->slotcap * sizeof(uint32_t));
    if (!w->Q || !w->B || !w->P || !w->skip || !w->used || !w->part || !w->mp || !w->ip || !w->priv || !w->S || !w->aug || !w->k1 ||
        !w->qk || !w->slot) { la_ws_free(w); return NULL; }
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
static inline int la_mine(const la_ws *w, size_t row, int b0, int b1)
{
    const int blk = (int)(row / w->bs);
    return blk >= b0 && blk < b1;
}
#define PART(buf, blk, j) part[((size_t)(buf) * LA_NB + (size_t)(blk)) * ncap + (j)]

/* Rows of the unfolding A (m x n, before the QR) that repeat an earlier row to round-off carry no
 * information for the pivoting (absorbing faces with a constant boundary cost produce whole families of
 * them); when the unfolding is rank-deficient the QR completes Q with arbitrary directions and maxvol would
 * happily pick such twins, which makes the NEXT unfolding rank-deficient as well.  skip[i] = 1 withholds
 * row i from the pivoting.  Twins are found through two fixed random projections of the rows (every thread
 * projects its own rows; thread 0 then walks the rows in order through a hash of the quantised projection while
 * the others go ahead into the QR, which does not need skip[]). */
static void team_twin_rows(la_ws *w, la_thr *th)
{
    const int tid = th->tid, b0 = th->b0, b1 = th->b1;
    const size_t m = w->m, n = w->n, ncap = w->ncap;
    const double *A = w->Q;
    double *k1 = w->k1, *k2 = w->k1 + m, *part = w->part;
    for (int blk = b0; blk < b1; blk++) {
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
        PART(1, blk, 0) = scale;                        /* buffer 1: the QR starts on buffer 0 */
    }
    LA_BARRIER();
    if (tid != 0) return;
    double scale = 0.0;
    for (int blk = 0; blk < LA_NB; blk++) if (PART(1, blk, 0) > scale) scale = PART(1, blk, 0);
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
 * Two barriers per reflector: one for the exact norm of the pivot column, one for its dots with the trailing
 * columns.  Row k of the trailing columns is read by everyone between the two and written by its owner after
 * the second; every thread keeps its own copy of tau and of the downdated column norms (LAPACK dgeqp3 style,
 * recomputed exactly once they have lost six digits). */
static void team_qr_basis(la_ws *w, la_thr *th)
{
    const int tid = th->tid, b0 = th->b0, b1 = th->b1;
    const size_t m = w->m, n = w->n, ncap = w->ncap;
    double *A = w->Q, *part = w->part;
    double *pv = w->priv + (size_t)tid * 8 * ncap;
    double *tau = pv, *vn2 = pv + ncap, *vref = pv + 2 * ncap, *rowk = pv + 3 * ncap, *sc = pv + 4 * ncap;
    size_t *redo = (size_t *)(pv + 5 * ncap);
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
    double ref = 0.0;
    for (size_t k = 0; k < n; k++) {
        size_t p = k; double best = -1.0;
        for (size_t j = k; j < n; j++) if (vn2[j] > best) { best = vn2[j]; p = j; }
        for (size_t j = k; j < p; j++) if (vn2[j] >= best * (1.0 - TIE_EPS)) { p = j; break; }
        double *ak = A + k * m;
        if (p != k) {
            double *ap = A + p * m;
            for (int blk = b0; blk < b1; blk++) {
                size_t lo, hi;
                la_rows(w, blk, 0, &lo, &hi);
                for (size_t i = lo; i < hi; i++) { const double t = ak[i]; ak[i] = ap[i]; ap[i] = t; }
            }
            double t = vn2[k]; vn2[k] = vn2[p]; vn2[p] = t;
            t = vref[k]; vref[k] = vref[p]; vref[p] = t;
        }
        for (int blk = b0; blk < b1; blk++) {
            size_t lo, hi;
            la_rows(w, blk, k, &lo, &hi);
            PART(pb, blk, 0) = dot8(ak + lo, ak + lo, hi - lo);
        }
        LA_BARRIER();
        double nrm = 0.0;
        for (int blk = 0; blk < LA_NB; blk++) nrm += PART(pb, blk, 0);
        pb ^= 1;
        nrm = sqrt(nrm);
        if (k == 0) ref = nrm;
        if (nrm <= RANK_EPS * ref || nrm == 0.0) { rank = k; break; }
        const double akk = ak[k];
        const double alpha = akk >= 0.0 ? -nrm : nrm;
        const double v0 = akk - alpha;
        tau[k] = -v0 / alpha;
        for (size_t j = k + 1; j < n; j++) rowk[j] = A[k + j * m];
        for (int blk = b0; blk < b1; blk++) {
            size_t lo, hi;
            la_rows(w, blk, k + 1, &lo, &hi);
            for (size_t i = lo; i < hi; i++) ak[i] /= v0;
            for (size_t j = k + 1; j < n; j++) PART(pb, blk, j) = dot8(ak + lo, A + j * m + lo, hi - lo);
        }
        LA_BARRIER();
        for (size_t j = k + 1; j < n; j++) {
            double s = 0.0;
            for (int blk = 0; blk < LA_NB; blk++) s += PART(pb, blk, j);
            sc[j] = (rowk[j] + s) * tau[k];
        }
        pb ^= 1;
        if (la_mine(w, k, b0, b1)) {
            ak[k] = alpha;
            for (size_t j = k + 1; j < n; j++) A[k + j * m] = rowk[j] - sc[j];
        }
        for (int blk = b0; blk < b1; blk++) {
            size_t lo, hi;
            la_rows(w, blk, k + 1, &lo, &hi);
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
    for (size_t k = rank; k < n; k++) tau[k] = 0.0;
    /* accumulate Q = H_0 .. H_{n-1} [I; 0] in place, last reflector first.  Column kk still holds its reflector
       below the diagonal while the columns after it (already Q columns, zero in rows <= kk) are reflected. */
    for (size_t kk = n; kk-- > 0;) {
        double *ak = A + kk * m;
        const double t = tau[kk];
        if (t != 0.0 && kk + 1 < n) {
            for (int blk = b0; blk < b1; blk++) {
                size_t lo, hi;
                la_rows(w, blk, kk + 1, &lo, &hi);
                for (size_t j = kk + 1; j < n; j++) PART(pb, blk, j) = dot8(ak + lo, A + j * m + lo, hi - lo);
            }
            LA_BARRIER();
            for (size_t j = kk + 1; j < n; j++) {
                double s = 0.0;
                for (int blk = 0; blk < LA_NB; blk++) s += PART(pb, blk, j);
                sc[j] = s * t;
            }
            pb ^= 1;
            for (int blk = b0; blk < b1; blk++) {
                size_t lo, hi;
                la_rows(w, blk, kk + 1, &lo, &hi);
                for (size_t j = kk + 1; j < n; j++) axpy(-sc[j], ak + lo, A + j * m + lo, hi - lo);
            }
            if (la_mine(w, kk, b0, b1))
                for (size_t j = kk + 1; j < n; j++) A[kk + j * m] = -sc[j];
        }
        for (int blk = b0; blk < b1; blk++) {           /* column kk itself: H_kk e_kk = e_kk - tau v */
            size_t lo, hi;
            la_rows(w, blk, 0, &lo, &hi);
            for (size_t i = lo; i < hi; i++) ak[i] = i < kk ? 0.0 : (i == kk ? 1.0 - t : (t != 0.0 ? -t * ak[i] : 0.0));
        }
    }
}

/* thread 0: S = inv(Q[P,:]) by Gauss-Jordan with partial pivoting on [Q[P] | I] */
static int invert_pivot_block(la_ws *w)
{
    const size_t m = w->m, n = w->n;
    const double *Q = w->Q;
    double *aug = w->aug, *S = w->S;                    /* row-major n x 2n */
    for (size_t a = 0; a < n; a++)
        for (size_t b = 0; b < n; b++) {
            aug[a * 2 * n + b] = Q[w->P[a] + b * m];
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

/* maxvol: rows P of Q (m x n) with |det Q[P]| locally maximal, and B = Q inv(Q[P]).
 * Symmetric problems (V(x) = V(-x)) make mirrored rows tie exactly in exact arithmetic; which one wins would
 * then depend on the last bits of the operator's values.  Entries within TIE_EPS of the maximum count as
 * tied and the first in scan order wins, so two operators that agree to round-off pick the same rows.
 * Start rows: Gaussian elimination with row pivoting on a copy (two barriers per column: the block maxima, then
 * the first near-maximal row); the chosen row is read by everyone and left alone by its owner from then on.
 * Swaps while some |B[i,j]| > 1 + delta: two barriers per swap (the first near-maximal entry in column-major
 * order; the per-block column maxima of the updated B).  The owner of the entering row holds its new values back
 * until the others have read the old ones. */
static void team_maxvol(la_ws *w, la_thr *th)
{
    const int tid = th->tid, b0 = th->b0, b1 = th->b1;
    const size_t m = w->m, n = w->n, ncap = w->ncap;
    const double *Q = w->Q;
    double *B = w->B, *part = w->part, *mp = w->mp;
    size_t *ip = w->ip;
    const char *skip = w->skip;
    char *used = w->used;
    double *pv = w->priv + (size_t)tid * 8 * ncap;
    double *f = pv, *cmax = pv + ncap, *col = pv + 2 * ncap, *newrow = pv + 3 * ncap;
    for (int blk = b0; blk < b1; blk++) {
        size_t lo, hi;
        la_rows(w, blk, 0, &lo, &hi);
        memset(used + lo, 0, hi - lo);
        for (size_t j = 0; j < n; j++) memcpy(B + j * m + lo, Q + j * m + lo, (hi - lo) * sizeof(double));
    }
    for (size_t j = 0; j < n; j++) {
        const double *bj = B + j * m;
        for (int blk = b0; blk < b1; blk++) {
            size_t lo, hi, piv = LA_NONE;
            la_rows(w, blk, 0, &lo, &hi);
            double best = -1.0, best_any = 0.0;
            for (size_t i = lo; i < hi; i++) {
                if (used[i]) continue;
                const double a = fabs(bj[i]);
                if (a > best_any) best_any = a;
                if (!skip[i] && a > best) { best = a; piv = i; }
            }
            mp[(0 * LA_NB + blk) * 2] = best_any; mp[(0 * LA_NB + blk) * 2 + 1] = best; ip[0 * LA_NB + blk] = piv;
        }
        LA_BARRIER();
        size_t piv = 0; double best = -1.0, best_any = 0.0;
        for (int blk = 0; blk < LA_NB; blk++) {
            if (mp[blk * 2] > best_any) best_any = mp[blk * 2];
            if (mp[blk * 2 + 1] > best) { best = mp[blk * 2 + 1]; piv = ip[blk]; }
        }
        const int all_rows = best < 1e-6 * best_any;    /* the distinct rows do not reach this direction */
        if (all_rows) {
            for (int blk = b0; blk < b1; blk++) {
                size_t lo, hi, pb_ = LA_NONE;
                la_rows(w, blk, 0, &lo, &hi);
                double bb = -1.0;
                for (size_t i = lo; i < hi; i++)
                    if (!used[i] && fabs(bj[i]) > bb) { bb = fabs(bj[i]); pb_ = i; }
                mp[(1 * LA_NB + blk) * 2] = bb; ip[1 * LA_NB + blk] = pb_;
            }
            LA_BARRIER();
            best = -1.0;
            for (int blk = 0; blk < LA_NB; blk++)
                if (mp[(LA_NB + blk) * 2] > best) { best = mp[(LA_NB + blk) * 2]; piv = ip[LA_NB + blk]; }
        }
        for (int blk = b0; blk < b1; blk++) {           /* TIE_EPS: first row among the near-maximal ones */
            size_t lo, hi, first = LA_NONE;
            la_rows(w, blk, 0, &lo, &hi);
            if (hi > piv) hi = piv;
            for (size_t i = lo; i < hi; i++)
                if (!used[i] && (all_rows || !skip[i]) && fabs(bj[i]) >= best * (1.0 - TIE_EPS)) { first = i; break; }
            ip[2 * LA_NB + blk] = first;
        }
        LA_BARRIER();
        for (int blk = 0; blk < LA_NB; blk++)
            if (ip[2 * LA_NB + blk] != LA_NONE) { piv = ip[2 * LA_NB + blk]; break; }
        if (tid == 0) w->P[j] = piv;
        if (la_mine(w, piv, b0, b1)) used[piv] = 1;
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
    if (tid == 0) w->rc = invert_pivot_block(w);
    LA_BARRIER();
    if (w->rc) return;
    const double *S = w->S;
    for (int blk = b0; blk < b1; blk++) {               /* B[:,b] = sum_a S[a,b] Q[:,a] */
        size_t lo, hi;
        la_rows(w, blk, 0, &lo, &hi);
        for (size_t b = 0; b < n; b++) {
            double *bb = B + b * m;
            for (size_t i = lo; i < hi; i++) bb[i] = 0.0;
            for (size_t a = 0; a < n; a++) axpy(S[a + b * n], Q + a * m + lo, bb + lo, hi - lo);
            PART(0, blk, b) = absmax(bb + lo, hi - lo);
        }
    }
    size_t held = LA_NONE;                               /* entering row whose new values this thread holds back */
    int it = 0;
    for (; it < 200; it++) {
        LA_BARRIER();
        if (held != LA_NONE) { for (size_t b = 0; b < n; b++) B[held + b * m] = newrow[b]; held = LA_NONE; }
        double best = 0.0;
        for (size_t j = 0; j < n; j++) {
            double c = 0.0;
            for (int blk = 0; blk < LA_NB; blk++) if (PART(0, blk, j) > c) c = PART(0, blk, j);
            cmax[j] = c;
            if (c > best) best = c;
        }
        if (best <= 1.0 + 1e-2) break;
        const double thr = best * (1.0 - TIE_EPS);
        for (int blk = b0; blk < b1; blk++) {           /* first near-maximal entry, column-major order */
            size_t lo, hi, key = LA_NONE;
            la_rows(w, blk, 0, &lo, &hi);
            for (size_t j = 0; j < n && key == LA_NONE; j++) {
                if (PART(0, blk, j) < thr) continue;
                for (size_t i = lo; i < hi; i++)
                    if (fabs(B[i + j * m]) >= thr && !skip[i]) { key = j * m + i; break; }
            }
            ip[blk] = key;
        }
        LA_BARRIER();
        size_t key = LA_NONE;
        for (int blk = 0; blk < LA_NB; blk++) if (ip[blk] < key) key = ip[blk];
        size_t bi, bj;
        if (key != LA_NONE) { bi = key % m; bj = key / m; }
        else {                                           /* the maxima sit on withheld rows: masked scan */
            if (tid == 0) {
                double bm = 0.0; size_t xi = 0, xj = 0;
                for (size_t j = 0; j < n; j++)
                    for (size_t i = 0; i < m; i++)
                        if (!skip[i] && fabs(B[i + j * m]) > bm * (1.0 + TIE_EPS)) { bm = fabs(B[i + j * m]); xi = i; xj = j; }
                w->sh_best = bm; w->sh_bi = xi; w->sh_bj = xj;
            }
            LA_BARRIER();
            if (w->sh_best <= 1.0 + 1e-2) break;
            bi = w->sh_bi; bj = w->sh_bj;
        }
        /* row bi replaces P[bj]:  B <- B - B[:,bj] (B[bi,:] - e_bj) / B[bi,bj] */
        const double pvv = B[bi + bj * m];
        for (size_t b = 0; b < n; b++) col[b] = (B[bi + b * m] - (b == bj ? 1.0 : 0.0)) / pvv;
        const double fs = 1.0 - col[bj];
        const int mine = la_mine(w, bi, b0, b1);
        if (mine) {
            for (size_t b = 0; b < n; b++)
                newrow[b] = b == bj ? pvv * fs : (col[b] == 0.0 ? B[bi + b * m] : B[bi + b * m] + (-col[b]) * pvv);
            held = bi;
        }
        for (int blk = b0; blk < b1; blk++) {
            size_t lo, hi;
            la_rows(w, blk, 0, &lo, &hi);
            const int has = bi >= lo && bi < hi;
            for (size_t b = 0; b < n; b++) {
                if (b == bj || col[b] == 0.0) continue;
                double c = has ? fabs(newrow[b]) : 0.0;
                for (int seg = 0; seg < 2; seg++) {
                    size_t s0 = lo, s1 = hi;
                    if (has) { if (seg == 0) s1 = bi; else s0 = bi + 1; }
                    else if (seg == 1) break;
                    const double cs = axpy_absmax(-col[b], B + bj * m + s0, B + b * m + s0, s1 - s0);
                    if (cs > c) c = cs;
                }
                PART(0, blk, b) = c;
            }
            double *bb = B + bj * m;
            for (size_t i = lo; i < hi; i++) if (i != bi) bb[i] *= fs;
            double c;                                    /* row bi still holds the old pivot; its new value is held back */
            if (has) {
                c = fabs(newrow[bj]);
                const double c0 = absmax(bb + lo, bi - lo), c1 = absmax(bb + bi + 1, hi - bi - 1);
                if (c0 > c) c = c0;
                if (c1 > c) c = c1;
            } else c = absmax(bb + lo, hi - lo);
            PART(0, blk, bj) = c;
        }
        if (tid == 0) w->P[bj] = bi;
    }
    if (it == 200) {                                     /* swap limit: the last entering row is still held back */
        LA_BARRIER();
        if (held != LA_NONE) for (size_t b = 0; b < n; b++) B[held + b * m] = newrow[b];
    }
}

/* twin rows + QR basis + maxvol of the unfolding in w->Q (m x n): w->Q <- basis, w->B <- cross core, w->P <- rows */
static int pivot_step(la_ws *w, size_t m, size_t n)
{
    if (m > w->mcap || n > w->ncap || n > m) return 1;
    w->m = m; w->n = n; w->rc = 0;
#ifdef _OPENMP
    atomic_store(&w->bar_count, 0); atomic_store(&w->bar_sense, 0);
#endif
    w->bs = ((m + LA_NB - 1) / LA_NB + 7) & ~(size_t)7;
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
        team_twin_rows(w, &th);
        team_qr_basis(w, &th);
        team_maxvol(w, &th);
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
/* the two unfoldings of a core, straight from the operator's fiber values, and the way back into the
 * valuef_precompute_cores layout (block j column-major: core[j*rk*rk1 + a + b*rk]) */
static void unfold_left(const double *vals, size_t ldo, size_t rk, size_t N, size_t rk1, double *Q /* rows (a,j) = a + j*rk, cols b */)
{
    for (size_t b = 0; b < rk1; b++)
        for (size_t a = 0; a < rk; a++) {
            const double *v = vals + (a + b * rk) * ldo;
            double *q = Q + a + b * rk * N;
            for (size_t j = 0; j < N; j++) q[j * rk] = v[j];
        }
}
static void unfold_right(const double *vals, size_t ldo, size_t rk, size_t N, size_t rk1, double *Q /* rows (j,b) = j + b*N, cols a */)
{
    for (size_t a = 0; a < rk; a++)
        for (size_t b = 0; b < rk1; b++) memcpy(Q + b * N + a * N * rk1, vals + (a + b * rk) * ldo, N * sizeof(double));
}
static void store_left(const double *B /* (a + j*rk) + b*rk*N */, size_t rk, size_t N, size_t rk1, double *core)
{
    for (size_t j = 0; j < N; j++)
        for (size_t b = 0; b < rk1; b++) memcpy(core + j * rk * rk1 + b * rk, B + j * rk + b * rk * N, rk * sizeof(double));
}
static void store_right(const double *B /* (j + b*N) + a*N*rk1 */, size_t rk, size_t N, size_t rk1, double *core)
{
    for (size_t a = 0; a < rk; a++)
        for (size_t b = 0; b < rk1; b++) {
            const double *s = B + b * N + a * N * rk1;
            double *t = core + a + b * rk;
            for (size_t j = 0; j < N; j++) t[j * rk * rk1] = s[j];
        }
}
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
    double *Q = ws ? ws->Q : NULL, *B = ws ? ws->B : NULL;
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
            unfold_left(vals, ldo, rk, N, rk1, Q);
            { const double t_ = now_s(); const int mv = pivot_step(ws, m, rk1); g_t_piv += now_s() - t_; if (mv) { rc = C3SC_ENUMERIC; goto done; } }
            for (size_t b = 0; b < rk1; b++) {                      /* new left set: row (a,j) = a + j*rk */
                const size_t a = P[b] % rk, j = P[b] / rk;
                for (uint32_t i = 0; i < k; i++) tmpI[b * d + i] = c->I[k][a * d + i];
                tmpI[b * d + k] = (int32_t)j;
            }
            for (size_t b = 0; b < rk1; b++)
                for (uint32_t i = 0; i <= k; i++) c->I[k + 1][b * d + i] = tmpI[b * d + i];
            store_left(B, rk, N, rk1, cores[k]);
        }
        rc = eval_core(c, d - 1, f, arg, dv, fi, vals, &nfib);
        if (rc) goto done;
        store_vals(vals, ldo, c->r[d - 1], c->n[d - 1], 1, cores[d - 1]);
        /* ---- right -> left: fix the right index sets J[1..d-1] ---------------------------------- */
        for (uint32_t k = d - 1; k >= 1; k--) {
            const size_t rk = c->r[k], rk1 = c->r[k + 1], N = c->n[k], m = N * rk1;
            rc = eval_core(c, k, f, arg, dv, fi, vals, &nfib);
            if (rc) goto done;
            unfold_right(vals, ldo, rk, N, rk1, Q);                 /* rows (j,b) = j + b*N, cols a */
            { const double t_ = now_s(); const int mv = pivot_step(ws, m, rk); g_t_piv += now_s() - t_; if (mv) { rc = C3SC_ENUMERIC; goto done; } }
            for (size_t a = 0; a < rk; a++) {
                const size_t j = P[a] % N, b = P[a] / N;
                tmpI[a * d + k] = (int32_t)j;
                for (uint32_t i = k + 1; i < d; i++) tmpI[a * d + i] = c->J[k + 1][b * d + i];
            }
            for (size_t a = 0; a < rk; a++)
                for (uint32_t i = k; i < d; i++) c->J[k][a * d + i] = tmpI[a * d + i];
            store_right(B, rk, N, rk1, cores[k]);
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
            if (verbose) fprintf(stderr, "c3sc_cross: sweep %u, |T|=%g, rel change %g, fibers %llu; cumulative s: operator %.4f pivoting (twin rows + qr + maxvol, %d threads) %.4f norms %.4f\n",
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
    int rc = c3sc_cross_run(c, pi_cb, &x, opts, cores, nfibers, rel_change);
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
    return c3sc_cross_run(c, vi_cb, &x, opts, cores, nfibers, rel_change);
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
    return c3sc_cross_run_adapt(c, vi_cb, &x, opts, aopts, ranks_out, cores, nfibers, rel_change);
}
