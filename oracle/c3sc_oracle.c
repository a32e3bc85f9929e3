/* c3sc_oracle.c -- CPU restatement of the c3sc Bellman-backup path.
 * TEST INFRASTRUCTURE ONLY (see c3sc_oracle.h).  Each function names the
 * /root/reference file:line whose arithmetic and operation ORDER it follows.
 * Compile with -ffp-contract=off (the reference is built -std=c99 on x86-64
 * without -march, i.e. no fused multiply-add, CMakeLists.txt:37).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "c3sc_oracle.h"

#define ORC_MAXD 32

/* ---- grid constants: bellman.c:1975-1986 and :181-186 ------------------ */
void orc_grid_constants(size_t dx, const size_t *ngrid, double *const *xgrid,
                        const double *lb, const double *ub,
                        double *h, double *hmin, double *h2, double *t)
{
    (void)ngrid;
    double hm = ub[0] - lb[0];
    for (size_t i = 0; i < dx; i++) {
        h[i] = xgrid[i][1] - xgrid[i][0];
        if (h[i] < hm) hm = h[i];
    }
    *hmin = hm;
    *h2 = hm * hm;
    for (size_t i = 0; i < dx; i++) {
        t[2 * i] = *h2 / h[i];
        t[2 * i + 1] = t[2 * i] / h[i];
    }
}

/* ---- x -> index: nodeutil.c:408-419 (linear scan, |x-g|<1e-14) ---------- */
size_t orc_x_to_ind(double x, size_t n, const double *grid)
{
    for (size_t j = 0; j < n; j++)
        if (fabs(x - grid[j]) < 1e-14) return j;
    return n;
}

/* ---- fiber decode: nodeutil.c:437-470 ---------------------------------- */
int orc_fiber_to_ind(size_t d, size_t N, const double *x, const size_t *ngrid,
                     double *const *xgrid, size_t *fixed_ind, size_t *dim_vary)
{
    for (size_t i = 0; i < d; i++) {
        fixed_ind[i] = orc_x_to_ind(x[i], ngrid[i], xgrid[i]);
        if (fixed_ind[i] == ngrid[i]) return 1;           /* off the grid */
    }
    *dim_vary = d;
    for (size_t i = 0; i < d; i++) {
        if (orc_x_to_ind(x[d + i], ngrid[i], xgrid[i]) != fixed_ind[i]) { *dim_vary = i; break; }
    }
    if (*dim_vary == d) return 1;
    return (N != ngrid[*dim_vary]) ? 2 : 0;
}

/* ---- obstacle test: boundary.c:329-344 (closed box), :668-680 ----------- */
int orc_in_obstacle(const orc_problem *p, const double *x)
{
    for (size_t o = 0; o < p->nobs; o++) {
        const double *lb = p->obs_lb + o * p->dx, *ub = p->obs_ub + o * p->dx;
        int inside = 1;
        for (size_t i = 0; i < p->dx; i++)
            if (x[i] < lb[i] || x[i] > ub[i]) { inside = 0; break; }
        if (inside) return 1;
    }
    return 0;
}

/* ---- flags + neighbour indices: nodeutil.c:489-627 ---------------------- */
int orc_fiber_neighbors(const orc_problem *p, const size_t *fi, size_t k,
                        const double *x, int *absorbed, size_t *nv, size_t *nf)
{
    const size_t d = p->dx, N = p->ngrid[k];
    for (size_t j = 0; j < N; j++)                                   /* :495-509 */
        absorbed[j] = orc_in_obstacle(p, x + j * d) ? -1 : 0;

    size_t slot = 0;
    for (size_t i = 0; i < d; i++) {                                 /* :512-566 */
        if (i == k) continue;
        const size_t i0 = fi[i], last = p->ngrid[i] - 1;
        size_t lo, hi;
        int wall = 0;
        if (i0 == 0) {                     /* tested first, as in the reference */
            if (p->bc[i] == ORC_ABSORB)        { lo = i0; hi = i0; wall = 1; }
            else if (p->bc[i] == ORC_REFLECT)  { lo = i0; hi = i0 + 1; }
            else if (p->bc[i] == ORC_PERIODIC) { lo = p->ngrid[i] - 2; hi = i0 + 1; }
            else return 3;
        } else if (i0 == last) {
            if (p->bc[i] == ORC_ABSORB)        { lo = i0; hi = i0; wall = 1; }
            else if (p->bc[i] == ORC_REFLECT)  { lo = i0 - 1; hi = i0; }
            else if (p->bc[i] == ORC_PERIODIC) { lo = i0 - 1; hi = 1; }
            else return 3;
        } else { lo = i0 - 1; hi = i0 + 1; }
        if (wall) for (size_t j = 0; j < N; j++) absorbed[j] = 1;
        nf[slot] = lo; nf[slot + 1] = hi;
        slot += 2;
    }

    /* the two ends of the varying dimension OVERWRITE the flags (:570-612) */
    const int bk = p->bc[k];
    if (bk == ORC_ABSORB)        { nv[0] = 0;     nv[1] = 0; absorbed[0] = 1; }
    else if (bk == ORC_REFLECT)  { nv[0] = 0;     nv[1] = 1; absorbed[0] = 0; }
    else if (bk == ORC_PERIODIC) { nv[0] = N - 2; nv[1] = 1; absorbed[0] = 0; }
    else return 3;
    const size_t e = N - 1;
    if (bk == ORC_ABSORB)        { nv[2 * e] = e;     nv[2 * e + 1] = e; absorbed[e] = 1; }
    else if (bk == ORC_REFLECT)  { nv[2 * e] = N - 2; nv[2 * e + 1] = e; absorbed[e] = 0; }
    else                         { nv[2 * e] = N - 2; nv[2 * e + 1] = 1; absorbed[e] = 0; }

    for (size_t j = 1; j + 1 < N; j++) {                             /* :615-624 */
        if (absorbed[j] == 0) { nv[2 * j] = j - 1; nv[2 * j + 1] = j + 1; }
        else                  { nv[2 * j] = j;     nv[2 * j + 1] = j; }
    }
    return 0;
}

/* ---- sequential BLAS-2/1 pieces (order of the reference's calls) -------- */
static void gemv_n(size_t m, size_t n, const double *A, const double *x, double *y)
{   /* y = A x, A column-major m x n */
    for (size_t a = 0; a < m; a++) {
        double s = 0.0;
        for (size_t b = 0; b < n; b++) s += A[a + b * m] * x[b];
        y[a] = s;
    }
}
static void gemv_t(size_t m, size_t n, const double *A, const double *x, double *y)
{   /* y = A^T x */
    for (size_t b = 0; b < n; b++) {
        double s = 0.0;
        for (size_t a = 0; a < m; a++) s += A[a + b * m] * x[a];
        y[b] = s;
    }
}
static double dot(size_t n, const double *x, const double *y)
{
    double s = 0.0;
    for (size_t i = 0; i < n; i++) s += x[i] * y[i];
    return s;
}

/* ---- FT values at a fiber's nodes and axis neighbours: valuefunc.c:369-585 */
int orc_ft_fiber_nn(const orc_ft *ft, const size_t *fi, size_t k,
                    const size_t *nf, const size_t *nv, double *out)
{
    const size_t d = ft->d, *r = ft->ranks, N = ft->n[k];
    size_t rmax = 1, nmax = 1;
    for (size_t i = 0; i <= d; i++) if (r[i] > rmax) rmax = r[i];
    for (size_t i = 0; i < d; i++) if (ft->n[i] > nmax) nmax = ft->n[i];
    if (d > ORC_MAXD) return 1;

    double *pool = calloc((2 * d * nmax + 1) * rmax, sizeof(double));
    if (!pool) return 1;
    double *fw[ORC_MAXD], *bw[ORC_MAXD], *tmp = pool + 2 * d * nmax * rmax;
    for (size_t i = 0; i < d; i++) {
        fw[i] = pool + i * nmax * rmax;
        bw[i] = pool + (d + i) * nmax * rmax;
    }
#define BLK(i, j) (ft->cores[i] + (j) * r[i] * r[(i) + 1])

    for (size_t i = 0; i < k; i++) {                     /* left chain :414-428 */
        if (i == 0) memcpy(fw[0], BLK(0, fi[0]), r[0] * r[1] * sizeof(double));
        else gemv_t(r[i], r[i + 1], BLK(i, fi[i]), fw[i - 1], fw[i]);
    }
    for (size_t i = d - 1; i > k; i--) {                 /* right chain :432-446 */
        if (i == d - 1) memcpy(bw[i], BLK(i, fi[i]), r[i] * r[i + 1] * sizeof(double));
        else gemv_n(r[i], r[i + 1], BLK(i, fi[i]), bw[i + 1], bw[i]);
    }
    {                                                    /* varying core :450-480 */
        const size_t m = r[k], n = r[k + 1];
        for (size_t j = 0; j < N; j++) {
            if (k == 0) memcpy(fw[k] + j * n, BLK(k, j), m * n * sizeof(double));
            else gemv_t(m, n, BLK(k, j), fw[k - 1], fw[k] + j * n);
            if (k == d - 1) memcpy(bw[k] + j * m, BLK(k, j), m * n * sizeof(double));
            else gemv_n(m, n, BLK(k, j), bw[k + 1], bw[k] + j * m);
        }
    }
    for (size_t i = k + 1; i < d; i++)                   /* forward :485-495 */
        for (size_t j = 0; j < N; j++)
            gemv_t(r[i], r[i + 1], BLK(i, fi[i]), fw[i - 1] + j * r[i], fw[i] + j * r[i + 1]);
    for (size_t i = k; i-- > 0;)                         /* backward :498-509 */
        for (size_t j = 0; j < N; j++)
            gemv_n(r[i], r[i + 1], BLK(i, fi[i]), bw[i + 1] + j * r[i + 1], bw[i] + j * r[i]);

    const size_t S = 2 * d + 1;
    for (size_t j = 0; j < N; j++) {                     /* along the fiber :514-519 */
        out[j * S + 2 * k]     = bw[0][nv[2 * j]];
        out[j * S + 2 * k + 1] = bw[0][nv[2 * j + 1]];
        out[j * S + 2 * d]     = bw[0][j];
    }
    for (size_t i = 0; i < k; i++) {                     /* fixed dims before :522-547 */
        const size_t m = r[i], n = r[i + 1];
        for (size_t j = 0; j < N; j++)
            for (size_t s = 0; s < 2; s++) {
                const size_t nb = nf[2 * i + s];
                if (i == 0) out[j * S + s] = dot(n, ft->cores[0] + nb * n, bw[1] + j * n);
                else {
                    gemv_n(m, n, BLK(i, nb), bw[i + 1] + j * n, tmp);
                    out[j * S + 2 * i + s] = dot(m, tmp, fw[i - 1]);
                }
            }
    }
    for (size_t i = k + 1; i < d; i++) {                 /* fixed dims after :549-582 */
        const size_t m = r[i], n = r[i + 1];
        for (size_t j = 0; j < N; j++)
            for (size_t s = 0; s < 2; s++) {
                const size_t nb = nf[2 * (i - 1) + s];
                if (i == d - 1) out[j * S + 2 * i + s] = dot(m, ft->cores[i] + nb * m, fw[i - 1] + j * m);
                else {
                    gemv_t(m, n, BLK(i, nb), fw[i - 1] + j * m, tmp);
                    out[j * S + 2 * i + s] = dot(n, tmp, bw[i + 1]);
                }
            }
    }
#undef BLK
    free(pool);
    return 0;
}

/* ---- nodeutil.c:647-713 ------------------------------------------------ */
int orc_neighbor_costs(const orc_problem *p, const orc_ft *ft, size_t N, const double *x,
                       size_t *fixed_ind, size_t *dim_vary, int *absorbed, double *costs)
{
    const size_t d = p->dx;
    for (size_t j = 0; j < N; j++) absorbed[j] = 0;
    memset(costs, 0, N * (2 * d + 1) * sizeof(double));
    int rc = orc_fiber_to_ind(d, N, x, p->ngrid, p->xgrid, fixed_ind, dim_vary);
    if (rc) return rc;
    size_t *nv = calloc(2 * N + 2 * d, sizeof(size_t)), *nf = nv + 2 * N;
    rc = orc_fiber_neighbors(p, fixed_ind, *dim_vary, x, absorbed, nv, nf);
    if (!rc) rc = orc_ft_fiber_nn(ft, fixed_ind, *dim_vary, nf, nv, costs);
    free(nv);
    return rc;
}

/* ---- upwind transition probabilities: nodeutil.c:284-309,:365-371,:396-402 */
int orc_transition(size_t dx, size_t dw, double h2, const double *t,
                   const double *drift, const double *ddiff, double *prob, double *dt)
{
    (void)dw;
    double norm = 0.0;
    for (size_t i = 0; i < dx; i++) {
        double s2 = ddiff[i * dx + i] * ddiff[i * dx + i];
        double q = t[2 * i + 1] * s2 / 2.0;
        prob[2 * i] = q;
        prob[2 * i + 1] = q;
        if (drift[i] < -1e-14)     prob[2 * i]     -= t[2 * i] * drift[i];
        else if (drift[i] > 1e-14) prob[2 * i + 1] += t[2 * i] * drift[i];
        norm += prob[2 * i];
        norm += prob[2 * i + 1];
    }
    if (norm < 1e-14) return 1;
    *dt = h2 / norm;
    prob[2 * dx] = 1.0;
    for (size_t i = 0; i < dx; i++) {
        prob[2 * i] /= norm;
        prob[2 * i + 1] /= norm;
        prob[2 * dx] -= prob[2 * i];
        prob[2 * dx] -= prob[2 * i + 1];
    }
    return 0;
}

/* ---- bellman.c:88-112 -------------------------------------------------- */
double orc_rhs(size_t dx, double stage, double beta, const double *prob, double dt,
               const double *cost)
{
    double ebt = exp(-beta * dt);
    double ctg = dot(2 * dx + 1, prob, cost);
    return dt * stage + ebt * ctg;
}

/* ---- bellman.c:367-480, absorbed==0, grad_u==NULL branch ---------------- */
double orc_control_value(const orc_problem *p, const double *x, const double *u,
                         const double *cost, double *prob, double *dt, double *stage,
                         int *status)
{
    double drift[ORC_MAXD], diff[ORC_MAXD * ORC_MAXD + 64];
    int rc = p->drift(0.0, x, u, drift, NULL, p->drift_arg);
    rc |= p->diff(0.0, x, u, diff, NULL, p->diff_arg);
    rc |= p->stage(0.0, x, u, stage, NULL);
    int ta = orc_transition(p->dx, p->dw, p->h2, p->t, drift, diff, prob, dt);
    if (status) *status = rc ? -1 : ta;
    if (ta) return NAN;        /* the reference asserts here (bellman.c:452) */
    return orc_rhs(p->dx, *stage, p->beta, prob, *dt, cost);
}

/* ---- bellman.c:504-543 (+ brute-force c3opt_minimize, first strict min) -- */
int orc_node_backup(const orc_problem *p, int absorbed, const double *x, const double *cost,
                    double *val, int *ubest)
{
    if (absorbed == 1)  { if (ubest) *ubest = -1; return p->boundcost(0.0, x, val); }
    if (absorbed == -1) { if (ubest) *ubest = -1; return p->obscost(x, val); }
    double prob[2 * ORC_MAXD + 1], dt, g, best = 0.0;
    int ib = 0, st = 0;
    for (size_t c = 0; c < p->nu; c++) {
        double v = orc_control_value(p, x, p->utab + c * p->du, cost, prob, &dt, &g, &st);
        if (st) return 10 + st;
        if (c == 0 || v < best) { best = v; ib = (int)c; }
    }
    *val = best;
    if (ubest) *ubest = ib;
    return 0;
}

/* ---- bellman.c:1295-1423 without the memo table -------------------------- */
int orc_vi_fiber(const orc_problem *p, const orc_ft *ft, size_t N, const double *x,
                 double *out, int *ubest, int *absorbed_out, double *costs_out)
{
    const size_t d = p->dx, S = 2 * d + 1;
    size_t fi[ORC_MAXD], k;
    int *absorbed = absorbed_out ? absorbed_out : malloc(N * sizeof(int));
    double *costs = costs_out ? costs_out : malloc(N * S * sizeof(double));
    int rc = orc_neighbor_costs(p, ft, N, x, fi, &k, absorbed, costs);
    for (size_t j = 0; j < N && !rc; j++)
        rc = orc_node_backup(p, absorbed[j], x + j * d, costs + j * S, out + j, ubest ? ubest + j : NULL);
    if (!absorbed_out) free(absorbed);
    if (!costs_out) free(costs);
    return rc;
}

/* ---- bellman.c:1702-1886 without the memo tables -------------------------- */
int orc_pi_fiber(const orc_problem *p, const orc_ft *ft_policy, const orc_ft *ft_iter,
                 size_t N, const double *x, int have_rows, double *rows, int *ubest,
                 double *out)
{
    const size_t d = p->dx, S = 2 * d + 1, R = 2 * d + 3;
    size_t fi[ORC_MAXD], k;
    int *abs_pol = malloc(2 * N * sizeof(int)), *abs_it = abs_pol + N;
    double *c_pol = malloc(2 * N * S * sizeof(double)), *c_it = c_pol + N * S;
    int rc = orc_neighbor_costs(p, ft_policy, N, x, fi, &k, abs_pol, c_pol);      /* :1742 */
    if (!rc) rc = orc_neighbor_costs(p, ft_iter, N, x, fi, &k, abs_it, c_it);      /* :1768 */
    for (size_t j = 0; j < N && !rc; j++) {
        const double *xj = x + j * d;
        double *row = rows + j * R;
        if (abs_pol[j] == 1)      { rc = p->boundcost(0.0, xj, out + j); if (ubest && !have_rows) ubest[j] = -1; continue; }
        if (abs_it[j] == -1)      { rc = p->obscost(xj, out + j);        if (ubest && !have_rows) ubest[j] = -1; continue; }
        if (!have_rows) {                                                 /* :1831-1860 */
            double v; int ib;
            rc = orc_node_backup(p, 0, xj, c_pol + j * S, &v, &ib);
            if (rc) break;
            int st;
            orc_control_value(p, xj, p->utab + (size_t)ib * p->du, c_pol + j * S,
                              row, row + S, row + S + 1, &st);
            if (st) { rc = 10 + st; break; }
            if (ubest) ubest[j] = ib;
        }
        out[j] = orc_rhs(d, row[S + 1], p->beta, row, row[S], c_it + j * S);   /* :1863-1871 */
    }
    free(abs_pol);
    free(c_pol);
    return rc;
}

/* ---- index-based batch conveniences -------------------------------------- */
void orc_fiber_points(const orc_problem *p, size_t k, const int *fixed_ind, double *x)
{
    const size_t d = p->dx, N = p->ngrid[k];
    for (size_t j = 0; j < N; j++)
        for (size_t i = 0; i < d; i++)
            x[j * d + i] = (i == k) ? p->xgrid[i][j] : p->xgrid[i][fixed_ind[i]];
}

int orc_vi_batch(const orc_problem *p, const orc_ft *ft, size_t F, const int *dim_vary,
                 const int *fixed_ind, size_t ldo, double *out, int *ubest, int nthreads)
{
    int err = 0;
    size_t nmax = 0;
    for (size_t i = 0; i < p->dx; i++) if (p->ngrid[i] > nmax) nmax = p->ngrid[i];
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        double *x = malloc(nmax * p->dx * sizeof(double));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (long f = 0; f < (long)F; f++) {
            size_t k = (size_t)dim_vary[f];
            orc_fiber_points(p, k, fixed_ind + f * p->dx, x);
            int rc = orc_vi_fiber(p, ft, p->ngrid[k], x, out + f * ldo,
                                  ubest ? ubest + f * ldo : NULL, NULL, NULL);
            if (rc) err = rc;
        }
        free(x);
    }
    (void)nthreads;
    return err;
}

int orc_pi_batch(const orc_problem *p, const orc_ft *ft_policy, const orc_ft *ft_iter,
                 size_t F, const int *dim_vary, const int *fixed_ind, size_t ldo,
                 int have_rows, double *rows, int *ubest, double *out, int nthreads)
{
    int err = 0;
    size_t nmax = 0;
    const size_t R = 2 * p->dx + 3;
    for (size_t i = 0; i < p->dx; i++) if (p->ngrid[i] > nmax) nmax = p->ngrid[i];
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        double *x = malloc(nmax * p->dx * sizeof(double));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (long f = 0; f < (long)F; f++) {
            size_t k = (size_t)dim_vary[f];
            orc_fiber_points(p, k, fixed_ind + f * p->dx, x);
            int rc = orc_pi_fiber(p, ft_policy, ft_iter, p->ngrid[k], x, have_rows,
                                  rows + f * ldo * R, ubest ? ubest + f * ldo : NULL,
                                  out + f * ldo);
            if (rc) err = rc;
        }
        free(x);
    }
    (void)nthreads;
    return err;
}

/* ---- multilinear evaluation of nodal cores (C3 LINELM function_train_eval) */
double orc_ft_eval_linear(const orc_ft *ft, double *const *xgrid, const double *x)
{
    const size_t d = ft->d, *r = ft->ranks;
    size_t rmax = 1;
    for (size_t i = 0; i <= d; i++) if (r[i] > rmax) rmax = r[i];
    double *v = calloc(3 * rmax * rmax, sizeof(double)), *w = v + rmax, *blk = v + 2 * rmax;
    v[0] = 1.0;
    for (size_t k = 0; k < d; k++) {
        const size_t n = ft->n[k], m = r[k], c = r[k + 1];
        const double *g = xgrid[k];
        size_t j = 0;
        while (j + 2 < n && x[k] > g[j + 1]) j++;
        double a = (x[k] - g[j]) / (g[j + 1] - g[j]);
        if (x[k] < g[0] || x[k] > g[n - 1]) { free(v); return 0.0; }   /* outside support */
        const double *b0 = ft->cores[k] + j * m * c, *b1 = b0 + m * c;
        for (size_t e = 0; e < m * c; e++) blk[e] = (1.0 - a) * b0[e] + a * b1[e];
        gemv_t(m, c, blk, v, w);
        memcpy(v, w, c * sizeof(double));
    }
    double res = v[0];
    free(v);
    return res;
}


/* ---- implicit policy at an arbitrary state (online controller) -------------------------------
 * mca_get_neighbor_node_costs, src/nodeutil.c:718-816: the value function at x -+ h e_i with the
 * boundary type deciding what stands in when the step leaves the grid; inside an obstacle every
 * entry is V(x) and absorbed = -1.  The reference leaves out[2d] (the node itself) unset outside
 * obstacles; it only meets the round-off probability p_self, and is set to V(x) here.        */
int orc_neighbor_node_costs(const orc_problem *p, const orc_ft *ft, const double *x, int *absorbed, double *out)
{
    const size_t d = p->dx;
    if (orc_in_obstacle(p, x)) {
        *absorbed = -1;
        const double val = orc_ft_eval_linear(ft, p->xgrid, x);
        for (size_t i = 0; i < 2 * d + 1; i++) out[i] = val;
        return 0;
    }
    *absorbed = 0;
    double xt[64];
    for (size_t i = 0; i < d; i++) xt[i] = x[i];
    for (size_t i = 0; i < d; i++) {
        const double lb = p->xgrid[i][0], ub = p->xgrid[i][p->ngrid[i] - 1];
        const double h = p->xgrid[i][1] - p->xgrid[i][0];
        if (((x[i] + h) < ub) && (x[i] - h > lb)) {                 /* standard case, :741-747 */
            xt[i] = x[i] - h; out[2 * i] = orc_ft_eval_linear(ft, p->xgrid, xt);
            xt[i] = x[i] + h; out[2 * i + 1] = orc_ft_eval_linear(ft, p->xgrid, xt);
        } else if ((x[i] - h) <= lb) {                              /* left boundary, :748-775 */
            xt[i] = x[i] + h; out[2 * i + 1] = orc_ft_eval_linear(ft, p->xgrid, xt);
            if (p->bc[i] == ORC_ABSORB || p->bc[i] == ORC_REFLECT) xt[i] = lb;
            else if (x[i] > lb) xt[i] = ub - (h - (x[i] - lb));
            else xt[i] = (ub - (lb - x[i])) - h;
            out[2 * i] = orc_ft_eval_linear(ft, p->xgrid, xt);
        } else {                                                    /* right boundary, :776-806 */
            xt[i] = x[i] - h; out[2 * i] = orc_ft_eval_linear(ft, p->xgrid, xt);
            if (p->bc[i] == ORC_ABSORB || p->bc[i] == ORC_REFLECT) xt[i] = ub;
            else if (x[i] < ub) xt[i] = lb + (h - (ub - x[i]));
            else xt[i] = (lb + (x[i] - ub)) + h;
            out[2 * i + 1] = orc_ft_eval_linear(ft, p->xgrid, xt);
        }
        xt[i] = x[i];
    }
    out[2 * d] = orc_ft_eval_linear(ft, p->xgrid, x);
    return 0;
}

/* c3control_policy_eval, src/bellman.c:2105-2151: neighbour values at x, then bellman_optimal
 * (brute force over the control table; absorbed = -1 gives the obstacle cost and u = 0).   */
int orc_policy_eval(const orc_problem *p, const orc_ft *ft, const double *x, double *u, double *val, int *absorbed,
                    double *costs)
{
    double cbuf[2 * 64 + 1];
    double *c = costs ? costs : cbuf;
    int ab = 0, ub = -1;
    int rc = orc_neighbor_node_costs(p, ft, x, &ab, c);
    if (rc) return rc;
    rc = orc_node_backup(p, ab, x, c, val, &ub);
    for (size_t i = 0; i < p->du; i++) u[i] = ub >= 0 ? p->utab[(size_t)ub * p->du + i] : 0.0;
    if (absorbed) *absorbed = ab;
    return rc;
}
