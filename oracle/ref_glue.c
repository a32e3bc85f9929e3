/* ref_glue.c -- links the reference's OWN object code into a testable library.
 * TEST INFRASTRUCTURE ONLY.
 *
 * oracle/Makefile compiles /root/reference/src/{bellman,nodeutil,boundary,
 * dynamics,hashgrid,util}.c unmodified, where they lie, with the reference's
 * flags, and links them with
 *   - c3shim/      : stand-ins for the absent C3/cdyn dependencies,
 *   - this file    : (1) the `struct ValueF` entry points those objects call
 *                    (valuefunc.c cannot be compiled without C3's
 *                    FunctionTrain internals, so its one hot function,
 *                    valuef_eval_fiber_ind_nn, is served by the restatement
 *                    orc_ft_fiber_nn), and (2) a small driver (`ref_*`) that
 *                    builds the reference structs through the reference's
 *                    public constructors and calls bellman_vi / bellman_pi
 *                    one fiber at a time -- the reference's granularity.
 * Output: oracle/_ref/libc3sc_ref.so (git-ignored, travels with gpurun).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "c3/array.h"
#include "c3/lib_optimization.h"
#include "util.h"
#include "boundary.h"
#include "dynamics.h"
#include "valuefunc.h"
#include "nodeutil.h"
#include "bellman.h"

#include "c3sc_oracle.h"
#include "models.h"

/* ------------------------------------------------------------------------
 * struct ValueF stand-in: nodal cores supplied directly.
 * ---------------------------------------------------------------------- */
struct ValueF {
    size_t d;
    size_t *N;
    size_t *ranks;
    double **cores;
};

struct ValueF *ref_valuef_create(size_t d, const size_t *n, const size_t *ranks, const double *flat)
{
    struct ValueF *vf = malloc(sizeof *vf);
    vf->d = d;
    vf->N = calloc_size_t(d);
    vf->ranks = calloc_size_t(d + 1);
    vf->cores = malloc_dd(d);
    memcpy(vf->N, n, d * sizeof(size_t));
    memcpy(vf->ranks, ranks, (d + 1) * sizeof(size_t));
    for (size_t k = 0; k < d; k++) {
        size_t len = n[k] * ranks[k] * ranks[k + 1];
        vf->cores[k] = calloc_double(len);
        memcpy(vf->cores[k], flat, len * sizeof(double));
        flat += len;
    }
    return vf;
}
void valuef_destroy(struct ValueF *vf)
{
    if (!vf) return;
    free(vf->N); free(vf->ranks); free_dd(vf->d, vf->cores); free(vf);
}
struct ValueF *valuef_copy(struct ValueF *vf)
{
    struct ValueF *c = malloc(sizeof *c);
    c->d = vf->d;
    c->N = calloc_size_t(vf->d);
    c->ranks = calloc_size_t(vf->d + 1);
    c->cores = malloc_dd(vf->d);
    memcpy(c->N, vf->N, vf->d * sizeof(size_t));
    memcpy(c->ranks, vf->ranks, (vf->d + 1) * sizeof(size_t));
    for (size_t k = 0; k < vf->d; k++) {
        size_t len = vf->N[k] * vf->ranks[k] * vf->ranks[k + 1];
        c->cores[k] = calloc_double(len);
        memcpy(c->cores[k], vf->cores[k], len * sizeof(double));
    }
    return c;
}
size_t *valuef_get_ranks(struct ValueF *vf) { return vf->ranks; }

static orc_ft as_ft(const struct ValueF *vf)
{
    orc_ft ft = { vf->d, vf->N, vf->ranks, vf->cores };
    return ft;
}
int valuef_eval_fiber_ind_nn(struct ValueF *vf, const size_t *fixed_ind, size_t dim_vary,
                             const size_t *neighbors, const size_t *neighbors_vary, double *out)
{
    orc_ft ft = as_ft(vf);
    return orc_ft_fiber_nn(&ft, fixed_ind, dim_vary, neighbors, neighbors_vary, out);
}

/* C3-only entry points that the reference objects reference but the hot path
 * never reaches.  Fail loudly if anything does.                           */
static void no_c3(const char *what)
{
    fprintf(stderr, "oracle/_ref: %s needs the C3 library, which is absent\n", what);
    abort();
}
static double *const *g_eval_grid = NULL;   /* set by ref_create for valuef_eval */
double valuef_eval(struct ValueF *vf, const double *x)
{
    if (!g_eval_grid) no_c3("valuef_eval");
    orc_ft ft = as_ft(vf);
    return orc_ft_eval_linear(&ft, g_eval_grid, x);
}
double valuef_norm(struct ValueF *vf) { (void)vf; no_c3("valuef_norm"); return 0; }
double valuef_norm2diff(struct ValueF *a, struct ValueF *b) { (void)a; (void)b; no_c3("valuef_norm2diff"); return 0; }
struct ValueF *valuef_interp(size_t d, int (*f)(size_t, const double *, double *, void *), void *args,
                             const size_t *N, double **grid, struct ValueF *vref,
                             struct ApproxArgs *aargs, int verbose)
{
    (void)d; (void)f; (void)args; (void)N; (void)grid; (void)vref; (void)aargs; (void)verbose;
    no_c3("valuef_interp (ftapprox_cross)");
    return NULL;
}

/* ------------------------------------------------------------------------
 * Driver: reference structs through the reference's public constructors.
 * ---------------------------------------------------------------------- */
typedef struct ref_ctx {
    size_t dx, du, dw, nmax;
    size_t *ngrid;
    struct C3Control *c3c;       /* owns xgrid (c3control_create, bellman.c:1962) */
    double **xgrid;
    double *h, hmin;
    struct Boundary *bound;
    struct MCAparam *mca;
    struct DPparam *dp;
    struct Workspace *work;
    struct c3Opt *opt;
    struct ControlParams *cp;
    struct PIparam *poli;
    double *xbuf;
} ref_ctx;

ref_ctx *ref_create(int model, size_t dx, const double *params, size_t nparams,
                    const double *lb, const double *ub, const size_t *ngrid, double beta,
                    const int *bc, size_t nobs, const double *obs_center, const double *obs_width,
                    size_t nu, const double *utab)
{
    size_t du, dw;
    if (orc_model_select(model, dx, params, nparams)) return NULL;
    orc_model_dims(model, dx, &du, &dw);
    ref_ctx *c = calloc(1, sizeof *c);
    c->dx = dx; c->du = du; c->dw = dw;
    c->ngrid = calloc_size_t(dx);
    memcpy(c->ngrid, ngrid, dx * sizeof(size_t));
    /* grid exactly as the reference builds it */
    c->c3c = c3control_create(dx, du, dw, (double *)lb, (double *)ub, c->ngrid, beta);
    c->xgrid = c3control_get_xgrid(c->c3c);
    g_eval_grid = c->xgrid;
    /* h, hmin as bellman.c:1975-1986 (not exposed by the facade) */
    c->h = calloc_double(dx);
    c->hmin = ub[0] - lb[0];
    c->nmax = ngrid[0];
    for (size_t i = 0; i < dx; i++) {
        c->h[i] = c->xgrid[i][1] - c->xgrid[i][0];
        if (c->h[i] < c->hmin) c->hmin = c->h[i];
        if (ngrid[i] > c->nmax) c->nmax = ngrid[i];
    }
    c->bound = boundary_alloc(dx, (double *)lb, (double *)ub);
    for (size_t i = 0; i < dx; i++) {
        if (bc[i] == PERIODIC) boundary_external_set_type(c->bound, i, "periodic");
        else if (bc[i] == REFLECT) boundary_external_set_type(c->bound, i, "reflect");
    }
    for (size_t o = 0; o < nobs; o++)
        boundary_add_obstacle(c->bound, (double *)obs_center + o * dx, (double *)obs_width + o * dx);
    c->mca = mca_param_create(dx, du);
    mca_add_grid_refs(c->mca, c->ngrid, c->xgrid, c->hmin, c->h);
    c->dp = dp_param_create(dx, du, dw, beta);
    dp_param_add_boundary(c->dp, c->bound);
    dp_param_add_drift(c->dp, orc_model_drift(), NULL);
    dp_param_add_diff(c->dp, orc_model_diff(), orc_model_diff_arg());
    dp_param_add_stagecost(c->dp, orc_model_stage());
    dp_param_add_boundcost(c->dp, orc_model_boundcost());
    dp_param_add_obscost(c->dp, orc_model_obscost());
    c->work = workspace_alloc(dx, du, dw, c->nmax);
    c->opt = c3opt_alloc(BRUTEFORCE, du);
    c3opt_set_brute_force_vals(c->opt, nu, (double *)utab);
    c->cp = control_params_create(dx, dw, c->dp, c->mca, c->work, c->opt);
    c->xbuf = calloc_double(c->nmax * dx);
    return c;
}

void ref_destroy(ref_ctx *c)
{
    if (!c) return;
    if (c->poli) pi_param_destroy(c->poli);
    control_params_destroy(c->cp);
    c3opt_free(c->opt);
    workspace_free(c->work);
    dp_param_destroy(c->dp);
    mca_param_destroy(c->mca);
    boundary_free(c->bound);
    c3control_destroy(c->c3c);
    free(c->h); free(c->ngrid); free(c->xbuf); free(c);
    g_eval_grid = NULL;
}

void ref_get_grid(const ref_ctx *c, size_t dim, double *out) { memcpy(out, c->xgrid[dim], c->ngrid[dim] * sizeof(double)); }
double ref_get_hmin(const ref_ctx *c) { return c->hmin; }
void ref_get_h(const ref_ctx *c, double *h) { memcpy(h, c->h, c->dx * sizeof(double)); }
void ref_get_obstacle(const ref_ctx *c, size_t o, double *lb, double *ub)
{
    memcpy(lb, boundary_obstacle_get_lb(c->bound, o), c->dx * sizeof(double));
    memcpy(ub, boundary_obstacle_get_ub(c->bound, o), c->dx * sizeof(double));
}

static void fiber_x(const ref_ctx *c, size_t k, const int *fi, double *x)
{
    for (size_t j = 0; j < c->ngrid[k]; j++)
        for (size_t i = 0; i < c->dx; i++)
            x[j * c->dx + i] = (i == k) ? c->xgrid[i][j] : c->xgrid[i][fi[i]];
}

/* process_fibers_neighbor (nodeutil.c:489) on an index-described fiber */
int ref_fiber_neighbors(ref_ctx *c, int dim_vary, const int *fixed_ind,
                        int *absorbed, size_t *nbr_vary, size_t *nbr_fixed)
{
    size_t fi[64];
    for (size_t i = 0; i < c->dx; i++) fi[i] = (size_t)fixed_ind[i];
    fiber_x(c, (size_t)dim_vary, fixed_ind, c->xbuf);
    return process_fibers_neighbor(c->dx, fi, (size_t)dim_vary, c->xbuf, absorbed,
                                   nbr_vary, nbr_fixed, c->ngrid, c->bound);
}

/* convert_fiber_to_ind (nodeutil.c:437) on point data */
int ref_fiber_to_ind(ref_ctx *c, size_t N, const double *x, size_t *fixed_ind, size_t *dim_vary)
{
    return convert_fiber_to_ind(c->dx, N, x, c->ngrid, c->xgrid, fixed_ind, dim_vary);
}

/* mca_get_neighbor_costs (nodeutil.c:647) */
int ref_neighbor_costs(ref_ctx *c, struct ValueF *vf, int dim_vary, const int *fixed_ind,
                       int *absorbed, double *costs)
{
    size_t fi[64], k;
    fiber_x(c, (size_t)dim_vary, fixed_ind, c->xbuf);
    return mca_get_neighbor_costs(c->dx, c->ngrid[dim_vary], c->xbuf, c->bound, vf, c->ngrid,
                                  c->xgrid, fi, &k, absorbed, costs);
}

/* mca_get_neighbor_node_costs (nodeutil.c:718) at one off-grid state */
int ref_neighbor_node_costs(ref_ctx *c, struct ValueF *vf, const double *x, int *absorbed, double *costs)
{
    return mca_get_neighbor_node_costs(c->dx, x, c->bound, vf, c->ngrid, c->xgrid, absorbed, costs);
}

/* transition_assemble (nodeutil.c:267), non-gradient branch, with the grid's h2/t */
int ref_transition(ref_ctx *c, const double *drift, const double *ddiff, double *prob, double *dt)
{
    double h2 = c->hmin * c->hmin, t[128];
    for (size_t i = 0; i < c->dx; i++) { t[2 * i] = h2 / c->h[i]; t[2 * i + 1] = t[2 * i] / c->h[i]; }
    return transition_assemble(c->dx, c->du, c->dw, h2, t, drift, NULL, ddiff, NULL,
                               prob, NULL, dt, NULL, NULL);
}
void ref_get_h2_t(const ref_ctx *c, double *h2, double *t)
{
    *h2 = c->hmin * c->hmin;
    for (size_t i = 0; i < c->dx; i++) { t[2 * i] = *h2 / c->h[i]; t[2 * i + 1] = t[2 * i] / c->h[i]; }
}

double ref_rhs(ref_ctx *c, double stage, double beta, const double *prob, double dt, const double *cost)
{
    return bellmanrhs(c->dx, c->du, stage, NULL, beta, prob, NULL, dt, NULL, cost, NULL);
}

/* One value-iteration pass over F fibers: bellman_vi (bellman.c:1295) once per
 * fiber, fresh vi_iter + emptied memo so nothing is short-circuited by an
 * earlier pass.  Returns wall seconds spent inside the bellman_vi calls
 * (negative on failure).                                                   */
double ref_vi_fibers(ref_ctx *c, struct ValueF *vf, size_t F, const int *dim_vary,
                     const int *fixed_ind, size_t ldo, double *out, int fresh_per_fiber)
{
    struct VIparam *vi = vi_param_create(1e-10);
    vi_param_add_cp(vi, c->cp);
    vi_param_add_value(vi, vf);
    workspace_reset_vi_htable(c->work);
    workspace_increment_vi_iter(c->work);
    struct timespec t0, t1;
    double secs = 0.0;
    for (size_t f = 0; f < F; f++) {
        size_t k = (size_t)dim_vary[f];
        fiber_x(c, k, fixed_ind + f * c->dx, c->xbuf);
        /* The memo key is (multi-index, vi_iter) (bellman.c:1334-1344) while the flags of a
         * node depend on the fiber it is reached through (end-node overwrite,
         * nodeutil.c:570-612): with a shared vi_iter a node reached twice keeps whichever
         * value was computed first.  fresh_per_fiber gives every fiber its own vi_iter so
         * each call is the pure per-fiber operator.                                      */
        if (fresh_per_fiber && f) workspace_increment_vi_iter(c->work);
        clock_gettime(CLOCK_MONOTONIC, &t0);
        int rc = bellman_vi(c->ngrid[k], c->xbuf, out + f * ldo, vi);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        secs += (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
        if (rc) { secs = -1.0; break; }
    }
    vi_param_destroy(vi);
    return secs;
}

/* Policy evaluation, mirroring c3control_pi_solve / step_pi (bellman.c:2343,
 * :2214): ref_pi_begin fixes the policy (new pi_iter, tables emptied);
 * every ref_pi_fibers call is one sub-iteration against vf_iter.          */
int ref_pi_begin(ref_ctx *c, struct ValueF *vf_policy)
{
    if (c->poli) pi_param_destroy(c->poli);
    c->poli = pi_param_create(1e-10, vf_policy);
    workspace_increment_pi_iter(c->work);
    workspace_reset_pi_prob_htable(c->work);
    workspace_reset_pi_htable(c->work);
    return 0;
}
double ref_pi_fibers(ref_ctx *c, struct ValueF *vf_iter, size_t F, const int *dim_vary,
                     const int *fixed_ind, size_t ldo, double *out)
{
    if (!c->poli) return -1.0;
    pi_param_add_cp(c->poli, c->cp);
    pi_param_add_value(c->poli, vf_iter);
    workspace_increment_pi_subiter(c->work);
    struct timespec t0, t1;
    double secs = 0.0;
    for (size_t f = 0; f < F; f++) {
        size_t k = (size_t)dim_vary[f];
        fiber_x(c, k, fixed_ind + f * c->dx, c->xbuf);
        clock_gettime(CLOCK_MONOTONIC, &t0);
        int rc = bellman_pi(c->ngrid[k], c->xbuf, out + f * ldo, c->poli);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        secs += (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
        if (rc) { secs = -1.0; break; }
    }
    return secs;
}

void ref_set_omp_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int ref_omp_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
