/* c3sc_host.c -- C host mirror of the reference API (include/c3sc_host.h).
 *
 * Containers and index decoding live here; every floating-point result of the
 * path comes back from the GPU through include/c3sc_b200.h.  Written from the
 * behaviour documented in SURVEY.md / the reference headers, not transcribed.
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../../include/c3sc_host.h"
#include "../../../include/c3sc_cross.h"

static void *xalloc(size_t n, size_t sz)
{
    void *p = calloc(n ? n : 1, sz);
    if (!p) { fprintf(stderr, "c3sc_b200: out of host memory\n"); exit(1); }
    return p;
}
static double *dupd(const double *s, size_t n)
{
    double *p = xalloc(n, sizeof(double));
    if (s) memcpy(p, s, n * sizeof(double));
    return p;
}
static void die(const char *what)
{
    fprintf(stderr, "c3sc_b200: %s: %s\n", what, c3sc_last_error());
    abort();                      /* the reference asserts on every failure of the path */
}

/* ========================== boundary ======================================= */
struct Boundary {
    size_t d, nobs;
    double *lo, *hi;              /* domain bounds                       */
    int *type;                    /* enum EBTYPE per dimension           */
    double *obs_lb[C3SC_MAXOBS], *obs_ub[C3SC_MAXOBS];
};
static int parse_type(const char *s)
{
    if (!strcmp(s, "absorb")) return ABSORB;
    if (!strcmp(s, "periodic")) return PERIODIC;
    if (!strcmp(s, "reflect")) return REFLECT;
    fprintf(stderr, "External boundary of type %s is unknown\n", s);
    exit(1);
}
struct Boundary *boundary_alloc(size_t d, double *lb, double *ub)
{
    struct Boundary *b = xalloc(1, sizeof *b);
    b->d = d;
    b->lo = dupd(lb, d);
    b->hi = dupd(ub, d);
    b->type = xalloc(d, sizeof(int));
    for (size_t i = 0; i < d; i++) b->type[i] = ABSORB;     /* default, boundary.c:386-388 */
    return b;
}
struct Boundary *boundary_copy_deep(struct Boundary *o)
{
    if (!o) return NULL;
    struct Boundary *b = boundary_alloc(o->d, o->lo, o->hi);
    memcpy(b->type, o->type, o->d * sizeof(int));
    b->nobs = o->nobs;
    for (size_t k = 0; k < o->nobs; k++) { b->obs_lb[k] = dupd(o->obs_lb[k], o->d); b->obs_ub[k] = dupd(o->obs_ub[k], o->d); }
    return b;
}
void boundary_free(struct Boundary *b)
{
    if (!b) return;
    for (size_t k = 0; k < b->nobs; k++) { free(b->obs_lb[k]); free(b->obs_ub[k]); }
    free(b->lo); free(b->hi); free(b->type); free(b);
}
size_t boundary_get_nobs(struct Boundary *b) { return b->nobs; }
double *boundary_obstacle_get_lb(struct Boundary *b, size_t k) { return b->obs_lb[k]; }
double *boundary_obstacle_get_ub(struct Boundary *b, size_t k) { return b->obs_ub[k]; }
void boundary_add_obstacle(struct Boundary *b, double *center, double *lengths)
{
    if (b->nobs == C3SC_MAXOBS) { fprintf(stderr, "Not enough space allocated for obstacles in boundary\n"); exit(1); }
    double *lb = xalloc(b->d, sizeof(double)), *ub = xalloc(b->d, sizeof(double));
    for (size_t i = 0; i < b->d; i++) {            /* boundary.c:264-267 */
        lb[i] = center[i] - lengths[i] / 2.0;
        ub[i] = center[i] + lengths[i] / 2.0;
    }
    b->obs_lb[b->nobs] = lb; b->obs_ub[b->nobs] = ub;
    b->nobs++;
}
void boundary_external_set_type(struct Boundary *b, size_t dim, char *type) { b->type[dim] = parse_type(type); }
enum EBTYPE boundary_type_dim(const struct Boundary *b, size_t dim, int right) { (void)right; return (enum EBTYPE)b->type[dim]; }
int boundary_in_obstacle(const struct Boundary *b, const double *x)
{   /* geometric predicate (closed boxes, first hit), boundary.c:329-344,668-680 */
    for (size_t k = 0; k < b->nobs; k++) {
        int in = 1;
        for (size_t i = 0; i < b->d && in; i++) in = !(x[i] < b->obs_lb[k][i] || x[i] > b->obs_ub[k][i]);
        if (in) return 1;
    }
    return 0;
}

double outer_bound_dim(const struct Boundary *b, size_t dim, double x, int *map)
{   /* boundary.c:577-597: only a PERIODIC face maps; left is tested first */
    *map = 0;
    if (x <= b->lo[dim]) {
        if (b->type[dim] == PERIODIC) { *map = 1; return b->hi[dim]; }
    } else if (x >= b->hi[dim]) {
        if (b->type[dim] == PERIODIC) { *map = 2; return b->lo[dim]; }
    }
    return x;
}

/* ========================== dynamics containers ============================ */
struct Drift { size_t dx, du; c3sc_dyn_cb b; void *bargs; };
struct Diff  { size_t dx, du, dw; c3sc_dyn_cb s; void *sargs; };
struct Dyn   { struct Drift *drift; struct Diff *diff; };
struct Drift *drift_alloc(size_t dx, size_t du) { struct Drift *d = xalloc(1, sizeof *d); d->dx = dx; d->du = du; return d; }
struct Drift *drift_copy(struct Drift *o) { if (!o) return NULL; struct Drift *d = xalloc(1, sizeof *d); *d = *o; return d; }
void drift_free(struct Drift *d) { free(d); }
void drift_add_func(struct Drift *d, c3sc_dyn_cb b, void *a) { d->b = b; d->bargs = a; }
size_t drift_get_dx(struct Drift *d) { return d->dx; }
int drift_eval(struct Drift *d, double t, const double *x, const double *u, double *out, double *jac)
{
    if (!d->b) { fprintf(stderr, "Warning: drift dynamics (Drift->b) are not\n         yet specified\n"); return 1; }
    return d->b(t, x, u, out, jac, d->bargs);
}
struct Diff *diff_alloc(size_t dx, size_t du, size_t dw) { struct Diff *d = xalloc(1, sizeof *d); d->dx = dx; d->du = du; d->dw = dw; return d; }
struct Diff *diff_copy(struct Diff *o) { if (!o) return NULL; struct Diff *d = xalloc(1, sizeof *d); *d = *o; return d; }
void diff_free(struct Diff *d) { free(d); }
void diff_add_func(struct Diff *d, c3sc_dyn_cb s, void *a) { d->s = s; d->sargs = a; }
int diff_eval(struct Diff *d, double t, const double *x, const double *u, double *out, double *jac)
{
    if (!d->s) { fprintf(stderr, "Warning: Diff dynamics (Diff->s) are not\n         yet specified\n"); return 1; }
    return d->s(t, x, u, out, jac, d->sargs);
}
size_t diff_get_dw(struct Diff *d) { return d->dw; }
struct Dyn *dyn_alloc(struct Drift *a, struct Diff *b) { struct Dyn *d = xalloc(1, sizeof *d); d->drift = a; d->diff = b; return d; }
void dyn_free(struct Dyn *d) { free(d); }
void dyn_free_deep(struct Dyn *d) { if (d) { drift_free(d->drift); diff_free(d->diff); free(d); } }
struct Dyn *dyn_copy_deep(struct Dyn *o) { return o ? dyn_alloc(drift_copy(o->drift), diff_copy(o->diff)) : NULL; }
void dyn_init_ref(struct Dyn *d, struct Drift *a, struct Diff *b) { d->drift = a; d->diff = b; }
size_t dyn_get_dx(struct Dyn *d) { return d->drift->dx; }
size_t dyn_get_dw(struct Dyn *d) { return d->diff->dw; }
size_t dyn_get_du(struct Dyn *d) { return d->drift->du; }
int dyn_eval(struct Dyn *d, double t, const double *x, const double *u, double *drift, double *jd, double *diff, double *jf)
{
    int rc = drift_eval(d->drift, t, x, u, drift, jd);
    return rc ? rc : diff_eval(d->diff, t, x, u, diff, jf);
}

/* ========================== c3opt (brute force) ============================ */
struct c3Opt { enum c3opt_alg alg; size_t d, n; double *vals; };
struct c3Opt *c3opt_alloc(enum c3opt_alg alg, size_t d) { struct c3Opt *o = xalloc(1, sizeof *o); o->alg = alg; o->d = d; return o; }
struct c3Opt *c3opt_copy(struct c3Opt *s) { if (!s) return NULL; struct c3Opt *o = xalloc(1, sizeof *o); *o = *s; o->vals = s->vals ? dupd(s->vals, s->n * s->d) : NULL; return o; }
void c3opt_free(struct c3Opt *o) { if (o) { free(o->vals); free(o); } }
int c3opt_is_bruteforce(const struct c3Opt *o) { return o->alg == BRUTEFORCE; }
void c3opt_set_brute_force_vals(struct c3Opt *o, size_t n, double *vals) { free(o->vals); o->n = n; o->vals = dupd(vals, n * o->d); }
size_t c3opt_get_d(const struct c3Opt *o) { return o->d; }

/* ========================== value function ================================= */
struct ValueF {
    size_t d, *N, *ranks;
    double **cores;               /* host copy, valuef_precompute_cores layout */
    c3sc_valuef *dev;
    c3sc_cross *cross;            /* isl / isr of the reference (src/valuefunc.c:62-78): ranks + index sets of the
                                     cross run that produced this function; NULL for valuef_from_cores */
    double **xgrid;               /* owned copy of the nodes (the reference keeps them inside the LINELM cores) */
    c3sc_problem *eval_dev;       /* geometry-only device problem for valuef_eval, built on first use */
    struct CrossIndex **isl;      /* valuef_get_isl's view of cross's left index sets, rebuilt per call */
};
static void isl_free(struct ValueF *v);
struct ValueF *valuef_from_cores(size_t d, const size_t *N, const size_t *ranks, double *const *cores)
{
    struct ValueF *v = xalloc(1, sizeof *v);
    v->d = d;
    v->N = xalloc(d, sizeof(size_t));
    v->ranks = xalloc(d + 1, sizeof(size_t));
    v->cores = xalloc(d, sizeof(double *));
    memcpy(v->N, N, d * sizeof(size_t));
    memcpy(v->ranks, ranks, (d + 1) * sizeof(size_t));
    uint64_t n64[C3SC_MAXD], r64[C3SC_MAXD + 1];
    for (size_t k = 0; k < d; k++) { v->cores[k] = dupd(cores[k], N[k] * ranks[k] * ranks[k + 1]); n64[k] = N[k]; }
    for (size_t k = 0; k <= d; k++) r64[k] = ranks[k];
    if (c3sc_valuef_create((uint32_t)d, n64, r64, (const double *const *)v->cores, &v->dev)) die("valuef_from_cores");
    return v;
}
int valuef_update_cores(struct ValueF *v, double *const *cores)
{
    for (size_t k = 0; k < v->d; k++) memcpy(v->cores[k], cores[k], v->N[k] * v->ranks[k] * v->ranks[k + 1] * sizeof(double));
    return c3sc_valuef_update(v->dev, (const double *const *)v->cores);
}
void valuef_destroy(struct ValueF *v)
{
    if (!v) return;
    isl_free(v);
    c3sc_valuef_destroy(v->dev);
    c3sc_cross_destroy(v->cross);
    c3sc_problem_destroy(v->eval_dev);
    if (v->xgrid) { for (size_t k = 0; k < v->d; k++) free(v->xgrid[k]); free(v->xgrid); }
    for (size_t k = 0; k < v->d; k++) free(v->cores[k]);
    free(v->cores); free(v->N); free(v->ranks); free(v);
}
void valuef_set_grid(struct ValueF *v, double *const *xgrid)
{
    if (v->xgrid) { for (size_t k = 0; k < v->d; k++) free(v->xgrid[k]); free(v->xgrid); }
    v->xgrid = xalloc(v->d, sizeof(double *));
    for (size_t k = 0; k < v->d; k++) v->xgrid[k] = dupd(xgrid[k], v->N[k]);
    c3sc_problem_destroy(v->eval_dev); v->eval_dev = NULL;
}
struct ValueF *valuef_copy(struct ValueF *v)
{
    struct ValueF *c = valuef_from_cores(v->d, v->N, v->ranks, v->cores);
    if (v->cross && c3sc_cross_copy(v->cross, &c->cross)) die("valuef_copy");
    if (v->xgrid) valuef_set_grid(c, v->xgrid);
    return c;
}
size_t *valuef_get_ranks(struct ValueF *v) { return v->ranks; }

/* src/valuefunc.c:218: the left index sets the cross run that produced vf ended with.  isl[k], k = 0..d-1, holds the
   r_k multi-indices over dimensions 0..k-1 that core k's fibers were sampled at (isl[0] is the empty prefix).
   C3's struct CrossIndex is absent; the stand-in in c3sc_host.h carries grid indices and, when the value function
   knows its grid, the node coordinates C3 stores.  NULL for a value function that did not come from a cross run. */
static void isl_free(struct ValueF *v)
{
    if (!v->isl) return;
    for (size_t k = 0; k < v->d; k++) if (v->isl[k]) { free(v->isl[k]->inds); free(v->isl[k]->vals); free(v->isl[k]); }
    free(v->isl); v->isl = NULL;
}
struct CrossIndex **valuef_get_isl(const struct ValueF *cv)
{
    struct ValueF *v = (struct ValueF *)cv;
    if (!v || !v->cross) return NULL;
    isl_free(v);
    uint64_t r[C3SC_MAXD + 1];
    if (c3sc_cross_ranks(v->cross, r)) return NULL;
    v->isl = xalloc(v->d, sizeof *v->isl);
    for (size_t k = 0; k < v->d; k++) {
        struct CrossIndex *ci = xalloc(1, sizeof *ci);
        ci->d = k; ci->n = (size_t)r[k];
        ci->inds = xalloc(ci->n * (k ? k : 1), sizeof(size_t));
        ci->vals = v->xgrid ? xalloc(ci->n * (k ? k : 1), sizeof(double)) : NULL;
        int32_t *left = xalloc(ci->n * v->d, sizeof(int32_t));
        if (c3sc_cross_index_sets(v->cross, (uint32_t)k, left, NULL)) die("valuef_get_isl");
        for (size_t a = 0; a < ci->n; a++)
            for (size_t i = 0; i < k; i++) {
                ci->inds[a * k + i] = (size_t)left[a * v->d + i];
                if (ci->vals) ci->vals[a * k + i] = v->xgrid[i][left[a * v->d + i]];
            }
        free(left);
        v->isl[k] = ci;
    }
    return v->isl;
}

int valuef_eval_fiber_ind_nn(struct ValueF *vf, const size_t *fixed_ind, size_t dim_vary,
                             const size_t *neighbors, const size_t *neighbors_vary, double *out)
{
    const size_t d = vf->d, N = vf->N[dim_vary];
    size_t nmax = 0;
    for (size_t i = 0; i < d; i++) if (vf->N[i] > nmax) nmax = vf->N[i];
    int32_t fi[C3SC_MAXD], nf[2 * C3SC_MAXD], dv = (int32_t)dim_vary;
    for (size_t i = 0; i < d; i++) fi[i] = (int32_t)fixed_ind[i];
    for (size_t i = 0; i + 2 < 2 * d; i++) nf[i] = (int32_t)neighbors[i];
    int32_t *nv = xalloc(2 * nmax, sizeof(int32_t));
    for (size_t j = 0; j < 2 * N; j++) nv[j] = (int32_t)neighbors_vary[j];
    double *tmp = xalloc(nmax * (2 * d + 1), sizeof(double));
    int rc = c3sc_ft_fiber_nn_batch(vf->dev, 1, &dv, fi, nf, nv, nmax, tmp);
    if (!rc) memcpy(out, tmp, N * (2 * d + 1) * sizeof(double));
    free(nv); free(tmp);
    return rc;
}

/* src/util.c:995-1006 */
size_t uniform_stride(size_t N, size_t M)
{
    size_t stride = 1;
    if (M < 2) return 0;
    while (stride * (M - 1) < N - 1) stride++;
    return stride - 1;
}

/* ========================== workspace ====================================== */
struct RowEntry { uint64_t hash; size_t pi_iter; int32_t key[C3SC_MAXD + 1]; int32_t slot; struct RowEntry *next; };   /* slot: the fiber's rows in the device store */
struct Workspace {
    size_t dx, du, dw, N;
    size_t active;                /* src/util.c: workspace_set_active */
    size_t index[11];             /* offsets of the slab fields, src/util.c:738-748 */
    double **mem;                 /* N per-node slabs: drift, grad_drift, diff, grad_diff, dt, grad_dt, prob, grad_prob,
                                     grad_stage, extra, u -- filled by the scalar entries bellman_control / bellman_optimal
                                     (one node per call); the batched kernels keep these quantities in registers */
    size_t vi_iter, pi_iter, pi_subiter;
    double *costs;                /* N*(2dx+1): neighbour costs of the fiber in flight */
    int *absorbed;                /* N */
    double *u;                    /* N*du */
    struct RowEntry **rows;       /* fiber -> slot of its policy rows (replaces pi_prob_htable); the rows themselves
                                     stay on the device, in `store` */
    size_t nbuckets;
    c3sc_rowstore *store;         /* created on first use (needs the grid's nmax) */
    size_t nslots;                /* slots handed out for the current policy */
};
struct Workspace *workspace_alloc(size_t dx, size_t du, size_t dw, size_t N)
{
    struct Workspace *w = xalloc(1, sizeof *w);
    w->dx = dx; w->du = du; w->dw = dw; w->N = N;
    w->costs = xalloc(N * (2 * dx + 1), sizeof(double));
    w->absorbed = xalloc(N, sizeof(int));
    w->u = xalloc(N * (du ? du : 1), sizeof(double));
    /* slab layout of src/util.c:738-748 */
    w->index[0] = dx;                                   /* drift */
    w->index[1] = w->index[0] + dx * du;                /* grad_drift */
    w->index[2] = w->index[1] + dx * dw;                /* diff */
    w->index[3] = w->index[2] + dx * dw * du;           /* grad_diff */
    w->index[4] = w->index[3] + 1;                      /* dt */
    w->index[5] = w->index[4] + du;                     /* grad_dt */
    w->index[6] = w->index[5] + 2 * dx + 1;             /* prob */
    w->index[7] = w->index[6] + du * (2 * dx + 1);      /* grad_prob */
    w->index[8] = w->index[7] + du;                     /* grad_stage */
    w->index[9] = w->index[8] + du;                     /* extra */
    w->index[10] = w->index[9] + du;                    /* u */
    w->mem = xalloc(N ? N : 1, sizeof(double *));
    for (size_t i = 0; i < N; i++) w->mem[i] = xalloc(w->index[10] ? w->index[10] : 1, sizeof(double));
    w->nbuckets = 1 << 16;
    w->rows = xalloc(w->nbuckets, sizeof(struct RowEntry *));
    return w;
}
void workspace_reset_pi_prob_htable(struct Workspace *w)
{
    for (size_t b = 0; b < w->nbuckets; b++) {
        struct RowEntry *e = w->rows[b];
        while (e) { struct RowEntry *n = e->next; free(e); e = n; }
        w->rows[b] = NULL;
    }
    w->nslots = 0;                /* the store keeps its memory; its slots are handed out again */
}
void workspace_reset_pi_htable(struct Workspace *w) { (void)w; }   /* value memo: dropped, backups are pure */
void workspace_reset_vi_htable(struct Workspace *w) { (void)w; }
void workspace_free(struct Workspace *w)
{
    if (!w) return;
    workspace_reset_pi_prob_htable(w);
    c3sc_rowstore_destroy(w->store);
    for (size_t i = 0; i < w->N; i++) free(w->mem[i]);
    free(w->mem);
    free(w->rows); free(w->costs); free(w->absorbed); free(w->u); free(w);
}
void workspace_increment_vi_iter(struct Workspace *w) { w->vi_iter++; }
size_t workspace_get_vi_iter(const struct Workspace *w) { return w->vi_iter; }
void workspace_increment_pi_iter(struct Workspace *w) { w->pi_iter++; }
size_t workspace_get_pi_iter(const struct Workspace *w) { return w->pi_iter; }
void workspace_increment_pi_subiter(struct Workspace *w) { w->pi_subiter++; }
size_t workspace_get_pi_subiter(const struct Workspace *w) { return w->pi_subiter; }
double *workspace_get_costs(struct Workspace *w, size_t node) { return w->costs + node * (2 * w->dx + 1); }
int *workspace_get_absorbed(struct Workspace *w, size_t node) { return w->absorbed + node; }
double *workspace_get_u(struct Workspace *w, size_t node) { return w->mem[node] + w->index[9]; }
/* src/util.c:843-906: the per-node scratch slabs */
void workspace_set_active(struct Workspace *w, size_t active) { w->active = active; }
size_t workspace_get_active(struct Workspace *w) { return w->active; }
double *workspace_get_drift(struct Workspace *w, size_t node) { return w->mem[node]; }
double *workspace_get_grad_drift(struct Workspace *w, size_t node) { return w->mem[node] + w->index[0]; }
double *workspace_get_diff(struct Workspace *w, size_t node) { return w->mem[node] + w->index[1]; }
double *workspace_get_grad_diff(struct Workspace *w, size_t node) { return w->mem[node] + w->index[2]; }
double *workspace_get_dt(struct Workspace *w, size_t node) { return w->mem[node] + w->index[3]; }
double *workspace_get_grad_dt(struct Workspace *w, size_t node) { return w->mem[node] + w->index[4]; }
double *workspace_get_prob(struct Workspace *w, size_t node) { return w->mem[node] + w->index[5]; }
double *workspace_get_grad_prob(struct Workspace *w, size_t node) { return w->mem[node] + w->index[6]; }
double *workspace_get_grad_stage(struct Workspace *w, size_t node) { return w->mem[node] + w->index[7]; }
double *workspace_get_control_size_extra(struct Workspace *w, size_t node) { return w->mem[node] + w->index[8]; }

static uint64_t fiber_hash(size_t pi_iter, int32_t k, const int32_t *fi, size_t d)
{
    uint64_t h = 1469598103934665603ull ^ pi_iter;
    h = (h ^ (uint64_t)(uint32_t)k) * 1099511628211ull;
    for (size_t i = 0; i < d; i++) h = (h ^ (uint64_t)(uint32_t)((size_t)k == i ? 0 : fi[i])) * 1099511628211ull;
    return h;
}
static struct RowEntry *rows_find(struct Workspace *w, int32_t k, const int32_t *fi)
{
    uint64_t h = fiber_hash(w->pi_iter, k, fi, w->dx);
    for (struct RowEntry *e = w->rows[h % w->nbuckets]; e; e = e->next) {
        if (e->hash != h || e->pi_iter != w->pi_iter || e->key[w->dx] != k) continue;
        int same = 1;
        for (size_t i = 0; i < w->dx && same; i++) same = ((size_t)k == i) || e->key[i] == fi[i];
        if (same) return e;
    }
    return NULL;
}
static struct RowEntry *rows_add(struct Workspace *w, int32_t k, const int32_t *fi)
{
    struct RowEntry *e = xalloc(1, sizeof *e);
    e->hash = fiber_hash(w->pi_iter, k, fi, w->dx);
    e->pi_iter = w->pi_iter;
    memcpy(e->key, fi, w->dx * sizeof(int32_t));
    e->key[w->dx] = k;
    e->slot = (int32_t)w->nslots++;
    e->next = w->rows[e->hash % w->nbuckets];
    w->rows[e->hash % w->nbuckets] = e;
    return e;
}

/* ========================== MCA / DP / ControlParams ======================= */
struct MCAparam { size_t dx, du; size_t *ngrid; double **xgrid; double hmin, *hvec, h2, *t; };
struct MCAparam *mca_param_create(size_t dx, size_t du) { struct MCAparam *m = xalloc(1, sizeof *m); m->dx = dx; m->du = du; return m; }
void mca_add_grid_refs(struct MCAparam *m, size_t *ngrid, double **xgrid, double hmin, double *hvec)
{   /* grid is BORROWED (bellman.c:171-178); derived constants as bellman.c:181-186 -- two divisions
       per dimension done once on the host and uploaded, never recomputed on the device */
    m->ngrid = ngrid; m->xgrid = xgrid; m->hmin = hmin; m->hvec = hvec;
    m->h2 = hmin * hmin;
    free(m->t);
    m->t = xalloc(2 * m->dx, sizeof(double));
    for (size_t i = 0; i < m->dx; i++) { m->t[2 * i] = m->h2 / hvec[i]; m->t[2 * i + 1] = m->t[2 * i] / hvec[i]; }
}
void mca_param_destroy(struct MCAparam *m) { if (m) { free(m->t); free(m); } }

struct DPparam {
    struct Drift *drift; struct Diff *diff; struct Boundary *bound;
    double discount;
    int (*stagecost)(double, const double *, const double *, double *, double *);
    int (*boundcost)(double, const double *, double *);
    int (*obscost)(const double *, double *);
    int model, arith; double params[8]; size_t nparams;
};
struct DPparam *dp_param_create(size_t dx, size_t du, size_t dw, double discount)
{
    struct DPparam *dp = xalloc(1, sizeof *dp);
    dp->drift = drift_alloc(dx, du);
    dp->diff = diff_alloc(dx, du, dw);
    dp->discount = discount;
    dp->arith = C3SC_ARITH_FAST;
    return dp;
}
void dp_param_destroy(struct DPparam *dp) { if (dp) { drift_free(dp->drift); diff_free(dp->diff); free(dp); } }
void dp_param_add_drift(struct DPparam *dp, c3sc_dyn_cb b, void *a) { drift_add_func(dp->drift, b, a); }
void dp_param_add_diff(struct DPparam *dp, c3sc_dyn_cb s, void *a) { diff_add_func(dp->diff, s, a); }
void dp_param_add_boundary(struct DPparam *dp, struct Boundary *b) { dp->bound = b; }
void dp_param_add_stagecost(struct DPparam *dp, int (*f)(double, const double *, const double *, double *, double *)) { dp->stagecost = f; }
void dp_param_add_boundcost(struct DPparam *dp, int (*f)(double, const double *, double *)) { dp->boundcost = f; }
void dp_param_add_obscost(struct DPparam *dp, int (*f)(const double *, double *)) { dp->obscost = f; }
void dp_param_set_device_model(struct DPparam *dp, int model, const double *params, size_t n)
{
    dp->model = model;
    dp->nparams = n > 8 ? 8 : n;
    if (params) memcpy(dp->params, params, dp->nparams * sizeof(double));
}
void dp_param_set_arith(struct DPparam *dp, int arith) { dp->arith = arith; }

static c3sc_problem *make_problem(struct DPparam *dp, struct MCAparam *m, struct c3Opt *opt, size_t dw)
{
    if (!dp->model) { fprintf(stderr, "c3sc_b200: no device model registered (dp_param_set_device_model): host callbacks cannot run inside a kernel\n"); abort(); }
    if (!opt || !c3opt_is_bruteforce(opt) || !opt->n) { fprintf(stderr, "c3sc_b200: only the BRUTEFORCE control set is supported on the device\n"); abort(); }
    if (!dp->bound || !m->xgrid) { fprintf(stderr, "c3sc_b200: boundary / grid missing\n"); abort(); }
    const size_t dx = m->dx;
    if (dx < 1 || dx > C3SC_MAXD || dp->bound->nobs > C3SC_MAXOBS) { fprintf(stderr, "c3sc_b200: dx=%zu outside [1,%d] or more than %d obstacles\n", dx, C3SC_MAXD, C3SC_MAXOBS); abort(); }
    uint64_t ng[C3SC_MAXD]; int32_t bc[C3SC_MAXD];
    double lb[C3SC_MAXOBS * C3SC_MAXD], ub[C3SC_MAXOBS * C3SC_MAXD];
    for (size_t i = 0; i < dx; i++) { ng[i] = m->ngrid[i]; bc[i] = dp->bound->type[i]; }
    for (size_t k = 0; k < dp->bound->nobs; k++) {
        memcpy(lb + k * dx, dp->bound->obs_lb[k], dx * sizeof(double));
        memcpy(ub + k * dx, dp->bound->obs_ub[k], dx * sizeof(double));
    }
    c3sc_problem_desc d;
    memset(&d, 0, sizeof d);
    d.dx = (uint32_t)dx; d.du = (uint32_t)m->du; d.dw = (uint32_t)dw;
    d.ngrid = ng; d.xgrid = (const double *const *)m->xgrid; d.h2 = m->h2; d.t = m->t; d.bc = bc;
    d.nobs = (uint32_t)dp->bound->nobs; d.obs_lb = lb; d.obs_ub = ub;
    d.discount = dp->discount; d.nu = (uint32_t)opt->n; d.controls = opt->vals;
    d.model = dp->model; d.model_params = dp->nparams ? dp->params : NULL; d.n_model_params = (uint32_t)dp->nparams;
    d.arith = dp->arith;
    c3sc_problem *p = NULL;
    if (c3sc_problem_create(&d, &p)) die("c3sc_problem_create");
    return p;
}

struct ControlParams {
    double time; size_t dx, dw, N; const double *x;
    struct DPparam *dp; struct MCAparam *mca; struct Workspace *work; struct c3Opt *opt;
    int res_last_grad;
    c3sc_problem *dev;            /* built on first use from everything above */
};
struct ControlParams *control_params_create(size_t dx, size_t dw, struct DPparam *dp, struct MCAparam *mca,
                                            struct Workspace *work, struct c3Opt *opt)
{
    struct ControlParams *c = xalloc(1, sizeof *c);
    c->dx = dx; c->dw = dw; c->dp = dp; c->mca = mca; c->work = work; c->opt = opt;
    return c;
}
static c3sc_problem *cp_dev(struct ControlParams *c)
{
    if (!c->dev) c->dev = make_problem(c->dp, c->mca, c->opt, c->dw);
    return c->dev;
}
void control_params_add_time_and_states(struct ControlParams *c, double t, size_t N, const double *x) { c->time = t; c->N = N; c->x = x; }
int control_params_get_last_res(const struct ControlParams *c) { return c->res_last_grad; }
void control_params_destroy(struct ControlParams *c) { if (c) { c3sc_problem_destroy(c->dev); free(c); } }

size_t dp_param_check_device_model(struct DPparam *dp, struct MCAparam *m, struct c3Opt *opt, size_t n, double tol)
{
    const size_t dx = m->dx, du = m->du, dw = dp->diff->dw;
    c3sc_problem *p = make_problem(dp, m, opt, dw);
    double *x = xalloc(n * dx, 8), *u = xalloc(n * du, 8), *dr = xalloc(n * dx, 8), *sg = xalloc(n * dx, 8);
    double *st = xalloc(n, 8), *bd = xalloc(n, 8), *ob = xalloc(n, 8), *hd = xalloc(dx, 8), *hs = xalloc(dx * dw + 64, 8);
    for (size_t e = 0; e < n; e++) {
        for (size_t i = 0; i < dx; i++) x[e * dx + i] = m->xgrid[i][(e * (2 * i + 3) + i) % m->ngrid[i]];
        memcpy(u + e * du, opt->vals + (e % opt->n) * du, du * sizeof(double));
    }
    if (c3sc_model_eval(p, n, x, u, dr, sg, st, bd, ob)) die("c3sc_model_eval");
    size_t bad = 0;
    for (size_t e = 0; e < n; e++) {
        double v;
        drift_eval(dp->drift, 0.0, x + e * dx, u + e * du, hd, NULL);
        diff_eval(dp->diff, 0.0, x + e * dx, u + e * du, hs, NULL);
        for (size_t i = 0; i < dx; i++) {
            bad += fabs(hd[i] - dr[e * dx + i]) > tol * fmax(1.0, fabs(hd[i]));
            bad += fabs(hs[i * dx + i] - sg[e * dx + i]) > tol * fmax(1.0, fabs(hs[i * dx + i]));
        }
        if (dp->stagecost) { dp->stagecost(0.0, x + e * dx, u + e * du, &v, NULL); bad += fabs(v - st[e]) > tol * fmax(1.0, fabs(v)); }
        if (dp->boundcost) { dp->boundcost(0.0, x + e * dx, &v); bad += fabs(v - bd[e]) > tol * fmax(1.0, fabs(v)); }
        if (dp->obscost) { dp->obscost(x + e * dx, &v); bad += fabs(v - ob[e]) > tol * fmax(1.0, fabs(v)); }
    }
    free(x); free(u); free(dr); free(sg); free(st); free(bd); free(ob); free(hd); free(hs);
    c3sc_problem_destroy(p);
    return bad;
}

/* ========================== scalar pieces of the path ====================== */
double bellmanrhs(size_t dx, size_t du, double stage, const double *stage_grad, double discount, const double *prob,
                  const double *prob_grad, double dt, const double *dtgrad, const double *cost, double *grad)
{
    (void)du; (void)stage_grad; (void)prob_grad; (void)dtgrad;
    if (grad) { fprintf(stderr, "c3sc_b200: bellmanrhs gradient (BFGS path) is out of scope\n"); abort(); }
    double out;
    if (c3sc_rhs_batch(C3SC_ARITH_EXACT, (uint32_t)dx, discount, 1, prob, &dt, &stage, cost, &out)) die("bellmanrhs");
    return out;
}
int transition_assemble(size_t dx, size_t du, size_t dw, double h, const double *hvec, const double *drift,
                        const double *grad_drift, const double *ddiff, const double *grad_ddiff, double *prob,
                        double *grad_prob, double *dt, double *grad_dt, double *space)
{
    (void)du; (void)dw; (void)grad_drift; (void)grad_ddiff; (void)grad_dt; (void)space;
    if (grad_prob) { fprintf(stderr, "c3sc_b200: transition_assemble gradients (BFGS path) are out of scope\n"); abort(); }
    double sig[C3SC_MAXD];
    for (size_t i = 0; i < dx; i++) sig[i] = ddiff[i * dx + i];         /* only the diagonal is read, nodeutil.c:294 */
    int32_t st;
    if (c3sc_transition_raw(C3SC_ARITH_EXACT, (uint32_t)dx, h, hvec, 1, drift, sig, prob, dt, &st)) die("transition_assemble");
    return st;
}
int convert_fiber_to_ind(size_t d, size_t N, const double *x, const size_t *Ngrid, double **xgrid,
                         size_t *fixed_ind, size_t *dim_vary)
{   /* index decode only: first point -> indices, second point -> first differing dimension */
    for (size_t i = 0; i < d; i++) {
        size_t hit = Ngrid[i];
        for (size_t j = 0; j < Ngrid[i]; j++) if (fabs(x[i] - xgrid[i][j]) < 1e-14) { hit = j; break; }
        if (hit == Ngrid[i]) {
            fprintf(stderr, "Error: evaluation is not on the grid\nx[%zu] = %3.15E\n", i, x[i]);
            return 1;
        }
        fixed_ind[i] = hit;
    }
    *dim_vary = d;
    for (size_t i = 0; i < d && *dim_vary == d; i++) {
        size_t hit = Ngrid[i];
        for (size_t j = 0; j < Ngrid[i]; j++) if (fabs(x[d + i] - xgrid[i][j]) < 1e-14) { hit = j; break; }
        if (hit != fixed_ind[i]) *dim_vary = i;
    }
    if (*dim_vary == d) return 1;
    return N != Ngrid[*dim_vary] ? 2 : 0;
}

/* node-level objective / argmin (struct Memory protocol of bellman.c:59-63,367,504) */
static void node_args(void *arg, struct ControlParams **cp, size_t *node, const double **x, int *ab, double **costs)
{
    struct c3sc_memory *mem = arg;
    *cp = mem->shared; *node = mem->private_;
    *x = (*cp)->x + *node * (*cp)->dx;
    *ab = *workspace_get_absorbed((*cp)->work, *node);
    *costs = workspace_get_costs((*cp)->work, *node);
}
double bellman_control(size_t du, const double *u, double *grad_u, void *args)
{
    struct ControlParams *cp; size_t node; const double *x; int ab; double *costs, val;
    (void)du;
    if (grad_u) { fprintf(stderr, "c3sc_b200: bellman_control gradient (BFGS path) is out of scope\n"); abort(); }
    node_args(args, &cp, &node, &x, &ab, &costs);
    if (ab != 0) {
        int32_t a32 = ab;
        if (c3sc_node_backup_batch(cp_dev(cp), 1, x, costs, &a32, &val, NULL)) die("bellman_control");
        return val;
    }
    if (c3sc_control_value_batch(cp_dev(cp), 1, x, u, costs, &val)) die("bellman_control");
    {   /* the node's slab as the reference leaves it (src/bellman.c:400-449): drift, diffusion (column-major
           dx x dw, diagonal), dt, probabilities -- from the device model and the device transition kernel */
        struct Workspace *w = cp->work;
        const size_t dx = cp->dx, dw = cp->dw;
        double sg[C3SC_MAXD], st, bd, ob;
        int32_t status;
        if (node < w->N && !c3sc_model_eval(cp_dev(cp), 1, x, u, workspace_get_drift(w, node), sg, &st, &bd, &ob) &&
            !c3sc_transition_batch(cp_dev(cp), 1, workspace_get_drift(w, node), sg, workspace_get_prob(w, node),
                                   workspace_get_dt(w, node), &status)) {
            double *df = workspace_get_diff(w, node);
            memset(df, 0, dx * dw * sizeof(double));
            for (size_t i = 0; i < dx && i < dw; i++) df[i * dx + i] = sg[i];
            cp->res_last_grad = status;
        }
    }
    return val;
}
int bellman_optimal(size_t du, double *u, double *val, void *arg)
{
    struct ControlParams *cp; size_t node; const double *x; int ab; double *costs;
    node_args(arg, &cp, &node, &x, &ab, &costs);
    int32_t a32 = ab, best = -1;
    int rc = c3sc_node_backup_batch(cp_dev(cp), 1, x, costs, &a32, val, &best);
    if (rc) return rc;
    for (size_t i = 0; i < du; i++) u[i] = best >= 0 ? cp->opt->vals[(size_t)best * du + i] : 0.0;
    if (node < cp->work->N) memcpy(workspace_get_u(cp->work, node), u, du * sizeof(double));
    return 0;
}

/* ========================== fiber operators ================================ */
struct VIparam { struct ControlParams *cp; struct ValueF *vf; size_t nstate_evals, nnode_evals; double convergence; };
struct VIparam *vi_param_create(double conv) { struct VIparam *v = xalloc(1, sizeof *v); v->convergence = conv; return v; }
void vi_param_destroy(struct VIparam *v) { free(v); }
void vi_param_add_cp(struct VIparam *v, struct ControlParams *cp) { v->cp = cp; }
void vi_param_add_value(struct VIparam *v, struct ValueF *vf) { v->vf = vf; v->nstate_evals = 0; v->nnode_evals = 0; }
size_t vi_param_get_nstate_evals(const struct VIparam *v) { return v->nstate_evals; }

struct PIparam { struct ControlParams *cp; struct ValueF *vf_iteration, *vf_policy; size_t npol_evals, niter_evals; double convergence; };
struct PIparam *pi_param_create(double conv, struct ValueF *policy)
{
    struct PIparam *p = xalloc(1, sizeof *p);
    p->convergence = conv; p->vf_policy = policy;
    return p;
}
void pi_param_destroy(struct PIparam *p) { free(p); }
void pi_param_add_cp(struct PIparam *p, struct ControlParams *cp) { p->cp = cp; }
void pi_param_add_value(struct PIparam *p, struct ValueF *vf) { p->vf_iteration = vf; p->niter_evals = 0; }
size_t pi_param_get_npol_evals(const struct PIparam *p) { return p->npol_evals; }

static size_t grid_nmax(const struct MCAparam *m)
{
    size_t n = 0;
    for (size_t i = 0; i < m->dx; i++) if (m->ngrid[i] > n) n = m->ngrid[i];
    return n;
}
/* decode F point-described fibers laid back to back; returns total node count or 0 on error */
static size_t decode_fibers(const struct MCAparam *m, size_t F, const double *x, int32_t *dv, int32_t *fi, size_t *offs)
{
    const size_t d = m->dx;
    size_t pos = 0, fixed[C3SC_MAXD], k;
    for (size_t f = 0; f < F; f++) {
        /* N is implied by the varying dimension: decode with the two leading points */
        size_t Ntry = 0;
        int rc = convert_fiber_to_ind(d, 0, x + pos * d, m->ngrid, m->xgrid, fixed, &k);
        if (rc == 1) return 0;
        Ntry = m->ngrid[k];
        for (size_t i = 0; i < d; i++) fi[f * d + i] = (int32_t)fixed[i];
        dv[f] = (int32_t)k;
        offs[f] = pos;
        pos += Ntry;
    }
    offs[F] = pos;
    return pos;
}

int bellman_vi_batch_ind(size_t F, const int32_t *dv, const int32_t *fi, double *out, void *arg)
{
    struct VIparam *v = arg;
    struct ControlParams *cp = v->cp;
    const size_t nmax = grid_nmax(cp->mca);
    int rc = c3sc_vi_batch(cp_dev(cp), v->vf->dev, F, dv, fi, nmax, out, NULL);
    if (rc) { fprintf(stderr, "c3sc_b200: bellman_vi: %s\n", c3sc_last_error()); return rc; }
    for (size_t f = 0; f < F; f++) { v->nstate_evals += cp->mca->ngrid[dv[f]]; v->nnode_evals += cp->mca->ngrid[dv[f]]; }
    return 0;
}
int bellman_vi_batch(size_t F, const double *x, double *out, void *arg)
{
    struct VIparam *v = arg;
    const struct MCAparam *m = v->cp->mca;
    const size_t d = m->dx, nmax = grid_nmax(m);
    int32_t *dv = xalloc(F, 4), *fi = xalloc(F * d, 4);
    size_t *offs = xalloc(F + 1, sizeof(size_t));
    double *tmp = xalloc(F * nmax, 8);
    int rc = decode_fibers(m, F, x, dv, fi, offs) ? 0 : 1;
    if (!rc) rc = bellman_vi_batch_ind(F, dv, fi, tmp, arg);
    for (size_t f = 0; f < F && !rc; f++) memcpy(out + offs[f], tmp + f * nmax, (offs[f + 1] - offs[f]) * sizeof(double));
    free(dv); free(fi); free(offs); free(tmp);
    return rc;
}
int bellman_vi(size_t N, const double *x, double *out, void *arg)
{
    struct VIparam *v = arg;
    const struct MCAparam *m = v->cp->mca;
    size_t fixed[C3SC_MAXD], k;
    int rc = convert_fiber_to_ind(m->dx, N, x, m->ngrid, m->xgrid, fixed, &k);
    if (rc) return rc;
    control_params_add_time_and_states(v->cp, 0.0, N, x);
    return bellman_vi_batch(1, x, out, arg);
}

int bellman_pi_batch_ind(size_t F, const int32_t *dv, const int32_t *fi, double *out, void *arg)
{
    struct PIparam *p = arg;
    struct ControlParams *cp = p->cp;
    struct Workspace *w = cp->work;
    const size_t d = cp->mca->dx, nmax = grid_nmax(cp->mca);
    /* The policy rows [p(2dx+1), dt, g] of src/bellman.c:1810 live in a device store owned by the Workspace, one slot
       per fiber of the current policy (pi_iter); only fiber descriptors go up and values come down.  Split the batch
       into fibers whose rows are filed (this pi_iter) and new ones. */
    if (!w->store && c3sc_rowstore_create((uint32_t)d, nmax, &w->store)) die("c3sc_rowstore_create");
    size_t nh = 0, nm = 0;
    size_t *ih = xalloc(F, sizeof(size_t)), *im = xalloc(F, sizeof(size_t));
    struct RowEntry **ent = xalloc(F, sizeof *ent);
    for (size_t f = 0; f < F; f++) {
        ent[f] = rows_find(w, dv[f], fi + f * d);
        if (ent[f]) ih[nh++] = f; else im[nm++] = f;
    }
    int rc = 0;
    for (int pass = 0; pass < 2 && !rc; pass++) {
        const size_t n = pass ? nh : nm, *idx = pass ? ih : im;
        if (!n) continue;
        int32_t *bdv = xalloc(n, 4), *bfi = xalloc(n * d, 4), *slot = xalloc(n, 4);
        double *val = xalloc(n * nmax, 8);
        for (size_t q = 0; q < n; q++) {
            bdv[q] = dv[idx[q]];
            memcpy(bfi + q * d, fi + idx[q] * d, d * 4);
            if (pass) slot[q] = ent[idx[q]]->slot;
            else {
                /* a fiber requested twice in one batch files its rows once */
                struct RowEntry *e = rows_find(w, bdv[q], bfi + q * d);
                if (!e) e = rows_add(w, bdv[q], bfi + q * d);
                slot[q] = e->slot;
            }
        }
        if (!pass) {
            size_t cap = 1024;
            while (cap < w->nslots) cap *= 2;
            rc = c3sc_rowstore_reserve(w->store, cap);
        }
        if (!rc) rc = c3sc_pi_batch_store(cp_dev(cp), p->vf_policy->dev, p->vf_iteration->dev, n, bdv, bfi, nmax, pass, w->store, slot, val);
        for (size_t q = 0; q < n && !rc; q++) {
            memcpy(out + idx[q] * nmax, val + q * nmax, nmax * 8);
            if (!pass) p->npol_evals += cp->mca->ngrid[bdv[q]];
            p->niter_evals += cp->mca->ngrid[bdv[q]];
        }
        free(bdv); free(bfi); free(slot); free(val);
    }
    if (rc) fprintf(stderr, "c3sc_b200: bellman_pi: %s\n", c3sc_last_error());
    free(ih); free(im); free(ent);
    return rc;
}
int bellman_pi_batch(size_t F, const double *x, double *out, void *arg)
{
    struct PIparam *p = arg;
    const struct MCAparam *m = p->cp->mca;
    const size_t d = m->dx, nmax = grid_nmax(m);
    int32_t *dv = xalloc(F, 4), *fi = xalloc(F * d, 4);
    size_t *offs = xalloc(F + 1, sizeof(size_t));
    double *tmp = xalloc(F * nmax, 8);
    int rc = decode_fibers(m, F, x, dv, fi, offs) ? 0 : 1;
    if (!rc) rc = bellman_pi_batch_ind(F, dv, fi, tmp, arg);
    for (size_t f = 0; f < F && !rc; f++) memcpy(out + offs[f], tmp + f * nmax, (offs[f + 1] - offs[f]) * sizeof(double));
    free(dv); free(fi); free(offs); free(tmp);
    return rc;
}
int bellman_pi(size_t N, const double *x, double *out, void *arg)
{
    struct PIparam *p = arg;
    const struct MCAparam *m = p->cp->mca;
    size_t fixed[C3SC_MAXD], k;
    int rc = convert_fiber_to_ind(m->dx, N, x, m->ngrid, m->xgrid, fixed, &k);
    if (rc) return rc;
    control_params_add_time_and_states(p->cp, 0.0, N, x);
    return bellman_pi_batch(1, x, out, arg);
}

/* geometry-only device problem (C3SC_MODEL_NONE): grid, boundary types, obstacle boxes -- what the flag and
   neighbour-value entries need; no dynamics, no control set */
static c3sc_problem *geometry_problem(size_t d, const struct Boundary *bound, const size_t *ngrid, double *const *xgrid)
{
    if (d < 1 || d > C3SC_MAXD || !bound || bound->nobs > C3SC_MAXOBS) return NULL;
    uint64_t ng[C3SC_MAXD]; int32_t bc[C3SC_MAXD];
    double t[2 * C3SC_MAXD], lb[C3SC_MAXOBS * C3SC_MAXD], ub[C3SC_MAXOBS * C3SC_MAXD];
    for (size_t i = 0; i < d; i++) { ng[i] = ngrid[i]; bc[i] = bound->type[i]; t[2 * i] = t[2 * i + 1] = 1.0; }
    for (size_t k = 0; k < bound->nobs; k++) {
        memcpy(lb + k * d, bound->obs_lb[k], d * sizeof(double));
        memcpy(ub + k * d, bound->obs_ub[k], d * sizeof(double));
    }
    c3sc_problem_desc ds;
    memset(&ds, 0, sizeof ds);
    ds.dx = (uint32_t)d; ds.du = 1; ds.dw = (uint32_t)d;
    ds.ngrid = ng; ds.xgrid = (const double *const *)xgrid; ds.h2 = 1.0; ds.t = t; ds.bc = bc;
    ds.nobs = (uint32_t)bound->nobs; ds.obs_lb = lb; ds.obs_ub = ub;
    ds.model = C3SC_MODEL_NONE; ds.arith = C3SC_ARITH_FAST;
    c3sc_problem *p = NULL;
    if (c3sc_problem_create(&ds, &p)) return NULL;
    return p;
}

/* src/nodeutil.c:647-713 */
int mca_get_neighbor_costs(size_t d, size_t N, const double *x, struct Boundary *bound, struct ValueF *vf,
                           const size_t *ngrid, double **xgrid, size_t *fixed_ind, size_t *dim_vary,
                           int *absorbed, double *out)
{
    int rc = convert_fiber_to_ind(d, N, x, ngrid, xgrid, fixed_ind, dim_vary);
    if (rc) return rc;
    c3sc_problem *p = geometry_problem(d, bound, ngrid, xgrid);
    if (!p) { fprintf(stderr, "c3sc_b200: mca_get_neighbor_costs: %s\n", c3sc_last_error()); return 1; }
    size_t nmax = 0;
    for (size_t i = 0; i < d; i++) if (ngrid[i] > nmax) nmax = ngrid[i];
    int32_t dv = (int32_t)*dim_vary, fi[C3SC_MAXD];
    for (size_t i = 0; i < d; i++) fi[i] = (int32_t)fixed_ind[i];
    int32_t *ab = xalloc(nmax, 4);
    double *costs = xalloc(nmax * (2 * d + 1), 8);
    rc = c3sc_neighbor_costs_batch(p, vf->dev, 1, &dv, fi, nmax, ab, costs, NULL, NULL);
    if (!rc) {
        for (size_t j = 0; j < N; j++) absorbed[j] = ab[j];
        memcpy(out, costs, N * (2 * d + 1) * sizeof(double));
    }
    free(ab); free(costs);
    c3sc_problem_destroy(p);
    return rc;
}

/* src/nodeutil.c:489-627.  The reference signature carries the fiber's points but no grids: the coordinates the
   obstacle test needs are all in x (the fixed ones in its first point, the varying one in every point), so a
   geometry-only problem is built whose grids hold exactly those coordinates at the fiber's indices. */
int process_fibers_neighbor(size_t d, const size_t *fixed_ind, size_t dim_vary, const double *x, int *absorbed,
                            size_t *neighbors_vary, size_t *neighbors_fixed, const size_t *ngrid,
                            const struct Boundary *bound)
{
    if (d < 1 || d > C3SC_MAXD || dim_vary >= d || !bound) return 1;
    const size_t N = ngrid[dim_vary];
    size_t nmax = 0;
    for (size_t i = 0; i < d; i++) if (ngrid[i] > nmax) nmax = ngrid[i];
    double *xg[C3SC_MAXD];
    for (size_t i = 0; i < d; i++) {
        xg[i] = xalloc(ngrid[i], sizeof(double));
        if (i == dim_vary) for (size_t j = 0; j < N; j++) xg[i][j] = x[j * d + i];
        else for (size_t j = 0; j < ngrid[i]; j++) xg[i][j] = x[i];        /* only entry fixed_ind[i] is read */
    }
    c3sc_problem *p = geometry_problem(d, bound, ngrid, xg);
    int rc = p ? 0 : 1;
    int32_t dv = (int32_t)dim_vary, fi[C3SC_MAXD];
    for (size_t i = 0; i < d; i++) fi[i] = (int32_t)(i == dim_vary ? 0 : fixed_ind[i]);
    int32_t *ab = xalloc(nmax, 4), *nv = xalloc(2 * nmax, 4), *nf = xalloc(2 * d, 4);
    if (!rc) rc = c3sc_fiber_flags_batch(p, 1, &dv, fi, nmax, ab, nv, nf);
    if (rc) fprintf(stderr, "c3sc_b200: process_fibers_neighbor: %s\n", c3sc_last_error());
    else {
        for (size_t j = 0; j < N; j++) { absorbed[j] = ab[j]; neighbors_vary[2 * j] = (size_t)nv[2 * j]; neighbors_vary[2 * j + 1] = (size_t)nv[2 * j + 1]; }
        for (size_t q = 0; q + 2 < 2 * d; q++) neighbors_fixed[q] = (size_t)nf[q];
    }
    free(ab); free(nv); free(nf);
    for (size_t i = 0; i < d; i++) free(xg[i]);
    c3sc_problem_destroy(p);
    return rc;
}

/* src/nodeutil.c:718-816 */
int mca_get_neighbor_node_costs(size_t d, const double *x, struct Boundary *bound, struct ValueF *vf,
                                const size_t *ngrid, double **xgrid, int *absorbed, double *out)
{
    c3sc_problem *p = geometry_problem(d, bound, ngrid, xgrid);
    if (!p) { fprintf(stderr, "c3sc_b200: mca_get_neighbor_node_costs: %s\n", c3sc_last_error()); return 1; }
    int32_t ab = 0;
    int rc = c3sc_neighbor_node_costs_batch(p, vf->dev, 1, x, &ab, out);
    if (rc) fprintf(stderr, "c3sc_b200: mca_get_neighbor_node_costs: %s\n", c3sc_last_error());
    *absorbed = ab;
    c3sc_problem_destroy(p);
    return rc;
}

/* ========================== C3Control facade =============================== */
struct C3Control {
    size_t dx, du, dw;
    size_t *ngrid; double **xgrid, *h, hmin;
    struct Boundary *bound; struct MCAparam *mca; struct DPparam *dp; struct Workspace *work;
    struct ValueF *policy_sim; struct c3Opt *opt_sim; void (*transform_sim)(size_t, const double *, double *);
    struct ControlParams *cp_sim; double *prevpol;
};
struct C3Control *c3control_create(size_t dx, size_t du, size_t dw, double *lb, double *ub, size_t *ngrid, double discount)
{
    struct C3Control *c = xalloc(1, sizeof *c);
    c->dx = dx; c->du = du; c->dw = dw; c->ngrid = ngrid;
    c->xgrid = xalloc(dx, sizeof(double *));
    c->h = xalloc(dx, sizeof(double));
    c->hmin = ub[0] - lb[0];
    size_t nmax = ngrid[0];
    for (size_t i = 0; i < dx; i++) {
        /* C3 linspace: running sum of the interval (bellman.c:1979) */
        double *g = xalloc(ngrid[i], sizeof(double));
        const double step = (ub[i] - lb[i]) / (double)(ngrid[i] - 1);
        g[0] = lb[i];
        for (size_t j = 1; j < ngrid[i]; j++) g[j] = g[j - 1] + step;
        c->xgrid[i] = g;
        c->h[i] = g[1] - g[0];
        if (c->h[i] < c->hmin) c->hmin = c->h[i];
        if (ngrid[i] > nmax) nmax = ngrid[i];
    }
    c->bound = boundary_alloc(dx, lb, ub);
    c->mca = mca_param_create(dx, du);
    mca_add_grid_refs(c->mca, c->ngrid, c->xgrid, c->hmin, c->h);
    c->dp = dp_param_create(dx, du, dw, discount);
    dp_param_add_boundary(c->dp, c->bound);
    c->work = workspace_alloc(dx, du, dw, nmax);
    return c;
}
void c3control_destroy(struct C3Control *c)
{
    if (!c) return;
    control_params_destroy(c->cp_sim); free(c->prevpol);
    boundary_free(c->bound); mca_param_destroy(c->mca); dp_param_destroy(c->dp); workspace_free(c->work);
    for (size_t i = 0; i < c->dx; i++) free(c->xgrid[i]);
    free(c->xgrid); free(c->h); free(c);
}
size_t *c3control_get_ngrid(struct C3Control *c) { return c ? c->ngrid : NULL; }
double **c3control_get_xgrid(struct C3Control *c) { return c ? c->xgrid : NULL; }
void c3control_set_external_boundary(struct C3Control *c, size_t dim, char *type) { boundary_external_set_type(c->bound, dim, type); }
void c3control_add_obstacle(struct C3Control *c, double *center, double *widths) { boundary_add_obstacle(c->bound, center, widths); }
void c3control_add_drift(struct C3Control *c, c3sc_dyn_cb b, void *a) { dp_param_add_drift(c->dp, b, a); }
void c3control_add_diff(struct C3Control *c, c3sc_dyn_cb s, void *a) { dp_param_add_diff(c->dp, s, a); }
void c3control_add_stagecost(struct C3Control *c, int (*f)(double, const double *, const double *, double *, double *)) { dp_param_add_stagecost(c->dp, f); }
void c3control_add_boundcost(struct C3Control *c, int (*f)(double, const double *, double *)) { dp_param_add_boundcost(c->dp, f); }
void c3control_add_obscost(struct C3Control *c, int (*f)(const double *, double *)) { dp_param_add_obscost(c->dp, f); }
void c3control_set_device_model(struct C3Control *c, int model, const double *params, size_t n) { dp_param_set_device_model(c->dp, model, params, n); }
struct DPparam *c3control_get_dp(struct C3Control *c) { return c->dp; }
struct MCAparam *c3control_get_mca(struct C3Control *c) { return c->mca; }
struct Workspace *c3control_get_work(struct C3Control *c) { return c->work; }
struct Boundary *c3control_get_boundary(struct C3Control *c) { return c->bound; }

int c3control_vi_fibers(struct C3Control *c, struct ValueF *vf, struct c3Opt *opt, size_t F, const int32_t *dv,
                        const int32_t *fi, double *out, size_t *nevals)
{
    struct ControlParams *cp = control_params_create(c->dx, c->dw, c->dp, c->mca, c->work, opt);
    struct VIparam *vi = vi_param_create(1e-10);
    vi_param_add_cp(vi, cp);
    vi_param_add_value(vi, vf);
    workspace_increment_vi_iter(c->work);
    int rc = bellman_vi_batch_ind(F, dv, fi, out, vi);
    if (nevals) *nevals = vi->nnode_evals;
    vi_param_destroy(vi);
    control_params_destroy(cp);
    return rc;
}


/* ========================== approximation arguments (src/util.c:105-230) ==== */
struct ApproxArgs { double cross_tol, round_tol; size_t kickrank, startrank, maxrank; int adapt; enum function_class fc; };
struct ApproxArgs *approx_args_init(void)
{
    struct ApproxArgs *a = xalloc(1, sizeof *a);
    a->cross_tol = 1e-10; a->round_tol = 1e-10; a->kickrank = 10; a->startrank = 5; a->maxrank = 40; a->adapt = 1; a->fc = LINELM;
    return a;
}
void approx_args_free(struct ApproxArgs *a) { free(a); }
void approx_args_set_function_class(struct ApproxArgs *a, enum function_class fc) { a->fc = fc; }
enum function_class approx_args_get_function_class(const struct ApproxArgs *a) { return a->fc; }
void approx_args_set_cross_tol(struct ApproxArgs *a, double v) { a->cross_tol = v; }
double approx_args_get_cross_tol(const struct ApproxArgs *a) { return a->cross_tol; }
void approx_args_set_round_tol(struct ApproxArgs *a, double v) { a->round_tol = v; }
double approx_args_get_round_tol(const struct ApproxArgs *a) { return a->round_tol; }
void approx_args_set_kickrank(struct ApproxArgs *a, size_t v) { a->kickrank = v; }
size_t approx_args_get_kickrank(const struct ApproxArgs *a) { return a->kickrank; }
void approx_args_set_maxrank(struct ApproxArgs *a, size_t v) { a->maxrank = v; }
size_t approx_args_get_maxrank(const struct ApproxArgs *a) { return a->maxrank; }
void approx_args_set_startrank(struct ApproxArgs *a, size_t v) { a->startrank = v; }
size_t approx_args_get_startrank(const struct ApproxArgs *a) { return a->startrank; }
void approx_args_set_adapt(struct ApproxArgs *a, int v) { a->adapt = v; }
int approx_args_get_adapt(const struct ApproxArgs *a) { return a->adapt; }

/* ========================== valuef_interp and friends ======================= */
/* The operator behind a cross run: bellman_vi / bellman_pi go to the device as whole-core batches; any other
 * fiber function (a start cost handed to c3control_init_value) is the caller's own host code and is called
 * once per fiber with the points the reference would hand it. */
struct interp_ctx {
    int (*f)(size_t, const double *, double *, void *); void *args;
    size_t d; const size_t *N; double **grid;
};
static int interp_cb(size_t F, const int32_t *dv, const int32_t *fi, size_t ldo, double *out, void *arg)
{
    struct interp_ctx *c = arg;
    if (c->f == bellman_vi) return bellman_vi_batch_ind(F, dv, fi, out, c->args);
    if (c->f == bellman_pi) return bellman_pi_batch_ind(F, dv, fi, out, c->args);
    double *x = xalloc(ldo * c->d, sizeof(double));
    int rc = 0;
    for (size_t f = 0; f < F && !rc; f++) {
        const size_t k = (size_t)dv[f], n = c->N[k];
        for (size_t j = 0; j < n; j++)
            for (size_t i = 0; i < c->d; i++) x[j * c->d + i] = c->grid[i][i == k ? j : (size_t)fi[f * c->d + i]];
        rc = c->f(n, x, out + f * ldo, c->args);
    }
    free(x);
    return rc;
}

/* valuef_interp (src/valuefunc.c:603-767) over the batched cross driver */
struct ValueF *valuef_interp(size_t d, int (*f)(size_t, const double *, double *, void *), void *args, const size_t *N,
                             double **grid, struct ValueF *vref, struct ApproxArgs *aargs, int verbose)
{
    if (aargs->fc != LINELM) die("valuef_interp: only nodal LINELM value functions are supported");
    uint64_t n64[C3SC_MAXD], r64[C3SC_MAXD + 1], minN = N[0];
    for (size_t k = 0; k < d; k++) { n64[k] = N[k]; if (N[k] < minN) minN = N[k]; }
    const uint64_t maxrank = aargs->maxrank < minN ? aargs->maxrank : minN;          /* :625-631 */
    c3sc_cross *cr = NULL;
    r64[0] = r64[d] = 1;
    if (vref && aargs->adapt == 1) {                                                  /* :637-648, :706-712 */
        for (size_t k = 1; k < d; k++) r64[k] = vref->ranks[k] + 1 >= maxrank ? maxrank : vref->ranks[k] + 1;
        if (vref->cross) {
            if (c3sc_cross_copy(vref->cross, &cr) || c3sc_cross_set_ranks(cr, r64)) die("valuef_interp: index sets");
        }
    } else {
        for (size_t k = 1; k < d; k++) r64[k] = aargs->startrank;
    }
    if (!cr && c3sc_cross_create((uint32_t)d, n64, r64, &cr)) die("valuef_interp: c3sc_cross_create");
    if (verbose > 0) {
        c3sc_cross_ranks(cr, r64);
        printf("Starting Ranks: ");
        for (size_t k = 0; k <= d; k++) printf("%llu ", (unsigned long long)r64[k]);
        printf("\n");
    }
    struct interp_ctx ctx = { f, args, d, N, grid };
    c3sc_cross_opts o = { 5, aargs->cross_tol, verbose > 1 };                          /* maxiter 5, :632 */
    c3sc_adapt_opts ao = { aargs->adapt == 1 ? (uint32_t)aargs->kickrank : 0, (uint32_t)maxrank, aargs->round_tol, 0 };
    uint64_t cap[C3SC_MAXD + 1], rout[C3SC_MAXD + 1];
    double *cores[C3SC_MAXD];
    if (aargs->adapt == 1) c3sc_cross_adapt_capacity(cr, &ao, cap);
    else c3sc_cross_ranks(cr, cap);
    for (size_t k = 0; k < d; k++) cores[k] = xalloc(N[k] * cap[k] * cap[k + 1], sizeof(double));
    int rc;
    if (aargs->adapt == 1) rc = c3sc_cross_run_adapt(cr, interp_cb, &ctx, &o, &ao, rout, cores, NULL, NULL);
    else { rc = c3sc_cross_run(cr, interp_cb, &ctx, &o, cores, NULL, NULL); c3sc_cross_ranks(cr, rout); }
    if (rc) die("valuef_interp: cross approximation failed");
    size_t ranks[C3SC_MAXD + 1];
    for (size_t k = 0; k <= d; k++) ranks[k] = (size_t)rout[k];
    if (verbose > 0) {
        printf("Final Ranks: ");
        for (size_t k = 0; k <= d; k++) printf("%zu ", ranks[k]);
        printf("\n");
    }
    struct ValueF *vf = valuef_from_cores(d, N, ranks, cores);
    vf->cross = cr;
    valuef_set_grid(vf, grid);
    for (size_t k = 0; k < d; k++) free(cores[k]);
    return vf;
}

static void vf_u64(const struct ValueF *v, uint64_t *n, uint64_t *r)
{
    for (size_t k = 0; k < v->d; k++) n[k] = v->N[k];
    for (size_t k = 0; k <= v->d; k++) r[k] = v->ranks[k];
}
/* valuef_norm / valuef_norm2diff (src/valuefunc.c:315-335): the reference's function_train_norm2 / norm2diff, i.e. the
   continuous L2 norm over the box of the piecewise-linear train (c3sc_cores_*_l2: mass matrices of the hat functions).
   A value function that does not know its grid (valuef_from_cores without valuef_set_grid) has no such norm: loud. */
double valuef_norm(struct ValueF *v)
{
    uint64_t n[C3SC_MAXD], r[C3SC_MAXD + 1];
    if (!v->xgrid) die("valuef_norm: the value function has no grid (valuef_set_grid); valuef_norm_nodal needs none");
    vf_u64(v, n, r);
    /* next to the cores on the device; ranks beyond the kernel's shared memory take the host restatement */
    double out = 0.0;
    if (v->dev && c3sc_valuef_norm_l2(v->dev, (const double *const *)v->xgrid, &out) == C3SC_OK) return out;
    return c3sc_cores_norm_l2((uint32_t)v->d, n, (const double *const *)v->xgrid, r, (const double *const *)v->cores);
}
double valuef_norm2diff(struct ValueF *a, struct ValueF *b)
{
    uint64_t n[C3SC_MAXD], ra[C3SC_MAXD + 1], rb[C3SC_MAXD + 1];
    double **xg = a->xgrid ? a->xgrid : b->xgrid;            /* both functions live on the same grid */
    if (!xg) die("valuef_norm2diff: neither value function has a grid (valuef_set_grid); valuef_norm2diff_nodal needs none");
    vf_u64(a, n, ra); vf_u64(b, n, rb);
    double out = 0.0;
    if (a->dev && b->dev && c3sc_valuef_norm2diff_l2(a->dev, b->dev, (const double *const *)xg, &out) == C3SC_OK) return out;
    return c3sc_cores_norm2diff_l2((uint32_t)a->d, n, (const double *const *)xg, ra, (const double *const *)a->cores, rb,
                                   (const double *const *)b->cores);
}
/* NEW names: the discrete l2 of the node values (what valuef_norm was in the first round of this library) */
double valuef_norm_nodal(struct ValueF *v)
{
    uint64_t n[C3SC_MAXD], r[C3SC_MAXD + 1];
    vf_u64(v, n, r);
    return c3sc_cores_norm((uint32_t)v->d, n, r, (const double *const *)v->cores);
}
double valuef_norm2diff_nodal(struct ValueF *a, struct ValueF *b)
{
    uint64_t n[C3SC_MAXD], ra[C3SC_MAXD + 1], rb[C3SC_MAXD + 1];
    vf_u64(a, n, ra); vf_u64(b, n, rb);
    return c3sc_cores_norm2diff((uint32_t)a->d, n, ra, (const double *const *)a->cores, rb, (const double *const *)b->cores);
}
/* valuef_eval (src/valuefunc.c:345-350): piecewise-linear interpolation of the nodal cores, on the device */
double valuef_eval(struct ValueF *v, const double *x)
{
    if (!v->xgrid) die("valuef_eval: the value function has no grid (valuef_set_grid)");
    if (!v->eval_dev) {
        double lo[C3SC_MAXD], hi[C3SC_MAXD];
        for (size_t i = 0; i < v->d; i++) { lo[i] = v->xgrid[i][0]; hi[i] = v->xgrid[i][v->N[i] - 1]; }
        struct Boundary *b = boundary_alloc(v->d, lo, hi);
        struct MCAparam m = { v->d, 1, v->N, v->xgrid, 1.0, NULL, 1.0, NULL };
        double t[2 * C3SC_MAXD], u0[C3SC_MAXD] = { 0 };
        for (size_t i = 0; i < 2 * v->d; i++) t[i] = 1.0;
        m.t = t;
        struct DPparam dp;
        memset(&dp, 0, sizeof dp);
        dp.bound = b;
        dp.arith = C3SC_ARITH_FAST;
        dp.model = (v->d % 2 == 0) ? C3SC_MODEL_LQGND : (v->d == 3 ? C3SC_MODEL_DUBINS : C3SC_MODEL_SKID5D);
        m.du = (dp.model == C3SC_MODEL_LQGND) ? v->d / 2 : 1;
        struct c3Opt opt = { BRUTEFORCE, m.du, 1, u0 };
        v->eval_dev = make_problem(&dp, &m, &opt, v->d);
        boundary_free(b);
    }
    double out = 0.0;
    if (c3sc_valuef_eval_batch(v->eval_dev, v->dev, 1, x, &out)) die("valuef_eval");
    return out;
}

/* ========================== checkpoint / resume (src/valuefunc.c:226-295) === */
/* The reference saves through C3's function_train_save / _savetxt; those formats belong to the absent
 * library, so the files written here are this library's own (not interchangeable with .c3 files): ranks, the
 * nodes the cores live on and the nodal cores, binary (valuef_save) or text with 21 digits (valuef_savetxt).
 * Loading re-samples the cores on the caller's grid when it differs from the saved one, which is what
 * function_train_create_nodal does in valuef_load (piecewise-linear cores, zero outside the saved nodes). */
static const char VF_MAGIC[8] = { 'C', '3', 'S', 'C', 'V', 'F', '0', '1' };
static int vf_write(struct ValueF *v, FILE *fp, int text)
{
    if (!v->xgrid) return 1;
    if (text) {
        fprintf(fp, "C3SCVF01 %zu\n", v->d);
        for (size_t k = 0; k < v->d; k++) fprintf(fp, "%zu ", v->N[k]);
        fprintf(fp, "\n");
        for (size_t k = 0; k <= v->d; k++) fprintf(fp, "%zu ", v->ranks[k]);
        fprintf(fp, "\n");
        for (size_t k = 0; k < v->d; k++) {
            for (size_t j = 0; j < v->N[k]; j++) fprintf(fp, "%.21g ", v->xgrid[k][j]);
            fprintf(fp, "\n");
            const size_t len = v->N[k] * v->ranks[k] * v->ranks[k + 1];
            for (size_t e = 0; e < len; e++) fprintf(fp, "%.21g ", v->cores[k][e]);
            fprintf(fp, "\n");
        }
        return ferror(fp) ? 1 : 0;
    }
    uint64_t hd = v->d;
    if (fwrite(VF_MAGIC, 1, 8, fp) != 8 || fwrite(&hd, 8, 1, fp) != 1) return 1;
    for (size_t k = 0; k < v->d; k++) { uint64_t n = v->N[k]; if (fwrite(&n, 8, 1, fp) != 1) return 1; }
    for (size_t k = 0; k <= v->d; k++) { uint64_t r = v->ranks[k]; if (fwrite(&r, 8, 1, fp) != 1) return 1; }
    for (size_t k = 0; k < v->d; k++) {
        const size_t len = v->N[k] * v->ranks[k] * v->ranks[k + 1];
        if (fwrite(v->xgrid[k], 8, v->N[k], fp) != v->N[k] || fwrite(v->cores[k], 8, len, fp) != len) return 1;
    }
    return 0;
}
static struct ValueF *vf_read(FILE *fp, int text, size_t *ngrid, double **xgrid)
{
    size_t d = 0, N[C3SC_MAXD], ranks[C3SC_MAXD + 1];
    double *g[C3SC_MAXD] = { 0 }, *c[C3SC_MAXD] = { 0 }, *cn[C3SC_MAXD] = { 0 };
    struct ValueF *out = NULL;
    int ok = 1;
    if (text) {
        char magic[16];
        ok = fscanf(fp, "%15s %zu", magic, &d) == 2 && !strcmp(magic, "C3SCVF01") && d >= 1 && d <= C3SC_MAXD;
        for (size_t k = 0; ok && k < d; k++) ok = fscanf(fp, "%zu", &N[k]) == 1;
        for (size_t k = 0; ok && k <= d; k++) ok = fscanf(fp, "%zu", &ranks[k]) == 1;
    } else {
        char magic[8];
        uint64_t t;
        ok = fread(magic, 1, 8, fp) == 8 && !memcmp(magic, VF_MAGIC, 8) && fread(&t, 8, 1, fp) == 1 && t >= 1 && t <= C3SC_MAXD;
        d = ok ? (size_t)t : 0;
        for (size_t k = 0; ok && k < d; k++) { ok = fread(&t, 8, 1, fp) == 1; N[k] = (size_t)t; }
        for (size_t k = 0; ok && k <= d; k++) { ok = fread(&t, 8, 1, fp) == 1; ranks[k] = (size_t)t; }
    }
    for (size_t k = 0; ok && k < d; k++) {
        ok = N[k] >= 2 && N[k] < ((size_t)1 << 24) && ranks[k] >= 1 && ranks[k + 1] >= 1 && ranks[k] < 4096 && ranks[k + 1] < 4096;
        if (!ok) break;
        const size_t len = N[k] * ranks[k] * ranks[k + 1];
        g[k] = xalloc(N[k], 8); c[k] = xalloc(len, 8);
        if (text) {
            for (size_t j = 0; ok && j < N[k]; j++) ok = fscanf(fp, "%lf", &g[k][j]) == 1;
            for (size_t e = 0; ok && e < len; e++) ok = fscanf(fp, "%lf", &c[k][e]) == 1;
        } else ok = fread(g[k], 8, N[k], fp) == N[k] && fread(c[k], 8, len, fp) == len;
    }
    if (ok) {
        /* cores on the caller's grid: copied when the nodes agree, else piecewise-linear in the node */
        for (size_t k = 0; k < d; k++) {
            const size_t blk = ranks[k] * ranks[k + 1], Nn = ngrid[k];
            int same = Nn == N[k];
            for (size_t j = 0; same && j < Nn; j++) same = xgrid[k][j] == g[k][j];
            cn[k] = xalloc(Nn * blk, 8);
            if (same) { memcpy(cn[k], c[k], Nn * blk * 8); continue; }
            for (size_t j = 0; j < Nn; j++) {
                const double x = xgrid[k][j];
                if (x < g[k][0] || x > g[k][N[k] - 1]) continue;               /* zero outside the saved nodes */
                size_t i = 0;
                while (i + 2 < N[k] && x > g[k][i + 1]) i++;
                const double w = (x - g[k][i]) / (g[k][i + 1] - g[k][i]);
                for (size_t e = 0; e < blk; e++) cn[k][j * blk + e] = (1.0 - w) * c[k][i * blk + e] + w * c[k][(i + 1) * blk + e];
            }
        }
        out = valuef_from_cores(d, ngrid, ranks, cn);
        valuef_set_grid(out, xgrid);
    }
    for (size_t k = 0; k < C3SC_MAXD; k++) { free(g[k]); free(c[k]); free(cn[k]); }
    return out;
}
int valuef_save(struct ValueF *v, char *filename)
{
    FILE *fp = fopen(filename, "wb");
    if (!fp) return 1;
    const int rc = vf_write(v, fp, 0);
    return fclose(fp) || rc;
}
int valuef_savetxt(struct ValueF *v, char *filename)
{
    FILE *fp = fopen(filename, "w");
    if (!fp) return 1;
    const int rc = vf_write(v, fp, 1);
    return fclose(fp) || rc;
}
struct ValueF *valuef_load(char *filename, size_t *ngrid, double **xgrid)
{
    FILE *fp = fopen(filename, "rb");
    if (!fp) return NULL;                                   /* the examples test for NULL to start afresh */
    struct ValueF *v = vf_read(fp, 0, ngrid, xgrid);
    fclose(fp);
    return v;
}
struct ValueF *valuef_loadtxt(char *filename, size_t *ngrid, double **xgrid)
{
    FILE *fp = fopen(filename, "r");
    if (!fp) return NULL;
    struct ValueF *v = vf_read(fp, 1, ngrid, xgrid);
    fclose(fp);
    return v;
}

/* ========================== solver loops (src/bellman.c:2105-2420) ========== */
void c3control_add_policy_sim(struct C3Control *c, struct ValueF *pol, struct c3Opt *opt_sim,
                              void (*transform)(size_t, const double *, double *))
{
    c->policy_sim = pol; c->opt_sim = opt_sim; c->transform_sim = transform;
    control_params_destroy(c->cp_sim); c->cp_sim = NULL;
}
int c3control_policy_eval(struct C3Control *c, double t, const double *x, double *u)
{
    if (!c->policy_sim || !c->opt_sim) die("c3control_policy_eval: no policy (c3control_add_policy_sim)");
    if (!c->cp_sim) c->cp_sim = control_params_create(c->dx, c->dw, c->dp, c->mca, c->work, c->opt_sim);
    if (!c->prevpol) c->prevpol = xalloc(c->du, sizeof(double));
    control_params_add_time_and_states(c->cp_sim, t, 1, x);
    int ab = 0;
    int rc = c3sc_policy_eval_batch(cp_dev(c->cp_sim), c->policy_sim->dev, 1, x, u, NULL, &ab,
                                    workspace_get_costs(c->work, 0));
    if (rc) { fprintf(stderr, "c3sc_b200: c3control_policy_eval: %s\n", c3sc_last_error()); return rc; }
    workspace_get_absorbed(c->work, 0)[0] = ab;
    memcpy(c->prevpol, u, c->du * sizeof(double));
    return 0;
}
int c3control_controller(double t, const double *x, double *u, void *args)
{
    struct C3Control *c = args;
    if (!c->transform_sim) return c3control_policy_eval(c, t, x, u);
    double *xin = xalloc(c->dx, sizeof(double));
    c->transform_sim(c->dx, x, xin);
    int rc = c3control_policy_eval(c, t, xin, u);
    free(xin);
    return rc;
}

struct ValueF *c3control_step_vi(struct C3Control *c, struct ValueF *vf, struct ApproxArgs *apargs, struct c3Opt *opt,
                                 int verbose, size_t *nevals)
{
    struct ControlParams *cp = control_params_create(c->dx, c->dw, c->dp, c->mca, c->work, opt);
    struct VIparam *vi = vi_param_create(1e-10);
    vi_param_add_cp(vi, cp);
    vi_param_add_value(vi, vf);
    workspace_increment_vi_iter(c->work);
    struct ValueF *next = valuef_interp(c->dx, bellman_vi, vi, c->ngrid, c->xgrid, vf, apargs, verbose);
    if (nevals) *nevals = vi->nnode_evals;
    vi_param_destroy(vi);
    control_params_destroy(cp);
    return next;
}
struct ValueF *c3control_step_pi(struct C3Control *c, struct ValueF *vf, struct PIparam *poli, struct ApproxArgs *apargs,
                                 struct c3Opt *opt, int verbose, size_t *niter_evals)
{
    struct ControlParams *cp = control_params_create(c->dx, c->dw, c->dp, c->mca, c->work, opt);
    pi_param_add_cp(poli, cp);
    pi_param_add_value(poli, vf);
    workspace_increment_pi_subiter(c->work);
    struct ValueF *next = valuef_interp(c->dx, bellman_pi, poli, c->ngrid, c->xgrid, vf, apargs, verbose);
    if (niter_evals) *niter_evals = poli->niter_evals;
    pi_param_add_cp(poli, NULL);
    control_params_destroy(cp);
    return next;
}
struct ValueF *c3control_init_value(struct C3Control *c, int (*f)(size_t, const double *, double *, void *), void *args,
                                    struct ApproxArgs *aargs, int verbose)
{
    return valuef_interp(c->dx, f, args, c->ngrid, c->xgrid, NULL, aargs, verbose);
}

/* ---- Diag: the per-iteration record list (src/bellman.c:2408-2540) ---- */
struct Diag { size_t iter; int type; double norm, abs_diff; size_t dim, *ranks; double frac; struct Diag *next; };
struct Diag *diag_create(size_t iter, int type, double norm, double abs_diff, size_t dim, size_t *ranks, double frac)
{
    struct Diag *g = xalloc(1, sizeof *g);
    g->iter = iter; g->type = type; g->norm = norm; g->abs_diff = abs_diff; g->dim = dim; g->frac = frac;
    g->ranks = xalloc(dim + 1, sizeof(size_t));
    memcpy(g->ranks, ranks, (dim + 1) * sizeof(size_t));
    return g;
}
void diag_destroy(struct Diag **head)
{
    if (!head) return;
    for (struct Diag *g = *head; g;) { struct Diag *nx = g->next; free(g->ranks); free(g); g = nx; }
    *head = NULL;
}
void diag_append(struct Diag **head, size_t iter, int type, double norm, double abs_diff, size_t dim, size_t *ranks, double frac)
{
    struct Diag *g = diag_create(iter, type, norm, abs_diff, dim, ranks, frac);
    if (!*head) { *head = g; return; }
    struct Diag *t = *head;
    while (t->next) t = t->next;
    t->next = g;
}
void diag_print(struct Diag *head, FILE *fp)
{
    fprintf(fp, "Type Iter Norm AbsDiff RelDiff FracEval AvgRank\n");
    for (struct Diag *g = head; g; g = g->next) {
        double avg = 0.0;
        for (size_t k = 1; k < g->dim; k++) avg += (double)g->ranks[k];
        if (g->dim > 1) avg /= (double)(g->dim - 1);
        fprintf(fp, "%d %zu %3.5E %3.5E %3.5E %3.5E %3.5E\n", g->type, g->iter, g->norm, g->abs_diff,
                g->norm != 0.0 ? g->abs_diff / g->norm : 0.0, g->frac, avg);
    }
}
int diag_save(struct Diag *head, char *filename)
{
    FILE *fp = fopen(filename, "w");
    if (!fp) return 1;
    diag_print(head, fp);
    fclose(fp);
    return 0;
}

static void solve_report(const char *what, size_t ii, size_t maxiter, double diff, double norm, double frac)
{
    printf("\t %s (%zu\\%zu):\n", what, ii + 1, maxiter);
    printf("\t \t L2 Difference between iterates    = %3.5E\n ", diff);
    printf("\t \t L2 Norm of current value function = %3.5E\n", norm);
    printf("\t \t Relative L2 Cauchy difference     = %3.5E\n", diff / norm);
    printf("\t \t Fraction of states evaluated      = %3.5E\n", frac);
}
struct ValueF *c3control_vi_solve(struct C3Control *c, size_t maxiter, double abs_conv_tol, struct ValueF *vo,
                                  struct ApproxArgs *apargs, struct c3Opt *opt, int verbose, struct Diag **diag)
{
    struct ValueF *start = valuef_copy(vo);
    if (!start->xgrid) valuef_set_grid(start, c->xgrid);   /* a value function built from bare cores lives on the solver's grid */
    workspace_reset_vi_htable(c->work);
    double stot = 1.0;
    for (size_t j = 0; j < c->dx; j++) stot *= (double)c->ngrid[j];
    for (size_t ii = 0; ii < maxiter; ii++) {
        size_t nev = 0;
        struct ValueF *next = c3control_step_vi(c, start, apargs, opt, verbose - 1, &nev);
        const double diff = valuef_norm2diff(start, next), norm = valuef_norm(next), frac = (double)nev / stot;
        if (verbose > 0) solve_report("Value Iteration", ii, maxiter, diff, norm, frac);
        if (diag) diag_append(diag, ii, 1, norm, diff, c->dx, valuef_get_ranks(next), frac);
        valuef_destroy(start);
        start = next;
        if (diff < abs_conv_tol) break;
    }
    return start;
}
struct ValueF *c3control_pi_solve(struct C3Control *c, size_t maxiter, double abs_conv_tol, struct ValueF *policy,
                                  struct ApproxArgs *apargs, struct c3Opt *opt, int verbose, struct Diag **diag)
{
    struct ValueF *start = valuef_copy(policy);
    if (!start->xgrid) valuef_set_grid(start, c->xgrid);
    struct PIparam *poli = pi_param_create(1e-10, policy);
    workspace_increment_pi_iter(c->work);
    workspace_reset_pi_prob_htable(c->work);
    workspace_reset_pi_htable(c->work);
    double stot = 1.0;
    for (size_t j = 0; j < c->dx; j++) stot *= (double)c->ngrid[j];
    for (size_t ii = 0; ii < maxiter; ii++) {
        size_t nev = 0;
        struct ValueF *next = c3control_step_pi(c, start, poli, apargs, opt, verbose - 1, &nev);
        const double diff = valuef_norm2diff(start, next), norm = valuef_norm(next), frac = (double)nev / stot;
        if (verbose > 0) solve_report("POLICY ITERATION", ii, maxiter, diff, norm, frac);
        if (diag) diag_append(diag, ii, 0, norm, diff, c->dx, valuef_get_ranks(next), frac);
        valuef_destroy(start);
        start = next;
        if (diff < abs_conv_tol) break;
    }
    pi_param_destroy(poli);
    return start;
}
