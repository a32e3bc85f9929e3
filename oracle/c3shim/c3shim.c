/* C3 stand-in implementation (test infrastructure only; see c3/array.h).
 * Plain-loop CBLAS, array helpers and the brute-force subset of c3opt, just
 * enough for /root/reference/src/{bellman,nodeutil,boundary,dynamics,
 * hashgrid,util}.c to link unmodified into oracle/_ref/libc3sc_ref.so.    */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include "c3/array.h"
#include "c3/lib_linalg.h"
#include "c3/lib_optimization.h"
#include "c3/stringmanip.h"

static void *xcalloc(size_t n, size_t sz)
{
    void *p = calloc(n ? n : 1, sz);
    if (!p) { fprintf(stderr, "c3shim: out of memory\n"); exit(1); }
    return p;
}
double  *calloc_double(size_t n) { return xcalloc(n, sizeof(double)); }
size_t  *calloc_size_t(size_t n) { return xcalloc(n, sizeof(size_t)); }
int     *calloc_int(size_t n)    { return xcalloc(n, sizeof(int)); }
double **malloc_dd(size_t n)     { return xcalloc(n, sizeof(double *)); }
void free_dd(size_t n, double **a)
{
    if (!a) return;
    for (size_t i = 0; i < n; i++) free(a[i]);
    free(a);
}
/* C3 array.c: first node = lb, then running sum of the interval. */
double *linspace(double lb, double ub, size_t n)
{
    if (n == 0) return NULL;
    double *g = calloc_double(n);
    g[0] = lb;
    if (n > 1) {
        double step = (ub - lb) / (double)(n - 1);
        for (size_t i = 1; i < n; i++) g[i] = g[i - 1] + step;
    }
    return g;
}
double randu(void) { return (double)rand() / (double)RAND_MAX; }
void dprint(size_t n, const double *a) { for (size_t i = 0; i < n; i++) printf("%3.15G ", a[i]); printf("\n"); }
void iprint(size_t n, const int *a)    { for (size_t i = 0; i < n; i++) printf("%d ", a[i]); printf("\n"); }
void iprint_sz(size_t n, const size_t *a) { for (size_t i = 0; i < n; i++) printf("%zu ", a[i]); printf("\n"); }

/* ---- CBLAS, strictly sequential accumulation ------------------------- */
double cblas_ddot(int n, const double *x, int incx, const double *y, int incy)
{
    double s = 0.0;
    for (int i = 0; i < n; i++) s += x[(size_t)i * incx] * y[(size_t)i * incy];
    return s;
}
void cblas_daxpy(int n, double a, const double *x, int incx, double *y, int incy)
{
    for (int i = 0; i < n; i++) y[(size_t)i * incy] += a * x[(size_t)i * incx];
}
void cblas_dgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int m, int n,
                 double alpha, const double *a, int lda, const double *x, int incx,
                 double beta, double *y, int incy)
{
    if (order != CblasColMajor) { fprintf(stderr, "c3shim: row-major dgemv unsupported\n"); exit(1); }
    int ny = (trans == CblasNoTrans) ? m : n, nx = (trans == CblasNoTrans) ? n : m;
    for (int i = 0; i < ny; i++) {
        double s = 0.0;
        for (int k = 0; k < nx; k++) {
            double aik = (trans == CblasNoTrans) ? a[i + (size_t)k * lda] : a[k + (size_t)i * lda];
            s += aik * x[(size_t)k * incx];
        }
        double *yi = y + (size_t)i * incy;
        *yi = (beta == 0.0) ? alpha * s : alpha * s + beta * (*yi);
    }
}

/* ---- c3opt, brute force only ------------------------------------------ */
struct c3Opt {
    enum c3opt_alg alg;
    size_t d, nvals;
    double *vals, *lb, *ub;
    double (*f)(size_t, const double *, double *, void *);
    void *farg;
    int verbose;
};
struct c3Opt *c3opt_alloc(enum c3opt_alg alg, size_t d)
{
    struct c3Opt *o = xcalloc(1, sizeof *o);
    o->alg = alg; o->d = d;
    return o;
}
static double *dupd(const double *s, size_t n)
{
    if (!s) return NULL;
    double *p = calloc_double(n);
    memcpy(p, s, n * sizeof(double));
    return p;
}
struct c3Opt *c3opt_copy(struct c3Opt *o)
{
    if (!o) return NULL;
    struct c3Opt *c = xcalloc(1, sizeof *c);
    *c = *o;
    c->vals = dupd(o->vals, o->nvals * o->d);
    c->lb = dupd(o->lb, o->d);
    c->ub = dupd(o->ub, o->d);
    return c;
}
void c3opt_free(struct c3Opt *o) { if (o) { free(o->vals); free(o->lb); free(o->ub); free(o); } }
void c3opt_add_objective(struct c3Opt *o, double (*f)(size_t, const double *, double *, void *), void *arg) { o->f = f; o->farg = arg; }
int  c3opt_is_bruteforce(const struct c3Opt *o) { return o->alg == BRUTEFORCE; }
void c3opt_set_brute_force_vals(struct c3Opt *o, size_t n, double *vals)
{
    free(o->vals);
    o->nvals = n;
    o->vals = dupd(vals, n * o->d);
}
int c3opt_minimize(struct c3Opt *o, double *x, double *val)
{
    if (o->alg != BRUTEFORCE || !o->nvals) { fprintf(stderr, "c3shim: only BRUTEFORCE c3opt is available\n"); exit(1); }
    size_t best = 0;
    double fbest = o->f(o->d, o->vals, NULL, o->farg);
    for (size_t i = 1; i < o->nvals; i++) {
        double fi = o->f(o->d, o->vals + i * o->d, NULL, o->farg);
        if (fi < fbest) { fbest = fi; best = i; }
    }
    memcpy(x, o->vals + best * o->d, o->d * sizeof(double));
    *val = fbest;
    return 0;
}
double *c3opt_get_lb(struct c3Opt *o) { return o->lb; }
double *c3opt_get_ub(struct c3Opt *o) { return o->ub; }
void c3opt_add_lb(struct c3Opt *o, double *lb) { free(o->lb); o->lb = dupd(lb, o->d); }
void c3opt_add_ub(struct c3Opt *o, double *ub) { free(o->ub); o->ub = dupd(ub, o->d); }
void c3opt_set_verbose(struct c3Opt *o, int v) { o->verbose = v; }
size_t c3opt_get_d(const struct c3Opt *o) { return o->d; }

char *serialize_double_to_text(double v)
{
    char *s = malloc(64);
    snprintf(s, 64, "%3.15E", v);
    return s;
}
double deserialize_double_from_text(char *s) { return strtod(s, NULL); }
