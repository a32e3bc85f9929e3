/* C3 stand-in: brute-force subset of c3opt (the only branch on the hot path,
 * bellman.c:539-543).  Candidate order = table order, strict '<' update so
 * the FIRST minimum wins -- this tie rule is pinned by no reference test
 * ("parity unpinned" for ties). */
#ifndef C3SHIM_OPT_H
#define C3SHIM_OPT_H
#include <stddef.h>
enum c3opt_alg { BFGS, LBFGS, BATCHGRAD, BRUTEFORCE, SGD };
struct c3Opt;
struct c3Opt *c3opt_alloc(enum c3opt_alg alg, size_t d);
struct c3Opt *c3opt_copy(struct c3Opt *o);
void   c3opt_free(struct c3Opt *o);
void   c3opt_add_objective(struct c3Opt *o, double (*f)(size_t, const double *, double *, void *), void *arg);
int    c3opt_is_bruteforce(const struct c3Opt *o);
void   c3opt_set_brute_force_vals(struct c3Opt *o, size_t n, double *vals);
int    c3opt_minimize(struct c3Opt *o, double *x, double *val);
double *c3opt_get_lb(struct c3Opt *o);
double *c3opt_get_ub(struct c3Opt *o);
void   c3opt_add_lb(struct c3Opt *o, double *lb);
void   c3opt_add_ub(struct c3Opt *o, double *ub);
void   c3opt_set_verbose(struct c3Opt *o, int v);
size_t c3opt_get_d(const struct c3Opt *o);
#endif
