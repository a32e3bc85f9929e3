/* c3sc_cross.h -- host cross-approximation driver with BATCHED fiber requests.
 *
 * Replaces, for the Bellman path, what the reference obtains from the (un-vendored) C3 library in
 * valuef_interp (src/valuefunc.c:603-767): ftapprox_cross on a nodal (LINELM) discretisation,
 * maxiter 5, index sets started from uniform_stride (src/util.c:995-1006).  The operator handed to
 * the driver is called once per CORE with all r_k*r_{k+1} fibers of that core, instead of once per
 * fiber (bellman_vi's signature, src/bellman.h:96); the result is in ValueF::cores layout
 * (src/valuefunc.c:165-189) and goes straight into c3sc_valuef_update.
 */
#ifndef C3SC_CROSS_H
#define C3SC_CROSS_H
#include "c3sc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* F fibers (dim_vary[F], fixed_ind[F*d]) -> out[F*ldo]; 0 on success.  Same contract as c3sc_vi_batch. */
typedef int (*c3sc_fiber_batch_fn)(size_t F, const int32_t *dim_vary, const int32_t *fixed_ind, size_t ldo,
                                   double *out, void *arg);

typedef struct c3sc_cross_opts {
    uint32_t maxiter;     /* sweep pairs (left->right + right->left); 0 = the reference's 5 */
    double tol;           /* stop when the relative change of the train drops below it; 0 = run maxiter */
    int verbose;
} c3sc_cross_opts;

typedef struct c3sc_cross c3sc_cross;   /* ranks + left/right index sets, kept between value-iteration steps */

/* ranks[d+1] with ranks[0] = ranks[d] = 1 (clipped to what the unfoldings allow) */
int  c3sc_cross_create(uint32_t d, const uint64_t *n, const uint64_t *ranks, c3sc_cross **out);
int  c3sc_cross_copy(const c3sc_cross *src, c3sc_cross **out);    /* ranks + index sets */
void c3sc_cross_destroy(c3sc_cross *c);
/* on != 0: the buffers the driver exchanges with the operator (fiber descriptors out, fiber values back) are
 * page-locked (c3sc_host_alloc), so a GPU operator copies at PCIe speed; c3sc_cross_run_vi / _pi and their
 * multi-GPU forms turn it on themselves.  Pageable memory is used where page-locking fails. */
int  c3sc_cross_pin_buffers(c3sc_cross *c, int on);
int  c3sc_cross_ranks(const c3sc_cross *c, uint64_t *ranks);
uint32_t c3sc_cross_dim(const c3sc_cross *c);
/* index sets at bond k (0..d): left[r_k*d] over dims 0..k-1, right[r_k*d] over dims k..d-1 (others 0);
 * what ValueF keeps as isl / isr between solver steps (src/valuefunc.c:706-712).  Either may be NULL. */
int  c3sc_cross_index_sets(const c3sc_cross *c, uint32_t k, int32_t *left, int32_t *right);

/* cores[k]: caller-allocated n[k]*r[k]*r[k+1] doubles.  nfibers / rel_change may be NULL. */
int c3sc_cross_run(c3sc_cross *c, c3sc_fiber_batch_fn f, void *arg, const c3sc_cross_opts *opts,
                   double *const *cores, uint64_t *nfibers, double *rel_change);

/* ---- fiber memo ---------------------------------------------------------------------------------------
 * The reference memoises backed-up values across the <= 5 sweeps of one cross approximation (hash tables keyed by
 * node, src/bellman.c:1334-1349, 1383).  Here a backup is a pure function of the fiber, so the
 * memo works on whole fibers: a wrapper around any fiber operator that computes each distinct
 * (dim_vary, fixed indices) once -- repeated requests, inside a batch or in a later sweep, are served from the
 * stored values (bit-identical: the same numbers).  c3sc_cross_run_vi / _pi and their multi-GPU forms use one for
 * the duration of a call.  The operator behind a memo must not change while the memo holds values. */
typedef struct c3sc_fiber_memo c3sc_fiber_memo;
int  c3sc_fiber_memo_create(uint32_t d, c3sc_fiber_batch_fn f, void *arg, c3sc_fiber_memo **out);
/* a c3sc_fiber_batch_fn: pass it with arg = the memo */
int  c3sc_fiber_memo_call(size_t F, const int32_t *dim_vary, const int32_t *fixed_ind, size_t ldo, double *out, void *memo);
void c3sc_fiber_memo_stats(const c3sc_fiber_memo *m, uint64_t *requested, uint64_t *computed);
void c3sc_fiber_memo_clear(c3sc_fiber_memo *m);
void c3sc_fiber_memo_destroy(c3sc_fiber_memo *m);
/* whether c3sc_cross_run_vi / _pi put a memo in front of the GPU operator for these options (more than one sweep pair) */
int  c3sc_cross_uses_memo(const c3sc_cross_opts *opts);

/* c3control_step_vi / c3control_step_pi (src/bellman.c:2177-2262) on the GPU path */
int c3sc_cross_run_vi(c3sc_cross *c, c3sc_problem *p, const c3sc_valuef *vf, const c3sc_cross_opts *opts,
                      double *const *cores, uint64_t *nfibers, double *rel_change);
int c3sc_cross_run_pi(c3sc_cross *c, c3sc_problem *p, const c3sc_valuef *vf_policy, const c3sc_valuef *vf_iter,
                      uint32_t dx, const c3sc_cross_opts *opts, double *const *cores, uint64_t *nfibers,
                      double *rel_change);

/* ---- rank adaptation (the adapt == 1 branch of valuef_interp, src/valuefunc.c:637-648,:706-730:
 * C3's ftapprox_cross_rankadapt with round_tol / kickrank / maxrank) ------------------------------- */
typedef struct c3sc_adapt_opts {
    uint32_t kickrank;       /* ranks the rounding did not reduce grow by this much; 0 = round only */
    uint32_t maxrank;        /* cap; 0 or > min N = min N (src/valuefunc.c:625-631) */
    double round_tol;        /* relative accuracy of the rounding */
    uint32_t maxiter_adapt;  /* cross runs at most; 0 = 5 */
} c3sc_adapt_opts;

/* function_train_round on nodal cores (discrete l2): ranks_out[d+1], cores_out[k] with the capacity of
 * cores_in[k].  ranks_in must respect the unfolding bounds (r_k <= r_{k-1} n_{k-1}, r_k <= r_{k+1} n_k). */
int c3sc_cores_round(uint32_t d, const uint64_t *n, const uint64_t *ranks_in, const double *const *cores_in,
                     double eps, uint64_t *ranks_out, double *const *cores_out);

/* largest rank per bond an adaptive run can return (sizes the caller's cores: n[k]*cap[k]*cap[k+1]) */
int c3sc_cross_adapt_capacity(const c3sc_cross *c, const c3sc_adapt_opts *aopts, uint64_t *cap);
/* resize the driver's ranks / index sets, e.g. to min(found+1, maxrank) before the next solver step */
int c3sc_cross_set_ranks(c3sc_cross *c, const uint64_t *ranks);
/* cross -> round -> kick -> cross ...; returns the ROUNDED train (ranks_out, cores) */
int c3sc_cross_run_adapt(c3sc_cross *c, c3sc_fiber_batch_fn f, void *arg, const c3sc_cross_opts *opts,
                         const c3sc_adapt_opts *aopts, uint64_t *ranks_out, double *const *cores, uint64_t *nfibers,
                         double *rel_change);
int c3sc_cross_run_vi_adapt(c3sc_cross *c, c3sc_problem *p, const c3sc_valuef *vf, const c3sc_cross_opts *opts,
                            const c3sc_adapt_opts *aopts, uint64_t *ranks_out, double *const *cores, uint64_t *nfibers,
                            double *rel_change);

/* c3control_vi_solve (src/bellman.c:2282-2340) on the GPU path: value iteration from the start train
 * (ranks0, cores0) until ||V_{t+1} - V_t|| < abs_conv_tol (nodal l2) or maxiter steps.  cores_out in the
 * driver's ranks (c3sc_cross_ranks).  iters_done / last_diff / nfibers may be NULL. */
int c3sc_vi_solve(c3sc_cross *c, c3sc_problem *p, const uint64_t *ranks0, const double *const *cores0,
                  uint32_t maxiter, double abs_conv_tol, const c3sc_cross_opts *opts, double *const *cores_out,
                  uint32_t *iters_done, double *last_diff, uint64_t *nfibers);

/* valuef_norm / valuef_norm2diff (src/valuefunc.c:315-335) on nodal cores in ValueF::cores layout:
 * discrete l2 over the grid nodes; the two trains may have different ranks. */
double c3sc_cores_dot(uint32_t d, const uint64_t *n, const uint64_t *ranks_a, const double *const *a,
                      const uint64_t *ranks_b, const double *const *b);
double c3sc_cores_norm(uint32_t d, const uint64_t *n, const uint64_t *ranks, const double *const *a);
double c3sc_cores_norm2diff(uint32_t d, const uint64_t *n, const uint64_t *ranks_a, const double *const *a,
                            const uint64_t *ranks_b, const double *const *b);

/* The reference's norms (src/valuefunc.c:315-335: function_train_norm2 / norm2diff on LINELM cores): the CONTINUOUS L2
 * inner product over the box of the multilinear interpolants of the nodal cores -- the train contraction with the
 * tridiagonal mass matrix of the hat functions of xgrid[k] in every dimension.  This is what abs_conv_tol of
 * c3control_vi_solve / pi_solve is measured in (src/bellman.c:2307-2337,2367); the nodal functions above are the
 * discrete l2 of the node values and are not the reference's quantity. */
double c3sc_cores_dot_l2(uint32_t d, const uint64_t *n, const double *const *xgrid, const uint64_t *ranks_a,
                         const double *const *a, const uint64_t *ranks_b, const double *const *b);
double c3sc_cores_norm_l2(uint32_t d, const uint64_t *n, const double *const *xgrid, const uint64_t *ranks,
                          const double *const *a);
double c3sc_cores_norm2diff_l2(uint32_t d, const uint64_t *n, const double *const *xgrid, const uint64_t *ranks_a,
                               const double *const *a, const uint64_t *ranks_b, const double *const *b);

#ifdef __cplusplus
}
#endif
#endif /* C3SC_CROSS_H */
