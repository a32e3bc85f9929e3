/* c3sc_host.h -- C host mirror of the reference's API for the Bellman-backup path.
 *
 * Same names, argument meaning and return conventions as the reference headers
 * (src/bellman.h, src/nodeutil.h, src/valuefunc.h, src/boundary.h,
 * src/dynamics.h, src/util.h under /root/reference), so a program written
 * against c3sc compiles against this header for everything on the path.  All
 * ARITHMETIC of the path runs on the GPU through include/c3sc_b200.h; the host
 * side holds containers, decodes fibers (x -> indices) and moves buffers.
 * There is no CPU fallback: compute entries return non-zero / abort loudly if
 * no CUDA device is usable.
 *
 * Additions (marked NEW) are the minimum the GPU path needs that the
 * reference API cannot express: a device model id next to the host callbacks,
 * batched fiber entries, and ValueF construction from nodal cores (the
 * reference builds ValueF only through the absent C3 library).
 *
 * The solver loops (valuef_interp, c3control_init_value / step_vi / step_pi / vi_solve / pi_solve, ApproxArgs,
 * Diag) and the online controller (c3control_add_policy_sim / policy_eval / controller) are here too; what they
 * obtain from C3's cross approximation in the reference is restated in include/c3sc_cross.h (batched core
 * requests, TT rounding, rank adaptation, nodal norms).  examples/lqg2d_b200.c is an examples/lqg2d_new-style
 * main() written against this header.
 *
 * NOT mirrored (out of scope): C3's own .c3 file formats (valuef_save / load write this
 * library's format), the CONSTELM function class, BoundInfo, HashGrid, process_fibers, the BFGS branch of bellman_optimal and every gradient output
 * (grad_* arguments must be NULL), the trajectory simulation of the example tails (cdyn).
 */
#ifndef C3SC_HOST_H
#define C3SC_HOST_H
#include <stddef.h>
#include <stdio.h>
#include "c3sc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- boundary.h ------------------------------------------------------------ */
enum EBTYPE { EB_NONE = 0, ABSORB = 1, PERIODIC = 2, REFLECT = 3 };   /* src/boundary.h:42-47 */
struct Boundary;
struct Boundary *boundary_alloc(size_t, double *, double *);            /* src/boundary.c:375 */
struct Boundary *boundary_copy_deep(struct Boundary *);                 /* :402 */
void boundary_free(struct Boundary *);                                   /* :429 */
size_t boundary_get_nobs(struct Boundary *);                             /* :443 */
double *boundary_obstacle_get_lb(struct Boundary *, size_t);             /* :452 */
double *boundary_obstacle_get_ub(struct Boundary *, size_t);             /* :461 */
void boundary_add_obstacle(struct Boundary *, double *, double *);       /* :470 */
void boundary_external_set_type(struct Boundary *, size_t, char *);      /* :486 */
enum EBTYPE boundary_type_dim(const struct Boundary *, size_t, int);     /* :604 */
int boundary_in_obstacle(const struct Boundary *, const double *);       /* :668 */
/* x mapped through a PERIODIC face: the opposite bound and *map = 1 (x <= left) / 2 (x >= right); else x, 0 */
double outer_bound_dim(const struct Boundary *, size_t, double, int *);  /* :577 */

/* ---- dynamics.h ------------------------------------------------------------ */
struct Drift;
struct Diff;
struct Dyn;
typedef int (*c3sc_dyn_cb)(double, const double *, const double *, double *, double *, void *);
struct Drift *drift_alloc(size_t, size_t);                               /* src/dynamics.c:70 */
struct Drift *drift_copy(struct Drift *);
void drift_free(struct Drift *);
void drift_add_func(struct Drift *, c3sc_dyn_cb, void *);                /* :107 */
size_t drift_get_dx(struct Drift *);
int drift_eval(struct Drift *, double, const double *, const double *, double *, double *);   /* :127 */
struct Diff *diff_alloc(size_t, size_t, size_t);                         /* :176 */
struct Diff *diff_copy(struct Diff *);
void diff_free(struct Diff *);
void diff_add_func(struct Diff *, c3sc_dyn_cb, void *);
int diff_eval(struct Diff *, double, const double *, const double *, double *, double *);     /* :224 */
size_t diff_get_dw(struct Diff *);
struct Dyn *dyn_alloc(struct Drift *, struct Diff *);                    /* :265 */
void dyn_free(struct Dyn *);
void dyn_free_deep(struct Dyn *);
struct Dyn *dyn_copy_deep(struct Dyn *);                                 /* :279 */
void dyn_init_ref(struct Dyn *, struct Drift *, struct Diff *);          /* :307 */
size_t dyn_get_dx(struct Dyn *);
size_t dyn_get_dw(struct Dyn *);
size_t dyn_get_du(struct Dyn *);
int dyn_eval(struct Dyn *, double, const double *, const double *, double *, double *, double *, double *);

/* ---- util.h --------------------------------------------------------------- */
/* largest stride s with s*(M-1) <= N-1 reached by the reference's search: the start index sets of the
 * cross approximation (src/util.c:995-1006).  M < 2 returns 0 (the reference does not terminate there). */
size_t uniform_stride(size_t, size_t);
/* ---- util.h: Workspace (iteration counters; the memo tables are dropped) ------ */
struct Workspace;
struct Workspace *workspace_alloc(size_t, size_t, size_t, size_t);       /* src/util.c:717 */
void workspace_free(struct Workspace *);
void workspace_reset_pi_prob_htable(struct Workspace *);                 /* frees the resident policy rows */
void workspace_reset_pi_htable(struct Workspace *);
void workspace_reset_vi_htable(struct Workspace *);
void workspace_increment_vi_iter(struct Workspace *);
size_t workspace_get_vi_iter(const struct Workspace *);
void workspace_increment_pi_iter(struct Workspace *);
size_t workspace_get_pi_iter(const struct Workspace *);
void workspace_increment_pi_subiter(struct Workspace *);
size_t workspace_get_pi_subiter(const struct Workspace *);
double *workspace_get_costs(struct Workspace *, size_t);
int *workspace_get_absorbed(struct Workspace *, size_t);
double *workspace_get_u(struct Workspace *, size_t);
/* per-node scratch slabs in the reference's layout (src/util.c:738-748,843-906).  The batched kernels keep these
 * quantities in registers; the slabs are filled by the scalar entries bellman_control (drift, diff, dt, prob of the
 * (node, u) it was called with) and bellman_optimal (u), which is how the reference's callers read them
 * (src/bellman.c:400-449).  The gradient slabs exist for layout compatibility and stay zero (BFGS path out of scope). */
void workspace_set_active(struct Workspace *, size_t);
size_t workspace_get_active(struct Workspace *);
double *workspace_get_drift(struct Workspace *, size_t);
double *workspace_get_grad_drift(struct Workspace *, size_t);
double *workspace_get_diff(struct Workspace *, size_t);
double *workspace_get_grad_diff(struct Workspace *, size_t);
double *workspace_get_dt(struct Workspace *, size_t);
double *workspace_get_grad_dt(struct Workspace *, size_t);
double *workspace_get_prob(struct Workspace *, size_t);
double *workspace_get_grad_prob(struct Workspace *, size_t);
double *workspace_get_grad_stage(struct Workspace *, size_t);
double *workspace_get_control_size_extra(struct Workspace *, size_t);

/* ---- minimal c3opt (brute force only; C3's lib_optimization.h is absent) ------- */
enum c3opt_alg { BFGS, LBFGS, BATCHGRAD, BRUTEFORCE, SGD };
struct c3Opt;
struct c3Opt *c3opt_alloc(enum c3opt_alg, size_t);
struct c3Opt *c3opt_copy(struct c3Opt *);
void c3opt_free(struct c3Opt *);
int c3opt_is_bruteforce(const struct c3Opt *);
void c3opt_set_brute_force_vals(struct c3Opt *, size_t, double *);       /* n x du, candidate-major */
size_t c3opt_get_d(const struct c3Opt *);

/* ---- valuefunc.h --------------------------------------------------------------- */
struct ValueF;
/* NEW: ValueF from nodal cores in the layout of valuef_precompute_cores
 * (src/valuefunc.c:165-189); uploads them to the device.                      */
struct ValueF *valuef_from_cores(size_t d, const size_t *N, const size_t *ranks, double *const *cores);
/* NEW: same shapes, new numbers (next iterate) */
int valuef_update_cores(struct ValueF *, double *const *cores);
void valuef_destroy(struct ValueF *);                                     /* src/valuefunc.c:104 */
struct ValueF *valuef_copy(struct ValueF *);                              /* :194 */
size_t *valuef_get_ranks(struct ValueF *);                                /* :300 */
/* Stand-in for C3's struct CrossIndex (absent): n multi-indices over the d leading dimensions, as grid indices
 * (inds[a*d + i]) and, when the value function knows its grid, node coordinates (vals, what C3 stores). */
struct CrossIndex { size_t d, n; size_t *inds; double *vals; };
struct CrossIndex **valuef_get_isl(const struct ValueF *);               /* :218; owned by the value function */
int valuef_eval_fiber_ind_nn(struct ValueF *, const size_t *, size_t, const size_t *, const size_t *,
                             double *);                                   /* :369 */

/* ---- nodeutil.h ---------------------------------------------------------------- */
int transition_assemble(size_t dx, size_t du, size_t dw, double h, const double *hvec,
                        const double *drift, const double *grad_drift, const double *ddiff,
                        const double *grad_ddiff, double *prob, double *grad_prob, double *dt,
                        double *grad_dt, double *space);                  /* src/nodeutil.c:267 */
int convert_fiber_to_ind(size_t d, size_t N, const double *x, const size_t *Ngrid, double **xgrid,
                         size_t *fixed_ind, size_t *dim_vary);            /* :437 (host: index decode) */
int process_fibers_neighbor(size_t d, const size_t *fixed_ind, size_t dim_vary, const double *x, int *absorbed,
                            size_t *neighbors_vary, size_t *neighbors_fixed, const size_t *ngrid,
                            const struct Boundary *bound);                /* :489 */
int mca_get_neighbor_node_costs(size_t d, const double *x, struct Boundary *bound, struct ValueF *vf,
                                const size_t *ngrid, double **xgrid, int *absorbed, double *out);   /* :718 */

/* ---- bellman.h ----------------------------------------------------------------- */
double bellmanrhs(size_t dx, size_t du, double stage_cost, const double *stage_grad, double discount,
                  const double *prob, const double *prob_grad, double dt, const double *dtgrad,
                  const double *cost, double *grad);                      /* src/bellman.c:88 */
struct MCAparam;
struct MCAparam *mca_param_create(size_t, size_t);                       /* :141 */
void mca_add_grid_refs(struct MCAparam *, size_t *, double **, double, double *);   /* :168 */
void mca_param_destroy(struct MCAparam *);
struct DPparam;
struct DPparam *dp_param_create(size_t, size_t, size_t, double);         /* :219 */
void dp_param_destroy(struct DPparam *);
void dp_param_add_drift(struct DPparam *, c3sc_dyn_cb, void *);
void dp_param_add_diff(struct DPparam *, c3sc_dyn_cb, void *);
void dp_param_add_boundary(struct DPparam *, struct Boundary *);
void dp_param_add_stagecost(struct DPparam *, int (*)(double, const double *, const double *, double *, double *));
void dp_param_add_boundcost(struct DPparam *, int (*)(double, const double *, double *));
void dp_param_add_obscost(struct DPparam *, int (*)(const double *, double *));
/* NEW: the device-resident twin of the callbacks above (enum c3sc_model) and the
 * arithmetic policy (enum c3sc_arith, default C3SC_ARITH_FAST).               */
void dp_param_set_device_model(struct DPparam *, int model, const double *params, size_t nparams);
void dp_param_set_arith(struct DPparam *, int arith);
/* NEW: evaluates host callbacks and device model on n sample nodes of the grid and
 * returns the number of values that differ by more than tol (0 = registration ok). */
size_t dp_param_check_device_model(struct DPparam *, struct MCAparam *, struct c3Opt *, size_t n, double tol);

struct ControlParams;
struct ControlParams *control_params_create(size_t, size_t, struct DPparam *, struct MCAparam *,
                                            struct Workspace *, struct c3Opt *);   /* :311 */
void control_params_add_time_and_states(struct ControlParams *, double, size_t, const double *);
int control_params_get_last_res(const struct ControlParams *);
void control_params_destroy(struct ControlParams *);
struct c3sc_memory { void *shared; size_t private_; };                  /* struct Memory, :59-63 */
double bellman_control(size_t, const double *, double *, void *);        /* :367 */
int bellman_optimal(size_t, double *, double *, void *);                 /* :504 */

struct VIparam;
struct VIparam *vi_param_create(double);                                 /* :1143 */
void vi_param_destroy(struct VIparam *);
void vi_param_add_cp(struct VIparam *, struct ControlParams *);
void vi_param_add_value(struct VIparam *, struct ValueF *);
int bellman_vi(size_t, const double *, double *, void *);                /* :1295 */
struct PIparam;
struct PIparam *pi_param_create(double, struct ValueF *);                /* :1446 */
void pi_param_destroy(struct PIparam *);
void pi_param_add_cp(struct PIparam *, struct ControlParams *);
void pi_param_add_value(struct PIparam *, struct ValueF *);
int bellman_pi(size_t, const double *, double *, void *);                /* :1702 */

int mca_get_neighbor_costs(size_t d, size_t N, const double *x, struct Boundary *bound,
                           struct ValueF *vf, const size_t *ngrid, double **xgrid, size_t *fixed_ind,
                           size_t *dim_vary, int *absorbed, double *out); /* src/nodeutil.c:647 */

/* NEW: the batched forms.  x = F fibers back to back, each N_f x dx point-major
 * exactly as bellman_vi receives one (N_f = ngrid[dim_vary_f]); out likewise.
 * All fibers go to the GPU in ONE launch.                                   */
int bellman_vi_batch(size_t F, const double *x, double *out, void *vi_param);
int bellman_pi_batch(size_t F, const double *x, double *out, void *pi_param);
/* NEW: index-described fibers (no x decode): dim_vary [F], fixed_ind [F*dx],
 * out [F*ldo] with ldo = max ngrid.                                         */
int bellman_vi_batch_ind(size_t F, const int32_t *dim_vary, const int32_t *fixed_ind, double *out, void *vi_param);
int bellman_pi_batch_ind(size_t F, const int32_t *dim_vary, const int32_t *fixed_ind, double *out, void *pi_param);
/* evaluation counters kept by the reference (bellman.c:1137-1138,1439-1441) */
size_t vi_param_get_nstate_evals(const struct VIparam *);
size_t pi_param_get_npol_evals(const struct PIparam *);

/* ---- C3Control facade (set-up part; the solver loops need the C3 cross driver) --- */
struct C3Control;
struct C3Control *c3control_create(size_t, size_t, size_t, double *, double *, size_t *, double);   /* :1962 */
void c3control_destroy(struct C3Control *);
size_t *c3control_get_ngrid(struct C3Control *);
double **c3control_get_xgrid(struct C3Control *);
void c3control_set_external_boundary(struct C3Control *, size_t, char *);
void c3control_add_obstacle(struct C3Control *, double *, double *);
void c3control_add_drift(struct C3Control *, c3sc_dyn_cb, void *);
void c3control_add_diff(struct C3Control *, c3sc_dyn_cb, void *);
void c3control_add_stagecost(struct C3Control *, int (*)(double, const double *, const double *, double *, double *));
void c3control_add_boundcost(struct C3Control *, int (*)(double, const double *, double *));
void c3control_add_obscost(struct C3Control *, int (*)(const double *, double *));
/* NEW */
void c3control_set_device_model(struct C3Control *, int model, const double *params, size_t nparams);
struct DPparam *c3control_get_dp(struct C3Control *);
struct MCAparam *c3control_get_mca(struct C3Control *);
struct Workspace *c3control_get_work(struct C3Control *);
struct Boundary *c3control_get_boundary(struct C3Control *);
/* NEW: one Bellman sweep over a caller-supplied fiber list, the part of
 * c3control_step_vi (src/bellman.c:2177-2212) that is on the hot path: creates
 * ControlParams + VIparam, bumps vi_iter, runs the batch, returns node count. */
int c3control_vi_fibers(struct C3Control *, struct ValueF *, struct c3Opt *, size_t F,
                        const int32_t *dim_vary, const int32_t *fixed_ind, double *out, size_t *nevals);

/* ---- approximation arguments (src/util.h:50-66, defaults src/util.c:116-132) ---- */
enum function_class { CONSTANT, PIECEWISE, POLYNOMIAL, LINELM, CONSTELM, KERNEL };   /* C3's enum; LINELM (nodal) only */
struct ApproxArgs;
struct ApproxArgs *approx_args_init(void);
void approx_args_free(struct ApproxArgs *);
void approx_args_set_function_class(struct ApproxArgs *, enum function_class);
enum function_class approx_args_get_function_class(const struct ApproxArgs *);
void approx_args_set_cross_tol(struct ApproxArgs *, double);
double approx_args_get_cross_tol(const struct ApproxArgs *);
void approx_args_set_round_tol(struct ApproxArgs *, double);
double approx_args_get_round_tol(const struct ApproxArgs *);
void approx_args_set_kickrank(struct ApproxArgs *, size_t);
size_t approx_args_get_kickrank(const struct ApproxArgs *);
void approx_args_set_maxrank(struct ApproxArgs *, size_t);
size_t approx_args_get_maxrank(const struct ApproxArgs *);
void approx_args_set_startrank(struct ApproxArgs *, size_t);
size_t approx_args_get_startrank(const struct ApproxArgs *);
void approx_args_set_adapt(struct ApproxArgs *, int);
int approx_args_get_adapt(const struct ApproxArgs *);

/* ---- value function: interpolation, norms, evaluation (src/valuefunc.h:57-76) ---- */
/* f == bellman_vi / bellman_pi: every core of the cross is ONE device batch; any other f is called per
 * fiber on the host, as the reference does.  Cross approximation: include/c3sc_cross.h.          */
struct ValueF *valuef_interp(size_t d, int (*f)(size_t, const double *, double *, void *), void *args, const size_t *N,
                             double **grid, struct ValueF *vref, struct ApproxArgs *aargs, int verbose);   /* :603 */
double valuef_norm(struct ValueF *);                                      /* :315: continuous L2 of the piecewise-linear train */
double valuef_norm2diff(struct ValueF *, struct ValueF *);                /* :325: likewise (the unit of abs_conv_tol) */
/* NEW names: discrete l2 of the node values (no grid needed); NOT the reference's quantity */
double valuef_norm_nodal(struct ValueF *);
double valuef_norm2diff_nodal(struct ValueF *, struct ValueF *);
double valuef_eval(struct ValueF *, const double *);                      /* :345 */
/* checkpoint / resume (src/valuefunc.h:51-54).  Own file formats (the reference's are C3's): binary and a
 * 21-digit text form; loading re-samples the cores on the given grid like function_train_create_nodal.
 * save: 0 on success.  load: NULL when the file is missing or not one of ours. */
int valuef_save(struct ValueF *, char *filename);
struct ValueF *valuef_load(char *filename, size_t *ngrid, double **xgrid);
int valuef_savetxt(struct ValueF *, char *filename);
struct ValueF *valuef_loadtxt(char *filename, size_t *ngrid, double **xgrid);
/* NEW: the nodes a train built with valuef_from_cores lives on (needed by valuef_eval) */
void valuef_set_grid(struct ValueF *, double *const *xgrid);

/* ---- solver loops and the online controller (src/bellman.h:129-194) ---- */
#include <stdio.h>
struct Diag;
void diag_destroy(struct Diag **);
struct Diag *diag_create(size_t iter, int type, double norm, double abs_diff, size_t dim, size_t *ranks, double frac);
void diag_append(struct Diag **, size_t iter, int type, double norm, double abs_diff, size_t dim, size_t *ranks, double frac);
void diag_print(struct Diag *, FILE *);
int diag_save(struct Diag *, char *filename);
void c3control_add_policy_sim(struct C3Control *, struct ValueF *, struct c3Opt *opt_sim,
                              void (*transform)(size_t, const double *, double *));
int c3control_policy_eval(struct C3Control *, double t, const double *x, double *u);     /* src/bellman.c:2105 */
int c3control_controller(double, const double *, double *, void *);                      /* :2158 */
struct ValueF *c3control_step_vi(struct C3Control *, struct ValueF *, struct ApproxArgs *, struct c3Opt *, int verbose,
                                 size_t *nevals);                                         /* :2177 */
struct ValueF *c3control_step_pi(struct C3Control *, struct ValueF *, struct PIparam *, struct ApproxArgs *,
                                 struct c3Opt *, int verbose, size_t *nevals_iter);       /* :2214 */
struct ValueF *c3control_init_value(struct C3Control *, int (*f)(size_t, const double *, double *, void *), void *args,
                                    struct ApproxArgs *, int verbose);                    /* :2264 */
struct ValueF *c3control_vi_solve(struct C3Control *, size_t maxiter, double abs_conv_tol, struct ValueF *vo,
                                  struct ApproxArgs *, struct c3Opt *, int verbose, struct Diag **);   /* :2282 */
struct ValueF *c3control_pi_solve(struct C3Control *, size_t maxiter, double abs_conv_tol, struct ValueF *policy,
                                  struct ApproxArgs *, struct c3Opt *, int verbose, struct Diag **);   /* :2342 */

#ifdef __cplusplus
}
#endif
#endif /* C3SC_HOST_H */
