/* c3sc_oracle.h -- CPU oracle for the Bellman-backup hot path of goroda/c3sc.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported
 * by or executed from the product library (c3sc_b200/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load it, and there only as the checker.
 *
 * This is a plain-C restatement of the reference algorithm; every function
 * cites the /root/reference file:line it follows.  Parity is PINNED: the
 * restatement is compared (bit-for-bit for indices/flags/probabilities,
 * 1e-15 for values) against the reference's own object code built in place
 * into oracle/_ref/ (see oracle/Makefile, tests/test_oracle_vs_ref.py) and
 * against the committed fixtures in tests/golden/ generated from that build.
 * Unpinned detail: the brute-force tie rule of C3's c3opt_minimize (C3 is an
 * absent, unpinned third-party dependency) -- defined here as "first strict
 * minimum in table order".
 */
#ifndef C3SC_ORACLE_H
#define C3SC_ORACLE_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* boundary types, values of enum EBTYPE (src/boundary.h:42-47) */
enum { ORC_ABSORB = 1, ORC_PERIODIC = 2, ORC_REFLECT = 3 };

typedef int (*orc_dyn_fn)(double, const double *, const double *, double *, double *, void *);
typedef int (*orc_stage_fn)(double, const double *, const double *, double *, double *);
typedef int (*orc_bound_fn)(double, const double *, double *);
typedef int (*orc_obs_fn)(const double *, double *);

/* Everything a backup needs besides the value function. */
typedef struct orc_problem {
    size_t dx, du, dw;
    const size_t *ngrid;      /* [dx]                                   */
    double *const *xgrid;     /* [dx][ngrid[i]]                         */
    double h2;                /* hmin^2            (bellman.c:181)      */
    const double *t;          /* [2dx] h2/h_i, h2/h_i^2 (bellman.c:182-186) */
    const int *bc;            /* [dx] ORC_ABSORB / PERIODIC / REFLECT   */
    size_t nobs;              /* axis-aligned box obstacles             */
    const double *obs_lb;     /* [nobs*dx]                              */
    const double *obs_ub;     /* [nobs*dx]                              */
    double beta;              /* discount                               */
    size_t nu;                /* discrete control candidates            */
    const double *utab;       /* [nu*du] candidate-major                */
    orc_dyn_fn drift;  void *drift_arg;
    orc_dyn_fn diff;   void *diff_arg;
    orc_stage_fn stage;
    orc_bound_fn boundcost;
    orc_obs_fn obscost;
} orc_problem;

/* Nodal function-train cores: block j of core k is an r_k x r_{k+1}
 * column-major matrix at cores[k] + j*r_k*r_{k+1} (valuefunc.c:165-189). */
typedef struct orc_ft {
    size_t d;
    const size_t *n;          /* [d]   */
    const size_t *ranks;      /* [d+1], ranks[0]=ranks[d]=1 */
    double *const *cores;     /* [d]   */
} orc_ft;

/* grid constants of c3control_create + mca_add_grid_refs
 * (bellman.c:1975-1986, :181-186): h[i], hmin, h2, t[2dx]. */
void orc_grid_constants(size_t dx, const size_t *ngrid, double *const *xgrid,
                        const double *lb, const double *ub,
                        double *h, double *hmin, double *h2, double *t);

/* nodeutil.c:408-419 / :437-470 */
size_t orc_x_to_ind(double x, size_t n, const double *grid);
int orc_fiber_to_ind(size_t d, size_t N, const double *x, const size_t *ngrid,
                     double *const *xgrid, size_t *fixed_ind, size_t *dim_vary);

/* boundary.c:329-344, :668-680 */
int orc_in_obstacle(const orc_problem *p, const double *x);

/* nodeutil.c:489-627.  x is the N x d point-major fiber. */
int orc_fiber_neighbors(const orc_problem *p, const size_t *fixed_ind, size_t dim_vary,
                        const double *x, int *absorbed, size_t *nbr_vary, size_t *nbr_fixed);

/* valuefunc.c:369-585, same gemv/dot order, sequential sums. out: N x (2d+1). */
int orc_ft_fiber_nn(const orc_ft *ft, const size_t *fixed_ind, size_t dim_vary,
                    const size_t *nbr_fixed, const size_t *nbr_vary, double *out);

/* nodeutil.c:647-713: zero, decode, flags, neighbour values. */
int orc_neighbor_costs(const orc_problem *p, const orc_ft *ft, size_t N, const double *x,
                       size_t *fixed_ind, size_t *dim_vary, int *absorbed, double *costs);

/* nodeutil.c:267-406, non-gradient branch.  returns 0, or 1 if norm<1e-14. */
int orc_transition(size_t dx, size_t dw, double h2, const double *t,
                   const double *drift, const double *ddiff, double *prob, double *dt);

/* bellman.c:88-112 without gradient. */
double orc_rhs(size_t dx, double stage, double beta, const double *prob, double dt,
               const double *cost);

/* bellman.c:367-480 (one candidate) and :504-543 (argmin / absorbed shortcut).
 * scratch: dx + dx*dw + 64 doubles.  ubest receives the index into utab
 * (-1 for absorbed nodes, whose control is 0).                              */
double orc_control_value(const orc_problem *p, const double *x, const double *u,
                         const double *cost, double *prob, double *dt, double *stage,
                         int *status);
int orc_node_backup(const orc_problem *p, int absorbed, const double *x, const double *cost,
                    double *val, int *ubest);

/* bellman.c:1295-1423 minus the memo: one fiber, x = N x dx. */
int orc_vi_fiber(const orc_problem *p, const orc_ft *ft, size_t N, const double *x,
                 double *out, int *ubest, int *absorbed_out, double *costs_out);

/* bellman.c:1702-1886 minus the memo.  rows: N x (2dx+3) = [p(2dx+1), dt, g];
 * have_rows != 0 reuses the rows (later sub-iterations).                    */
int orc_pi_fiber(const orc_problem *p, const orc_ft *ft_policy, const orc_ft *ft_iter,
                 size_t N, const double *x, int have_rows, double *rows, int *ubest,
                 double *out);

/* index-based conveniences used by the tests and the CPU baseline:
 * build the fiber's x from the grid, then call the functions above.      */
void orc_fiber_points(const orc_problem *p, size_t dim_vary, const int *fixed_ind, double *x);
int orc_vi_batch(const orc_problem *p, const orc_ft *ft, size_t F, const int *dim_vary,
                 const int *fixed_ind, size_t ldo, double *out, int *ubest, int nthreads);
int orc_pi_batch(const orc_problem *p, const orc_ft *ft_policy, const orc_ft *ft_iter,
                 size_t F, const int *dim_vary, const int *fixed_ind, size_t ldo,
                 int have_rows, double *rows, int *ubest, double *out, int nthreads);

/* C3 function_train_eval for LINELM cores = multilinear interpolation of the
 * nodal cores (valuefunc.c:345-350 -> C3).  Used for held-out-grid checks. */
double orc_ft_eval_linear(const orc_ft *ft, double *const *xgrid, const double *x);

/* nodeutil.c:718-816 and bellman.c:2105-2151: the implicit policy at an off-grid state. */
int orc_neighbor_node_costs(const orc_problem *p, const orc_ft *ft, const double *x, int *absorbed, double *out);
int orc_policy_eval(const orc_problem *p, const orc_ft *ft, const double *x, double *u, double *val, int *absorbed,
                    double *costs);

#ifdef __cplusplus
}
#endif
#endif
