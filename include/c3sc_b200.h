/* c3sc_b200.h -- thin C-ABI CUDA layer for the c3sc Bellman-backup hot path.
 *
 * Plain pointers and sizes only (no torch / C++ types): this is the surface
 * the reference's C host code binds (see INTEGRATION.md for the stubs a
 * maintainer adds to src/bellman.c, src/valuefunc.c).  Every entry point
 * names the reference interface it replaces (path:line under
 * /root/reference).  All functions return 0 on success and a non-zero
 * C3SC_E* code otherwise; c3sc_last_error() gives the message.  There is no
 * CPU fallback: without a CUDA device every compute entry fails loudly.
 *
 * Conventions
 *   - a FIBER is (dim_vary, fixed_ind[dx]): all nodes of the grid that share
 *     fixed_ind except along dim_vary (what C3's cross approximation hands
 *     to bellman_vi as N points, src/bellman.c:1295).  fixed_ind[dim_vary]
 *     is ignored.
 *   - per-node outputs of fiber f live at [f*ldo + j], j < ngrid[dim_vary];
 *     ldo >= max ngrid.
 *   - index outputs are int32; values fp64.
 */
#ifndef C3SC_B200_H
#define C3SC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define C3SC_MAXD 16          /* state dimensions supported by the kernels    */
#define C3SC_MAXOBS 10        /* obstacle boxes (src/boundary.c:393)          */
#define C3SC_MAXPEERS 8       /* GPUs of one box for the fused value all-gather */

enum c3sc_status {
    C3SC_OK = 0,
    C3SC_EINVAL = 1,          /* bad argument                                 */
    C3SC_ECUDA = 2,           /* CUDA runtime error (message has the detail)  */
    C3SC_ENODEV = 3,          /* no usable CUDA device                        */
    C3SC_ENUMERIC = 4,        /* transition normaliser < 1e-14 somewhere: the
                                 reference asserts here (src/bellman.c:452,
                                 src/nodeutil.c:365-367)                      */
    C3SC_EUNSUPPORTED = 5     /* model / dimension not instantiated           */
};

/* enum EBTYPE of the reference (src/boundary.h:42-47), same values */
enum c3sc_bc { C3SC_ABSORB = 1, C3SC_PERIODIC = 2, C3SC_REFLECT = 3 };

/* Device-resident dynamics + cost models.  The reference takes host function
 * pointers (src/dynamics.c:66,172; src/bellman.c:215-217) that a kernel
 * cannot call; the examples' definitions are therefore built in and selected
 * by id next to the unchanged host pointers.
 *   LQGND      examples/lqgnd/lqgnd.c:80-186 (== examples/lqg2d_new for dx=2)
 *              params [ss0, ss1, boundcost, obscost]
 *   DOUBLE_INT examples/double_int/double_int.c:80-162, same params
 *   DUBINS     examples/dubinscar_new/dubinscar.c:40-121
 *              params [s_xy, s_theta, stage, boundcost, obscost]
 *   SKID5D     examples/skidding5d/scar.c:39-176, params [obscost]         */
enum c3sc_model {
    C3SC_MODEL_NONE = 0,      /* geometry only: grid, boundary types, obstacles.  Enough for the entries that do
                                 not evaluate dynamics (c3sc_neighbor_costs_batch, c3sc_neighbor_node_costs_batch,
                                 c3sc_valuef_eval_batch); nu / controls may be 0 / NULL; backups are refused   */
    C3SC_MODEL_LQGND = 1,
    C3SC_MODEL_DOUBLE_INT = 2,
    C3SC_MODEL_DUBINS = 3,
    C3SC_MODEL_SKID5D = 4,
    C3SC_MODEL_USER = 5       /* struct c3sc::UserModel of the header the library was built with (make USER_MODEL=...;
                                 default examples/user_model_vdp.cuh: controlled Van der Pol oscillator, params
                                 [mu, s0, s1, boundcost, obscost]) -- INTEGRATION.md, "Adding a device model"     */
};

/* EXACT reproduces the reference's operation order without fused
 * multiply-add and with IEEE division per probability: transition
 * probabilities and dt are bit-identical to src/nodeutil.c:267-406.
 * FAST reassociates (one reciprocal per control, FMA, control-independent
 * dimensions hoisted): <= a few ulp from EXACT, far inside the 1e-12 bar. */
enum c3sc_arith { C3SC_ARITH_EXACT = 0, C3SC_ARITH_FAST = 1 };

typedef struct c3sc_problem c3sc_problem;   /* device mirror of MCAparam+DPparam+Boundary+c3opt table */
typedef struct c3sc_valuef c3sc_valuef;     /* device mirror of ValueF::cores */

/* What c3control_create / mca_add_grid_refs / dp_param_* / boundary_* hold on
 * the host (src/bellman.c:118-285,1962-1999; src/boundary.c:353-489), flattened. */
typedef struct c3sc_problem_desc {
    uint32_t dx, du, dw;
    const uint64_t *ngrid;          /* [dx]                                   */
    const double *const *xgrid;     /* [dx][ngrid[i]], uploaded as given      */
    double h2;                      /* MCAparam::h2                           */
    const double *t;                /* [2dx] MCAparam::t                      */
    const int32_t *bc;              /* [dx] enum c3sc_bc                      */
    uint32_t nobs;
    const double *obs_lb;           /* [nobs*dx] BoundRect::lb                */
    const double *obs_ub;           /* [nobs*dx] BoundRect::ub                */
    double discount;                /* DPparam::discount                      */
    uint32_t nu;                    /* brute-force candidates                 */
    const double *controls;         /* [nu*du] candidate-major (c3opt table)  */
    int32_t model;                  /* enum c3sc_model                        */
    const double *model_params;     /* may be NULL                            */
    uint32_t n_model_params;
    int32_t arith;                  /* enum c3sc_arith                        */
} c3sc_problem_desc;

/* Optional per-batch outputs (device pointers, NULL = not wanted). */
typedef struct c3sc_batch_out {
    double  *value;       /* [F*ldo]          backed-up values (bellman_vi `out`)        */
    int32_t *argmin;      /* [F*ldo]          index into the control table, -1 absorbed  */
    int32_t *absorbed;    /* [F*ldo]          0 / 1 / -1 (process_fibers_neighbor)        */
    double  *costs;       /* [F*ldo*(2dx+1)]  neighbour values (valuef_eval_fiber_ind_nn) */
    double  *rows;        /* [F*ldo*(2dx+3)]  [p(2dx+1), dt, g] at the argmin (bellman_pi) */
    int32_t *nbr_vary;    /* [F*ldo*2]        neighbours along the fiber                  */
    int32_t *nbr_fixed;   /* [F*2*(dx-1)]     neighbours in the fixed dimensions          */
    /* Fused all-gather (one process per GPU, peer-mapped device buffers, e.g. cudaIpcOpenMemHandle):
     * the control kernel stores every backed-up value ALSO to value_peers[g][peer_offset + f*ldo + j]
     * for g < n_peers, i.e. straight into every rank's gathered buffer over NVLink while it computes.
     * The caller orders completion across ranks (a barrier) before reading.  n_peers = 0: off.      */
    double  *value_peers[C3SC_MAXPEERS];
    uint32_t n_peers;
    uint64_t peer_offset;
    /* how the values reach the peers: 0 = stores from the control kernel (each value crosses NVLink once per peer as
     * an 8-byte store; no extra pass); 1 = one bulk copy per pipeline chunk and peer on the library's copy stream
     * (copy engines, no SM involvement, overlapped with the next chunk's kernels; needs `value`).  A peer pointer that
     * equals value - peer_offset (the rank's own gathered buffer used as its output) is skipped.  2 = the same, but the
     * chunk travels by ONE small kernel (48 CTAs, 16-byte stores to every peer) on the copy stream instead of one copy
     * per peer: a twelfth of the launches at 8 GPUs.                                                               */
    uint32_t peer_mode;
} c3sc_batch_out;

/* ---- runtime ---------------------------------------------------------- */
int c3sc_cuda_init(int device);                  /* select device, create context; ONE device per
                                                    process (multi-GPU = one process per GPU)     */
int c3sc_cuda_device_count(void);
const char *c3sc_last_error(void);
const char *c3sc_version(void);
/* kernels launched by this library since load (bench.py's gpu_launches) */
uint64_t c3sc_launch_count(void);

/* Best-of-`repeats` throughput of a pure DFMA loop on the current device, in
 * TFLOP/s (FMA = 2 flop): the measured FP64-pipe roofline denominator.      */
int c3sc_measure_fp64_peak(double *tflops, int iters, int repeats);
/* The same for the FP64 tensor path: a pure DMMA (mma.sync.m8n8k4.f64) loop.  On B200 the two figures are within
 * 10 % of each other and do not add up (one execution resource, profiles/r01_fp64_pipe_microbench.md); this one is
 * the denominator for kernels whose flops run as DMMAs (stage 1).                                            */
int c3sc_measure_fp64_tensor_peak(double *tflops, int iters, int repeats);

/* The library's own out-of-bounds check (the GPU pool it is developed on has compute-sanitizer switched off).  With
 * C3SC_GUARD=1 in the environment before the first allocation, every device allocation of the library sits between two
 * 64 kB zones of 0xFF bytes (NaN as doubles, -1 as ints).  c3sc_guard_check() returns how many zones were damaged (a write
 * outside an allocation); a READ outside an allocation that reaches a result turns it into NaN, which the parity tests
 * catch (tests/test_gpu_paths.py::test_guard_zones_*).  Returns 0 when guards are off.                          */
int c3sc_guard_check(void);

/* Page-locked (pinned, portable) host memory.  The host-buffer entries accept any host pointer, but only page-locked
 * buffers move at PCIe speed and overlap with the kernels (cudaMemcpyAsync from pageable memory is staged by the
 * driver): allocate the fiber descriptors and result arrays of large batches with these.                     */
int c3sc_host_alloc(size_t bytes, void **ptr);
int c3sc_host_free(void *ptr);

/* Peer-mapped device buffers for the fused all-gather (c3sc_batch_out::value_peers), one process per
 * GPU: every rank creates its gathered buffer (cudaMalloc + cudaIpcGetMemHandle), exchanges the 64-byte
 * handles out of band, and opens the other ranks' buffers (cudaIpcOpenMemHandle with lazy peer access).
 * close: opened != 0 for buffers obtained from _open, 0 for the rank's own.                          */
int c3sc_peer_buffer_create(size_t bytes, void **dev, unsigned char handle[64]);
int c3sc_peer_buffer_open(const unsigned char handle[64], void **dev);
int c3sc_peer_buffer_close(void *dev, int opened);

/* ---- problem / value function ------------------------------------------ */
int  c3sc_problem_create(const c3sc_problem_desc *desc, c3sc_problem **out);
void c3sc_problem_destroy(c3sc_problem *p);
/* returns C3SC_ENUMERIC if any launch since the last check hit norm<1e-14, C3SC_EINVAL if
 * one was given a fiber descriptor outside the grid; synchronises the device.               */
int  c3sc_problem_check(c3sc_problem *p);
/* how stage 2 walks the control table in FAST arithmetic: 0 = plain table walk, 1 = candidates grouped by
 * their share of the normaliser, 2 = shared-prefix walk of a full {lo, 0, hi}^du grid */
int  c3sc_problem_control_path(const c3sc_problem *p);

/* valuef_precompute_cores layout (src/valuefunc.c:165-189): block j of core k
 * is an r_k x r_{k+1} column-major matrix at cores[k] + j*r_k*r_{k+1}.      */
int  c3sc_valuef_create(uint32_t d, const uint64_t *n, const uint64_t *ranks,
                        const double *const *cores, c3sc_valuef **out);
/* same shapes, new numbers (next VI iterate); host pointers               */
int  c3sc_valuef_update(c3sc_valuef *vf, const double *const *cores);
/* contiguous device buffer holding all cores (for ncclBroadcast)           */
int  c3sc_valuef_device_buffer(c3sc_valuef *vf, double **dev, size_t *count);
/* after writing that buffer on the device (e.g. a broadcast): rebuild the
 * derived copies the kernels read; asynchronous on `stream`.  The value function remembers the event of its last commit: a
 * c3sc_vi_batch_dev on ANOTHER stream runs its grouping and chain plan (they read the descriptors only) and waits for that
 * event right before its first kernel that reads the cores -- so a broadcast + commit on a side stream overlaps the plan of
 * the batch.  (The caller still orders the broadcast after the previous batch's kernels: they read the old cores.)       */
int  c3sc_valuef_commit(c3sc_valuef *vf, void *stream);
void c3sc_valuef_destroy(c3sc_valuef *vf);

/* valuef_norm / valuef_norm2diff (/root/reference/src/valuefunc.c:315-335 -> C3 function_train_norm2 / norm2diff on LINELM cores)
 * NEXT TO THE CORES: the continuous L2 inner product of two piecewise-linear function trains whose cores are on the device
 * (SURVEY 8(f)-3).  xgrid[k]: the n[k] node coordinates of dimension k (host).  One launch per dimension, the node range of a
 * dimension split over 32 CTAs, partial results summed in a fixed order (run-to-run reproducible); norm2diff shares its launches
 * between <a,a>, <a,b> and <b,b>.  The host restatement of the same formula is c3sc_cores_dot_l2 (c3sc_cross.h). */
int c3sc_valuef_dot_l2(const c3sc_valuef *a, const c3sc_valuef *b, const double *const *xgrid, double *out);
int c3sc_valuef_norm_l2(const c3sc_valuef *a, const double *const *xgrid, double *out);
int c3sc_valuef_norm2diff_l2(const c3sc_valuef *a, const c3sc_valuef *b, const double *const *xgrid, double *out);

/* ---- fiber descriptors --------------------------------------------------- */
/* A fiber is (dim_vary, fixed_ind[dx]): what convert_fiber_to_ind (src/nodeutil.c:437-470) decodes from
 * the reference's point list.  Valid means 0 <= dim_vary < dx and 0 <= fixed_ind[i] < ngrid[i] (the slot of
 * the varying dimension is ignored by the kernels but must be in range too; the reference stores the first
 * node's index there, i.e. 0).  EVERY batch entry validates its descriptors ON THE DEVICE, in the grouping
 * kernel that reads them anyway (no host scan, no extra pass): an invalid descriptor makes the call return
 * C3SC_EINVAL with the smallest offending fiber id in c3sc_last_error() -- the batch analogue of
 * convert_fiber_to_ind's non-zero return -- and is clamped into the grid wherever a kernel uses it, so it is
 * never a read outside the cores.  The host-buffer entries report it from the call itself; the asynchronous
 * *_dev entries report it from the next c3sc_problem_check().  c3sc_fibers_check is the same test as a scan
 * of HOST arrays, with no device work, for callers who want the answer before queueing anything.        */
int c3sc_fibers_check(const c3sc_problem *p, size_t F, const int32_t *dim_vary, const int32_t *fixed_ind);

/* ---- the hot path, device-resident arguments ---------------------------- */
/* Concurrency contract (as for the reference's Workspace, src/util.c:689-964, which one C3Control owns and
 * bellman_vi / bellman_pi mutate): ONE batch in flight per c3sc_problem.  A problem owns the pipeline scratch
 * (grouping, cost scratch, lanes) that its batches reuse; queueing a second batch of the SAME problem on another
 * stream or thread before the first has finished is a data race.  Different problems are independent.  The
 * handle-free entries further down (c3sc_rhs_batch, c3sc_transition_raw, c3sc_ft_fiber_nn_batch) share static
 * scratch and serialise themselves on a mutex.  c3sc_last_error() is per thread.                        */
/* bellman_vi over F fibers (src/bellman.c:1295-1423, memo dropped: backups are
 * pure).  d_dim_vary [F], d_fixed_ind [F*dx] int32 device arrays.  stream is
 * a cudaStream_t (NULL = default stream).  Asynchronous.                   */
int c3sc_vi_batch_dev(c3sc_problem *p, const c3sc_valuef *vf, size_t F,
                      const int32_t *d_dim_vary, const int32_t *d_fixed_ind,
                      size_t ldo, const c3sc_batch_out *out, void *stream);

/* Measurement entry: stage 1 of the pipeline alone (neighbour values into the library's own scratch, nothing
 * returned), launched exactly as c3sc_vi_batch_dev launches it.  bench.py times it for roofline.stage1_live. */
int c3sc_stage1_batch_dev(c3sc_problem *p, const c3sc_valuef *vf, size_t F,
                          const int32_t *d_dim_vary, const int32_t *d_fixed_ind, size_t ldo, void *stream);

/* bellman_pi over F fibers (src/bellman.c:1702-1886).  have_rows == 0: pick
 * u* against vf_policy, store the policy rows in d_rows (and argmin if
 * wanted), then evaluate against vf_iter; have_rows != 0: later sub-iteration,
 * reuse d_rows.  d_rows [F*ldo*(2dx+3)].  When vf_policy and vf_iter are the
 * SAME object (the first sub-iteration of a solver step) the evaluation takes
 * its neighbour values from the improvement's scratch: one pass of stage 1,
 * bit-identical results.                                                    */
int c3sc_pi_batch_dev(c3sc_problem *p, const c3sc_valuef *vf_policy, const c3sc_valuef *vf_iter,
                      size_t F, const int32_t *d_dim_vary, const int32_t *d_fixed_ind,
                      size_t ldo, int have_rows, double *d_rows, int32_t *d_argmin,
                      double *d_value, void *stream);

/* ---- the hot path, host buffers (H2D + kernel + D2H inside) -------------- */
/* What the reference-facing wrappers call.  value [F*ldo]; argmin may be NULL. */
int c3sc_vi_batch(c3sc_problem *p, const c3sc_valuef *vf, size_t F,
                  const int32_t *dim_vary, const int32_t *fixed_ind,
                  size_t ldo, double *value, int32_t *argmin);
/* One rank's block of a SHARDED batch (one process per GPU): as c3sc_vi_batch, and the values are also gathered into
 * every rank's peer-mapped device buffer -- peers->value_peers / n_peers / peer_offset / peer_mode as in
 * c3sc_batch_out, its other fields ignored.  The caller orders completion across ranks (a barrier) before reading. */
int c3sc_vi_batch_peers(c3sc_problem *p, const c3sc_valuef *vf, size_t F,
                        const int32_t *dim_vary, const int32_t *fixed_ind,
                        size_t ldo, double *value, int32_t *argmin, const c3sc_batch_out *peers);
/* Host-side debugging / parity variant that returns every intermediate.    */
int c3sc_vi_batch_debug(c3sc_problem *p, const c3sc_valuef *vf, size_t F,
                        const int32_t *dim_vary, const int32_t *fixed_ind, size_t ldo,
                        double *value, int32_t *argmin, int32_t *absorbed, double *costs,
                        double *rows, int32_t *nbr_vary, int32_t *nbr_fixed);
/* rows: host buffer [F*ldo*(2dx+3)], in/out exactly as d_rows above.       */
int c3sc_pi_batch(c3sc_problem *p, const c3sc_valuef *vf_policy, const c3sc_valuef *vf_iter,
                  size_t F, const int32_t *dim_vary, const int32_t *fixed_ind,
                  size_t ldo, int have_rows, double *rows, int32_t *argmin, double *value);

/* bellman_pi with the policy rows RESIDENT on the device: only descriptors go up and values come down, the
 * 184 B/node row record of src/bellman.c:1810 never crosses PCIe.
 *   _resident  rows of the whole batch in a caller-owned DEVICE buffer d_rows [F*ldo*(2dx+3)] (written when
 *              have_rows == 0, read otherwise);
 *   _store     rows filed per fiber in a c3sc_rowstore -- a device object of its own, because it outlives the
 *              per-step problems: the reference keeps a policy's rows in the Workspace (pi_prob_htable) across the
 *              c3control_step_pi calls of one c3control_pi_solve.  Fiber f of a call owns slot row_id[f]
 *              (< reserved capacity; reserving more keeps what is filed).                                        */
int c3sc_pi_batch_resident(c3sc_problem *p, const c3sc_valuef *vf_policy, const c3sc_valuef *vf_iter, size_t F,
                           const int32_t *dim_vary, const int32_t *fixed_ind, size_t ldo, int have_rows,
                           double *d_rows, double *value);
typedef struct c3sc_rowstore c3sc_rowstore;
int  c3sc_rowstore_create(uint32_t dx, size_t ldo, c3sc_rowstore **out);
int  c3sc_rowstore_reserve(c3sc_rowstore *s, size_t capacity_fibers);
void c3sc_rowstore_destroy(c3sc_rowstore *s);
int c3sc_pi_batch_store(c3sc_problem *p, const c3sc_valuef *vf_policy, const c3sc_valuef *vf_iter, size_t F,
                        const int32_t *dim_vary, const int32_t *fixed_ind, size_t ldo, int have_rows,
                        c3sc_rowstore *store, const int32_t *row_id, double *value);

/* mca_get_neighbor_costs (src/nodeutil.c:647-713) over F fibers: flags, neighbour
 * indices and FT neighbour values only (process_fibers_neighbor + 
 * valuef_eval_fiber_ind_nn).  nbr_vary / nbr_fixed may be NULL.               */
int c3sc_neighbor_costs_batch(c3sc_problem *p, const c3sc_valuef *vf, size_t F,
                              const int32_t *dim_vary, const int32_t *fixed_ind, size_t ldo,
                              int32_t *absorbed, double *costs, int32_t *nbr_vary, int32_t *nbr_fixed);
/* process_fibers_neighbor (src/nodeutil.c:489-627) over F fibers: absorbed [F*ldo] (0 / 1 / -1), nbr_vary
 * [F*ldo*2], nbr_fixed [F*2*(dx-1)] (either may be NULL).  No value function, no dynamics (C3SC_MODEL_NONE ok). */
int c3sc_fiber_flags_batch(c3sc_problem *p, size_t F, const int32_t *dim_vary, const int32_t *fixed_ind,
                           size_t ldo, int32_t *absorbed, int32_t *nbr_vary, int32_t *nbr_fixed);
/* bellman_optimal (src/bellman.c:504-543) at n nodes with caller-supplied
 * neighbour costs [n*(2dx+1)] and flags (NULL = all 0): the building block of
 * the online controller (src/bellman.c:2105-2175).                           */
int c3sc_node_backup_batch(c3sc_problem *p, size_t n, const double *x, const double *costs,
                           const int32_t *absorbed, double *value, int32_t *argmin);

/* ---- the implicit policy at off-grid states (online controller) ------------- */
/* valuef_eval (src/valuefunc.c:345-350 -> C3 function_train_eval on LINELM cores): piecewise-
 * linear interpolation of the nodal cores at n points x[n*dx]; 0 outside the grid.            */
int c3sc_valuef_eval_batch(c3sc_problem *p, const c3sc_valuef *vf, size_t n, const double *x, double *out);
/* mca_get_neighbor_node_costs (src/nodeutil.c:718-816) at n off-grid states x[n*dx]: absorbed [n],
 * costs [n*(2dx+1)] (slot 2i / 2i+1 = V at x -+ h_i e_i with the boundary stand-ins; everything V(x) and
 * flag -1 inside an obstacle; slot 2dx = V(x)).  No dynamics involved: a C3SC_MODEL_NONE problem will do. */
int c3sc_neighbor_node_costs_batch(c3sc_problem *p, const c3sc_valuef *vf, size_t n, const double *x,
                                   int32_t *absorbed, double *costs);
/* c3control_policy_eval (src/bellman.c:2105-2151) at n states: mca_get_neighbor_node_costs
 * (src/nodeutil.c:718-816: V at x -+ h e_i with the boundary stand-ins, all V(x) and flag -1 inside
 * an obstacle) then bellman_optimal.  u [n*du]; value [n], absorbed [n], costs [n*(2dx+1)] may be
 * NULL.  costs[2dx] (the state itself, left unset by the reference) is V(x).                   */
int c3sc_policy_eval_batch(c3sc_problem *p, const c3sc_valuef *vf, size_t n, const double *x, double *u,
                           double *value, int32_t *absorbed, double *costs);

/* bellman_control (src/bellman.c:367-480, grad_u == NULL) at n (x, u, costs)
 * triples; u need not be in the control table.                               */
int c3sc_control_value_batch(c3sc_problem *p, size_t n, const double *x, const double *u,
                             const double *costs, double *value);
/* bellmanrhs (src/bellman.c:88-112, no gradient) on n raw tuples; no problem handle. */
int c3sc_rhs_batch(int arith, uint32_t dx, double discount, size_t n, const double *prob,
                   const double *dt, const double *stage, const double *cost, double *out);
/* transition_assemble (src/nodeutil.c:267-406, non-gradient branch) with the raw
 * (h2, t[2dx]) of its signature; no problem handle.                          */
int c3sc_transition_raw(int arith, uint32_t dx, double h2, const double *t, size_t n,
                        const double *drift, const double *sigma_diag, double *prob, double *dt,
                        int32_t *status);
/* valuef_eval_fiber_ind_nn (src/valuefunc.c:369-585) with caller-supplied neighbour
 * indices: nbr_fixed [F*2*(d-1)], nbr_vary [F*ldo*2]; costs [F*ldo*(2d+1)].  */
int c3sc_ft_fiber_nn_batch(const c3sc_valuef *vf, size_t F, const int32_t *dim_vary,
                           const int32_t *fixed_ind, const int32_t *nbr_fixed,
                           const int32_t *nbr_vary, size_t ldo, double *costs);

/* ---- pieces of the path exposed for parity tests -------------------------- */
/* transition_assemble (src/nodeutil.c:267-406, non-gradient branch) for n
 * independent (drift[dx], diag sigma[dx]) pairs given on the host;
 * prob [n*(2dx+1)], dt [n], status [n] (0 ok, 1 norm<1e-14).               */
int c3sc_transition_batch(c3sc_problem *p, size_t n, const double *drift, const double *sigma_diag,
                          double *prob, double *dt, int32_t *status);
/* drift_eval / diff_eval / stagecost / boundcost / obscost of the device model
 * at n (x,u) pairs: the device-vs-host-callback equality check.
 * x [n*dx], u [n*du] -> drift [n*dx], sigma_diag [n*dx], stage/bound/obs [n]. */
int c3sc_model_eval(c3sc_problem *p, size_t n, const double *x, const double *u,
                    double *drift, double *sigma_diag, double *stage, double *bound, double *obs);

#ifdef __cplusplus
}
#endif
#endif /* C3SC_B200_H */
