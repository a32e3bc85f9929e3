/* c3sc_multi.h -- the Bellman-backup hot path on several GPUs of one box from ONE process.
 *
 * The reference is a single-process C library (OpenMP inside bellman_vi, src/bellman.c:1390-1404); its drop-in
 * replacement must therefore be able to use the box's GPUs without the caller becoming an MPI / torchrun job.
 * Fibers are independent given the FT cores: a batch is cut into contiguous blocks of
 * ceil(F / G) fibers, block g runs on device g through that device's own pipeline (c3sc_b200.h), driven by a host
 * thread per device (a pipeline enqueues ~100 launches per batch; one thread for G devices would be launch-bound).
 * The fiber -> device map is a function of (F, G) only, so the policy rows of bellman_pi stay on the device that
 * produced them across sub-iterations.  The cores are uploaded once to device 0 and broadcast with ncclBroadcast
 * over NVLink (NCCL is loaded at run time, libnccl.so.2; without it: peer copies); the device-resident variant
 * ends with an ncclAllGather of the fiber values.  Every function returns a c3sc_status; c3sc_last_error() has
 * the message.
 */
#ifndef C3SC_MULTI_H
#define C3SC_MULTI_H
#include "c3sc_b200.h"
#include "c3sc_cross.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct c3sc_multi c3sc_multi;                 /* G problems (one per device) + worker threads + communicators */
typedef struct c3sc_multi_valuef c3sc_multi_valuef;   /* G device copies of one value function */

/* ndev = 0: every visible device.  devices = NULL: 0 .. ndev-1.  SURVEY 8(b)-4 `c3sc_cuda_init(ndev)`. */
int  c3sc_multi_create(const c3sc_problem_desc *desc, int ndev, const int *devices, c3sc_multi **out);
void c3sc_multi_destroy(c3sc_multi *m);
int  c3sc_multi_device_count(const c3sc_multi *m);
/* 1 if the cores travel by ncclBroadcast / the values by ncclAllGather, 0 if by peer copies (NCCL not loadable) */
int  c3sc_multi_uses_nccl(const c3sc_multi *m);
/* the per-device problem (for the single-device entries of c3sc_b200.h); g < device count */
c3sc_problem *c3sc_multi_problem(c3sc_multi *m, int g);

int  c3sc_multi_valuef_create(c3sc_multi *m, uint32_t d, const uint64_t *n, const uint64_t *ranks,
                              const double *const *cores, c3sc_multi_valuef **out);
/* same shapes, new numbers: host -> device 0 -> broadcast -> derived copies rebuilt on every device */
int  c3sc_multi_valuef_update(c3sc_multi_valuef *vf, const double *const *cores);
void c3sc_multi_valuef_destroy(c3sc_multi_valuef *vf);
const c3sc_valuef *c3sc_multi_valuef_get(const c3sc_multi_valuef *vf, int g);

/* fibers [begin, end) of a batch of F that device g of G owns (contiguous blocks of ceil(F / G)) */
void c3sc_multi_shard(size_t F, int G, int g, size_t *begin, size_t *end);

/* bellman_vi over F fibers, host buffers, sharded over the devices; same contract as c3sc_vi_batch */
int c3sc_multi_vi_batch(c3sc_multi *m, const c3sc_multi_valuef *vf, size_t F, const int32_t *dim_vary,
                        const int32_t *fixed_ind, size_t ldo, double *value, int32_t *argmin);

/* bellman_pi over F fibers with the policy rows RESIDENT on the owning devices.  `slot` (< 64) names the batch:
 * a sub-iteration (have_rows != 0) must pass the same slot, F and fibers as the improvement call (have_rows == 0)
 * that filled it -- the cross driver uses one slot per (core, sweep direction).  Nothing but the fiber descriptors
 * goes up and nothing but the values comes down (src/bellman.c:1803-1880 keeps the rows in pi_prob_htable).  */
int c3sc_multi_pi_batch(c3sc_multi *m, const c3sc_multi_valuef *vf_policy, const c3sc_multi_valuef *vf_iter,
                        uint32_t slot, size_t F, const int32_t *dim_vary, const int32_t *fixed_ind, size_t ldo,
                        int have_rows, double *value);
/* drop the resident rows (a new policy: c3control_pi_solve resets pi_prob_htable, src/bellman.c:2352-2355) */
int c3sc_multi_pi_reset(c3sc_multi *m);

/* Device-resident variant: fiber descriptors on the host, the values of ALL fibers gathered on EVERY device
 * (ncclAllGather over NVLink): d_gathered[g] is a device-g buffer of c3sc_multi_gathered_count(F, G, ldo) doubles,
 * fiber f's values at [f*ldo, f*ldo + ldo) (blocks are padded to ceil(F/G) fibers, so the layout is that of one
 * big batch).  Synchronous. */
size_t c3sc_multi_gathered_count(size_t F, int G, size_t ldo);
int c3sc_multi_vi_batch_gathered(c3sc_multi *m, const c3sc_multi_valuef *vf, size_t F, const int32_t *dim_vary,
                                 const int32_t *fixed_ind, size_t ldo, double *const *d_gathered);

/* c3control_step_vi / _pi (src/bellman.c:2177-2262) through the host cross driver with every core batch sharded
 * over the devices; between steps the new cores go back up with c3sc_multi_valuef_update. */
int c3sc_cross_run_vi_multi(c3sc_cross *c, c3sc_multi *m, const c3sc_multi_valuef *vf, const c3sc_cross_opts *opts,
                            double *const *cores, uint64_t *nfibers, double *rel_change);
int c3sc_cross_run_pi_multi(c3sc_cross *c, c3sc_multi *m, const c3sc_multi_valuef *vf_policy,
                            const c3sc_multi_valuef *vf_iter, const c3sc_cross_opts *opts, double *const *cores,
                            uint64_t *nfibers, double *rel_change);

#ifdef __cplusplus
}
#endif
#endif /* C3SC_MULTI_H */
