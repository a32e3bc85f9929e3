// explicit instantiations: LQG chain-of-integrators, dx = 2,4,6
#include "control_kernel.cuh"
namespace c3sc {
__device__ void fused_walk_lqg_lo(int dx, const CtlArgs &c, const FusedCta &w)
{
    switch (dx) {
    case 2: fused_walk_m<LqgNd<2>>(c, w); break;
    case 4: fused_walk_m<LqgNd<4>>(c, w); break;
    case 6: fused_walk_m<LqgNd<6>>(c, w); break;
    }
}
int fused_ok_lqg_lo(int dx, int arith, const CtlArgs &c, int pi_eval)
{
    switch (dx) {
    case 2: return fused_ok_m<LqgNd<2>>(arith, c, pi_eval);
    case 4: return fused_ok_m<LqgNd<4>>(arith, c, pi_eval);
    case 6: return fused_ok_m<LqgNd<6>>(arith, c, pi_eval);
    }
    return 0;
}
int launch_control_lqg_lo(int dx, int arith, const CtlArgs &a, int pi_eval, cudaStream_t st)
{
    switch (dx) {
    case 2: return launch_control_m<LqgNd<2>>(arith, a, pi_eval, st);
    case 4: return launch_control_m<LqgNd<4>>(arith, a, pi_eval, st);
    case 6: return launch_control_m<LqgNd<6>>(arith, a, pi_eval, st);
    }
    return -1;
}
int launch_model_eval_lqg_lo(int dx, const DevProblem &P, int n, const double *x, const double *u, double *drift,
                             double *sig, double *stage, double *bound, double *obs, cudaStream_t st)
{
    switch (dx) {
    case 2: return launch_model_eval_t<LqgNd<2>>(P, n, x, u, drift, sig, stage, bound, obs, st);
    case 4: return launch_model_eval_t<LqgNd<4>>(P, n, x, u, drift, sig, stage, bound, obs, st);
    case 6: return launch_model_eval_t<LqgNd<6>>(P, n, x, u, drift, sig, stage, bound, obs, st);
    }
    return -1;
}
int build_ctab_lqg_lo(int dx, const DevProblem &P, double *ctab, cudaStream_t st)
{
    switch (dx) {
    case 2: return build_ctab_t<LqgNd<2>>(P, ctab, st);
    case 4: return build_ctab_t<LqgNd<4>>(P, ctab, st);
    case 6: return build_ctab_t<LqgNd<6>>(P, ctab, st);
    }
    return -1;
}
int launch_node_backup_lqg_lo(int dx, int arith, const DevProblem &P, int n, const double *x, const double *costs,
                              const int *absorbed, double *value, int *argmin, cudaStream_t st)
{
    switch (dx) {
    case 2: return launch_node_backup_t<LqgNd<2>>(arith, P, n, x, costs, absorbed, value, argmin, st);
    case 4: return launch_node_backup_t<LqgNd<4>>(arith, P, n, x, costs, absorbed, value, argmin, st);
    case 6: return launch_node_backup_t<LqgNd<6>>(arith, P, n, x, costs, absorbed, value, argmin, st);
    }
    return -1;
}
int launch_control_value_lqg_lo(int dx, int arith, const DevProblem &P, int n, const double *x, const double *u,
                                const double *costs, double *value, cudaStream_t st)
{
    switch (dx) {
    case 2: return launch_control_value_t<LqgNd<2>>(arith, P, n, x, u, costs, value, st);
    case 4: return launch_control_value_t<LqgNd<4>>(arith, P, n, x, u, costs, value, st);
    case 6: return launch_control_value_t<LqgNd<6>>(arith, P, n, x, u, costs, value, st);
    }
    return -1;
}
}  // namespace c3sc
