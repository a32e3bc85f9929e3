// explicit instantiations: LQG chain-of-integrators, dx = 8,10,12
#include "control_kernel.cuh"
namespace c3sc {
__device__ void fused_walk_lqg_hi(int dx, const CtlArgs &c, const FusedCta &w)
{
    switch (dx) {
    case 8: fused_walk_m<LqgNd<8>>(c, w); break;
    case 10: fused_walk_m<LqgNd<10>>(c, w); break;
    case 12: fused_walk_m<LqgNd<12>>(c, w); break;
    }
}
int fused_ok_lqg_hi(int dx, int arith, const CtlArgs &c, int pi_eval)
{
    switch (dx) {
    case 8: return fused_ok_m<LqgNd<8>>(arith, c, pi_eval);
    case 10: return fused_ok_m<LqgNd<10>>(arith, c, pi_eval);
    case 12: return fused_ok_m<LqgNd<12>>(arith, c, pi_eval);
    }
    return 0;
}
int launch_control_lqg_hi(int dx, int arith, const CtlArgs &a, int pi_eval, cudaStream_t st)
{
    switch (dx) {
    case 8: return launch_control_m<LqgNd<8>>(arith, a, pi_eval, st);
    case 10: return launch_control_m<LqgNd<10>>(arith, a, pi_eval, st);
    case 12: return launch_control_m<LqgNd<12>>(arith, a, pi_eval, st);
    }
    return -1;
}
int launch_model_eval_lqg_hi(int dx, const DevProblem &P, int n, const double *x, const double *u, double *drift,
                             double *sig, double *stage, double *bound, double *obs, cudaStream_t st)
{
    switch (dx) {
    case 8: return launch_model_eval_t<LqgNd<8>>(P, n, x, u, drift, sig, stage, bound, obs, st);
    case 10: return launch_model_eval_t<LqgNd<10>>(P, n, x, u, drift, sig, stage, bound, obs, st);
    case 12: return launch_model_eval_t<LqgNd<12>>(P, n, x, u, drift, sig, stage, bound, obs, st);
    }
    return -1;
}
int build_ctab_lqg_hi(int dx, const DevProblem &P, double *ctab, cudaStream_t st)
{
    switch (dx) {
    case 8: return build_ctab_t<LqgNd<8>>(P, ctab, st);
    case 10: return build_ctab_t<LqgNd<10>>(P, ctab, st);
    case 12: return build_ctab_t<LqgNd<12>>(P, ctab, st);
    }
    return -1;
}
int launch_node_backup_lqg_hi(int dx, int arith, const DevProblem &P, int n, const double *x, const double *costs,
                              const int *absorbed, double *value, int *argmin, cudaStream_t st)
{
    switch (dx) {
    case 8: return launch_node_backup_t<LqgNd<8>>(arith, P, n, x, costs, absorbed, value, argmin, st);
    case 10: return launch_node_backup_t<LqgNd<10>>(arith, P, n, x, costs, absorbed, value, argmin, st);
    case 12: return launch_node_backup_t<LqgNd<12>>(arith, P, n, x, costs, absorbed, value, argmin, st);
    }
    return -1;
}
int launch_control_value_lqg_hi(int dx, int arith, const DevProblem &P, int n, const double *x, const double *u,
                                const double *costs, double *value, cudaStream_t st)
{
    switch (dx) {
    case 8: return launch_control_value_t<LqgNd<8>>(arith, P, n, x, u, costs, value, st);
    case 10: return launch_control_value_t<LqgNd<10>>(arith, P, n, x, u, costs, value, st);
    case 12: return launch_control_value_t<LqgNd<12>>(arith, P, n, x, u, costs, value, st);
    }
    return -1;
}
}  // namespace c3sc
