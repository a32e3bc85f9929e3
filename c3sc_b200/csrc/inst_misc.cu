// explicit instantiations: double integrator (dx 2,3,4), Dubins car, skidding car; transition test kernel
#include "control_kernel.cuh"
namespace c3sc {
__device__ void fused_walk_misc(int model, int dx, const CtlArgs &c, const FusedCta &w)
{
    switch (model) {
    case C3SC_MODEL_DOUBLE_INT:
        switch (dx) {
        case 2: fused_walk_m<DoubleInt<2>>(c, w); break;
        case 3: fused_walk_m<DoubleInt<3>>(c, w); break;
        case 4: fused_walk_m<DoubleInt<4>>(c, w); break;
        }
        break;
    case C3SC_MODEL_DUBINS: fused_walk_m<Dubins>(c, w); break;
    case C3SC_MODEL_SKID5D: fused_walk_m<Skid5d>(c, w); break;
#ifdef C3SC_USER_MODEL_HEADER
    case C3SC_MODEL_USER: fused_walk_m<UserModel>(c, w); break;
#endif
    }
}
// control dimensions of the USER model's candidate table (0: not separable, no table); -1: no user model compiled in
int user_model_nud()
{
#ifdef C3SC_USER_MODEL_HEADER
    return UserModel::SEP ? UserModel::NUD : 0;
#else
    return -1;
#endif
}
int user_model_dims(int *dx, int *du)
{
#ifdef C3SC_USER_MODEL_HEADER
    *dx = UserModel::DX; *du = UserModel::DU;
    return 0;
#else
    (void)dx; (void)du;
    return -1;
#endif
}
int fused_ok_misc(int model, int dx, int arith, const CtlArgs &c, int pi_eval)
{
    switch (model) {
    case C3SC_MODEL_DOUBLE_INT:
        switch (dx) {
        case 2: return fused_ok_m<DoubleInt<2>>(arith, c, pi_eval);
        case 3: return fused_ok_m<DoubleInt<3>>(arith, c, pi_eval);
        case 4: return fused_ok_m<DoubleInt<4>>(arith, c, pi_eval);
        }
        return 0;
    case C3SC_MODEL_DUBINS: return dx == 3 ? fused_ok_m<Dubins>(arith, c, pi_eval) : 0;
    case C3SC_MODEL_SKID5D: return dx == 5 ? fused_ok_m<Skid5d>(arith, c, pi_eval) : 0;
#ifdef C3SC_USER_MODEL_HEADER
    case C3SC_MODEL_USER: return dx == UserModel::DX ? fused_ok_m<UserModel>(arith, c, pi_eval) : 0;
#endif
    }
    return 0;
}
int launch_control_misc(int model, int dx, int arith, const CtlArgs &a, int pi_eval, cudaStream_t st)
{
    switch (model) {
    case C3SC_MODEL_DOUBLE_INT:
        switch (dx) {
        case 2: return launch_control_m<DoubleInt<2>>(arith, a, pi_eval, st);
        case 3: return launch_control_m<DoubleInt<3>>(arith, a, pi_eval, st);
        case 4: return launch_control_m<DoubleInt<4>>(arith, a, pi_eval, st);
        }
        return -1;
    case C3SC_MODEL_DUBINS: return dx == 3 ? launch_control_m<Dubins>(arith, a, pi_eval, st) : -1;
    case C3SC_MODEL_SKID5D: return dx == 5 ? launch_control_m<Skid5d>(arith, a, pi_eval, st) : -1;
#ifdef C3SC_USER_MODEL_HEADER
    case C3SC_MODEL_USER: return dx == UserModel::DX ? launch_control_m<UserModel>(arith, a, pi_eval, st) : -1;
#endif
    }
    return -1;
}
int launch_model_eval_misc(int model, int dx, const DevProblem &P, int n, const double *x, const double *u,
                           double *drift, double *sig, double *stage, double *bound, double *obs, cudaStream_t st)
{
    switch (model) {
    case C3SC_MODEL_DOUBLE_INT:
        switch (dx) {
        case 2: return launch_model_eval_t<DoubleInt<2>>(P, n, x, u, drift, sig, stage, bound, obs, st);
        case 3: return launch_model_eval_t<DoubleInt<3>>(P, n, x, u, drift, sig, stage, bound, obs, st);
        case 4: return launch_model_eval_t<DoubleInt<4>>(P, n, x, u, drift, sig, stage, bound, obs, st);
        }
        return -1;
    case C3SC_MODEL_DUBINS: return dx == 3 ? launch_model_eval_t<Dubins>(P, n, x, u, drift, sig, stage, bound, obs, st) : -1;
    case C3SC_MODEL_SKID5D: return dx == 5 ? launch_model_eval_t<Skid5d>(P, n, x, u, drift, sig, stage, bound, obs, st) : -1;
#ifdef C3SC_USER_MODEL_HEADER
    case C3SC_MODEL_USER: return dx == UserModel::DX ? launch_model_eval_t<UserModel>(P, n, x, u, drift, sig, stage, bound, obs, st) : -1;
#endif
    }
    return -1;
}

int build_ctab_misc(int model, int dx, const DevProblem &P, double *ctab, cudaStream_t st)
{
    switch (model) {
    case C3SC_MODEL_DOUBLE_INT:
        switch (dx) {
        case 2: return build_ctab_t<DoubleInt<2>>(P, ctab, st);
        case 3: return build_ctab_t<DoubleInt<3>>(P, ctab, st);
        case 4: return build_ctab_t<DoubleInt<4>>(P, ctab, st);
        }
        return -1;
    case C3SC_MODEL_DUBINS: return dx == 3 ? build_ctab_t<Dubins>(P, ctab, st) : -1;
    case C3SC_MODEL_SKID5D: return dx == 5 ? 0 : -1;
#ifdef C3SC_USER_MODEL_HEADER
    case C3SC_MODEL_USER: return dx == UserModel::DX ? build_ctab_t<UserModel>(P, ctab, st) : -1;
#endif
    }
    return -1;
}

int launch_node_backup_misc(int model, int dx, int arith, const DevProblem &P, int n, const double *x, const double *costs,
                            const int *absorbed, double *value, int *argmin, cudaStream_t st)
{
    switch (model) {
    case C3SC_MODEL_DOUBLE_INT:
        switch (dx) {
        case 2: return launch_node_backup_t<DoubleInt<2>>(arith, P, n, x, costs, absorbed, value, argmin, st);
        case 3: return launch_node_backup_t<DoubleInt<3>>(arith, P, n, x, costs, absorbed, value, argmin, st);
        case 4: return launch_node_backup_t<DoubleInt<4>>(arith, P, n, x, costs, absorbed, value, argmin, st);
        }
        return -1;
    case C3SC_MODEL_DUBINS: return dx == 3 ? launch_node_backup_t<Dubins>(arith, P, n, x, costs, absorbed, value, argmin, st) : -1;
    case C3SC_MODEL_SKID5D: return dx == 5 ? launch_node_backup_t<Skid5d>(arith, P, n, x, costs, absorbed, value, argmin, st) : -1;
#ifdef C3SC_USER_MODEL_HEADER
    case C3SC_MODEL_USER: return dx == UserModel::DX ? launch_node_backup_t<UserModel>(arith, P, n, x, costs, absorbed, value, argmin, st) : -1;
#endif
    }
    return -1;
}

int launch_control_value_misc(int model, int dx, int arith, const DevProblem &P, int n, const double *x, const double *u,
                              const double *costs, double *value, cudaStream_t st)
{
    switch (model) {
    case C3SC_MODEL_DOUBLE_INT:
        switch (dx) {
        case 2: return launch_control_value_t<DoubleInt<2>>(arith, P, n, x, u, costs, value, st);
        case 3: return launch_control_value_t<DoubleInt<3>>(arith, P, n, x, u, costs, value, st);
        case 4: return launch_control_value_t<DoubleInt<4>>(arith, P, n, x, u, costs, value, st);
        }
        return -1;
    case C3SC_MODEL_DUBINS: return dx == 3 ? launch_control_value_t<Dubins>(arith, P, n, x, u, costs, value, st) : -1;
    case C3SC_MODEL_SKID5D: return dx == 5 ? launch_control_value_t<Skid5d>(arith, P, n, x, u, costs, value, st) : -1;
#ifdef C3SC_USER_MODEL_HEADER
    case C3SC_MODEL_USER: return dx == UserModel::DX ? launch_control_value_t<UserModel>(arith, P, n, x, u, costs, value, st) : -1;
#endif
    }
    return -1;
}
int launch_rhs(int arith, int dx, double beta, int n, const double *prob, const double *dt, const double *stage,
               const double *cost, double *out, void *stream)
{
    if (n <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int g = (n + 127) / 128;
    if (arith == C3SC_ARITH_EXACT) k_rhs<Exact><<<g, 128, 0, st>>>(dx, beta, n, prob, dt, stage, cost, out);
    else k_rhs<Fast><<<g, 128, 0, st>>>(dx, beta, n, prob, dt, stage, cost, out);
    return (int)cudaGetLastError();
}

template <int DX>
static int tr(int arith, const DevProblem &P, int n, const double *drift, const double *sig, double *prob,
              double *dt, int *status, cudaStream_t st)
{
    if (n <= 0) return 0;
    const int g = (n + 127) / 128;
    if (arith == C3SC_ARITH_EXACT) k_transition<DX, Exact><<<g, 128, 0, st>>>(P, n, drift, sig, prob, dt, status);
    else k_transition<DX, Fast><<<g, 128, 0, st>>>(P, n, drift, sig, prob, dt, status);
    return (int)cudaGetLastError();
}
int launch_transition(int arith, const DevProblem &P, int n, const double *drift, const double *sig,
                      double *prob, double *dt, int *status, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    switch (P.dx) {
    case 1: return tr<1>(arith, P, n, drift, sig, prob, dt, status, st);
    case 2: return tr<2>(arith, P, n, drift, sig, prob, dt, status, st);
    case 3: return tr<3>(arith, P, n, drift, sig, prob, dt, status, st);
    case 4: return tr<4>(arith, P, n, drift, sig, prob, dt, status, st);
    case 5: return tr<5>(arith, P, n, drift, sig, prob, dt, status, st);
    case 6: return tr<6>(arith, P, n, drift, sig, prob, dt, status, st);
    case 8: return tr<8>(arith, P, n, drift, sig, prob, dt, status, st);
    case 10: return tr<10>(arith, P, n, drift, sig, prob, dt, status, st);
    case 12: return tr<12>(arith, P, n, drift, sig, prob, dt, status, st);
    }
    return -1;
}
}  // namespace c3sc
