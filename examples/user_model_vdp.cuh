// user_model_vdp.cuh -- the example USER device model: how a problem that is not one of the reference's four
// examples gets onto the device ("Adding a device model").
//
// The reference accepts ANY host callback (src/dynamics.c:107, src/bellman.c:215-217); a kernel cannot call one, so
// a new problem states its callbacks once more as a struct like this.  The library is rebuilt with
//     make -C c3sc_b200/csrc USER_MODEL=path/to/this_header.cuh
// and the model is selected as C3SC_MODEL_USER (id 5) next to the unchanged host pointers
// (dp_param_set_device_model / c3control_set_device_model).  dp_param_check_device_model compares the two on grid nodes.
//
// A controlled Van der Pol oscillator with additive noise:
//     dx0 = x1 dt + s0 dW0,     dx1 = (mu (1 - x0^2) x1 - x0 + u) dt + s1 dW1,
//     stage cost x0^2 + x1^2 + u^2,  mp = [mu, s0, s1, boundcost, obscost]   (defaults in api.cu: 1, 0.5, 0.5, 50, 0)
//
// What a model struct provides (see models.cuh for the built-in ones):
//   DX, DU              state and control dimension
//   u_dep(i)            does drift_i (or sigma_ii) depend on the control?  false: evaluated once per node
//   SEP, NUD, ud(m)     SEP = true promises that the NUD control-dependent dimensions ud(m) have a drift that depends on u
//                       ONLY, a control-independent sigma and a stage cost stage_x(x) + stage_u(u): then stage 2 tabulates
//                       the candidates once per problem.  Here drift_1 mixes x and u: SEP = false, the general walk
//   drift<A>, sigma<A>  b(x,u) and the DIAGONAL of the diffusion (all transition_assemble reads, src/nodeutil.c:294);
//   stage<A>            A = arithmetic policy: A::mul / add / sub / div / mad keep the host's operation order without
//                       FMA contraction in EXACT mode -- write them in the order of the host callback
//   stage_x, stage_u    the split stage cost (only used when SEP)
//   boundcost, obscost  cost of absorbing-boundary and obstacle nodes
#pragma once

namespace c3sc {

struct UserModel {
    static constexpr int DX = 2, DU = 1, ID = C3SC_MODEL_USER;
    static constexpr bool SEP = false;
    static constexpr int NUD = 1;
    __host__ __device__ static constexpr bool u_dep(int i) { return i == 1; }
    __host__ __device__ static constexpr int ud(int) { return 1; }
    template <class A>
    __device__ __forceinline__ static void drift(const double *x, const double *u, const double *mp, double *b)
    {
        b[0] = x[1];
        // mu * (1 - x0*x0) * x1 - x0 + u, left to right like the host callback
        b[1] = A::add(A::sub(A::mul(A::mul(mp[0], A::sub(1.0, A::mul(x[0], x[0]))), x[1]), x[0]), u[0]);
    }
    template <class A>
    __device__ __forceinline__ static void sigma(const double *, const double *, const double *mp, double *s)
    {
        s[0] = mp[1]; s[1] = mp[2];
    }
    template <class A>
    __device__ __forceinline__ static double stage(const double *x, const double *u, const double *)
    {
        return A::add(A::add(A::mul(x[0], x[0]), A::mul(x[1], x[1])), A::mul(u[0], u[0]));
    }
    __device__ __forceinline__ static double stage_x(const double *x, const double *) { return x[0] * x[0] + x[1] * x[1]; }
    __device__ __forceinline__ static double stage_u(const double *u, const double *) { return u[0] * u[0]; }
    __device__ __forceinline__ static double boundcost(const double *, const double *mp) { return mp[3]; }
    __device__ __forceinline__ static double obscost(const double *, const double *mp) { return mp[4]; }
};

}  // namespace c3sc
