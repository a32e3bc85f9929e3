// models.cuh -- device-resident dynamics / cost models (enum c3sc_model).
//
// The reference evaluates host callbacks (src/dynamics.c:127-139,224-239;
// src/bellman.c:428-467).  Each struct below restates one example's callbacks
// with the example's operation order (A = arithmetic policy: Exact keeps every
// rounding of the CPU build).  Only the DIAGONAL of the diffusion is produced
// because that is all transition_assemble reads (src/nodeutil.c:294).
//
// u_dep(i): does drift_i or sigma_ii depend on the control?  Dimensions with
// u_dep==false are evaluated once per node instead of once per candidate.
//
// Separable models (SEP): for the NUD control-dependent dimensions ud(m) the
// drift depends on u ONLY, sigma does not depend on u, and the stage cost
// splits as stage_x(x) + stage_u(u).  Then everything control-dependent is a
// property of the candidate alone and is tabulated once per problem
// (k_build_ctab); FAST arithmetic uses the table, EXACT never does.
#pragma once
#include "arith.cuh"

namespace c3sc {

// examples/lqgnd/lqgnd.c:80-186 (chain of double integrators; == lqg2d.c:71-142 for DX=2)
// mp = [ss0, ss1, boundcost, obscost]
template <int DX_>
struct LqgNd {
    static constexpr int DX = DX_, DU = DX_ / 2, ID = C3SC_MODEL_LQGND;
    static constexpr bool SEP = true;            // see "separable models" below
    static constexpr int NUD = DX_ / 2;
    __host__ __device__ static constexpr bool u_dep(int i) { return (i & 1) != 0; }
    __host__ __device__ static constexpr int ud(int m) { return 2 * m + 1; }
    template <class A>
    __device__ __forceinline__ static void drift(const double *x, const double *u, const double *, double *b)
    {
#pragma unroll
        for (int i = 0; i < DX; i++) b[i] = (i & 1) ? u[i >> 1] : x[i + 1];     // lqgnd.c:86-95
    }
    template <class A>
    __device__ __forceinline__ static void sigma(const double *, const double *, const double *mp, double *s)
    {
#pragma unroll
        for (int i = 0; i < DX; i++) s[i] = (i & 1) ? mp[1] : mp[0];            // lqgnd.c:122-129
    }
    template <class A>
    __device__ __forceinline__ static double stage(const double *x, const double *u, const double *)
    {
        double g = 0.0;                                                          // lqgnd.c:146-159
#pragma unroll
        for (int i = 0; i < DX; i++) g = A::mad(x[i], x[i], g);
#pragma unroll
        for (int i = 0; i < DU; i++) g = A::mad(u[i], u[i], g);
        return g;
    }
    __device__ __forceinline__ static double stage_x(const double *x, const double *)
    {
        double g = 0.0;
#pragma unroll
        for (int i = 0; i < DX; i++) g = fma(x[i], x[i], g);
        return g;
    }
    __device__ __forceinline__ static double stage_u(const double *u, const double *)
    {
        double g = 0.0;
#pragma unroll
        for (int i = 0; i < DU; i++) g = fma(u[i], u[i], g);
        return g;
    }
    __device__ __forceinline__ static double boundcost(const double *, const double *mp) { return mp[2]; }
    __device__ __forceinline__ static double obscost(const double *, const double *mp) { return mp[3]; }
};

// examples/double_int/double_int.c:80-162; mp = [ss0, ss1, boundcost, obscost]
template <int DX_>
struct DoubleInt {
    static constexpr int DX = DX_, DU = 1, ID = C3SC_MODEL_DOUBLE_INT;
    static constexpr bool SEP = true;
    static constexpr int NUD = 1;
    __host__ __device__ static constexpr bool u_dep(int i) { return i == DX_ - 1; }
    __host__ __device__ static constexpr int ud(int) { return DX_ - 1; }
    __device__ __forceinline__ static double stage_x(const double *, const double *) { return 1.0; }
    __device__ __forceinline__ static double stage_u(const double *, const double *) { return 0.0; }
    template <class A>
    __device__ __forceinline__ static void drift(const double *x, const double *u, const double *, double *b)
    {
#pragma unroll
        for (int i = 0; i < DX - 1; i++) b[i] = x[i + 1];
        b[DX - 1] = u[0];
    }
    template <class A>
    __device__ __forceinline__ static void sigma(const double *, const double *, const double *mp, double *s)
    {
#pragma unroll
        for (int i = 0; i < DX - 1; i++) s[i] = mp[0];
        s[DX - 1] = mp[1];
    }
    template <class A>
    __device__ __forceinline__ static double stage(const double *, const double *, const double *) { return 1.0; }
    __device__ __forceinline__ static double boundcost(const double *, const double *mp) { return mp[2]; }
    __device__ __forceinline__ static double obscost(const double *, const double *mp) { return mp[3]; }
};

// examples/dubinscar_new/dubinscar.c:40-121; mp = [s_xy, s_theta, stage, boundcost, obscost]
struct Dubins {
    static constexpr int DX = 3, DU = 1, ID = C3SC_MODEL_DUBINS;
    static constexpr bool SEP = true;
    static constexpr int NUD = 1;
    __host__ __device__ static constexpr bool u_dep(int i) { return i == 2; }
    __host__ __device__ static constexpr int ud(int) { return 2; }
    __device__ __forceinline__ static double stage_x(const double *, const double *mp) { return mp[2]; }
    __device__ __forceinline__ static double stage_u(const double *, const double *) { return 0.0; }
    template <class A>
    __device__ __forceinline__ static void drift(const double *x, const double *u, const double *, double *b)
    {
        b[0] = cos(x[2]);
        b[1] = sin(x[2]);
        b[2] = u[0];
    }
    template <class A>
    __device__ __forceinline__ static void sigma(const double *, const double *, const double *mp, double *s)
    {
        s[0] = mp[0]; s[1] = mp[0]; s[2] = mp[1];
    }
    template <class A>
    __device__ __forceinline__ static double stage(const double *, const double *, const double *mp) { return mp[2]; }
    __device__ __forceinline__ static double boundcost(const double *, const double *mp) { return mp[3]; }
    __device__ __forceinline__ static double obscost(const double *, const double *mp) { return mp[4]; }
};

// examples/skidding5d/scar.c:39-176 (order = {0,1,2,3,4}); mp = [obscost]
struct Skid5d {
    static constexpr int DX = 5, DU = 1, ID = C3SC_MODEL_SKID5D;
    static constexpr bool SEP = false;           // drift_3, drift_4 mix x and u
    static constexpr int NUD = 2;
    __host__ __device__ static constexpr bool u_dep(int i) { return i >= 3; }
    __host__ __device__ static constexpr int ud(int m) { return 3 + m; }
    __device__ __forceinline__ static double stage_x(const double *, const double *) { return 0.0; }
    __device__ __forceinline__ static double stage_u(const double *, const double *) { return 0.0; }
    template <class A>
    __device__ __forceinline__ static void drift(const double *x, const double *u, const double *, double *b)
    {
        const double orient = x[2], angvel = x[3], speed = x[4], steering = u[0];
        const double m = 1460.0, cf = 17000.0, ct = 20000.0, a1 = 1.2, b1 = 1.5, In = 2170.0, s = 27.0;
        const double co = cos(orient), so = sin(orient);
        const double ff = A::mul(cf, A::add(A::div(A::add(speed, A::mul(a1, angvel)), s), steering));   // scar.c:78
        const double ft = A::div(A::mul(ct, A::sub(speed, A::mul(b1, angvel))), s);                       // scar.c:79
        b[0] = A::sub(A::mul(s, co), A::mul(speed, so));
        b[1] = A::add(A::mul(s, so), A::mul(speed, co));
        b[2] = angvel;
        b[3] = A::div(A::sub(A::mul(a1, ff), A::mul(b1, ft)), In);
        b[4] = A::add(A::mul(-s, angvel), A::div(A::add(ff, ft), m));
    }
    template <class A>
    __device__ __forceinline__ static void sigma(const double *, const double *, const double *, double *s)
    {
        // scar.c:118-130: the example's 4th entry lands outside the 5x5 block, so the
        // diagonal transition_assemble sees is (1e-5,1e-5,1e-5,0,1e-5).
        s[0] = 1e-5; s[1] = 1e-5; s[2] = 1e-5; s[3] = 0.0; s[4] = 1e-5;
    }
    template <class A>
    __device__ __forceinline__ static double stage(const double *x, const double *, const double *)
    {
        double g = A::add(A::add(1.0, A::mul(0.02, A::mul(x[0], x[0]))), A::mul(0.02, A::mul(x[1], x[1])));  // scar.c:149
        g = A::add(A::add(g, A::mul(x[3], x[3])), A::mul(x[4], x[4]));                                        // scar.c:150
        return g;
    }
    __device__ __forceinline__ static double boundcost(const double *x, const double *)
    {
        double g = __dadd_rn(__dmul_rn(0.1, __dmul_rn(x[0], x[0])), __dmul_rn(0.1, __dmul_rn(x[1], x[1])));     // scar.c:165
        g = __dadd_rn(__dadd_rn(g, __dmul_rn(0.1, __dmul_rn(x[3], x[3]))), __dmul_rn(0.1, __dmul_rn(x[4], x[4]))); // scar.c:166
        return g;
    }
    __device__ __forceinline__ static double obscost(const double *, const double *mp) { return mp[0]; }
};

}  // namespace c3sc

// The USER model slot (C3SC_MODEL_USER): a header named at build time (make USER_MODEL=...; default
// examples/user_model_vdp.cuh) that defines struct c3sc::UserModel with the interface above.
#ifdef C3SC_USER_MODEL_HEADER
#include C3SC_USER_MODEL_HEADER
#endif
