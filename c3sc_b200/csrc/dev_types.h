// dev_types.h -- POD structs passed BY VALUE to the kernels (constant bank).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/c3sc_b200.h"

namespace c3sc {

constexpr int MAXD = C3SC_MAXD;

// Device mirror of MCAparam + DPparam + Boundary + the c3opt brute-force table
// (reference src/bellman.c:118-285, src/boundary.c:353-489).
struct DevProblem {
    int dx, du, dw, nu, nobs, nmax;
    int ngrid[MAXD];
    int bc[MAXD];
    int xoff[MAXD];          // offset of xgrid[i] inside `xgrid`
    double h2, beta;
    double t[2 * MAXD];      // t[2i]=h2/h_i, t[2i+1]=h2/h_i^2 (bellman.c:181-186)
    double mp[8];            // model parameters
    const double *xgrid;     // device, concatenated
    const double *obs;       // device, [nobs][2][dx]: lb then ub
    const double *utab;      // device, [nu*du]
    const double *ctab;      // device, candidate table of a separable model (FAST), see k_build_ctab
    const double *gtab;      // device, the same rows regrouped by normaliser share: [W.., h2*gu, table index]; NULL = not grouped
    double amin;             // min over candidates of the table's normaliser share
    int *err;                // device error word (norm < 1e-14 seen)
};

// Device mirror of ValueF::cores (reference src/valuefunc.c:62-78,165-189).
struct DevFT {
    int d, rmax;
    int n[MAXD];
    int r[MAXD + 1];
    long long off[MAXD];     // offset of core k inside `base` (doubles)
    const double *base;
    const double *baseT;     // same offsets, every block transposed (b + a*r_{k+1}); see k_pack_cores
    // zero-padded tile copy for the tensor-core node kernel (ranks <= 32, else NULL): block j of core k
    // at baseP + offP[k] + j*cpp[k]*ldp[k], element (a,b) at b*ldp[k] + a, rows a >= r_k / cols b >= r_{k+1} zero
    const double *baseP;
    long long offP[MAXD];
    int ldp[MAXD], cpp[MAXD];
    // compact tile copy for the node kernel's TMA: block j of core k at baseQ + offQ[k] + j*ldq[k]*r_{k+1},
    // element (a,b) at b*ldq[k] + a with ldq = r_k rounded up to even, so every block starts 16-byte aligned.
    // Fragment reads past r_k / r_{k+1} land in neighbouring finite data and meet zero operands.
    const double *baseQ;
    long long offQ[MAXD];
    int ldq[MAXD];
};

struct DevOut {
    double *value;
    int *argmin;
    int *absorbed;
    double *costs;
    double *rows;
    int *nbr_vary;
    int *nbr_fixed;
};

// Per-device caches (function attributes are per device; one process may drive several GPUs, one host thread each)
constexpr int C3SC_MAXDEV = 16;
inline int c3sc_cur_dev()
{
    int dv = 0;
    cudaGetDevice(&dv);
    return dv < 0 ? 0 : (dv >= C3SC_MAXDEV ? C3SC_MAXDEV - 1 : dv);
}

// neighbour pair of node j ALONG the fiber (nodeutil.c:570-624); ab = the node's flag after the
// end-node overwrite (0 / 1 / -1).  Interior absorbed nodes point at themselves.
__host__ __device__ inline void ft_vary_pair(int bk, int N, int j, int ab, int &lo, int &hi)
{
    lo = j - 1; hi = j + 1;
    if (j == 0) {
        if (bk == C3SC_ABSORB)       { lo = 0; hi = 0; }
        else if (bk == C3SC_REFLECT) { lo = 0; hi = 1; }
        else                         { lo = N - 2; hi = 1; }
    } else if (j == N - 1) {
        if (bk == C3SC_ABSORB)       { lo = N - 1; hi = N - 1; }
        else if (bk == C3SC_REFLECT) { lo = N - 2; hi = N - 1; }
        else                         { lo = N - 2; hi = 1; }
    } else if (ab != 0) { lo = j; hi = j; }
}

// implemented in inst_misc.cu; returns cudaError_t as int, or -1 if dx is not instantiated
int launch_transition(int arith, const DevProblem &P, int n, const double *drift, const double *sig,
                      double *prob, double *dt, int *status, void *stream);
// bellmanrhs (bellman.c:88-112) for n independent (prob[2dx+1], dt, stage, cost[2dx+1]) tuples
int launch_rhs(int arith, int dx, double beta, int n, const double *prob, const double *dt, const double *stage,
               const double *cost, double *out, void *stream);

}  // namespace c3sc
