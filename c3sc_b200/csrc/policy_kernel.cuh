// policy_kernel.cuh -- the implicit policy at arbitrary (off-grid) states: SURVEY §8(f)-4.
//
//   valuef_eval                  src/valuefunc.c:345-350 -> C3 function_train_eval on LINELM cores
//                                = piecewise-linear interpolation of the nodal cores, 0 outside the grid
//   mca_get_neighbor_node_costs  src/nodeutil.c:718-816   V at x -+ h e_i, boundary type decides the
//                                stand-in when the step leaves the grid; inside an obstacle all = V(x)
// The backup itself is k_node_backup (control_kernel.cuh) on these neighbour values
// (c3control_policy_eval, src/bellman.c:2105-2151).  Latency-oriented, model independent.
#pragma once
#include <cuda_runtime.h>
#include "dev_types.h"

namespace c3sc {

// one thread per state: flag + the 2d+1 evaluation points [n][2d+1][d]
__global__ void k_policy_points(const DevProblem P, int n, const double *x, double *pts, int *absorbed)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int d = P.dx, np = 2 * d + 1;
    const double *xe = x + (size_t)e * d;
    double *pe = pts + (size_t)e * np * d;
    int ab = 0;
    for (int o = 0; o < P.nobs && ab == 0; o++) {                       // boundary.c:329-344,668-680
        const double *lb = P.obs + (size_t)o * 2 * d, *ub = lb + d;
        bool inside = true;
        for (int i = 0; i < d; i++) inside = inside && !(xe[i] < lb[i] || xe[i] > ub[i]);
        if (inside) ab = -1;
    }
    absorbed[e] = ab;
    for (int m = 0; m < np; m++)
        for (int i = 0; i < d; i++) pe[(size_t)m * d + i] = xe[i];
    if (ab) return;                                                     // every entry is V(x), nodeutil.c:726-733
    for (int i = 0; i < d; i++) {
        const double *g = P.xgrid + P.xoff[i];
        const double lb = g[0], ub = g[P.ngrid[i] - 1], h = g[1] - g[0], xi = xe[i];
        double lo, hi;
        if (((xi + h) < ub) && (xi - h > lb)) { lo = xi - h; hi = xi + h; }                 // :741-747
        else if ((xi - h) <= lb) {                                                          // :748-775
            hi = xi + h;
            if (P.bc[i] == C3SC_ABSORB || P.bc[i] == C3SC_REFLECT) lo = lb;
            else if (xi > lb) lo = ub - (h - (xi - lb));
            else lo = (ub - (lb - xi)) - h;
        } else {                                                                            // :776-806
            lo = xi - h;
            if (P.bc[i] == C3SC_ABSORB || P.bc[i] == C3SC_REFLECT) hi = ub;
            else if (xi < ub) hi = lb + (h - (ub - xi));
            else hi = (lb + (xi - ub)) + h;
        }
        pe[(size_t)(2 * i) * d + i] = lo;
        pe[(size_t)(2 * i + 1) * d + i] = hi;
    }
}

// one warp per point: v <- v * ((1-th) G_k[j] + th G_k[j+1]) core by core, lanes over the output rank index
// (transposed core copy: coalesced).  Dynamic shared memory: 2 * rmax doubles per warp.
__global__ void __launch_bounds__(256) k_ft_eval_points(const DevProblem P, const DevFT ft, int npts, const double *pts,
                                                        double *out)
{
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int d = ft.d;
    int rs = 1;
    for (int i = 0; i <= d; i++) rs = ft.r[i] > rs ? ft.r[i] : rs;
    double *v = smem + warp * 2 * rs, *w = v + rs;
    for (int pt = blockIdx.x * nw + warp; pt < npts; pt += gridDim.x * nw) {
        bool outside = false;
        __syncwarp();
        if (lane == 0) v[0] = 1.0;
        __syncwarp();
        for (int k = 0; k < d; k++) {
            const double xk = pts[(size_t)pt * d + k];
            const double *g = P.xgrid + P.xoff[k];
            const int n = P.ngrid[k];
            int lo = 1, hi = n;                                    // count of i in [1, n-1] with g[i] < xk
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (g[mid] < xk) lo = mid + 1; else hi = mid; }
            const int cnt = lo - 1, j = cnt < n - 2 ? cnt : n - 2;
            if (xk < g[0] || xk > g[n - 1]) outside = true;
            const double th = (xk - g[j]) / (g[j + 1] - g[j]);
            const int r0 = ft.r[k], r1 = ft.r[k + 1], blk = r0 * r1;
            const double *b0 = ft.baseT + ft.off[k] + (size_t)j * blk, *b1 = b0 + blk;
            for (int b = lane; b < r1; b += 32) {
                double acc = 0.0;
                for (int a = 0; a < r0; a++) {
                    const double gv = __dadd_rn(__dmul_rn(1.0 - th, b0[b + a * r1]), __dmul_rn(th, b1[b + a * r1]));
                    acc = __dadd_rn(acc, __dmul_rn(gv, v[a]));
                }
                w[b] = acc;
            }
            __syncwarp();
            double *t = v; v = w; w = t;
        }
        if (lane == 0) out[pt] = outside ? 0.0 : v[0];
    }
}

}  // namespace c3sc
