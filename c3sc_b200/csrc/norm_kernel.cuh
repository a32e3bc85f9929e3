// norm_kernel.cuh -- valuef_norm / valuef_norm2diff (src/valuefunc.c:315-335 -> C3 function_train_norm2 / norm2diff on LINELM
// cores) next to the cores: the CONTINUOUS L2 inner product of two piecewise-linear function trains on the device.
//
//   <f, g> = integral over the box of f g;  per dimension the nodal values meet through the mass matrix of the hat functions,
//   M_ii = (h_{i-1} + h_i)/3, M_{i,i+1} = h_i/6, so the train contraction is
//       Z_{k+1} = sum_i  X_k[i]^T  Z_k  ( M_ii Y_k[i] + M_{i,i-1} Y_k[i-1] + M_{i,i+1} Y_k[i+1] ).
//   One launch per dimension; CTA (g, p) takes nodes g, g + G, .. of pair p (the three pairs of a norm2diff -- <a,a>, <a,b>,
//   <b,b> -- share the launches) and writes its partial Z_{k+1}; the next launch starts by summing the G partials in a FIXED
//   order.  Same formula as the host restatement c3sc_cores_dot_l2 (c3sc_cross.c), which the GPU test compares against.
#pragma once
#include "dev_types.h"

namespace c3sc {

constexpr int NRM_NT = 256;      // threads per CTA
constexpr int NRM_G = 32;        // CTAs (node slices) per pair

struct NormTrain { const double *base; long long off[MAXD]; int r[MAXD + 1]; };
struct NormArgs {
    NormTrain t[2];
    int n[MAXD], xoff[MAXD];     // nodes per dimension, start of the dimension's node coordinates in x
    const double *x;
    int d, npairs;               // 1: <t0, t1>;  3: <t0,t0>, <t0,t1>, <t1,t1>
    int zstride;                 // doubles per (pair, CTA) partial
};

// zin / zout: [npairs][NRM_G][zstride]
__global__ void __launch_bounds__(NRM_NT) k_train_dot_l2(const NormArgs a, int k, const double *zin, double *zout)
{
    extern __shared__ double nsm[];
    const int p = blockIdx.y, tid = threadIdx.x;
    const NormTrain &X = a.t[a.npairs == 1 ? 0 : (p == 2 ? 1 : 0)], &Y = a.t[a.npairs == 1 ? 1 : (p == 0 ? 0 : 1)];
    const int x1 = X.r[k], x2 = X.r[k + 1], y1 = Y.r[k], y2 = Y.r[k + 1], N = a.n[k];
    double *Z = nsm, *Bw = Z + x1 * y1, *H = Bw + y1 * y2, *Zn = H + x1 * y2;
    for (int e = tid; e < x1 * y1; e += NRM_NT) {
        double s = 0.0;
        if (k == 0) s = 1.0;
        else {
            const double *zi = zin + (size_t)p * NRM_G * a.zstride + e;
            for (int g = 0; g < NRM_G; g++) s += zi[(size_t)g * a.zstride];
        }
        Z[e] = s;
    }
    for (int e = tid; e < x2 * y2; e += NRM_NT) Zn[e] = 0.0;
    const double *xs = a.x + a.xoff[k];
    const double *Xk = X.base + X.off[k], *Yk = Y.base + Y.off[k];
    for (int i = blockIdx.x; i < N; i += NRM_G) {
        const double hl = i > 0 ? xs[i] - xs[i - 1] : 0.0, hr = i + 1 < N ? xs[i + 1] - xs[i] : 0.0;
        const double mc = (hl + hr) / 3.0, ml = hl / 6.0, mr = hr / 6.0;
        const double *Yc = Yk + (size_t)i * y1 * y2, *Yl = i > 0 ? Yc - y1 * y2 : Yc, *Yr = i + 1 < N ? Yc + y1 * y2 : Yc;
        __syncthreads();                                    // Z ready / the previous node's H consumed
        for (int e = tid; e < y1 * y2; e += NRM_NT) Bw[e] = mc * Yc[e] + ml * Yl[e] + mr * Yr[e];
        __syncthreads();
        for (int e = tid; e < x1 * y2; e += NRM_NT) {        // H[a, c] = sum_b Z[a, b] Bw[b, c]
            const int aa = e % x1, c = e / x1;
            double s = 0.0;
            for (int b = 0; b < y1; b++) s += Z[aa + b * x1] * Bw[b + c * y1];
            H[e] = s;
        }
        __syncthreads();
        const double *Xi = Xk + (size_t)i * x1 * x2;
        for (int e = tid; e < x2 * y2; e += NRM_NT) {        // Zn[q, c] += sum_a X_i[a, q] H[a, c]   (a thread owns its entries)
            const int q = e % x2, c = e / x2;
            double s = 0.0;
            for (int aa = 0; aa < x1; aa++) s += Xi[aa + q * x1] * H[aa + c * x1];
            Zn[e] += s;
        }
    }
    __syncthreads();
    double *zo = zout + ((size_t)p * NRM_G + blockIdx.x) * a.zstride;
    for (int e = tid; e < x2 * y2; e += NRM_NT) zo[e] = Zn[e];
}

}  // namespace c3sc
