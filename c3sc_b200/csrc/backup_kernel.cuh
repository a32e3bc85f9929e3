// backup_kernel.cuh -- the fused Bellman-backup kernel (one CTA walks fibers).
//
// Per fiber (dim_vary k, fixed indices f) a CTA does, entirely on chip:
//   1. flags + neighbour indices      process_fibers_neighbor  src/nodeutil.c:489-627
//   2. FT values at every node and its 2d axis neighbours
//                                      valuef_eval_fiber_ind_nn src/valuefunc.c:369-585
//      (re-associated: fiber-constant prefix/suffix/neighbour vectors are built once,
//       each node then costs two r x r mat-vecs + (2d+1) length-r dots)
//   3. per node, min over the control table of
//        dt*g + exp(-beta*dt) * <p, V_nbr>       bellman_control/bellmanrhs
//                                                src/bellman.c:88-112,367-480,504-543
//      with p, dt from the upwind Kushner-Dupuis construction
//                                                transition_assemble src/nodeutil.c:267-406
//   4. absorbed nodes take boundcost / obscost   src/bellman.c:513-532
// Only the N values (and optional diagnostics) go back to HBM.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include "dev_types.h"
#include "arith.cuh"
#include "models.cuh"

namespace c3sc {

constexpr int NT = 256;              // threads per CTA
constexpr int NW = NT / 32;          // warps per CTA

__host__ __device__ inline int odd_up(int v) { return v | 1; }

// ---------------------------------------------------------------------------
// shared-memory carve-up (all offsets in doubles / ints); identical on host+device
struct SmemPlan {
    int rs;        // padded (odd) vector stride
    int cs;        // cost-row stride (2dx+1, odd already)
    int oL, oR, oNb, oTmp, oW, oU, oC, oV, oUtab, nDoubles;
    int oAbs, oNv, oAct, oNf, oFix, oMisc, nInts;
    __host__ __device__ SmemPlan(int dx, int du, int nu, int nmax, int rmax)
    {
        rs = odd_up(rmax);
        cs = 2 * dx + 1;
        int o = 0;
        oL = o;    o += (dx + 1) * rs;
        oR = o;    o += (dx + 1) * rs;
        oNb = o;   o += 2 * dx * rs;
        oTmp = o;  o += NW * 2 * rs;
        oW = o;    o += nmax * rs;
        oU = o;    o += nmax * rs;
        oC = o;    o += nmax * cs;
        oV = o;    o += nmax;
        oUtab = o; o += nu * du;
        nDoubles = o;
        int q = 0;
        oAbs = q;  q += nmax;
        oNv = q;   q += 2 * nmax;
        oAct = q;  q += nmax;
        oNf = q;   q += 2 * dx;
        oFix = q;  q += dx;
        oMisc = q; q += 4;
        nInts = q;
    }
    __host__ __device__ size_t bytes() const { return (size_t)nDoubles * 8 + (size_t)nInts * 4; }
};

// ---------------------------------------------------------------------------
// transition_assemble (src/nodeutil.c:284-309,365-371,396-402), all dims in order.
// prob = [pl_0, pr_0, ..., pl_{d-1}, pr_{d-1}, pself]; returns 0 or 1 (norm < 1e-14).
template <int DX, class A>
__device__ __forceinline__ int transition_row(const DevProblem &P, const double *b, const double *s,
                                              double *prob, double &dt)
{
    double norm = 0.0;
#pragma unroll
    for (int i = 0; i < DX; i++) {
        const double s2 = A::mul(s[i], s[i]);
        const double q = A::mul(A::mul(P.t[2 * i + 1], s2), 0.5);     // t2*diff/2.0
        double pl = q, pr = q;
        if (b[i] < -1e-14)      pl = A::sub(pl, A::mul(P.t[2 * i], b[i]));
        else if (b[i] > 1e-14)  pr = A::add(pr, A::mul(P.t[2 * i], b[i]));
        prob[2 * i] = pl;
        prob[2 * i + 1] = pr;
        norm = A::add(norm, pl);
        norm = A::add(norm, pr);
    }
    if (norm < 1e-14) { dt = 0.0; return 1; }
    dt = A::div(P.h2, norm);
    double ps = 1.0;
#pragma unroll
    for (int i = 0; i < DX; i++) {
        prob[2 * i] = A::div(prob[2 * i], norm);
        prob[2 * i + 1] = A::div(prob[2 * i + 1], norm);
        ps = A::sub(ps, prob[2 * i]);
        ps = A::sub(ps, prob[2 * i + 1]);
    }
    prob[2 * DX] = ps;
    return 0;
}

// bellmanrhs (src/bellman.c:88-112): dt*g + exp(-beta*dt) * <p, c>, sequential dot.
template <int DX, class A>
__device__ __forceinline__ double rhs(const DevProblem &P, const double *prob, double dt, double g,
                                      const double *c)
{
    const double ebt = exp(A::mul(-P.beta, dt));
    double ctg = 0.0;
#pragma unroll
    for (int m = 0; m < 2 * DX + 1; m++) ctg = A::mad(prob[m], c[m], ctg);
    return A::add(A::mul(dt, g), A::mul(ebt, ctg));
}

// Per-node state hoisted out of the candidate loop.
template <class M>
struct NodeInv {
    double b0[M::DX], s0[M::DX];   // drift / sigma at the first candidate (valid for !u_dep dims)
    double norm0, S0;              // Fast: partial normaliser and partial <raw, c>
};

template <class M, class A>
__device__ __forceinline__ void node_prepare(const DevProblem &P, const double *utab, const double *x,
                                             const double *c, NodeInv<M> &inv)
{
    constexpr int DX = M::DX;
    double u0[M::DU];
#pragma unroll
    for (int i = 0; i < M::DU; i++) u0[i] = utab[i];
    M::template drift<A>(x, u0, P.mp, inv.b0);
    M::template sigma<A>(x, u0, P.mp, inv.s0);
    inv.norm0 = 0.0;
    inv.S0 = 0.0;
    if (!A::exact) {
#pragma unroll
        for (int i = 0; i < DX; i++) {
            if (M::u_dep(i)) continue;
            const double q = (P.t[2 * i + 1] * 0.5) * (inv.s0[i] * inv.s0[i]);
            const double tb = P.t[2 * i] * inv.b0[i];
            const double rl = q - ((inv.b0[i] < -1e-14) ? tb : 0.0);
            const double rr = q + ((inv.b0[i] > 1e-14) ? tb : 0.0);
            inv.norm0 += rl + rr;
            inv.S0 = fma(rl, c[2 * i], inv.S0);
            inv.S0 = fma(rr, c[2 * i + 1], inv.S0);
        }
    }
}

// value of one candidate control (bellman_control, src/bellman.c:367-480, absorbed==0 branch)
template <class M, class A>
__device__ __forceinline__ double candidate_value(const DevProblem &P, const double *x, const double *u,
                                                  const double *c, const NodeInv<M> &inv, int &bad)
{
    constexpr int DX = M::DX;
    double b[DX], s[DX];
    M::template drift<A>(x, u, P.mp, b);
    M::template sigma<A>(x, u, P.mp, s);
    const double g = M::template stage<A>(x, u, P.mp);
    if (A::exact) {
#pragma unroll
        for (int i = 0; i < DX; i++)
            if (!M::u_dep(i)) { b[i] = inv.b0[i]; s[i] = inv.s0[i]; }
        double prob[2 * DX + 1], dt;
        if (transition_row<DX, A>(P, b, s, prob, dt)) { bad = 1; return CUDART_INF; }
        return rhs<DX, A>(P, prob, dt, g, c);
    } else {
        double norm = inv.norm0, S = inv.S0;
#pragma unroll
        for (int i = 0; i < DX; i++) {
            if (!M::u_dep(i)) continue;
            const double q = (P.t[2 * i + 1] * 0.5) * (s[i] * s[i]);
            const double tb = P.t[2 * i] * b[i];
            const double rl = q - ((b[i] < -1e-14) ? tb : 0.0);
            const double rr = q + ((b[i] > 1e-14) ? tb : 0.0);
            norm += rl + rr;
            S = fma(rl, c[2 * i], S);
            S = fma(rr, c[2 * i + 1], S);
        }
        if (norm < 1e-14) { bad = 1; return CUDART_INF; }
        const double rinv = 1.0 / norm;
        const double dt = P.h2 * rinv;
        const double ebt = exp(-P.beta * dt);
        return fma(dt, g, ebt * (S * rinv));
    }
}

// ---------------------------------------------------------------------------
// y[b] = sum_a v[a] * G[a + b*m]   (row vector times column-major m x n block), one warp
__device__ __forceinline__ void warp_vecmat(int m, int n, const double *__restrict__ G,
                                            const double *v, double *y, int lane)
{
    for (int b = lane; b < n; b += 32) {
        const double *col = G + (size_t)b * m;
        double acc = 0.0;
        for (int a = 0; a < m; a++) acc = fma(v[a], __ldg(col + a), acc);
        y[b] = acc;
    }
}
// y[a] = sum_b G[a + b*m] * v[b], one warp
__device__ __forceinline__ void warp_matvec(int m, int n, const double *__restrict__ G,
                                            const double *v, double *y, int lane)
{
    for (int a = lane; a < m; a += 32) {
        double acc = 0.0;
        for (int b = 0; b < n; b++) acc = fma(__ldg(G + a + (size_t)b * m), v[b], acc);
        y[a] = acc;
    }
}

// ---------------------------------------------------------------------------
template <class M, class A>
__global__ void __launch_bounds__(NT) k_backup(const LaunchArgs a, const int G)
{
    constexpr int DX = M::DX, DU = M::DU, CS = 2 * DX + 1, RW = 2 * DX + 3;
    const DevProblem &P = a.P;
    const DevFT &ft = a.ft;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    extern __shared__ double smem[];
    const SmemPlan sp(DX, DU, P.nu, P.nmax, ft.rmax);
    double *sL = smem + sp.oL, *sR = smem + sp.oR, *sNb = smem + sp.oNb, *sTmp = smem + sp.oTmp;
    double *sW = smem + sp.oW, *sU = smem + sp.oU, *sC = smem + sp.oC, *sV = smem + sp.oV;
    double *sUtab = smem + sp.oUtab;
    int *ismem = reinterpret_cast<int *>(smem + sp.nDoubles);
    int *sAbs = ismem + sp.oAbs, *sNv = ismem + sp.oNv, *sAct = ismem + sp.oAct, *sNf = ismem + sp.oNf;
    int *sMisc = ismem + sp.oMisc, *sFix = ismem + sp.oFix;
    const int rs = sp.rs;

    for (int i = tid; i < P.nu * DU; i += NT) sUtab[i] = P.utab[i];

    for (int f = blockIdx.x; f < a.F; f += gridDim.x) {
        __syncthreads();                       // previous fiber fully consumed
        const int k = a.dim_vary[f];
        int fix[DX];
#pragma unroll
        for (int i = 0; i < DX; i++) fix[i] = a.fixed_ind[(size_t)f * DX + i];
        const int N = P.ngrid[k];
        const size_t obase = (size_t)f * a.ldo;
        if (tid == 0) sMisc[0] = 0;            // active-node counter
        if (tid < DX) sFix[tid] = a.fixed_ind[(size_t)f * DX + tid];   // dynamically indexed copy

        // ---- 1. flags and neighbour indices (nodeutil.c:489-627) -------------
        // fixed-dimension pairs + "wall" (a fixed index on an ABSORB face)
        bool wall = false;
        {
            int slot = 0;
#pragma unroll
            for (int i = 0; i < DX; i++) {
                if (i == k) continue;
                const int i0 = fix[i], last = P.ngrid[i] - 1, bc = P.bc[i];
                int lo, hi;
                if (i0 == 0) {
                    if (bc == C3SC_ABSORB)        { lo = i0; hi = i0; wall = true; }
                    else if (bc == C3SC_REFLECT)  { lo = i0; hi = i0 + 1; }
                    else                          { lo = P.ngrid[i] - 2; hi = i0 + 1; }
                } else if (i0 == last) {
                    if (bc == C3SC_ABSORB)        { lo = i0; hi = i0; wall = true; }
                    else if (bc == C3SC_REFLECT)  { lo = i0 - 1; hi = i0; }
                    else                          { lo = i0 - 1; hi = 1; }
                } else { lo = i0 - 1; hi = i0 + 1; }
                if (tid == 0) { sNf[slot] = lo; sNf[slot + 1] = hi; }
                slot += 2;
            }
        }
        for (int j = tid; j < N; j += NT) {
            double x[DX];
#pragma unroll
            for (int i = 0; i < DX; i++) x[i] = P.xgrid[P.xoff[i] + (i == k ? j : fix[i])];
            int ab = 0;
            for (int o = 0; o < P.nobs && ab == 0; o++) {           // boundary.c:329-344,668-680
                const double *lb = P.obs + (size_t)o * 2 * DX, *ub = lb + DX;
                bool inside = true;
#pragma unroll
                for (int i = 0; i < DX; i++) inside = inside && !(x[i] < lb[i] || x[i] > ub[i]);
                if (inside) ab = -1;
            }
            if (wall) ab = 1;
            int lo = j - 1, hi = j + 1;
            const int bk = P.bc[k];
            if (j == 0) {                                           // ends overwrite (nodeutil.c:570-612)
                if (bk == C3SC_ABSORB)       { lo = 0; hi = 0; ab = 1; }
                else if (bk == C3SC_REFLECT) { lo = 0; hi = 1; ab = 0; }
                else                         { lo = N - 2; hi = 1; ab = 0; }
            } else if (j == N - 1) {
                if (bk == C3SC_ABSORB)       { lo = N - 1; hi = N - 1; ab = 1; }
                else if (bk == C3SC_REFLECT) { lo = N - 2; hi = N - 1; ab = 0; }
                else                         { lo = N - 2; hi = 1; ab = 0; }
            } else if (ab != 0) { lo = j; hi = j; }
            sAbs[j] = ab;
            sNv[2 * j] = lo;
            sNv[2 * j + 1] = hi;
        }

        __syncthreads();                       // sFix / sNf / flags visible

        // ---- 2. function-train neighbour values (valuefunc.c:369-585) ---------
        const int rk = ft.r[k], rk1 = ft.r[k + 1];
#define CORE_BLK(i, j) (ft.base + ft.off[i] + (size_t)(j) * ft.r[i] * ft.r[(i) + 1])
        // 2a. prefix row vectors L_i = G_0[f]..G_{i-1}[f] (warp 0), suffix columns R_i = G_i[f]..G_{d-1}[f] (warp 1)
        if (warp == 0) {
            if (lane == 0) sL[0] = 1.0;
            __syncwarp();
            for (int i = 0; i < k; i++) {
                warp_vecmat(ft.r[i], ft.r[i + 1], CORE_BLK(i, sFix[i]), sL + i * rs, sL + (i + 1) * rs, lane);
                __syncwarp();
            }
        } else if (warp == 1) {
            if (lane == 0) sR[DX * rs] = 1.0;
            __syncwarp();
            for (int i = DX - 1; i > k; i--) {
                warp_matvec(ft.r[i], ft.r[i + 1], CORE_BLK(i, sFix[i]), sR + (i + 1) * rs, sR + i * rs, lane);
                __syncwarp();
            }
        }
        __syncthreads();
        // 2b. neighbour vectors: for i<k  a_i^s = L_i G_i[nb] G_{i+1}[f]..G_{k-1}[f]   (length r_k)
        //                        for i>k  c_i^s = G_{k+1}[f]..G_{i-1}[f] G_i[nb] R_{i+1} (length r_{k+1})
        for (int task = warp; task < 2 * (DX - 1); task += NW) {
            const int slot = task >> 1, side = task & 1;
            const int i = slot < k ? slot : slot + 1;
            const int nb = sNf[2 * slot + side];
            double *t0 = sTmp + (warp * 2) * rs, *t1 = t0 + rs;
            double *dst = sNb + (2 * i + side) * rs;
            if (i < k) {
                double *cur = (i + 1 == k) ? dst : t0;
                warp_vecmat(ft.r[i], ft.r[i + 1], CORE_BLK(i, nb), sL + i * rs, cur, lane);
                __syncwarp();
                for (int m = i + 1; m < k; m++) {
                    double *nxt = (m + 1 == k) ? dst : (cur == t0 ? t1 : t0);
                    warp_vecmat(ft.r[m], ft.r[m + 1], CORE_BLK(m, sFix[m]), cur, nxt, lane);
                    __syncwarp();
                    cur = nxt;
                }
            } else {
                double *cur = (i - 1 == k) ? dst : t0;
                warp_matvec(ft.r[i], ft.r[i + 1], CORE_BLK(i, nb), sR + (i + 1) * rs, cur, lane);
                __syncwarp();
                for (int m = i - 1; m > k; m--) {
                    double *nxt = (m - 1 == k) ? dst : (cur == t0 ? t1 : t0);
                    warp_matvec(ft.r[m], ft.r[m + 1], CORE_BLK(m, sFix[m]), cur, nxt, lane);
                    __syncwarp();
                    cur = nxt;
                }
            }
        }
        // 2c. per node: w_j = G_k[j] R_{k+1} (length r_k), u_j = L_k G_k[j] (length r_{k+1})
        {
            const double *Lk = sL + k * rs, *Rk1 = sR + (k + 1) * rs;
            const double *Gk = ft.base + ft.off[k];
            const int blk = rk * rk1;
            for (int e = tid; e < N * rk; e += NT) {
                const int j = e / rk, aa = e - j * rk;
                const double *g = Gk + (size_t)j * blk + aa;
                double acc = 0.0;
                for (int b = 0; b < rk1; b++) acc = fma(__ldg(g + (size_t)b * rk), Rk1[b], acc);
                sW[j * rs + aa] = acc;
            }
            for (int e = tid; e < N * rk1; e += NT) {
                const int j = e / rk1, bb = e - j * rk1;
                const double *g = Gk + (size_t)j * blk + (size_t)bb * rk;
                double acc = 0.0;
                for (int aa = 0; aa < rk; aa++) acc = fma(Lk[aa], __ldg(g + aa), acc);
                sU[j * rs + bb] = acc;
            }
        }
        __syncthreads();
        // 2d. dots: slot 2d = self, slots of fixed dims; slots 2k,2k+1 gathered below
        {
            const double *Lk = sL + k * rs;
            for (int e = tid; e < N * (2 * DX - 1); e += NT) {
                const int q = e / N, j = e - q * N;        // q: 0 = self, then fixed-dim slots in order
                double acc = 0.0;
                int oslot;
                if (q == 0) {
                    for (int aa = 0; aa < rk; aa++) acc = fma(Lk[aa], sW[j * rs + aa], acc);
                    oslot = 2 * DX;
                    sV[j] = acc;
                } else {
                    const int fs = q - 1, slot = fs >> 1, side = fs & 1;
                    const int i = slot < k ? slot : slot + 1;
                    const double *nbv = sNb + (2 * i + side) * rs;
                    if (i < k) for (int aa = 0; aa < rk; aa++) acc = fma(nbv[aa], sW[j * rs + aa], acc);
                    else       for (int bb = 0; bb < rk1; bb++) acc = fma(sU[j * rs + bb], nbv[bb], acc);
                    oslot = 2 * i + side;
                }
                sC[j * CS + oslot] = acc;
            }
        }
        __syncthreads();
        for (int j = tid; j < N; j += NT) {                 // along the fiber (valuefunc.c:514-519)
            sC[j * CS + 2 * k] = sV[sNv[2 * j]];
            sC[j * CS + 2 * k + 1] = sV[sNv[2 * j + 1]];
            if (sAbs[j] == 0 && a.mode == MODE_VI) sAct[atomicAdd(&sMisc[0], 1)] = j;
        }
        __syncthreads();
#undef CORE_BLK

        // optional diagnostics
        if (a.out.absorbed) for (int j = tid; j < N; j += NT) a.out.absorbed[obase + j] = sAbs[j];
        if (a.out.nbr_vary) for (int e = tid; e < 2 * N; e += NT) a.out.nbr_vary[2 * obase + e] = sNv[e];
        if (a.out.nbr_fixed) for (int e = tid; e < 2 * (DX - 1); e += NT) a.out.nbr_fixed[(size_t)f * 2 * (DX - 1) + e] = sNf[e];
        if (a.out.costs) for (int e = tid; e < N * CS; e += NT) a.out.costs[obase * CS + e] = sC[e];

        if (a.mode == MODE_PI_EVAL) {
            // ---- policy evaluation (bellman.c:1774-1828,1863-1871) ------------
            for (int j = tid; j < N; j += NT) {
                double x[DX];
#pragma unroll
                for (int i = 0; i < DX; i++) x[i] = P.xgrid[P.xoff[i] + (i == k ? j : fix[i])];
                const int ab = sAbs[j];
                double v;
                if (ab == 1) v = M::boundcost(x, P.mp);
                else if (ab == -1) v = M::obscost(x, P.mp);
                else {
                    const double *row = a.rows_in + (obase + j) * RW;
                    double prob[CS], c[CS];
#pragma unroll
                    for (int m = 0; m < CS; m++) { prob[m] = row[m]; c[m] = sC[j * CS + m]; }
                    v = rhs<DX, A>(P, prob, row[CS], row[CS + 1], c);
                }
                a.out.value[obase + j] = v;
            }
            continue;
        }

        // ---- 4. absorbed nodes (bellman.c:513-532) -----------------------------
        for (int j = tid; j < N; j += NT) {
            const int ab = sAbs[j];
            if (ab == 0) continue;
            double x[DX];
#pragma unroll
            for (int i = 0; i < DX; i++) x[i] = P.xgrid[P.xoff[i] + (i == k ? j : fix[i])];
            const double v = (ab == 1) ? M::boundcost(x, P.mp) : M::obscost(x, P.mp);
            if (a.write_value) a.out.value[obase + j] = v;
            if (a.out.argmin) a.out.argmin[obase + j] = -1;
            if (a.out.rows) {
                double *row = a.out.rows + (obase + j) * RW;
                for (int m = 0; m < RW; m++) row[m] = 0.0;
            }
        }

        // ---- 3. min over the control table, G lanes per node -------------------
        const int nact = sMisc[0];
        const int ngroups = NT / G, grp = tid / G, gl = tid - grp * G;
        const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((lane / G) * G));
        for (int it = grp; it < nact; it += ngroups) {
            const int j = sAct[it];
            double x[DX], c[CS];
#pragma unroll
            for (int i = 0; i < DX; i++) x[i] = P.xgrid[P.xoff[i] + (i == k ? j : fix[i])];
#pragma unroll
            for (int m = 0; m < CS; m++) c[m] = sC[j * CS + m];
            NodeInv<M> inv;
            node_prepare<M, A>(P, sUtab, x, c, inv);
            double best = CUDART_INF;
            int ibest = 0x7fffffff, bad = 0;
            for (int cand = gl; cand < P.nu; cand += G) {
                double u[DU];
#pragma unroll
                for (int i = 0; i < DU; i++) u[i] = sUtab[cand * DU + i];
                const double v = candidate_value<M, A>(P, x, u, c, inv, bad);
                if (v < best) { best = v; ibest = cand; }
            }
            for (int off = G >> 1; off > 0; off >>= 1) {     // first strict minimum in table order
                const double ov = __shfl_xor_sync(gmask, best, off);
                const int oi = __shfl_xor_sync(gmask, ibest, off);
                if (ov < best || (ov == best && oi < ibest)) { best = ov; ibest = oi; }
            }
            if (bad) atomicOr(P.err, 1);
            if (gl == 0) {
                if (a.write_value) a.out.value[obase + j] = best;
                if (a.out.argmin) a.out.argmin[obase + j] = ibest;
                if (a.out.rows) {                           // policy row at u* (bellman.c:1851-1860)
                    double u[DU], b[DX], s[DX], prob[CS], dt;
#pragma unroll
                    for (int i = 0; i < DU; i++) u[i] = sUtab[(ibest < P.nu ? ibest : 0) * DU + i];
                    M::template drift<A>(x, u, P.mp, b);
                    M::template sigma<A>(x, u, P.mp, s);
                    const double g = M::template stage<A>(x, u, P.mp);
                    if (transition_row<DX, A>(P, b, s, prob, dt)) atomicOr(P.err, 1);
                    double *row = a.out.rows + (obase + j) * RW;
#pragma unroll
                    for (int m = 0; m < CS; m++) row[m] = prob[m];
                    row[CS] = dt;
                    row[CS + 1] = g;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// choose lanes-per-node: minimise rounds(controls) x rounds(nodes), prefer wider groups
inline int pick_group(int nu, int nmax)
{
    int best = 32, bestc = 1 << 30;
    for (int G = 32; G >= 1; G >>= 1) {
        const int cr = (nu + G - 1) / G, groups = NT / G;
        const int nodes = (nmax * 9 + 9) / 10;
        const int nr = (nodes + groups - 1) / groups;
        const int cost = cr * nr;
        if (cost < bestc) { bestc = cost; best = G; }
    }
    return best;
}

template <class M, class A>
int launch_backup_t(const LaunchArgs &a, cudaStream_t st)
{
    static int sms = 0, max_optin = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    const SmemPlan sp(M::DX, M::DU, a.P.nu, a.P.nmax, a.ft.rmax);
    const size_t smem = sp.bytes();
    if (smem > (size_t)max_optin) return (int)cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k_backup<M, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_backup<M, A>, NT, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    int grid = sms * per_sm;
    if (grid > a.F) grid = a.F;
    if (grid < 1) return 0;
    const int G = pick_group(a.P.nu, a.P.nmax);
    k_backup<M, A><<<grid, NT, smem, st>>>(a, G);
    return (int)cudaGetLastError();
}

template <class M>
int launch_backup_m(int arith, const LaunchArgs &a, cudaStream_t st)
{
    if (arith == C3SC_ARITH_EXACT) return launch_backup_t<M, Exact>(a, st);
    return launch_backup_t<M, Fast>(a, st);
}

// ---- small test kernels -------------------------------------------------------
template <class M>
__global__ void k_model_eval(const DevProblem P, int n, const double *x, const double *u, double *drift,
                             double *sig, double *stage, double *bound, double *obs)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double xx[M::DX], uu[M::DU], b[M::DX], s[M::DX];
    for (int i = 0; i < M::DX; i++) xx[i] = x[(size_t)e * M::DX + i];
    for (int i = 0; i < M::DU; i++) uu[i] = u[(size_t)e * M::DU + i];
    M::template drift<Exact>(xx, uu, P.mp, b);
    M::template sigma<Exact>(xx, uu, P.mp, s);
    for (int i = 0; i < M::DX; i++) { drift[(size_t)e * M::DX + i] = b[i]; sig[(size_t)e * M::DX + i] = s[i]; }
    stage[e] = M::template stage<Exact>(xx, uu, P.mp);
    bound[e] = M::boundcost(xx, P.mp);
    obs[e] = M::obscost(xx, P.mp);
}

template <class M>
int launch_model_eval_t(const DevProblem &P, int n, const double *x, const double *u, double *drift,
                        double *sig, double *stage, double *bound, double *obs, cudaStream_t st)
{
    if (n <= 0) return 0;
    k_model_eval<M><<<(n + 127) / 128, 128, 0, st>>>(P, n, x, u, drift, sig, stage, bound, obs);
    return (int)cudaGetLastError();
}

// transition_assemble on caller-supplied (drift, diag sigma) pairs: parity-test entry
template <int DX, class A>
__global__ void k_transition(const DevProblem P, int n, const double *drift, const double *sig,
                             double *prob, double *dt, int *status)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double b[DX], s[DX], p[2 * DX + 1], t;
    for (int i = 0; i < DX; i++) { b[i] = drift[(size_t)e * DX + i]; s[i] = sig[(size_t)e * DX + i]; }
    const int st = transition_row<DX, A>(P, b, s, p, t);
    for (int m = 0; m < 2 * DX + 1; m++) prob[(size_t)e * (2 * DX + 1) + m] = p[m];
    dt[e] = t;
    status[e] = st;
}

}  // namespace c3sc
