// control_kernel.cuh -- stage 2 of the Bellman backup: per node, the min over the discretised
// control set of
//        dt*g + exp(-beta*dt) * <p, V_nbr>        bellman_control / bellmanrhs
//                                                 src/bellman.c:88-112,367-480,504-543
// with p, dt from the upwind Kushner-Dupuis construction (transition_assemble,
// src/nodeutil.c:267-406) and V_nbr the neighbour values stage 1 left in the slot-major
// scratch.  One thread owns one (node, candidate chunk); the chunks of a node sit in adjacent
// lanes and are merged in table order with strict '<', so the result is the FIRST strict
// minimum in table order -- the brute-force c3opt_minimize rule (bellman.c:539-543).
// Absorbed nodes take boundcost / obscost (bellman.c:513-532).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include "dev_types.h"
#include "ctl_types.cuh"
#include "arith.cuh"
#include "models.cuh"

namespace c3sc {

// backed-up value of node `id`: the local output and, when the all-gather is fused, every rank's copy
__device__ __forceinline__ void store_value(const CtlArgs &c, long long id, double v)
{
    if (c.value) c.value[id] = v;
    for (int g = 0; g < c.npeer; g++) c.vpeer[g][c.peer_off + id] = v;
}

// ---------------------------------------------------------------------------
// fast reciprocal and exp for the FAST policy (both ~1 ulp)
__device__ __forceinline__ double rcp_pos(double x)
{   // x > 0, normal: MUFU.RCP64H seed (~2^-20) + two Newton steps
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
__device__ __forceinline__ double exp_nonpos(double x)
{   // exp(x) for x <= 0: n = rint(x*log2 e), r = x - n ln2 (Cody-Waite), degree-13 Taylor on
    // |r| <= ln2/2 (truncation 4e-18), scale by adding n to the exponent field.  No overflow
    // path is needed for x <= 0; below -708 the result is flushed to 0.
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);
    const double n = t - 6755399441055744.0;
    double r = fma(n, -6.93147180369123816490e-01, x);
    r = fma(n, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;            // 1/13!
    p = fma(p, r, 2.08767569878681e-09);          // 1/12!
    p = fma(p, r, 2.505210838544172e-08);         // 1/11!
    p = fma(p, r, 2.755731922398589e-07);         // 1/10!
    p = fma(p, r, 2.7557319223985893e-06);        // 1/9!
    p = fma(p, r, 2.48015873015873e-05);          // 1/8!
    p = fma(p, r, 1.984126984126984e-04);         // 1/7!
    p = fma(p, r, 1.388888888888889e-03);         // 1/6!
    p = fma(p, r, 8.333333333333333e-03);         // 1/5!
    p = fma(p, r, 4.1666666666666664e-02);        // 1/4!
    p = fma(p, r, 1.6666666666666666e-01);        // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int ni = __double2loint(t);
    const double res = __hiloint2double(__double2hiint(p) + (ni << 20), __double2loint(p));
    return (x < -708.0) ? 0.0 : res;
}
__device__ __forceinline__ double exp_tiny(double x)
{   // exp(x) for -2^-8 <= x <= 0: degree-5 Taylor, truncation x^6/720 < 5e-18 (relative, e^x ~ 1)
    double p = 8.333333333333333e-03;
    p = fma(p, x, 4.1666666666666664e-02);
    p = fma(p, x, 1.6666666666666666e-01);
    p = fma(p, x, 0.5);
    p = fma(p, x, 1.0);
    return fma(p, x, 1.0);
}

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
        const double rinv = rcp_pos(norm);
        const double dt = P.h2 * rinv;
        const double ebt = exp_nonpos(-P.beta * dt);
        return fma(dt, g, ebt * (S * rinv));
    }
}

// node id -> state
template <int DX>
__device__ __forceinline__ void node_state(const CtlArgs &c, int id, double *x)
{
    const int f = id / c.ldo, j = id - f * c.ldo;
    int k = c.dim_vary[f];
    k = k < 0 ? 0 : (k >= DX ? DX - 1 : k);                     // stage 1 clamps the same way
#pragma unroll
    for (int i = 0; i < DX; i++) {
        int i0 = c.fixed_ind[(size_t)f * DX + i];                   // clamped like stage 1 (bad descriptors are reported by k_group_fibers)
        i0 = i0 < 0 ? 0 : (i0 >= c.P.ngrid[i] ? c.P.ngrid[i] - 1 : i0);
        x[i] = c.P.xgrid[c.P.xoff[i] + (i == k ? j : i0)];
    }
}

#ifndef C3SC_CST_STREAM
#define C3SC_CST_STREAM 1
#endif
#if C3SC_CST_STREAM
#define C3SC_CST_LD(p) __ldcs(p)
#else
#define C3SC_CST_LD(p) (*(p))
#endif
// Neighbour values of node `id` from the slot-major scratch.  Stage 1 stores the fixed-dimension
// neighbours and the node's own value (slot 2dx); the two neighbours ALONG the fiber are the own
// values of other nodes of the same fiber (valuefunc.c:514-519) and are fetched here.
template <int DX>
__device__ __forceinline__ void load_costs(const CtlArgs &c, int id, double *cc)
{
    const int f = id / c.ldo, j = id - f * c.ldo;
    int k = c.dim_vary[f];
    k = k < 0 ? 0 : (k >= DX ? DX - 1 : k);                     // stage 1 clamps the same way
    int lo, hi;
    ft_vary_pair(c.P.bc[k], c.P.ngrid[k], j, c.flag[id], lo, hi);
    const double *self = c.cst + (size_t)(2 * DX) * c.NS + (size_t)f * c.ldo;
    const double vlo = self[lo], vhi = self[hi];
#pragma unroll
    for (int i = 0; i < DX; i++) {
        // read once, dead afterwards: evict-first loads (ld.global.cs) keep the chain records and rows in L2 instead
        cc[2 * i] = i == k ? vlo : C3SC_CST_LD(c.cst + (size_t)(2 * i) * c.NS + id);
        cc[2 * i + 1] = i == k ? vhi : C3SC_CST_LD(c.cst + (size_t)(2 * i + 1) * c.NS + id);
    }
    cc[2 * DX] = self[j];
}

// ---------------------------------------------------------------------------
template <class M, class A>
__global__ void __launch_bounds__(CT_NT, 3) k_control(const CtlArgs c)
{
    constexpr int DX = M::DX, DU = M::DU, CS = 2 * DX + 1, RW = 2 * DX + 3;
    constexpr bool TAB = M::SEP && !A::exact;
    constexpr int NUD = M::NUD, CTW = 2 * NUD + 2;
    const DevProblem &P = c.P;
    const int tid = threadIdx.x, lane = tid & 31;
    extern __shared__ __align__(16) double smem[];

    // candidate table (separable models, FAST) or the raw control table, staged once per CTA
    const bool grouped = TAB && c.ng > 0;
    {
        const double *src = TAB ? (grouped ? P.gtab : P.ctab) : P.utab;
        const int cnt = P.nu * (TAB ? CTW : DU);
        for (int e = tid; e < cnt; e += CT_NT) smem[e] = src[e];
        __syncthreads();
    }
    const double *tab = smem;

    const int pl2 = c.parts_log2, parts = 1 << pl2;
    const int chunk = (P.nu + parts - 1) >> pl2;
    const int nact = *c.act_count;
    const long long total = (long long)nact << pl2;
    const long long stride = (long long)gridDim.x * CT_NT;
    // warp-uniform trip count: every lane of a warp stays in the loop for the shuffles
    for (long long it0 = (long long)blockIdx.x * CT_NT + (tid & ~31); it0 < total; it0 += stride) {
        const long long it = it0 + lane;
        const bool valid = it < total;
        const int idx = valid ? (int)(it >> pl2) : 0, part = (int)(it & (parts - 1));
        const int id = c.act[idx];
        const int c0 = part * chunk, c1 = valid ? ((c0 + chunk < P.nu) ? c0 + chunk : P.nu) : c0;
        double x[DX], cc[CS];
        node_state<DX>(c, id, x);
        load_costs<DX>(c, id, cc);
        double best = CUDART_INF;
        int ibest = 0x7fffffff;
        if (TAB) {
            // per-node invariants: norm0 = control-independent part of the normaliser, S0 = the
            // matching part of <raw p, V_nbr>, gx = stage_x; cu = the 2*NUD neighbour values the
            // candidates weight.
            double u0[DU], b0[DX], s0[DX];
#pragma unroll
            for (int i = 0; i < DU; i++) u0[i] = P.utab[i];
            M::template drift<A>(x, u0, P.mp, b0);
            M::template sigma<A>(x, u0, P.mp, s0);
            double norm0 = 0.0, S0 = 0.0;
#pragma unroll
            for (int i = 0; i < DX; i++) {
                const double q = (P.t[2 * i + 1] * 0.5) * (s0[i] * s0[i]);
                double rl = q, rr = q;
                if (!M::u_dep(i)) {
                    const double tb = P.t[2 * i] * b0[i];
                    rl = q - ((b0[i] < -1e-14) ? tb : 0.0);
                    rr = q + ((b0[i] > 1e-14) ? tb : 0.0);
                }
                norm0 += rl + rr;
                S0 = fma(rl, cc[2 * i], S0);
                S0 = fma(rr, cc[2 * i + 1], S0);
            }
            const double gx = M::stage_x(x, P.mp);
            double cu[2 * NUD];
#pragma unroll
            for (int m = 0; m < NUD; m++) { cu[2 * m] = cc[2 * M::ud(m)]; cu[2 * m + 1] = cc[2 * M::ud(m) + 1]; }
            if (valid && norm0 + P.amin < 1e-14) atomicOr(P.err, 1);
            const double nbh = -P.beta * P.h2;
            const bool disc = P.beta != 0.0;
            // exp(-beta*dt) with dt <= h2/(norm0+amin): short series when the node's bound is tiny.  Decided per
            // NODE (not by a warp vote): which nodes share a warp depends on the atomically filled active list,
            // and the result of a node must not depend on its neighbours in the list.
            const bool tiny = !valid || (P.beta * P.h2 <= 0.00390625 * (norm0 + P.amin));
            if (grouped) {
                // Candidates with the same normaliser share A_g share dt and the discount: per group
                // one reciprocal + one exp, per candidate only S_c and  t_c = ebt*S_c + h2*gu_c.
                // value_c = rinv_g * (t_c + h2*gx).  Rows of a group keep table order (strict '<'),
                // groups are merged on (value, table index).
                const double hgx = P.h2 * gx;
                for (int g = 0; g < c.ng; g++) {
                    const int lo = c.gstart[g] > c0 ? c.gstart[g] : c0;
                    const int hi = c.gstart[g + 1] < c1 ? c.gstart[g + 1] : c1;
                    if (lo >= hi) continue;
                    const double rinv = rcp_pos(norm0 + c.gA[g]);
                    const double ebt = !disc ? 1.0 : (tiny ? exp_tiny(nbh * rinv) : exp_nonpos(nbh * rinv));
                    double bt = CUDART_INF;
                    int bi = 0x7fffffff;
#pragma unroll 2
                    for (int pos = lo; pos < hi; pos++) {
                        const double2 *row = reinterpret_cast<const double2 *>(tab + pos * CTW);
                        double Sa = S0, Sb = 0.0;
#pragma unroll
                        for (int m = 0; m < NUD; m++) {
                            const double2 w = row[m];
                            Sa = fma(w.x, cu[2 * m], Sa);
                            Sb = fma(w.y, cu[2 * m + 1], Sb);
                        }
                        const double2 hi2 = row[NUD];           // (h2*gu_c, table index in the low word)
                        const double t = fma(ebt, Sa + Sb, hi2.x);
                        if (t < bt) { bt = t; bi = __double2loint(hi2.y); }
                    }
                    const double v = rinv * (bt + hgx);
                    if (v < best || (v == best && bi < ibest)) { best = v; ibest = bi; }
                }
            } else if (tiny) {
#pragma unroll 2
                for (int cand = c0; cand < c1; cand++) {
                    const double2 *row = reinterpret_cast<const double2 *>(tab + cand * CTW);
                    double S = S0;
#pragma unroll
                    for (int m = 0; m < NUD; m++) {
                        const double2 w = row[m];
                        S = fma(w.x, cu[2 * m], S);
                        S = fma(w.y, cu[2 * m + 1], S);
                    }
                    const double2 ag = row[NUD];
                    const double rinv = rcp_pos(norm0 + ag.x);
                    const double ebt = disc ? exp_tiny(nbh * rinv) : 1.0;
                    const double v = rinv * fma(P.h2, gx + ag.y, ebt * S);
                    if (v < best) { best = v; ibest = cand; }
                }
            } else {
                for (int cand = c0; cand < c1; cand++) {
                    const double2 *row = reinterpret_cast<const double2 *>(tab + cand * CTW);
                    double S = S0;
#pragma unroll
                    for (int m = 0; m < NUD; m++) {
                        const double2 w = row[m];
                        S = fma(w.x, cu[2 * m], S);
                        S = fma(w.y, cu[2 * m + 1], S);
                    }
                    const double2 ag = row[NUD];
                    const double rinv = rcp_pos(norm0 + ag.x);
                    const double ebt = disc ? exp_nonpos(nbh * rinv) : 1.0;
                    const double v = rinv * fma(P.h2, gx + ag.y, ebt * S);
                    if (v < best) { best = v; ibest = cand; }
                }
            }
        } else {
            NodeInv<M> inv;
            node_prepare<M, A>(P, P.utab, x, cc, inv);
            int bad = 0;
            for (int cand = c0; cand < c1; cand++) {
                double u[DU];
#pragma unroll
                for (int i = 0; i < DU; i++) u[i] = tab[(size_t)cand * DU + i];
                const double v = candidate_value<M, A>(P, x, u, cc, inv, bad);
                if (v < best) { best = v; ibest = cand; }
            }
            if (bad) atomicOr(P.err, 1);
        }
        // merge the chunks of a node: lower part = earlier in the table, wins ties
        for (int o = 1; o < parts; o <<= 1) {
            const double vb = __shfl_down_sync(0xffffffffu, best, o);
            const int ib = __shfl_down_sync(0xffffffffu, ibest, o);
            if (vb < best || (vb == best && ib < ibest)) { best = vb; ibest = ib; }
        }
        if (!valid || part != 0) continue;
        store_value(c, id, best);
        if (c.argmin) c.argmin[id] = ibest;
        if (c.rows) {                               // policy row at u* (bellman.c:1851-1860)
            double u[DU], b[DX], s[DX], prob[CS], dt;
            node_state<DX>(c, id, x);               // recomputed: keeps x out of the candidate loop's live set
#pragma unroll
            for (int i = 0; i < DU; i++) u[i] = P.utab[(size_t)(ibest < P.nu ? ibest : 0) * DU + i];
            M::template drift<A>(x, u, P.mp, b);
            M::template sigma<A>(x, u, P.mp, s);
            const double g = M::template stage<A>(x, u, P.mp);
            if (transition_row<DX, A>(P, b, s, prob, dt)) atomicOr(P.err, 1);
            double *row = c.rows + (size_t)id * RW;
#pragma unroll
            for (int m = 0; m < CS; m++) row[m] = prob[m];
            row[CS] = dt;
            row[CS + 1] = g;
        }
    }

    // absorbed nodes (bellman.c:513-532): boundary / obstacle cost, u = 0
    for (long long id = (long long)blockIdx.x * CT_NT + tid; id < c.NS; id += stride) {
        const int ab = c.flag[id];
        if (ab == 0) continue;
        double v = 0.0;                                     // ab == 2: padding entry j >= ngrid[dim_vary], defined as 0 / -1
        if (ab != 2) {
            double x[DX];
            node_state<DX>(c, (int)id, x);
            v = (ab == 1) ? M::boundcost(x, P.mp) : M::obscost(x, P.mp);
        }
        store_value(c, id, v);
        if (c.argmin) c.argmin[id] = -1;
        if (c.rows) {
            double *row = c.rows + (size_t)id * RW;
            for (int m = 0; m < RW; m++) row[m] = 0.0;
        }
    }
}

// Large batches of a separable model (FAST, grouped candidates): TWO nodes per thread, so every
// candidate row read from shared memory feeds two independent FMA chains (half the LDS traffic per
// candidate-node, twice the instruction-level parallelism).  Same arithmetic as k_control's grouped
// walk; nodes it and it+ceil(nact/2) of the active list share a thread.
template <class M>
struct Node2 {
    double S0, norm0, hgx, cu[2 * M::NUD];
    double best;
    int ibest;
};
template <class M>
__device__ __forceinline__ void node2_from(const CtlArgs &c, const double *x, const double *cc, Node2<M> &n);
template <class M>
__device__ __forceinline__ void node2_prepare(const CtlArgs &c, int id, Node2<M> &n)
{
    constexpr int DX = M::DX;
    double x[DX], cc[2 * DX + 1];
    node_state<DX>(c, id, x);
    load_costs<DX>(c, id, cc);
    node2_from<M>(c, x, cc, n);
}
// per-node invariants of the separable FAST walk from the node's state and its 2dx+1 neighbour values
template <class M>
__device__ __forceinline__ void node2_from(const CtlArgs &c, const double *x, const double *cc, Node2<M> &n)
{
    constexpr int DX = M::DX, DU = M::DU, NUD = M::NUD;
    const DevProblem &P = c.P;
    double u0[DU], b0[DX], s0[DX];
#pragma unroll
    for (int i = 0; i < DU; i++) u0[i] = P.utab[i];
    M::template drift<Fast>(x, u0, P.mp, b0);
    M::template sigma<Fast>(x, u0, P.mp, s0);
    double norm0 = 0.0, S0 = 0.0;
#pragma unroll
    for (int i = 0; i < DX; i++) {
        const double cl = cc[2 * i], cr = cc[2 * i + 1];
        const double q = (P.t[2 * i + 1] * 0.5) * (s0[i] * s0[i]);
        double rl = q, rr = q;
        if (!M::u_dep(i)) {
            const double tb = P.t[2 * i] * b0[i];
            rl = q - ((b0[i] < -1e-14) ? tb : 0.0);
            rr = q + ((b0[i] > 1e-14) ? tb : 0.0);
        }
        norm0 += rl + rr;
        S0 = fma(rl, cl, S0);
        S0 = fma(rr, cr, S0);
#pragma unroll
        for (int m = 0; m < NUD; m++)
            if (M::ud(m) == i) { n.cu[2 * m] = cl; n.cu[2 * m + 1] = cr; }
    }
    n.S0 = S0; n.norm0 = norm0; n.hgx = P.h2 * M::stage_x(x, P.mp);
    n.best = CUDART_INF; n.ibest = 0x7fffffff;
}

#ifndef C3SC_C2_UNROLL
#define C3SC_C2_UNROLL 2
#endif
#ifndef C3SC_C2_MINB
#define C3SC_C2_MINB 2
#endif
#ifndef C3SC_C2_NODES
#define C3SC_C2_NODES 2
#endif
#ifndef C3SC_C2_NT
#define C3SC_C2_NT CT_NT
#endif
constexpr int C2U = C3SC_C2_UNROLL;
constexpr int C2N = C3SC_C2_NODES;         // nodes per thread
constexpr int C2_NT = C3SC_C2_NT;          // threads per CTA for large batches
template <class M>
__global__ void __launch_bounds__(C2_NT, C3SC_C2_MINB) k_control2(const CtlArgs c)
{
    constexpr int DX = M::DX, DU = M::DU, CS = 2 * DX + 1, RW = 2 * DX + 3;
    constexpr int NUD = M::NUD, CTW = 2 * NUD + 2;
    const DevProblem &P = c.P;
    const int tid = threadIdx.x, lane = tid & 31;
    extern __shared__ __align__(16) double smem[];
    {
        const int cnt = P.nu * CTW;
        for (int e = tid; e < cnt; e += (int)blockDim.x) smem[e] = P.gtab[e];
        __syncthreads();
    }
    const double *tab = smem;
    const int nact = *c.act_count, nper = (nact + C2N - 1) / C2N;    // node q of a thread: active position it + q*nper
    const long long stride = (long long)gridDim.x * blockDim.x;      // launched with C2_NT or (small batches) 128 threads
    const double nbh = -P.beta * P.h2;
    const bool disc = P.beta != 0.0;
    for (long long it0 = (long long)blockIdx.x * blockDim.x + (tid & ~31); it0 < nper; it0 += stride) {
        const int it = (int)it0 + lane;
        bool valid[C2N];
        int id[C2N];
        Node2<M> nd[C2N];
        bool bad = false, tiny[C2N];
#pragma unroll
        for (int q = 0; q < C2N; q++) {
            valid[q] = it < nper && it + q * nper < nact;
            id[q] = c.act[valid[q] ? it + q * nper : 0];
            node2_prepare<M>(c, id[q], nd[q]);
            bad = bad || (valid[q] && nd[q].norm0 + P.amin < 1e-14);
            tiny[q] = !valid[q] || P.beta * P.h2 <= 0.00390625 * (nd[q].norm0 + P.amin);      // per node: see k_control
        }
        if (bad) atomicOr(P.err, 1);
        for (int g = 0; g < c.ng; g++) {
            const int lo = c.gstart[g], hi = valid[0] ? c.gstart[g + 1] : lo;
            double rinv[C2N], ebt[C2N], bt[C2N];
            int bi[C2N];
#pragma unroll
            for (int q = 0; q < C2N; q++) {
                rinv[q] = rcp_pos(nd[q].norm0 + c.gA[g]);
                ebt[q] = !disc ? 1.0 : (tiny[q] ? exp_tiny(nbh * rinv[q]) : exp_nonpos(nbh * rinv[q]));
                bt[q] = CUDART_INF;
                bi[q] = 0x7fffffff;
            }
#pragma unroll C2U
            for (int pos = lo; pos < hi; pos++) {
                const double2 *row = reinterpret_cast<const double2 *>(tab + pos * CTW);
                double S0[C2N], S1[C2N];
#pragma unroll
                for (int q = 0; q < C2N; q++) { S0[q] = nd[q].S0; S1[q] = 0.0; }
#pragma unroll
                for (int m = 0; m < NUD; m++) {
                    const double2 w = row[m];
#pragma unroll
                    for (int q = 0; q < C2N; q++) {
                        S0[q] = fma(w.x, nd[q].cu[2 * m], S0[q]);
                        S1[q] = fma(w.y, nd[q].cu[2 * m + 1], S1[q]);
                    }
                }
                const double2 hi2 = row[NUD];                   // (h2*gu_c, table index in the low word)
                const int ci = __double2loint(hi2.y);
#pragma unroll
                for (int q = 0; q < C2N; q++) {
                    const double t = fma(ebt[q], S0[q] + S1[q], hi2.x);
                    if (t < bt[q]) { bt[q] = t; bi[q] = ci; }
                }
            }
#pragma unroll
            for (int q = 0; q < C2N; q++) {
                const double v = rinv[q] * (bt[q] + nd[q].hgx);
                if (v < nd[q].best || (v == nd[q].best && bi[q] < nd[q].ibest)) { nd[q].best = v; nd[q].ibest = bi[q]; }
            }
        }
#pragma unroll
        for (int q = 0; q < C2N; q++) {
            if (!valid[q]) continue;
            const int ibest = nd[q].ibest;
            store_value(c, id[q], nd[q].best);
            if (c.argmin) c.argmin[id[q]] = ibest;
            if (c.rows) {                               // policy row at u* (bellman.c:1851-1860)
                double x[DX], u[DU], b[DX], s[DX], prob[CS], dt;
                node_state<DX>(c, id[q], x);
#pragma unroll
                for (int i = 0; i < DU; i++) u[i] = P.utab[(size_t)(ibest < P.nu ? ibest : 0) * DU + i];
                M::template drift<Fast>(x, u, P.mp, b);
                M::template sigma<Fast>(x, u, P.mp, s);
                const double g = M::template stage<Fast>(x, u, P.mp);
                if (transition_row<DX, Fast>(P, b, s, prob, dt)) atomicOr(P.err, 1);
                double *row = c.rows + (size_t)id[q] * RW;
#pragma unroll
                for (int m = 0; m < CS; m++) row[m] = prob[m];
                row[CS] = dt;
                row[CS + 1] = g;
            }
        }
    }
    // absorbed nodes (bellman.c:513-532): boundary / obstacle cost, u = 0
    for (long long id = (long long)blockIdx.x * blockDim.x + tid; id < c.NS; id += stride) {
        const int ab = c.flag[id];
        if (ab == 0) continue;
        double v = 0.0;                                     // ab == 2: padding entry j >= ngrid[dim_vary], defined as 0 / -1
        if (ab != 2) {
            double x[DX];
            node_state<DX>(c, (int)id, x);
            v = (ab == 1) ? M::boundcost(x, P.mp) : M::obscost(x, P.mp);
        }
        store_value(c, id, v);
        if (c.argmin) c.argmin[id] = -1;
        if (c.rows) {
            double *row = c.rows + (size_t)id * RW;
            for (int m = 0; m < RW; m++) row[m] = 0.0;
        }
    }
}

// ---------------------------------------------------------------------------
// Grid-structured control tables ({lo, 0, hi}^NUD, C order): every candidate is still formed and compared,
// but candidates that share their first controls share the partial sum S0 + sum_{m' < m} T_{m'}(k_{m'}), so a
// candidate costs one DADD (amortised 2/3 of one: the zero level adds nothing) and one compare instead of
// 2*NUD DFMAs.  The walk is fully unrolled (3^NUD leaves, every index static), candidates are visited in
// table order, and the running minimum is kept per number of non-zero controls -- the groups of the
// grouped walk, whose reciprocal/exp are shared.  ARG = false (value iteration: no argmin, no policy rows)
// drops the index bookkeeping.
template <int NUD, int m, int cnt, int idx, bool ARG>
struct GridWalk {
    static __device__ __forceinline__ void go(double s, const double (&Tl)[NUD], const double (&Th)[NUD],
                                              double (&mn)[NUD + 1], int (&am)[NUD + 1])
    {
        if constexpr (m == NUD) {
            if (s < mn[cnt]) {                   // DSETP + predicated moves (fmin's NaN handling costs three more slots)
                mn[cnt] = s;
                if constexpr (ARG) am[cnt] = idx;
            }
        } else {
            constexpr int stride = GridWalk<NUD, m + 1, 0, 0, ARG>::SPAN;      // 3^(NUD-1-m): last control fastest
            GridWalk<NUD, m + 1, cnt + 1, idx, ARG>::go(s + Tl[m], Tl, Th, mn, am);
            GridWalk<NUD, m + 1, cnt, idx + stride, ARG>::go(s, Tl, Th, mn, am);
            GridWalk<NUD, m + 1, cnt + 1, idx + 2 * stride, ARG>::go(s + Th[m], Tl, Th, mn, am);
        }
    }
    static constexpr int span()
    {
        int v = 1;
        for (int i = m; i < NUD; i++) v *= 3;
        return v;
    }
    static constexpr int SPAN = span();      // candidates below a node of depth m
};

// The same walk for Q nodes of one thread in lock step: Q independent dependency chains per leaf.  Used by the fused
// walk, which runs inside the node kernel at 16 warps per SM and 128 registers per thread -- latency has to be covered
// by instruction-level parallelism there, not by occupancy.
template <int NUD, int m, int cnt, int idx, bool ARG, int Q>
struct GridWalkQ {
    static __device__ __forceinline__ void go(const double (&s)[Q], const double (&Tl)[Q][NUD], const double (&Th)[Q][NUD],
                                              double (&mn)[Q][NUD + 1], int (&am)[Q][NUD + 1])
    {
        if constexpr (m == NUD) {
#pragma unroll
            for (int q = 0; q < Q; q++)
                if (s[q] < mn[q][cnt]) {
                    mn[q][cnt] = s[q];
                    if constexpr (ARG) am[q][cnt] = idx;
                }
        } else {
            constexpr int stride = GridWalk<NUD, m + 1, 0, 0, ARG>::SPAN;
            double sl[Q], sh[Q];
#pragma unroll
            for (int q = 0; q < Q; q++) { sl[q] = s[q] + Tl[q][m]; sh[q] = s[q] + Th[q][m]; }
            GridWalkQ<NUD, m + 1, cnt + 1, idx, ARG, Q>::go(sl, Tl, Th, mn, am);
            GridWalkQ<NUD, m + 1, cnt, idx + stride, ARG, Q>::go(s, Tl, Th, mn, am);
            GridWalkQ<NUD, m + 1, cnt + 1, idx + 2 * stride, ARG, Q>::go(sh, Tl, Th, mn, am);
        }
    }
};

// Occupancy beats registers here: the walk has no loop-carried state beyond the du+1 running minima, and the
// 21 neighbour values of a node arrive from L2 -- one node per thread at 4 CTAs/SM (64 registers, ~100 B of
// spills) measured 4 % faster end to end than two nodes per thread at 2 CTAs/SM (profiles/r01_lanes.md).
#ifndef C3SC_GRID_MINB
#define C3SC_GRID_MINB 3          // re-measured with the 10 922-fiber chunks of round 2: 3 CTAs/SM (85 registers, no spills) 1.735 ms per
#endif                             // 65 536-fiber step, 4 (64 registers) 1.751, 2: 1.801, 5: 1.958 (profiles/r02b_stage1.md)
#ifndef C3SC_GRID_Q
#define C3SC_GRID_Q 1          // nodes per thread for large batches
#endif
constexpr int grid_minb(bool arg) { return arg && C3SC_GRID_MINB > 3 ? 3 : C3SC_GRID_MINB; }   // the index bookkeeping needs more registers
template <class M, bool ARG, int Q>
__global__ void __launch_bounds__(CT_NT, grid_minb(ARG)) k_control_grid(const CtlArgs c)
{
    constexpr int DX = M::DX, DU = M::DU, CS = 2 * DX + 1, RW = 2 * DX + 3;
    constexpr int NUD = M::NUD, NG = NUD + 1;
    const DevProblem &P = c.P;
    const int tid = threadIdx.x, lane = tid & 31;
    const int nact = *c.act_count, nper = (nact + Q - 1) / Q;
    const long long stride = (long long)gridDim.x * blockDim.x;
    constexpr bool STAGE = ARG && (size_t)CT_NT * RW * sizeof(double) <= 47 * 1024;     // (dx <= 10; larger records go out directly)
    __shared__ double srows[STAGE ? CT_NT * RW : 1];        // policy rows of the CTA's nodes on their way out
    const double nbh = -P.beta * P.h2;
    const bool disc = P.beta != 0.0;
    for (long long it0 = (long long)blockIdx.x * blockDim.x + (tid & ~31); it0 < nper; it0 += stride) {
        const int it = (int)it0 + lane;
        bool valid[Q], bad = false, tiny[Q];
        int id[Q];
        Node2<M> nd[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) {
            valid[q] = it < nper && it + q * nper < nact;
            id[q] = c.act[valid[q] ? it + q * nper : 0];
            node2_prepare<M>(c, id[q], nd[q]);
            bad = bad || (valid[q] && nd[q].norm0 + P.amin < 1e-14);
            tiny[q] = !valid[q] || P.beta * P.h2 <= 0.00390625 * (nd[q].norm0 + P.amin);      // per node: see k_control
        }
        if (bad) atomicOr(P.err, 1);
#pragma unroll
        for (int q = 0; q < Q; q++) {
            double Tl[NUD], Th[NUD], mn[NG];
            int am[NG];
#pragma unroll
            for (int m = 0; m < NUD; m++) { Tl[m] = c.gWlo[m] * nd[q].cu[2 * m]; Th[m] = c.gWhi[m] * nd[q].cu[2 * m + 1]; }
#pragma unroll
            for (int g = 0; g < NG; g++) { mn[g] = CUDART_INF; am[g] = 0x7fffffff; }
            GridWalk<NUD, 0, 0, 0, ARG>::go(nd[q].S0, Tl, Th, mn, am);
#pragma unroll
            for (int g = 0; g < NG; g++) {
                const double rinv = rcp_pos(nd[q].norm0 + c.gAg[g]);
                const double ebt = !disc ? 1.0 : (tiny[q] ? exp_tiny(nbh * rinv) : exp_nonpos(nbh * rinv));
                const double v = rinv * (fma(ebt, mn[g], c.gHg[g]) + nd[q].hgx);
                if (v < nd[q].best || (ARG && v == nd[q].best && am[g] < nd[q].ibest)) { nd[q].best = v; nd[q].ibest = am[g]; }
            }
        }
#pragma unroll
        for (int q = 0; q < Q; q++) {
            if (valid[q]) {
                store_value(c, id[q], nd[q].best);
                if constexpr (ARG) { if (c.argmin) c.argmin[id[q]] = nd[q].ibest; }
            }
            if constexpr (ARG) {
                if (c.rows) {                               // policy row at u* (bellman.c:1851-1860)
                    // The rows are node-major records of 2dx+3 doubles: a thread storing its own record issues 2dx+3 stores that
                    // each touch 32 different sectors.  The warp's records go through shared memory instead and leave row by
                    // row, 2dx+3 consecutive lanes writing one record (184 contiguous bytes at dx = 10).
                    double *wr = STAGE ? srows + tid * RW : c.rows + (size_t)(valid[q] ? id[q] : 0) * RW;
                    if (valid[q]) {
                        const int ibest = nd[q].ibest;
                        double x[DX], u[DU], b[DX], sg[DX], prob[CS], dt;
                        node_state<DX>(c, id[q], x);
#pragma unroll
                        for (int i = 0; i < DU; i++) u[i] = P.utab[(size_t)(ibest < P.nu ? ibest : 0) * DU + i];
                        M::template drift<Fast>(x, u, P.mp, b);
                        M::template sigma<Fast>(x, u, P.mp, sg);
                        const double g = M::template stage<Fast>(x, u, P.mp);
                        if (transition_row<DX, Fast>(P, b, sg, prob, dt)) atomicOr(P.err, 1);
#pragma unroll
                        for (int m = 0; m < CS; m++) wr[m] = prob[m];
                        wr[CS] = dt;
                        wr[CS + 1] = g;
                    }
                    if constexpr (STAGE) {
                        __syncwarp();
                        const int idme = valid[q] ? id[q] : -1;
                        const double *wsrc = srows + (tid & ~31) * RW;
                        for (int r = 0; r < 32; r++) {
                            const int idr = __shfl_sync(0xffffffffu, idme, r);
                            if (idr >= 0) for (int m = lane; m < RW; m += 32) c.rows[(size_t)idr * RW + m] = wsrc[r * RW + m];
                        }
                        __syncwarp();
                    }
                }
            }
        }
    }
    // absorbed nodes (bellman.c:513-532): boundary / obstacle cost, u = 0
    for (long long id = (long long)blockIdx.x * blockDim.x + tid; id < c.NS; id += stride) {
        const int ab = c.flag[id];
        if (ab == 0) continue;
        double v = 0.0;                                     // ab == 2: padding entry j >= ngrid[dim_vary], defined as 0 / -1
        if (ab != 2) {
            double x[DX];
            node_state<DX>(c, (int)id, x);
            v = (ab == 1) ? M::boundcost(x, P.mp) : M::obscost(x, P.mp);
        }
        store_value(c, id, v);
        if (c.argmin) c.argmin[id] = -1;
        if (c.rows) {
            double *row = c.rows + (size_t)id * RW;
            for (int m = 0; m < RW; m++) row[m] = 0.0;
        }
    }
}

// ---------------------------------------------------------------------------
// FUSED stage 2: the node kernel of stage 1 (ft_mma_kernel.cuh) leaves the neighbour values of ITS OWN fibers in a
// small per-CTA region (slot-major, written minutes of nanoseconds ago: L2-hot) and calls this walk before it exits --
// no batch-sized cost scratch through HBM, no second launch.  The call crosses translation units (the node kernel is
// compiled per rank geometry, the walk per dynamics model): relocatable device code, one dispatcher per model family.

#ifndef C3SC_FUSED_Q
#define C3SC_FUSED_Q 2         // nodes per thread of the fused walk
#endif
template <class M, bool ARG>
__device__ void fused_walk_t(const CtlArgs &c, const FusedCta &w)
{
    constexpr int DX = M::DX, DU = M::DU, CS = 2 * DX + 1, RW = 2 * DX + 3;
    constexpr int NUD = M::NUD > 0 ? M::NUD : 1, NG = NUD + 1;
    constexpr bool GRID = M::SEP && M::NUD >= 1 && M::NUD <= CT_NUDMAX;
    constexpr int Q = GRID ? C3SC_FUSED_Q : 1;
    const DevProblem &P = c.P;
    const int nj = w.je - w.jb, total = w.nf * nj, N = P.ngrid[w.k], bk = P.bc[w.k];
    const double nbh = -P.beta * P.h2;
    const bool disc = P.beta != 0.0;
    const int per = (total + Q - 1) / Q;                    // thread t of round r owns nodes it, it + per, ..
    for (int it = threadIdx.x; it < per; it += blockDim.x) {
        bool live[Q], tiny[Q];                              // live: a real, non-absorbed node
        long long id[Q];
        // what the lock-step walk keeps per node: the control-independent sum, the per-control terms, the running minima
        double s0[Q], norm0[Q], hgx[Q], Tl[Q][NUD], Th[Q][NUD], mn[Q][NG];
        int am[Q][NG];
#pragma unroll
        for (int u = 0; u < Q; u++) {
            const int q = it + u * per;
            live[u] = false; tiny[u] = true;
            id[u] = 0;
            s0[u] = 0.0; norm0[u] = 1.0; hgx[u] = 0.0;
#pragma unroll
            for (int m = 0; m < NUD; m++) { Tl[u][m] = 0.0; Th[u][m] = 0.0; }
#pragma unroll
            for (int gq = 0; gq < NG; gq++) { mn[u][gq] = CUDART_INF; am[u][gq] = 0x7fffffff; }
            if (q >= total) continue;
            const int g = q / nj, jj = q - g * nj, j = w.jb + jj;
            const int ab = w.sAbs[g * w.nmax + j];
            id[u] = (long long)w.sFid[g] * c.ldo + j;
            double x[DX];
            node_state<DX>(c, (int)id[u], x);
            if (ab != 0) {                                  // absorbed (bellman.c:513-532): boundary / obstacle cost, u = 0
                store_value(c, id[u], (ab == 1) ? M::boundcost(x, P.mp) : M::obscost(x, P.mp));
                if (c.argmin) c.argmin[id[u]] = -1;
                if (c.rows) { double *row = c.rows + (size_t)id[u] * RW; for (int m = 0; m < RW; m++) row[m] = 0.0; }
                continue;
            }
            double cc[CS];
            int lo, hi;
            ft_vary_pair(bk, N, j, 0, lo, hi);
            const double *base = w.reg + g * w.njp + jj, *self = w.reg + (size_t)(2 * DX) * w.RN + g * w.njp;
#pragma unroll
            for (int i = 0; i < DX; i++) {
                cc[2 * i] = __ldcg(i == w.k ? self + (lo - w.jb) : base + (size_t)(2 * i) * w.RN);     // L2, not this SM's L1: the
                cc[2 * i + 1] = __ldcg(i == w.k ? self + (hi - w.jb) : base + (size_t)(2 * i + 1) * w.RN);   // region is recycled
            }
            cc[2 * DX] = __ldcg(self + jj);
            if (w.pi_eval) {                                // bellman.c:1863-1871: the stored row against the new values
                const double *row = c.rows_in + (size_t)id[u] * RW;
                double prob[CS];
#pragma unroll
                for (int m = 0; m < CS; m++) prob[m] = row[m];
                store_value(c, id[u], rhs<DX, Fast>(P, prob, row[CS], row[CS + 1], cc));
                continue;
            }
            if constexpr (GRID) {
                Node2<M> nd;
                node2_from<M>(c, x, cc, nd);
                if (nd.norm0 + P.amin < 1e-14) atomicOr(P.err, 1);
                live[u] = true;
                tiny[u] = P.beta * P.h2 <= 0.00390625 * (nd.norm0 + P.amin);
                s0[u] = nd.S0; norm0[u] = nd.norm0; hgx[u] = nd.hgx;
#pragma unroll
                for (int m = 0; m < NUD; m++) { Tl[u][m] = c.gWlo[m] * nd.cu[2 * m]; Th[u][m] = c.gWhi[m] * nd.cu[2 * m + 1]; }
            }
        }
        if constexpr (GRID) {
            if (w.pi_eval) continue;
            GridWalkQ<NUD, 0, 0, 0, ARG, Q>::go(s0, Tl, Th, mn, am);
#pragma unroll
            for (int u = 0; u < Q; u++) {
                if (!live[u]) continue;
                double best = CUDART_INF;
                int ibest = 0x7fffffff;
#pragma unroll
                for (int gq = 0; gq < NG; gq++) {
                    const double rinv = rcp_pos(norm0[u] + c.gAg[gq]);
                    const double ebt = !disc ? 1.0 : (tiny[u] ? exp_tiny(nbh * rinv) : exp_nonpos(nbh * rinv));
                    const double v = rinv * (fma(ebt, mn[u][gq], c.gHg[gq]) + hgx[u]);
                    if (v < best || (ARG && v == best && am[u][gq] < ibest)) { best = v; ibest = am[u][gq]; }
                }
                store_value(c, id[u], best);
                if constexpr (ARG) {
                    if (c.argmin) c.argmin[id[u]] = ibest;
                    if (c.rows) {                           // policy row at u* (bellman.c:1851-1860)
                        double x[DX], uu[DU], b[DX], sg[DX], prob[CS], dt;
                        node_state<DX>(c, (int)id[u], x);
#pragma unroll
                        for (int i = 0; i < DU; i++) uu[i] = P.utab[(size_t)(ibest < P.nu ? ibest : 0) * DU + i];
                        M::template drift<Fast>(x, uu, P.mp, b);
                        M::template sigma<Fast>(x, uu, P.mp, sg);
                        const double gs = M::template stage<Fast>(x, uu, P.mp);
                        if (transition_row<DX, Fast>(P, b, sg, prob, dt)) atomicOr(P.err, 1);
                        double *row = c.rows + (size_t)id[u] * RW;
#pragma unroll
                        for (int m = 0; m < CS; m++) row[m] = prob[m];
                        row[CS] = dt;
                        row[CS + 1] = gs;
                    }
                }
            }
        }
    }
    // padding entries j >= ngrid[k] of ragged grids: defined content
    const int pad = c.ldo - N;
    if (pad > 0 && w.jb == 0)
        for (int q = threadIdx.x; q < w.nf * pad; q += blockDim.x) {
            const int g = q / pad;
            const long long id = (long long)w.sFid[g] * c.ldo + N + (q - g * pad);
            store_value(c, id, 0.0);
            if (c.argmin) c.argmin[id] = -1;
            if (c.rows) { double *row = c.rows + (size_t)id * RW; for (int m = 0; m < RW; m++) row[m] = 0.0; }
        }
}
template <class M>
__device__ void fused_walk_m(const CtlArgs &c, const FusedCta &w)
{
    if (w.arg && !w.pi_eval) fused_walk_t<M, true>(c, w);
    else fused_walk_t<M, false>(c, w);
}
// can a problem's stage 2 run fused?  (host side; the same test picks the kernel in launch_control_t)
template <class M>
inline bool fused_ok_m(int arith, const CtlArgs &c, int pi_eval)
{
    if (pi_eval) return arith == C3SC_ARITH_FAST;
    return arith == C3SC_ARITH_FAST && M::SEP && M::NUD >= 1 && M::NUD <= CT_NUDMAX && c.grid_on && !getenv("C3SC_NO_GRID");
}

// policy evaluation (bellman.c:1774-1828,1863-1871): stored rows against the new neighbour values.
// The rows are node-major records of 2dx+3 doubles (the reference's layout).  The records of PE_NT consecutive nodes are ONE
// contiguous piece (23.5 kB at dx = 10): a CTA fetches it with one TMA bulk copy (cp.async.bulk + mbarrier) into one of two
// shared-memory buffers, one tile ahead of the tile it works on, and every thread takes its node's record from there (odd
// stride in doubles: no bank conflicts).  The kernel is HBM-bound (184 B of rows + 168 B of neighbour values per node); the
// first version staged the tile with a load -> store loop per thread, four or five loads in flight, and reached 47 % of the
// DRAM throughput with 81 % of its stall samples on those loads (profiles/r02b_logs/r02b_stalls_k_pi_eval.txt).
constexpr int PE_NT = 128;
__device__ __forceinline__ unsigned pe_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <class M, class A>
__global__ void __launch_bounds__(PE_NT) k_pi_eval(const CtlArgs c)
{
    constexpr int DX = M::DX, CS = 2 * DX + 1, RW = 2 * DX + 3;
    constexpr unsigned TILE_BYTES = PE_NT * RW * sizeof(double);
    constexpr bool BULK = TILE_BYTES % 16 == 0 && 2 * TILE_BYTES + 64 <= 48 * 1024;
    __shared__ __align__(16) double srow[(BULK ? 2 : 1) * PE_NT * RW];
    __shared__ unsigned long long bar[2];
    const DevProblem &P = c.P;
    const int tid = threadIdx.x;
    const long long tstride = (long long)gridDim.x * PE_NT;
    const bool aligned = ((size_t)c.rows_in & 15) == 0;
    auto fetch = [&](long long base, int buf) {             // tid 0: the full tile at `base` -> buffer buf
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pe_smem_u32(bar + buf)), "r"(TILE_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(pe_smem_u32(srow + buf * PE_NT * RW)), "l"(c.rows_in + (size_t)base * RW), "r"(TILE_BYTES),
                       "r"(pe_smem_u32(bar + buf)) : "memory");
    };
    if (BULK && tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pe_smem_u32(bar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pe_smem_u32(bar + 1)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const long long b0 = (long long)blockIdx.x * PE_NT;
        if (aligned && b0 + PE_NT <= c.NS) fetch(b0, 0);
    }
    __syncthreads();
    unsigned phase = 0;
    int buf = 0;
    for (long long base = (long long)blockIdx.x * PE_NT; base < c.NS; base += tstride, buf ^= 1) {
        const long long left = c.NS - base;
        const int cnt = left < PE_NT ? (int)left : PE_NT;
        const bool bulk = BULK && aligned && cnt == PE_NT;
        const double *tile = srow + (BULK ? buf : 0) * PE_NT * RW;
        if (BULK) {
            // the other buffer was read by the previous trip (barrier at its end): the next full tile goes there now
            const long long nb = base + tstride;
            if (tid == 0 && aligned && nb + PE_NT <= c.NS) fetch(nb, buf ^ 1);
        }
        if (bulk) {
            const unsigned par = (phase >> buf) & 1u;
            asm volatile("{\n.reg .pred P1;\nPE_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra PE_DONE;\nbra PE_WAIT;\nPE_DONE:\n}"
                         ::"r"(pe_smem_u32(bar + buf)), "r"(par) : "memory");
            phase ^= 1u << buf;
        } else {                                            // the ragged last tile (or an unaligned buffer): plain loads
            const double *src = c.rows_in + (size_t)base * RW;
            double *dstt = srow + (BULK ? buf : 0) * PE_NT * RW;
            if (!BULK) __syncthreads();
            for (int e = tid; e < cnt * RW; e += PE_NT) dstt[e] = src[e];
            __syncthreads();
        }
        const long long id = base + tid;
        if (tid < cnt) {
            const int ab = c.flag[id];
            double v;
            if (ab == 2) v = 0.0;                           // padding entry: defined content
            else if (ab != 0) {
                double x[DX];
                node_state<DX>(c, (int)id, x);
                v = (ab == 1) ? M::boundcost(x, P.mp) : M::obscost(x, P.mp);
            } else {
                const double *row = tile + tid * RW;
                double prob[CS], cc[CS];
                load_costs<DX>(c, (int)id, cc);
#pragma unroll
                for (int m = 0; m < CS; m++) prob[m] = row[m];
                v = rhs<DX, A>(P, prob, row[CS], row[CS + 1], cc);
            }
            store_value(c, id, v);
        }
        if (BULK) __syncthreads();                          // every thread is done with this buffer
    }
}

// Candidate table of a separable model (FAST only): row c = [Wl_0, Wr_0, .., Wl_{NUD-1}, Wr_{NUD-1}, A, gu]
//   Wl_m / Wr_m = t_i*|b_i(u_c)| on the side the upwind scheme adds it to (0 inside the 1e-14 dead band),
//   A = sum of the W's (the candidate's share of the normaliser), gu = h2-free stage_u(u_c).
template <class M>
__global__ void k_build_ctab(const DevProblem P, double *ctab)
{
    constexpr int DX = M::DX, DU = M::DU, NUD = M::NUD, CT = 2 * NUD + 2;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.nu) return;
    double x[DX], u[DU], b[DX];
    for (int i = 0; i < DX; i++) x[i] = 0.0;           // SEP: drift of ud(m) does not read x
    for (int i = 0; i < DU; i++) u[i] = P.utab[(size_t)c * DU + i];
    M::template drift<Fast>(x, u, P.mp, b);
    double A = 0.0;
    for (int m = 0; m < NUD; m++) {
        const int i = M::ud(m);
        const double tb = P.t[2 * i] * b[i];
        const double wl = (b[i] < -1e-14) ? -tb : 0.0, wr = (b[i] > 1e-14) ? tb : 0.0;
        ctab[(size_t)c * CT + 2 * m] = wl;
        ctab[(size_t)c * CT + 2 * m + 1] = wr;
        A += wl + wr;
    }
    ctab[(size_t)c * CT + 2 * NUD] = A;
    ctab[(size_t)c * CT + 2 * NUD + 1] = M::stage_u(u, P.mp);
}

template <class M>
int build_ctab_t(const DevProblem &P, double *ctab, cudaStream_t st)
{
    if (!M::SEP) return 0;
    k_build_ctab<M><<<(P.nu + 127) / 128, 128, 0, st>>>(P, ctab);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------
struct CtlLaunchInfo { int sms, max_optin; };
inline const CtlLaunchInfo &ctl_info()
{
    static CtlLaunchInfo info = {0, 0};
    if (!info.sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&info.sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&info.max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    return info;
}

template <class M, class A>
int launch_control_t(const CtlArgs &c_in, int pi_eval, cudaStream_t st)
{
    const CtlLaunchInfo &info = ctl_info();
    CtlArgs c = c_in;
    if (pi_eval) {
        long long g = (c.NS + PE_NT - 1) / PE_NT;
        if (g > (long long)info.sms * 8) g = (long long)info.sms * 8;
        if (g < 1) return 0;
        k_pi_eval<M, A><<<(int)g, PE_NT, 0, st>>>(c);
        return (int)cudaGetLastError();
    }
    constexpr bool TAB = M::SEP && !A::exact;
    const size_t smem = (size_t)c.P.nu * (TAB ? 2 * M::NUD + 2 : M::DU) * sizeof(double);
    if (smem > (size_t)info.max_optin) return (int)cudaErrorInvalidValue;      // control table must fit on chip
    static size_t attr_dev[C3SC_MAXDEV] = {0};
    size_t &attr_set = attr_dev[c3sc_cur_dev()];
    if (attr_set == 0) attr_set = 48 * 1024;
    if (smem > attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_control<M, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = smem;
    }
    // candidate chunks per node: enough (node, chunk) items to occupy every SM; 1 for large batches
    int pl2 = 0;
    while (pl2 < 5 && (c.NS << pl2) < (long long)info.sms * CT_NT * 8 && (2 << pl2) <= c.P.nu) pl2++;
    c.parts_log2 = pl2;
    if constexpr (TAB && M::NUD >= 1 && M::NUD <= CT_NUDMAX) {
        if (c.grid_on && !getenv("C3SC_NO_GRID")) {
            // grid-structured table: the shared-prefix walk, one node per thread until the batch fills the part twice
            const bool arg = c.argmin || c.rows;
            const bool two = C3SC_GRID_Q == 2 && c.NS >= (long long)info.sms * 4 * CT_NT;
            const int q = two ? 2 : 1;
            const int nt = (c.NS / q >= (long long)info.sms * 2 * CT_NT) ? CT_NT : 128;
            long long need = ((c.NS + q - 1) / q + nt - 1) / nt;
            long long g = (long long)info.sms * grid_minb(arg) * (CT_NT / nt);
            if (g > need) g = need;
            if (g < 1) return 0;
            if (arg) { if (two) k_control_grid<M, true, 2><<<(int)g, nt, 0, st>>>(c); else k_control_grid<M, true, 1><<<(int)g, nt, 0, st>>>(c); }
            else     { if (two) k_control_grid<M, false, 2><<<(int)g, nt, 0, st>>>(c); else k_control_grid<M, false, 1><<<(int)g, nt, 0, st>>>(c); }
            return (int)cudaGetLastError();
        }
    }
    if (TAB && c.ng > 0 && c.NS >= (long long)info.sms * 128) {
        // separable model, at least ~a warp pair of nodes per SM: two nodes per thread walk the whole
        // grouped table (no per-chunk re-derivation of the node invariants); 128-thread CTAs while the
        // batch is too small to give every SM a 256-thread one
        static size_t attr2_dev[C3SC_MAXDEV] = {0};
        size_t &attr2 = attr2_dev[c3sc_cur_dev()];
        if (attr2 == 0) attr2 = 48 * 1024;
        if (smem > attr2) {
            cudaError_t e2 = cudaFuncSetAttribute(k_control2<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e2 != cudaSuccess) return (int)e2;
            attr2 = smem;
        }
        const int nt2 = (c.NS / C2N >= (long long)info.sms * 2 * C2_NT) ? C2_NT : (C2_NT < 128 ? C2_NT : 128);
        int per2 = 1;
        cudaError_t e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per2, k_control2<M>, nt2, smem);
        if (e2 != cudaSuccess) return (int)e2;
        if (per2 < 1) per2 = 1;
        long long need2 = ((c.NS + C2N - 1) / C2N + nt2 - 1) / nt2;
        long long g2 = (long long)info.sms * per2;
        if (g2 > need2) g2 = need2;
        if (g2 < 1) return 0;
        k_control2<M><<<(int)g2, nt2, smem, st>>>(c);
        return (int)cudaGetLastError();
    }
    int per_sm = 1;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_control<M, A>, CT_NT, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    long long need = ((c.NS << pl2) + CT_NT - 1) / CT_NT;
    long long g = (long long)info.sms * per_sm;
    if (g > need) g = need;
    if (g < 1) return 0;
    k_control<M, A><<<(int)g, CT_NT, smem, st>>>(c);
    return (int)cudaGetLastError();
}

template <class M>
int launch_control_m(int arith, const CtlArgs &c, int pi_eval, cudaStream_t st)
{
    if (arith == C3SC_ARITH_EXACT) return launch_control_t<M, Exact>(c, pi_eval, st);
    return launch_control_t<M, Fast>(c, pi_eval, st);
}

// ---- node-level entry: bellman_optimal on caller-supplied (x, neighbour costs, flag) --------
// One thread per node, candidates walked in table order (bellman.c:504-543).
template <class M, class A>
__global__ void k_node_backup(const DevProblem P, int n, const double *x, const double *costs, const int *absorbed,
                              double *value, int *argmin)
{
    constexpr int DX = M::DX, DU = M::DU, CS = 2 * DX + 1;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double xx[DX], c[CS];
#pragma unroll
    for (int i = 0; i < DX; i++) xx[i] = x[(size_t)e * DX + i];
    const int ab = absorbed ? absorbed[e] : 0;
    if (ab == 1) { value[e] = M::boundcost(xx, P.mp); if (argmin) argmin[e] = -1; return; }
    if (ab == -1) { value[e] = M::obscost(xx, P.mp); if (argmin) argmin[e] = -1; return; }
#pragma unroll
    for (int m = 0; m < CS; m++) c[m] = costs[(size_t)e * CS + m];
    NodeInv<M> inv;
    node_prepare<M, A>(P, P.utab, xx, c, inv);
    double best = CUDART_INF;
    int ibest = 0x7fffffff, bad = 0;
    for (int cand = 0; cand < P.nu; cand++) {
        double u[DU];
#pragma unroll
        for (int i = 0; i < DU; i++) u[i] = P.utab[(size_t)cand * DU + i];
        const double v = candidate_value<M, A>(P, xx, u, c, inv, bad);
        if (v < best) { best = v; ibest = cand; }
    }
    if (bad) atomicOr(P.err, 1);
    value[e] = best;
    if (argmin) argmin[e] = ibest;
}

template <class M>
int launch_node_backup_t(int arith, const DevProblem &P, int n, const double *x, const double *costs,
                         const int *absorbed, double *value, int *argmin, cudaStream_t st)
{
    if (n <= 0) return 0;
    const int g = (n + 127) / 128;
    if (arith == C3SC_ARITH_EXACT) k_node_backup<M, Exact><<<g, 128, 0, st>>>(P, n, x, costs, absorbed, value, argmin);
    else k_node_backup<M, Fast><<<g, 128, 0, st>>>(P, n, x, costs, absorbed, value, argmin);
    return (int)cudaGetLastError();
}

// bellman_control (bellman.c:367-480, grad_u == NULL) at n (x, u, costs) triples, any u
template <class M, class A>
__global__ void k_control_value(const DevProblem P, int n, const double *x, const double *u, const double *costs,
                                double *value)
{
    constexpr int DX = M::DX, DU = M::DU, CS = 2 * DX + 1;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double xx[DX], uu[DU], c[CS], b[DX], s[DX], prob[CS], dt;
    for (int i = 0; i < DX; i++) xx[i] = x[(size_t)e * DX + i];
    for (int i = 0; i < DU; i++) uu[i] = u[(size_t)e * DU + i];
    for (int m = 0; m < CS; m++) c[m] = costs[(size_t)e * CS + m];
    M::template drift<A>(xx, uu, P.mp, b);
    M::template sigma<A>(xx, uu, P.mp, s);
    const double g = M::template stage<A>(xx, uu, P.mp);
    if (transition_row<DX, A>(P, b, s, prob, dt)) { atomicOr(P.err, 1); value[e] = CUDART_NAN; return; }
    value[e] = rhs<DX, A>(P, prob, dt, g, c);
}
template <class M>
int launch_control_value_t(int arith, const DevProblem &P, int n, const double *x, const double *u, const double *costs,
                           double *value, cudaStream_t st)
{
    if (n <= 0) return 0;
    const int g = (n + 127) / 128;
    if (arith == C3SC_ARITH_EXACT) k_control_value<M, Exact><<<g, 128, 0, st>>>(P, n, x, u, costs, value);
    else k_control_value<M, Fast><<<g, 128, 0, st>>>(P, n, x, u, costs, value);
    return (int)cudaGetLastError();
}

// bellmanrhs on raw tuples (runtime dx)
template <class A>
__global__ void k_rhs(int dx, double beta, int n, const double *prob, const double *dt, const double *stage,
                      const double *cost, double *out)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int cs = 2 * dx + 1;
    const double ebt = exp(A::mul(-beta, dt[e]));
    double ctg = 0.0;
    for (int m = 0; m < cs; m++) ctg = A::mad(prob[(size_t)e * cs + m], cost[(size_t)e * cs + m], ctg);
    out[e] = A::add(A::mul(dt[e], stage[e]), A::mul(ebt, ctg));
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
