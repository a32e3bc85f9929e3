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
#include <cstdlib>
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
constexpr int PMAX = 8;              // max candidate chunks ("parts") per node
constexpr int TMAX = 16;             // max nodes per FT tile
constexpr int NSTAGE = 3;            // depth of the TMA staging ring

struct SmemPlan {
    int rs;        // padded (odd) vector stride
    int cs;        // cost-row stride (2dx+1, odd already)
    int ns;        // per-node invariant stride (odd)
    int nv;        // vectors per chain set: 1 + 2(dx-1)
    int wst;       // stride of one node's (w,u) pair in the FT tile
    int oVL, oVR, oC, oV, oScr, nDoubles;
    int oWt;                       // FT phase view of the scratch region
    int oNode, oBestV, oBestI;     // control phase view of the scratch region (oBestI in doubles)
    int oRing, oBar, cb, slot;     // TMA staging ring (staged plan only): NSTAGE slots of `slot` doubles
    int oAbs, oNv, oAct, oNf, oFix, oMisc, nInts;
    __host__ __device__ SmemPlan(int dx, int nud, int nmax, int rmax, bool staged)
    {
        rs = odd_up(rmax);
        cs = 2 * dx + 1;
        ns = odd_up(2 * nud + 3);
        nv = 2 * dx - 1;
        wst = 2 * rs + 1 - ((2 * rs + 1) & 1) + 1;     // odd
        int o = 0;
        oVL = o;   o += 2 * nv * rs;
        oVR = o;   o += 2 * nv * rs;
        oC = o;    o += nmax * cs;
        oV = o;    o += nmax;
        oScr = o;
        oWt = oScr;
        const int ft_need = TMAX * wst;
        oNode = oScr; oBestV = oNode + nmax * ns; oBestI = oBestV + PMAX * nmax;
        const int ctl_need = nmax * ns + PMAX * nmax + (PMAX * nmax + 1) / 2;
        o += ft_need > ctl_need ? ft_need : ctl_need;
        o += o & 1;                                     // 16-byte alignment of the ring
        cb = (rmax * rmax + 2 + 1) & ~1;               // one block + alignment slack, even
        slot = 6 * cb;                                  // a chain step stages <= 6 blocks
        oRing = o;
        oBar = o;
        if (staged) { o += NSTAGE * slot; oBar = o; o += NSTAGE + (NSTAGE & 1); }
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

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers -------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity)
{
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "LAB_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 "@P1 bra DONE;\n"
                 "bra LAB_WAIT;\n"
                 "DONE:\n"
                 "}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
// length-n dot with strides, four independent partial sums (breaks the DFMA dependency chain so
// the shared-memory loads of four terms are in flight together)
__device__ __forceinline__ double dot4(const double *x, int sx, const double *y, int sy, int n)
{
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int i = 0;
    for (; i + 4 <= n; i += 4) {
        a0 = fma(x[i * sx], y[i * sy], a0);
        a1 = fma(x[(i + 1) * sx], y[(i + 1) * sy], a1);
        a2 = fma(x[(i + 2) * sx], y[(i + 2) * sy], a2);
        a3 = fma(x[(i + 3) * sx], y[(i + 3) * sy], a3);
    }
    for (; i < n; i++) a0 = fma(x[i * sx], y[i * sy], a0);
    return (a0 + a1) + (a2 + a3);
}
__host__ __device__ inline int gcd16(int r) { int g = 1; while (g < 16 && (r % (2 * g)) == 0) g *= 2; return g; }

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

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
#define PHASE_MARK(slot)                                                        \
    do {                                                                        \
        if (a.prof && tid == 0) {                                               \
            const long long now_ = clock64();                                   \
            pacc[slot] += (unsigned long long)(now_ - tmark);                   \
            tmark = now_;                                                       \
        }                                                                       \
    } while (0)

// Arm one ring slot: the bulk copies of item `item` of the current fiber (thread 0 only).
// Chain step s stages, per active side, the centre block and the two neighbour blocks of the
// dimension it consumes; a node tile stages nt consecutive blocks of the varying core.  Sources
// are widened to 16-byte boundaries (cp.async.bulk needs it); consumers add (offset & 1).
__device__ __forceinline__ void ring_issue(const DevFT &ft, int item, int nsteps, int k, int DX, int N, int T,
                                        const int *sFix, const int *sNf, double *dst, int cb,
                                        unsigned long long *bar)
{
    unsigned total = 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) mbar_expect_tx(bar, total);
        if (item < nsteps) {
#pragma unroll 1
            for (int c = 0; c < 6; c++) {
                const int side = c / 3, w = c - 3 * side;
                const bool on = side == 0 ? (item < k) : (item < DX - 1 - k);
                if (!on) continue;
                const int m = side == 0 ? item : DX - 1 - item;
                const int bl = ft.r[m] * ft.r[m + 1];
                const int slotn = m < k ? m : m - 1;
                const int jn = w == 0 ? sFix[m] : sNf[2 * slotn + (w - 1)];
                const long long s0 = ft.off[m] + (long long)jn * bl;
                const long long sa = s0 & ~1LL, ea = (s0 + bl + 1) & ~1LL;
                if (pass == 0) total += (unsigned)((ea - sa) * 8);
                else bulk_g2s(dst + c * cb, ft.base + sa, (unsigned)((ea - sa) * 8), bar);
            }
        } else {
            const int blk = ft.r[k] * ft.r[k + 1];
            const int j0 = (item - nsteps) * T;
            const int nt = (N - j0 < T) ? N - j0 : T;
            const long long s0 = ft.off[k] + (long long)j0 * blk;
            const long long sa = s0 & ~1LL, ea = (s0 + (long long)nt * blk + 1) & ~1LL;
            if (pass == 0) total += (unsigned)((ea - sa) * 8);
            else bulk_g2s(dst, ft.base + sa, (unsigned)((ea - sa) * 8), bar);
        }
    }
}

template <class M, class A>
__global__ void __launch_bounds__(NT, A::exact ? 1 : 2) k_backup(const LaunchArgs a)
{
    unsigned long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tmark = clock64();
    constexpr int DX = M::DX, DU = M::DU, CS = 2 * DX + 1, RW = 2 * DX + 3;
    const DevProblem &P = a.P;
    const DevFT &ft = a.ft;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    extern __shared__ __align__(16) double smem[];
    const SmemPlan sp(DX, M::NUD, P.nmax, ft.rmax, a.staged != 0);
    double *sVL = smem + sp.oVL, *sVR = smem + sp.oVR, *sC = smem + sp.oC, *sV = smem + sp.oV;
    double *sWt = smem + sp.oWt;
    const double *sLset = sVL, *sRset = sVR;
    const int NV = sp.nv, wst = sp.wst;
    double *ring = smem + sp.oRing;
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem + sp.oBar);
    unsigned git = 0;                      // running item counter of the staging ring (CTA-uniform)
    if (a.staged) {
        if (tid == 0) {
            for (int q = 0; q < NSTAGE; q++) mbar_init(mbar + q, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    double *sNode = smem + sp.oNode, *sBestV = smem + sp.oBestV;
    int *sBestI = reinterpret_cast<int *>(smem + sp.oBestI);
    int *ismem = reinterpret_cast<int *>(smem + sp.nDoubles);
    int *sAbs = ismem + sp.oAbs, *sNv = ismem + sp.oNv, *sAct = ismem + sp.oAct, *sNf = ismem + sp.oNf;
    int *sMisc = ismem + sp.oMisc, *sFix = ismem + sp.oFix;
    const int rs = sp.rs;

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

        if (a.nbr_fixed_in) {                  // valuef_eval_fiber_ind_nn: indices come from the caller
            __syncthreads();
            for (int e = tid; e < 2 * (DX - 1); e += NT) sNf[e] = a.nbr_fixed_in[(size_t)f * 2 * (DX - 1) + e];
            for (int e = tid; e < 2 * N; e += NT) sNv[e] = a.nbr_vary_in[2 * obase + e];
        }
        __syncthreads();                       // sFix / sNf / flags visible
        PHASE_MARK(0);

        // ---- 2. function-train neighbour values (valuefunc.c:369-585, re-associated) ----
        // 2a. stepped chains.  Left set after step m:  {L_{m+1}} U {a_i^-, a_i^+ : i <= m}, all row
        //     vectors of length r_{m+1};  a_i^s = L_i G_i[nb_s] G_{i+1}[f] .. G_m[f].
        //     Right set (dimensions walked from d-1 down): {R_m} U {c_i^s : i >= m}, columns of
        //     length r_m;  c_i^s = G_m[f] .. G_{i-1}[f] G_i[nb_s] R_{i+1}.
        //     One step = every vector of a set times one centre block + two neighbour blocks of
        //     the prefix/suffix vector: a small GEMM spread over all threads, depth max(k, d-1-k).
        const int rk = ft.r[k], rk1 = ft.r[k + 1];
#define CORE_BLK(i, j) (ft.base + ft.off[i] + (size_t)(j) * ft.r[i] * ft.r[(i) + 1])
        if (a.staged) {
            // ---- staged plan: every FT operand arrives in shared memory by TMA bulk copy --------
            // item stream of a fiber = chain steps 0..nsteps-1, then node tiles; item i lives in ring
            // slot (git0+i) % NSTAGE; thread 0 re-arms a slot as soon as the CTA is done with it.
            const int nsteps = (k > DX - 1 - k) ? k : DX - 1 - k;
            const int blk = rk * rk1, CBs = sp.cb, SL = sp.slot;
            int T = (SL - 2) / blk;
            T = T < 1 ? 1 : (T > TMAX ? TMAX : T);
            const int ntiles = (N + T - 1) / T, nitems = nsteps + ntiles;
            const unsigned git0 = git;
            auto issue = [&](int item) {
                const unsigned g = git0 + (unsigned)item;
                ring_issue(ft, item, nsteps, k, DX, N, T, sFix, sNf, ring + (g % NSTAGE) * SL, CBs, mbar + (g % NSTAGE));
            };
            if (tid == 0) {
                sVL[0] = 1.0; sVR[0] = 1.0;
                for (int it = 0; it < NSTAGE && it < nitems; it++) issue(it);
            }
            int cur = 0;
            __syncthreads();
            for (int s = 0; s < nsteps; s++, git++) {
                const double *inL = sVL + cur * NV * rs, *inR = sVR + cur * NV * rs;
                double *outL = sVL + (cur ^ 1) * NV * rs, *outR = sVR + (cur ^ 1) * NV * rs;
                const bool doL = s < k, doR = s < DX - 1 - k;
                const int mL = s, mR = DX - 1 - s, nin = 1 + 2 * s;
                const int nl = doL ? (nin + 2) * ft.r[mL + 1] : 0;
                const int nr = doR ? (nin + 2) * ft.r[mR] : 0;
                const double *sl = ring + (git % NSTAGE) * SL;
                mbar_wait(mbar + (git % NSTAGE), (git / NSTAGE) & 1);
                for (int e = tid; e < nl + nr; e += NT) {
                    if (e < nl) {
                        const int m = mL, rm = ft.r[m], nout = nin + 2, bl = rm * ft.r[m + 1];
                        const int b = e / nout, v = e - b * nout;
                        const int w = v < nin ? 0 : 1 + (v - nin);
                        const int jn = w == 0 ? sFix[m] : sNf[2 * m + (w - 1)];
                        const double *G = sl + w * CBs + (int)((ft.off[m] + (long long)jn * bl) & 1);
                        const double *vec = inL + (v < nin ? v : 0) * rs;
                        outL[v * rs + b] = dot4(vec, 1, G + b * rm, 1, rm);
                    } else {
                        const int e2 = e - nl, m = mR, rm = ft.r[m], rm1 = ft.r[m + 1], nout = nin + 2, bl = rm * rm1;
                        const int aa = e2 / nout, v = e2 - aa * nout;
                        const int w = v < nin ? 0 : 1 + (v - nin);
                        const int jn = w == 0 ? sFix[m] : sNf[2 * (m - 1) + (w - 1)];
                        const double *G = sl + (3 + w) * CBs + (int)((ft.off[m] + (long long)jn * bl) & 1);
                        const double *vec = inR + (v < nin ? v : 0) * rs;
                        outR[v * rs + aa] = dot4(G + aa, rm, vec, 1, rm1);
                    }
                }
                if (!doL && k > 0) for (int e = tid; e < (1 + 2 * k) * rs; e += NT) outL[e] = inL[e];
                if (!doR && DX - 1 - k > 0) for (int e = tid; e < (1 + 2 * (DX - 1 - k)) * rs; e += NT) outR[e] = inR[e];
                if (!doL && k == 0 && tid == 0) outL[0] = 1.0;
                if (!doR && DX - 1 - k == 0 && tid == 0) outR[0] = 1.0;
                cur ^= 1;
                __syncthreads();
                if (tid == 0 && s + NSTAGE < nitems) issue(s + NSTAGE);
            }
            sLset = sVL + cur * NV * rs;
            sRset = sVR + cur * NV * rs;
            PHASE_MARK(1);
            {
                const double *Lk = sLset, *Rk1 = sRset;
                const int gk = gcd16(rk), per16 = 16 / gk;
                for (int t = 0; t < ntiles; t++, git++) {
                    const int j0 = t * T, nt = (N - j0 < T) ? N - j0 : T;
                    const double *gt = ring + (git % NSTAGE) * SL + (int)((ft.off[k] + (long long)j0 * blk) & 1);
                    long long t0_ = (a.prof && tid == 0) ? clock64() : 0;
                    mbar_wait(mbar + (git % NSTAGE), (git / NSTAGE) & 1);
                    if (a.prof && tid == 0) { const long long t1_ = clock64(); pacc[6] += (unsigned long long)(t1_ - t0_); t0_ = t1_; }
                    for (int e = tid; e < nt * rk; e += NT) {          // w[a] = sum_b G[a + b*rk] R[b]
                        const int jl = e / rk, aa = e - jl * rk;
                        sWt[jl * wst + aa] = dot4(gt + jl * blk + aa, rk, Rk1, 1, rk1);
                    }
                    for (int e = tid; e < nt * rk1; e += NT) {         // u[b] = sum_a L[a] G[a + b*rk]
                        const int jl = e / rk1, b = e - jl * rk1;
                        const double *col = gt + jl * blk + b * rk;
                        const int q = (b / per16) % gk;                 // start skew: conflict-free column walks
                        sWt[jl * wst + rs + b] = dot4(Lk + q, 1, col + q, 1, rk - q) + dot4(Lk, 1, col, 1, q);
                    }
                    __syncthreads();
                    if (a.prof && tid == 0) { pacc[7] += (unsigned long long)(clock64() - t0_); }
                    if (tid == 0 && nsteps + t + NSTAGE < nitems) issue(nsteps + t + NSTAGE);
                    for (int e = tid; e < nt * (2 * DX - 1); e += NT) {
                        const int q = e / nt, jl = e - q * nt, j = j0 + jl;
                        const double *w = sWt + jl * wst, *u = w + rs;
                        double acc = 0.0;
                        int oslot;
                        if (q == 0) {
                            acc = dot4(Lk, 1, w, 1, rk);
                            oslot = 2 * DX;
                            sV[j] = acc;
                        } else {
                            const int fs = q - 1, slot = fs >> 1, side = fs & 1;
                            const int i = slot < k ? slot : slot + 1;
                            if (i < k) acc = dot4(sLset + (1 + 2 * i + side) * rs, 1, w, 1, rk);
                            else       acc = dot4(u, 1, sRset + (1 + 2 * (DX - 1 - i) + side) * rs, 1, rk1);
                            oslot = 2 * i + side;
                        }
                        sC[j * CS + oslot] = acc;
                    }
                    __syncthreads();
                }
            }
        } else {
        {
            if (tid == 0) { sVL[0] = 1.0; sVR[0] = 1.0; }
            const int nsteps = (k > DX - 1 - k) ? k : DX - 1 - k;
            int cur = 0;
            __syncthreads();
            for (int s = 0; s < nsteps; s++) {
                const double *inL = sVL + cur * NV * rs, *inR = sVR + cur * NV * rs;
                double *outL = sVL + (cur ^ 1) * NV * rs, *outR = sVR + (cur ^ 1) * NV * rs;
                // left side, dimension m = s
                const bool doL = s < k, doR = s < DX - 1 - k;
                const int mL = s, mR = DX - 1 - s;
                const int ninL = 1 + 2 * s, ninR = 1 + 2 * s;
                const int nl = doL ? (ninL + 2) * ft.r[mL + 1] : 0;
                const int nr = doR ? (ninR + 2) * ft.r[mR] : 0;
                if (s + 1 < nsteps) {              // pull the next step's blocks towards L1
                    const int nxt = s + 1;
                    if (nxt < k) {
                        const int bl = ft.r[nxt] * ft.r[nxt + 1];
                        const int slot = nxt;     // nxt < k
                        const double *b0 = CORE_BLK(nxt, sFix[nxt]), *b1 = CORE_BLK(nxt, sNf[2 * slot]), *b2 = CORE_BLK(nxt, sNf[2 * slot + 1]);
                        for (int e = tid * 16; e < bl; e += NT * 16) { prefetch_l1(b0 + e); prefetch_l1(b1 + e); prefetch_l1(b2 + e); }
                    }
                    const int nxr = DX - 1 - nxt;
                    if (nxt < DX - 1 - k) {
                        const int bl = ft.r[nxr] * ft.r[nxr + 1];
                        const int slot = nxr - 1;  // nxr > k
                        const double *b0 = CORE_BLK(nxr, sFix[nxr]), *b1 = CORE_BLK(nxr, sNf[2 * slot]), *b2 = CORE_BLK(nxr, sNf[2 * slot + 1]);
                        for (int e = tid * 16; e < bl; e += NT * 16) { prefetch_l1(b0 + e); prefetch_l1(b1 + e); prefetch_l1(b2 + e); }
                    }
                }
                for (int e = tid; e < nl + nr; e += NT) {
                    if (e < nl) {
                        const int m = mL, rm = ft.r[m], nout = ninL + 2;
                        const int b = e / nout, v = e - b * nout;            // v fastest: block reads broadcast
                        const double *G = (v < ninL) ? CORE_BLK(m, sFix[m]) : CORE_BLK(m, sNf[2 * m + (v - ninL)]);
                        const double *vec = inL + ((v < ninL) ? v : 0) * rs;
                        const double *col = G + (size_t)b * rm;
                        double acc = 0.0;
                        for (int a = 0; a < rm; a++) acc = fma(vec[a], __ldg(col + a), acc);
                        outL[v * rs + b] = acc;
                    } else {
                        const int e2 = e - nl, m = mR, rm = ft.r[m], rm1 = ft.r[m + 1], nout = ninR + 2;
                        const int aa = e2 / nout, v = e2 - aa * nout;
                        const double *G = (v < ninR) ? CORE_BLK(m, sFix[m]) : CORE_BLK(m, sNf[2 * (m - 1) + (v - ninR)]);
                        const double *vec = inR + ((v < ninR) ? v : 0) * rs;
                        double acc = 0.0;
                        for (int b = 0; b < rm1; b++) acc = fma(__ldg(G + aa + (size_t)b * rm), vec[b], acc);
                        outR[v * rs + aa] = acc;
                    }
                }
                // a side that is already finished carries its set over unchanged
                if (!doL && k > 0) for (int e = tid; e < (1 + 2 * k) * rs; e += NT) outL[e] = inL[e];
                if (!doR && DX - 1 - k > 0) for (int e = tid; e < (1 + 2 * (DX - 1 - k)) * rs; e += NT) outR[e] = inR[e];
                if (!doL && k == 0 && tid == 0) outL[0] = 1.0;
                if (!doR && DX - 1 - k == 0 && tid == 0) outR[0] = 1.0;
                cur ^= 1;
                __syncthreads();
            }
            sLset = sVL + cur * NV * rs;
            sRset = sVR + cur * NV * rs;
        }
        PHASE_MARK(1);
        // 2b. per node tile: w_j = G_k[j] R (length r_k), u_j = L G_k[j] (length r_{k+1}), then the dots.
        //     left set index of dim i<k:  1+2i+side;  right set index of dim i>k: 1+2(d-1-i)+side
        {
            const double *Lk = sLset, *Rk1 = sRset;
            const double *Gk = ft.base + ft.off[k];
            const int blk = rk * rk1, per = rk + rk1;
            int T = (2 * NT) / per;
            T = T < 1 ? 1 : (T > TMAX ? TMAX : T);
            for (int j0 = 0; j0 < N; j0 += T) {
                const int nt = (N - j0 < T) ? N - j0 : T;
                if (j0 + T < N) {                   // next tile towards L1
                    const double *nx = Gk + (size_t)(j0 + T) * blk;
                    const int cnt = ((N - j0 - T < T) ? N - j0 - T : T) * blk;
                    for (int e = tid * 16; e < cnt; e += NT * 16) prefetch_l1(nx + e);
                }
                for (int e = tid; e < nt * per; e += NT) {
                    const int jl = e / per, q = e - jl * per;
                    const double *g = Gk + (size_t)(j0 + jl) * blk;
                    double acc = 0.0;
                    if (q < rk) {                   // w[q] = sum_b G[q + b*rk] R[b]
                        for (int b = 0; b < rk1; b++) acc = fma(__ldg(g + q + (size_t)b * rk), Rk1[b], acc);
                        sWt[jl * wst + q] = acc;
                    } else {                        // u[b] = sum_a L[a] G[a + b*rk]
                        const int b = q - rk;
                        const double *col = g + (size_t)b * rk;
                        for (int a = 0; a < rk; a++) acc = fma(Lk[a], __ldg(col + a), acc);
                        sWt[jl * wst + rs + b] = acc;
                    }
                }
                __syncthreads();
                for (int e = tid; e < nt * (2 * DX - 1); e += NT) {
                    const int q = e / nt, jl = e - q * nt, j = j0 + jl;    // q: 0 = self, then fixed-dim slots
                    const double *w = sWt + jl * wst, *u = w + rs;
                    double acc = 0.0;
                    int oslot;
                    if (q == 0) {
                        for (int a = 0; a < rk; a++) acc = fma(Lk[a], w[a], acc);
                        oslot = 2 * DX;
                        sV[j] = acc;
                    } else {
                        const int fs = q - 1, slot = fs >> 1, side = fs & 1;
                        const int i = slot < k ? slot : slot + 1;
                        if (i < k) {
                            const double *av = sLset + (1 + 2 * i + side) * rs;
                            for (int a = 0; a < rk; a++) acc = fma(av[a], w[a], acc);
                        } else {
                            const double *cv = sRset + (1 + 2 * (DX - 1 - i) + side) * rs;
                            for (int b = 0; b < rk1; b++) acc = fma(u[b], cv[b], acc);
                        }
                        oslot = 2 * i + side;
                    }
                    sC[j * CS + oslot] = acc;
                }
                __syncthreads();
            }
        }
        }   // direct plan

        for (int j = tid; j < N; j += NT) {                 // along the fiber (valuefunc.c:514-519)
            sC[j * CS + 2 * k] = sV[sNv[2 * j]];
            sC[j * CS + 2 * k + 1] = sV[sNv[2 * j + 1]];
            if (sAbs[j] == 0 && a.mode == MODE_VI) sAct[atomicAdd(&sMisc[0], 1)] = j;
        }
        __syncthreads();
#undef CORE_BLK

        PHASE_MARK(2);
        // optional diagnostics
        if (a.out.absorbed) for (int j = tid; j < N; j += NT) a.out.absorbed[obase + j] = sAbs[j];
        if (a.out.nbr_vary) for (int e = tid; e < 2 * N; e += NT) a.out.nbr_vary[2 * obase + e] = sNv[e];
        if (a.out.nbr_fixed) for (int e = tid; e < 2 * (DX - 1); e += NT) a.out.nbr_fixed[(size_t)f * 2 * (DX - 1) + e] = sNf[e];
        if (a.out.costs) for (int e = tid; e < N * CS; e += NT) a.out.costs[obase * CS + e] = sC[e];

        if (a.mode == MODE_COSTS) continue;      // mca_get_neighbor_costs only (nodeutil.c:647-713)

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

        // ---- 3. min over the control table -------------------------------------
        // work item = (candidate chunk "part", active node); a thread walks the candidates of
        // its chunk sequentially, so the first strict minimum in table order is kept
        // (bellman.c:539-543 + brute-force c3opt_minimize).  Items are ordered part-major: a warp
        // sees one chunk, so candidate-table reads are warp-uniform.
        const int nact = sMisc[0];
        int parts = nact > 0 ? (2 * NT) / nact : 1;
        parts = parts < 1 ? 1 : (parts > PMAX ? PMAX : parts);
        if (parts > P.nu) parts = P.nu;
        const int chunk = (P.nu + parts - 1) / parts;
        constexpr bool TAB = M::SEP && !A::exact;
        constexpr int NUD = M::NUD, CT = 2 * NUD + 2;
        const int NS = sp.ns;
        if (TAB) {
            // per-node invariants: norm0 = sum over all dims of the control-independent part of
            // (pl+pr), S0 = the matching part of <raw p, V_nbr>, gx = stage_x, then the 2*NUD
            // neighbour values the candidates weight.
            __syncthreads();                    // scratch region changes owner (sW/sU -> sNode)
            for (int idx = tid; idx < nact; idx += NT) {
                const int j = sAct[idx];
                double x[DX], u0[DU], b0[DX], s0[DX];
#pragma unroll
                for (int i = 0; i < DX; i++) x[i] = P.xgrid[P.xoff[i] + (i == k ? j : fix[i])];
#pragma unroll
                for (int i = 0; i < DU; i++) u0[i] = __ldg(P.utab + i);
                M::template drift<A>(x, u0, P.mp, b0);
                M::template sigma<A>(x, u0, P.mp, s0);
                double norm0 = 0.0, S0 = 0.0;
                const double *c = sC + j * CS;
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
                    S0 = fma(rl, c[2 * i], S0);
                    S0 = fma(rr, c[2 * i + 1], S0);
                }
                double *nd = sNode + idx * NS;
                nd[0] = norm0; nd[1] = S0; nd[2] = M::stage_x(x, P.mp);
#pragma unroll
                for (int m = 0; m < NUD; m++) { nd[3 + 2 * m] = c[2 * M::ud(m)]; nd[4 + 2 * m] = c[2 * M::ud(m) + 1]; }
                if (norm0 + P.amin < 1e-14) atomicOr(P.err, 1);
            }
        }
        __syncthreads();
        PHASE_MARK(3);
        for (int it = tid; it < nact * parts; it += NT) {
            const int part = it / nact, idx = it - part * nact;
            const int c0 = part * chunk, c1 = (c0 + chunk < P.nu) ? c0 + chunk : P.nu;
            double best = CUDART_INF;
            int ibest = 0x7fffffff;
            if (TAB) {
                const double *nd = sNode + idx * NS;
                const double norm0 = nd[0], S0 = nd[1], gx = nd[2];
                double cu[2 * NUD];
#pragma unroll
                for (int m = 0; m < 2 * NUD; m++) cu[m] = nd[3 + m];
                const bool disc = P.beta != 0.0;
                for (int cand = c0; cand < c1; cand++) {
                    const double2 *row = reinterpret_cast<const double2 *>(P.ctab + (size_t)cand * CT);
                    double S = S0;
#pragma unroll
                    for (int m = 0; m < NUD; m++) {
                        const double2 w = __ldg(row + m);
                        S = fma(w.x, cu[2 * m], S);
                        S = fma(w.y, cu[2 * m + 1], S);
                    }
                    const double2 ag = __ldg(row + NUD);
                    const double rinv = rcp_pos(norm0 + ag.x);
                    const double dt = P.h2 * rinv;
                    const double ebt = disc ? exp_nonpos(-P.beta * dt) : 1.0;
                    const double v = fma(dt, gx + ag.y, ebt * (S * rinv));
                    if (v < best) { best = v; ibest = cand; }
                }
            } else {
                const int j = sAct[idx];
                double x[DX], c[CS];
#pragma unroll
                for (int i = 0; i < DX; i++) x[i] = P.xgrid[P.xoff[i] + (i == k ? j : fix[i])];
#pragma unroll
                for (int m = 0; m < CS; m++) c[m] = sC[j * CS + m];
                NodeInv<M> inv;
                node_prepare<M, A>(P, P.utab, x, c, inv);
                int bad = 0;
                for (int cand = c0; cand < c1; cand++) {
                    double u[DU];
#pragma unroll
                    for (int i = 0; i < DU; i++) u[i] = __ldg(P.utab + (size_t)cand * DU + i);
                    const double v = candidate_value<M, A>(P, x, u, c, inv, bad);
                    if (v < best) { best = v; ibest = cand; }
                }
                if (bad) atomicOr(P.err, 1);
            }
            sBestV[part * P.nmax + idx] = best;
            sBestI[part * P.nmax + idx] = ibest;
        }
        __syncthreads();
        PHASE_MARK(4);
        for (int idx = tid; idx < nact; idx += NT) {
            const int j = sAct[idx];
            double best = sBestV[idx];
            int ibest = sBestI[idx];
            for (int pp = 1; pp < parts; pp++) {          // chunks in table order, strict '<'
                const double v = sBestV[pp * P.nmax + idx];
                if (v < best) { best = v; ibest = sBestI[pp * P.nmax + idx]; }
            }
            if (a.write_value) a.out.value[obase + j] = best;
            if (a.out.argmin) a.out.argmin[obase + j] = ibest;
            if (a.out.rows) {                               // policy row at u* (bellman.c:1851-1860)
                double x[DX], u[DU], b[DX], s[DX], prob[CS], dt;
#pragma unroll
                for (int i = 0; i < DX; i++) x[i] = P.xgrid[P.xoff[i] + (i == k ? j : fix[i])];
#pragma unroll
                for (int i = 0; i < DU; i++) u[i] = P.utab[(size_t)(ibest < P.nu ? ibest : 0) * DU + i];
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
        PHASE_MARK(5);
    }
    if (a.prof && tid == 0)
        for (int q = 0; q < 8; q++) atomicAdd(a.prof + q, pacc[q]);
}

// Candidate table of a separable model (FAST only): row c = [Wl_0, Wr_0, .., Wl_{NUD-1}, Wr_{NUD-1}, A, gu]
//   Wl_m / Wr_m = t_i*|b_i(u_c)| on the side the upwind scheme adds it to (0 inside the 1e-14 dead band),
//   A = sum of the W's (the candidate's share of the normaliser), gu = stage_u(u_c).
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
template <class M>
constexpr int ctab_stride() { return 2 * M::NUD + 2; }

// ---------------------------------------------------------------------------
template <class M, class A>
int launch_backup_t(const LaunchArgs &a_in, cudaStream_t st)
{
    static int sms = 0, max_optin = 0, max_sm = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaDeviceGetAttribute(&max_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    }
    // staged plan (TMA ring) when two CTAs of it fit on an SM, else the direct-load plan
    LaunchArgs a = a_in;
    a.staged = 1;
    size_t smem = SmemPlan(M::DX, M::NUD, a.P.nmax, a.ft.rmax, true).bytes();
    static const bool want_staged = getenv("C3SC_STAGED") != nullptr;   // experiment: measured slower than direct loads
    if (!want_staged || a.force_direct || 2 * (smem + 1024) > (size_t)max_sm) {
        a.staged = 0;
        smem = SmemPlan(M::DX, M::NUD, a.P.nmax, a.ft.rmax, false).bytes();
    }
    if (smem > (size_t)max_optin) return (int)cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k_backup<M, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_backup<M, A>, NT, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    if (const char *env = getenv("C3SC_CTAS_PER_SM")) { int v = atoi(env); if (v >= 1 && v < per_sm) per_sm = v; }   // experiments only
    int grid = sms * per_sm;
    if (grid > a.F) grid = a.F;
    if (grid < 1) return 0;
    k_backup<M, A><<<grid, NT, smem, st>>>(a);
    return (int)cudaGetLastError();
}

template <class M>
int launch_backup_m(int arith, const LaunchArgs &a, cudaStream_t st)
{
    if (arith == C3SC_ARITH_EXACT) return launch_backup_t<M, Exact>(a, st);
    return launch_backup_t<M, Fast>(a, st);
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
