// ft_mma_kernel.cuh -- stage 1 for ranks <= 32, as two kernels:
//
//   k_ft_chains   one WARP per (fiber, side): the prefix / suffix vector sets of
//                 valuef_eval_fiber_ind_nn (src/valuefunc.c:414-446 and the neighbour variants
//                 of :522-582), stepped block by block with coalesced loads straight from L2
//                 (right side: G, left side: the transposed copy).  No CTA barrier anywhere;
//                 results go to a per-fiber global scratch  sets[f][SETW].
//   k_ft_nodes    one CTA per group of <= 8 fibers that share dim_vary k.  Fibers of a group share
//                 all but the fixed coordinates, so the per-node work is a dense contraction and
//                 runs on the FP64 tensor cores (DMMA, mma.sync m8n8k4):
//                     W[a][g] = sum_b G_k[j][a,b] R_g[b]      (r_k x r_{k+1}) x (r_{k+1} x 8 fibers)
//                     U[g][b] = sum_a L_g[a] G_k[j][a,b]      (8 fibers x r_k) x (r_k x r_{k+1})
//                     left  variants:  C[v][j] = sum_a A_g[v][a] W_g[a][j]     per fiber, 8 nodes a tile
//                     right variants:  C[j][v] = sum_b U_g[j][b] C_g[b][v]
//                 G_k tiles are staged in shared memory with an odd column stride (both fragment
//                 orientations conflict-light) and prefetched one tile ahead through registers.
//
// Same outputs as k_ft_costs (ft_kernel.cuh), which stays the general path for larger ranks.
#pragma once
#include "ft_kernel.cuh"
#include "chain_kernel.cuh"
#include "ctl_types.cuh"

namespace c3sc {

constexpr int FTC_NT = 256;      // chain kernel: 8 warps = 8 tasks in flight per CTA
constexpr int FTN_NT = 256;      // node kernel: 8 warps
#ifndef C3SC_FTN_PAIR
#define C3SC_FTN_PAIR 0          // 1: phase 1 on node pairs (no half-empty rank tile when ceil(r/4) is odd): 17 % fewer phase-1 DMMAs at
                                 // r = 20 but 23 % more instructions (stacked-row addressing); measured 2 % SLOWER on B200
                                 // (0.831 vs 0.815 ms of node kernels per 65 536 fibers, profiles/r02b_node_kernel.md) -- opt-in
#endif
#ifndef C3SC_FTN_SELFSIDE
#define C3SC_FTN_SELFSIDE 1      // the self value from whichever side saves a variant tile
#endif
constexpr int FTN_T = 8;         // nodes per tile (one per warp in the w/u phase)
constexpr int FTN_TP = 8;        // row stride of a (rank index) row of w / u: 8 nodes, XOR-swizzled (ftn_swz): element
                                 // (rank row a, node j) of a fiber's tile sits at a*8 + (j ^ swz(a)).  With the odd
                                 // per-fiber stride sw below, the D-fragment stores of w take 2 shared-memory wavefronts
                                 // (the minimum for 64-bit accesses), those of u 4, a fragment load 2 -- against 8 / 8 / 4
                                 // unswizzled and 4 / 4 / 4 for a plain stride of 9 (layouts enumerated off line)
__host__ __device__ constexpr int ftn_swz(int a) { return (5 * ((a >> 1) & 1)) ^ ((a >> 2) & 1); }

// vectors-per-row stride of the chain sets in shared memory: >= 2d and = 4 (mod 8), so the A-fragment
// reads (row v, col q) of the stepped products are bank-conflict free
__host__ __device__ inline int ft_chain_nvp(int d)
{
    int n = 2 * d;
    while ((n & 7) != 4) n++;
    return n;
}

// Both kernels are instantiated on KS = ceil(rmax/4), the number of MMA k-steps a rank index needs;
// MT = ceil(KS/2) 8-wide tiles.  Every core is padded to that geometry (ft_padded_layout), so all
// fragment loops have compile-time trip counts and no predicates (zero padding does the masking).
template <int KS>
struct FtChainPlan {            // shared memory of k_ft_chains, per warp: two set buffers + indices
    int nvp, bufDoubles, perWarpDoubles, perWarpInts;
    __host__ __device__ FtChainPlan(const DevFT &ft)
    {
        nvp = ft_chain_nvp(ft.d);
        bufDoubles = 8 * ((KS + 1) / 2) * nvp;       // [rank row q][vector v], rows padded to the MMA tile
        perWarpDoubles = 2 * bufDoubles;
        perWarpInts = 4 * ft.d;                      // fixed indices + neighbour pairs
    }
    __host__ __device__ size_t bytes() const { return (size_t)(FTC_NT / 32) * (perWarpDoubles * 8 + perWarpInts * 4); }
};

// ---------------------------------------------------------------------------
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// One warp per (fiber, side).  A step multiplies the whole vector set by one core block on the
// tensor cores:  out[v][o] = sum_q in[v][q] B[q][o]  (left: B = G_m[f_m], right: B = G_m[f_m]^T),
// B fragments straight from the zero-padded core copy in L2 (all loads of a step are independent),
// and appends the two neighbour variants  in[0] . G_m[nb]  with plain FMAs.
#ifndef C3SC_FTC_MINB
#define C3SC_FTC_MINB 2
#endif
template <int KS>
__global__ void __launch_bounds__(FTC_NT, KS <= 6 ? C3SC_FTC_MINB : 1) k_ft_chains(const FtArgs a, double *sets)
{
    constexpr int NTL = (KS + 1) / 2, VT = (2 * MAXD + 7) / 8;
    const DevProblem &P = a.P;
    const DevFT &ft = a.ft;
    const int d = ft.d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    constexpr int NW = FTC_NT / 32;
    extern __shared__ __align__(16) double smem[];
    const FtChainPlan<KS> cp(ft);
    const int NVP = cp.nvp;
    double *buf0 = smem + warp * cp.perWarpDoubles, *buf1 = buf0 + cp.bufDoubles;
    int *iw = reinterpret_cast<int *>(smem + NW * cp.perWarpDoubles) + warp * cp.perWarpInts;
    int *sFix = iw, *sNf = iw + d;                   // sNf[2*i], sNf[2*i+1]: pair of dimension i
    const int SETW = a.setw, RS = a.rs;

    // programmatic dependent launch: the node kernel behind this one may start its prologue (tile buffers, flags, first TMA
    // fetches) now; it waits with griddepcontrol.wait before it reads the records written here
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // every value a fragment may touch must be finite (padding rows meet zero rows of the block)
    for (int e = lane; e < cp.perWarpDoubles; e += 32) buf0[e] = 0.0;

    // Tasks differ in length (a left chain has k steps, a right chain d-1-k): warps take them from a counter,
    // longest first -- even tasks are right chains over the k-grouped order ascending (k = 0 first), odd tasks
    // left chains over it descending (k = d-1 first).
    for (;;) {
        int task = 0;
        if (lane == 0) task = atomicAdd(a.task_count, 1);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= 2 * a.F) break;
        const bool left = (task & 1) != 0;
        const int f = a.perm[left ? a.F - 1 - (task >> 1) : (task >> 1)];
        int k = a.dim_vary[f];
        k = k < 0 ? 0 : (k >= d ? d - 1 : k);
        const int nsteps = left ? k : d - 1 - k;
        __syncwarp();
        if (lane < d) {
            const int i = lane, i0 = ft_clamp_index(a.fixed_ind[(size_t)f * d + i], P.ngrid[i]);
            sFix[i] = i0;
            if (i != k) {
                int lo, hi;
                ft_fixed_pair(P, i, i0, lo, hi);
                if (a.nbr_fixed_in) {
                    const int slot = i < k ? i : i - 1;
                    lo = a.nbr_fixed_in[(size_t)f * 2 * (d - 1) + 2 * slot];
                    hi = a.nbr_fixed_in[(size_t)f * 2 * (d - 1) + 2 * slot + 1];
                }
                sNf[2 * i] = lo;
                sNf[2 * i + 1] = hi;
            }
        }
        if (lane == 0) buf0[0] = 1.0;
        __syncwarp();
        double *in = buf0, *out = buf1;
        for (int s = 0; s < nsteps; s++) {
            const int m = left ? s : d - 1 - s;
            const int rq = left ? ft.r[m] : ft.r[m + 1], ro = left ? ft.r[m + 1] : ft.r[m];
            const int ld = ft.ldp[m], pblk = ft.ldp[m] * ft.cpp[m];
            const double *base = ft.baseP + ft.offP[m];
            const double *bc = base + (size_t)sFix[m] * pblk;
            const double *bl = base + (size_t)sNf[2 * m] * pblk, *bh = base + (size_t)sNf[2 * m + 1] * pblk;
            const int nin = 1 + 2 * s;
            constexpr int nks = KS, ntl = NTL;
            const int mtn = (nin + 7) >> 3;
            (void)rq; (void)ro;
            // B fragments (row q = 4ks+tig, col o = 8nt+gid); element (a,b) of a padded block at b*ld + a.
            // All loads of the centre and the first neighbour block are issued before any use.
            const int sq = left ? 1 : ld, so = left ? ld : 1;        // strides of q and of o inside the block
            const int foff = tig * sq + gid * so;
            double Bc[KS][NTL], Bl[KS][NTL], Bh[KS][NTL];
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
#pragma unroll
                for (int nt = 0; nt < NTL; nt++)
                    Bc[ks][nt] = (ks < nks && nt < ntl) ? __ldg(bc + foff + 4 * ks * sq + 8 * nt * so) : 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
#pragma unroll
                for (int nt = 0; nt < NTL; nt++)
                    Bl[ks][nt] = (ks < nks && nt < ntl) ? __ldg(bl + foff + 4 * ks * sq + 8 * nt * so) : 0.0;
            // the existing vectors against the centre block
            double Af0[KS];                                          // A fragments of the first vector tile (row v = gid)
#pragma unroll
            for (int ks = 0; ks < KS; ks++) Af0[ks] = (ks < nks) ? in[(4 * ks + tig) * NVP + gid] : 0.0;
#pragma unroll
            for (int mt = 0; mt < VT; mt++) {
                if (mt < mtn) {
                    double acc[NTL][2];
#pragma unroll
                    for (int nt = 0; nt < NTL; nt++) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll
                    for (int ks = 0; ks < KS; ks++) {
                        if (ks < nks) {
                            const double af = mt == 0 ? Af0[ks] : in[(4 * ks + tig) * NVP + 8 * mt + gid];   // A: row v, col q
#pragma unroll
                            for (int nt = 0; nt < NTL; nt++)
                                if (nt < ntl) dmma_m8n8k4(acc[nt][0], acc[nt][1], af, Bc[ks][nt]);
                        }
                    }
                    const int v = 8 * mt + gid;                      // D: row v, cols o = 8nt+2tig, +1
                    if (v < nin) {
#pragma unroll
                        for (int nt = 0; nt < NTL; nt++) {
                            if (nt < ntl) {
                                out[(8 * nt + 2 * tig) * NVP + v] = acc[nt][0];
                                out[(8 * nt + 2 * tig + 1) * NVP + v] = acc[nt][1];
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
#pragma unroll
                for (int nt = 0; nt < NTL; nt++)
                    Bh[ks][nt] = (ks < nks && nt < ntl) ? __ldg(bh + foff + 4 * ks * sq + 8 * nt * so) : 0.0;
            // the two new variants: row 0 of the first vector tile (the prefix / suffix itself) against the
            // neighbour blocks; only D row 0 (lanes with gid == 0) is kept
#pragma unroll
            for (int h = 0; h < 2; h++) {
                double acc[NTL][2];
#pragma unroll
                for (int nt = 0; nt < NTL; nt++) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < KS; ks++) {
                    if (ks < nks) {
#pragma unroll
                        for (int nt = 0; nt < NTL; nt++)
                            if (nt < ntl) dmma_m8n8k4(acc[nt][0], acc[nt][1], Af0[ks], h ? Bh[ks][nt] : Bl[ks][nt]);
                    }
                }
                if (gid == 0) {
#pragma unroll
                    for (int nt = 0; nt < NTL; nt++) {
                        if (nt < ntl) {
                            out[(8 * nt + 2 * tig) * NVP + nin + h] = acc[nt][0];
                            out[(8 * nt + 2 * tig + 1) * NVP + nin + h] = acc[nt][1];
                        }
                    }
                }
            }
            __syncwarp();
            double *t = in; in = out; out = t;
        }
        // final set -> global record of the fiber (chain_kernel.cuh: row v = vector v, RS doubles, zero beyond the
        // rank; the right set follows the 1 + 2k rows of the left one)
        {
            const int rl = left ? ft.r[k] : ft.r[k + 1];
            const int nv = left ? 1 + 2 * k : 1 + 2 * (d - 1 - k);
            double *dst = sets + (size_t)f * SETW + (size_t)(left ? 0 : 1 + 2 * k) * RS;
            for (int e = lane; e < nv * RS; e += 32) {
                const int v = e / RS, q = e - v * RS;
                dst[e] = q < rl ? in[q * NVP + v] : 0.0;
            }
        }
    }
}


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

// w / u of one node for the (up to) 8 fibers of the group (first phase of k_ft_nodes):
//   W[a][g] = sum_b G[a,b] R_g[b],  U[g][b] = sum_a L_g[a] G[a,b];  k-step outermost, so the MT accumulator
//   tiles of either product are independent DMMA chains.
template <int KS, bool DO_W, bool DO_U>
__device__ __forceinline__ void node_wu(const double *gj, int offW, int offU, int ldk, const double (&Rf)[KS], const double (&Lf)[KS],
                                        double *sW, double *sU, int SW, int tig, int gid, int warp)
{
    constexpr int MT = (KS + 1) / 2;
    double dw[MT][2], du[MT][2];
#pragma unroll
    for (int t = 0; t < MT; t++) { dw[t][0] = dw[t][1] = 0.0; du[t][0] = du[t][1] = 0.0; }
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
        if constexpr (DO_W) {
#pragma unroll
            for (int mt = 0; mt < MT; mt++) dmma_m8n8k4(dw[mt][0], dw[mt][1], gj[offW + ks * 4 * ldk + mt * 8], Rf[ks]);
        }
        if constexpr (DO_U) {
#pragma unroll
            // u TRANSPOSED, u^T = G^T L^T: the same G elements (the B fragment of G is the A fragment of G^T) and the same L
            // registers with the operands swapped; the D fragment then is (rank row b, fiber columns) like w's, whose stores
            // take 2 shared-memory wavefronts instead of 4
            for (int nb = 0; nb < MT; nb++) dmma_m8n8k4(du[nb][0], du[nb][1], gj[offU + nb * 8 * ldk + ks * 4], Lf[ks]);
        }
    }
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
        if constexpr (DO_W) {                                    // D: row a = 8mt+gid, cols fiber 2*tig, 2*tig+1
            const int jw = warp ^ ftn_swz(gid);                  // swz(8mt + gid) = swz(gid)
            sW[(2 * tig) * SW + (8 * mt + gid) * FTN_TP + jw] = dw[mt][0];
            sW[(2 * tig + 1) * SW + (8 * mt + gid) * FTN_TP + jw] = dw[mt][1];
        }
        if constexpr (DO_U) {                                    // D of u^T: row b = 8mt+gid, cols fiber 2*tig, 2*tig+1
            const int ju = warp ^ ftn_swz(gid);
            sU[(2 * tig) * SW + (8 * mt + gid) * FTN_TP + ju] = du[mt][0];
            sU[(2 * tig + 1) * SW + (8 * mt + gid) * FTN_TP + ju] = du[mt][1];
        }
    }
}

// w (or u) of a PAIR of nodes: the two blocks stacked along the free index of the product.  A rank r = 4 (mod 8)
// pads every node's 8-wide tiles by a half (20 -> 24 rows); 2r rows need ceil(2r/8) <= KS tiles against 2*ceil(r/8),
// one tile less whenever KS is odd (r = 17..20: 5 instead of 6).  Row R of the stack is row R of the first block for
// R < r and row R - r of the second; rows beyond 2r read finite data and are not stored.
// u is computed TRANSPOSED, u^T = G^T L^T: the same instruction stream as w = G R with other strides (row stride ldk and
// k stride 1 instead of 1 and ldk), the other register operand and the other tile buffer -- one code path for both
// roles (warps 0..3: w of pairs 0..3, warps 4..7: u), and u's D fragments land in the (rank row, fiber column) form
// whose stores are conflict-free like w's.
template <int KS>
__device__ __forceinline__ void node_pair_wu(const double *g2, int pblk, int rr, int rowstride, int kstride, int tigoff, int jn,
                                             const double (&Xf)[KS], double *dst, int SW, int tig, int gid)
{
    double acc[KS][2];
    int ld[KS];
#pragma unroll
    for (int mt = 0; mt < KS; mt++) {
        acc[mt][0] = acc[mt][1] = 0.0;
        const int R = 8 * mt + gid;
        ld[mt] = tigoff + (R < rr ? R * rowstride : (R < 2 * rr ? pblk + (R - rr) * rowstride : 0));
    }
#pragma unroll
    for (int ks = 0; ks < KS; ks++)
#pragma unroll
        for (int mt = 0; mt < KS; mt++) dmma_m8n8k4(acc[mt][0], acc[mt][1], g2[ld[mt] + ks * kstride], Xf[ks]);
#pragma unroll
    for (int mt = 0; mt < KS; mt++) {                          // D: stacked row 8mt+gid, cols fiber 2*tig, 2*tig+1
        const int R = 8 * mt + gid;
        if (R < 2 * rr) {
            const int nd = R >= rr, a = R - nd * rr;
            const int o = a * FTN_TP + ((jn + nd) ^ ftn_swz(a));
            dst[(2 * tig) * SW + o] = acc[mt][0];
            dst[(2 * tig + 1) * SW + o] = acc[mt][1];
        }
    }
}

// Shared memory of k_ft_nodes.  Everything the MMA fragments read is zero-padded to the fragment
// shape, so no fragment load is predicated.  The chain records are NOT staged: every fragment of a fiber's
// variant sets is loaded from L2 once per CTA and lives in registers for all node tiles.
template <int KS>
struct FtNodePlan {
    int nmax, sw, gbuf;
    int oG, oW, oU, oBar, nDoubles;
    int oFix, oNf, oFid, oWall, nInts;
    __host__ __device__ FtNodePlan(const DevFT &ft, int nmax_)
    {
        nmax = nmax_;
        sw = 8 * ((KS + 1) / 2) * FTN_TP + 1;        // one fiber's w (or u) tile: [rank index][8 nodes, swizzled], odd stride
        int gt = 0;
        for (int k = 0; k < ft.d; k++) {
            // 8 blocks + zero slack for the fragment rows / columns that run past the last block: the padded
            // MMA geometry reads at most 7 columns and 4 rows beyond a block (ranks 17..24 on 24-wide tiles)
            const int g = FTN_T * ft.ldq[k] * (int)ft.r[k + 1] + 8 * ft.ldq[k] + 8;
            gt = g > gt ? g : gt;
        }
        gt = (gt + 1) & ~1;
        gbuf = gt;
        int o = 0;
        oG = o;    o += 2 * gt;                      // two tile buffers, each 16-byte aligned
        oW = o;    o += 2 * FT_FBMAX * sw;           // w tiles of two consecutive node tiles (one barrier per tile, see the loop)
        oU = o;    o += 2 * FT_FBMAX * sw;
        o = ft_even_up(o);
        oBar = o;  o += 2;                           // two mbarriers
        nDoubles = o;
        int q = 0;
        oFix = q;  q += FT_FBMAX * ft.d;
        oNf = q;   q += FT_FBMAX * 2 * ft.d;
        oFid = q;  q += FT_FBMAX;
        oWall = q; q += FT_FBMAX;
        nInts = q;                                   // then 8*nmax bytes (sAbs)
    }
    __host__ __device__ size_t bytes() const { return (size_t)nDoubles * 8 + (size_t)nInts * 4 + (size_t)FT_FBMAX * nmax; }
};

// Everything the tile loop of k_ft_nodes needs besides its compile-time tile counts.
struct FtNodeCtx {
    const double *sG; double *sW, *sU;
    unsigned long long *mbar;
    const double *Gp;
    const int *sFid;
    int GB, SW, pblk, ldk, jb, je, nf, nvL, nvR, d, CS, RS, setw;
    long long NS;
    double *cst, *costs;
    const double *sets;
    int ldo;
    int vL0, vR0;            // first vector of the left / right set that takes part in the dots: exactly one side keeps its
                             // vector 0 (the prefix / suffix alone, which gives the SELF value), whichever needs fewer 8-wide tiles
    int rk, rk1;             // ranks of the varying core
    int regionStride;        // > 0: costs go to the CTA's region, fiber g at g*regionStride (fused stage 2); 0: fiber id * ldo
};

// The node-tile loop with compile-time numbers of 8-wide variant tiles: ML over the left set (1 + 2k vectors, 0 when
// the left side has no variants and the self value comes from the right), NR over the right set.  A DMMA that is
// merely predicated off still costs its issue slot and fragment load, hence the dispatch on (ML, NR).
//   phase 1 (warp = node of the tile)   W[a][g] = sum_b G[a,b] R_g[b],  U[g][b] = sum_a L_g[a] G[a,b]   for the 8 fibers
//   phase 2 (warp = fiber of the group) left  C[v][jl] = sum_a A[v][a] W[a][jl],  right C[jl][v] = sum_b U[jl][b] Cv[b][v]
// The A / Cv fragments (the fiber's variant sets) and R / L are loaded from the chain records ONCE and stay in
// registers; per tile phase 2 reads only the w / u fragments from shared memory.
// The slot-major cost scratch is written once here and read once by stage 2 (which reads it with evict-first loads,
// control_kernel.cuh).  Streaming STORES here as well (-DC3SC_CST_STREAM_ST=1, st.global.cs) were measured and are off: the node
// kernel gets 1 % slower (875 against 869 us per step, 975 against 953 MB of DRAM writes) for no gain elsewhere.
#ifndef C3SC_CST_STREAM_ST
#define C3SC_CST_STREAM_ST 0
#endif
#if C3SC_CST_STREAM_ST
#define C3SC_CST_ST(p, v) __stcs((p), (v))
#define C3SC_CST_ST2(p, v) __stcs((p), (v))
#else
#define C3SC_CST_ST(p, v) (*(p) = (v))
#define C3SC_CST_ST2(p, v) (*(p) = (v))
#endif
template <int KS, int ML, int NR>
__device__ __forceinline__ void ftn_tile_loop(const FtNodeCtx &c)
{
    constexpr int VT = (2 * MAXD + 7) / 8;
    constexpr bool DO_W = ML > 0, DO_U = NR > 0;
    constexpr bool PAIR = DO_W && DO_U && (KS & 1) && C3SC_FTN_PAIR;     // node pairs in phase 1 (node_pair_w / node_pair_u)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int d = c.d, RS = c.RS;

    // The chain records are the first thing of the kernel that the preceding launch (chain stage) writes: the kernel is launched
    // with programmatic stream serialization, everything above (tile buffers, flags, the first tile fetches) overlaps its tail.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // ---- register-resident operands ---------------------------------------------------------------
    // phase 1: B fragment of R (row b = 4ks+tig, col fiber gid), A fragment of L (row fiber gid, col a = 4ks+tig)
    // (node pairs: a warp computes w OR u, so it keeps only R or only L, in Rf)
    double Rf[KS], Lf[PAIR ? 1 : KS];
    {
        const bool on = gid < c.nf;
        const double *rec = c.sets + (size_t)(on ? c.sFid[gid] : 0) * c.setw;
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
            if constexpr (PAIR) Rf[ks] = on ? __ldg(rec + (warp < 4 ? (size_t)c.nvL * RS : (size_t)0) + 4 * ks + tig) : 0.0;
            else {
                Rf[ks] = on ? __ldg(rec + (size_t)c.nvL * RS + 4 * ks + tig) : 0.0;
                Lf[ks] = on ? __ldg(rec + 4 * ks + tig) : 0.0;
            }
        }
    }
    // phase 2, fiber g = warp: A fragments of the left set (row v = 8mt+gid, col a = 4ks+tig), B fragments of the
    // right set (row b = 4ks+tig, col v = 8nb+gid); both are rows v of the record, element 4ks+tig
    double aL[ML > 0 ? ML : 1][KS], bR[NR > 0 ? NR : 1][KS];
    {
        const bool on = warp < c.nf;
        const double *rec = c.sets + (size_t)(on ? c.sFid[warp] : 0) * c.setw;
#pragma unroll
        for (int mt = 0; mt < ML; mt++) {
            const int v = c.vL0 + 8 * mt + gid;
#pragma unroll
            for (int ks = 0; ks < KS; ks++) aL[mt][ks] = (on && v < c.nvL) ? __ldg(rec + (size_t)v * RS + 4 * ks + tig) : 0.0;
        }
#pragma unroll
        for (int nb = 0; nb < NR; nb++) {
            const int v = c.vR0 + 8 * nb + gid;
#pragma unroll
            for (int ks = 0; ks < KS; ks++) bR[nb][ks] = (on && v < c.nvR) ? __ldg(rec + (size_t)(c.nvL + v) * RS + 4 * ks + tig) : 0.0;
        }
    }
    // output slots of this lane's accumulator rows / columns (-1 = no such variant)
    int slotL[VT], slotR[VT][2];
#pragma unroll
    for (int t = 0; t < VT; t++) {
        const int v = c.vL0 + 8 * t + gid;
        slotL[t] = v >= c.nvL ? -1 : (v == 0 ? 2 * d : 2 * ((v - 1) >> 1) + ((v - 1) & 1));
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int vv = c.vR0 + 8 * t + 2 * tig + h;
            slotR[t][h] = vv >= c.nvR ? -1 : (vv == 0 ? 2 * d                               // vector 0 is R itself: u . R = the self value
                                                       : 2 * (d - 1 - ((vv - 1) >> 1)) + ((vv - 1) & 1));
        }
    }
    // element offsets of this lane's outputs inside the slot-major scratch, relative to cst + fiber row + tile
    // start (left: nodes 2tig, 2tig+1 of variant row gid; right: node gid of variant columns 2tig, 2tig+1)
    int outL[VT], outR[VT][2];
    const bool off32 = c.cst && (long long)c.CS * c.NS < 0x7fffffffLL && !c.costs && ((c.NS | (long long)c.ldo) & 1) == 0 &&
                       ((size_t)c.cst & 15) == 0;
#pragma unroll
    for (int t = 0; t < VT; t++) {
        outL[t] = (off32 && slotL[t] >= 0) ? (int)(slotL[t] * c.NS) + 2 * tig : -1;
#pragma unroll
        for (int h = 0; h < 2; h++) outR[t][h] = (off32 && slotR[t][h] >= 0) ? (int)(slotR[t][h] * c.NS) + gid : -1;
    }
    const int ldk = c.ldk, SW = c.SW, pblk = c.pblk, GB = c.GB;
    const int offW = tig * ldk + gid, offU = gid * ldk + tig;           // fragment origins inside a node block
    // node pairs: w = G R walks rows with stride 1 and the contraction index with stride ldk, u^T = G^T L^T the other way round
    const bool urole = warp >= 4;
    const int prr = urole ? c.rk1 : c.rk, prow = urole ? ldk : 1, pks = urole ? 4 : 4 * ldk, ptig = urole ? tig : tig * ldk;
    const double *wg0 = c.sW + warp * SW + tig * FTN_TP, *ug0 = c.sU + warp * SW + tig * FTN_TP;
    const int WB = FT_FBMAX * SW;                           // one buffer of w (or u) tiles
    const int jl0 = gid ^ ftn_swz(tig), jl1 = gid ^ ftn_swz(4 + tig);
    const size_t idf = c.regionStride > 0 ? (size_t)warp * c.regionStride : (warp < c.nf ? (size_t)c.sFid[warp] * c.ldo : 0);
    const int CS = c.CS;

    auto fetch = [&](int j1, int buf) {                     // tid 0: tile starting at node j1 -> buffer buf
        const int nt1 = (c.je - j1 < FTN_T) ? c.je - j1 : FTN_T;
        const unsigned bytes = (unsigned)(((size_t)nt1 * pblk * 8 + 15) & ~(size_t)15);
        mbar_expect_tx(c.mbar + buf, bytes);
        bulk_g2s(const_cast<double *>(c.sG) + buf * GB, c.Gp + (size_t)j1 * pblk, bytes, c.mbar + buf);
    };

    unsigned ph0 = 0, ph1 = 0;                             // mbarrier phase parity of the two buffers
    int buf = 0;
    for (int j0 = c.jb; j0 < c.je; j0 += FTN_T, buf ^= 1) {
        const int nt = (c.je - j0 < FTN_T) ? c.je - j0 : FTN_T;
        mbar_wait(c.mbar + buf, buf ? ph1 : ph0);
        if (buf) ph1 ^= 1; else ph0 ^= 1;
        // ---- w / u of node jl = warp, all fibers of the group (k-step outermost: independent DMMA chains) ----
        // The w / u tiles are double-buffered like the G tiles: phase 1 of tile i+1 writes the other buffer, so the
        // barrier after phase 1 of tile i+1 is also the one that orders phase 2 of tile i before phase 1 of tile i+2
        // re-uses its buffer: ONE CTA barrier per node tile.
        if constexpr (PAIR) {
            const int jn = 2 * (warp & 3);                  // first node of this warp's pair
            if (jn < nt)
                node_pair_wu<KS>(c.sG + buf * GB + jn * pblk, pblk, prr, prow, pks, ptig, jn, Rf, (warp < 4 ? c.sW : c.sU) + buf * WB, SW, tig, gid);
        } else {
            if (warp < nt) node_wu<KS, DO_W, DO_U>(c.sG + buf * GB + warp * pblk, offW, offU, ldk, Rf, Lf, c.sW + buf * WB, c.sU + buf * WB, SW, tig, gid, warp);
        }
        __syncthreads();
        const double *wg = wg0 + buf * WB, *ug = ug0 + buf * WB;
        if (tid == 0 && j0 + 2 * FTN_T < c.je) fetch(j0 + 2 * FTN_T, buf);     // this buffer is free: fetch the tile after next
        // ---- variant dots of fiber g = warp over the tile's nodes -----------------------------
        if (warp < c.nf) {
            double dl[ML > 0 ? ML : 1][2], dr[NR > 0 ? NR : 1][2];
#pragma unroll
            for (int t = 0; t < ML; t++) dl[t][0] = dl[t][1] = 0.0;
#pragma unroll
            for (int t = 0; t < NR; t++) dr[t][0] = dr[t][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
                if constexpr (ML > 0) {
                    const double bw = wg[ks * 4 * FTN_TP + ((ks & 1) ? jl1 : jl0)];   // B fragment (row a = 4ks+tig, col jl = gid)
#pragma unroll
                    for (int mt = 0; mt < ML; mt++) dmma_m8n8k4(dl[mt][0], dl[mt][1], aL[mt][ks], bw);
                }
                if constexpr (NR > 0) {
                    const double au = ug[ks * 4 * FTN_TP + ((ks & 1) ? jl1 : jl0)];   // A fragment (row jl = gid, col b = 4ks+tig)
#pragma unroll
                    for (int nb = 0; nb < NR; nb++) dmma_m8n8k4(dr[nb][0], dr[nb][1], au, bR[nb][ks]);
                }
            }
            const size_t idb = idf + j0;
            if (off32 && nt == FTN_T && (j0 & 1) == 0) {
                // full tile, slot-major scratch only (the pipeline's case): one 32-bit element offset per output
                double *cb = c.cst + idb;
#pragma unroll
                for (int mt = 0; mt < ML; mt++)
                    if (outL[mt] >= 0) C3SC_CST_ST2(reinterpret_cast<double2 *>(cb + outL[mt]), make_double2(dl[mt][0], dl[mt][1]));
#pragma unroll
                for (int nb = 0; nb < NR; nb++) {
#pragma unroll
                    for (int h = 0; h < 2; h++)
                        if (outR[nb][h] >= 0) C3SC_CST_ST(cb + outR[nb][h], dr[nb][h]);
                }
            } else {
#pragma unroll
                for (int mt = 0; mt < ML; mt++) {
                    const int slot = slotL[mt];
                    if (slot >= 0) {                                         // D: row v, cols jl = 2tig, 2tig+1
                        const int jl = 2 * tig;
                        const double d0 = dl[mt][0], d1 = dl[mt][1];
                        if (c.cst) {
                            double *o = c.cst + (size_t)slot * c.NS + idb + jl;
                            if (jl < nt) o[0] = d0;
                            if (jl + 1 < nt) o[1] = d1;
                        }
                        if (c.costs) {
                            if (jl < nt) c.costs[(idb + jl) * CS + slot] = d0;
                            if (jl + 1 < nt) c.costs[(idb + jl + 1) * CS + slot] = d1;
                        }
                    }
                }
                if (gid < nt) {
#pragma unroll
                    for (int nb = 0; nb < NR; nb++) {                        // D: row jl = gid, cols v = 8nb+2tig, +1
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int slot = slotR[nb][h];
                            if (slot < 0) continue;
                            const double val = dr[nb][h];
                            if (c.cst) c.cst[(size_t)slot * c.NS + idb + gid] = val;
                            if (c.costs) c.costs[(idb + gid) * CS + slot] = val;
                        }
                    }
                }
            }
        }
    }
}

template <int KS, bool FUSED>
__device__ __forceinline__ void ftn_body(const FtArgs &a, const double *sets, const CtlArgs *ctl)
{
    const DevProblem &P = a.P;
    const DevFT &ft = a.ft;
    const int d = ft.d;
    const int tid = threadIdx.x;

    int k, gstart, nf;
    ft_find_group(a, k, gstart, nf);
    if (k < 0) return;

    extern __shared__ __align__(16) double smem[];
    const FtNodePlan<KS> sp(ft, P.nmax);
    const int nmax = sp.nmax;
    double *sG = smem + sp.oG, *sW = smem + sp.oW, *sU = smem + sp.oU;
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem + sp.oBar);
    int *ismem = reinterpret_cast<int *>(smem + sp.nDoubles);
    int *sFix = ismem + sp.oFix, *sNf = ismem + sp.oNf;
    int *sFid = ismem + sp.oFid, *sWall = ismem + sp.oWall;
    signed char *sAbs = reinterpret_cast<signed char *>(ismem + sp.nInts);

    const int N = P.ngrid[k];
    const int rk1 = ft.r[k + 1];
    const int pblk = ft.ldq[k] * rk1;                                  // compact block (rows even-padded)
    const int nvL = 1 + 2 * k, nvR = 1 + 2 * (d - 1 - k);
    // Vector 0 of a side is the prefix / suffix alone: against the other side's w / u it gives the SELF value, so only one
    // side needs it in the dots.  The side whose tile count does not grow keeps it (d = 10: three 8-wide tiles for every k,
    // against 3.5 on average when the left side always kept it); a side with nothing left to do is dropped altogether.
    const int tA = ((nvL + 7) >> 3) + ((nvR - 1 + 7) >> 3), tB = ((nvL - 1 + 7) >> 3) + ((nvR + 7) >> 3);
    const bool selfLeft = C3SC_FTN_SELFSIDE ? (tA <= tB) : (nvL > 1 || nvR <= 1);
    const int vL0 = selfLeft ? 0 : 1, vR0 = selfLeft ? 1 : 0;
    const int mtL = (nvL - vL0 + 7) >> 3, ntR = (nvR - vR0 + 7) >> 3;      // 8-wide tiles over the variant vectors

    // this CTA's share of the fiber: a contiguous range of node tiles
    const int ntiles = (N + FTN_T - 1) / FTN_T, nsp = a.nsplit > 0 ? a.nsplit : 1;
    const int tper = (ntiles + nsp - 1) / nsp;
    const int jb = (int)blockIdx.y * tper * FTN_T, je = (jb + tper * FTN_T < N) ? jb + tper * FTN_T : N;
    if (jb >= N) return;

    // ---- tile buffers: zero once (fragment rows / columns past the ranks read finite data that meets
    //      zero operands), then the first two tiles are on their way while the flags are computed -----
    const double *Gp = ft.baseQ + ft.offQ[k];
    const int GB = sp.gbuf;
    for (int e = tid; e < 2 * GB; e += FTN_NT) sG[e] = 0.0;
    for (int e = tid; e < 4 * FT_FBMAX * sp.sw; e += FTN_NT) sW[e] = 0.0;       // sW and sU are adjacent
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        mbar_init(mbar, 1);
        mbar_init(mbar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int b = 0; b < 2; b++) {
            const int j1 = jb + b * FTN_T;
            if (j1 >= je) break;
            const int nt1 = (je - j1 < FTN_T) ? je - j1 : FTN_T;
            const unsigned bytes = (unsigned)(((size_t)nt1 * pblk * 8 + 15) & ~(size_t)15);
            mbar_expect_tx(mbar + b, bytes);
            bulk_g2s(sG + b * GB, Gp + (size_t)j1 * pblk, bytes, mbar + b);
        }
    }

    // fused stage 2: take a free region of the ring (more regions than CTAs can be resident, so one is always free)
    __shared__ int s_region;
    if constexpr (FUSED) {
        if (tid == 0) {
            int r;
            for (;;) {
                r = (int)((unsigned)atomicAdd(a.ring_flag + a.nring, 1) % (unsigned)a.nring);
                if (atomicCAS(a.ring_flag + r, 0, 1) == 0) break;
            }
            s_region = r;
        }
    }

    ft_flags_and_indices(a, k, nf, gstart, jb, je, sFid, sWall, sFix, sNf, sAbs, nmax);      // ends with a barrier

    const int njp = ft_even_up(nmax);
    double *region = FUSED ? a.ring + (size_t)s_region * a.region_doubles : nullptr;
    FtNodeCtx c;
    c.sG = sG; c.sW = sW; c.sU = sU; c.mbar = mbar; c.Gp = Gp; c.sFid = sFid;
    c.GB = GB; c.SW = sp.sw; c.pblk = pblk; c.ldk = ft.ldq[k]; c.jb = jb; c.je = je; c.nf = nf;
    c.nvL = nvL; c.nvR = nvR; c.d = d; c.CS = 2 * d + 1; c.RS = a.rs; c.setw = a.setw;
    c.NS = a.NS; c.cst = a.cst; c.costs = a.costs; c.sets = sets; c.ldo = a.ldo;
    c.regionStride = 0;
    if constexpr (FUSED) { c.cst = region; c.NS = (long long)FT_FBMAX * njp; c.costs = nullptr; c.regionStride = njp; }
    c.vL0 = vL0; c.vR0 = vR0; c.rk = ft.r[k]; c.rk1 = rk1;
    // the tile counts of the variant sets are CTA-uniform run-time values: dispatch to a loop with compile-time
    // counts, so that no predicated-off DMMA (and its fragment load) is issued at all
    switch (mtL * 8 + ntR) {
#define C3SC_ND(L, R) case (L) * 8 + (R): ftn_tile_loop<KS, L, R>(c); break;
        C3SC_ND(0, 1) C3SC_ND(0, 2) C3SC_ND(0, 3) C3SC_ND(0, 4)
        C3SC_ND(1, 0) C3SC_ND(1, 1) C3SC_ND(1, 2) C3SC_ND(1, 3) C3SC_ND(1, 4)
        C3SC_ND(2, 0) C3SC_ND(2, 1) C3SC_ND(2, 2) C3SC_ND(2, 3) C3SC_ND(2, 4)
        C3SC_ND(3, 0) C3SC_ND(3, 1) C3SC_ND(3, 2) C3SC_ND(3, 3) C3SC_ND(3, 4)
        C3SC_ND(4, 0) C3SC_ND(4, 1) C3SC_ND(4, 2) C3SC_ND(4, 3) C3SC_ND(4, 4)
#undef C3SC_ND
    }

    if constexpr (!FUSED) ft_active_list(a, nf, jb, je, sFid, sAbs, nmax);
    else {
        // The walk is a real call into another translation unit: its arguments must be addressable.  A kernel parameter
        // passed by reference would be copied to every thread's local stack (1.7 kB each); one shared copy per CTA instead.
        __shared__ CtlArgs sctl;
        {
            const int *src = reinterpret_cast<const int *>(ctl);
            int *dst = reinterpret_cast<int *>(&sctl);
            for (int e = tid; e < (int)(sizeof(CtlArgs) / sizeof(int)); e += FTN_NT) dst[e] = src[e];
        }
        __syncthreads();                                    // every neighbour value of the group is in the region
        FusedCta w;
        w.reg = region; w.RN = FT_FBMAX * njp; w.njp = njp; w.nf = nf; w.jb = jb; w.je = je; w.k = k; w.nmax = nmax;
        w.sFid = sFid; w.sAbs = sAbs; w.pi_eval = a.fuse_pi; w.arg = a.fuse_arg;
        switch (a.family) {
        case 0: fused_walk_lqg_lo(P.dx, sctl, w); break;
        case 1: fused_walk_lqg_hi(P.dx, sctl, w); break;
        default: fused_walk_misc(a.model, P.dx, sctl, w); break;
        }
        __syncthreads();
        if (tid == 0) { __threadfence(); atomicExch(a.ring_flag + s_region, 0); }
    }
}

template <int KS>
__global__ void __launch_bounds__(FTN_NT, 2) k_ft_nodes(const FtArgs a, const double *sets)
{
    ftn_body<KS, false>(a, sets, nullptr);
}
template <int KS>
__global__ void __launch_bounds__(FTN_NT, 2) k_ft_nodes_fused(const FtArgs a, const CtlArgs ctl, const double *sets)
{
    ftn_body<KS, true>(a, sets, &ctl);
}

}  // namespace c3sc
