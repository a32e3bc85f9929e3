// ft_kernel.cuh -- stage 1 of the Bellman backup: flags, neighbour indices and the
// function-train neighbour values of every node of every fiber of a batch.
//
//   process_fibers_neighbor   src/nodeutil.c:489-627   (flags + neighbour indices)
//   valuef_eval_fiber_ind_nn  src/valuefunc.c:369-585  (FT values at the node and its 2d
//                                                       axis neighbours, re-associated)
//   = mca_get_neighbor_costs  src/nodeutil.c:647-713   for F fibers at once.
//
// Independent of the dynamics / cost model.  Fibers are first grouped by their varying
// dimension k (k_group_fibers); one CTA then evaluates up to FB fibers that share k, so the
// blocks G_k[j] of the varying core are brought on chip once per group.
//
//   chains   (per fiber)  L = G_0[f_0]..G_{k-1}[f_{k-1}], R = G_{k+1}[f_{k+1}]..G_{d-1}[f_{d-1}] and the
//            2(d-1) neighbour variants of them (one centre block replaced by a neighbour's),
//            built as stepped vector-set x block products.  Operands come straight from
//            L2 with coalesced loads: the right side walks G (a + b r_m, lanes over a), the
//            left side walks the TRANSPOSED copy of the cores (b + a r_{m+1}, lanes over b)
//            that the value function keeps next to the original.
//   nodes    per tile of T nodes: stage G_k[j] in shared memory (odd column stride, so both
//            orientations are bank-conflict free), w_j = G_k[j] R and u_j = L G_k[j] for
//            all fibers of the group, then the (2d-1) length-r dots per (fiber, node).
//   output   costs in a slot-major scratch  cst[slot*NS + node]  (what the control kernel
//            reads coalesced), optionally the reference's node-major rows, flags, the
//            compacted list of non-absorbed nodes.
#pragma once
#include <cuda_runtime.h>
#include "dev_types.h"

namespace c3sc {

constexpr int FT_NT = 256;        // threads per CTA
constexpr int FT_FBMAX = 8;       // fibers per group
constexpr int FT_TMAX = 8;        // nodes per tile
constexpr int FT_VG = 8;          // chain: vectors per work item
constexpr int FT_DG = 4;          // dots: vectors per work item

struct FtArgs {
    DevProblem P;             // grid, boundary types, obstacles (model fields unused)
    DevFT ft;
    int F;                    // fibers in this launch (a chunk of the caller's batch)
    const int *dim_vary;      // [F]
    const int *fixed_ind;     // [F*d]
    int ldo;
    int FB;                   // fibers per group (<= FT_FBMAX)
    const int *perm;          // [F] fiber ids grouped by k
    const int *kcount;        // [d] fibers per k
    const int *kstart;        // [d] first position of k in perm
    // outputs
    double *cst;              // [(2d+1) * NS] slot-major scratch, may be NULL
    long long NS;             // F*ldo
    signed char *flag;        // [NS] 0 / 1 / -1, 2 = padding (j >= ngrid[k]); may be NULL
    int *act;                 // [NS] ids of nodes with flag 0; may be NULL
    int *act_count;
    double *costs;            // optional node-major rows [NS*(2d+1)]
    int *absorbed;            // optional [NS]
    int *nbr_vary;            // optional [NS*2]
    int *nbr_fixed;           // optional [F*2*(d-1)]
    const int *nbr_fixed_in;  // caller-supplied neighbour indices (valuef_eval_fiber_ind_nn)
    const int *nbr_vary_in;
    double *sets;             // [F * setw] chain records of the tensor-core path (chain_kernel.cuh); NULL = general kernel
    int setw, rs;             // record width and row stride in doubles
    int chains_done;          // the records were filled by the bucketed chain steps: skip k_ft_chains
    int *task_count;          // chain kernel: next (fiber, side) task, zeroed by k_group_fibers
    int nsplit;               // k_ft_nodes: CTAs per group, each owning a contiguous range of node tiles (blockIdx.y)
    // fused stage 2 (k_ft_nodes_fused): the CTA keeps its neighbour values in a region of a small ring and walks the control
    // set itself before it exits
    int fused;                // 0 = off
    int family, model;        // dispatcher of the walk (ctl_types.cuh) and the model id it switches on
    int fuse_pi, fuse_arg;    // policy evaluation against CtlArgs::rows_in; argmin / policy rows wanted
    double *ring;             // [nring][region_doubles]
    int *ring_flag;           // [nring] 0 = free; [nring] = the allocation counter
    int nring;
    long long region_doubles; // (2d+1) * FT_FBMAX * even(nmax)
};

__host__ __device__ inline int ft_even_up(int v) { return (v + 1) & ~1; }

// Chunks of a batch.  A super-chunk is FS = FC * SC fibers = SC equal chunks of FC fibers, in m >= SC chunk slots (unused slots are
// empty).  The host-buffer entries copy a chunk out while the next ones compute: the link is busy from the first finished chunk on
// and only the copy of a lane's last chunk is exposed, so
//   * super-chunks before `head_to` (each lane's first) have their FIRST chunk cut into nh + 1 pieces starting at
//     head[0] = 0 < head[1] < ... (small pieces first: results start to flow early),
//   * super-chunks from `taper_from` on (each lane's last) have their LAST chunk cut into nt + 1 pieces starting at
//     tail[0] = 0 < tail[1] < ... fibers past the first SC - 1 chunks (small pieces last).
// Chunk `ord` (= super-chunk * m + slot) owns perm[c0 ..] and the ord-th 64 ints of `cnt`.
struct ChunkLayout { int FS, FC, SC, m, head_to, taper_from, nh, nt; int head[4], tail[4]; };
__host__ __device__ inline void ft_chunk_range(const ChunkLayout &L, int F, int ord, int *c0, int *Fc)
{
    const int si = ord / L.m, ci = ord - si * L.m;
    const int H = (L.nh > 0 && si < L.head_to) ? L.nh : 0, T = (L.nt > 0 && si >= L.taper_from) ? L.nt : 0;
    int lo = L.FS, hi = L.FS;
    if (H > 0 && ci <= H) { lo = L.head[ci]; hi = ci < H ? L.head[ci + 1] : L.FC; }
    else {
        const int cj = ci - H;                               // position among the SC equal chunks
        if (T > 0 && cj >= L.SC - 1) {
            const int t = cj - (L.SC - 1), base = (L.SC - 1) * L.FC;
            if (t <= T) { lo = base + L.tail[t]; hi = t < T ? base + L.tail[t + 1] : L.FS; }
        } else if (cj < L.SC) { lo = cj * L.FC; hi = lo + L.FC; }
    }
    const long long b = (long long)si * L.FS;
    long long a0 = b + lo, a1 = b + hi;
    if (a1 > F) a1 = F;
    if (a0 > F) a0 = F;
    *c0 = (int)a0; *Fc = (int)(a1 - a0);
}

// node-tile size for a varying core with block r_k x r_{k+1}
__host__ __device__ inline int ft_tile_nodes(int rk, int rk1)
{
    int T = FT_NT / (rk + rk1);
    return T < 1 ? 1 : (T > FT_TMAX ? FT_TMAX : T);
}

// shared-memory carve-up; identical on host and device
struct FtPlan {
    int d, FB, rs, nvt, tp, nmax;
    int oSetA, oUni, oLt, oRt, nDoubles;          // doubles
    int setDoubles, gTile, oW, oU;                // inside the union region
    int oFix, oNf, oFid, oWall, nInts;            // ints
    __host__ __device__ FtPlan(const DevFT &ft, int nmax_, int FB_)
    {
        d = ft.d; FB = FB_; nmax = nmax_;
        rs = 1;
        for (int i = 0; i <= d; i++) rs = ft.r[i] > rs ? ft.r[i] : rs;
        nvt = 2 * d + 2;                          // vectors of both chain sets of one fiber (even-padded)
        tp = FT_TMAX | 1;                         // odd stride of a (fiber, rank index) row of w / u
        setDoubles = FB * rs * nvt;
        gTile = 0;
        int wu = 0;
        for (int k = 0; k < d; k++) {
            const int rk = ft.r[k], rk1 = ft.r[k + 1];
            const int g = ft_tile_nodes(rk, rk1) * (rk | 1) * rk1;
            gTile = g > gTile ? g : gTile;
            wu = (rk + rk1) > wu ? (rk + rk1) : wu;
        }
        gTile = ft_even_up(gTile);
        int o = 0;
        oSetA = o; o += setDoubles;
        oUni = o;
        oW = gTile;                                // union region: [G tile][w rows][u rows]  or  chain buffer B
        oU = oW;                                   // u rows follow the w rows (offset set per group: rk*FB*tp)
        const int node_need = gTile + wu * FB * tp;
        o += node_need > setDoubles ? node_need : setDoubles;
        o = ft_even_up(o);
        oLt = o; o += rs * FT_FBMAX;
        oRt = o; o += rs * FT_FBMAX;
        nDoubles = ft_even_up(o);
        int q = 0;
        oFix = q;  q += FB * d;
        oNf = q;   q += FB * 2 * d;
        oFid = q;  q += FT_FBMAX;
        oWall = q; q += FT_FBMAX;
        nInts = q;                                 // then FB*nmax bytes (sAbs)
    }
    __host__ __device__ size_t bytes() const { return (size_t)nDoubles * 8 + (size_t)nInts * 4 + (((size_t)FB * nmax + 7) & ~(size_t)7); }
};

#ifndef C3SC_FT_TYPES_ONLY
// ---------------------------------------------------------------------------
// Group the fibers of every chunk of a batch by varying dimension: perm = fiber ids (relative to the
// chunk), k-major.  One CTA per chunk (blockIdx.x); chunk c covers the fibers ft_chunk_range gives it and
// owns perm[c0 ..], and 64 ints of `cnt`: [0,16) kcount, [16,32) kstart, [32] active-node counter (cleared).
//
// The same pass validates the descriptors (the batch analogue of convert_fiber_to_ind's error returns,
// src/nodeutil.c:437-470): a dim_vary outside [0, d) or a fixed index outside its grid sets err[1] and records
// the smallest offending fiber id in err[2]; every consumer clamps what it loads, so a bad descriptor is an
// error return of the batch, never a read outside the cores.
struct GridDims { int n[MAXD]; };
__device__ __forceinline__ int ft_clamp_index(int i0, int n) { return i0 < 0 ? 0 : (i0 >= n ? n - 1 : i0); }
#ifndef C3SC_FT_KS_UNIT        // compiled once, in ft.cu (ft_ks.cu holds the per-rank-geometry templates)
__global__ void __launch_bounds__(1024) k_group_fibers(int F, ChunkLayout lay, int d, const int *dim_vary, const int *fixed_ind,
                                                       GridDims ng, int *err, int *perm, int *cnt_all)
{
    __shared__ int cnt[MAXD], pos[MAXD];
    const int tid = threadIdx.x;
    int c0, Fc;
    ft_chunk_range(lay, F, (int)blockIdx.x, &c0, &Fc);
    if (err && fixed_ind) {                                 // validation: the gridDim.y CTAs of a chunk share the scan
        int bad = 0x7fffffff;
        for (int e = blockIdx.y * blockDim.x + tid; e < Fc * d; e += gridDim.y * blockDim.x) {
            const int f = e / d, i = e - f * d;
            const int v = fixed_ind[(size_t)c0 * d + e];
            bool b = (unsigned)v >= (unsigned)ng.n[i];
            if (i == 0) b = b || (unsigned)dim_vary[c0 + f] >= (unsigned)d;
            if (b && c0 + f < bad) bad = c0 + f;
        }
        if (bad != 0x7fffffff) { atomicOr(err + 1, 1); atomicMin(err + 2, bad); }
    }
    if (blockIdx.y != 0) return;
    dim_vary += c0; perm += c0;
    int *kcount = cnt_all + 64 * blockIdx.x, *kstart = kcount + 16, *act_count = kcount + 32;
    if (tid < MAXD) cnt[tid] = 0;
    __syncthreads();
    // a core batch of the cross driver has ONE k: the counters are bumped once per warp and value (match_any), not 1024 times
    for (int f0 = 0; f0 < Fc; f0 += blockDim.x) {
        const int f = f0 + tid;
        int k = f < Fc ? dim_vary[f] : -1;
        k = f < Fc ? (k < 0 ? 0 : (k >= d ? d - 1 : k)) : MAXD;
        const unsigned peers = __match_any_sync(0xffffffffu, k);
        if (f < Fc && (int)(__ffs(peers) - 1) == (tid & 31)) atomicAdd(&cnt[k], __popc(peers));
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int k = 0; k < d; k++) { pos[k] = run; kstart[k] = run; kcount[k] = cnt[k]; run += cnt[k]; }
        *act_count = 0;
        kcount[33] = 0;                                     // task counter of the chain kernel (FtArgs::task_count)
    }
    __syncthreads();
    // the order inside a group does not change any result
    for (int f0 = 0; f0 < Fc; f0 += blockDim.x) {
        const int f = f0 + tid, lane = tid & 31;
        int k = f < Fc ? dim_vary[f] : -1;
        k = f < Fc ? (k < 0 ? 0 : (k >= d ? d - 1 : k)) : MAXD;
        const unsigned peers = __match_any_sync(0xffffffffu, k);
        const int leader = __ffs(peers) - 1;
        int base = 0;
        if (f < Fc && leader == lane) base = atomicAdd(&pos[k], __popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (f < Fc) perm[base + __popc(peers & ((1u << lane) - 1u))] = f;
    }
}
#endif

// Derived copies of the cores, rebuilt whenever the cores change (c3sc_valuef_commit):
//   baseT  every block transposed: baseT[off_k + j*blk + b + a*r_{k+1}] = base[off_k + j*blk + a + b*r_k]
//   baseP  (optional) zero-padded blocks, leading dimension ldp[k], cpp[k] columns
#ifndef C3SC_FT_KS_UNIT        // compiled once, in ft.cu (ft_ks.cu holds the per-rank-geometry templates)
__global__ void k_pack_cores(DevFT ft, double *baseT, double *baseP, double *baseQ)
{
    const int k = blockIdx.y;
    const int rk = ft.r[k], rk1 = ft.r[k + 1], blk = rk * rk1;
    const long long total = (long long)ft.n[k] * blk;
    const long long step = (long long)gridDim.x * blockDim.x, t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (long long e = t0; e < total; e += step) {
        const long long j = e / blk;
        const int rem = (int)(e - j * blk);
        const int a = rem / rk1, b = rem - a * rk1;            // e enumerates the transposed layout
        baseT[ft.off[k] + e] = ft.base[ft.off[k] + j * blk + a + (long long)b * rk];
    }
    if (baseP) {
        const int ld = ft.ldp[k], cp = ft.cpp[k], pb = ld * cp;
        const long long ptotal = (long long)ft.n[k] * pb;
        for (long long e = t0; e < ptotal; e += step) {
            const long long j = e / pb;
            const int rem = (int)(e - j * pb);
            const int b = rem / ld, a = rem - b * ld;
            baseP[ft.offP[k] + e] = (a < rk && b < rk1) ? ft.base[ft.off[k] + j * blk + a + (long long)b * rk] : 0.0;
        }
    }
    if (baseQ) {
        const int ld = ft.ldq[k], pb = ld * rk1;
        const long long qtotal = (long long)ft.n[k] * pb;
        for (long long e = t0; e < qtotal; e += step) {
            const long long j = e / pb;
            const int rem = (int)(e - j * pb);
            const int b = rem / ld, a = rem - b * ld;
            baseQ[ft.offQ[k] + e] = a < rk ? ft.base[ft.off[k] + j * blk + a + (long long)b * rk] : 0.0;
        }
    }
}
#endif

// ---------------------------------------------------------------------------
// pieces shared by the general kernel (k_ft_costs) and the tensor-core pair (ft_mma_kernel.cuh)

// which same-k group does this CTA own: groups are enumerated k-major, FB fibers each
__device__ __forceinline__ void ft_find_group(const FtArgs &a, int &k, int &gstart, int &nf)
{
    k = -1; gstart = 0; nf = 0;
    int b = blockIdx.x;
    for (int kk = 0; kk < a.ft.d; kk++) {
        const int c = a.kcount[kk];
        const int ng = (c + a.FB - 1) / a.FB;
        if (b < ng) { k = kk; gstart = a.kstart[kk] + b * a.FB; nf = c - b * a.FB; nf = nf > a.FB ? a.FB : nf; return; }
        b -= ng;
    }
}

// neighbour pair of a FIXED dimension i at index i0 (nodeutil.c:527-565); returns true when the
// index sits on an ABSORB face ("wall": every node of the fiber is absorbed)
__device__ __forceinline__ bool ft_fixed_pair(const DevProblem &P, int i, int i0, int &lo, int &hi)
{
    const int last = P.ngrid[i] - 1, bc = P.bc[i];
    bool wall = false;
    if (i0 == 0) {
        if (bc == C3SC_ABSORB)        { lo = i0; hi = i0; wall = true; }
        else if (bc == C3SC_REFLECT)  { lo = i0; hi = i0 + 1; }
        else                          { lo = P.ngrid[i] - 2; hi = i0 + 1; }
    } else if (i0 == last) {
        if (bc == C3SC_ABSORB)        { lo = i0; hi = i0; wall = true; }
        else if (bc == C3SC_REFLECT)  { lo = i0 - 1; hi = i0; }
        else                          { lo = i0 - 1; hi = 1; }
    } else { lo = i0 - 1; hi = i0 + 1; }
    return wall;
}

// descriptors, flags, neighbour indices of the nf fibers of a group (nodeutil.c:489-627).
// Whole CTA; ends with a barrier.
__device__ __forceinline__ void ft_flags_and_indices(const FtArgs &a, int k, int nf, int gstart, int jb, int je,
                                                     int *sFid, int *sWall, int *sFix, int *sNf, signed char *sAbs, int nmax)
{
    // nodes [jb, je) of every fiber of the group; the CTA that owns node 0 also marks the padding entries
    const DevProblem &P = a.P;
    const int d = a.ft.d, tid = threadIdx.x, NT = blockDim.x;
    const int N = P.ngrid[k], nj = je - jb;
    // one thread per (fiber, dimension): fiber id, fixed index and the neighbour pair of that dimension in one
    // dependent chain of loads (perm -> fixed_ind), one barrier before anything is consumed
    if (tid < FT_FBMAX) sWall[tid] = 0;
    __syncthreads();
    for (int e = tid; e < nf * d; e += NT) {
        const int g = e / d, i = e - g * d;
        const int fid = a.perm[gstart + g];
        const int i0 = ft_clamp_index(a.fixed_ind[(size_t)fid * d + i], P.ngrid[i]);
        if (i == 0) sFid[g] = fid;
        sFix[g * d + i] = i0;
        if (i == k) continue;
        const int slot = i < k ? i : i - 1;
        int lo, hi;
        if (ft_fixed_pair(P, i, i0, lo, hi)) sWall[g] = 1;
        if (a.nbr_fixed_in) {
            lo = a.nbr_fixed_in[(size_t)fid * 2 * (d - 1) + 2 * slot];
            hi = a.nbr_fixed_in[(size_t)fid * 2 * (d - 1) + 2 * slot + 1];
        }
        sNf[g * 2 * d + 2 * slot] = lo;
        sNf[g * 2 * d + 2 * slot + 1] = hi;
        if (a.nbr_fixed && jb == 0) {
            a.nbr_fixed[(size_t)fid * 2 * (d - 1) + 2 * slot] = lo;
            a.nbr_fixed[(size_t)fid * 2 * (d - 1) + 2 * slot + 1] = hi;
        }
    }
    if (tid >= nf && tid < FT_FBMAX) sFid[tid] = -1;
    __syncthreads();
    for (int e = tid; e < nf * nj; e += NT) {
        const int g = e / nj, j = jb + (e - g * nj);
        int ab = 0;
        for (int o = 0; o < P.nobs && ab == 0; o++) {                   // boundary.c:329-344,668-680
            const double *lb = P.obs + (size_t)o * 2 * d, *ub = lb + d;
            bool inside = true;
            for (int i = 0; i < d; i++) {
                const double x = P.xgrid[P.xoff[i] + (i == k ? j : sFix[g * d + i])];
                inside = inside && !(x < lb[i] || x > ub[i]);
            }
            if (inside) ab = -1;
        }
        if (sWall[g]) ab = 1;
        const int bk = P.bc[k];
        if (j == 0 || j == N - 1) ab = bk == C3SC_ABSORB ? 1 : 0;       // ends overwrite (nodeutil.c:570-612)
        const size_t id = (size_t)sFid[g] * a.ldo + j;
        sAbs[g * nmax + j] = (signed char)ab;
        if (a.flag) a.flag[id] = (signed char)ab;
        if (a.absorbed) a.absorbed[id] = ab;
        if (a.nbr_vary) {
            int lo, hi;
            ft_vary_pair(bk, N, j, ab, lo, hi);
            if (a.nbr_vary_in) { lo = a.nbr_vary_in[2 * id]; hi = a.nbr_vary_in[2 * id + 1]; }
            a.nbr_vary[2 * id] = lo; a.nbr_vary[2 * id + 1] = hi;
        }
    }
    if (a.flag && jb == 0)
        for (int e = tid; e < nf * (a.ldo - N); e += NT) {              // padding entries of ragged grids
            const int g = e / (a.ldo - N), j = N + (e - g * (a.ldo - N));
            a.flag[(size_t)sFid[g] * a.ldo + j] = 2;
        }
    __syncthreads();
}

// Compacted list of the non-absorbed nodes [jb, je) of the group's fibers: one warp per fiber counts,
// reserves a run of the list and writes the ids in node order.  sAbs must be visible.
__device__ __forceinline__ void ft_active_list(const FtArgs &a, int nf, int jb, int je, const int *sFid,
                                               const signed char *sAbs, int nmax)
{
    if (!a.act) return;
    const int tid = threadIdx.x, NT = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31;
    for (int g = warp; g < nf; g += NT / 32) {
        int cnt = 0;
        for (int j = jb + lane; j < je; j += 32) cnt += sAbs[g * nmax + j] == 0;
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        int base = 0;
        if (lane == 0 && cnt > 0) base = atomicAdd(a.act_count, cnt);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int j0 = jb; j0 < je; j0 += 32) {
            const int j = j0 + lane;
            const bool on = j < je && sAbs[g * nmax + j] == 0;
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (on) a.act[base + __popc(m & ((1u << lane) - 1))] = sFid[g] * a.ldo + j;
            base += __popc(m);
        }
    }
}

// The two neighbours ALONG the fiber (valuefunc.c:514-519) are the fiber's own values at other
// nodes: stage 1 stores only the self value (slot 2d); the slot-major scratch is completed by its
// consumer (control_kernel.cuh: load_costs), the node-major `costs` output by this kernel.
#ifndef C3SC_FT_KS_UNIT        // compiled once, in ft.cu (ft_ks.cu holds the per-rank-geometry templates)
__global__ void k_costs_along(const FtArgs a)
{
    const int d = a.ft.d, CS = 2 * d + 1;
    for (long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x; id < a.NS; id += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(id / a.ldo), j = (int)(id - (long long)f * a.ldo);
        int k = a.dim_vary[f];
        k = k < 0 ? 0 : (k >= d ? d - 1 : k);
        const int N = a.P.ngrid[k];
        if (j >= N) continue;
        int lo, hi;
        if (a.nbr_vary_in) { lo = a.nbr_vary_in[2 * id]; hi = a.nbr_vary_in[2 * id + 1]; }
        else ft_vary_pair(a.P.bc[k], N, j, a.flag ? a.flag[id] : a.absorbed[id], lo, hi);
        const size_t fb = (size_t)f * a.ldo;
        a.costs[id * CS + 2 * k] = a.costs[(fb + lo) * CS + 2 * d];
        a.costs[id * CS + 2 * k + 1] = a.costs[(fb + hi) * CS + 2 * d];
    }
}
#endif

// ---------------------------------------------------------------------------
#ifndef C3SC_FT_KS_UNIT        // compiled once, in ft.cu (ft_ks.cu holds the per-rank-geometry templates)
__global__ void __launch_bounds__(FT_NT, 2) k_ft_costs(const FtArgs a)
{
    const DevProblem &P = a.P;
    const DevFT &ft = a.ft;
    const int d = ft.d, CS = 2 * d + 1;
    const int tid = threadIdx.x;
    const int FB = a.FB;

    int k, gstart, nf;
    ft_find_group(a, k, gstart, nf);
    if (k < 0) return;

    extern __shared__ __align__(16) double smem[];
    const FtPlan sp(ft, P.nmax, FB);
    const int rs = sp.rs, TP = sp.tp, nmax = sp.nmax;
    double *bufA = smem + sp.oSetA, *bufB = smem + sp.oUni;
    double *sLt = smem + sp.oLt, *sRt = smem + sp.oRt;
    int *ismem = reinterpret_cast<int *>(smem + sp.nDoubles);
    int *sFix = ismem + sp.oFix, *sNf = ismem + sp.oNf;
    int *sFid = ismem + sp.oFid, *sWall = ismem + sp.oWall;
    signed char *sAbs = reinterpret_cast<signed char *>(ismem + sp.nInts);

    const int N = P.ngrid[k];
    const int NVL = ft_even_up(1 + 2 * k), NVR = ft_even_up(1 + 2 * (d - 1 - k));
    const int setStride = rs * sp.nvt;                 // one fiber's two sets inside a buffer
    const int offR = rs * NVL;                         // right set follows the left set

    ft_flags_and_indices(a, k, nf, gstart, 0, N, sFid, sWall, sFix, sNf, sAbs, nmax);

    // ---- 1. chains -----------------------------------------------------------------
    // set layout: element (q, v) of a set at q*NV + v  (q = rank index, v = vector, v fastest).
    // left set after step m:  v=0 the prefix L_{m+1}, v=1+2i+s the variant with G_i[nb_s], i<=m.
    // right set after step m' (dimension d-1-m'): v=0 the suffix, v=1+2m''+s the variant of
    // dimension d-1-m''.
    for (int e = tid; e < nf; e += FT_NT) {
        bufA[e * setStride] = 1.0;
        bufA[e * setStride + offR] = 1.0;
    }
    __syncthreads();
    const int nsteps = (k > d - 1 - k) ? k : d - 1 - k;
    for (int s = 0; s < nsteps; s++) {
        const bool doL = s < k, doR = s < d - 1 - k;
        const int nin = 1 + 2 * s, nvg = (nin + FT_VG - 1) / FT_VG;
        const double *in = (s & 1) ? bufB : bufA;
        double *out = (s & 1) ? bufA : bufB;
        const int mL = s, mR = d - 1 - s;
        const int roL = doL ? ft.r[mL + 1] : 0, roR = doR ? ft.r[mR] : 0;
        const int perL = nvg * roL, perR = nvg * roR;
        const int nitems = nf * (perL + perR);
        for (int e = tid; e < nitems; e += FT_NT) {
            const int g = e / (perL + perR);
            int rem = e - g * (perL + perR);
            const bool left = rem < perL;
            if (!left) rem -= perL;
            const int ro = left ? roL : roR;
            const int vg = rem / ro, o = rem - vg * ro;
            const int m = left ? mL : mR;
            const int rq = left ? ft.r[m] : ft.r[m + 1];
            const int NV = left ? NVL : NVR;
            const int blk = ft.r[m] * ft.r[m + 1];
            const double *base = (left ? ft.baseT : ft.base) + ft.off[m];
            const int slot = m < k ? m : m - 1;
            const double *gc = base + (size_t)sFix[g * d + m] * blk + o;
            const double *vin = in + g * setStride + (left ? 0 : offR);
            double *vout = out + g * setStride + (left ? 0 : offR);
            const int v0 = vg * FT_VG;
            double acc[FT_VG];
#pragma unroll
            for (int i = 0; i < FT_VG; i++) acc[i] = 0.0;
            if (vg == 0) {
                const double *glo = base + (size_t)sNf[g * 2 * d + 2 * slot] * blk + o;
                const double *ghi = base + (size_t)sNf[g * 2 * d + 2 * slot + 1] * blk + o;
                double alo = 0.0, ahi = 0.0;
#pragma unroll 4
                for (int q = 0; q < rq; q++) {
                    const double c = __ldg(gc + (size_t)q * ro), l = __ldg(glo + (size_t)q * ro), h = __ldg(ghi + (size_t)q * ro);
                    const double2 *row = reinterpret_cast<const double2 *>(vin + q * NV);
#pragma unroll
                    for (int i = 0; i < FT_VG / 2; i++) {
                        const double2 x = row[i];
                        acc[2 * i] = fma(x.x, c, acc[2 * i]);
                        acc[2 * i + 1] = fma(x.y, c, acc[2 * i + 1]);
                        if (i == 0) { alo = fma(x.x, l, alo); ahi = fma(x.x, h, ahi); }
                    }
                }
                vout[o * NV + nin] = alo;
                vout[o * NV + nin + 1] = ahi;
            } else {
#pragma unroll 4
                for (int q = 0; q < rq; q++) {
                    const double c = __ldg(gc + (size_t)q * ro);
                    const double2 *row = reinterpret_cast<const double2 *>(vin + q * NV + v0);
#pragma unroll
                    for (int i = 0; i < FT_VG / 2; i++) {
                        const double2 x = row[i];
                        acc[2 * i] = fma(x.x, c, acc[2 * i]);
                        acc[2 * i + 1] = fma(x.y, c, acc[2 * i + 1]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < FT_VG; i++)
                if (v0 + i < nin) vout[o * NV + v0 + i] = acc[i];
        }
        // a side that has already finished carries its set over unchanged
        if (!doL) {
            const int cnt = ft.r[k] * NVL;
            for (int e = tid; e < nf * cnt; e += FT_NT) { const int g = e / cnt, q = e - g * cnt; out[g * setStride + q] = in[g * setStride + q]; }
        }
        if (!doR) {
            const int cnt = ft.r[k + 1] * NVR;
            for (int e = tid; e < nf * cnt; e += FT_NT) { const int g = e / cnt, q = e - g * cnt; out[g * setStride + offR + q] = in[g * setStride + offR + q]; }
        }
        __syncthreads();
    }
    if (nsteps & 1) {                                   // final sets must live in buffer A (B is the node scratch)
        for (int e = tid; e < nf * setStride; e += FT_NT) bufA[e] = bufB[e];
        __syncthreads();
    }
    const int rk = ft.r[k], rk1 = ft.r[k + 1];
    // transposed prefix / suffix for the node mat-vecs: Lt[a*FBMAX + g], Rt[b*FBMAX + g] (0 for g >= nf)
    for (int e = tid; e < rk * FT_FBMAX; e += FT_NT) {
        const int aa = e / FT_FBMAX, g = e - aa * FT_FBMAX;
        sLt[e] = g < nf ? bufA[g * setStride + aa * NVL] : 0.0;
    }
    for (int e = tid; e < rk1 * FT_FBMAX; e += FT_NT) {
        const int b = e / FT_FBMAX, g = e - b * FT_FBMAX;
        sRt[e] = g < nf ? bufA[g * setStride + offR + b * NVR] : 0.0;
    }
    __syncthreads();

    // ---- 2. node tiles ----------------------------------------------------------------
    {
        const int ldk = rk | 1, blk = rk * rk1;
        const int T = ft_tile_nodes(rk, rk1);
        double *sG = bufB, *sW = bufB + sp.gTile, *sU = sW + rk * FB * TP;
        const double *Gk = ft.base + ft.off[k];
        const unsigned magic = (unsigned)((0x100000000ULL + rk - 1) / rk);        // e / rk for e < 2^32 / rk
        const int ngL = (1 + 2 * k + FT_DG - 1) / FT_DG;
        const int ngR = (d - 1 - k > 0) ? (1 + 2 * (d - 1 - k) + FT_DG - 1) / FT_DG : 0;
        for (int j0 = 0; j0 < N; j0 += T) {
            const int nt = (N - j0 < T) ? N - j0 : T;
            // stage the tile: column c = jl*rk1 + b of length rk -> sG[c*ldk + a]
            {
                const double *src = Gk + (size_t)j0 * blk;
                const int cnt = nt * blk;
                for (int e = tid; e < cnt; e += FT_NT) {
                    const int c = rk == 1 ? e : (int)__umulhi((unsigned)e, magic);
                    const int aa = e - c * rk;
                    sG[c * ldk + aa] = __ldg(src + e);
                }
            }
            __syncthreads();
            // w[g] = G_k[j] R_g (row a), u[g] = L_g G_k[j] (column b), all fibers of the group at once.
            // Items are ordered all-w then all-u so a warp runs one of the two loops, not both.
            {
                const int nW = nt * rk, nU = nt * rk1;
                for (int e = tid; e < nW + nU; e += FT_NT) {
                    double acc[FT_FBMAX];
#pragma unroll
                    for (int g = 0; g < FT_FBMAX; g++) acc[g] = 0.0;
                    if (e < nW) {
                        const int jl = e / rk, q = e - jl * rk;
                        const double *gp = sG + jl * rk1 * ldk + q;
#pragma unroll 2
                        for (int b = 0; b < rk1; b++) {
                            const double gv = gp[b * ldk];
                            const double2 *rr = reinterpret_cast<const double2 *>(sRt + b * FT_FBMAX);
#pragma unroll
                            for (int g = 0; g < FT_FBMAX / 2; g++) {
                                const double2 x = rr[g];
                                acc[2 * g] = fma(gv, x.x, acc[2 * g]);
                                acc[2 * g + 1] = fma(gv, x.y, acc[2 * g + 1]);
                            }
                        }
#pragma unroll
                        for (int g = 0; g < FT_FBMAX; g++)
                            if (g < nf) sW[(g * rk + q) * TP + jl] = acc[g];
                    } else {
                        const int e2 = e - nW;
                        const int jl = e2 / rk1, b = e2 - jl * rk1;
                        const double *gp = sG + (jl * rk1 + b) * ldk;
#pragma unroll 2
                        for (int aa = 0; aa < rk; aa++) {
                            const double gv = gp[aa];
                            const double2 *ll = reinterpret_cast<const double2 *>(sLt + aa * FT_FBMAX);
#pragma unroll
                            for (int g = 0; g < FT_FBMAX / 2; g++) {
                                const double2 x = ll[g];
                                acc[2 * g] = fma(gv, x.x, acc[2 * g]);
                                acc[2 * g + 1] = fma(gv, x.y, acc[2 * g + 1]);
                            }
                        }
#pragma unroll
                        for (int g = 0; g < FT_FBMAX; g++)
                            if (g < nf) sU[(g * rk1 + b) * TP + jl] = acc[g];
                    }
                }
            }
            __syncthreads();
            // dots: left variants against w, right variants against u
            {
                const int perf = (ngL + ngR) * nt;
                for (int e = tid; e < nf * perf; e += FT_NT) {
                    const int g = e / perf;
                    int rem = e - g * perf;
                    const int grp = rem / nt, jl = rem - grp * nt, j = j0 + jl;
                    const bool left = grp < ngL;
                    const int v0 = (left ? grp : grp - ngL) * FT_DG;
                    const int NV = left ? NVL : NVR, rr = left ? rk : rk1;
                    const double *vec = bufA + g * setStride + (left ? 0 : offR) + v0;
                    const double *wu = (left ? sW + g * rk * TP : sU + g * rk1 * TP) + jl;
                    double acc[FT_DG];
#pragma unroll
                    for (int i = 0; i < FT_DG; i++) acc[i] = 0.0;
#pragma unroll 2
                    for (int q = 0; q < rr; q++) {
                        const double x = wu[q * TP];
                        const double2 *row = reinterpret_cast<const double2 *>(vec + q * NV);
#pragma unroll
                        for (int i = 0; i < FT_DG / 2; i++) {
                            const double2 y = row[i];
                            acc[2 * i] = fma(y.x, x, acc[2 * i]);
                            acc[2 * i + 1] = fma(y.y, x, acc[2 * i + 1]);
                        }
                    }
                    const size_t id = (size_t)sFid[g] * a.ldo + j;
                    const int nvec = left ? 1 + 2 * k : 1 + 2 * (d - 1 - k);
#pragma unroll
                    for (int i = 0; i < FT_DG; i++) {
                        const int v = v0 + i;
                        if (v >= nvec) continue;
                        int slot;
                        if (v == 0) {
                            if (!left) continue;                       // suffix . u == self again
                            slot = 2 * d;
                        } else {
                            const int st = (v - 1) >> 1, side = (v - 1) & 1;
                            slot = 2 * (left ? st : d - 1 - st) + side;
                        }
                        if (a.cst) a.cst[(size_t)slot * a.NS + id] = acc[i];
                        if (a.costs) a.costs[id * CS + slot] = acc[i];
                    }
                }
            }
            __syncthreads();
        }
    }

    ft_active_list(a, nf, 0, N, sFid, sAbs, nmax);
}
#endif

#endif  // C3SC_FT_TYPES_ONLY

inline int ft_pick_fb(const DevFT &ft, int nmax, int F, int sms, size_t smem_budget)
{
    // enough groups to fill the machine twice over, and a carve-up that lets two CTAs share an SM
    int fb = FT_FBMAX;
    while (fb > 1 && (F / fb) < 2 * sms) fb >>= 1;
    while (fb > 1 && FtPlan(ft, nmax, fb).bytes() > smem_budget) fb >>= 1;
    return fb;
}

}  // namespace c3sc
