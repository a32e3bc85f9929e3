// chain_kernel.cuh -- stage 1a for LARGE batches: the prefix / suffix vector sets of valuef_eval_fiber_ind_nn
// (src/valuefunc.c:414-446 and the neighbour variants of :522-582) for all fibers of a chunk at once, advanced
// one dimension per launch as a streaming GEMM per core block.
//
// A fiber's left set after step s holds 1 + 2(s+1) row vectors:  v = 0 the prefix  L_{s+1} = G_0[f_0] .. G_s[f_s],
// v = 1 + 2i + side the variant in which G_i[f_i] is replaced by the neighbour block G_i[nb_side], i <= s.
// Step s multiplies every existing vector by the centre block G_s[f_s] and appends  L_s . G_s[nb_lo|hi].  The right
// set is the mirror image (suffixes, dimensions d-1, d-2, ..).  Random fibers share no chain, but they do share
// BLOCKS: with F fibers and N nodes per dimension F/N fibers use the same G_m[j] as centre and 2F/N as neighbour.
// So step s is organised per block: all rows (vectors of whatever fiber) that meet G_m[j] form the A operand of one
//      [rows x r_in] . [r_in x r_out]      FP64 tensor-core product (DMMA m8n8k4)
// whose B fragments sit in registers for the whole bucket and whose A rows stream from L2 and back, 8 rows per warp
// step.  Every DMMA row is a useful vector (the per-fiber kernel k_ft_chains spends a whole 8-row tile on each of
// the two neighbour products), and a block is read once per warp instead of once per fiber.
//
//   k_chain_count / k_chain_scan / k_chain_scatter
//                  counting sort of the chunk's (fiber, role) entries by (side, dimension, block); roles: centre,
//                  lower neighbour, upper neighbour.  Also the inverse: the ROW a fiber's centre block / lower /
//                  upper slot has in the dimension's bucket order.
//   k_chain_link   per row of every launch: where the row's product goes (the fiber's rows in the NEXT dimension's
//                  bucket order, or the fiber's record when the side is done), and per tile its bucket.
//   k_chain_step   launch t = 0..d-2 advances the left sets through dimension t and the right sets through d-1-t.
//
// SCATTER ON WRITE, STREAM ON READ.  The rows a launch works on lie in bucket order in a row buffer X[t & 1]: tile i of
// the launch is rows [8i, 8i+8) -- one contiguous 8*RS-double piece, fetched by ONE TMA bulk copy (cp.async.bulk +
// mbarrier) into a per-warp ring of shared-memory slots, four tiles ahead.  Nothing a load needs depends on another
// load.  The products go where launch t+1 will read them (X[(t+1) & 1], at the rows k_chain_link recorded; a prefix is
// written three times: into the fiber's centre block and into the two neighbour slots), or, when the fiber's side is
// complete, into the fiber's record.  (The first version gathered rows through a tag -> row address chain of two
// dependent L2 round trips per tile and spent 48 % of its issue slots waiting for them, profiles/r02b_chain.md.)
//
// Record of fiber f (FtArgs::sets + f*setw, rows of RS = 4*KS doubles, zero beyond the rank):
//   rows [0, 1+2k)        left set,  row v = vector v          (k = dim_vary)
//   rows [1+2k, 2d)       right set
//   rows 2d .. 2d+3       unused (the gathering version kept the double-buffered prefix / suffix there)
#pragma once
#include "ft_kernel.cuh"

namespace c3sc {

#ifndef C3SC_CH_NT
#define C3SC_CH_NT 128
#endif
constexpr int CH_NT = C3SC_CH_NT;     // step kernel: 4 warps per CTA, six CTAs per SM (256 x 2 per SM: +1 % time)
constexpr int CH_PREFIX = 1 << 30;    // ChainArgs::rowd[]: the row (a prefix) also goes to the two slots in rowp[]
constexpr int CH_NMAX = 768;          // nodes per dimension the plan kernel's shared-memory histogram covers (14 ints each)

__host__ __device__ inline int ft_rec_rs(int rmax) { return (rmax + 3) & ~3; }                // row stride of a record
__host__ __device__ inline int ft_rec_width(int d, int rmax) { return (2 * d + 4) * ft_rec_rs(rmax); }

struct ChainArgs {
    DevProblem P;
    DevFT ft;
    int F;                    // fibers of this (super-)chunk
    const int *dim_vary;      // [F]
    const int *fixed_ind;     // [F*d]
    const int *nbr_fixed_in;  // optional caller-supplied neighbour pairs [F*2*(d-1)]
    double *sets;             // [F*setw]
    int setw, rs;
    // plan of this chunk (device, written by k_chain_count / _scan / _scatter / _link):
    int *kst;                 // [d][2][nmax*3 + 1]  start of (block j, role) inside the dimension's entry list
    int *tst;                 // [d][2][nmax + 1]    first 8-row tile of block j; [..][N] = number of tiles
    int *ent;                 // [d][entstride]      entries: fiber | k << 24
    int nmax, entstride;      // entstride = 3 * (fibers of a full chunk)
    int *inv;                 // [d][invstride][3]   row of the fiber's centre block / lower slot / upper slot in dimension m's
                              //                     bucket order, relative to its side's first row
    int *rowd;                // [d-1][xrows]        per row of launch t's buffer, where the product goes: >= 0 a row of launch
                              //                     t+1's buffer (+ CH_PREFIX when the row is a prefix that two neighbour slots of
                              //                     launch t+1 take as well), < 0: ~(row of the chunk's records, in units of RS
                              //                     doubles), INT_MIN: padding row
    int2 *rowp;               // [d-1][xrows]        those two slots (rows of launch t+1's buffer), prefix rows only
    int *tileb;               // [d-1][xrows / 8]    per tile of launch t: side * 65536 + block of the tile's bucket
    int invstride;            // fibers of a full chunk
    double *x[2];             // row buffers of the launches, X[t & 1] read by launch t; xrows rows of RS doubles each
    long long xrows;
};
// rows a launch's buffer must hold: a fiber has at most 2d rows in one launch, every bucket is padded to 8 rows
__host__ __device__ inline long long chain_x_rows(int d, int nmax, long long F) { return F * 2 * d + 16LL * nmax + 8; }

#ifndef C3SC_FT_TYPES_ONLY
// exclusive scan of n ints in shared memory by one CTA (n <= 8 * blockDim.x); returns the total
__device__ __forceinline__ int cta_exclusive_scan(int *v, int n, int *wsum)
{
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + NT - 1) / NT;
    const int b = tid * per, e = (b + per < n) ? b + per : n;
    int s = 0;
    for (int i = b; i < e; i++) s += v[i];
    int x = s;
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = lane < NT / 32 ? wsum[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
        wsum[lane] = w;                                   // inclusive over warps
    }
    __syncthreads();
    int run = (warp ? wsum[warp - 1] : 0) + x - s;          // exclusive prefix of this thread's range
    const int total = wsum[NT / 32 - 1];
    for (int i = b; i < e; i++) { const int c = v[i]; v[i] = run; run += c; }
    __syncthreads();
    return total;
}

// The plan of a batch: chunk c covers fibers [c*FC, min(F, (c+1)*FC)); its plan arrays are the c-th slices of the buffers
// in `a` (strides given).  Three launches:
//   k_chain_count    one CTA per (slice of 1024 fibers, dimension): the (side, block, role) bins, counted in shared memory
//                    (the first version ran one 1024-thread CTA per (chunk, dimension) with a shared-memory histogram: 40 CTAs
//                    on 148 SMs and two latency-bound passes over the descriptors, 60 us per 65 536 fibers);
//   k_chain_scan     one CTA per (chunk, dimension): tiles per block, the exclusive scans -> kst, tst, the running positions;
//   k_chain_scatter  the same slices again: entries into their bins, and the inverse (rows of the fiber).
struct ChainPlanStrides { long long kst, tst, ent, inv, rowd; };

// roles of fiber f (batch index) in dimension m; false when m is the varying dimension
__device__ __forceinline__ bool chain_roles(const ChainArgs &a, int f, int m, int i0raw, int &k, int &side, int &i0, int &lo, int &hi)
{
    const int d = a.ft.d, N = a.P.ngrid[m];
    k = a.dim_vary[f];
    k = k < 0 ? 0 : (k >= d ? d - 1 : k);
    if (k == m) return false;
    side = m < k ? 0 : 1;
    i0 = ft_clamp_index(i0raw, N);
    ft_fixed_pair(a.P, m, i0, lo, hi);
    if (a.nbr_fixed_in) {
        const int slot = m < k ? m : m - 1;
        lo = ft_clamp_index(a.nbr_fixed_in[(size_t)f * 2 * (d - 1) + 2 * slot], N);
        hi = ft_clamp_index(a.nbr_fixed_in[(size_t)f * 2 * (d - 1) + 2 * slot + 1], N);
    }
    return true;
}

#ifndef C3SC_FT_KS_UNIT
constexpr int CHP_NT = 512, CHP_U = 2, CHP_SLICE = CHP_NT * CHP_U;      // fibers per CTA of the count / scatter kernels
// grid (slices of a chunk, d, chunks), CHP_NT threads, 6 * nmax ints of shared memory (12 * nmax for the scatter).
// cntg: [chunk][d][2][nmax*3 + 1] like kst, zero on entry.  A CTA counts its slice of the chunk's fibers in shared memory
// and adds every non-empty bin to the global count with one atomic.
__global__ void __launch_bounds__(CHP_NT) k_chain_count(ChainArgs a, int FC, ChainPlanStrides S, int *cntg)
{
    extern __shared__ int sh[];
    const int d = a.ft.d, nmax = a.nmax, m = blockIdx.y, c = blockIdx.z, tid = threadIdx.x;
    const int N = a.P.ngrid[m], nb = 6 * N;
    const int f0 = c * FC + blockIdx.x * CHP_SLICE;
    int fend = (c + 1) * FC < a.F ? (c + 1) * FC : a.F;
    fend = f0 + CHP_SLICE < fend ? f0 + CHP_SLICE : fend;
    for (int e = tid; e < nb; e += CHP_NT) sh[e] = 0;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < CHP_U; u++) {
        const int f = f0 + u * CHP_NT + tid;
        int k, side, i0, lo, hi;
        if (f < fend && chain_roles(a, f, m, a.fixed_ind[(size_t)f * d + m], k, side, i0, lo, hi)) {
            atomicAdd(&sh[(side * N + i0) * 3], 1);
            atomicAdd(&sh[(side * N + lo) * 3 + 1], 1);
            atomicAdd(&sh[(side * N + hi) * 3 + 2], 1);
        }
    }
    __syncthreads();
    int *cg = cntg + c * S.kst + (size_t)m * 2 * (nmax * 3 + 1);
    for (int e = tid; e < nb; e += CHP_NT)
        if (sh[e]) atomicAdd(cg + (e / (3 * N)) * (nmax * 3 + 1) + (e % (3 * N)), sh[e]);
}

// grid (chunks, d), 1024 threads, 14 * nmax ints of shared memory
__global__ void __launch_bounds__(1024) k_chain_scan(ChainArgs a, ChainPlanStrides S, int *cntg)
{
    extern __shared__ int sh[];
    __shared__ int wsum[32];
    const int d = a.ft.d, m = blockIdx.y, tid = threadIdx.x, NT = blockDim.x;
    const int N = a.P.ngrid[m], nmax = a.nmax;
    int *kst = a.kst + blockIdx.x * S.kst + (size_t)m * 2 * (nmax * 3 + 1);
    int *tst = a.tst + blockIdx.x * S.tst + (size_t)m * 2 * (nmax + 1);
    int *cg = cntg + blockIdx.x * S.kst + (size_t)m * 2 * (nmax * 3 + 1);
    int *cnt = sh, *til = sh + 2 * 3 * N;                   // [side][j][role]; tiles per (side, j)
    const int nb = 2 * 3 * N;
    for (int e = tid; e < nb; e += NT) cnt[e] = cg[(e / (3 * N)) * (nmax * 3 + 1) + (e % (3 * N))];
    __syncthreads();
    // tiles per block: ceil((centre * nin + lo + hi) / 8), nin = 1 + 2s vectors before the step; the left sets
    // reach dimension m at step m, the right sets at step d-1-m
    for (int e = tid; e < 2 * N; e += NT) {
        const int side = e / N;
        const int nin = 1 + 2 * (side ? d - 1 - m : m);
        const int rows = cnt[e * 3] * nin + cnt[e * 3 + 1] + cnt[e * 3 + 2];
        til[e] = (rows + 7) >> 3;
    }
    __syncthreads();
    const int total = cta_exclusive_scan(cnt, nb, wsum);    // both sides in one list: side 1 starts where side 0 ends
    for (int side = 0; side < 2; side++) {                  // tile prefix per side
        int *t = til + side * N;
        const int tt = cta_exclusive_scan(t, N, wsum);
        for (int e = tid; e < N; e += NT) tst[side * (nmax + 1) + e] = t[e];
        if (tid == 0) tst[side * (nmax + 1) + N] = tt;
    }
    for (int e = tid; e < nb; e += NT) {
        const int o = (e / (3 * N)) * (nmax * 3 + 1) + (e % (3 * N));
        kst[o] = cnt[e];
        cg[o] = cnt[e];                                     // running positions of the scatter
    }
    if (tid == 0) { kst[3 * N] = cnt[3 * N]; kst[(nmax * 3 + 1) + 3 * N] = total; }   // end of side 0 = start of side 1; end of side 1
}

// Same grid.  The CTA counts its slice again, reserves a range of every non-empty bin with one global atomic, and places
// its entries inside the ranges with shared-memory atomics; with the position it knows the row, and writes the inverse.
__global__ void __launch_bounds__(CHP_NT) k_chain_scatter(ChainArgs a, int FC, ChainPlanStrides S, int *cntg)
{
    extern __shared__ int sh[];
    const int d = a.ft.d, nmax = a.nmax, m = blockIdx.y, c = blockIdx.z, tid = threadIdx.x;
    const int N = a.P.ngrid[m], nb = 6 * N;
    int *run = sh, *base = sh + nb;
    const int f0 = c * FC + blockIdx.x * CHP_SLICE;
    int fend = (c + 1) * FC < a.F ? (c + 1) * FC : a.F;
    fend = f0 + CHP_SLICE < fend ? f0 + CHP_SLICE : fend;
    for (int e = tid; e < nb; e += CHP_NT) run[e] = 0;
    __syncthreads();
    int k[CHP_U], side[CHP_U], i0[CHP_U], lo[CHP_U], hi[CHP_U];
    bool on[CHP_U];
#pragma unroll
    for (int u = 0; u < CHP_U; u++) {
        const int f = f0 + u * CHP_NT + tid;
        on[u] = f < fend && chain_roles(a, f, m, a.fixed_ind[(size_t)f * d + m], k[u], side[u], i0[u], lo[u], hi[u]);
        if (on[u]) {
            atomicAdd(&run[(side[u] * N + i0[u]) * 3], 1);
            atomicAdd(&run[(side[u] * N + lo[u]) * 3 + 1], 1);
            atomicAdd(&run[(side[u] * N + hi[u]) * 3 + 2], 1);
        }
    }
    __syncthreads();
    int *cg = cntg + c * S.kst + (size_t)m * 2 * (nmax * 3 + 1);
    for (int e = tid; e < nb; e += CHP_NT) {
        const int n = run[e];
        base[e] = n ? atomicAdd(cg + (e / (3 * N)) * (nmax * 3 + 1) + (e % (3 * N)), n) : 0;
        run[e] = 0;
    }
    __syncthreads();
    const int *kst = a.kst + c * S.kst + (size_t)m * 2 * (nmax * 3 + 1);
    const int *tst = a.tst + c * S.tst + (size_t)m * 2 * (nmax + 1);
    int *ent = a.ent + c * S.ent + (size_t)m * a.entstride;
#pragma unroll
    for (int u = 0; u < CHP_U; u++) {
        if (!on[u]) continue;
        const int fl = f0 + u * CHP_NT + tid - c * FC;
        int *inv = a.inv + c * S.inv + ((size_t)m * a.invstride + fl) * 3;
        const int tag = fl | (k[u] << 24);
        const int nin = 1 + 2 * (side[u] ? d - 1 - m : m);
        const int *ks = kst + side[u] * (nmax * 3 + 1), *tpre = tst + side[u] * (nmax + 1);     // first entry / tile of block j
        {   // centre: the fiber's nin rows start at (entry - first centre entry) * nin inside the block
            const int b = (side[u] * N + i0[u]) * 3, pos = base[b] + atomicAdd(&run[b], 1);
            ent[pos] = tag;
            inv[0] = 8 * tpre[i0[u]] + (pos - ks[3 * i0[u]]) * nin;
        }
        {   // neighbour slots follow the block's centre rows, one row per entry, lower entries first
            const int b = (side[u] * N + lo[u]) * 3 + 1, pos = base[b] + atomicAdd(&run[b], 1);
            ent[pos] = tag;
            inv[1] = 8 * tpre[lo[u]] + (ks[3 * lo[u] + 1] - ks[3 * lo[u]]) * nin + (pos - ks[3 * lo[u] + 1]);
        }
        {
            const int b = (side[u] * N + hi[u]) * 3 + 2, pos = base[b] + atomicAdd(&run[b], 1);
            ent[pos] = tag;
            inv[2] = 8 * tpre[hi[u]] + (ks[3 * hi[u] + 1] - ks[3 * hi[u]]) * nin + (pos - ks[3 * hi[u] + 1]);
        }
    }
}
#endif

// grid (chunks, d, Z), 256 threads: the row descriptors.  CTA (c, m, z) walks buckets z, z+Z, .. of dimension m (both sides)
// and writes, for every row of the bucket, where the row's product goes: the fiber's rows in the next dimension of its side
// (m+1 left, m-1 right; k_chain_scatter's inverse rows), or the fiber's record when the side ends here.
#ifndef C3SC_FT_KS_UNIT
constexpr int CH_ROW_PAD = (int)0x80000000;
constexpr int CH_LINK_E = 1024;       // centre entries of a bucket staged per round
__global__ void __launch_bounds__(256) k_chain_link(ChainArgs a, ChainPlanStrides S)
{
    const int d = a.ft.d, m = blockIdx.y, nmax = a.nmax;
    const int N = a.P.ngrid[m];
    const int *kstm = a.kst + blockIdx.x * S.kst + (size_t)m * 2 * (nmax * 3 + 1);
    const int *tstm = a.tst + blockIdx.x * S.tst + (size_t)m * 2 * (nmax + 1);
    const int *tst0 = a.tst + blockIdx.x * S.tst;
    const int *ent = a.ent + blockIdx.x * S.ent + (size_t)m * a.entstride;
    const int *inv = a.inv + blockIdx.x * S.inv;
    int *rowd = a.rowd + blockIdx.x * S.rowd, *tileb = a.tileb + blockIdx.x * (S.rowd / 8);
    int2 *rowp = a.rowp + blockIdx.x * S.rowd;
    const int recrows = a.setw / a.rs;
    __shared__ int s_i0[CH_LINK_E];
    __shared__ bool s_ct[CH_LINK_E];
    for (int b = blockIdx.z; b < 2 * N; b += gridDim.z) {
        const int side = b / N, j = b - side * N;
        const int t = side ? d - 1 - m : m;                 // the launch that works on this side of dimension m
        if (t > d - 2) continue;                            // (left side of the last dimension / right side of the first: empty)
        const int nin = 1 + 2 * t;
        const int *ks = kstm + side * (nmax * 3 + 1) + 3 * j;
        const int csc = ks[0], cslo = ks[1], cshi = ks[2], csend = ks[3];
        const int t0 = tstm[side * (nmax + 1) + j], t1 = tstm[side * (nmax + 1) + j + 1];
        const int nc = (cslo - csc) * nin, rows = nc + (csend - cslo);
        // rows of the right side lie behind the left side's tiles, in this launch's buffer and in the next one's
        const long long here = side ? 8LL * tst0[((size_t)t * 2 + 0) * (nmax + 1) + a.P.ngrid[t]] : 0;
        const int next = (side && t + 1 <= d - 2) ? 8 * tst0[((size_t)(t + 1) * 2 + 0) * (nmax + 1) + a.P.ngrid[t + 1]] : 0;
        int *out = rowd + (size_t)t * a.xrows + here + 8LL * t0;
        int2 *outp = rowp + (size_t)t * a.xrows + here + 8LL * t0;
        int *outb = tileb + (size_t)t * (a.xrows / 8) + here / 8 + t0;
        const int mn = side ? m - 1 : m + 1;
        const int bw = side * 65536 + j;
        // where an entry's fiber goes next: loaded once per ENTRY (tag, then the inverse rows of the next dimension)
        auto entry = [&](int e, bool centre, int &i0, bool &cont) {
            const int tag = ent[e], f = tag & 0xffffff, k = tag >> 24;
            cont = side ? mn > k : mn < k;
            if (cont) {
                const int *iv = inv + ((size_t)mn * a.invstride + f) * 3;
                i0 = next + iv[0];
                if (centre) outp[(size_t)(e - csc) * nin] = make_int2(next + iv[1], next + iv[2]);      // the prefix's two extra slots
            } else i0 = ~(f * recrows + (side ? 1 + 2 * k : 0));
        };
        // centre entries, CH_LINK_E at a time: per entry into shared memory, then the entries' nin rows each with consecutive
        // threads on consecutive rows (products keep their vector index v: row i0 + v of the next buffer, i0 - v of a record)
        for (int eb = 0; eb < cslo - csc; eb += CH_LINK_E) {
            const int n = (cslo - csc - eb < CH_LINK_E) ? cslo - csc - eb : CH_LINK_E;
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                int i0; bool cont;
                entry(csc + eb + i, true, i0, cont);
                s_i0[i] = i0; s_ct[i] = cont;
            }
            __syncthreads();
            int *o = out + (size_t)eb * nin;
            for (int r = threadIdx.x; r < n * nin; r += blockDim.x) {
                const int q = r / nin, v = r - q * nin;
                const bool cont = s_ct[q];
                o[r] = cont ? (s_i0[q] + v) | (v == 0 ? CH_PREFIX : 0) : s_i0[q] - v;
            }
        }
        // neighbour entries: one row each, the new vector nin (lower) or nin + 1 (upper)
        for (int i = threadIdx.x; i < csend - cslo; i += blockDim.x) {
            int i0; bool cont;
            entry(cslo + i, false, i0, cont);
            const int v = cslo + i < cshi ? nin : nin + 1;
            out[nc + i] = cont ? i0 + v : i0 - v;
        }
        for (int r = rows + threadIdx.x; r < 8 * (t1 - t0); r += blockDim.x) out[r] = CH_ROW_PAD;
        for (int tl = threadIdx.x; tl < t1 - t0; tl += blockDim.x) outb[tl] = bw;
    }
}
#endif

__device__ __forceinline__ void ch_dmma(double &d0, double &d1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// TMA bulk copy + mbarrier (the node kernel has its own copies of these in ft_mma_kernel.cuh)
__device__ __forceinline__ unsigned ch_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ch_mbar_init(unsigned long long *b, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ch_smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void ch_fetch(void *dst, const void *src, unsigned bytes, unsigned long long *b)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ch_smem_u32(b)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(ch_smem_u32(dst)), "l"(src), "r"(bytes), "r"(ch_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void ch_mbar_wait(unsigned long long *b, unsigned parity)
{
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "CH_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 "@P1 bra CH_DONE;\n"
                 "bra CH_WAIT;\n"
                 "CH_DONE:\n"
                 "}" ::"r"(ch_smem_u32(b)), "r"(parity) : "memory");
}

// One launch = one step of both sides.  Warps own contiguous ranges of 8-row tiles of the concatenated tile list
// [left buckets of dimension t | right buckets of dimension d-1-t] = contiguous pieces of the row buffer X[t & 1].
// A warp keeps CH_SLOTS tiles in flight (TMA bulk copies into its own ring) and the row descriptors of the next tile
// (only the stores and the choice of the block need them) one tile ahead.  No table, no decoding: everything a row
// needs is in its descriptor (k_chain_link).
#ifndef C3SC_CH_SLOTS
#define C3SC_CH_SLOTS 4
#endif
constexpr int CH_SLOTS = C3SC_CH_SLOTS;

// dynamic shared memory of k_chain_step: the rings (16-byte aligned), then the mbarriers
__host__ __device__ inline size_t chain_step_smem(int rs, int *bar_off = nullptr)
{
    size_t o = (size_t)(CH_NT / 32) * CH_SLOTS * 8 * rs * sizeof(double);
    if (bar_off) *bar_off = (int)o;
    o += (size_t)(CH_NT / 32) * CH_SLOTS * sizeof(unsigned long long);
    return o;
}

template <int KS>
__global__ void __launch_bounds__(CH_NT, 512 / CH_NT) k_chain_step(const ChainArgs a, int t)
{
    constexpr int NT8 = (KS + 1) / 2;                       // 8-wide output tiles
    extern __shared__ __align__(16) double shd[];
    const DevFT &ft = a.ft;
    const int d = ft.d, nmax = a.nmax, RS = a.rs;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    const int W = gridDim.x * (CH_NT / 32), w = blockIdx.x * (CH_NT / 32) + warp;
    const int mL = t, mR = d - 1 - t;
    int bar_off;
    chain_step_smem(RS, &bar_off);
    const int TB = 8 * RS;                                  // doubles of a tile
    double *ring = shd + warp * CH_SLOTS * TB;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(shd) + bar_off) + warp * CH_SLOTS;
    const int TL = __ldg(a.tst + ((size_t)mL * 2 + 0) * (nmax + 1) + a.P.ngrid[mL]);
    const int TR = __ldg(a.tst + ((size_t)mR * 2 + 1) * (nmax + 1) + a.P.ngrid[mR]);
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < CH_SLOTS; q++) ch_mbar_init(bar + q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // Programmatic dependent launch: steps t >= 1 are launched while step t-1 still runs (ft.cu); everything above
    // reads the plan only.  Let the next step start its own prologue, then wait for the previous step's rows.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (t == 0) {
        // records whose side takes no step at all keep the empty product e_1 in row 0: the left set of k = 0, the
        // right set of k = d-1 (rows that no step of this or a later launch touches)
        const long long total = (long long)a.F * RS, stride = (long long)gridDim.x * CH_NT;
        for (long long e = blockIdx.x * (long long)CH_NT + threadIdx.x; e < total; e += stride) {
            const int f = (int)(e / RS), q = (int)(e - (long long)f * RS);
            int k = a.dim_vary[f];
            k = k < 0 ? 0 : (k >= d ? d - 1 : k);
            if (k == 0) a.sets[(size_t)f * a.setw + q] = q == 0 ? 1.0 : 0.0;
            if (k == d - 1) a.sets[(size_t)f * a.setw + (size_t)(1 + 2 * k) * RS + q] = q == 0 ? 1.0 : 0.0;
        }
    }
    const long long T = (long long)TL + TR;
    int tile = (int)(T * w / W);
    const int tend = (int)(T * (w + 1) / W);
    if (tile >= tend) return;
    const double *xin = a.x[t & 1];
    double *xout = a.x[(t + 1) & 1];
    const int *rowd = a.rowd + (size_t)t * a.xrows + gid, *tileb = a.tileb + (size_t)t * (a.xrows / 8);
    const int2 *rowp = a.rowp + (size_t)t * a.xrows + gid;
    const unsigned tbytes = (unsigned)(TB * sizeof(double));

    // ring fill: the first CH_SLOTS tiles of this warp's range
    if (t > 0 && lane == 0) {
#pragma unroll
        for (int q = 0; q < CH_SLOTS; q++)
            if (tile + q < tend) ch_fetch(ring + q * TB, xin + (size_t)(tile + q) * TB, tbytes, bar + q);
    }
    double B[KS][NT8];
    int Bsj = -1;
    int D0 = __ldg(rowd + (size_t)tile * 8), J0 = __ldg(tileb + tile);
    unsigned phase = 0;                                     // bit q: parity of slot q's next completion
    for (int it = 0; tile < tend; tile++, it++) {
        int D1 = D0, J1 = J0;
        if (tile + 1 < tend) { D1 = __ldg(rowd + (size_t)(tile + 1) * 8); J1 = __ldg(tileb + tile + 1); }
        const int sj0 = J0;
        const bool prefix = D0 >= 0 && (D0 & CH_PREFIX);
        int2 P0 = make_int2(-1, -1);                        // a prefix row's two extra destinations: needed by its stores only
        if (prefix) P0 = __ldg(rowp + (size_t)tile * 8);
        if (sj0 != Bsj) {                                   // (the bucket is the same for the eight rows of a tile)
            Bsj = sj0;
            const int side = Bsj >> 16, j = Bsj & 0xffff, m = side ? mR : mL;
            const int rin = side ? ft.r[m + 1] : ft.r[m], rout = side ? ft.r[m] : ft.r[m + 1];
            // B fragments (row q = 4ks+tig of the contraction, column o = 8nt+gid of the output) of block G_m[j]
            // (element (a,b) at a + b*r_m): left  out[o=b] = sum_a in[a] G[a,b],  right  out[o=a] = sum_b G[a,b] in[b]
            const double *g = ft.base + ft.off[m] + (size_t)j * ft.r[m] * ft.r[m + 1];
            const int rm = ft.r[m];
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
#pragma unroll
                for (int nt = 0; nt < NT8; nt++) {
                    const int q = 4 * ks + tig, o = 8 * nt + gid;
                    B[ks][nt] = (q < rin && o < rout) ? __ldg(g + (side ? o + q * rm : q + o * rm)) : 0.0;
                }
        }
        // this tile's rows: A fragment (row gid, column 4ks+tig) from the ring slot
        double A[KS];
        const int slot = it % CH_SLOTS;
        if (t == 0) {
#pragma unroll
            for (int ks = 0; ks < KS; ks++) A[ks] = (4 * ks + tig == 0) ? 1.0 : 0.0;      // every vector of step 0 is e_1
        } else {
            ch_mbar_wait(bar + slot, (phase >> slot) & 1u);
            phase ^= 1u << slot;
            const double *row = ring + slot * TB + gid * RS + tig;
#pragma unroll
            for (int ks = 0; ks < KS; ks++) A[ks] = row[4 * ks];
            __syncwarp();                                   // every lane has read the slot: refill it with tile + CH_SLOTS
            if (lane == 0 && tile + CH_SLOTS < tend)
                ch_fetch(ring + slot * TB, xin + (size_t)(tile + CH_SLOTS) * TB, tbytes, bar + slot);
        }
        double acc[NT8][2];
#pragma unroll
        for (int nt = 0; nt < NT8; nt++) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ks++)
#pragma unroll
            for (int nt = 0; nt < NT8; nt++) ch_dmma(acc[nt][0], acc[nt][1], A[ks], B[ks][nt]);
        if (D0 != (int)0x80000000) {
            // launch t+1's buffer (plus the two neighbour slots for a prefix), or the fiber's record
            double *dst = (D0 >= 0 ? xout + (size_t)(D0 & (CH_PREFIX - 1)) * RS : a.sets + (size_t)(~D0) * RS) + 2 * tig;
#pragma unroll
            for (int nt = 0; nt < NT8; nt++)
                if (8 * nt + 2 * tig < RS) *reinterpret_cast<double2 *>(dst + 8 * nt) = make_double2(acc[nt][0], acc[nt][1]);
            if (prefix) {
                double *dlo = xout + (size_t)P0.x * RS + 2 * tig, *dhi = xout + (size_t)P0.y * RS + 2 * tig;
#pragma unroll
                for (int nt = 0; nt < NT8; nt++)
                    if (8 * nt + 2 * tig < RS) {
                        *reinterpret_cast<double2 *>(dlo + 8 * nt) = make_double2(acc[nt][0], acc[nt][1]);
                        *reinterpret_cast<double2 *>(dhi + 8 * nt) = make_double2(acc[nt][0], acc[nt][1]);
                    }
            }
        }
        D0 = D1; J0 = J1;
    }
}

#endif  // C3SC_FT_TYPES_ONLY

}  // namespace c3sc
