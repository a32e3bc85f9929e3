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
//   k_chain_plan   counting sort of the chunk's (fiber, role) entries by (side, dimension, block); roles: centre,
//                  lower neighbour, upper neighbour.  One CTA per (chunk, dimension).
//   k_chain_init   row 0 of both sets and the prefix buffers of every record := e_1.
//   k_chain_step   launch t = 0..d-2 advances the left sets through dimension t and the right sets through d-1-t.
//
// Record of fiber f (FtArgs::sets + f*setw, rows of RS = 4*KS doubles, zero beyond the rank):
//   rows [0, 1+2k)        left set,  row v = vector v          (k = dim_vary)
//   rows [1+2k, 2d)       right set
//   rows 2d .. 2d+3       P[side][parity]: the prefix / suffix alone, double-buffered -- the centre product
//                         rewrites a fiber's rows in place while the neighbour products (other buckets, other
//                         warps) still read its prefix of the previous step.
#pragma once
#include "ft_kernel.cuh"

namespace c3sc {

constexpr int CH_NT = 256;            // step kernel: 8 warps
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
    // plan of this chunk (device, written by k_chain_plan):
    int *kst;                 // [d][2][nmax*3 + 1]  start of (block j, role) inside the dimension's entry list
    int *tst;                 // [d][2][nmax + 1]    first 8-row tile of block j; [..][N] = number of tiles
    int *ent;                 // [d][entstride]      entries: fiber | k << 24
    int nmax, entstride;      // entstride = 3 * (fibers of a full chunk)
};

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

// grid (chunks, d), 1024 threads.  Chunk c covers fibers [c*FC, min(F, (c+1)*FC)) of the batch; its plan arrays and
// records are the c-th slices of the buffers in `a` (strides given).
struct ChainPlanStrides { long long kst, tst, ent; };
__global__ void __launch_bounds__(1024) k_chain_plan(ChainArgs a, int FC, ChainPlanStrides S)
{
    extern __shared__ int sh[];
    __shared__ int wsum[32];
    const int d = a.ft.d, m = blockIdx.y, tid = threadIdx.x, NT = blockDim.x;
    const int c0 = blockIdx.x * FC, Fc = (a.F - c0 < FC) ? a.F - c0 : FC;
    const int N = a.P.ngrid[m], nmax = a.nmax;
    const int *dv = a.dim_vary + c0, *fi = a.fixed_ind + (size_t)c0 * d;
    const int *nfi = a.nbr_fixed_in ? a.nbr_fixed_in + (size_t)c0 * 2 * (d - 1) : nullptr;
    int *kst = a.kst + blockIdx.x * S.kst + (size_t)m * 2 * (nmax * 3 + 1);
    int *tst = a.tst + blockIdx.x * S.tst + (size_t)m * 2 * (nmax + 1);
    int *ent = a.ent + blockIdx.x * S.ent + (size_t)m * a.entstride;
    int *cnt = sh, *fill = sh + 2 * 3 * N;                  // [side][j][role]; fill = running positions of the scatter
    const int nb = 2 * 3 * N;
    for (int e = tid; e < nb; e += NT) cnt[e] = 0;
    __syncthreads();
    auto roles = [&](int f, int &side, int &i0, int &lo, int &hi) -> int {      // returns k, or -1 when m is the varying dim
        int k = dv[f];
        k = k < 0 ? 0 : (k >= d ? d - 1 : k);
        if (k == m) return -1;
        side = m < k ? 0 : 1;
        i0 = ft_clamp_index(fi[(size_t)f * d + m], N);
        ft_fixed_pair(a.P, m, i0, lo, hi);
        if (nfi) {
            const int slot = m < k ? m : m - 1;
            lo = ft_clamp_index(nfi[(size_t)f * 2 * (d - 1) + 2 * slot], N);
            hi = ft_clamp_index(nfi[(size_t)f * 2 * (d - 1) + 2 * slot + 1], N);
        }
        return k;
    };
    for (int f = tid; f < Fc; f += NT) {
        int side, i0, lo, hi;
        if (roles(f, side, i0, lo, hi) < 0) continue;
        atomicAdd(&cnt[(side * N + i0) * 3], 1);
        atomicAdd(&cnt[(side * N + lo) * 3 + 1], 1);
        atomicAdd(&cnt[(side * N + hi) * 3 + 2], 1);
    }
    __syncthreads();
    // tiles per block: ceil((centre * nin + lo + hi) / 8), nin = 1 + 2s vectors before the step; the left sets
    // reach dimension m at step m, the right sets at step d-1-m
    for (int e = tid; e < 2 * N; e += NT) {
        const int side = e / N;
        const int nin = 1 + 2 * (side ? d - 1 - m : m);
        const int rows = cnt[e * 3] * nin + cnt[e * 3 + 1] + cnt[e * 3 + 2];
        fill[nb + e] = (rows + 7) >> 3;                     // tile counts behind the two histograms
    }
    __syncthreads();
    cta_exclusive_scan(cnt, nb, wsum);                      // both sides in one list: side 1 starts where side 0 ends
    for (int side = 0; side < 2; side++) {                  // tile prefix per side
        int *t = fill + nb + side * N;
        const int total = cta_exclusive_scan(t, N, wsum);
        for (int e = tid; e < N; e += NT) tst[side * (nmax + 1) + e] = t[e];
        if (tid == 0) tst[side * (nmax + 1) + N] = total;
    }
    for (int e = tid; e < nb; e += NT) { fill[e] = cnt[e]; kst[(e / (3 * N)) * (nmax * 3 + 1) + (e % (3 * N))] = cnt[e]; }
    if (tid == 0) kst[3 * N] = cnt[3 * N];                  // end of side 0 = start of side 1
    __syncthreads();
    for (int f = tid; f < Fc; f += NT) {
        int side, i0, lo, hi;
        const int k = roles(f, side, i0, lo, hi);
        if (k < 0) continue;
        const int tag = f | (k << 24);
        ent[atomicAdd(&fill[(side * N + i0) * 3], 1)] = tag;
        ent[atomicAdd(&fill[(side * N + lo) * 3 + 1], 1)] = tag;
        ent[atomicAdd(&fill[(side * N + hi) * 3 + 2], 1)] = tag;
    }
    __syncthreads();
    if (tid == 0) kst[(nmax * 3 + 1) + 3 * N] = fill[(2 * N - 1) * 3 + 2];       // end of side 1's last bucket
}

// Records of a chunk before its first step: row 0 of both sets and both P[.][0] buffers start as e_1 (the empty
// product).  Launched per chunk on the chunk's stream (the record buffer is reused by the lane's next chunk).
__global__ void __launch_bounds__(256) k_chain_init(const ChainArgs a)
{
    const int d = a.ft.d, RS = a.rs;
    const long long total = (long long)a.F * 4 * RS, step = (long long)gridDim.x * blockDim.x;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += step) {
        const int f = (int)(e / (4 * RS)), rem = (int)(e - (long long)f * 4 * RS), w = rem / RS, q = rem - w * RS;
        int k = a.dim_vary[f];
        k = k < 0 ? 0 : (k >= d ? d - 1 : k);
        const int row = w == 0 ? 0 : (w == 1 ? 1 + 2 * k : (w == 2 ? 2 * d : 2 * d + 2));   // L row 0, R row 0, P[0][0], P[1][0]
        a.sets[(size_t)f * a.setw + (size_t)row * RS + q] = q == 0 ? 1.0 : 0.0;
    }
}

__device__ __forceinline__ void ch_dmma(double &d0, double &d1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// One launch = one step of both sides.  Warps own contiguous ranges of 8-row tiles of the concatenated tile list
// [left buckets of dimension t | right buckets of dimension d-1-t].
template <int KS>
__global__ void __launch_bounds__(CH_NT, 2) k_chain_step(const ChainArgs a, int t)
{
    constexpr int NT8 = (KS + 1) / 2;                       // 8-wide output tiles
    const DevFT &ft = a.ft;
    const int d = ft.d, nmax = a.nmax, RS = a.rs;
    const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
    const int W = gridDim.x * (CH_NT / 32), w = blockIdx.x * (CH_NT / 32) + (threadIdx.x >> 5);
    const int mS[2] = {t, d - 1 - t};
    const int nin = 1 + 2 * t, par = t & 1;
    const int TL = a.tst[((size_t)mS[0] * 2 + 0) * (nmax + 1) + a.P.ngrid[mS[0]]];
    const int TR = a.tst[((size_t)mS[1] * 2 + 1) * (nmax + 1) + a.P.ngrid[mS[1]]];
    const long long T = (long long)TL + TR;
    int tile = (int)(T * w / W);
    const int tend = (int)(T * (w + 1) / W);
    if (tile >= tend) return;
    const unsigned magic = (unsigned)((0x100000000ULL + nin - 1) / nin);          // r / nin for r < 2^32 / nin

    int side = -1, j = 0, N = 0, m = 0, rin = 0, rout = 0;
    const int *tst = nullptr, *kst = nullptr, *ent = nullptr;
    int jt0 = 0, jt1 = 0;                                   // tile range of the current bucket (side-relative)
    int sc = 0, slo = 0, shi = 0, send = 0;                 // entry ranges of the bucket: centre, lo, hi
    double B[KS][NT8];
    for (; tile < tend; tile++) {
        const int sd = tile < TL ? 0 : 1;
        const int tl = tile - (sd ? TL : 0);
        bool newb = false;
        if (sd != side) {
            side = sd; m = mS[side]; N = a.P.ngrid[m];
            tst = a.tst + ((size_t)m * 2 + side) * (nmax + 1);
            kst = a.kst + ((size_t)m * 2 + side) * (nmax * 3 + 1);
            ent = a.ent + (size_t)m * a.entstride;
            rin = side ? ft.r[m + 1] : ft.r[m];
            rout = side ? ft.r[m] : ft.r[m + 1];
            int lo = 0, hi = N;                             // largest j with tst[j] <= tl
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (__ldg(tst + mid) <= tl) lo = mid; else hi = mid; }
            j = lo; newb = true;
        }
        if (!newb && tl >= jt1) { j++; newb = true; }
        if (newb) {
            jt0 = __ldg(tst + j); jt1 = __ldg(tst + j + 1);
            while (tl >= jt1) { j++; jt0 = jt1; jt1 = __ldg(tst + j + 1); }       // empty buckets
            sc = __ldg(kst + 3 * j); slo = __ldg(kst + 3 * j + 1); shi = __ldg(kst + 3 * j + 2); send = __ldg(kst + 3 * j + 3);
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
        // this lane's row of the tile
        const int r = (tl - jt0) * 8 + gid;
        const int nc = (slo - sc) * nin, rows = nc + (send - slo);
        const bool valid = r < rows;
        int e, v;                                           // entry position, vector index (centre) / new row (neighbours)
        if (r < nc) { const int q = (int)__umulhi((unsigned)r, magic); e = sc + q; v = r - q * nin; }
        else { e = slo + (r - nc); v = e < shi ? nin : nin + 1; }
        const int tag = valid ? __ldg(ent + e) : 0;
        const int f = tag & 0xffffff, k = tag >> 24;
        double *rec = a.sets + (size_t)f * a.setw;
        double *set = rec + (size_t)(side ? 1 + 2 * k : 0) * RS;
        double *Pc = rec + (size_t)(2 * d + 2 * side + par) * RS, *Pn = rec + (size_t)(2 * d + 2 * side + (par ^ 1)) * RS;
        const bool centre = r < nc;
        const double *src = (centre && v > 0) ? set + (size_t)v * RS : Pc;
        double *dst = (centre && v == 0) ? Pn : set + (size_t)v * RS;
        double A[KS];
#pragma unroll
        for (int ks = 0; ks < KS; ks++) A[ks] = valid ? src[4 * ks + tig] : 0.0;
        double acc[NT8][2];
#pragma unroll
        for (int nt = 0; nt < NT8; nt++) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ks++)
#pragma unroll
            for (int nt = 0; nt < NT8; nt++) ch_dmma(acc[nt][0], acc[nt][1], A[ks], B[ks][nt]);
        if (valid) {
#pragma unroll
            for (int nt = 0; nt < NT8; nt++) {
                const int o = 8 * nt + 2 * tig;                 // D: row r, columns o, o+1; zero beyond the rank
                if (o < RS) {
                    const double2 val = make_double2(acc[nt][0], acc[nt][1]);
                    *reinterpret_cast<double2 *>(dst + o) = val;
                    if (centre && v == 0) *reinterpret_cast<double2 *>(set + o) = val;      // the prefix also lives in row 0
                }
            }
        }
    }
}

#endif  // C3SC_FT_TYPES_ONLY

}  // namespace c3sc
