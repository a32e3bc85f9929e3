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
#ifndef C3SC_FT_KS_UNIT        // compiled once, in ft.cu (ft_ks.cu holds the per-rank-geometry templates)
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
    // four fibers per thread and trip: the descriptor loads (a dependent pair per fiber, strided by d) are what this
    // kernel waits for
    constexpr int U = 4;
    for (int f0 = tid; f0 < Fc; f0 += U * NT) {
        int kk[U], side[U], i0[U], lo[U], hi[U];
#pragma unroll
        for (int u = 0; u < U; u++) kk[u] = (f0 + u * NT < Fc) ? roles(f0 + u * NT, side[u], i0[u], lo[u], hi[u]) : -1;
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (kk[u] < 0) continue;
            atomicAdd(&cnt[(side[u] * N + i0[u]) * 3], 1);
            atomicAdd(&cnt[(side[u] * N + lo[u]) * 3 + 1], 1);
            atomicAdd(&cnt[(side[u] * N + hi[u]) * 3 + 2], 1);
        }
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
    for (int f0 = tid; f0 < Fc; f0 += U * NT) {
        int kk[U], side[U], i0[U], lo[U], hi[U];
#pragma unroll
        for (int u = 0; u < U; u++) kk[u] = (f0 + u * NT < Fc) ? roles(f0 + u * NT, side[u], i0[u], lo[u], hi[u]) : -1;
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (kk[u] < 0) continue;
            const int tag = (f0 + u * NT) | (kk[u] << 24);
            ent[atomicAdd(&fill[(side[u] * N + i0[u]) * 3], 1)] = tag;
            ent[atomicAdd(&fill[(side[u] * N + lo[u]) * 3 + 1], 1)] = tag;
            ent[atomicAdd(&fill[(side[u] * N + hi[u]) * 3 + 2], 1)] = tag;
        }
    }
    __syncthreads();
    if (tid == 0) kst[(nmax * 3 + 1) + 3 * N] = fill[(2 * N - 1) * 3 + 2];       // end of side 1's last bucket
}
#endif

__device__ __forceinline__ void ch_dmma(double &d0, double &d1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// One launch = one step of both sides.  Warps own contiguous ranges of 8-row tiles of the concatenated tile list
// [left buckets of dimension t | right buckets of dimension d-1-t].  The bucket tables of the two dimensions are
// staged in shared memory once per CTA (one coalesced round trip instead of a dependent chain of L2 loads per warp);
// a warp's loop is software-pipelined: the entry of tile i+2 and the rows of tile i+1 are in flight while the
// products of tile i run.
struct ChainDec { int e, v, sj; unsigned flags; };          // entry position, vector / new row, side*65536 + bucket; flags: 1 valid, 2 centre

template <int KS>
__global__ void __launch_bounds__(CH_NT, 2) k_chain_step(const ChainArgs a, int t)
{
    constexpr int NT8 = (KS + 1) / 2;                       // 8-wide output tiles
    extern __shared__ int shs[];
    const DevFT &ft = a.ft;
    const int d = ft.d, nmax = a.nmax, RS = a.rs;
    const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
    const int W = gridDim.x * (CH_NT / 32), w = blockIdx.x * (CH_NT / 32) + (threadIdx.x >> 5);
    const int mL = t, mR = d - 1 - t;
    const int NL = a.P.ngrid[mL], NR = a.P.ngrid[mR];
    const int nin = 1 + 2 * t, par = t & 1;
    // shared: tst of both sides (N+1 each), kst of both sides (3N+1 each)
    int *sT[2], *sK[2];
    sT[0] = shs; sT[1] = sT[0] + NL + 1; sK[0] = sT[1] + NR + 1; sK[1] = sK[0] + 3 * NL + 1;
    {
        const int *gT0 = a.tst + ((size_t)mL * 2 + 0) * (nmax + 1), *gT1 = a.tst + ((size_t)mR * 2 + 1) * (nmax + 1);
        const int *gK0 = a.kst + ((size_t)mL * 2 + 0) * (nmax * 3 + 1), *gK1 = a.kst + ((size_t)mR * 2 + 1) * (nmax * 3 + 1);
        for (int e = threadIdx.x; e <= NL; e += CH_NT) sT[0][e] = __ldg(gT0 + e);
        for (int e = threadIdx.x; e <= NR; e += CH_NT) sT[1][e] = __ldg(gT1 + e);
        for (int e = threadIdx.x; e <= 3 * NL; e += CH_NT) sK[0][e] = __ldg(gK0 + e);
        for (int e = threadIdx.x; e <= 3 * NR; e += CH_NT) sK[1][e] = __ldg(gK1 + e);
    }
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
    __syncthreads();
    const int TL = sT[0][NL], TR = sT[1][NR];
    const long long T = (long long)TL + TR;
    int tile = (int)(T * w / W);
    const int tend = (int)(T * (w + 1) / W);
    if (tile >= tend) return;
    const unsigned magic = (unsigned)((0x100000000ULL + nin - 1) / nin);          // r / nin for r < 2^32 / nin (nin > 1)

    // cursor of the decoder: the bucket that holds the tile most recently decoded
    int cside = -1, cj = 0, cjt0 = 0, cjt1 = 0, csc = 0, cslo = 0, cshi = 0, csend = 0;
    auto decode = [&](int tl_abs) -> ChainDec {
        const int sd = tl_abs < TL ? 0 : 1;
        const int tl = tl_abs - (sd ? TL : 0);
        bool newb = false;
        if (sd != cside) {
            cside = sd;
            const int N = sd ? NR : NL;
            int lo = 0, hi = N;                             // largest j with tst[j] <= tl
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sT[sd][mid] <= tl) lo = mid; else hi = mid; }
            cj = lo; newb = true;
        } else if (tl >= cjt1) { cj++; newb = true; }
        if (newb) {
            cjt0 = sT[sd][cj]; cjt1 = sT[sd][cj + 1];
            while (tl >= cjt1) { cj++; cjt0 = cjt1; cjt1 = sT[sd][cj + 1]; }      // empty buckets
            csc = sK[sd][3 * cj]; cslo = sK[sd][3 * cj + 1]; cshi = sK[sd][3 * cj + 2]; csend = sK[sd][3 * cj + 3];
        }
        const int r = (tl - cjt0) * 8 + gid;
        const int nc = (cslo - csc) * nin, rows = nc + (csend - cslo);
        ChainDec D;
        D.flags = (r < rows ? 1u : 0u) | (r < nc ? 2u : 0u);
        if (r < nc) { const int q = nin == 1 ? r : (int)__umulhi((unsigned)r, magic); D.e = csc + q; D.v = r - q * nin; }
        else { D.e = cslo + (r - nc); D.v = D.e < cshi ? nin : nin + 1; }
        D.sj = sd * 65536 + cj;
        return D;
    };
    const int *entS[2] = {a.ent + (size_t)mL * a.entstride, a.ent + (size_t)mR * a.entstride};
    auto load_tag = [&](const ChainDec &D) -> int { return (D.flags & 1u) ? __ldg(entS[D.sj >> 16] + D.e) : 0; };
    // pointers of a row: src (read), dst (written), and for the prefix row also row 0 of the set
    auto row_ptrs = [&](const ChainDec &D, int tag, const double *&src, double *&dst, double *&dst0) {
        const int f = tag & 0xffffff, k = tag >> 24, side = D.sj >> 16;
        double *rec = a.sets + (size_t)f * a.setw;
        double *set = rec + (size_t)(side ? 1 + 2 * k : 0) * RS;
        const bool centre = (D.flags & 2u) != 0;
        src = (centre && D.v > 0) ? set + (size_t)D.v * RS : rec + (size_t)(2 * d + 2 * side + par) * RS;
        dst = (centre && D.v == 0) ? rec + (size_t)(2 * d + 2 * side + (par ^ 1)) * RS : set + (size_t)D.v * RS;
        dst0 = (centre && D.v == 0) ? set : nullptr;
    };
    auto load_rows = [&](const ChainDec &D, const double *src, double (&A)[KS]) {
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
            if (t == 0) A[ks] = (4 * ks + tig == 0 && (D.flags & 1u)) ? 1.0 : 0.0;       // every vector of step 0 is e_1
            else A[ks] = (D.flags & 1u) ? src[4 * ks + tig] : 0.0;
        }
    };

    double B[KS][NT8];
    int Bsj = -1;
    // pipeline fill: tile i decoded + tag + rows; tile i+1 decoded + tag
    ChainDec D0 = decode(tile), D1;
    int tag0 = load_tag(D0), tag1 = 0;
    const double *src0; double *dst0, *dsz0;
    row_ptrs(D0, tag0, src0, dst0, dsz0);
    double A0[KS];
    load_rows(D0, src0, A0);
    bool have1 = tile + 1 < tend;
    if (have1) { D1 = decode(tile + 1); tag1 = load_tag(D1); }
    for (; tile < tend; tile++) {
        // next tile: pointers from its tag, rows in flight; the tile after: decode + tag in flight
        const double *src1 = nullptr; double *dst1 = nullptr, *dsz1 = nullptr;
        double A1[KS];
        ChainDec D2;
        int tag2 = 0;
        bool have2 = false;
        if (have1) {
            row_ptrs(D1, tag1, src1, dst1, dsz1);
            load_rows(D1, src1, A1);
            have2 = tile + 2 < tend;
            if (have2) { D2 = decode(tile + 2); tag2 = load_tag(D2); }
        }
        // current tile
        if (D0.sj != Bsj) {
            Bsj = D0.sj;
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
        double acc[NT8][2];
#pragma unroll
        for (int nt = 0; nt < NT8; nt++) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ks++)
#pragma unroll
            for (int nt = 0; nt < NT8; nt++) ch_dmma(acc[nt][0], acc[nt][1], A0[ks], B[ks][nt]);
        if (D0.flags & 1u) {
#pragma unroll
            for (int nt = 0; nt < NT8; nt++) {
                const int o = 8 * nt + 2 * tig;                 // D: row r, columns o, o+1; zero beyond the rank
                if (o < RS) {
                    const double2 val = make_double2(acc[nt][0], acc[nt][1]);
                    *reinterpret_cast<double2 *>(dst0 + o) = val;
                    if (dsz0) *reinterpret_cast<double2 *>(dsz0 + o) = val;                   // the prefix also lives in row 0
                }
            }
        }
        // shift the pipeline
        D0 = D1; tag0 = tag1; src0 = src1; dst0 = dst1; dsz0 = dsz1;
#pragma unroll
        for (int ks = 0; ks < KS; ks++) A0[ks] = A1[ks];
        D1 = D2; tag1 = tag2; have1 = have2;
    }
}

#endif  // C3SC_FT_TYPES_ONLY

}  // namespace c3sc
