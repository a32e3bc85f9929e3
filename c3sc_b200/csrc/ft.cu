// ft.cu -- host launchers of the model-independent stage-1 kernels (ft_kernel.cuh)
#include <cstdlib>
#include "ft_mma_kernel.cuh"
#include "policy_kernel.cuh"

namespace c3sc {

static int g_sms = 0, g_max_optin = 0, g_max_sm = 0;
static void ft_device_info()
{
    if (g_sms) return;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g_max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&g_max_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
}

// The kernels that are compiled per rank geometry KS = ceil(r_max / 4) (chain steps, per-fiber chains, node kernel) live in
// ft_ks.cu, one translation unit per KS (built in parallel): their launchers, by name.
#define C3SC_KS_DECL(n) \
    int launch_chain_steps_ks##n(const ChainArgs &a, cudaStream_t st); \
    int launch_mma_ks##n(const FtArgs &a, const CtlArgs *fused, int nsplit, cudaStream_t st);
C3SC_KS_DECL(1) C3SC_KS_DECL(2) C3SC_KS_DECL(3) C3SC_KS_DECL(4) C3SC_KS_DECL(5) C3SC_KS_DECL(6) C3SC_KS_DECL(7) C3SC_KS_DECL(8)
#undef C3SC_KS_DECL

int ft_sm_count() { ft_device_info(); return g_sms; }
int ft_max_optin_smem() { ft_device_info(); return g_max_optin; }

// perm / kcount / kstart / cleared active counters of ALL chunks of a batch.  1 launch.
int launch_group_fibers(const DevProblem &P, int F, const ChunkLayout &lay, const int *dim_vary, const int *fixed_ind, int *perm,
                        int *cnt_all, cudaStream_t st)
{
    if (F <= 0) return 0;
    GridDims ng;
    for (int i = 0; i < MAXD; i++) ng.n[i] = P.ngrid[i];
    const int FC = lay.FC;
    const int vy = (FC * P.dx + 8191) / 8192 > 16 ? 16 : (FC * P.dx + 8191) / 8192;     // CTAs per chunk for the descriptor check
    const unsigned nord = (unsigned)((F + lay.FS - 1) / lay.FS) * (unsigned)lay.m;      // chunk slots of all super-chunks
    k_group_fibers<<<dim3(nord, (unsigned)(vy < 1 ? 1 : vy)), 1024, 0, st>>>(F, lay, P.dx, dim_vary, fixed_ind, ng, P.err, perm, cnt_all);
    return (int)cudaGetLastError();
}

int launch_pack_cores(const DevFT &ft, double *baseT, double *baseP, double *baseQ, cudaStream_t st)
{
    long long most = 0;
    for (int k = 0; k < ft.d; k++) {
        long long len = (long long)ft.n[k] * ft.r[k] * ft.r[k + 1];
        if (baseP) { const long long pl = (long long)ft.n[k] * ft.ldp[k] * ft.cpp[k]; len = pl > len ? pl : len; }
        most = len > most ? len : most;
    }
    if (most <= 0) return 0;
    long long gx = (most + 255) / 256;
    if (gx > 1024) gx = 1024;
    k_pack_cores<<<dim3((unsigned)gx, (unsigned)ft.d), 256, 0, st>>>(ft, baseT, baseP, baseQ);
    return (int)cudaGetLastError();
}

// padded-tile geometry for the tensor-core kernels, the same for every core: 8*MT columns of 8*MT+4
// rows (leading dimension = 4 mod 8: fragment reads conflict-free in both orientations, 16-byte
// columns for TMA).  Fills ft.ldp / cpp / offP, returns the doubles needed.
long long ft_padded_layout(DevFT &ft)
{
    long long total = 0;
    int rmax = 1;
    for (int k = 0; k <= ft.d; k++) rmax = ft.r[k] > rmax ? ft.r[k] : rmax;
    const int m8 = 8 * (((rmax + 3) / 4 + 1) / 2);          // 8*MT, MT = ceil(KS/2), KS = ceil(rmax/4): one geometry for all cores
    for (int k = 0; k < ft.d; k++) {
        ft.ldp[k] = m8 + 4;
        ft.cpp[k] = m8;
        ft.offP[k] = total;
        total += (long long)ft.n[k] * ft.ldp[k] * ft.cpp[k];
    }
    return total;
}

// compact tile copy (baseQ): rows padded to even only; + one zero block of slack at the end
long long ft_compact_layout(DevFT &ft)
{
    long long total = 0;
    for (int k = 0; k < ft.d; k++) {
        ft.ldq[k] = (ft.r[k] + 1) & ~1;
        ft.offQ[k] = total;
        total += (long long)ft.n[k] * ft.ldq[k] * ft.r[k + 1];
        total = (total + 1) & ~1LL;
    }
    return total + 2;
}

// ranks <= 32: chains by warp tasks + DMMA node kernel; larger ranks (or C3SC_FT_GENERAL=1): k_ft_costs
int ft_uses_mma(const DevFT &ft)
{
    static const bool general = getenv("C3SC_FT_GENERAL") != nullptr;
    if (general) return 0;
    for (int i = 0; i <= ft.d; i++)
        if (ft.r[i] > 32) return 0;
    return 1;
}
static int ft_rmax(const DevFT &ft)
{
    int rmax = 1;
    for (int i = 0; i <= ft.d; i++) rmax = ft.r[i] > rmax ? ft.r[i] : rmax;
    return rmax;
}
size_t ft_sets_bytes(const DevFT &ft, size_t F) { return (size_t)ft_rec_width(ft.d, ft_rmax(ft)) * F * sizeof(double); }
void ft_record_geometry(const DevFT &ft, int *setw, int *rs) { *setw = ft_rec_width(ft.d, ft_rmax(ft)); *rs = ft_rec_rs(ft_rmax(ft)); }

// ---- bucketed chain stage (chain_kernel.cuh) ------------------------------------------------------------------
// ints of plan storage per chunk of FC fibers
void chain_plan_sizes(const DevFT &ft, int nmax, size_t FC, size_t *kst, size_t *tst, size_t *ent)
{
    *kst = (size_t)ft.d * 2 * ((size_t)nmax * 3 + 1);
    *tst = (size_t)ft.d * 2 * ((size_t)nmax + 1);
    *ent = (size_t)ft.d * 3 * FC;                           // the inverse rows take another ent ints per chunk
}
int chain_bucketed_ok(const DevFT &ft, int nmax) { return ft.d >= 2 && nmax <= CH_NMAX && ft_uses_mma(ft); }

// plan of every chunk of a batch in one launch: a.F = fibers of the batch, FC = fibers per chunk; a.kst / tst / ent /
// sets point at chunk 0's slices, consecutive chunks follow at the strides of chain_plan_sizes / FC * setw
int launch_chain_plan(const ChainArgs &a, int FC, int *cntg, cudaStream_t st)
{
    if (a.F <= 0) return 0;
    size_t k, t, e;
    chain_plan_sizes(a.ft, a.nmax, (size_t)FC, &k, &t, &e);
    ChainPlanStrides S = {(long long)k, (long long)t, (long long)e, (long long)e, (long long)(a.ft.d - 1) * a.xrows};
    const unsigned nch = (unsigned)((a.F + FC - 1) / FC);
    cudaError_t err = cudaMemsetAsync(cntg, 0, (size_t)nch * k * sizeof(int), st);
    if (err != cudaSuccess) return (int)err;
    const int fc = a.F < FC ? a.F : FC;
    const dim3 gs((unsigned)((fc + CHP_SLICE - 1) / CHP_SLICE), (unsigned)a.ft.d, nch);
    k_chain_count<<<gs, CHP_NT, (size_t)6 * a.nmax * sizeof(int), st>>>(a, FC, S, cntg);
    k_chain_scan<<<dim3(nch, (unsigned)a.ft.d), 1024, (size_t)8 * a.nmax * sizeof(int), st>>>(a, S, cntg);
    k_chain_scatter<<<gs, CHP_NT, (size_t)12 * a.nmax * sizeof(int), st>>>(a, FC, S, cntg);
    // ... and where every row's product goes in the next dimension's bucket order (needs the inverse rows of ALL dimensions;
    // every bucket of a dimension is a few hundred rows)
    unsigned gz = (unsigned)(2 * a.nmax);                   // one CTA per bucket: a bucket is a chain of dependent loads
    k_chain_link<<<dim3(nch, (unsigned)a.ft.d, gz), 256, 0, st>>>(a, S);
    return (int)cudaGetLastError();
}
constexpr int CHAIN_PLAN_LAUNCHES = 4;

// the d-1 steps of one chunk (a.F = fibers of the chunk, pointers at the chunk's slices); returns the launches via *n
int launch_chain_steps(const ChainArgs &a, cudaStream_t st, int *n)
{
    *n = a.ft.d - 1;
    if (getenv("C3SC_DBG_SKIP_CHAIN")) { *n = 0; return 0; }     // timing experiments only: the records keep stale (finite) data
    switch ((ft_rmax(a.ft) + 3) / 4) {
    case 1: return launch_chain_steps_ks1(a, st);
    case 2: return launch_chain_steps_ks2(a, st);
    case 3: return launch_chain_steps_ks3(a, st);
    case 4: return launch_chain_steps_ks4(a, st);
    case 5: return launch_chain_steps_ks5(a, st);
    case 6: return launch_chain_steps_ks6(a, st);
    case 7: return launch_chain_steps_ks7(a, st);
    default: return launch_chain_steps_ks8(a, st);
    }
}

// CTAs per group of the node kernel: small batches split every group over node-tile ranges until the machine is covered
static int ft_nodes_nsplit_of(const FtArgs &a)
{
    ft_device_info();
    const int grid = (a.F + FT_FBMAX - 1) / FT_FBMAX + a.ft.d;
    int ntl = 1;
    for (int k = 0; k < a.ft.d; k++) { const int t = (a.P.ngrid[k] + FTN_T - 1) / FTN_T; ntl = t > ntl ? t : ntl; }
    int ns = 1;
    while (ns < ntl && grid * ns * 2 <= 2 * g_sms) ns *= 2;      // stay within one wave of 2 CTAs/SM
    return ns > ntl ? ntl : ns;
}
int ft_nodes_nsplit(const FtArgs &a) { return ft_nodes_nsplit_of(a); }
// doubles of one region of the fused ring, and the ring size (more regions than CTAs can be resident)
long long ft_region_doubles(const DevProblem &P) { return (long long)(2 * P.dx + 1) * FT_FBMAX * ft_even_up(P.nmax); }
int ft_ring_regions() { ft_device_info(); return 2 * g_sms + 32; }

static int launch_ft_mma(FtArgs a, const CtlArgs *fused, cudaStream_t st)
{
    // group size: 8 fibers fill the DMMA tile (a tile costs the same for 1 fiber as for 8); small
    // batches cover the SMs by splitting every group over node-tile ranges instead (nsplit)
    a.FB = FT_FBMAX;
    int rmax = 1;
    for (int i = 0; i <= a.ft.d; i++) rmax = a.ft.r[i] > rmax ? a.ft.r[i] : rmax;
    const int nsplit = ft_nodes_nsplit_of(a);
    switch ((rmax + 3) / 4) {                                // KS
    case 1: return launch_mma_ks1(a, fused, nsplit, st);
    case 2: return launch_mma_ks2(a, fused, nsplit, st);
    case 3: return launch_mma_ks3(a, fused, nsplit, st);
    case 4: return launch_mma_ks4(a, fused, nsplit, st);
    case 5: return launch_mma_ks5(a, fused, nsplit, st);
    case 6: return launch_mma_ks6(a, fused, nsplit, st);
    case 7: return launch_mma_ks7(a, fused, nsplit, st);
    default: return launch_mma_ks8(a, fused, nsplit, st);
    }
}

static int launch_ft_stage(const FtArgs &a_in, const CtlArgs *fused, cudaStream_t st);

// Stage 1 over one chunk.  a.FB == 0: pick the group size here.  fused != NULL: the node kernel also runs stage 2.
int launch_ft_costs(const FtArgs &a_in, const CtlArgs *fused, cudaStream_t st)
{
    int rc = launch_ft_stage(a_in, fused, st);
    if (rc || !a_in.costs || a_in.F <= 0) return rc;
    long long g = (a_in.NS + 255) / 256;
    if (g > 4096) g = 4096;
    k_costs_along<<<(unsigned)g, 256, 0, st>>>(a_in);
    return (int)cudaGetLastError();
}

static int launch_ft_stage(const FtArgs &a_in, const CtlArgs *fused, cudaStream_t st)
{
    ft_device_info();
    FtArgs a = a_in;
    if (a.F <= 0) return 0;
    if (a.sets && ft_uses_mma(a.ft)) return launch_ft_mma(a, fused, st);
    if (fused) return (int)cudaErrorInvalidValue;               // only the tensor-core node kernel fuses stage 2
    // two CTAs per SM when the carve-up allows it
    const size_t budget = (size_t)g_max_sm / 2 - 1024;
    if (a.FB <= 0) a.FB = ft_pick_fb(a.ft, a.P.nmax, a.F, g_sms, budget);
    size_t smem = FtPlan(a.ft, a.P.nmax, a.FB).bytes();
    while (smem > (size_t)g_max_optin && a.FB > 1) { a.FB >>= 1; smem = FtPlan(a.ft, a.P.nmax, a.FB).bytes(); }
    if (smem > (size_t)g_max_optin) return (int)cudaErrorInvalidValue;
    static size_t attr_dev[C3SC_MAXDEV] = {0};
    size_t &attr = attr_dev[c3sc_cur_dev()];
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(k_ft_costs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr = smem;
    }
    const int grid = (a.F + a.FB - 1) / a.FB + a.ft.d;       // upper bound on the number of groups
    k_ft_costs<<<grid, FT_NT, smem, st>>>(a);
    return (int)cudaGetLastError();
}

// policy rows between a batch-contiguous buffer and the per-fiber row store: fiber f <-> slot idx[f], `per` doubles each
__global__ void k_rows_move(double *dst, const double *src, const int *idx, int F, long long per, int scatter)
{
    const long long total = (long long)F * per, step = (long long)gridDim.x * blockDim.x;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += step) {
        const long long f = e / per, q = e - f * per;
        const long long s = (long long)idx[f] * per + q;
        if (scatter) dst[s] = src[e]; else dst[e] = src[s];
    }
}
int launch_rows_move(double *dst, const double *src, const int *idx, int F, long long per, int scatter, cudaStream_t st)
{
    if (F <= 0) return 0;
    ft_device_info();
    long long g = ((long long)F * per + 255) / 256;
    if (g > g_sms * 8) g = g_sms * 8;
    k_rows_move<<<(unsigned)g, 256, 0, st>>>(dst, src, idx, F, per, scatter);
    return (int)cudaGetLastError();
}

// One chunk's values into every peer's gathered buffer with 16-byte stores from a few CTAs (c3sc_batch_out::peer_mode 2):
// runs on the library's peer stream next to the following chunk's kernels and leaves the control kernel's own stores local.
struct PeerList { double *p[C3SC_MAXPEERS]; int n; };
__global__ void __launch_bounds__(256) k_peer_scatter(const double *src, long long n, PeerList peers, long long off)
{
    const long long n2 = n >> 1, step = (long long)gridDim.x * blockDim.x;
    const double2 *s2 = reinterpret_cast<const double2 *>(src);
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n2; e += step) {
        const double2 v = s2[e];
#pragma unroll 1
        for (int g = 0; g < peers.n; g++) reinterpret_cast<double2 *>(peers.p[g] + off)[e] = v;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0)
        for (int g = 0; g < peers.n; g++) peers.p[g][off + n - 1] = src[n - 1];
}
int launch_peer_scatter(const double *src, long long n, double *const *peers, int npeer, long long off, cudaStream_t st)
{
    if (n <= 0 || npeer <= 0) return 0;
    PeerList pl;
    pl.n = 0;
    for (int g = 0; g < npeer && g < C3SC_MAXPEERS; g++)
        if (peers[g] + off != src) pl.p[pl.n++] = peers[g];         // the rank's own slot when it is the output itself
    if (pl.n == 0) return 0;
    if (((size_t)src & 15) || (off & 1)) return (int)cudaErrorMisalignedAddress;
    k_peer_scatter<<<48, 256, 0, st>>>(src, n, pl, off);
    return (int)cudaGetLastError();
}

// valuef_eval at npts points (device arrays): piecewise-linear interpolation of the nodal cores
int launch_ft_eval_points(const DevProblem &P, const DevFT &ft, int npts, const double *pts, double *out, cudaStream_t st)
{
    if (npts <= 0) return 0;
    int rs = 1;
    for (int i = 0; i <= ft.d; i++) rs = ft.r[i] > rs ? ft.r[i] : rs;
    const size_t smem = (size_t)8 * 2 * rs * sizeof(double);
    int grid = (npts + 7) / 8;
    ft_device_info();
    if (grid > g_sms * 8) grid = g_sms * 8;
    k_ft_eval_points<<<grid, 256, smem, st>>>(P, ft, npts, pts, out);
    return (int)cudaGetLastError();
}

// flags + the 2d+1 evaluation points of n states (mca_get_neighbor_node_costs)
int launch_policy_points(const DevProblem &P, int n, const double *x, double *pts, int *absorbed, cudaStream_t st)
{
    if (n <= 0) return 0;
    k_policy_points<<<(n + 127) / 128, 128, 0, st>>>(P, n, x, pts, absorbed);
    return (int)cudaGetLastError();
}

}  // namespace c3sc
