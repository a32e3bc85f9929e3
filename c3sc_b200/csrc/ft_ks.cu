// ft_ks.cu -- the stage-1 kernels that are instantiated on the rank geometry KS = ceil(r_max / 4) and their launchers:
// k_chain_step<KS> (chain_kernel.cuh), k_ft_chains<KS>, k_ft_nodes<KS>, k_ft_nodes_fused<KS> (ft_mma_kernel.cuh).
// Compiled once per KS (-DC3SC_KS=1..8), so the eight geometries build in parallel; ft.cu dispatches by name.
#include <cstdlib>
#define C3SC_FT_KS_UNIT
#include "ft_mma_kernel.cuh"

#ifndef C3SC_KS
#error "compile with -DC3SC_KS=1..8"
#endif

namespace c3sc {

int ft_sm_count();
int ft_max_optin_smem();

template <int KS>
static int launch_chain_steps_t(const ChainArgs &a, cudaStream_t st)
{
    const bool g_no_pdl = getenv("C3SC_NO_PDL") != nullptr;
    // CTAs of a step: the kernel is latency-bound (dependent L2 round trips, a handful of tiles per warp), and while it
    // holds an SM's registers the other lane's node kernel cannot use that SM
    int grid = ft_sm_count() * (1536 / CH_NT) / 2;         // six 128-thread CTAs per SM
    { const char *e = getenv("C3SC_CHAIN_GRID"); if (e && atoi(e) > 0) grid = atoi(e); }
    size_t most = 0;
    for (int t = 0; t + 1 < a.ft.d; t++) {
        const size_t sm = chain_step_smem(a.rs);
        most = sm > most ? sm : most;
    }
    if (most > (size_t)ft_max_optin_smem()) return (int)cudaErrorInvalidValue;
    static size_t sattr_dev[C3SC_MAXDEV] = {0};
    size_t &sattr = sattr_dev[c3sc_cur_dev()];
    if (most > 48 * 1024 && most > sattr) {
        cudaError_t e = cudaFuncSetAttribute(k_chain_step<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)most);
        if (e != cudaSuccess) return (int)e;
        sattr = most;
    }
    for (int t = 0; t + 1 < a.ft.d; t++) {
        const size_t smem = chain_step_smem(a.rs);
        // steps t >= 1 overlap their launch and table prologue with the tail of step t-1 (programmatic dependent
        // launch; the kernel waits with griddepcontrol.wait before it touches the records)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(CH_NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = (t > 0 && !g_no_pdl) ? 1 : 0;
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_chain_step<KS>, a, t);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

template <int KS>
static int launch_mma_t(const FtArgs &a, const CtlArgs *fused, int nsplit, cudaStream_t st)
{
    // chains: one warp per (fiber, side)
    const size_t csm = FtChainPlan<KS>(a.ft).bytes();
    if (csm > (size_t)ft_max_optin_smem()) return (int)cudaErrorInvalidValue;
    static size_t cattr_dev[C3SC_MAXDEV] = {0};
    size_t &cattr = cattr_dev[c3sc_cur_dev()];
    if (csm > cattr) {
        cudaError_t e = cudaFuncSetAttribute(k_ft_chains<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm);
        if (e != cudaSuccess) return (int)e;
        cattr = csm;
    }
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ft_chains<KS>, FTC_NT, csm);
    if (per_sm < 1) per_sm = 1;
    int cgrid = (2 * a.F + FTC_NT / 32 - 1) / (FTC_NT / 32);
    if (cgrid > ft_sm_count() * per_sm) cgrid = ft_sm_count() * per_sm;
    cudaError_t e = cudaSuccess;
    if (!a.chains_done) {
        k_ft_chains<KS><<<cgrid, FTC_NT, csm, st>>>(a, a.sets);
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    // nodes: one CTA per same-k group
    const size_t smem = FtNodePlan<KS>(a.ft, a.P.nmax).bytes();
    if (smem > (size_t)ft_max_optin_smem()) return (int)cudaErrorInvalidValue;
    static size_t attr_dev[C3SC_MAXDEV] = {0}, attrf_dev[C3SC_MAXDEV] = {0};
    size_t &attr = fused ? attrf_dev[c3sc_cur_dev()] : attr_dev[c3sc_cur_dev()];
    if (smem > attr) {
        e = fused ? cudaFuncSetAttribute(k_ft_nodes_fused<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                  : cudaFuncSetAttribute(k_ft_nodes<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr = smem;
    }
    if (getenv("C3SC_DBG_SKIP_NODES")) return 0;                 // timing experiments only
    const int grid = (a.F + a.FB - 1) / a.FB + a.ft.d;
    // small batches: several CTAs per group, each a range of node tiles, until the machine is covered twice
    FtArgs b = a;
    b.nsplit = nsplit;
    if (fused) {
        if (b.nsplit != 1) return (int)cudaErrorInvalidValue;       // the caller asked ft_nodes_nsplit first
        k_ft_nodes_fused<KS><<<dim3((unsigned)grid, 1u), FTN_NT, smem, st>>>(b, *fused, b.sets);
    } else {
        // programmatic dependent launch behind the chain stage (k_ft_chains / the last k_chain_step trigger early; the kernel waits
        // before it reads their records).  Behind any other kernel the attribute changes nothing: no early trigger, no early start.
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid, (unsigned)b.nsplit); cfg.blockDim = dim3(FTN_NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = getenv("C3SC_NO_PDL") ? 0 : 1;
        const double *setsp = b.sets;
        e = cudaLaunchKernelEx(&cfg, k_ft_nodes<KS>, b, setsp);
        if (e != cudaSuccess) return (int)e;
    }
    return (int)cudaGetLastError();
}

#define C3SC_CAT2(a, b) a##b
#define C3SC_CAT(a, b) C3SC_CAT2(a, b)
int C3SC_CAT(launch_chain_steps_ks, C3SC_KS)(const ChainArgs &a, cudaStream_t st) { return launch_chain_steps_t<C3SC_KS>(a, st); }
int C3SC_CAT(launch_mma_ks, C3SC_KS)(const FtArgs &a, const CtlArgs *fused, int nsplit, cudaStream_t st)
{
    return launch_mma_t<C3SC_KS>(a, fused, nsplit, st);
}

}  // namespace c3sc
