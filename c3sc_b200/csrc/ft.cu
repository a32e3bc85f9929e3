// ft.cu -- host launchers of the model-independent stage-1 kernels (ft_kernel.cuh)
#include "ft_kernel.cuh"

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

int ft_sm_count() { ft_device_info(); return g_sms; }

// perm / kcount / kstart of one chunk; clears *act_count.  1 launch.
int launch_group_fibers(int F, int d, const int *dim_vary, int *perm, int *kcount, int *kstart, int *act_count,
                        cudaStream_t st)
{
    if (F <= 0) return 0;
    k_group_fibers<<<1, 1024, 0, st>>>(F, d, dim_vary, perm, kcount, kstart, act_count);
    return (int)cudaGetLastError();
}

int launch_transpose_cores(const DevFT &ft, double *baseT, cudaStream_t st)
{
    long long most = 0;
    for (int k = 0; k < ft.d; k++) {
        const long long len = (long long)ft.n[k] * ft.r[k] * ft.r[k + 1];
        most = len > most ? len : most;
    }
    if (most <= 0) return 0;
    long long gx = (most + 255) / 256;
    if (gx > 1024) gx = 1024;
    k_transpose_cores<<<dim3((unsigned)gx, (unsigned)ft.d), 256, 0, st>>>(ft, baseT);
    return (int)cudaGetLastError();
}

// Stage 1 over one chunk.  a.FB == 0: pick the group size here.  1 launch.
int launch_ft_costs(const FtArgs &a_in, cudaStream_t st)
{
    ft_device_info();
    FtArgs a = a_in;
    if (a.F <= 0) return 0;
    // two CTAs per SM when the carve-up allows it
    const size_t budget = (size_t)g_max_sm / 2 - 1024;
    if (a.FB <= 0) a.FB = ft_pick_fb(a.ft, a.P.nmax, a.F, g_sms, budget);
    size_t smem = FtPlan(a.ft, a.P.nmax, a.FB).bytes();
    while (smem > (size_t)g_max_optin && a.FB > 1) { a.FB >>= 1; smem = FtPlan(a.ft, a.P.nmax, a.FB).bytes(); }
    if (smem > (size_t)g_max_optin) return (int)cudaErrorInvalidValue;
    static size_t attr = 0;
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(k_ft_costs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr = smem;
    }
    const int grid = (a.F + a.FB - 1) / a.FB + a.ft.d;       // upper bound on the number of groups
    k_ft_costs<<<grid, FT_NT, smem, st>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace c3sc
