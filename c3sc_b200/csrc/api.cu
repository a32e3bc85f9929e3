// api.cu -- the C-ABI of include/c3sc_b200.h: device mirrors + launch wrappers.
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>
#include "dev_types.h"
#define C3SC_FT_TYPES_ONLY
#include "chain_kernel.cuh"     // FtArgs, ChainArgs (host side; the kernels are compiled in ft.cu)
#include "control_kernel.cuh"   // CtlArgs
#include "norm_kernel.cuh"      // k_train_dot_l2 (compiled here)

namespace c3sc {
int launch_control_lqg_lo(int dx, int arith, const CtlArgs &a, int pi_eval, cudaStream_t st);
int launch_control_lqg_hi(int dx, int arith, const CtlArgs &a, int pi_eval, cudaStream_t st);
int launch_control_misc(int model, int dx, int arith, const CtlArgs &a, int pi_eval, cudaStream_t st);
int launch_group_fibers(const DevProblem &P, int F, const ChunkLayout &lay, const int *dim_vary, const int *fixed_ind, int *perm,
                        int *cnt_all, cudaStream_t st);
int launch_pack_cores(const DevFT &ft, double *baseT, double *baseP, double *baseQ, cudaStream_t st);
long long ft_padded_layout(DevFT &ft);
long long ft_compact_layout(DevFT &ft);
int launch_ft_costs(const FtArgs &a, const CtlArgs *fused, cudaStream_t st);
int ft_nodes_nsplit(const FtArgs &a);
long long ft_region_doubles(const DevProblem &P);
int ft_ring_regions();
int fused_ok_lqg_lo(int dx, int arith, const CtlArgs &c, int pi_eval);
int fused_ok_lqg_hi(int dx, int arith, const CtlArgs &c, int pi_eval);
int fused_ok_misc(int model, int dx, int arith, const CtlArgs &c, int pi_eval);
int user_model_nud();
int user_model_dims(int *dx, int *du);
int ft_uses_mma(const DevFT &ft);
int launch_ft_eval_points(const DevProblem &P, const DevFT &ft, int npts, const double *pts, double *out, cudaStream_t st);
int launch_policy_points(const DevProblem &P, int n, const double *x, double *pts, int *absorbed, cudaStream_t st);
size_t ft_sets_bytes(const DevFT &ft, size_t F);
void ft_record_geometry(const DevFT &ft, int *setw, int *rs);
void chain_plan_sizes(const DevFT &ft, int nmax, size_t FC, size_t *kst, size_t *tst, size_t *ent);
int chain_bucketed_ok(const DevFT &ft, int nmax);
int launch_chain_plan(const ChainArgs &a, int FC, int *cntg, cudaStream_t st);
int launch_chain_steps(const ChainArgs &a, cudaStream_t st, int *n);
int launch_rows_move(double *dst, const double *src, const int *idx, int F, long long per, int scatter, cudaStream_t st);
int launch_peer_scatter(const double *src, long long n, double *const *peers, int npeer, long long off, cudaStream_t st);
int launch_node_backup_lqg_lo(int dx, int arith, const DevProblem &P, int n, const double *x, const double *costs,
                              const int *absorbed, double *value, int *argmin, cudaStream_t st);
int launch_node_backup_lqg_hi(int dx, int arith, const DevProblem &P, int n, const double *x, const double *costs,
                              const int *absorbed, double *value, int *argmin, cudaStream_t st);
int launch_node_backup_misc(int model, int dx, int arith, const DevProblem &P, int n, const double *x, const double *costs,
                            const int *absorbed, double *value, int *argmin, cudaStream_t st);
int launch_control_value_lqg_lo(int dx, int arith, const DevProblem &P, int n, const double *x, const double *u,
                                const double *costs, double *value, cudaStream_t st);
int launch_control_value_lqg_hi(int dx, int arith, const DevProblem &P, int n, const double *x, const double *u,
                                const double *costs, double *value, cudaStream_t st);
int launch_control_value_misc(int model, int dx, int arith, const DevProblem &P, int n, const double *x, const double *u,
                              const double *costs, double *value, cudaStream_t st);
int build_ctab_lqg_lo(int dx, const DevProblem &P, double *ctab, cudaStream_t st);
int build_ctab_lqg_hi(int dx, const DevProblem &P, double *ctab, cudaStream_t st);
int build_ctab_misc(int model, int dx, const DevProblem &P, double *ctab, cudaStream_t st);
int launch_model_eval_lqg_lo(int dx, const DevProblem &P, int n, const double *x, const double *u, double *drift,
                             double *sig, double *stage, double *bound, double *obs, cudaStream_t st);
int launch_model_eval_lqg_hi(int dx, const DevProblem &P, int n, const double *x, const double *u, double *drift,
                             double *sig, double *stage, double *bound, double *obs, cudaStream_t st);
int launch_model_eval_misc(int model, int dx, const DevProblem &P, int n, const double *x, const double *u,
                           double *drift, double *sig, double *stage, double *bound, double *obs, cudaStream_t st);
}  // namespace c3sc

using namespace c3sc;

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) return fail(C3SC_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// ---- guard zones (C3SC_GUARD=1): the pool this library is developed on has compute-sanitizer switched off, so the
// library carries its own check.  Every device allocation gets a 64 kB zone of 0xFF bytes on either side (NaN as a
// double, -1 as an int): a write outside an allocation damages a zone (c3sc_guard_check counts damaged zones), and a
// read outside an allocation that reaches a result turns it into NaN, which the parity tests catch.
constexpr size_t GUARD_BYTES = 64 << 10;
struct GuardRec { char *raw; size_t bytes; int device; };
static std::mutex g_guard_mu;
static std::vector<GuardRec> g_guards;
static bool guard_on() { static const bool on = getenv("C3SC_GUARD") != nullptr; return on; }
static cudaError_t dev_malloc(void **p, size_t bytes)
{
    if (!guard_on()) return cudaMalloc(p, bytes);
    char *raw = nullptr;
    const size_t padded = (bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&raw, padded + 2 * GUARD_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaMemset(raw, 0xFF, padded + 2 * GUARD_BYTES);
    if (e != cudaSuccess) { cudaFree(raw); return e; }
    GuardRec r = {raw, padded, 0};
    cudaGetDevice(&r.device);
    { std::lock_guard<std::mutex> lk(g_guard_mu); g_guards.push_back(r); }
    *p = raw + GUARD_BYTES;
    return cudaSuccess;
}
template <class T> static cudaError_t dev_malloc(T **p, size_t bytes) { return dev_malloc((void **)p, bytes); }
static void dev_free(void *p)
{
    if (!p) return;
    if (!guard_on()) { cudaFree(p); return; }
    char *raw = (char *)p - GUARD_BYTES;
    { std::lock_guard<std::mutex> lk(g_guard_mu);
      for (size_t i = 0; i < g_guards.size(); i++) if (g_guards[i].raw == raw) { g_guards.erase(g_guards.begin() + i); break; } }
    cudaFree(raw);
}

// grow-only device scratch
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) dev_free(p);
        p = nullptr; cap = 0;
        if (dev_malloc(&p, bytes) != cudaSuccess) return 1;
        cap = bytes;
        return 0;
    }
    void release() { if (p) dev_free(p); p = nullptr; cap = 0; }
};

// per-problem scratch of the two-stage pipeline (one batch in flight per problem, like the
// reference's Workspace, src/util.c:689-964)
// Pipeline scratch.  perm / cnt belong to the batch; cst / flag / act / sets belong to a chunk and exist once
// per LANE: consecutive chunks of a large batch alternate between the caller's stream and a second one, so
// the latency-bound chain kernel of one chunk and the tails of every kernel overlap the other chunk's work.
constexpr int MAXLANES = 4;
struct LaneScratch {
    DevBuf cst, flag, act, sets, xa, xb;    // xa / xb: row buffers of the bucketed chain stage (chain_kernel.cuh)
    cudaStream_t stream = nullptr;          // lane 0 runs on the caller's stream
    cudaEvent_t join = nullptr;
    // the chain steps of the lane's NEXT super-chunk run on a stream of their own at the highest priority: they are latency-bound
    // (a few resident warps per SM for 10 us a launch), so their CTAs slip in between the node / control CTAs of the other lane
    // instead of queueing behind them
    cudaStream_t chain_stream = nullptr;
    cudaEvent_t chain_done = nullptr, sets_free = nullptr;
    bool sets_busy = false;                 // sets_free was recorded for an earlier super-chunk of this batch
    // host-buffer entries: lanes above 0 copy their finished chunks out on a stream of their own (lane 0 uses the problem's copy
    // stream).  On ONE copy stream the copies wait for their chunks in the order they were queued -- lane 0's whole super-chunk
    // before lane 1's first chunk -- and the second lane's copies end up behind the batch instead of under it.
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_done = nullptr;
    void release()
    {
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (copy_done) cudaEventDestroy(copy_done);
        copy_stream = nullptr; copy_done = nullptr;
        cst.release(); flag.release(); act.release(); sets.release(); xa.release(); xb.release();
        if (stream) cudaStreamDestroy(stream);
        if (join) cudaEventDestroy(join);
        if (chain_stream) cudaStreamDestroy(chain_stream);
        if (chain_done) cudaEventDestroy(chain_done);
        if (sets_free) cudaEventDestroy(sets_free);
        stream = nullptr; join = nullptr; chain_stream = nullptr; chain_done = nullptr; sets_free = nullptr;
    }
};
struct Scratch {
    DevBuf perm, cnt, plan_k, plan_t, plan_e, plan_l, plan_i, plan_c;     // plan_l: row descriptors, plan_i: inverse rows (chain_kernel.cuh)
    DevBuf ring, ring_flag;                 // fused stage 2: per-CTA regions of neighbour values + their in-use flags
    bool ring_ready = false;
    LaneScratch lane[MAXLANES];
    cudaEvent_t fork = nullptr;
    cudaEvent_t start = nullptr, grouped = nullptr;     // the grouping kernel on the second lane, next to the chain plan
    void release()
    {
        if (start) cudaEventDestroy(start);
        if (grouped) cudaEventDestroy(grouped);
        start = grouped = nullptr;
        perm.release(); cnt.release(); plan_k.release(); plan_t.release(); plan_e.release(); plan_l.release(); plan_i.release(); plan_c.release(); ring.release(); ring_flag.release();
        ring_ready = false;
        for (LaneScratch &l : lane) l.release();
        if (fork) cudaEventDestroy(fork);
        fork = nullptr;
    }
};

// candidates of a separable model regrouped by normaliser share (see k_control)
struct CtlGroups {
    int ng = 0;
    int gstart[CT_NGMAX + 1];
    double gA[CT_NGMAX];
    // full {lo, 0, hi}^nud control grid in C order (k_control_grid)
    int grid_on = 0;
    double gWlo[CT_NUDMAX], gWhi[CT_NUDMAX], gAg[CT_NUDMAX + 1], gHg[CT_NUDMAX + 1];
};

/* Is the control table the full grid {lo_m, mid_m, hi_m}^nud, candidate c = sum_m k_m 3^(nud-1-m), whose
 * middle level sits in the dead band of the upwind scheme (no extra weight), whose lo / hi levels load only
 * the left / right neighbour, and whose normaliser share and stage cost depend on the number of non-middle
 * levels only?  tab: the candidate table read back from the device (row stride ct).  All comparisons are
 * exact: the shared-prefix walk must reproduce the per-candidate numbers, not approximate them. */
static void detect_control_grid(const c3sc_problem_desc *d, const std::vector<double> &tab, int nud, int ct, CtlGroups &G)
{
    G.grid_on = 0;
    if (nud < 1 || nud > CT_NUDMAX || (int)d->du != nud) return;
    size_t want = 1;
    for (int m = 0; m < nud; m++) want *= 3;
    if (d->nu != want) return;
    double lev[CT_NUDMAX][3];
    for (int m = 0; m < nud; m++) {
        int nl = 0;
        for (uint32_t c = 0; c < d->nu; c++) {
            const double u = d->controls[(size_t)c * d->du + m];
            int k = 0;
            while (k < nl && lev[m][k] != u) k++;
            if (k < nl) continue;
            if (nl == 3) return;
            lev[m][nl++] = u;
        }
        if (nl != 3) return;
        for (int a = 0; a < 3; a++)
            for (int b = a + 1; b < 3; b++)
                if (lev[m][b] < lev[m][a]) { const double t = lev[m][a]; lev[m][a] = lev[m][b]; lev[m][b] = t; }
    }
    bool haveW[CT_NUDMAX][2] = {}, haveG[CT_NUDMAX + 1] = {};
    for (uint32_t c = 0; c < d->nu; c++) {
        size_t idx = 0;
        int cnt = 0;
        for (int m = 0; m < nud; m++) {
            const double u = d->controls[(size_t)c * d->du + m];
            const int k = u == lev[m][0] ? 0 : (u == lev[m][1] ? 1 : 2);
            idx = idx * 3 + (size_t)k;
            const double wl = tab[(size_t)c * ct + 2 * m], wr = tab[(size_t)c * ct + 2 * m + 1];
            if (k == 1) { if (wl != 0.0 || wr != 0.0) return; continue; }
            cnt++;
            if (k == 0) {
                if (wr != 0.0 || !(wl > 0.0)) return;
                if (!haveW[m][0]) { G.gWlo[m] = wl; haveW[m][0] = true; } else if (G.gWlo[m] != wl) return;
            } else {
                if (wl != 0.0 || !(wr > 0.0)) return;
                if (!haveW[m][1]) { G.gWhi[m] = wr; haveW[m][1] = true; } else if (G.gWhi[m] != wr) return;
            }
        }
        if (idx != c) return;
        const double A = tab[(size_t)c * ct + 2 * nud], H = d->h2 * tab[(size_t)c * ct + 2 * nud + 1];
        if (!haveG[cnt]) { G.gAg[cnt] = A; G.gHg[cnt] = H; haveG[cnt] = true; }
        else if (G.gAg[cnt] != A || G.gHg[cnt] != H) return;
    }
    G.grid_on = 1;
}

struct c3sc_problem {
    DevProblem P;
    int model, arith;
    int device = 0;                          // the device the problem lives on (current device at creation)
    Scratch scr;
    CtlGroups grp;
    double *d_gtab = nullptr;
    std::vector<double> h_utab;              // host copy of the control table (policy entry returns u, not its index)
    double *d_xgrid = nullptr, *d_obs = nullptr, *d_utab = nullptr, *d_ctab = nullptr;
    int *d_err = nullptr;
    int *h_err = nullptr;                    // page-locked landing place of the error word (finish_begin / finish_end)
    cudaStream_t stream = nullptr;           // host-buffer entry points run here
    cudaStream_t copy_stream = nullptr;      // device->host copies of finished chunks
    cudaStream_t peer_stream = nullptr;      // bulk copies of finished chunks into the peers' gathered buffers
    cudaEvent_t chunk_done = nullptr, copies_done = nullptr;
    DevBuf b_dv, b_fi, b_val, b_arg, b_abs, b_costs, b_rows, b_nv, b_nf, b_misc[8];
    c3sc_valuef *vf_flags = nullptr;         // rank-1 zero train for the flags-only entry (c3sc_fiber_flags_batch)
};

struct c3sc_valuef {
    DevFT ft;
    int device = 0;
    double *d_base = nullptr, *d_baseT = nullptr, *d_baseP = nullptr, *d_baseQ = nullptr;
    size_t count = 0;
    // c3sc_valuef_commit on a stream of its own: the event the derived copies are complete at; a batch on ANOTHER stream waits
    // for it right before its first kernel that reads the cores (the plan of the batch does not, and runs ahead)
    cudaEvent_t ready = nullptr;
    cudaStream_t ready_stream = nullptr;
    bool ready_set = false;
    std::vector<size_t> len;
};

// switch the calling thread to a handle's device for the duration of an entry point
struct DeviceScope {
    int prev = -1;
    explicit DeviceScope(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};

// device error word of a problem: [0] transition normaliser < 1e-14 seen, [1] a fiber descriptor outside the
// grid seen (k_group_fibers), [2] smallest such fiber id, [3] spare
static const int k_err_clear[4] = {0, 0, 0x7fffffff, 0};
static int report_error_word(c3sc_problem *p, const int *w)
{
    if (!w[0] && !w[1]) return C3SC_OK;
    cudaMemcpy(p->d_err, k_err_clear, sizeof k_err_clear, cudaMemcpyHostToDevice);
    if (w[1])
        return fail(C3SC_EINVAL, "fiber %d: dim_vary or a fixed index lies outside the grid (src/nodeutil.c:437-470 returns non-zero here); "
                    "the batch's results are undefined", w[2]);
    return fail(C3SC_ENUMERIC, "transition normaliser < 1e-14 at some (node, control): the reference asserts here");
}
static int read_error_word(c3sc_problem *p)
{
    int w[4] = {0, 0, 0, 0};
    CK(cudaMemcpy(w, p->d_err, sizeof w, cudaMemcpyDeviceToHost));
    return report_error_word(p, w);
}

extern "C" {

const char *c3sc_last_error(void) { return g_err; }
/* internal (multi.cu): set the calling thread's message */
int c3sc_set_error(int code, const char *msg) { return fail(code, "%s", msg ? msg : ""); }
const char *c3sc_version(void) { return "c3sc_b200 0.1 (sm_100a)"; }
uint64_t c3sc_launch_count(void) { return g_launches.load(); }

int c3sc_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int c3sc_cuda_init(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(C3SC_ENODEV, "no CUDA device (%s); the Bellman backup has no CPU fallback",
                    e == cudaSuccess ? "count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(C3SC_EINVAL, "device %d out of range [0,%d)", device, n);
    // a process may use several devices (include/c3sc_multi.h: one host thread per device); every handle remembers the
    // device it was created on and its entry points switch to it
    CK(cudaSetDevice(device));
    CK(cudaFree(0));
    return C3SC_OK;
}

/* C3SC_GUARD=1 (set before the first allocation): number of guard zones around the library's device allocations that no
 * longer hold their fill pattern, i.e. allocations something wrote outside of; synchronises the devices.  0 without guards. */
int c3sc_guard_check(void)
{
    if (!guard_on()) return 0;
    std::lock_guard<std::mutex> lk(g_guard_mu);
    int prev = 0, damaged = 0;
    cudaGetDevice(&prev);
    std::vector<unsigned char> h(GUARD_BYTES);
    for (const GuardRec &r : g_guards) {
        cudaSetDevice(r.device);
        cudaDeviceSynchronize();
        for (int side = 0; side < 2; side++) {
            const char *z = side ? r.raw + GUARD_BYTES + r.bytes : r.raw;
            if (cudaMemcpy(h.data(), z, GUARD_BYTES, cudaMemcpyDeviceToHost) != cudaSuccess) { damaged++; continue; }
            bool ok = true;
            for (size_t i = 0; i < GUARD_BYTES && ok; i++) ok = h[i] == 0xFF;
            if (!ok) damaged++;
        }
    }
    cudaSetDevice(prev);
    return damaged;
}

/* page-locked host memory, usable from every device of the process: host buffers handed to the batch entries
 * move at PCIe speed only when they are page-locked (pageable memory goes through the driver's staging copies) */
int c3sc_host_alloc(size_t bytes, void **ptr)
{
    if (!ptr || bytes == 0) return fail(C3SC_EINVAL, "null argument");
    CK(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable));
    return C3SC_OK;
}
int c3sc_host_free(void *ptr)
{
    if (ptr) CK(cudaFreeHost(ptr));
    return C3SC_OK;
}

/* ---- peer-mapped buffers for the fused all-gather (one process per GPU) --------------------------- */
int c3sc_peer_buffer_create(size_t bytes, void **dev, unsigned char handle[64])
{
    if (!dev || !handle || bytes == 0) return fail(C3SC_EINVAL, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    void *p = nullptr;
    CK(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(C3SC_ECUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    memcpy(handle, &h, 64);
    *dev = p;
    return C3SC_OK;
}

int c3sc_peer_buffer_open(const unsigned char handle[64], void **dev)
{
    if (!dev || !handle) return fail(C3SC_EINVAL, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CK(cudaIpcOpenMemHandle(dev, h, cudaIpcMemLazyEnablePeerAccess));     /* maps the peer's allocation, enables P2P */
    return C3SC_OK;
}

int c3sc_peer_buffer_close(void *dev, int opened)
{
    if (!dev) return C3SC_OK;
    if (opened) CK(cudaIpcCloseMemHandle(dev));
    else CK(cudaFree(dev));
    return C3SC_OK;
}

int c3sc_problem_create(const c3sc_problem_desc *d, c3sc_problem **out)
{
    if (!d || !out) return fail(C3SC_EINVAL, "null argument");
    if (d->dx < 1 || d->dx > C3SC_MAXD) return fail(C3SC_EINVAL, "dx=%u outside [1,%d]", d->dx, C3SC_MAXD);
    if (d->nobs > C3SC_MAXOBS) return fail(C3SC_EINVAL, "nobs=%u > %d", d->nobs, C3SC_MAXOBS);
    const bool geometry_only = d->model == C3SC_MODEL_NONE;       // flags + neighbour values only: no dynamics, no control set
    if (!geometry_only && (d->nu < 1 || !d->controls)) return fail(C3SC_EINVAL, "empty control table");
    if (c3sc_cuda_device_count() == 0)
        return fail(C3SC_ENODEV, "no CUDA device; the Bellman backup has no CPU fallback");
    c3sc_problem *p = new c3sc_problem();
    cudaGetDevice(&p->device);
    DevProblem &P = p->P;
    memset(&P, 0, sizeof P);
    P.dx = (int)d->dx; P.du = (int)d->du; P.dw = (int)d->dw; P.nu = (int)d->nu; P.nobs = (int)d->nobs;
    P.h2 = d->h2; P.beta = d->discount;
    p->model = d->model;
    p->arith = d->arith;
    size_t total = 0;
    for (uint32_t i = 0; i < d->dx; i++) {
        if (d->ngrid[i] < 2) { delete p; return fail(C3SC_EINVAL, "ngrid[%u] < 2", i); }
        if (d->bc[i] < C3SC_ABSORB || d->bc[i] > C3SC_REFLECT) { delete p; return fail(C3SC_EINVAL, "bc[%u]=%d unknown", i, d->bc[i]); }
        P.ngrid[i] = (int)d->ngrid[i];
        P.bc[i] = d->bc[i];
        P.xoff[i] = (int)total;
        total += d->ngrid[i];
        if ((int)d->ngrid[i] > P.nmax) P.nmax = (int)d->ngrid[i];
        P.t[2 * i] = d->t[2 * i];
        P.t[2 * i + 1] = d->t[2 * i + 1];
    }
    // model parameter defaults = the reference examples' constants
    static const double defaults[6][8] = {{0}, {1.0, 1.0, 100.0, 0.0}, {1.0, 1.0, 1000.0, 0.0},
                                          {1.0, 1e-2, 1.0, 10.0, 0.0}, {0.0}, {1.0, 0.5, 0.5, 50.0, 0.0}};
    if (d->model < 0 || d->model > 5) { delete p; return fail(C3SC_EUNSUPPORTED, "model %d unknown", d->model); }
    if (d->model == C3SC_MODEL_USER) {
        int udx = 0, udu = 0;
        if (user_model_dims(&udx, &udu)) { delete p; return fail(C3SC_EUNSUPPORTED, "the library was built without a user model (make USER_MODEL=...)"); }
        if ((int)d->dx != udx || (int)d->du != udu) { delete p; return fail(C3SC_EINVAL, "the user model is %d-dimensional with %d controls", udx, udu); }
    }
    if (geometry_only) P.nu = 0;
    memcpy(P.mp, defaults[d->model], sizeof P.mp);
    for (uint32_t i = 0; i < d->n_model_params && i < 8; i++) P.mp[i] = d->model_params[i];

    std::vector<double> xg(total);
    for (uint32_t i = 0; i < d->dx; i++) memcpy(xg.data() + P.xoff[i], d->xgrid[i], d->ngrid[i] * sizeof(double));
    std::vector<double> obs((size_t)d->nobs * 2 * d->dx + 1);
    for (uint32_t o = 0; o < d->nobs; o++) {
        memcpy(obs.data() + (size_t)o * 2 * d->dx, d->obs_lb + (size_t)o * d->dx, d->dx * sizeof(double));
        memcpy(obs.data() + (size_t)o * 2 * d->dx + d->dx, d->obs_ub + (size_t)o * d->dx, d->dx * sizeof(double));
    }
#define CKP(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) { c3sc_problem_destroy(p); return fail(C3SC_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); } \
    } while (0)
    CKP(dev_malloc(&p->d_xgrid, total * sizeof(double)));
    CKP(cudaMemcpy(p->d_xgrid, xg.data(), total * sizeof(double), cudaMemcpyHostToDevice));
    CKP(dev_malloc(&p->d_obs, obs.size() * sizeof(double)));
    CKP(cudaMemcpy(p->d_obs, obs.data(), obs.size() * sizeof(double), cudaMemcpyHostToDevice));
    if (!geometry_only) {
        CKP(dev_malloc(&p->d_utab, (size_t)d->nu * d->du * sizeof(double)));
        CKP(cudaMemcpy(p->d_utab, d->controls, (size_t)d->nu * d->du * sizeof(double), cudaMemcpyHostToDevice));
    }
    CKP(dev_malloc(&p->d_err, 4 * sizeof(int)));
    CKP(cudaMemcpy(p->d_err, k_err_clear, sizeof k_err_clear, cudaMemcpyHostToDevice));
    CKP(cudaHostAlloc((void **)&p->h_err, 4 * sizeof(int), cudaHostAllocPortable));
    CKP(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    CKP(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    CKP(cudaStreamCreateWithFlags(&p->peer_stream, cudaStreamNonBlocking));
    CKP(cudaEventCreateWithFlags(&p->chunk_done, cudaEventDisableTiming));
    CKP(cudaEventCreateWithFlags(&p->copies_done, cudaEventDisableTiming));
    P.xgrid = p->d_xgrid; P.obs = p->d_obs; P.utab = p->d_utab; P.err = p->d_err;
    if (!geometry_only) p->h_utab.assign(d->controls, d->controls + (size_t)d->nu * d->du);
    // candidate table of separable models (row stride 2*NUD+2; NUD = dx/2 for LQG, 1 otherwise)
    if (!geometry_only) {
        const int nud = (d->model == C3SC_MODEL_LQGND) ? (int)d->dx / 2 : (d->model == C3SC_MODEL_SKID5D ? 0 : (d->model == C3SC_MODEL_USER ? user_model_nud() : 1));
        const int ct = 2 * nud + 2;
        CKP(dev_malloc(&p->d_ctab, (size_t)d->nu * ct * sizeof(double)));
        CKP(cudaMemset(p->d_ctab, 0, (size_t)d->nu * ct * sizeof(double)));
        P.ctab = p->d_ctab;
        P.amin = 0.0;
        int rc;
        if (d->model == C3SC_MODEL_LQGND) rc = (P.dx <= 6) ? build_ctab_lqg_lo(P.dx, P, p->d_ctab, 0) : build_ctab_lqg_hi(P.dx, P, p->d_ctab, 0);
        else rc = build_ctab_misc(d->model, P.dx, P, p->d_ctab, 0);
        if (rc == -1) { c3sc_problem_destroy(p); return fail(C3SC_EUNSUPPORTED, "model %d with dx=%u is not instantiated", d->model, d->dx); }
        if (rc != 0) { c3sc_problem_destroy(p); return fail(C3SC_ECUDA, "candidate table kernel: %s", cudaGetErrorString((cudaError_t)rc)); }
        if (nud > 0) {
            g_launches++;
            std::vector<double> tab((size_t)d->nu * ct);
            CKP(cudaMemcpy(tab.data(), p->d_ctab, tab.size() * sizeof(double), cudaMemcpyDeviceToHost));
            double amin = tab[2 * nud];
            for (uint32_t c = 1; c < d->nu; c++) amin = tab[(size_t)c * ct + 2 * nud] < amin ? tab[(size_t)c * ct + 2 * nud] : amin;
            P.amin = amin;
            // regroup by normaliser share A_c (exact equality, first-appearance order); rows of a group
            // keep table order.  Grouped row: [Wl_0, Wr_0, .., h2*gu_c, table index in the low word].
            std::vector<double> shares;
            std::vector<int> gid(d->nu);
            for (uint32_t c = 0; c < d->nu; c++) {
                const double a = tab[(size_t)c * ct + 2 * nud];
                size_t g = 0;
                while (g < shares.size() && shares[g] != a) g++;
                if (g == shares.size()) shares.push_back(a);
                gid[c] = (int)g;
            }
            detect_control_grid(d, tab, nud, ct, p->grp);
            if (shares.size() <= (size_t)CT_NGMAX) {
                std::vector<double> gt((size_t)d->nu * ct);
                CtlGroups &G = p->grp;
                size_t pos = 0;
                for (size_t g = 0; g < shares.size(); g++) {
                    G.gstart[g] = (int)pos;
                    G.gA[g] = shares[g];
                    for (uint32_t c = 0; c < d->nu; c++) {
                        if (gid[c] != (int)g) continue;
                        memcpy(&gt[pos * ct], &tab[(size_t)c * ct], 2 * nud * sizeof(double));
                        gt[pos * ct + 2 * nud] = d->h2 * tab[(size_t)c * ct + 2 * nud + 1];
                        const long long idx = (long long)c;
                        memcpy(&gt[pos * ct + 2 * nud + 1], &idx, sizeof(double));
                        pos++;
                    }
                }
                G.gstart[shares.size()] = (int)pos;
                G.ng = (int)shares.size();
                CKP(dev_malloc(&p->d_gtab, gt.size() * sizeof(double)));
                CKP(cudaMemcpy(p->d_gtab, gt.data(), gt.size() * sizeof(double), cudaMemcpyHostToDevice));
                P.gtab = p->d_gtab;
            }
        }
    }
#undef CKP
    *out = p;
    return C3SC_OK;
}

void c3sc_problem_destroy(c3sc_problem *p)
{
    if (!p) return;
    DeviceScope ds_(p->device);
    dev_free(p->d_xgrid); dev_free(p->d_obs); dev_free(p->d_utab); dev_free(p->d_err); dev_free(p->d_ctab); dev_free(p->d_gtab);
    if (p->h_err) cudaFreeHost(p->h_err);
    DevBuf *bufs[] = {&p->b_dv, &p->b_fi, &p->b_val, &p->b_arg, &p->b_abs, &p->b_costs, &p->b_rows, &p->b_nv, &p->b_nf};
    for (DevBuf *b : bufs) b->release();
    for (DevBuf &b : p->b_misc) b.release();
    p->scr.release();
    if (p->vf_flags) c3sc_valuef_destroy(p->vf_flags);
    if (p->stream) cudaStreamDestroy(p->stream);
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    if (p->peer_stream) cudaStreamDestroy(p->peer_stream);
    if (p->chunk_done) cudaEventDestroy(p->chunk_done);
    if (p->copies_done) cudaEventDestroy(p->copies_done);
    delete p;
}

int c3sc_problem_control_path(const c3sc_problem *p)
{
    if (!p) return -1;
    if (p->arith != C3SC_ARITH_FAST) return 0;
    return p->grp.grid_on ? 2 : (p->grp.ng > 0 ? 1 : 0);
}

int c3sc_problem_check(c3sc_problem *p)
{
    DeviceScope ds_(p ? p->device : c3sc_cur_dev());
    if (!p) return fail(C3SC_EINVAL, "null problem");
    CK(cudaDeviceSynchronize());
    return read_error_word(p);
}

int c3sc_valuef_create(uint32_t d, const uint64_t *n, const uint64_t *ranks, const double *const *cores,
                       c3sc_valuef **out)
{
    if (!n || !ranks || !out || d < 1 || d > C3SC_MAXD) return fail(C3SC_EINVAL, "bad value-function shape");
    if (ranks[0] != 1 || ranks[d] != 1) return fail(C3SC_EINVAL, "boundary ranks must be 1");
    if (c3sc_cuda_device_count() == 0)
        return fail(C3SC_ENODEV, "no CUDA device; the Bellman backup has no CPU fallback");
    c3sc_valuef *vf = new c3sc_valuef();
    cudaGetDevice(&vf->device);
    DevFT &ft = vf->ft;
    memset(&ft, 0, sizeof ft);
    ft.d = (int)d;
    size_t total = 0;
    for (uint32_t k = 0; k < d; k++) {
        ft.n[k] = (int)n[k];
        ft.r[k] = (int)ranks[k];
        ft.off[k] = (long long)total;
        size_t len = n[k] * ranks[k] * ranks[k + 1];
        vf->len.push_back(len);
        total += len;
        if ((int)ranks[k] > ft.rmax) ft.rmax = (int)ranks[k];
    }
    ft.r[d] = 1;
    if (ft.rmax < 1) ft.rmax = 1;
    vf->count = total;
    cudaError_t e = dev_malloc(&vf->d_base, total * sizeof(double));
    if (e == cudaSuccess) e = dev_malloc(&vf->d_baseT, total * sizeof(double));
    if (e != cudaSuccess) { dev_free(vf->d_base); delete vf; return fail(C3SC_ECUDA, "cudaMalloc cores: %s", cudaGetErrorString(e)); }
    ft.base = vf->d_base;
    ft.baseT = vf->d_baseT;
    if (ft_uses_mma(ft)) {
        const long long np = ft_padded_layout(ft);
        e = dev_malloc(&vf->d_baseP, (size_t)np * sizeof(double));
        if (e != cudaSuccess) { dev_free(vf->d_base); dev_free(vf->d_baseT); delete vf; return fail(C3SC_ECUDA, "cudaMalloc padded cores: %s", cudaGetErrorString(e)); }
        ft.baseP = vf->d_baseP;
        const long long nq = ft_compact_layout(ft);
        e = dev_malloc(&vf->d_baseQ, (size_t)nq * sizeof(double));
        if (e == cudaSuccess) e = cudaMemset(vf->d_baseQ, 0, (size_t)nq * sizeof(double));
        if (e != cudaSuccess) { dev_free(vf->d_base); dev_free(vf->d_baseT); dev_free(vf->d_baseP); delete vf; return fail(C3SC_ECUDA, "cudaMalloc tile cores: %s", cudaGetErrorString(e)); }
        ft.baseQ = vf->d_baseQ;
    }
    *out = vf;
    if (cores) {
        int rc = c3sc_valuef_update(vf, cores);
        if (rc) { c3sc_valuef_destroy(vf); *out = nullptr; return rc; }
    }
    return C3SC_OK;
}

int c3sc_valuef_update(c3sc_valuef *vf, const double *const *cores)
{
    DeviceScope ds_(vf ? vf->device : c3sc_cur_dev());
    if (!vf || !cores) return fail(C3SC_EINVAL, "null argument");
    for (int k = 0; k < vf->ft.d; k++)
        CK(cudaMemcpy(vf->d_base + vf->ft.off[k], cores[k], vf->len[k] * sizeof(double), cudaMemcpyHostToDevice));
    int rc = c3sc_valuef_commit(vf, nullptr);
    if (rc) return rc;
    CK(cudaDeviceSynchronize());
    return C3SC_OK;
}

int c3sc_valuef_commit(c3sc_valuef *vf, void *stream)
{
    DeviceScope ds_(vf ? vf->device : c3sc_cur_dev());
    if (!vf) return fail(C3SC_EINVAL, "null argument");
    int rc = launch_pack_cores(vf->ft, vf->d_baseT, vf->d_baseP, vf->d_baseQ, (cudaStream_t)stream);
    if (rc) return fail(C3SC_ECUDA, "core packing kernel: %s", cudaGetErrorString((cudaError_t)rc));
    g_launches++;
    if (!vf->ready) CK(cudaEventCreateWithFlags(&vf->ready, cudaEventDisableTiming));
    CK(cudaEventRecord(vf->ready, (cudaStream_t)stream));
    vf->ready_stream = (cudaStream_t)stream;
    vf->ready_set = true;
    return C3SC_OK;
}

int c3sc_valuef_device_buffer(c3sc_valuef *vf, double **dev, size_t *count)
{
    if (!vf || !dev || !count) return fail(C3SC_EINVAL, "null argument");
    *dev = vf->d_base;
    *count = vf->count;
    return C3SC_OK;
}

}  // extern "C"

// ---- valuef_norm / valuef_norm2diff on the device (norm_kernel.cuh) --------------------------------------------
namespace {
struct NormScratch { DevBuf z, x; std::mutex mu; };
NormScratch g_norm[C3SC_MAXDEV];

// <a,b> (npairs == 1: out[0]) or <a,a>, <a,b>, <b,b> (npairs == 3: out[0..2]) of two value functions on one grid
int train_dots(const c3sc_valuef *va, const c3sc_valuef *vb, const double *const *xgrid, int npairs, double *out)
{
    if (!va || !vb || !xgrid || !out) return fail(C3SC_EINVAL, "null argument");
    if (va->device != vb->device) return fail(C3SC_EINVAL, "the two value functions live on different devices");
    const DevFT &fa = va->ft, &fb = vb->ft;
    if (fa.d != fb.d) return fail(C3SC_EINVAL, "value functions of different dimension");
    DeviceScope ds_(va->device);
    NormArgs na;
    memset(&na, 0, sizeof na);
    na.d = fa.d; na.npairs = npairs;
    int rmax = 1, ntot = 0;
    for (int k = 0; k < fa.d; k++) {
        if (fa.n[k] != fb.n[k]) return fail(C3SC_EINVAL, "grid size mismatch in dim %d", k);
        if (!xgrid[k]) return fail(C3SC_EINVAL, "no node coordinates for dim %d", k);
        na.n[k] = fa.n[k]; na.xoff[k] = ntot; ntot += fa.n[k];
        na.t[0].off[k] = fa.off[k]; na.t[1].off[k] = fb.off[k];
    }
    for (int k = 0; k <= fa.d; k++) {
        na.t[0].r[k] = fa.r[k]; na.t[1].r[k] = fb.r[k];
        rmax = fa.r[k] > rmax ? fa.r[k] : rmax; rmax = fb.r[k] > rmax ? fb.r[k] : rmax;
    }
    na.t[0].base = fa.base; na.t[1].base = fb.base;
    na.zstride = rmax * rmax;
    const size_t smem = (size_t)4 * rmax * rmax * sizeof(double);
    int dev = 0, optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (smem > (size_t)optin) return fail(C3SC_EUNSUPPORTED, "train inner product on the device: rank %d needs %zu bytes of shared memory", rmax, smem);
    NormScratch &ns = g_norm[c3sc_cur_dev()];
    std::lock_guard<std::mutex> lock(ns.mu);
    static size_t attr_dev[C3SC_MAXDEV] = {0};
    if (smem > 48 * 1024 && smem > attr_dev[c3sc_cur_dev()]) {
        CK(cudaFuncSetAttribute(k_train_dot_l2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_dev[c3sc_cur_dev()] = smem;
    }
    const size_t zbytes = (size_t)2 * npairs * NRM_G * na.zstride * sizeof(double);
    if (ns.z.reserve(zbytes) || ns.x.reserve((size_t)ntot * sizeof(double))) return fail(C3SC_ECUDA, "cudaMalloc of the norm scratch failed");
    std::vector<double> hx((size_t)ntot);
    for (int k = 0; k < fa.d; k++) memcpy(hx.data() + na.xoff[k], xgrid[k], (size_t)fa.n[k] * sizeof(double));
    CK(cudaMemcpyAsync(ns.x.p, hx.data(), (size_t)ntot * sizeof(double), cudaMemcpyHostToDevice, 0));
    na.x = (const double *)ns.x.p;
    double *z0 = (double *)ns.z.p, *z1 = z0 + (size_t)npairs * NRM_G * na.zstride;
    for (int k = 0; k < fa.d; k++) {
        k_train_dot_l2<<<dim3(NRM_G, (unsigned)npairs), NRM_NT, smem, 0>>>(na, k, (k & 1) ? z1 : z0, (k & 1) ? z0 : z1);
        g_launches++;
    }
    CK(cudaGetLastError());
    const double *zl = (fa.d & 1) ? z1 : z0;                // launch d-1 wrote z1 when d-1 is even
    std::vector<double> part((size_t)npairs * NRM_G);
    CK(cudaMemcpy2D(part.data(), sizeof(double), zl, (size_t)na.zstride * sizeof(double), sizeof(double), (size_t)npairs * NRM_G,
                    cudaMemcpyDeviceToHost));
    for (int p = 0; p < npairs; p++) {
        double s = 0.0;
        for (int g = 0; g < NRM_G; g++) s += part[(size_t)p * NRM_G + g];
        out[p] = s;
    }
    return C3SC_OK;
}
}  // namespace

extern "C" {
int c3sc_valuef_dot_l2(const c3sc_valuef *a, const c3sc_valuef *b, const double *const *xgrid, double *out)
{
    return train_dots(a, b, xgrid, 1, out);
}
int c3sc_valuef_norm_l2(const c3sc_valuef *a, const double *const *xgrid, double *out)
{
    double v = 0.0;
    const int rc = train_dots(a, a, xgrid, 1, &v);
    if (rc == C3SC_OK && out) *out = sqrt(fabs(v));
    return rc;
}
int c3sc_valuef_norm2diff_l2(const c3sc_valuef *a, const c3sc_valuef *b, const double *const *xgrid, double *out)
{
    double v[3] = {0.0, 0.0, 0.0};
    const int rc = train_dots(a, b, xgrid, 3, v);
    if (rc == C3SC_OK && out) *out = sqrt(fabs(v[0] - 2.0 * v[1] + v[2]));
    return rc;
}
}  // extern "C"

extern "C" {
void c3sc_valuef_destroy(c3sc_valuef *vf)
{
    if (!vf) return;
    DeviceScope ds_(vf->device);
    dev_free(vf->d_base);
    dev_free(vf->d_baseT);
    dev_free(vf->d_baseP);
    dev_free(vf->d_baseQ);
    if (vf->ready) cudaEventDestroy(vf->ready);
    delete vf;
}

}  // extern "C"

// ---------------------------------------------------------------------------
enum Mode { MODE_VI = 0, MODE_PI_EVAL = 1, MODE_COSTS = 2, MODE_STAGE1 = 3 };   // STAGE1: stage 1 into the pipeline's scratch, no stage 2 (measurement)

// One batch through the two-stage pipeline, in chunks whose slot-major cost scratch stays L2-sized:
//   k_group_fibers -> k_ft_costs -> k_control (MODE_VI) | k_pi_eval (MODE_PI_EVAL) | nothing (MODE_COSTS)
struct BatchArgs {
    size_t F, ldo;
    const int *dim_vary, *fixed_ind;      // device
    DevOut out;                           // device, optional
    const double *rows_in;                // MODE_PI_EVAL
    const int *nbr_fixed_in, *nbr_vary_in;
    int mode;
    // host-buffer entries: copy each chunk's results out on a second stream while the next chunk computes
    cudaStream_t copy_stream;
    cudaEvent_t chunk_done;
    double *h_value;
    int32_t *h_argmin;
    // fused all-gather into peer-mapped buffers
    double *value_peers[C3SC_MAXPEERS];
    int n_peers;
    size_t peer_offset;
    int peer_copy;                        // 1: one bulk copy per chunk and peer on peer_stream (copy engines) instead of stores from the kernel
    cudaStream_t peer_stream;
    cudaEvent_t copies_done;
    cudaEvent_t cores_ready;              // the value function was committed on another stream: wait here before the cores are read
    double *eval_value;                   // MODE_VI with rows: after a chunk's improvement, evaluate the rows just written against the
                                          // SAME neighbour values (policy and iterate are one value function) into this buffer
};

static thread_local size_t g_chunk_bytes = (size_t)192 << 20;   // slot-major cost scratch in flight (all lanes) when stage 1a runs per fiber
static thread_local size_t g_chunk_bytes_b = (size_t)384 << 20; // ... when it runs bucketed (measured: smaller, L2-resident scratch chunks lose more to launch
                                                    // granularity than they gain, gpurun_out/r02_tune*.log -> profiles/r02_tuning.md)
static thread_local size_t g_chain_fibers = 32768;               // fibers per chain super-chunk (bucketed stage 1a): 126 MB of records (16384: +1.3 % time)
static thread_local size_t g_chain_min = 4096;                   // smaller batches keep the per-fiber chain kernel
static thread_local int g_lanes = -1;                            // 2 = alternate (super-)chunks between two streams (default), 1 = one stream

static void read_tuning()
{   // re-read on every batch (four getenv calls): tests and tuning runs switch paths inside one process
    const char *e = getenv("C3SC_LANES");
    g_lanes = e ? atoi(e) : 2;
    if (g_lanes < 1) g_lanes = 1;
    if (g_lanes > MAXLANES) g_lanes = MAXLANES;
    const char *m = getenv("C3SC_CHUNK_MB");                          // tuning aids: cost scratch of all lanes together
    g_chunk_bytes = (size_t)192 << 20; g_chunk_bytes_b = (size_t)384 << 20;
    if (m && atoi(m) > 0) { g_chunk_bytes = (size_t)atoi(m) << 20; g_chunk_bytes_b = g_chunk_bytes; }
    const char *cf = getenv("C3SC_CHAIN_FIBERS");
    g_chain_fibers = (cf && atoi(cf) > 0) ? (size_t)atoi(cf) : 32768;
    const char *cm = getenv("C3SC_CHAIN_MIN");
    g_chain_min = (cm && atoi(cm) >= 0) ? (size_t)atoi(cm) : 4096;
}

// Timing aid (C3SC_DBG_TIMELINE=1): events after the plan, every chunk's kernels and every chunk's copy-out; the host-buffer entry
// prints them (ms since the batch's first event) after the batch.  Not for production: the events are created per batch.
struct TimelineMark { cudaEvent_t ev; char what[40]; };
static thread_local std::vector<TimelineMark> g_timeline;
static bool timeline_on() { static const bool on = getenv("C3SC_DBG_TIMELINE") != nullptr; return on; }
static void timeline_mark(cudaStream_t st, const char *fmt, int a = 0, int b = 0)
{
    if (!timeline_on()) return;
    TimelineMark m;
    cudaEventCreate(&m.ev);
    cudaEventRecord(m.ev, st);
    snprintf(m.what, sizeof m.what, fmt, a, b);
    g_timeline.push_back(m);
}
static void timeline_print()
{
    if (!timeline_on() || g_timeline.empty()) return;
    for (auto &m : g_timeline) {
        float ms = 0;
        cudaEventSynchronize(m.ev);
        cudaEventElapsedTime(&ms, g_timeline[0].ev, m.ev);
        fprintf(stderr, "  [timeline] %8.3f ms  %s\n", ms, m.what);
    }
    for (auto &m : g_timeline) cudaEventDestroy(m.ev);
    g_timeline.clear();
}

// One batch through the pipeline.  Units:
//   chunk        FC fibers: grouping by varying dimension, node kernel, control kernel, cost scratch
//   super-chunk  SC consecutive chunks: the bucketed chain stage (chain_kernel.cuh) works on all of them at once,
//                its records stay in L2 until the node kernels of its chunks have consumed them
//   lane         consecutive super-chunks alternate between the caller's stream and a second one
static int run_batch(const DevProblem &P, int model, int arith, Scratch &scr, const CtlGroups *grp, const DevFT &ft,
                     const BatchArgs &b, cudaStream_t st)
{
    const size_t d = (size_t)P.dx, CS = 2 * d + 1, RW = 2 * d + 3;
    read_tuning();
    timeline_mark(st, "batch starts on the device");
    const bool need_cst = b.mode != MODE_COSTS;
    const int mma = ft_uses_mma(ft);
    const bool bucketed = mma && chain_bucketed_ok(ft, P.nmax) && b.F >= g_chain_min && !getenv("C3SC_NO_BUCKETS");
    // several lanes as soon as each (super-)chunk still fills the machine
    const bool multi = g_lanes > 1 && b.mode != MODE_COSTS && b.F * b.ldo >= (size_t)g_lanes * 148 * 1024;
    const size_t L = multi ? (size_t)g_lanes : 1;
    // chain steps of the lane's next super-chunk on a high-priority stream of their own: measured no gain (1.952 vs 1.939 ms per
    // 65 536-fiber step, profiles/r02_cross_step.md: the steps are L2-bandwidth work, not idle latency), so opt-in: C3SC_CHAIN_PRIO=1
    const bool chain_prio = bucketed && multi && getenv("C3SC_CHAIN_PRIO") && atoi(getenv("C3SC_CHAIN_PRIO")) == 1;
    // host-buffer entries copy a chunk's results out while the next chunks compute, each lane on a copy stream of its own.  The link
    // needs ~1.0 ms for a step's values (52 MB at 52 GB/s under load, tools/d2h_under_load.py) and a step's first results appear
    // after ~1.0 ms (plan, both lanes' chain stages, the first chunk): from there on the link is the bound (C3SC_DBG_TIMELINE=1,
    // tools/e2e_timeline.py), so these entries keep EQUAL chunks of the smaller size (first results early): 1.97 ms per 65 536-fiber
    // step.  Measured and not better (profiles/r02c_e2e.md): the large chunks of the device-resident entries with the last chunk of
    // each lane cut into shrinking pieces (C3SC_TAPER=1: 1.99 ms), the first chunk cut into growing pieces as well
    // (C3SC_HEAD_CUTS: 2.01-2.06 ms, the small pieces cost the kernels more than the link gains), lanes staggered by one chain stage
    // (2.05 ms).
    const bool host_out = b.copy_stream && (b.h_value || b.h_argmin);
    const bool taper_on = getenv("C3SC_TAPER") && atoi(getenv("C3SC_TAPER")) == 1;
    const bool taper_want = host_out && multi && taper_on;
    size_t per_chunk = ((bucketed && (!host_out || taper_want)) ? g_chunk_bytes_b : g_chunk_bytes) / (size_t)g_lanes / (b.ldo * CS * 8);
    if (!multi) per_chunk *= (size_t)g_lanes;
    { const char *pf = getenv("C3SC_CHUNK_FIBERS"); if (pf && atoi(pf) > 0) per_chunk = (size_t)atoi(pf); }    // tests: exact chunk size
    if (per_chunk < 1) per_chunk = 1;
    size_t SC = 1;                                          // chunks per super-chunk
    if (bucketed) { SC = (g_chain_fibers + per_chunk / 2) / per_chunk; if (SC < 1) SC = 1; }
    // super-chunks: a multiple of the lane count (equal load per lane), then equal chunks inside
    size_t nsup = (b.F + per_chunk * SC - 1) / (per_chunk * SC);
    if (nsup < L) nsup = L;
    nsup = (nsup + L - 1) / L * L;
    const size_t FC = ((b.F + nsup - 1) / nsup + SC - 1) / SC;        // fibers per chunk
    const size_t FS = FC * SC;                                        // fibers per super-chunk
    nsup = (b.F + FS - 1) / FS;
    const size_t NSmax = FC * b.ldo;
    ChunkLayout lay;
    memset(&lay, 0, sizeof lay);
    lay.FS = (int)FS; lay.FC = (int)FC; lay.SC = (int)SC; lay.m = (int)SC;
    if (taper_want && FC >= 4096) {                         // pieces stay above ~350 fibers
        // The last chunk's pieces shrink by ~0.55, the ratio of the pipeline's fibers per second to the link's: each piece's copy
        // ends about when the next piece's kernels do.  The first chunk's pieces grow the same way: the link moves a step's values
        // in ~1 ms of a ~1.8 ms step and must start long before the first large chunk is through (tools/d2h_under_load.py).
        // C3SC_TAPER_CUTS="a,b,c" / C3SC_HEAD_CUTS="a,b,c": the cuts in thousandths of a chunk (tuning aids; HEAD "0" = no head).
        auto cuts = [](const char *env, int *cut, int ncut) {
            const char *tc = getenv(env);
            if (!tc) return ncut;
            int a = 0, b_ = 0, c_ = 0;
            const int got = sscanf(tc, "%d,%d,%d", &a, &b_, &c_);
            if (got >= 1 && a == 0) return 0;
            if (got >= 1 && a > 0 && a < 1000 && (got < 2 || (b_ > a && b_ < 1000)) && (got < 3 || (c_ > b_ && c_ < 1000))) {
                cut[0] = a; cut[1] = b_; cut[2] = c_;
                return got;
            }
            return ncut;
        };
        int tcut[3] = {494, 766, 915}, hcut[3] = {85, 234, 506};
        lay.nt = cuts("C3SC_TAPER_CUTS", tcut, 3);
        lay.nh = (SC >= 2 && getenv("C3SC_HEAD_CUTS")) ? cuts("C3SC_HEAD_CUTS", hcut, 3) : 0;    // (one chunk per super-chunk: it cannot be both)
        lay.m = (int)SC + lay.nh + lay.nt;
        lay.taper_from = nsup > L ? (int)(nsup - L) : 0;
        lay.head_to = (int)L;
        for (int i = 0; i < lay.nt; i++) lay.tail[i + 1] = (int)((FC * (size_t)tcut[i] / 1000 + 7) & ~(size_t)7);
        for (int i = 0; i < lay.nh; i++) lay.head[i + 1] = (int)((FC * (size_t)hcut[i] / 1000 + 7) & ~(size_t)7);
    }
    const size_t nord = nsup * (size_t)lay.m;               // chunk slots of the batch (some empty)
    int setw = 0, rs = 0;
    if (mma) ft_record_geometry(ft, &setw, &rs);
    if (scr.perm.reserve(b.F * 4) || scr.cnt.reserve(nord * 64 * 4))
        return fail(C3SC_ECUDA, "cudaMalloc pipeline scratch failed");
    size_t pk = 0, pt = 0, pe = 0;
    if (bucketed) {
        chain_plan_sizes(ft, P.nmax, FS, &pk, &pt, &pe);
        if (scr.plan_k.reserve(nsup * pk * 4) || scr.plan_t.reserve(nsup * pt * 4) || scr.plan_e.reserve(nsup * pe * 4) ||
            scr.plan_l.reserve(nsup * (d - 1) * (size_t)chain_x_rows((int)d, P.nmax, (long long)FS) * 16) || scr.plan_i.reserve(nsup * pe * 4) || scr.plan_c.reserve(nsup * pk * 4))
            return fail(C3SC_ECUDA, "cudaMalloc chain plan failed");
    }
    for (size_t l = 0; l < L; l++) {
        LaneScratch &ln = scr.lane[l];
        if ((need_cst && ln.cst.reserve(NSmax * CS * 8)) || ln.flag.reserve(NSmax) || ln.act.reserve(NSmax * 4) ||
            (mma && ln.sets.reserve((size_t)setw * FS * 8)) ||
            (bucketed && (ln.xa.reserve((size_t)chain_x_rows((int)d, P.nmax, (long long)FS) * rs * 8) ||
                          ln.xb.reserve((size_t)chain_x_rows((int)d, P.nmax, (long long)FS) * rs * 8))))
            return fail(C3SC_ECUDA, "cudaMalloc pipeline scratch failed");
        if (l > 0 && !ln.stream) {
            CK(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&ln.join, cudaEventDisableTiming));
        }
        if (chain_prio && !ln.chain_stream) {
            int least = 0, greatest = 0;
            CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
            CK(cudaStreamCreateWithPriority(&ln.chain_stream, cudaStreamNonBlocking, greatest));
            CK(cudaEventCreateWithFlags(&ln.chain_done, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ln.sets_free, cudaEventDisableTiming));
        }
        if (l > 0 && host_out && !ln.copy_stream) {
            CK(cudaStreamCreateWithFlags(&ln.copy_stream, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&ln.copy_done, cudaEventDisableTiming));
        }
        ln.sets_busy = false;
    }
    if (L > 1 && !scr.fork) {
        CK(cudaEventCreateWithFlags(&scr.fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&scr.start, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&scr.grouped, cudaEventDisableTiming));
    }
    auto chain_args = [&](size_t si, size_t s0, size_t Fs) {    // the chain stage's view of super-chunk si = fibers [s0, s0 + Fs)
        ChainArgs ca;
        memset(&ca, 0, sizeof ca);
        ca.P = P; ca.ft = ft; ca.F = (int)Fs; ca.dim_vary = b.dim_vary + s0; ca.fixed_ind = b.fixed_ind + s0 * d;
        ca.nbr_fixed_in = (b.nbr_fixed_in && d > 1) ? b.nbr_fixed_in + s0 * 2 * (d - 1) : nullptr;
        ca.setw = setw; ca.rs = rs;
        ca.kst = (int *)scr.plan_k.p + si * pk; ca.tst = (int *)scr.plan_t.p + si * pt; ca.ent = (int *)scr.plan_e.p + si * pe;
        ca.xrows = chain_x_rows((int)d, P.nmax, (long long)FS);
        {   // plan_l: [nsup] row descriptors (4 B per row) | [nsup] prefix slots (8 B per row) | [nsup] tile buckets (4 B per 8 rows)
            const size_t per = (d - 1) * (size_t)ca.xrows;
            char *pl = (char *)scr.plan_l.p;
            ca.rowd = (int *)pl + si * per;
            ca.rowp = (int2 *)(pl + nsup * per * 4) + si * per;
            ca.tileb = (int *)(pl + nsup * per * 12) + si * (per / 8);
        }
        ca.inv = (int *)scr.plan_i.p + si * pe; ca.invstride = (int)FS;
        ca.nmax = P.nmax; ca.entstride = (int)(3 * FS);
        return ca;
    };
    {   // every chunk's grouping in one launch, every super-chunk's chain plan in four more.  Both read the descriptors only:
        // with a second lane the grouping runs there, next to the plan (19 of the 123 us before the first chain step can start).
        // (Tried and measured slower: the plan of a super-chunk queued on its own lane right before its chain steps, 1.30 against
        // 1.26 ms of stage 1 per 65 536 fibers, and all plans in a row on a third stream, 1.40 ms.)
        const bool aside = L > 1 && bucketed;
        cudaStream_t gst = st;
        if (aside) {
            gst = scr.lane[1].stream;
            CK(cudaEventRecord(scr.start, st));
            CK(cudaStreamWaitEvent(gst, scr.start, 0));
        }
        int rc = launch_group_fibers(P, (int)b.F, lay, b.dim_vary, b.fixed_ind, (int *)scr.perm.p, (int *)scr.cnt.p, gst);
        if (rc) return fail(C3SC_ECUDA, "grouping kernel: %s", cudaGetErrorString((cudaError_t)rc));
        g_launches++;
        if (aside) CK(cudaEventRecord(scr.grouped, gst));
        if (bucketed) {
            ChainArgs ca = chain_args(0, 0, b.F);
            rc = launch_chain_plan(ca, (int)FS, (int *)scr.plan_c.p, st);
            if (rc) return fail(C3SC_ECUDA, "chain plan kernel: %s", cudaGetErrorString((cudaError_t)rc));
            g_launches += 4;
        }
        if (aside) CK(cudaStreamWaitEvent(st, scr.grouped, 0));
    }
    timeline_mark(st, "plan + grouping done");
    if (b.cores_ready) CK(cudaStreamWaitEvent(st, b.cores_ready, 0));       // (grouping and plan above read the descriptors only)
    if (L > 1) {                                            // the other lanes start after everything queued on st so far
        CK(cudaEventRecord(scr.fork, st));
        for (size_t l = 1; l < L; l++) CK(cudaStreamWaitEvent(scr.lane[l].stream, scr.fork, 0));
        if (chain_prio) for (size_t l = 0; l < L; l++) CK(cudaStreamWaitEvent(scr.lane[l].chain_stream, scr.fork, 0));
    }
    const cudaStream_t st0 = st;
    for (size_t s0 = 0, si = 0; s0 < b.F; s0 += FS, si++) {
        const size_t Fs = (b.F - s0 < FS) ? b.F - s0 : FS;
        LaneScratch &ln = scr.lane[si % L];
        st = ln.stream ? ln.stream : st0;
        if (bucketed) {
            ChainArgs ca = chain_args(si, s0, Fs);
            ca.sets = (double *)ln.sets.p;
            ca.x[0] = (double *)ln.xa.p; ca.x[1] = (double *)ln.xb.p;
            int nl = 0;
            cudaStream_t cst_ = st;
            if (chain_prio) {                               // records of this lane are free once its previous super-chunk's nodes are done
                cst_ = ln.chain_stream;
                if (ln.sets_busy) CK(cudaStreamWaitEvent(cst_, ln.sets_free, 0));
            }
            int rc = launch_chain_steps(ca, cst_, &nl);
            if (rc) return fail(C3SC_ECUDA, "chain step kernel: %s", cudaGetErrorString((cudaError_t)rc));
            g_launches += nl;
            timeline_mark(cst_, "lane %d: chain stage of %d fibers", (int)(si % L), (int)Fs);
            if (chain_prio) {
                CK(cudaEventRecord(ln.chain_done, cst_));
                CK(cudaStreamWaitEvent(st, ln.chain_done, 0));
            }
        }
        for (size_t ord = si * (size_t)lay.m; ord < (si + 1) * (size_t)lay.m; ord++) {
        int c0i, Fci;
        ft_chunk_range(lay, (int)b.F, (int)ord, &c0i, &Fci);
        if (Fci <= 0) continue;
        const size_t c0 = (size_t)c0i, Fc = (size_t)Fci;
        const bool last_of_super = c0 + Fc >= s0 + Fs;
        const size_t n0 = c0 * b.ldo;
        DevBuf &bcst = ln.cst, &bflag = ln.flag, &bact = ln.act, &bsets = ln.sets;
        int *cnt = (int *)scr.cnt.p + 64 * ord;         // [0,16) kcount, [16,32) kstart, [32] act_count
        int rc;
        FtArgs a;
        memset(&a, 0, sizeof a);
        a.P = P; a.ft = ft; a.F = (int)Fc; a.dim_vary = b.dim_vary + c0; a.fixed_ind = b.fixed_ind + c0 * d;
        a.ldo = (int)b.ldo; a.FB = 0;
        a.perm = (const int *)scr.perm.p + c0; a.kcount = cnt; a.kstart = cnt + 16;
        a.NS = (long long)(Fc * b.ldo);
        a.cst = need_cst ? (double *)bcst.p : nullptr;
        a.flag = (signed char *)bflag.p;
        a.act = b.mode == MODE_VI ? (int *)bact.p : nullptr;
        a.act_count = cnt + 32;
        a.costs = b.out.costs ? b.out.costs + n0 * CS : nullptr;
        a.absorbed = b.out.absorbed ? b.out.absorbed + n0 : nullptr;
        a.nbr_vary = b.out.nbr_vary ? b.out.nbr_vary + 2 * n0 : nullptr;
        a.nbr_fixed = (b.out.nbr_fixed && d > 1) ? b.out.nbr_fixed + c0 * 2 * (d - 1) : nullptr;
        a.nbr_fixed_in = (b.nbr_fixed_in && d > 1) ? b.nbr_fixed_in + c0 * 2 * (d - 1) : nullptr;
        a.nbr_vary_in = b.nbr_vary_in ? b.nbr_vary_in + 2 * n0 : nullptr;
        a.sets = mma ? (double *)bsets.p + (c0 - s0) * (size_t)setw : nullptr;
        a.setw = setw; a.rs = rs; a.chains_done = bucketed ? 1 : 0;
        a.task_count = cnt + 33;
        if (b.mode == MODE_COSTS || b.mode == MODE_STAGE1) {
            rc = launch_ft_costs(a, nullptr, st);
            if (rc) return fail(C3SC_ECUDA, "FT kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
            g_launches += 1 + (mma && !bucketed);
            if (chain_prio && last_of_super) { CK(cudaEventRecord(ln.sets_free, st)); ln.sets_busy = true; }
            continue;
        }
        CtlArgs c;
        memset(&c, 0, sizeof c);
        c.P = P; c.F = (int)Fc; c.dim_vary = a.dim_vary; c.fixed_ind = a.fixed_ind; c.ldo = (int)b.ldo; c.NS = a.NS;
        c.cst = a.cst; c.flag = a.flag; c.act = a.act; c.act_count = cnt + 32;
        c.value = b.out.value ? b.out.value + n0 : nullptr;
        c.argmin = b.out.argmin ? b.out.argmin + n0 : nullptr;
        c.rows = b.out.rows ? b.out.rows + n0 * RW : nullptr;
        c.rows_in = b.rows_in ? b.rows_in + n0 * RW : nullptr;
        c.npeer = b.peer_copy ? 0 : b.n_peers;
        for (int g = 0; g < b.n_peers; g++) c.vpeer[g] = b.value_peers[g];
        c.peer_off = (long long)(b.peer_offset + n0);
        if (grp && grp->grid_on) {
            c.grid_on = 1;
            memcpy(c.gWlo, grp->gWlo, sizeof c.gWlo); memcpy(c.gWhi, grp->gWhi, sizeof c.gWhi);
            memcpy(c.gAg, grp->gAg, sizeof c.gAg); memcpy(c.gHg, grp->gHg, sizeof c.gHg);
        }
        if (grp && grp->ng > 0 && P.gtab) {
            c.ng = grp->ng;
            memcpy(c.gstart, grp->gstart, sizeof c.gstart);
            memcpy(c.gA, grp->gA, sizeof c.gA);
        }
        const int pe_ = b.mode == MODE_PI_EVAL;
        // Fused stage 2: the node kernel walks the control set of its own fibers from a per-CTA region of neighbour values
        // (no batch-sized scratch through HBM, no second launch) when the model's walk has a fused form, the node kernel
        // owns whole fibers (large batches) and no node-major cost output is wanted.
        // MEASURED SLOWER than the two-kernel pipeline on B200 (2.61 vs 1.94 ms per 65 536-fiber step, profiles/r02_fusion.md:
        // the fully unrolled walk and the node loop evict each other from the instruction cache and the walk runs at half
        // the standalone kernel's occupancy), so it is opt-in: C3SC_FUSE=1.
        bool fuse = mma && !a.costs && getenv("C3SC_FUSE") && ft_nodes_nsplit(a) == 1;
        if (fuse) {
            const int fam = model == C3SC_MODEL_LQGND ? (P.dx <= 6 ? 0 : 1) : 2;
            fuse = fam == 0 ? fused_ok_lqg_lo(P.dx, arith, c, pe_) : (fam == 1 ? fused_ok_lqg_hi(P.dx, arith, c, pe_) : fused_ok_misc(model, P.dx, arith, c, pe_));
            if (fuse) {
                const int nring = ft_ring_regions();
                const long long rd = ft_region_doubles(P);
                if (scr.ring.cap < (size_t)nring * rd * 8 || !scr.ring_ready) {
                    CK(cudaStreamSynchronize(st0));                 // (re)allocation: nothing of this problem may be in flight
                    for (size_t l = 1; l < L; l++) CK(cudaStreamSynchronize(scr.lane[l].stream));
                    if (scr.ring.reserve((size_t)nring * rd * 8) || scr.ring_flag.reserve((size_t)(nring + 1) * 4))
                        return fail(C3SC_ECUDA, "cudaMalloc of the fused stage-2 ring failed");
                    CK(cudaMemset(scr.ring_flag.p, 0, (size_t)(nring + 1) * 4));
                    scr.ring_ready = true;
                }
                a.fused = 1; a.family = fam; a.model = model; a.fuse_pi = pe_; a.fuse_arg = (c.argmin || c.rows) ? 1 : 0;
                a.ring = (double *)scr.ring.p; a.ring_flag = (int *)scr.ring_flag.p; a.nring = nring; a.region_doubles = rd;
                a.cst = nullptr; a.flag = nullptr; a.act = nullptr;
            }
        }
        rc = launch_ft_costs(a, fuse ? &c : nullptr, st);
        if (rc) return fail(C3SC_ECUDA, "FT kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
        g_launches += 1 + (mma && !bucketed);
        if (chain_prio && last_of_super) { CK(cudaEventRecord(ln.sets_free, st)); ln.sets_busy = true; }   // the last reader of the lane's records
        if (fuse) rc = 0;
        else if (model == C3SC_MODEL_LQGND) rc = (P.dx <= 6) ? launch_control_lqg_lo(P.dx, arith, c, pe_, st)
                                                        : launch_control_lqg_hi(P.dx, arith, c, pe_, st);
        else rc = launch_control_misc(model, P.dx, arith, c, pe_, st);
        if (rc == -1) return fail(C3SC_EUNSUPPORTED, "model %d with dx=%d is not instantiated", model, P.dx);
        if (rc != 0) return fail(C3SC_ECUDA, "control kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
        if (!fuse) g_launches++;
        if (b.eval_value && b.mode == MODE_VI && c.rows && !fuse) {
            // bellman_pi's first visit of a node (src/bellman.c:1831-1871) is an improvement against the policy function followed
            // by an evaluation of the new row against the iterate; when the two are the SAME value function the evaluation finds
            // its neighbour values still in the chunk's scratch -- no second pass of stage 1
            CtlArgs e = c;
            e.rows_in = c.rows; e.rows = nullptr; e.argmin = nullptr;
            e.value = b.eval_value + n0;
            e.npeer = 0;
            if (model == C3SC_MODEL_LQGND) rc = (P.dx <= 6) ? launch_control_lqg_lo(P.dx, arith, e, 1, st) : launch_control_lqg_hi(P.dx, arith, e, 1, st);
            else rc = launch_control_misc(model, P.dx, arith, e, 1, st);
            if (rc == -1) return fail(C3SC_EUNSUPPORTED, "model %d with dx=%d is not instantiated", model, P.dx);
            if (rc != 0) return fail(C3SC_ECUDA, "policy evaluation kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
            g_launches++;
        }
        timeline_mark(st, "lane %d: kernels of a chunk of %d fibers", (int)(si % L), (int)Fc);
        if ((b.copy_stream || b.peer_copy) && b.chunk_done) {
            CK(cudaEventRecord(b.chunk_done, st));
            if (b.peer_copy && c.value && b.peer_stream) {  // the chunk's values into every peer's gathered buffer, off the SMs
                CK(cudaStreamWaitEvent(b.peer_stream, b.chunk_done, 0));
                if (b.peer_copy == 2 && ((b.peer_offset + n0) & 1) == 0 && ((size_t)c.value & 15) == 0) {
                    int rs_ = launch_peer_scatter(c.value, (long long)(Fc * b.ldo), b.value_peers, b.n_peers, (long long)(b.peer_offset + n0), b.peer_stream);
                    if (rs_) return fail(C3SC_ECUDA, "peer scatter kernel: %s", cudaGetErrorString((cudaError_t)rs_));
                    g_launches++;
                } else
                for (int g = 0; g < b.n_peers; g++) {
                    double *dst = b.value_peers[g] + b.peer_offset + n0;
                    if (dst != c.value) CK(cudaMemcpyAsync(dst, c.value, Fc * b.ldo * 8, cudaMemcpyDefault, b.peer_stream));
                }
            }
        }
        if (b.copy_stream && b.chunk_done && (b.h_value || b.h_argmin)) {
            const cudaStream_t cps = ln.copy_stream ? ln.copy_stream : b.copy_stream;      // the lane's own copy stream
            CK(cudaStreamWaitEvent(cps, b.chunk_done, 0));
            static const bool dbg_nocopy = getenv("C3SC_DBG_NO_COPYOUT") != nullptr;       // timing aid only: results stay on the device
            if (dbg_nocopy) continue;
            if (b.h_value && c.value) CK(cudaMemcpyAsync(b.h_value + n0, c.value, Fc * b.ldo * 8, cudaMemcpyDeviceToHost, cps));
            if (b.h_argmin && c.argmin) CK(cudaMemcpyAsync(b.h_argmin + n0, c.argmin, Fc * b.ldo * 4, cudaMemcpyDeviceToHost, cps));
            timeline_mark(cps, "lane %d: copy-out of %d fibers", (int)(si % L), (int)Fc);
        }
        }
    }
    for (size_t l = 1; l < L; l++) {                        // the caller's stream continues after every lane
        CK(cudaEventRecord(scr.lane[l].join, scr.lane[l].stream));
        CK(cudaStreamWaitEvent(st0, scr.lane[l].join, 0));
        if (host_out && scr.lane[l].copy_stream) {          // ... and the problem's copy stream after every lane's copies
            CK(cudaEventRecord(scr.lane[l].copy_done, scr.lane[l].copy_stream));
            CK(cudaStreamWaitEvent(b.copy_stream, scr.lane[l].copy_done, 0));
        }
    }
    if (b.peer_copy && b.peer_stream && b.copies_done) {    // ... and after the last peer copy
        CK(cudaEventRecord(b.copies_done, b.peer_stream));
        CK(cudaStreamWaitEvent(st0, b.copies_done, 0));
    }
    return C3SC_OK;
}

static int need_model(const c3sc_problem *p)
{
    if (p && p->model == C3SC_MODEL_NONE)
        return fail(C3SC_EUNSUPPORTED, "geometry-only problem (C3SC_MODEL_NONE): no dynamics / control set for a backup");
    return C3SC_OK;
}

static int check_shapes(const c3sc_problem *p, const c3sc_valuef *vf, size_t F, size_t ldo)
{
    if (!p || !vf) return fail(C3SC_EINVAL, "null problem / value function");
    if (vf->ft.d != p->P.dx) return fail(C3SC_EINVAL, "value function has d=%d, problem dx=%d", vf->ft.d, p->P.dx);
    for (int i = 0; i < p->P.dx; i++)
        if (vf->ft.n[i] != p->P.ngrid[i]) return fail(C3SC_EINVAL, "grid size mismatch in dim %d", i);
    if (ldo < (size_t)p->P.nmax) return fail(C3SC_EINVAL, "ldo=%zu < max ngrid=%d", ldo, p->P.nmax);
    if (F > 0x7fffffffu) return fail(C3SC_EINVAL, "too many fibers in one batch");
    return C3SC_OK;
}

extern "C" {

int c3sc_vi_batch_dev(c3sc_problem *p, const c3sc_valuef *vf, size_t F, const int32_t *d_dim_vary,
                      const int32_t *d_fixed_ind, size_t ldo, const c3sc_batch_out *out, void *stream)
{
    DeviceScope ds_(p ? p->device : c3sc_cur_dev());
    int rc = check_shapes(p, vf, F, ldo);
    if (rc) return rc;
    if (!out || (!out->value && !out->rows && !out->argmin && !out->costs && !out->absorbed && !out->n_peers))
        return fail(C3SC_EINVAL, "no output requested");
    if (F == 0) return C3SC_OK;
    BatchArgs b;
    memset(&b, 0, sizeof b);
    b.F = F; b.ldo = ldo; b.dim_vary = d_dim_vary; b.fixed_ind = d_fixed_ind;
    b.out.value = out->value; b.out.argmin = out->argmin; b.out.absorbed = out->absorbed;
    b.out.costs = out->costs; b.out.rows = out->rows; b.out.nbr_vary = out->nbr_vary; b.out.nbr_fixed = out->nbr_fixed;
    b.mode = (out->value || out->argmin || out->rows || out->n_peers) ? MODE_VI : MODE_COSTS;
    if (b.mode == MODE_VI && (rc = need_model(p))) return rc;
    if (out->n_peers > C3SC_MAXPEERS) return fail(C3SC_EINVAL, "n_peers=%u > %d", out->n_peers, C3SC_MAXPEERS);
    b.n_peers = (int)out->n_peers;
    for (uint32_t g = 0; g < out->n_peers; g++) b.value_peers[g] = out->value_peers[g];
    b.peer_offset = (size_t)out->peer_offset;
    if (out->n_peers && out->peer_mode >= 1) {
        if (!out->value) return fail(C3SC_EINVAL, "peer_mode 1 / 2 (bulk copies) needs the local value buffer");
        b.peer_copy = (int)out->peer_mode;
        b.peer_stream = p->peer_stream; b.chunk_done = p->chunk_done; b.copies_done = p->copies_done;
    }
    if (vf->ready_set && vf->ready_stream != (cudaStream_t)stream) b.cores_ready = vf->ready;
    return run_batch(p->P, p->model, p->arith, p->scr, &p->grp, vf->ft, b, (cudaStream_t)stream);
}

/* Stage 1 alone, exactly as the pipeline runs it (grouping, chain stage, node kernel into the lanes' slot-major
 * scratch; no stage 2, no output): the measurement entry behind bench.py's roofline.stage1_live. */
int c3sc_stage1_batch_dev(c3sc_problem *p, const c3sc_valuef *vf, size_t F, const int32_t *d_dim_vary,
                          const int32_t *d_fixed_ind, size_t ldo, void *stream)
{
    DeviceScope ds_(p ? p->device : c3sc_cur_dev());
    int rc = check_shapes(p, vf, F, ldo);
    if (rc) return rc;
    if (F == 0) return C3SC_OK;
    BatchArgs b;
    memset(&b, 0, sizeof b);
    b.F = F; b.ldo = ldo; b.dim_vary = d_dim_vary; b.fixed_ind = d_fixed_ind;
    b.mode = MODE_STAGE1;
    return run_batch(p->P, p->model, p->arith, p->scr, &p->grp, vf->ft, b, (cudaStream_t)stream);
}

int c3sc_pi_batch_dev(c3sc_problem *p, const c3sc_valuef *vf_policy, const c3sc_valuef *vf_iter, size_t F,
                      const int32_t *d_dim_vary, const int32_t *d_fixed_ind, size_t ldo, int have_rows,
                      double *d_rows, int32_t *d_argmin, double *d_value, void *stream)
{
    DeviceScope ds_(p ? p->device : c3sc_cur_dev());
    int rc = check_shapes(p, vf_iter, F, ldo);
    if (!rc) rc = need_model(p);
    if (rc) return rc;
    if (!d_rows || !d_value) return fail(C3SC_EINVAL, "rows and value buffers are required");
    if (F == 0) return C3SC_OK;
    BatchArgs b;
    memset(&b, 0, sizeof b);
    b.F = F; b.ldo = ldo; b.dim_vary = d_dim_vary; b.fixed_ind = d_fixed_ind;
    if (!have_rows) {                       // policy improvement against vf_policy (bellman.c:1831-1860)
        rc = check_shapes(p, vf_policy, F, ldo);
        if (rc) return rc;
        b.mode = MODE_VI;
        b.out.rows = d_rows;
        b.out.argmin = d_argmin;
        if (vf_policy == vf_iter && !getenv("C3SC_FUSE") && !getenv("C3SC_PI_TWO_PASSES")) {
            b.eval_value = d_value;             // one pass: every chunk's rows are evaluated against the chunk's own scratch
            return run_batch(p->P, p->model, p->arith, p->scr, &p->grp, vf_policy->ft, b, (cudaStream_t)stream);
        }
        rc = run_batch(p->P, p->model, p->arith, p->scr, &p->grp, vf_policy->ft, b, (cudaStream_t)stream);
        if (rc) return rc;
    }
    memset(&b.out, 0, sizeof b.out);        // evaluation against vf_iter (bellman.c:1863-1871)
    b.mode = MODE_PI_EVAL;
    b.out.value = d_value;
    b.rows_in = d_rows;
    return run_batch(p->P, p->model, p->arith, p->scr, &p->grp, vf_iter->ft, b, (cudaStream_t)stream);
}

static int upload_fibers(c3sc_problem *p, size_t F, const int32_t *dim_vary, const int32_t *fixed_ind)
{
    if (p->b_dv.reserve(F * sizeof(int32_t)) || p->b_fi.reserve(F * p->P.dx * sizeof(int32_t)))
        return fail(C3SC_ECUDA, "cudaMalloc fiber descriptors failed");
    CK(cudaMemcpyAsync(p->b_dv.p, dim_vary, F * sizeof(int32_t), cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(p->b_fi.p, fixed_ind, F * p->P.dx * sizeof(int32_t), cudaMemcpyHostToDevice, p->stream));
    return C3SC_OK;
}

/* the batch's error word comes back behind the batch's kernels on the same stream (one round trip, not a
   synchronize followed by a blocking copy); finish_begin may be issued before waiting on other streams */
static int finish_begin(c3sc_problem *p)
{
    CK(cudaMemcpyAsync(p->h_err, p->d_err, 4 * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    return C3SC_OK;
}
static int finish_end(c3sc_problem *p)
{
    CK(cudaStreamSynchronize(p->stream));
    return report_error_word(p, p->h_err);
}
static int finish(c3sc_problem *p)
{
    int rc = finish_begin(p);
    return rc ? rc : finish_end(p);
}

int c3sc_fibers_check(const c3sc_problem *p, size_t F, const int32_t *dim_vary, const int32_t *fixed_ind)
{
    if (!p) return fail(C3SC_EINVAL, "null problem");
    if (F && (!dim_vary || !fixed_ind)) return fail(C3SC_EINVAL, "null fiber descriptors");
    const uint32_t d = (uint32_t)p->P.dx;
    for (size_t f = 0; f < F; f++) {
        unsigned bad = (uint32_t)dim_vary[f] >= d;
        const int32_t *fi = fixed_ind + f * d;
        for (uint32_t i = 0; i < d; i++) bad |= (uint32_t)fi[i] >= (uint32_t)p->P.ngrid[i];
        if (!bad) continue;
        if ((uint32_t)dim_vary[f] >= d)
            return fail(C3SC_EINVAL, "fiber %zu: dim_vary=%d outside [0, %u)", f, (int)dim_vary[f], d);
        for (uint32_t i = 0; i < d; i++)
            if ((uint32_t)fi[i] >= (uint32_t)p->P.ngrid[i])
                return fail(C3SC_EINVAL, "fiber %zu: fixed_ind[%u]=%d outside [0, %d)", f, i, (int)fi[i], p->P.ngrid[i]);
    }
    return C3SC_OK;
}

int c3sc_vi_batch_debug(c3sc_problem *p, const c3sc_valuef *vf, size_t F, const int32_t *dim_vary,
                        const int32_t *fixed_ind, size_t ldo, double *value, int32_t *argmin, int32_t *absorbed,
                        double *costs, double *rows, int32_t *nbr_vary, int32_t *nbr_fixed)
{
    DeviceScope ds_(p ? p->device : c3sc_cur_dev());
    int rc = check_shapes(p, vf, F, ldo);
    if (rc) return rc;
    if (!dim_vary || !fixed_ind || !value) return fail(C3SC_EINVAL, "null fiber descriptors / value buffer");
    if (F == 0) return C3SC_OK;
    rc = c3sc_fibers_check(p, F, dim_vary, fixed_ind);
    if (rc) return rc;
    const size_t dx = p->P.dx, n = F * ldo;
    rc = upload_fibers(p, F, dim_vary, fixed_ind);
    if (rc) return rc;
    c3sc_batch_out o;
    memset(&o, 0, sizeof o);
    int bad = p->b_val.reserve(n * 8);
    o.value = (double *)p->b_val.p;
    if (argmin)    { bad |= p->b_arg.reserve(n * 4); o.argmin = (int32_t *)p->b_arg.p; }
    if (absorbed)  { bad |= p->b_abs.reserve(n * 4); o.absorbed = (int32_t *)p->b_abs.p; }
    if (costs)     { bad |= p->b_costs.reserve(n * (2 * dx + 1) * 8); o.costs = (double *)p->b_costs.p; }
    if (rows)      { bad |= p->b_rows.reserve(n * (2 * dx + 3) * 8); o.rows = (double *)p->b_rows.p; }
    if (nbr_vary)  { bad |= p->b_nv.reserve(n * 2 * 4); o.nbr_vary = (int32_t *)p->b_nv.p; }
    if (nbr_fixed) { bad |= p->b_nf.reserve(F * 2 * (dx > 1 ? dx - 1 : 1) * 4); o.nbr_fixed = (int32_t *)p->b_nf.p; }
    if (bad) return fail(C3SC_ECUDA, "cudaMalloc batch outputs failed");
    // deterministic content for the padding entries j >= ngrid[dim_vary]
    CK(cudaMemsetAsync(o.value, 0, n * 8, p->stream));
    if (argmin)   CK(cudaMemsetAsync(o.argmin, 0xff, n * 4, p->stream));
    if (absorbed) CK(cudaMemsetAsync(o.absorbed, 0, n * 4, p->stream));
    if (costs)    CK(cudaMemsetAsync(o.costs, 0, n * (2 * dx + 1) * 8, p->stream));
    if (rows)     CK(cudaMemsetAsync(o.rows, 0, n * (2 * dx + 3) * 8, p->stream));
    if (nbr_vary) CK(cudaMemsetAsync(o.nbr_vary, 0, n * 2 * 4, p->stream));
    rc = c3sc_vi_batch_dev(p, vf, F, (const int32_t *)p->b_dv.p, (const int32_t *)p->b_fi.p, ldo, &o, p->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(value, o.value, n * 8, cudaMemcpyDeviceToHost, p->stream));
    if (argmin)    CK(cudaMemcpyAsync(argmin, o.argmin, n * 4, cudaMemcpyDeviceToHost, p->stream));
    if (absorbed)  CK(cudaMemcpyAsync(absorbed, o.absorbed, n * 4, cudaMemcpyDeviceToHost, p->stream));
    if (costs)     CK(cudaMemcpyAsync(costs, o.costs, n * (2 * dx + 1) * 8, cudaMemcpyDeviceToHost, p->stream));
    if (rows)      CK(cudaMemcpyAsync(rows, o.rows, n * (2 * dx + 3) * 8, cudaMemcpyDeviceToHost, p->stream));
    if (nbr_vary)  CK(cudaMemcpyAsync(nbr_vary, o.nbr_vary, n * 2 * 4, cudaMemcpyDeviceToHost, p->stream));
    if (nbr_fixed && dx > 1) CK(cudaMemcpyAsync(nbr_fixed, o.nbr_fixed, F * 2 * (dx - 1) * 4, cudaMemcpyDeviceToHost, p->stream));
    return finish(p);
}

static int vi_batch_host(c3sc_problem *p, const c3sc_valuef *vf, size_t F, const int32_t *dim_vary, const int32_t *fixed_ind,
                         size_t ldo, double *value, int32_t *argmin, const c3sc_batch_out *peers);

int c3sc_vi_batch(c3sc_problem *p, const c3sc_valuef *vf, size_t F, const int32_t *dim_vary,
                  const int32_t *fixed_ind, size_t ldo, double *value, int32_t *argmin)
{
    return vi_batch_host(p, vf, F, dim_vary, fixed_ind, ldo, value, argmin, nullptr);
}

/* c3sc_vi_batch of one rank's block of a sharded batch: host descriptors in, host values out, AND the values gathered
 * into every rank's peer-mapped device buffer (peers->value_peers / n_peers / peer_offset / peer_mode; the other
 * fields of *peers are ignored).  The caller orders completion across ranks before reading the gathered buffers. */
int c3sc_vi_batch_peers(c3sc_problem *p, const c3sc_valuef *vf, size_t F, const int32_t *dim_vary,
                        const int32_t *fixed_ind, size_t ldo, double *value, int32_t *argmin, const c3sc_batch_out *peers)
{
    if (!peers || peers->n_peers == 0 || peers->n_peers > C3SC_MAXPEERS) return fail(C3SC_EINVAL, "no peers given");
    return vi_batch_host(p, vf, F, dim_vary, fixed_ind, ldo, value, argmin, peers);
}

static int vi_batch_host(c3sc_problem *p, const c3sc_valuef *vf, size_t F, const int32_t *dim_vary, const int32_t *fixed_ind,
                         size_t ldo, double *value, int32_t *argmin, const c3sc_batch_out *peers)
{
    DeviceScope ds_(p ? p->device : c3sc_cur_dev());
    int rc = check_shapes(p, vf, F, ldo);
    if (!rc) rc = need_model(p);
    if (rc) return rc;
    if (!dim_vary || !fixed_ind || !value) return fail(C3SC_EINVAL, "null fiber descriptors / value buffer");
    if (F == 0) return C3SC_OK;
    const size_t n = F * ldo;
    rc = upload_fibers(p, F, dim_vary, fixed_ind);
    if (rc) return rc;
    // Results are staged on the device and copied chunk by chunk while later chunks compute.  With
    // C3SC_ZEROCOPY=1 and a page-locked, device-mapped result buffer (cudaHostAlloc / cudaHostRegister, e.g.
    // torch pin_memory) the control kernels store the values straight into it over PCIe instead -- no
    // staging buffer, no copy tail; measured SLOWER on B200 + PCIe 5 (2.16 vs 2.45 G node-backups/s end to
    // end: the stores of a control kernel arrive in bursts of ~100 GB/s and stall it), so it is opt-in.
    double *mapped = nullptr;
    {
        const char *zc = getenv("C3SC_ZEROCOPY");
        cudaPointerAttributes at;
        if (zc && zc[0] == '1' && cudaPointerGetAttributes(&at, value) == cudaSuccess && at.type == cudaMemoryTypeHost &&
            at.devicePointer)
            mapped = (double *)at.devicePointer;
        cudaGetLastError();
    }
    c3sc_batch_out o;
    memset(&o, 0, sizeof o);
    int bad = 0;
    if (!mapped) { bad |= p->b_val.reserve(n * 8); o.value = (double *)p->b_val.p; }
    if (argmin) { bad |= p->b_arg.reserve(n * 4); o.argmin = (int32_t *)p->b_arg.p; }
    if (bad) return fail(C3SC_ECUDA, "cudaMalloc batch outputs failed");
    BatchArgs b;
    memset(&b, 0, sizeof b);
    b.F = F; b.ldo = ldo; b.dim_vary = (const int *)p->b_dv.p; b.fixed_ind = (const int *)p->b_fi.p;
    b.out.value = o.value; b.out.argmin = o.argmin;
    b.mode = MODE_VI;
    if (mapped) { b.value_peers[0] = mapped; b.n_peers = 1; b.peer_offset = 0; }
    b.copy_stream = p->copy_stream; b.chunk_done = p->chunk_done; b.h_value = mapped ? nullptr : value; b.h_argmin = argmin;
    if (peers && !mapped) {
        b.n_peers = (int)peers->n_peers;
        for (uint32_t g = 0; g < peers->n_peers; g++) b.value_peers[g] = peers->value_peers[g];
        b.peer_offset = (size_t)peers->peer_offset;
        b.peer_copy = peers->peer_mode >= 1 ? (int)peers->peer_mode : 0;
        b.peer_stream = p->peer_stream; b.copies_done = p->copies_done;
    }
    rc = run_batch(p->P, p->model, p->arith, p->scr, &p->grp, vf->ft, b, p->stream);
    if (rc) return rc;
    rc = finish_begin(p);
    if (rc) return rc;
    CK(cudaStreamSynchronize(p->copy_stream));
    timeline_print();
    return finish_end(p);
}

int c3sc_pi_batch(c3sc_problem *p, const c3sc_valuef *vf_policy, const c3sc_valuef *vf_iter, size_t F,
                  const int32_t *dim_vary, const int32_t *fixed_ind, size_t ldo, int have_rows, double *rows,
                  int32_t *argmin, double *value)
{
    DeviceScope ds_(p ? p->device : c3sc_cur_dev());
    int rc = check_shapes(p, vf_iter, F, ldo);
    if (rc) return rc;
    if (!dim_vary || !fixed_ind || !value || !rows) return fail(C3SC_EINVAL, "null argument");
    if (F == 0) return C3SC_OK;
    const size_t dx = p->P.dx, n = F * ldo, rowb = n * (2 * dx + 3) * 8;
    rc = upload_fibers(p, F, dim_vary, fixed_ind);
    if (rc) return rc;
    if (p->b_val.reserve(n * 8) || p->b_rows.reserve(rowb) || p->b_arg.reserve(n * 4))
        return fail(C3SC_ECUDA, "cudaMalloc batch outputs failed");
    if (have_rows) CK(cudaMemcpyAsync(p->b_rows.p, rows, rowb, cudaMemcpyHostToDevice, p->stream));
    else {
        CK(cudaMemsetAsync(p->b_rows.p, 0, rowb, p->stream));
        CK(cudaMemsetAsync(p->b_arg.p, 0xff, n * 4, p->stream));
    }
    rc = c3sc_pi_batch_dev(p, vf_policy, vf_iter, F, (const int32_t *)p->b_dv.p, (const int32_t *)p->b_fi.p, ldo,
                           have_rows, (double *)p->b_rows.p, (int32_t *)p->b_arg.p, (double *)p->b_val.p, p->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(value, p->b_val.p, n * 8, cudaMemcpyDeviceToHost, p->stream));
    if (!have_rows) {
        CK(cudaMemcpyAsync(rows, p->b_rows.p, rowb, cudaMemcpyDeviceToHost, p->stream));
        if (argmin) CK(cudaMemcpyAsync(argmin, p->b_arg.p, n * 4, cudaMemcpyDeviceToHost, p->stream));
    }
    return finish(p);
}

/* bellman_pi with the policy rows RESIDENT on the device: host fiber descriptors in, host values out, rows in a
 * caller-owned device buffer d_rows [F*ldo*(2dx+3)] that is written when have_rows == 0 and only read otherwise.
 * Nothing of the 184 B/node row record crosses PCIe (the reference keeps it in pi_prob_htable, src/bellman.c:1803-1880). */
int c3sc_pi_batch_resident(c3sc_problem *p, const c3sc_valuef *vf_policy, const c3sc_valuef *vf_iter, size_t F,
                           const int32_t *dim_vary, const int32_t *fixed_ind, size_t ldo, int have_rows, double *d_rows,
                           double *value)
{
    DeviceScope ds_(p ? p->device : c3sc_cur_dev());
    int rc = check_shapes(p, vf_iter, F, ldo);
    if (!rc) rc = need_model(p);
    if (rc) return rc;
    if (!dim_vary || !fixed_ind || !value || !d_rows) return fail(C3SC_EINVAL, "null argument");
    if (F == 0) return C3SC_OK;
    const size_t n = F * ldo;
    rc = upload_fibers(p, F, dim_vary, fixed_ind);
    if (rc) return rc;
    if (p->b_val.reserve(n * 8)) return fail(C3SC_ECUDA, "cudaMalloc batch outputs failed");
    rc = c3sc_pi_batch_dev(p, vf_policy, vf_iter, F, (const int32_t *)p->b_dv.p, (const int32_t *)p->b_fi.p, ldo, have_rows,
                           d_rows, nullptr, (double *)p->b_val.p, p->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(value, p->b_val.p, n * 8, cudaMemcpyDeviceToHost, p->stream));
    return finish(p);
}

/* Device-resident policy-row store, addressed per fiber: capacity fibers x ldo nodes x (2dx+3) doubles.  It is an
 * object of its own because it outlives the per-step device problems: the reference keeps the rows of a policy in
 * the Workspace (pi_prob_htable) across the c3control_step_pi calls of one c3control_pi_solve.
 * c3sc_pi_batch_store runs one bellman_pi batch whose fiber f owns row slot row_id[f]: have_rows == 0 computes the
 * rows and files them there, have_rows != 0 evaluates against the filed rows. */
struct c3sc_rowstore {
    int device = 0;
    uint32_t dx = 0;
    size_t ldo = 0, cap_fibers = 0;
    double *p = nullptr;
};

int c3sc_rowstore_create(uint32_t dx, size_t ldo, c3sc_rowstore **out)
{
    if (!out || dx < 1 || dx > C3SC_MAXD || ldo < 1) return fail(C3SC_EINVAL, "bad row-store shape");
    if (c3sc_cuda_device_count() == 0) return fail(C3SC_ENODEV, "no CUDA device; the Bellman backup has no CPU fallback");
    c3sc_rowstore *s = new c3sc_rowstore();
    cudaGetDevice(&s->device);
    s->dx = dx; s->ldo = ldo;
    *out = s;
    return C3SC_OK;
}

int c3sc_rowstore_reserve(c3sc_rowstore *s, size_t capacity_fibers)
{
    if (!s) return fail(C3SC_EINVAL, "null row store");
    if (capacity_fibers <= s->cap_fibers) return C3SC_OK;
    DeviceScope ds_(s->device);
    const size_t per = s->ldo * (2 * (size_t)s->dx + 3) * sizeof(double);
    double *np = nullptr;
    if (dev_malloc(&np, capacity_fibers * per) != cudaSuccess) {
        cudaGetLastError();
        return fail(C3SC_ECUDA, "cudaMalloc of the policy-row store (%zu MB) failed", (capacity_fibers * per) >> 20);
    }
    if (s->p) {                                             // growing keeps what is filed
        CK(cudaMemcpy(np, s->p, s->cap_fibers * per, cudaMemcpyDeviceToDevice));
        dev_free(s->p);
    }
    s->p = np; s->cap_fibers = capacity_fibers;
    return C3SC_OK;
}

void c3sc_rowstore_destroy(c3sc_rowstore *s)
{
    if (!s) return;
    DeviceScope ds_(s->device);
    dev_free(s->p);
    delete s;
}

int c3sc_pi_batch_store(c3sc_problem *p, const c3sc_valuef *vf_policy, const c3sc_valuef *vf_iter, size_t F,
                        const int32_t *dim_vary, const int32_t *fixed_ind, size_t ldo, int have_rows,
                        c3sc_rowstore *store, const int32_t *row_id, double *value)
{
    DeviceScope ds_(p ? p->device : c3sc_cur_dev());
    int rc = check_shapes(p, vf_iter, F, ldo);
    if (!rc) rc = need_model(p);
    if (rc) return rc;
    if (!dim_vary || !fixed_ind || !value || !row_id || !store) return fail(C3SC_EINVAL, "null argument");
    if (F == 0) return C3SC_OK;
    if (store->ldo != ldo || (int)store->dx != p->P.dx || store->device != p->device)
        return fail(C3SC_EINVAL, "policy-row store was created for dx=%u, ldo=%zu on device %d", store->dx, store->ldo, store->device);
    const size_t RW = 2 * (size_t)p->P.dx + 3, per = ldo * RW, n = F * ldo;
    for (size_t f = 0; f < F; f++)
        if (row_id[f] < 0 || (size_t)row_id[f] >= store->cap_fibers)
            return fail(C3SC_EINVAL, "fiber %zu: row slot %d outside the store (%zu)", f, row_id[f], store->cap_fibers);
    rc = upload_fibers(p, F, dim_vary, fixed_ind);
    if (rc) return rc;
    if (p->b_val.reserve(n * 8) || p->b_rows.reserve(n * RW * 8) || p->b_misc[7].reserve(F * 4))
        return fail(C3SC_ECUDA, "cudaMalloc batch outputs failed");
    CK(cudaMemcpyAsync(p->b_misc[7].p, row_id, F * 4, cudaMemcpyHostToDevice, p->stream));
    if (have_rows) {
        rc = launch_rows_move((double *)p->b_rows.p, store->p, (const int *)p->b_misc[7].p, (int)F, (long long)per, 0, p->stream);
        if (rc) return fail(C3SC_ECUDA, "row gather kernel: %s", cudaGetErrorString((cudaError_t)rc));
        g_launches++;
    }
    rc = c3sc_pi_batch_dev(p, vf_policy, vf_iter, F, (const int32_t *)p->b_dv.p, (const int32_t *)p->b_fi.p, ldo, have_rows,
                           (double *)p->b_rows.p, nullptr, (double *)p->b_val.p, p->stream);
    if (rc) return rc;
    if (!have_rows) {
        rc = launch_rows_move(store->p, (const double *)p->b_rows.p, (const int *)p->b_misc[7].p, (int)F, (long long)per, 1, p->stream);
        if (rc) return fail(C3SC_ECUDA, "row scatter kernel: %s", cudaGetErrorString((cudaError_t)rc));
        g_launches++;
    }
    CK(cudaMemcpyAsync(value, p->b_val.p, n * 8, cudaMemcpyDeviceToHost, p->stream));
    return finish(p);
}

int c3sc_neighbor_costs_batch(c3sc_problem *p, const c3sc_valuef *vf, size_t F, const int32_t *dim_vary,
                              const int32_t *fixed_ind, size_t ldo, int32_t *absorbed, double *costs,
                              int32_t *nbr_vary, int32_t *nbr_fixed)
{
    DeviceScope ds_(p ? p->device : c3sc_cur_dev());
    int rc = check_shapes(p, vf, F, ldo);
    if (rc) return rc;
    if (!dim_vary || !fixed_ind || !absorbed || !costs) return fail(C3SC_EINVAL, "null argument");
    if (F == 0) return C3SC_OK;
    const size_t dx = p->P.dx, n = F * ldo;
    rc = upload_fibers(p, F, dim_vary, fixed_ind);
    if (rc) return rc;
    int bad = p->b_abs.reserve(n * 4) | p->b_costs.reserve(n * (2 * dx + 1) * 8);
    if (nbr_vary) bad |= p->b_nv.reserve(n * 2 * 4);
    if (nbr_fixed) bad |= p->b_nf.reserve(F * 2 * (dx > 1 ? dx - 1 : 1) * 4);
    if (bad) return fail(C3SC_ECUDA, "cudaMalloc batch outputs failed");
    CK(cudaMemsetAsync(p->b_abs.p, 0, n * 4, p->stream));
    CK(cudaMemsetAsync(p->b_costs.p, 0, n * (2 * dx + 1) * 8, p->stream));
    if (nbr_vary) CK(cudaMemsetAsync(p->b_nv.p, 0, n * 2 * 4, p->stream));
    BatchArgs b;
    memset(&b, 0, sizeof b);
    b.F = F; b.ldo = ldo; b.dim_vary = (const int *)p->b_dv.p; b.fixed_ind = (const int *)p->b_fi.p;
    b.mode = MODE_COSTS;
    b.out.absorbed = (int *)p->b_abs.p; b.out.costs = (double *)p->b_costs.p;
    b.out.nbr_vary = nbr_vary ? (int *)p->b_nv.p : nullptr;
    b.out.nbr_fixed = nbr_fixed ? (int *)p->b_nf.p : nullptr;
    rc = run_batch(p->P, p->model, p->arith, p->scr, &p->grp, vf->ft, b, p->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(absorbed, p->b_abs.p, n * 4, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(costs, p->b_costs.p, n * (2 * dx + 1) * 8, cudaMemcpyDeviceToHost, p->stream));
    if (nbr_vary) CK(cudaMemcpyAsync(nbr_vary, p->b_nv.p, n * 2 * 4, cudaMemcpyDeviceToHost, p->stream));
    if (nbr_fixed && dx > 1) CK(cudaMemcpyAsync(nbr_fixed, p->b_nf.p, F * 2 * (dx - 1) * 4, cudaMemcpyDeviceToHost, p->stream));
    return finish(p);
}

/* process_fibers_neighbor (src/nodeutil.c:489-627) over F fibers: flags and neighbour indices only.  The flags are
 * produced by the stage-1 kernels, which want a value function: a rank-1 zero train of the problem's shape is kept
 * inside the problem for this entry. */
int c3sc_fiber_flags_batch(c3sc_problem *p, size_t F, const int32_t *dim_vary, const int32_t *fixed_ind, size_t ldo,
                           int32_t *absorbed, int32_t *nbr_vary, int32_t *nbr_fixed)
{
    DeviceScope ds_(p ? p->device : c3sc_cur_dev());
    if (!p || !dim_vary || !fixed_ind || !absorbed) return fail(C3SC_EINVAL, "null argument");
    if (ldo < (size_t)p->P.nmax) return fail(C3SC_EINVAL, "ldo=%zu < max ngrid=%d", ldo, p->P.nmax);
    if (F == 0) return C3SC_OK;
    if (!p->vf_flags) {
        uint64_t n[C3SC_MAXD], r[C3SC_MAXD + 1];
        std::vector<std::vector<double>> z(p->P.dx);
        const double *cp[C3SC_MAXD];
        for (int i = 0; i < p->P.dx; i++) { n[i] = (uint64_t)p->P.ngrid[i]; r[i] = 1; z[i].assign(n[i], 0.0); cp[i] = z[i].data(); }
        r[p->P.dx] = 1;
        int rc = c3sc_valuef_create((uint32_t)p->P.dx, n, r, cp, &p->vf_flags);
        if (rc) return rc;
    }
    const size_t dx = p->P.dx, n = F * ldo;
    int rc = upload_fibers(p, F, dim_vary, fixed_ind);
    if (rc) return rc;
    int bad = p->b_abs.reserve(n * 4);
    if (nbr_vary) bad |= p->b_nv.reserve(n * 2 * 4);
    if (nbr_fixed) bad |= p->b_nf.reserve(F * 2 * (dx > 1 ? dx - 1 : 1) * 4);
    if (bad) return fail(C3SC_ECUDA, "cudaMalloc batch outputs failed");
    CK(cudaMemsetAsync(p->b_abs.p, 0, n * 4, p->stream));
    if (nbr_vary) CK(cudaMemsetAsync(p->b_nv.p, 0, n * 2 * 4, p->stream));
    BatchArgs b;
    memset(&b, 0, sizeof b);
    b.F = F; b.ldo = ldo; b.dim_vary = (const int *)p->b_dv.p; b.fixed_ind = (const int *)p->b_fi.p;
    b.mode = MODE_COSTS;
    b.out.absorbed = (int *)p->b_abs.p;
    b.out.nbr_vary = nbr_vary ? (int *)p->b_nv.p : nullptr;
    b.out.nbr_fixed = nbr_fixed ? (int *)p->b_nf.p : nullptr;
    rc = run_batch(p->P, p->model, p->arith, p->scr, &p->grp, p->vf_flags->ft, b, p->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(absorbed, p->b_abs.p, n * 4, cudaMemcpyDeviceToHost, p->stream));
    if (nbr_vary) CK(cudaMemcpyAsync(nbr_vary, p->b_nv.p, n * 2 * 4, cudaMemcpyDeviceToHost, p->stream));
    if (nbr_fixed && dx > 1) CK(cudaMemcpyAsync(nbr_fixed, p->b_nf.p, F * 2 * (dx - 1) * 4, cudaMemcpyDeviceToHost, p->stream));
    return finish(p);
}

int c3sc_node_backup_batch(c3sc_problem *p, size_t n, const double *x, const double *costs, const int32_t *absorbed,
                           double *value, int32_t *argmin)
{
    DeviceScope ds_(p ? p->device : c3sc_cur_dev());
    if (!p || !x || !costs || !value) return fail(C3SC_EINVAL, "null argument");
    if (need_model(p)) return C3SC_EUNSUPPORTED;
    if (n == 0) return C3SC_OK;
    const size_t dx = p->P.dx;
    DevBuf *b = p->b_misc;
    if (b[0].reserve(n * dx * 8) || b[1].reserve(n * (2 * dx + 1) * 8) || b[2].reserve(n * 4) || b[3].reserve(n * 8) || b[4].reserve(n * 4))
        return fail(C3SC_ECUDA, "cudaMalloc failed");
    CK(cudaMemcpyAsync(b[0].p, x, n * dx * 8, cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(b[1].p, costs, n * (2 * dx + 1) * 8, cudaMemcpyHostToDevice, p->stream));
    if (absorbed) CK(cudaMemcpyAsync(b[2].p, absorbed, n * 4, cudaMemcpyHostToDevice, p->stream));
    else CK(cudaMemsetAsync(b[2].p, 0, n * 4, p->stream));
    int rc;
    if (p->model == C3SC_MODEL_LQGND)
        rc = (p->P.dx <= 6) ? launch_node_backup_lqg_lo(p->P.dx, p->arith, p->P, (int)n, (const double *)b[0].p, (const double *)b[1].p, (const int *)b[2].p, (double *)b[3].p, (int *)b[4].p, p->stream)
                            : launch_node_backup_lqg_hi(p->P.dx, p->arith, p->P, (int)n, (const double *)b[0].p, (const double *)b[1].p, (const int *)b[2].p, (double *)b[3].p, (int *)b[4].p, p->stream);
    else rc = launch_node_backup_misc(p->model, p->P.dx, p->arith, p->P, (int)n, (const double *)b[0].p, (const double *)b[1].p, (const int *)b[2].p, (double *)b[3].p, (int *)b[4].p, p->stream);
    if (rc == -1) return fail(C3SC_EUNSUPPORTED, "model %d with dx=%d is not instantiated", p->model, p->P.dx);
    if (rc) return fail(C3SC_ECUDA, "kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
    g_launches++;
    CK(cudaMemcpyAsync(value, b[3].p, n * 8, cudaMemcpyDeviceToHost, p->stream));
    if (argmin) CK(cudaMemcpyAsync(argmin, b[4].p, n * 4, cudaMemcpyDeviceToHost, p->stream));
    return finish(p);
}

/* valuef_eval (src/valuefunc.c:345-350) at n arbitrary points */
int c3sc_valuef_eval_batch(c3sc_problem *p, const c3sc_valuef *vf, size_t n, const double *x, double *out)
{
    DeviceScope ds_(p ? p->device : c3sc_cur_dev());
    if (!p || !vf || !x || !out) return fail(C3SC_EINVAL, "null argument");
    if (vf->ft.d != p->P.dx) return fail(C3SC_EINVAL, "value function has d=%d, problem dx=%d", vf->ft.d, p->P.dx);
    if (n == 0) return C3SC_OK;
    const size_t dx = p->P.dx;
    DevBuf *b = p->b_misc;
    if (b[0].reserve(n * dx * 8) || b[1].reserve(n * 8)) return fail(C3SC_ECUDA, "cudaMalloc failed");
    CK(cudaMemcpyAsync(b[0].p, x, n * dx * 8, cudaMemcpyHostToDevice, p->stream));
    int rc = launch_ft_eval_points(p->P, vf->ft, (int)n, (const double *)b[0].p, (double *)b[1].p, p->stream);
    if (rc) return fail(C3SC_ECUDA, "kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
    g_launches++;
    CK(cudaMemcpyAsync(out, b[1].p, n * 8, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return C3SC_OK;
}

/* mca_get_neighbor_node_costs (src/nodeutil.c:718-816) at n off-grid states: V at x -+ h e_i with the boundary
 * stand-ins, all V(x) and flag -1 inside an obstacle.  Needs no dynamics: works on a geometry-only problem. */
int c3sc_neighbor_node_costs_batch(c3sc_problem *p, const c3sc_valuef *vf, size_t n, const double *x, int32_t *absorbed,
                                   double *costs)
{
    DeviceScope ds_(p ? p->device : c3sc_cur_dev());
    if (!p || !vf || !x || !absorbed || !costs) return fail(C3SC_EINVAL, "null argument");
    if (vf->ft.d != p->P.dx) return fail(C3SC_EINVAL, "value function has d=%d, problem dx=%d", vf->ft.d, p->P.dx);
    if (n == 0) return C3SC_OK;
    const size_t dx = p->P.dx, np = 2 * dx + 1;
    DevBuf *b = p->b_misc;
    if (b[0].reserve(n * dx * 8) || b[1].reserve(n * np * dx * 8) || b[2].reserve(n * 4) || b[3].reserve(n * np * 8))
        return fail(C3SC_ECUDA, "cudaMalloc failed");
    CK(cudaMemcpyAsync(b[0].p, x, n * dx * 8, cudaMemcpyHostToDevice, p->stream));
    int rc = launch_policy_points(p->P, (int)n, (const double *)b[0].p, (double *)b[1].p, (int *)b[2].p, p->stream);
    if (!rc) rc = launch_ft_eval_points(p->P, vf->ft, (int)(n * np), (const double *)b[1].p, (double *)b[3].p, p->stream);
    if (rc) return fail(C3SC_ECUDA, "kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
    g_launches += 2;
    CK(cudaMemcpyAsync(absorbed, b[2].p, n * 4, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(costs, b[3].p, n * np * 8, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return C3SC_OK;
}

/* c3control_policy_eval (src/bellman.c:2105-2151) at n states: mca_get_neighbor_node_costs
 * (src/nodeutil.c:718-816) + bellman_optimal.  u [n*du]; value / absorbed / costs may be NULL. */
int c3sc_policy_eval_batch(c3sc_problem *p, const c3sc_valuef *vf, size_t n, const double *x, double *u, double *value,
                           int32_t *absorbed, double *costs)
{
    DeviceScope ds_(p ? p->device : c3sc_cur_dev());
    if (!p || !vf || !x || !u) return fail(C3SC_EINVAL, "null argument");
    if (vf->ft.d != p->P.dx) return fail(C3SC_EINVAL, "value function has d=%d, problem dx=%d", vf->ft.d, p->P.dx);
    if (need_model(p)) return C3SC_EUNSUPPORTED;
    if (n == 0) return C3SC_OK;
    const size_t dx = p->P.dx, du = p->P.du, np = 2 * dx + 1;
    DevBuf *b = p->b_misc;
    if (b[0].reserve(n * dx * 8) || b[1].reserve(n * np * dx * 8) || b[2].reserve(n * 4) || b[3].reserve(n * np * 8) ||
        b[4].reserve(n * 8) || b[5].reserve(n * 4))
        return fail(C3SC_ECUDA, "cudaMalloc failed");
    const double *d_x = (const double *)b[0].p;
    double *d_pts = (double *)b[1].p, *d_costs = (double *)b[3].p, *d_val = (double *)b[4].p;
    int *d_abs = (int *)b[2].p, *d_arg = (int *)b[5].p;
    CK(cudaMemcpyAsync(b[0].p, x, n * dx * 8, cudaMemcpyHostToDevice, p->stream));
    int rc = launch_policy_points(p->P, (int)n, d_x, d_pts, d_abs, p->stream);
    if (!rc) rc = launch_ft_eval_points(p->P, vf->ft, (int)(n * np), d_pts, d_costs, p->stream);
    if (rc) return fail(C3SC_ECUDA, "kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
    if (p->model == C3SC_MODEL_LQGND)
        rc = (p->P.dx <= 6) ? launch_node_backup_lqg_lo(p->P.dx, p->arith, p->P, (int)n, d_x, d_costs, d_abs, d_val, d_arg, p->stream)
                            : launch_node_backup_lqg_hi(p->P.dx, p->arith, p->P, (int)n, d_x, d_costs, d_abs, d_val, d_arg, p->stream);
    else rc = launch_node_backup_misc(p->model, p->P.dx, p->arith, p->P, (int)n, d_x, d_costs, d_abs, d_val, d_arg, p->stream);
    if (rc == -1) return fail(C3SC_EUNSUPPORTED, "model %d with dx=%d is not instantiated", p->model, p->P.dx);
    if (rc) return fail(C3SC_ECUDA, "kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
    g_launches += 3;
    std::vector<int> arg(n);
    CK(cudaMemcpyAsync(arg.data(), d_arg, n * 4, cudaMemcpyDeviceToHost, p->stream));
    if (value) CK(cudaMemcpyAsync(value, d_val, n * 8, cudaMemcpyDeviceToHost, p->stream));
    if (absorbed) CK(cudaMemcpyAsync(absorbed, d_abs, n * 4, cudaMemcpyDeviceToHost, p->stream));
    if (costs) CK(cudaMemcpyAsync(costs, d_costs, n * np * 8, cudaMemcpyDeviceToHost, p->stream));
    rc = finish(p);
    if (rc) return rc;
    for (size_t e = 0; e < n; e++)
        for (size_t i = 0; i < du; i++) u[e * du + i] = arg[e] >= 0 ? p->h_utab[(size_t)arg[e] * du + i] : 0.0;   /* absorbed: u = 0 */
    return C3SC_OK;
}

int c3sc_control_value_batch(c3sc_problem *p, size_t n, const double *x, const double *u, const double *costs,
                             double *value)
{
    DeviceScope ds_(p ? p->device : c3sc_cur_dev());
    if (!p || !x || !u || !costs || !value) return fail(C3SC_EINVAL, "null argument");
    if (need_model(p)) return C3SC_EUNSUPPORTED;
    if (n == 0) return C3SC_OK;
    const size_t dx = p->P.dx, du = p->P.du;
    DevBuf *b = p->b_misc;
    if (b[0].reserve(n * dx * 8) || b[1].reserve(n * du * 8) || b[2].reserve(n * (2 * dx + 1) * 8) || b[3].reserve(n * 8))
        return fail(C3SC_ECUDA, "cudaMalloc failed");
    CK(cudaMemcpyAsync(b[0].p, x, n * dx * 8, cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(b[1].p, u, n * du * 8, cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(b[2].p, costs, n * (2 * dx + 1) * 8, cudaMemcpyHostToDevice, p->stream));
    int rc;
    const double *a0 = (const double *)b[0].p, *a1 = (const double *)b[1].p, *a2 = (const double *)b[2].p;
    if (p->model == C3SC_MODEL_LQGND)
        rc = (p->P.dx <= 6) ? launch_control_value_lqg_lo(p->P.dx, p->arith, p->P, (int)n, a0, a1, a2, (double *)b[3].p, p->stream)
                            : launch_control_value_lqg_hi(p->P.dx, p->arith, p->P, (int)n, a0, a1, a2, (double *)b[3].p, p->stream);
    else rc = launch_control_value_misc(p->model, p->P.dx, p->arith, p->P, (int)n, a0, a1, a2, (double *)b[3].p, p->stream);
    if (rc == -1) return fail(C3SC_EUNSUPPORTED, "model %d with dx=%d is not instantiated", p->model, p->P.dx);
    if (rc) return fail(C3SC_ECUDA, "kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
    g_launches++;
    CK(cudaMemcpyAsync(value, b[3].p, n * 8, cudaMemcpyDeviceToHost, p->stream));
    return finish(p);
}

// ---- entry points that need no problem handle (raw reference signatures) -------------------
// They share static scratch: one call at a time (g_scratch_mu).
static DevBuf g_scratch[6];
static std::mutex g_scratch_mu;

int c3sc_rhs_batch(int arith, uint32_t dx, double discount, size_t n, const double *prob, const double *dt,
                   const double *stage, const double *cost, double *out)
{
    std::lock_guard<std::mutex> lock(g_scratch_mu);
    if (!prob || !dt || !stage || !cost || !out) return fail(C3SC_EINVAL, "null argument");
    if (c3sc_cuda_device_count() == 0) return fail(C3SC_ENODEV, "no CUDA device; no CPU fallback");
    if (n == 0) return C3SC_OK;
    const size_t cs = 2 * (size_t)dx + 1;
    DevBuf *b = g_scratch;
    if (b[0].reserve(n * cs * 8) || b[1].reserve(n * 8) || b[2].reserve(n * 8) || b[3].reserve(n * cs * 8) || b[4].reserve(n * 8))
        return fail(C3SC_ECUDA, "cudaMalloc failed");
    CK(cudaMemcpy(b[0].p, prob, n * cs * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b[1].p, dt, n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b[2].p, stage, n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b[3].p, cost, n * cs * 8, cudaMemcpyHostToDevice));
    int rc = launch_rhs(arith, (int)dx, discount, (int)n, (const double *)b[0].p, (const double *)b[1].p,
                        (const double *)b[2].p, (const double *)b[3].p, (double *)b[4].p, nullptr);
    if (rc) return fail(C3SC_ECUDA, "kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
    g_launches++;
    CK(cudaMemcpy(out, b[4].p, n * 8, cudaMemcpyDeviceToHost));
    return C3SC_OK;
}

int c3sc_transition_raw(int arith, uint32_t dx, double h2, const double *t, size_t n, const double *drift,
                        const double *sigma_diag, double *prob, double *dt, int32_t *status)
{
    std::lock_guard<std::mutex> lock(g_scratch_mu);
    if (!t || !drift || !sigma_diag || !prob || !dt || !status) return fail(C3SC_EINVAL, "null argument");
    if (dx < 1 || dx > C3SC_MAXD) return fail(C3SC_EINVAL, "dx out of range");
    if (c3sc_cuda_device_count() == 0) return fail(C3SC_ENODEV, "no CUDA device; no CPU fallback");
    if (n == 0) return C3SC_OK;
    DevProblem P;
    memset(&P, 0, sizeof P);
    P.dx = (int)dx; P.h2 = h2;
    for (uint32_t i = 0; i < 2 * dx; i++) P.t[i] = t[i];
    DevBuf *b = g_scratch;
    if (b[0].reserve(n * dx * 8) || b[1].reserve(n * dx * 8) || b[2].reserve(n * (2 * dx + 1) * 8) || b[3].reserve(n * 8) || b[4].reserve(n * 4))
        return fail(C3SC_ECUDA, "cudaMalloc failed");
    CK(cudaMemcpy(b[0].p, drift, n * dx * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b[1].p, sigma_diag, n * dx * 8, cudaMemcpyHostToDevice));
    int rc = launch_transition(arith, P, (int)n, (const double *)b[0].p, (const double *)b[1].p, (double *)b[2].p,
                               (double *)b[3].p, (int *)b[4].p, nullptr);
    if (rc == -1) return fail(C3SC_EUNSUPPORTED, "dx=%u not instantiated for the transition kernel", dx);
    if (rc) return fail(C3SC_ECUDA, "kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
    g_launches++;
    CK(cudaMemcpy(prob, b[2].p, n * (2 * dx + 1) * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(dt, b[3].p, n * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(status, b[4].p, n * 4, cudaMemcpyDeviceToHost));
    return C3SC_OK;
}

int c3sc_ft_fiber_nn_batch(const c3sc_valuef *vf, size_t F, const int32_t *dim_vary, const int32_t *fixed_ind,
                           const int32_t *nbr_fixed, const int32_t *nbr_vary, size_t ldo, double *costs)
{
    DeviceScope ds_(vf ? vf->device : c3sc_cur_dev());
    std::lock_guard<std::mutex> lock(g_scratch_mu);
    if (!vf || !dim_vary || !fixed_ind || !nbr_fixed || !nbr_vary || !costs) return fail(C3SC_EINVAL, "null argument");
    if (F == 0) return C3SC_OK;
    const int dx = vf->ft.d;
    const int model = 0;   // stage 1 is independent of the dynamics model
    struct { DevProblem P; } tmp;
    memset(&tmp.P, 0, sizeof tmp.P);
    tmp.P.dx = dx; tmp.P.nu = 0;
    for (int i = 0; i < dx; i++) { tmp.P.ngrid[i] = vf->ft.n[i]; tmp.P.bc[i] = C3SC_REFLECT; if (vf->ft.n[i] > tmp.P.nmax) tmp.P.nmax = vf->ft.n[i]; }
    if (ldo < (size_t)tmp.P.nmax) return fail(C3SC_EINVAL, "ldo too small");
    const size_t n = F * ldo, cs = 2 * (size_t)dx + 1, nfix = 2 * (size_t)(dx > 1 ? dx - 1 : 1);
    DevBuf *b = g_scratch;
    if (b[0].reserve(F * 4) || b[1].reserve(F * dx * 4) || b[2].reserve(F * nfix * 4) || b[3].reserve(n * 2 * 4) ||
        b[4].reserve(n * cs * 8) || b[5].reserve(8 * (size_t)dx * tmp.P.nmax + 64))
        return fail(C3SC_ECUDA, "cudaMalloc failed");
    CK(cudaMemcpy(b[0].p, dim_vary, F * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b[1].p, fixed_ind, F * dx * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b[2].p, nbr_fixed, F * 2 * (size_t)(dx - 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b[3].p, nbr_vary, n * 2 * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(b[4].p, 0, n * cs * 8));
    CK(cudaMemset(b[5].p, 0, 8 * (size_t)dx * tmp.P.nmax + 64));
    tmp.P.xgrid = (const double *)b[5].p;           // coordinates are irrelevant here (no obstacles)
    for (int i = 0; i < dx; i++) tmp.P.xoff[i] = i * tmp.P.nmax;
    tmp.P.err = (int *)((char *)b[5].p + 8 * (size_t)dx * tmp.P.nmax);
    CK(cudaMemcpy(tmp.P.err, k_err_clear, sizeof k_err_clear, cudaMemcpyHostToDevice));
    BatchArgs ba;
    memset(&ba, 0, sizeof ba);
    ba.F = F; ba.ldo = ldo; ba.dim_vary = (const int *)b[0].p; ba.fixed_ind = (const int *)b[1].p;
    ba.mode = MODE_COSTS; ba.out.costs = (double *)b[4].p;
    ba.nbr_fixed_in = (const int *)b[2].p; ba.nbr_vary_in = (const int *)b[3].p;
    static Scratch scr;
    int rc = run_batch(tmp.P, model, C3SC_ARITH_FAST, scr, nullptr, vf->ft, ba, nullptr);
    if (rc) return rc;
    CK(cudaDeviceSynchronize());
    int w[4];
    CK(cudaMemcpy(w, tmp.P.err, sizeof w, cudaMemcpyDeviceToHost));
    if (w[1]) return fail(C3SC_EINVAL, "fiber %d: dim_vary or a fixed index lies outside the grid", w[2]);
    CK(cudaMemcpy(costs, b[4].p, n * cs * 8, cudaMemcpyDeviceToHost));
    return C3SC_OK;
}

int c3sc_transition_batch(c3sc_problem *p, size_t n, const double *drift, const double *sigma_diag,
                          double *prob, double *dt, int32_t *status)
{
    DeviceScope ds_(p ? p->device : c3sc_cur_dev());
    if (!p || !drift || !sigma_diag || !prob || !dt || !status) return fail(C3SC_EINVAL, "null argument");
    if (need_model(p)) return C3SC_EUNSUPPORTED;
    if (n == 0) return C3SC_OK;
    const size_t dx = p->P.dx;
    DevBuf *b = p->b_misc;
    if (b[0].reserve(n * dx * 8) || b[1].reserve(n * dx * 8) || b[2].reserve(n * (2 * dx + 1) * 8) ||
        b[3].reserve(n * 8) || b[4].reserve(n * 4))
        return fail(C3SC_ECUDA, "cudaMalloc failed");
    CK(cudaMemcpyAsync(b[0].p, drift, n * dx * 8, cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(b[1].p, sigma_diag, n * dx * 8, cudaMemcpyHostToDevice, p->stream));
    int rc = launch_transition(p->arith, p->P, (int)n, (const double *)b[0].p, (const double *)b[1].p,
                               (double *)b[2].p, (double *)b[3].p, (int *)b[4].p, p->stream);
    if (rc == -1) return fail(C3SC_EUNSUPPORTED, "dx=%d not instantiated for the transition kernel", p->P.dx);
    if (rc) return fail(C3SC_ECUDA, "kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
    g_launches++;
    CK(cudaMemcpyAsync(prob, b[2].p, n * (2 * dx + 1) * 8, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(dt, b[3].p, n * 8, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(status, b[4].p, n * 4, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return C3SC_OK;
}

int c3sc_model_eval(c3sc_problem *p, size_t n, const double *x, const double *u, double *drift,
                    double *sigma_diag, double *stage, double *bound, double *obs)
{
    DeviceScope ds_(p ? p->device : c3sc_cur_dev());
    if (!p || !x || !u || !drift || !sigma_diag || !stage || !bound || !obs) return fail(C3SC_EINVAL, "null argument");
    if (need_model(p)) return C3SC_EUNSUPPORTED;
    if (n == 0) return C3SC_OK;
    const size_t dx = p->P.dx, du = p->P.du;
    DevBuf *b = p->b_misc;
    if (b[0].reserve(n * dx * 8) || b[1].reserve(n * du * 8) || b[2].reserve(n * dx * 8) || b[3].reserve(n * dx * 8) ||
        b[4].reserve(n * 8) || b[5].reserve(n * 8) || b[6].reserve(n * 8))
        return fail(C3SC_ECUDA, "cudaMalloc failed");
    CK(cudaMemcpyAsync(b[0].p, x, n * dx * 8, cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(b[1].p, u, n * du * 8, cudaMemcpyHostToDevice, p->stream));
    int rc;
    const double *dxp = (const double *)b[0].p, *dup = (const double *)b[1].p;
    double *o0 = (double *)b[2].p, *o1 = (double *)b[3].p, *o2 = (double *)b[4].p, *o3 = (double *)b[5].p, *o4 = (double *)b[6].p;
    if (p->model == C3SC_MODEL_LQGND)
        rc = (p->P.dx <= 6) ? launch_model_eval_lqg_lo(p->P.dx, p->P, (int)n, dxp, dup, o0, o1, o2, o3, o4, p->stream)
                            : launch_model_eval_lqg_hi(p->P.dx, p->P, (int)n, dxp, dup, o0, o1, o2, o3, o4, p->stream);
    else rc = launch_model_eval_misc(p->model, p->P.dx, p->P, (int)n, dxp, dup, o0, o1, o2, o3, o4, p->stream);
    if (rc == -1) return fail(C3SC_EUNSUPPORTED, "model %d with dx=%d is not instantiated", p->model, p->P.dx);
    if (rc) return fail(C3SC_ECUDA, "kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
    g_launches++;
    CK(cudaMemcpyAsync(drift, o0, n * dx * 8, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(sigma_diag, o1, n * dx * 8, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(stage, o2, n * 8, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(bound, o3, n * 8, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(obs, o4, n * 8, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return C3SC_OK;
}

}  // extern "C"
