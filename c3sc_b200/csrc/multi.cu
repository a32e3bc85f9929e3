// multi.cu -- include/c3sc_multi.h: one process, several GPUs of one box.
//
// A worker thread per device (bound to it with cudaSetDevice once) executes the per-device share of every call through
// the single-device entries of api.cu; the calling thread only cuts the batch and waits.  Cores: host -> device 0 ->
// ncclBroadcast (NCCL loaded with dlopen, so the library has no link-time dependency on it; peer copies otherwise).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include "../../include/c3sc_multi.h"

extern "C" int c3sc_set_error(int code, const char *msg);     // api.cu: message of the calling thread

namespace {

// ---- the few NCCL entry points used, resolved at run time ---------------------------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;
enum { ncclFloat64 = 8 };
struct Nccl {
    void *h = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load()
    {
        if (getenv("C3SC_NO_NCCL")) return false;
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(h, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
        GroupStart = (decltype(GroupStart))dlsym(h, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(h, "ncclGroupEnd");
        Broadcast = (decltype(Broadcast))dlsym(h, "ncclBroadcast");
        AllGather = (decltype(AllGather))dlsym(h, "ncclAllGather");
        GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
        return CommInitAll && CommDestroy && GroupStart && GroupEnd && Broadcast && AllGather;
    }
};

// ---- one worker thread per device ----------------------------------------------------------------------------------
struct Worker {
    int device = 0;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = true, quit = false;
    int rc = 0;
    char err[512] = "";
    void run()
    {
        cudaSetDevice(device);
        cudaFree(0);
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return has_job || quit; });
            if (quit) return;
            std::function<int()> j = std::move(job);
            has_job = false;
            lk.unlock();
            const int r = j();
            if (r) snprintf(err, sizeof err, "device %d: %s", device, c3sc_last_error());
            lk.lock();
            rc = r;
            done = true;
            cv.notify_all();
        }
    }
    void submit(std::function<int()> j)
    {
        std::lock_guard<std::mutex> lk(mu);
        job = std::move(j); has_job = true; done = false;
        cv.notify_all();
    }
    int wait()
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done; });
        return rc;
    }
};

}  // namespace

struct c3sc_multi {
    int G = 0;
    std::vector<int> dev;
    std::vector<c3sc_problem *> prob;
    std::vector<Worker *> work;
    std::vector<cudaStream_t> stream;             // per device: broadcast / gather
    Nccl nccl;
    std::vector<ncclComm_t> comm;
    bool use_nccl = false;
    uint32_t dx = 0;
    // resident policy rows: [slot][device]
    struct Rows { void *p = nullptr; size_t cap = 0; };
    std::vector<std::vector<Rows>> rows;
    // staging of the gathered variant
    std::vector<void *> d_dv, d_fi;
    std::vector<size_t> cap_dv, cap_fi;
};

struct c3sc_multi_valuef {
    c3sc_multi *m = nullptr;
    std::vector<c3sc_valuef *> vf;
    uint32_t d = 0;
    std::vector<size_t> len;
};

static int run_all(c3sc_multi *m, const std::function<int(int)> &f)
{
    for (int g = 0; g < m->G; g++) m->work[g]->submit([f, g] { return f(g); });
    int rc = 0;
    const char *msg = nullptr;
    for (int g = 0; g < m->G; g++) {
        const int r = m->work[g]->wait();
        if (r && !rc) { rc = r; msg = m->work[g]->err; }
    }
    if (rc) return c3sc_set_error(rc, msg);
    return C3SC_OK;
}

extern "C" {

void c3sc_multi_shard(size_t F, int G, int g, size_t *begin, size_t *end)
{
    const size_t per = G > 0 ? (F + (size_t)G - 1) / (size_t)G : F;
    size_t b = per * (size_t)g, e = b + per;
    if (b > F) b = F;
    if (e > F) e = F;
    if (begin) *begin = b;
    if (end) *end = e;
}

int c3sc_multi_create(const c3sc_problem_desc *desc, int ndev, const int *devices, c3sc_multi **out)
{
    if (!desc || !out) return c3sc_set_error(C3SC_EINVAL, "null argument");
    int have = c3sc_cuda_device_count();
    if (have == 0) return c3sc_set_error(C3SC_ENODEV, "no CUDA device; the Bellman backup has no CPU fallback");
    if (ndev <= 0) ndev = have;
    if (ndev > have || ndev > C3SC_MAXPEERS) return c3sc_set_error(C3SC_EINVAL, "more devices requested than visible (or than C3SC_MAXPEERS)");
    c3sc_multi *m = new c3sc_multi();
    m->G = ndev;
    m->dx = desc->dx;
    for (int g = 0; g < ndev; g++) m->dev.push_back(devices ? devices[g] : g);
    m->prob.assign(ndev, nullptr);
    m->stream.assign(ndev, nullptr);
    m->d_dv.assign(ndev, nullptr); m->d_fi.assign(ndev, nullptr);
    m->cap_dv.assign(ndev, 0); m->cap_fi.assign(ndev, 0);
    m->rows.assign(64, std::vector<c3sc_multi::Rows>(ndev));
    for (int g = 0; g < ndev; g++) {
        Worker *w = new Worker();
        w->device = m->dev[g];
        w->th = std::thread([w] { w->run(); });
        m->work.push_back(w);
    }
    int rc = run_all(m, [m, desc](int g) {
        int r = c3sc_problem_create(desc, &m->prob[g]);
        if (r) return r;
        if (cudaStreamCreateWithFlags(&m->stream[g], cudaStreamNonBlocking) != cudaSuccess) return (int)C3SC_ECUDA;
        for (int h = 0; h < m->G; h++) {                      // peer access for the copy fallback and for NCCL's P2P transport
            if (h == g) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, m->dev[g], m->dev[h]);
            if (can) { cudaDeviceEnablePeerAccess(m->dev[h], 0); cudaGetLastError(); }
        }
        return (int)C3SC_OK;
    });
    if (rc) { c3sc_multi_destroy(m); return rc; }
    if (ndev > 1 && m->nccl.load()) {
        m->comm.assign(ndev, nullptr);
        if (m->nccl.CommInitAll(m->comm.data(), ndev, m->dev.data()) == 0) m->use_nccl = true;
        else m->comm.clear();
    }
    *out = m;
    return C3SC_OK;
}

void c3sc_multi_destroy(c3sc_multi *m)
{
    if (!m) return;
    if (!m->work.empty() && (int)m->work.size() == m->G) {
        run_all(m, [m](int g) {
            for (auto &slot : m->rows) if (slot[g].p) { cudaFree(slot[g].p); slot[g].p = nullptr; }
            if (m->d_dv[g]) cudaFree(m->d_dv[g]);
            if (m->d_fi[g]) cudaFree(m->d_fi[g]);
            if (m->prob[g]) c3sc_problem_destroy(m->prob[g]);
            if (m->stream[g]) cudaStreamDestroy(m->stream[g]);
            return 0;
        });
    }
    if (m->use_nccl) for (ncclComm_t c : m->comm) if (c) m->nccl.CommDestroy(c);
    for (Worker *w : m->work) {
        { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; w->cv.notify_all(); }
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    delete m;
}

int c3sc_multi_device_count(const c3sc_multi *m) { return m ? m->G : 0; }
int c3sc_multi_uses_nccl(const c3sc_multi *m) { return m && m->use_nccl ? 1 : 0; }
c3sc_problem *c3sc_multi_problem(c3sc_multi *m, int g) { return (m && g >= 0 && g < m->G) ? m->prob[g] : nullptr; }
const c3sc_valuef *c3sc_multi_valuef_get(const c3sc_multi_valuef *vf, int g) { return (vf && g >= 0 && g < (int)vf->vf.size()) ? vf->vf[g] : nullptr; }

int c3sc_multi_valuef_create(c3sc_multi *m, uint32_t d, const uint64_t *n, const uint64_t *ranks, const double *const *cores,
                             c3sc_multi_valuef **out)
{
    if (!m || !n || !ranks || !out) return c3sc_set_error(C3SC_EINVAL, "null argument");
    c3sc_multi_valuef *v = new c3sc_multi_valuef();
    v->m = m; v->d = d;
    v->vf.assign(m->G, nullptr);
    for (uint32_t k = 0; k < d; k++) v->len.push_back((size_t)(n[k] * ranks[k] * ranks[k + 1]));
    int rc = run_all(m, [&](int g) { return c3sc_valuef_create(d, n, ranks, nullptr, &v->vf[g]); });
    if (rc) { c3sc_multi_valuef_destroy(v); return rc; }
    if (cores) {
        rc = c3sc_multi_valuef_update(v, cores);
        if (rc) { c3sc_multi_valuef_destroy(v); return rc; }
    }
    *out = v;
    return C3SC_OK;
}

int c3sc_multi_valuef_update(c3sc_multi_valuef *v, const double *const *cores)
{
    if (!v || !cores) return c3sc_set_error(C3SC_EINVAL, "null argument");
    c3sc_multi *m = v->m;
    // host -> device 0 (c3sc_valuef_update also commits there)
    int rc = run_all(m, [&](int g) { return g == 0 ? c3sc_valuef_update(v->vf[0], cores) : (int)C3SC_OK; });
    if (rc || m->G == 1) return rc;
    std::vector<double *> buf(m->G);
    std::vector<size_t> cnt(m->G);
    for (int g = 0; g < m->G; g++) {
        rc = c3sc_valuef_device_buffer(v->vf[g], &buf[g], &cnt[g]);
        if (rc) return rc;
    }
    if (m->use_nccl) {
        // single-process group call: one ncclBroadcast per device between GroupStart / GroupEnd (root = device 0)
        m->nccl.GroupStart();
        int bad = 0;
        for (int g = 0; g < m->G; g++) {
            cudaSetDevice(m->dev[g]);
            bad |= m->nccl.Broadcast(buf[0], buf[g], cnt[0], ncclFloat64, 0, m->comm[g], m->stream[g]);
        }
        bad |= m->nccl.GroupEnd();
        if (bad) return c3sc_set_error(C3SC_ECUDA, "ncclBroadcast of the cores failed");
    }
    return run_all(m, [&](int g) {
        if (g == 0 && m->use_nccl) {
            if (cudaStreamSynchronize(m->stream[0]) != cudaSuccess) return (int)C3SC_ECUDA;
            return (int)C3SC_OK;
        }
        if (g == 0) return (int)C3SC_OK;
        if (!m->use_nccl && cudaMemcpyPeerAsync(buf[g], m->dev[g], buf[0], m->dev[0], cnt[0] * sizeof(double), m->stream[g]) != cudaSuccess)
            return c3sc_set_error(C3SC_ECUDA, "peer copy of the cores failed");
        int r = c3sc_valuef_commit(v->vf[g], m->stream[g]);
        if (r) return r;
        if (cudaStreamSynchronize(m->stream[g]) != cudaSuccess) return c3sc_set_error(C3SC_ECUDA, "broadcast stream failed");
        return (int)C3SC_OK;
    });
}

void c3sc_multi_valuef_destroy(c3sc_multi_valuef *v)
{
    if (!v) return;
    c3sc_multi *m = v->m;
    run_all(m, [&](int g) { if (v->vf[g]) c3sc_valuef_destroy(v->vf[g]); return 0; });
    delete v;
}

int c3sc_multi_vi_batch(c3sc_multi *m, const c3sc_multi_valuef *vf, size_t F, const int32_t *dim_vary, const int32_t *fixed_ind,
                        size_t ldo, double *value, int32_t *argmin)
{
    if (!m || !vf || !dim_vary || !fixed_ind || !value) return c3sc_set_error(C3SC_EINVAL, "null argument");
    const size_t d = m->dx;
    return run_all(m, [=](int g) {
        size_t b, e;
        c3sc_multi_shard(F, m->G, g, &b, &e);
        if (e <= b) return (int)C3SC_OK;
        return c3sc_vi_batch(m->prob[g], vf->vf[g], e - b, dim_vary + b, fixed_ind + b * d, ldo, value + b * ldo,
                             argmin ? argmin + b * ldo : nullptr);
    });
}

int c3sc_multi_pi_reset(c3sc_multi *m)
{
    if (!m) return c3sc_set_error(C3SC_EINVAL, "null argument");
    return run_all(m, [m](int g) {
        for (auto &slot : m->rows) if (slot[g].p) { cudaFree(slot[g].p); slot[g].p = nullptr; slot[g].cap = 0; }
        return 0;
    });
}

int c3sc_multi_pi_batch(c3sc_multi *m, const c3sc_multi_valuef *vf_policy, const c3sc_multi_valuef *vf_iter, uint32_t slot,
                        size_t F, const int32_t *dim_vary, const int32_t *fixed_ind, size_t ldo, int have_rows, double *value)
{
    if (!m || !vf_iter || !dim_vary || !fixed_ind || !value) return c3sc_set_error(C3SC_EINVAL, "null argument");
    if (!have_rows && !vf_policy) return c3sc_set_error(C3SC_EINVAL, "the improvement step needs the policy value function");
    if (slot >= m->rows.size()) return c3sc_set_error(C3SC_EINVAL, "slot out of range");
    const size_t d = m->dx, RW = 2 * d + 3;
    return run_all(m, [=](int g) {
        size_t b, e;
        c3sc_multi_shard(F, m->G, g, &b, &e);
        if (e <= b) return (int)C3SC_OK;
        c3sc_multi::Rows &R = m->rows[slot][g];
        const size_t need = (e - b) * ldo * RW * sizeof(double);
        if (have_rows && (!R.p || R.cap < need)) return c3sc_set_error(C3SC_EINVAL, "no resident rows in this slot for a batch of this size");
        if (!have_rows && R.cap < need) {
            if (R.p) cudaFree(R.p);
            R.p = nullptr; R.cap = 0;
            if (cudaMalloc(&R.p, need) != cudaSuccess) return c3sc_set_error(C3SC_ECUDA, "cudaMalloc of the resident policy rows failed");
            R.cap = need;
        }
        return c3sc_pi_batch_resident(m->prob[g], vf_policy ? vf_policy->vf[g] : nullptr, vf_iter->vf[g], e - b, dim_vary + b,
                                      fixed_ind + b * d, ldo, have_rows, (double *)R.p, value + b * ldo);
    });
}

size_t c3sc_multi_gathered_count(size_t F, int G, size_t ldo)
{
    const size_t per = G > 0 ? (F + (size_t)G - 1) / (size_t)G : F;
    return per * (size_t)(G > 0 ? G : 1) * ldo;
}

int c3sc_multi_vi_batch_gathered(c3sc_multi *m, const c3sc_multi_valuef *vf, size_t F, const int32_t *dim_vary,
                                 const int32_t *fixed_ind, size_t ldo, double *const *d_gathered)
{
    if (!m || !vf || !dim_vary || !fixed_ind || !d_gathered) return c3sc_set_error(C3SC_EINVAL, "null argument");
    const size_t d = m->dx, per = (F + (size_t)m->G - 1) / (size_t)m->G;
    // 1. every device backs up its block into its slot of its own gathered buffer
    int rc = run_all(m, [=](int g) {
        size_t b, e;
        c3sc_multi_shard(F, m->G, g, &b, &e);
        const size_t nf = e > b ? e - b : 0;
        if (nf * 4 > m->cap_dv[g]) { if (m->d_dv[g]) cudaFree(m->d_dv[g]); if (cudaMalloc(&m->d_dv[g], per * 4) != cudaSuccess) return (int)C3SC_ECUDA; m->cap_dv[g] = per * 4; }
        if (nf * d * 4 > m->cap_fi[g]) { if (m->d_fi[g]) cudaFree(m->d_fi[g]); if (cudaMalloc(&m->d_fi[g], per * d * 4) != cudaSuccess) return (int)C3SC_ECUDA; m->cap_fi[g] = per * d * 4; }
        cudaStream_t st = m->stream[g];
        if (nf) {
            if (cudaMemcpyAsync(m->d_dv[g], dim_vary + b, nf * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) return (int)C3SC_ECUDA;
            if (cudaMemcpyAsync(m->d_fi[g], fixed_ind + b * d, nf * d * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) return (int)C3SC_ECUDA;
            c3sc_batch_out o;
            memset(&o, 0, sizeof o);
            o.value = d_gathered[g] + (size_t)g * per * ldo;
            int r = c3sc_vi_batch_dev(m->prob[g], vf->vf[g], nf, (const int32_t *)m->d_dv[g], (const int32_t *)m->d_fi[g], ldo, &o, st);
            if (r) return r;
        }
        if (nf < per) cudaMemsetAsync(d_gathered[g] + ((size_t)g * per + nf) * ldo, 0, (per - nf) * ldo * sizeof(double), st);
        if (!m->use_nccl && cudaStreamSynchronize(st) != cudaSuccess) return (int)C3SC_ECUDA;
        return (int)C3SC_OK;
    });
    if (rc || m->G == 1) {
        if (!rc) rc = run_all(m, [=](int g) { return cudaStreamSynchronize(m->stream[g]) == cudaSuccess ? (int)C3SC_OK : (int)C3SC_ECUDA; });
        return rc;
    }
    // 2. all-gather in place
    if (m->use_nccl) {
        m->nccl.GroupStart();
        int bad = 0;
        for (int g = 0; g < m->G; g++) {
            cudaSetDevice(m->dev[g]);
            bad |= m->nccl.AllGather(d_gathered[g] + (size_t)g * per * ldo, d_gathered[g], per * ldo, ncclFloat64, m->comm[g], m->stream[g]);
        }
        bad |= m->nccl.GroupEnd();
        if (bad) return c3sc_set_error(C3SC_ECUDA, "ncclAllGather of the fiber values failed");
    }
    rc = run_all(m, [=](int g) {
        cudaStream_t st = m->stream[g];
        if (!m->use_nccl)
            for (int h = 0; h < m->G; h++) {
                if (h == g) continue;
                if (cudaMemcpyPeerAsync(d_gathered[g] + (size_t)h * per * ldo, m->dev[g], d_gathered[h] + (size_t)h * per * ldo, m->dev[h],
                                        per * ldo * sizeof(double), st) != cudaSuccess)
                    return c3sc_set_error(C3SC_ECUDA, "peer copy of the fiber values failed");
            }
        return cudaStreamSynchronize(st) == cudaSuccess ? (int)C3SC_OK : c3sc_set_error(C3SC_ECUDA, "gather stream failed");
    });
    if (rc) return rc;
    return run_all(m, [=](int g) { return c3sc_problem_check(m->prob[g]); });
}

// ---- the host cross driver over all devices ------------------------------------------------------------------------
struct mvi_ctx { c3sc_multi *m; const c3sc_multi_valuef *vf; };
static int mvi_cb(size_t F, const int32_t *dv, const int32_t *fi, size_t ldo, double *out, void *arg)
{
    mvi_ctx *x = (mvi_ctx *)arg;
    return c3sc_multi_vi_batch(x->m, x->vf, F, dv, fi, ldo, out, nullptr);
}
struct mpi_ctx { c3sc_multi *m; const c3sc_multi_valuef *pol, *iter; };
static int mpi_cb(size_t F, const int32_t *dv, const int32_t *fi, size_t ldo, double *out, void *arg)
{
    mpi_ctx *x = (mpi_ctx *)arg;
    // the index sets move between sweeps: every request is an improvement + evaluation; the rows it produces stay on
    // the devices (slot 0) and are never transferred
    return c3sc_multi_pi_batch(x->m, x->pol, x->iter, 0, F, dv, fi, ldo, 0, out);
}

int c3sc_cross_run_vi_multi(c3sc_cross *c, c3sc_multi *m, const c3sc_multi_valuef *vf, const c3sc_cross_opts *opts,
                            double *const *cores, uint64_t *nfibers, double *rel_change)
{
    if (!c || !m || !vf) return c3sc_set_error(C3SC_EINVAL, "null argument");
    mvi_ctx x = {m, vf};
    c3sc_cross_pin_buffers(c, 1);
    if (!c3sc_cross_uses_memo(opts)) return c3sc_cross_run(c, mvi_cb, &x, opts, cores, nfibers, rel_change);
    c3sc_fiber_memo *memo = nullptr;
    int rc = c3sc_fiber_memo_create(c3sc_cross_dim(c), mvi_cb, &x, &memo);
    if (rc == C3SC_OK) rc = c3sc_cross_run(c, c3sc_fiber_memo_call, memo, opts, cores, nfibers, rel_change);
    c3sc_fiber_memo_destroy(memo);
    return rc;
}

int c3sc_cross_run_pi_multi(c3sc_cross *c, c3sc_multi *m, const c3sc_multi_valuef *vf_policy, const c3sc_multi_valuef *vf_iter,
                            const c3sc_cross_opts *opts, double *const *cores, uint64_t *nfibers, double *rel_change)
{
    if (!c || !m || !vf_policy || !vf_iter) return c3sc_set_error(C3SC_EINVAL, "null argument");
    mpi_ctx x = {m, vf_policy, vf_iter};
    c3sc_cross_pin_buffers(c, 1);
    if (!c3sc_cross_uses_memo(opts)) return c3sc_cross_run(c, mpi_cb, &x, opts, cores, nfibers, rel_change);
    c3sc_fiber_memo *memo = nullptr;
    int rc = c3sc_fiber_memo_create(c3sc_cross_dim(c), mpi_cb, &x, &memo);
    if (rc == C3SC_OK) rc = c3sc_cross_run(c, c3sc_fiber_memo_call, memo, opts, cores, nfibers, rel_change);
    c3sc_fiber_memo_destroy(memo);
    return rc;
}

}  // extern "C"
