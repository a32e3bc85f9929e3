/* multi_gpu_b200.c -- a plain C program drives every GPU of the box through include/c3sc_multi.h.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/multi_gpu_b200.c -Lc3sc_b200/lib -lc3sc_b200 \
 *       -Wl,-rpath,$PWD/c3sc_b200/lib -lm -o build/multi_gpu_b200
 *   build/multi_gpu_b200 [devices (0 = all)] [fibers] [d] [nodes per dim] [rank] [timed repetitions]
 *
 * A d-dimensional LQG problem (examples/lqgnd/lqgnd.c: chain of double integrators, {-1,0,1}^(d/2) control grid,
 * reflecting faces), a random value function and F random fibers.  Checks, bit for bit against ONE device:
 *   - bellman_vi sharded over G devices (c3sc_multi_vi_batch) == c3sc_vi_batch on device 0
 *   - the device-resident, all-gathered variant (ncclAllGather) on every device == the same numbers
 *   - bellman_pi with resident rows: improvement + sub-iteration == c3sc_pi_batch on device 0
 *   - one cross-approximation value-iteration step over all devices == the same step on one device
 * and prints the sharded throughput.  Exit code 0 and a final "OK ..." line on success.
 */
#define _POSIX_C_SOURCE 199309L
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <cuda_runtime_api.h>
#include "c3sc_multi.h"

static uint64_t sm_state;
static uint64_t splitmix(void)
{
    uint64_t z = (sm_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double u01(void) { return (double)(splitmix() >> 11) * (1.0 / 9007199254740992.0); }
static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc_ = (call);                                                             \
        if (rc_) { fprintf(stderr, "%s: %s\n", #call, c3sc_last_error()); return rc_ == C3SC_ENODEV ? 2 : 1; } \
    } while (0)

int main(int argc, char **argv)
{
    int ndev = argc > 1 ? atoi(argv[1]) : 0;
    const size_t F = argc > 2 ? (size_t)atol(argv[2]) : 4000;
    const uint32_t d = argc > 3 ? (uint32_t)atoi(argv[3]) : 6;
    const uint64_t N = argc > 4 ? (uint64_t)atol(argv[4]) : 24;
    const uint64_t r = argc > 5 ? (uint64_t)atol(argv[5]) : 6;
    const int reps = argc > 6 ? atoi(argv[6]) : 3;
    if (d < 2 || d > 12 || d % 2 || N < 4) { fprintf(stderr, "d must be even in [2, 12], N >= 4\n"); return 1; }
    if (c3sc_cuda_device_count() == 0) { fprintf(stderr, "no CUDA device: the Bellman backup has no CPU fallback\n"); return 2; }
    sm_state = 0xC35C0000u;

    /* grid, MCA constants (src/bellman.c:1975-1986,181-186), control grid */
    uint64_t ngrid[C3SC_MAXD], ranks[C3SC_MAXD + 1];
    double *xg[C3SC_MAXD], t[2 * C3SC_MAXD];
    int32_t bc[C3SC_MAXD];
    const double lb = -2.0, ub = 2.0, h = (ub - lb) / (double)(N - 1);
    for (uint32_t i = 0; i < d; i++) {
        ngrid[i] = N; bc[i] = C3SC_REFLECT;
        xg[i] = malloc(N * sizeof(double));
        xg[i][0] = lb;
        for (uint64_t j = 1; j < N; j++) xg[i][j] = xg[i][j - 1] + h;
        t[2 * i] = h; t[2 * i + 1] = 1.0;                        /* h2/h_i, h2/h_i^2 with equal steps */
    }
    const uint32_t du = d / 2;
    uint32_t nu = 1;
    for (uint32_t m = 0; m < du; m++) nu *= 3;
    double *controls = malloc((size_t)nu * du * sizeof(double));
    for (uint32_t c = 0; c < nu; c++) {
        uint32_t q = c;
        for (int m = (int)du - 1; m >= 0; m--) { controls[(size_t)c * du + m] = (double)(q % 3) - 1.0; q /= 3; }
    }
    c3sc_problem_desc desc;
    memset(&desc, 0, sizeof desc);
    desc.dx = d; desc.du = du; desc.dw = d; desc.ngrid = ngrid; desc.xgrid = (const double *const *)xg;
    desc.h2 = h * h; desc.t = t; desc.bc = bc; desc.discount = 0.1; desc.nu = nu; desc.controls = controls;
    desc.model = C3SC_MODEL_LQGND; desc.arith = C3SC_ARITH_FAST;

    /* random value functions and fibers */
    double *cores[C3SC_MAXD], *cores2[C3SC_MAXD];
    ranks[0] = ranks[d] = 1;
    for (uint32_t k = 1; k < d; k++) ranks[k] = r;
    for (uint32_t k = 0; k < d; k++) {
        const size_t len = N * ranks[k] * ranks[k + 1];
        cores[k] = malloc(len * sizeof(double)); cores2[k] = malloc(len * sizeof(double));
        for (size_t e = 0; e < len; e++) { cores[k][e] = (2.0 * u01() - 1.0) / sqrt((double)ranks[k]); cores2[k][e] = (2.0 * u01() - 1.0) / sqrt((double)ranks[k]); }
    }
    /* descriptors and results of the timed calls in page-locked memory (c3sc_host_alloc): PCIe speed, overlapped copies */
    int32_t *dv, *fi;
    CHECK(c3sc_host_alloc(F * sizeof(int32_t), (void **)&dv));
    CHECK(c3sc_host_alloc(F * d * sizeof(int32_t), (void **)&fi));
    for (size_t f = 0; f < F; f++) {
        dv[f] = (int32_t)(f % d);
        for (uint32_t i = 0; i < d; i++) {
            const double u = u01();
            fi[f * d + i] = u < 0.05 ? 0 : (u < 0.1 ? (int32_t)N - 1 : (int32_t)(u01() * (double)N) % (int32_t)N);
        }
        fi[f * d + dv[f]] = 0;
    }
    const size_t ldo = N, nval = F * ldo;

    /* one device: the reference numbers */
    CHECK(c3sc_cuda_init(0));
    c3sc_problem *p1; c3sc_valuef *v1, *v1b;
    CHECK(c3sc_problem_create(&desc, &p1));
    CHECK(c3sc_valuef_create(d, ngrid, ranks, (const double *const *)cores, &v1));
    CHECK(c3sc_valuef_create(d, ngrid, ranks, (const double *const *)cores2, &v1b));
    double *ref = malloc(nval * 8), *refp1 = malloc(nval * 8), *refp2 = malloc(nval * 8), *rows = malloc(nval * (2 * d + 3) * 8);
    CHECK(c3sc_vi_batch(p1, v1, F, dv, fi, ldo, ref, NULL));
    CHECK(c3sc_pi_batch(p1, v1, v1b, F, dv, fi, ldo, 0, rows, NULL, refp1));
    CHECK(c3sc_pi_batch(p1, NULL, v1, F, dv, fi, ldo, 1, rows, NULL, refp2));

    /* all devices */
    c3sc_multi *m; c3sc_multi_valuef *mv, *mvb;
    CHECK(c3sc_multi_create(&desc, ndev, NULL, &m));
    const int G = c3sc_multi_device_count(m);
    CHECK(c3sc_multi_valuef_create(m, d, ngrid, ranks, (const double *const *)cores2, &mv));     /* wrong numbers first ... */
    CHECK(c3sc_multi_valuef_update(mv, (const double *const *)cores));                           /* ... the update broadcasts the right ones */
    CHECK(c3sc_multi_valuef_create(m, d, ngrid, ranks, (const double *const *)cores2, &mvb));
    double *got;
    CHECK(c3sc_host_alloc(nval * 8, (void **)&got));
    memset(got, 0xff, nval * 8);
    CHECK(c3sc_multi_vi_batch(m, mv, F, dv, fi, ldo, got, NULL));
    if (memcmp(got, ref, nval * 8)) { fprintf(stderr, "sharded bellman_vi differs from one device\n"); return 1; }
    double best = 1e30;
    for (int it = 0; it < reps; it++) {
        const double t0 = now();
        CHECK(c3sc_multi_vi_batch(m, mv, F, dv, fi, ldo, got, NULL));
        const double dt = now() - t0;
        if (dt < best) best = dt;
    }
    /* policy iteration with resident rows */
    CHECK(c3sc_multi_pi_batch(m, mv, mvb, 3, F, dv, fi, ldo, 0, got));
    if (memcmp(got, refp1, nval * 8)) { fprintf(stderr, "sharded bellman_pi (improvement) differs from one device\n"); return 1; }
    CHECK(c3sc_multi_pi_batch(m, NULL, mv, 3, F, dv, fi, ldo, 1, got));
    if (memcmp(got, refp2, nval * 8)) { fprintf(stderr, "sharded bellman_pi (sub-iteration on resident rows) differs from one device\n"); return 1; }
    CHECK(c3sc_multi_pi_reset(m));
    if (c3sc_multi_pi_batch(m, NULL, mv, 3, F, dv, fi, ldo, 1, got) != C3SC_EINVAL) { fprintf(stderr, "sub-iteration without rows must be refused\n"); return 1; }
    /* gathered on every device */
    const size_t gc = c3sc_multi_gathered_count(F, G, ldo);
    double *dg[C3SC_MAXPEERS];
    for (int g = 0; g < G; g++) { cudaSetDevice(g); if (cudaMalloc((void **)&dg[g], gc * 8) != cudaSuccess) { fprintf(stderr, "cudaMalloc\n"); return 1; } }
    CHECK(c3sc_multi_vi_batch_gathered(m, mv, F, dv, fi, ldo, dg));
    for (int g = 0; g < G; g++) {
        cudaSetDevice(g);
        if (cudaMemcpy(got, dg[g], nval * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { fprintf(stderr, "cudaMemcpy\n"); return 1; }
        if (memcmp(got, ref, nval * 8)) { fprintf(stderr, "gathered values on device %d differ\n", g); return 1; }
        cudaFree(dg[g]);
    }
    cudaSetDevice(0);
    /* one cross-approximation value-iteration step: all devices vs one */
    c3sc_cross *ca, *cb;
    c3sc_cross_opts co = {2, 0.0, 0};
    CHECK(c3sc_cross_create(d, ngrid, ranks, &ca));
    CHECK(c3sc_cross_create(d, ngrid, ranks, &cb));
    double *oa[C3SC_MAXD], *ob[C3SC_MAXD];
    for (uint32_t k = 0; k < d; k++) { oa[k] = malloc(N * ranks[k] * ranks[k + 1] * 8); ob[k] = malloc(N * ranks[k] * ranks[k + 1] * 8); }
    uint64_t nfa = 0, nfb = 0;
    CHECK(c3sc_cross_run_vi(ca, p1, v1, &co, oa, &nfa, NULL));
    CHECK(c3sc_cross_run_vi_multi(cb, m, mv, &co, ob, &nfb, NULL));
    uint64_t ra[C3SC_MAXD + 1], rb[C3SC_MAXD + 1];
    c3sc_cross_ranks(ca, ra); c3sc_cross_ranks(cb, rb);
    int same = nfa == nfb;
    for (uint32_t k = 0; k <= d; k++) same = same && ra[k] == rb[k];
    for (uint32_t k = 0; k < d && same; k++) same = !memcmp(oa[k], ob[k], N * ra[k] * ra[k + 1] * 8);
    if (!same) { fprintf(stderr, "cross step over all devices differs from one device\n"); return 1; }
    printf("OK devices %d nccl %d fibers %zu nodes %zu sharded_vi_seconds %.6f node_backups_per_s %.4e cross_fibers %llu\n", G,
           c3sc_multi_uses_nccl(m), F, nval, best, (double)nval / best, (unsigned long long)nfb);
    c3sc_host_free(dv); c3sc_host_free(fi); c3sc_host_free(got);
    c3sc_cross_destroy(ca); c3sc_cross_destroy(cb);
    c3sc_multi_valuef_destroy(mv); c3sc_multi_valuef_destroy(mvb); c3sc_multi_destroy(m);
    c3sc_valuef_destroy(v1); c3sc_valuef_destroy(v1b); c3sc_problem_destroy(p1);
    return 0;
}
