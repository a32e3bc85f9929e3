/* micro-benchmark of the host cross driver's dense kernels (QR, maxvol, twin-row marking, TT dot) at the
 * bench shape: unfolding 2000 x 20, d = 10, N = 100, r = 20.
 *   gcc -std=gnu99 -O3 -fopenmp tools/cross_linalg_bench.c -o build/cross_linalg_bench -lm && C3SC_HOST_THREADS=8 build/cross_linalg_bench [m n] */
#include "../c3sc_b200/csrc/host/c3sc_cross.c"
int c3sc_vi_batch(c3sc_problem *p, const c3sc_valuef *v, size_t F, const int32_t *a, const int32_t *b, size_t l, double *o, int32_t *g)
{ (void)p; (void)v; (void)F; (void)a; (void)b; (void)l; (void)o; (void)g; return 1; }
int c3sc_pi_batch(c3sc_problem *p, const c3sc_valuef *v, const c3sc_valuef *w, size_t F, const int32_t *a, const int32_t *b, size_t l,
                  int h, double *r, int32_t *g, double *o)
{ (void)p; (void)v; (void)w; (void)F; (void)a; (void)b; (void)l; (void)h; (void)r; (void)g; (void)o; return 1; }
int c3sc_valuef_create(uint32_t d, const uint64_t *n, const uint64_t *r, const double *const *c, c3sc_valuef **o)
{ (void)d; (void)n; (void)r; (void)c; (void)o; return 1; }
int c3sc_valuef_update(c3sc_valuef *v, const double *const *c) { (void)v; (void)c; return 1; }
void c3sc_valuef_destroy(c3sc_valuef *v) { (void)v; }
int c3sc_host_alloc(size_t b, void **p) { (void)b; (void)p; return 1; }
int c3sc_host_free(void *p) { (void)p; return 0; }

int main(int argc, char **argv)
{
    const size_t m = argc > 1 ? (size_t)atol(argv[1]) : 2000, n = argc > 2 ? (size_t)atol(argv[2]) : 20;
    la_ws *w = la_ws_create(m, n);
    double *A = malloc(m * n * 8);
    srand(1);
    for (size_t i = 0; i < m * n; i++) A[i] = rand() / (double)RAND_MAX - 0.5;
    double tp = 0, t0;
    const int reps = 200;
    unsigned long long h = 1469598103934665603ull;
    for (int rep = 0; rep < reps; rep++) {
        memcpy(w->Q, A, m * n * 8);
        t0 = now_s(); pivot_step(w, m, n); tp += now_s() - t0;
    }
    for (size_t i = 0; i < m * n; i++) { unsigned long long u; memcpy(&u, w->B + i, 8); h = (h ^ u) * 1099511628211ull; }
    for (size_t j = 0; j < n; j++) h = (h ^ w->P[j]) * 1099511628211ull;
    printf("unfolding %zux%zu, %d threads: twin rows + qr + maxvol %.3f ms; hash of (B, P) %016llx\n", m, n, w->threads, tp / reps * 1e3, h);
#ifdef LA_PROFILE
    for (int t = 0; t < w->threads; t++)
        printf("  thread %d: %.1f barriers per step, %.3f ms per step waiting in them\n", t, (double)g_bar_n[t] / reps, g_bar_wait[t] / reps * 1e3);
#endif
    /* TT dot at d = 10, N = 100, r = 20 */
    uint64_t nn[10], rr[11];
    double *cores[10], *wk = malloc(3 * 400 * 8);
    for (int k = 0; k <= 10; k++) rr[k] = (k == 0 || k == 10) ? 1 : 20;
    for (int k = 0; k < 10; k++) {
        nn[k] = 100;
        cores[k] = malloc(100 * rr[k] * rr[k + 1] * 8);
        for (size_t e = 0; e < 100 * rr[k] * rr[k + 1]; e++) cores[k][e] = rand() / (double)RAND_MAX - 0.5;
    }
    t0 = now_s();
    double acc = 0;
    for (int rep = 0; rep < 20; rep++) acc += tt_dot(10, nn, rr, cores, cores, wk, wk + 400, wk + 800);
    printf("tt_dot: %.3f ms (%.17g)\n", (now_s() - t0) / 20 * 1e3, acc);
    la_ws_free(w);
    return 0;
}
