/* micro-benchmark of the host cross driver's dense kernels (QR, maxvol, twin-row marking, TT dot) at the
 * bench shape: unfolding 2000 x 20, d = 10, N = 100, r = 20.
 *   gcc -std=c99 -O3 tools/cross_linalg_bench.c -o build/cross_linalg_bench -lm && build/cross_linalg_bench */
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

int main(void)
{
    const size_t m = 2000, n = 20;
    double *A = malloc(m * n * 8), *Q = malloc(m * n * 8), *B = malloc(m * n * 8), *work = malloc((n * n * 3 + n + m + 16) * 8);
    size_t P[64];
    char *skip = calloc(m, 1);
    srand(1);
    for (size_t i = 0; i < m * n; i++) A[i] = rand() / (double)RAND_MAX - 0.5;
    double tq = 0, tm = 0, tw = 0, t0;
    const int reps = 50;
    for (int rep = 0; rep < reps; rep++) {
        memcpy(Q, A, m * n * 8);
        t0 = now_s(); mark_twin_rows(Q, m, n, skip); tw += now_s() - t0;
        t0 = now_s(); qr_explicit_q(Q, m, n, work); tq += now_s() - t0;
        t0 = now_s(); maxvol(Q, m, n, skip, P, B, work); tm += now_s() - t0;
    }
    printf("unfolding %zux%zu: twin rows %.3f ms, qr %.3f ms, maxvol %.3f ms\n", m, n, tw / reps * 1e3, tq / reps * 1e3, tm / reps * 1e3);
    /* TT dot at d = 10, N = 100, r = 20 */
    uint64_t nn[10], rr[11];
    double *cores[10], *w = malloc(3 * 400 * 8);
    for (int k = 0; k <= 10; k++) rr[k] = (k == 0 || k == 10) ? 1 : 20;
    for (int k = 0; k < 10; k++) {
        nn[k] = 100;
        cores[k] = malloc(100 * rr[k] * rr[k + 1] * 8);
        for (size_t e = 0; e < 100 * rr[k] * rr[k + 1]; e++) cores[k][e] = rand() / (double)RAND_MAX - 0.5;
    }
    t0 = now_s();
    double acc = 0;
    for (int rep = 0; rep < 20; rep++) acc += tt_dot(10, nn, rr, cores, cores, w, w + 400, w + 800);
    printf("tt_dot: %.3f ms (%g)\n", (now_s() - t0) / 20 * 1e3, acc);
    return 0;
}
