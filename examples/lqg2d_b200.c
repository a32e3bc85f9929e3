/* lqg2d_b200.c -- the solver part of an examples/lqg2d_new-style program, written against
 * include/c3sc_host.h: the reference's own call sequence (c3control_create, add_drift / diff / costs,
 * boundaries, c3opt brute-force control set, ApproxArgs, c3control_init_value, c3control_pi_solve +
 * c3control_vi_solve outer loop, c3control_controller), plus the ONE added line a port needs:
 * c3control_set_device_model next to the host callbacks.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/lqg2d_b200.c -Lc3sc_b200/lib -lc3sc_b200 \
 *       -Wl,-rpath,$PWD/c3sc_b200/lib -lm -o build/lqg2d_b200
 *   build/lqg2d_b200 [nodes per dimension] [outer iterations] [checkpoint file]
 *
 * With a checkpoint file the program resumes from it when it exists and saves the value function into it
 * at the end, the way the reference examples handle their cost.c3 (examples/dubinscar_new/dubinscar.c:324-354).
 *
 * Prints one line per outer iteration and a final "RESULT norm <nodal l2> u0 <control at (0.5,-0.5)>".
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "c3sc_host.h"

/* dynamics and costs of the 2-D LQG problem (double integrator with unit noise, quadratic stage cost) */
static int drift(double t, const double *x, const double *u, double *out, double *jac, void *args)
{
    (void)t; (void)args; (void)jac;
    out[0] = x[1];
    out[1] = u[0];
    return 0;
}
static int diffusion(double t, const double *x, const double *u, double *out, double *grad, void *args)
{
    (void)t; (void)x; (void)u; (void)args; (void)grad;
    out[0] = 1.0; out[1] = 0.0; out[2] = 0.0; out[3] = 1.0;
    return 0;
}
static int stagecost(double t, const double *x, const double *u, double *out, double *grad)
{
    (void)t; (void)grad;
    *out = x[0] * x[0] + x[1] * x[1] + u[0] * u[0];
    return 0;
}
static int boundcost(double t, const double *x, double *out) { (void)t; (void)x; *out = 100.0; return 0; }
static int obscost(const double *x, double *out) { (void)x; *out = 0.0; return 0; }
static int startcost(size_t N, const double *x, double *out, void *arg)
{
    (void)arg;
    for (size_t i = 0; i < N; i++) out[i] = x[2 * i] * x[2 * i] + x[2 * i + 1] * x[2 * i + 1];
    return 0;
}

int main(int argc, char **argv)
{
    const size_t n = argc > 1 ? (size_t)atoi(argv[1]) : 40;
    const size_t outer = argc > 2 ? (size_t)atoi(argv[2]) : 5;
    size_t dx = 2, du = 1, dw = 2, ngrid[2] = { n, n };
    double lb[2] = { -2.0, -2.0 }, ub[2] = { 2.0, 2.0 }, discount = 0.1;

    if (c3sc_cuda_init(0)) { fprintf(stderr, "lqg2d_b200: %s\n", c3sc_last_error()); return 2; }

    struct C3Control *c3c = c3control_create(dx, du, dw, lb, ub, ngrid, discount);
    c3control_add_drift(c3c, drift, NULL);
    c3control_add_diff(c3c, diffusion, NULL);
    c3control_add_stagecost(c3c, stagecost);
    c3control_add_boundcost(c3c, boundcost);
    c3control_add_obscost(c3c, obscost);
    c3control_set_external_boundary(c3c, 0, "reflect");
    c3control_set_external_boundary(c3c, 1, "reflect");
    const double model_params[4] = { 1.0, 1.0, 100.0, 0.0 };       /* diffusion diag, boundary cost, obstacle cost */
    c3control_set_device_model(c3c, C3SC_MODEL_LQGND, model_params, 4);      /* the added line */

    double uvals[3] = { -1.0, 0.0, 1.0 };
    struct c3Opt *opt = c3opt_alloc(BRUTEFORCE, du);
    c3opt_set_brute_force_vals(opt, 3, uvals);
    /* the registered device model against the host callbacks on 200 grid nodes: must agree exactly */
    if (dp_param_check_device_model(c3control_get_dp(c3c), c3control_get_mca(c3c), opt, 200, 0.0) != 0) {
        fprintf(stderr, "lqg2d_b200: the device model does not match the host callbacks\n");
        return 3;
    }

    struct ApproxArgs *aargs = approx_args_init();
    approx_args_set_cross_tol(aargs, 1e-8);
    approx_args_set_round_tol(aargs, 1e-7);
    approx_args_set_kickrank(aargs, 2);
    approx_args_set_startrank(aargs, 3);
    approx_args_set_maxrank(aargs, 12);
    approx_args_set_adapt(aargs, 1);

    char *checkpoint = argc > 3 ? argv[3] : NULL;
    struct ValueF *cost = checkpoint ? valuef_load(checkpoint, ngrid, c3control_get_xgrid(c3c)) : NULL;
    if (cost) printf("resumed from %s\n", checkpoint);
    else cost = c3control_init_value(c3c, startcost, NULL, aargs, 0);
    struct Diag *diag = NULL;
    for (size_t it = 0; it < outer; it++) {
        struct ValueF *next = c3control_pi_solve(c3c, 5, 1e-7, cost, aargs, opt, 0, &diag);
        struct ValueF *temp = c3control_vi_solve(c3c, 1, 1e-7, next, aargs, opt, 0, &diag);
        const double diff = valuef_norm2diff(next, temp);
        printf("outer %zu: norm %.10e diff %.3e ranks %zu\n", it, valuef_norm(temp), diff, valuef_get_ranks(temp)[1]);
        valuef_destroy(next);
        valuef_destroy(cost);
        cost = temp;
    }
    c3control_add_policy_sim(c3c, cost, opt, NULL);
    double x[2] = { 0.5, -0.5 }, u[1] = { 0.0 };
    if (c3control_controller(0.0, x, u, c3c)) return 4;
    printf("RESULT norm %.12e u0 %.3f\n", valuef_norm(cost), u[0]);
    if (checkpoint && valuef_save(cost, checkpoint)) fprintf(stderr, "lqg2d_b200: could not write %s\n", checkpoint);

    diag_destroy(&diag);
    valuef_destroy(cost);
    approx_args_free(aargs);
    c3opt_free(opt);
    c3control_destroy(c3c);
    return 0;
}
