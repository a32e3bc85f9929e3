/* models.c -- the four problem definitions behind BASELINE.json's configs,
 * restated from the reference examples (operation order kept).
 * TEST INFRASTRUCTURE ONLY.
 *
 * params layout (all optional, defaults = the example's constants):
 *   LQGND      [ss0, ss1, boundcost, obscost]          lqgnd.c:80-186, lqg2d.c:71-142
 *   DOUBLE_INT [ss0, ss1, boundcost, obscost]          double_int.c:80-162
 *   DUBINS     [s_xy, s_theta, stage, boundcost, obscost]  dubinscar.c:40-121
 *   SKID5D     [obscost]                               scar.c:39-176
 */
#include <math.h>
#include <string.h>
#include "models.h"

static int    g_model = 0;
static size_t g_dim = 0;
static double g_par[8];

static const double k_defaults[6][8] = {
    {0},
    {1.0, 1.0, 100.0, 0.0},          /* lqgnd.c:236 ss={1,1}; :164 boundcost 100; :175 ocost 0 */
    {1.0, 1.0, 1000.0, 0.0},         /* double_int.c:147 boundcost 1000 */
    {1.0, 1e-2, 1.0, 10.0, 0.0},     /* dubinscar.c:67-68, :93, :110, :121 */
    {0.0},
    {1.0, 0.5, 0.5, 50.0, 0.0},      /* user model (Van der Pol): mu, s0, s1, boundcost, obscost */
};

int orc_model_dims(int model, size_t dx, size_t *du, size_t *dw)
{
    switch (model) {
    case ORC_MODEL_LQGND:      if (dx % 2) return 1; *du = dx / 2; *dw = dx; return 0;  /* lqgnd.c:296-298 */
    case ORC_MODEL_DOUBLE_INT: *du = 1; *dw = dx; return 0;                             /* double_int.c:259-261 */
    case ORC_MODEL_DUBINS:     if (dx != 3) return 1; *du = 1; *dw = 3; return 0;
    case ORC_MODEL_SKID5D:     if (dx != 5) return 1; *du = 1; *dw = 5; return 0;
    case ORC_MODEL_USER:       if (dx != 2) return 1; *du = 1; *dw = 2; return 0;
    }
    return 1;
}

int orc_model_select(int model, size_t dx, const double *params, size_t nparams)
{
    size_t du, dw;
    if (orc_model_dims(model, dx, &du, &dw)) return 1;
    g_model = model;
    g_dim = dx;
    memcpy(g_par, k_defaults[model], sizeof g_par);
    for (size_t i = 0; i < nparams && i < 8; i++) g_par[i] = params[i];
    return 0;
}

/* ---- drift --------------------------------------------------------------- */
static int drift_cb(double t, const double *x, const double *u, double *out, double *jac, void *a)
{
    (void)t; (void)jac; (void)a;
    const size_t d = g_dim;
    switch (g_model) {
    case ORC_MODEL_LQGND: {                 /* lqgnd.c:86-95: chain of double integrators */
        size_t c = 0;
        for (size_t i = 0; i < d; i++) {
            if (i % 2 == 0) out[i] = x[i + 1];
            else out[i] = u[c++];
        }
        return 0;
    }
    case ORC_MODEL_DOUBLE_INT:              /* double_int.c:86-89 */
        for (size_t i = 0; i + 1 < d; i++) out[i] = x[i + 1];
        out[d - 1] = u[0];
        return 0;
    case ORC_MODEL_DUBINS:                  /* dubinscar.c:48-50 */
        out[0] = cos(x[2]);
        out[1] = sin(x[2]);
        out[2] = u[0];
        return 0;
    case ORC_MODEL_SKID5D: {                /* scar.c:61-87, order = {0,1,2,3,4} */
        double orient = x[2], angvel = x[3], speed = x[4], steering = u[0];
        double m = 1460.0, cf = 17000.0, ct = 20000.0, a1 = 1.2, b1 = 1.5, In = 2170.0, s = 27.0;
        double co = cos(orient), so = sin(orient);
        double ff = cf * ((speed + a1 * angvel) / s + steering);
        double ft = ct * (speed - b1 * angvel) / s;
        out[0] = s * co - speed * so;
        out[1] = s * so + speed * co;
        out[2] = angvel;
        out[3] = (a1 * ff - b1 * ft) / In;
        out[4] = -s * angvel + (ff + ft) / m;
        return 0;
    }
    case ORC_MODEL_USER:                    /* what a user's host drift callback looks like: controlled Van der Pol */
        out[0] = x[1];
        out[1] = g_par[0] * (1.0 - x[0] * x[0]) * x[1] - x[0] + u[0];
        return 0;
    }
    return 1;
}

/* ---- diffusion (dx x dw column-major; only the diagonal is consumed,
 *      nodeutil.c:294) ------------------------------------------------------ */
static int diff_cb(double t, const double *x, const double *u, double *out, double *grad, void *a)
{
    (void)t; (void)x; (void)u; (void)grad; (void)a;
    const size_t d = g_dim;
    for (size_t i = 0; i < d * d; i++) out[i] = 0.0;
    switch (g_model) {
    case ORC_MODEL_LQGND:                   /* lqgnd.c:122-129 */
        for (size_t i = 0; i < d; i++) out[i * d + i] = (i % 2 == 0) ? g_par[0] : g_par[1];
        return 0;
    case ORC_MODEL_DOUBLE_INT:              /* double_int.c:113-117 */
        for (size_t i = 0; i + 1 < d; i++) out[i * d + i] = g_par[0];
        out[(d - 1) * d + (d - 1)] = g_par[1];
        return 0;
    case ORC_MODEL_DUBINS:                  /* dubinscar.c:67-71 */
        out[0] = g_par[0]; out[4] = g_par[0]; out[8] = g_par[1];
        return 0;
    case ORC_MODEL_SKID5D:                  /* scar.c:118-130.  The example stores its 4th
                                               entry at out[28] (outside the 5x5 block), so the
                                               diagonal seen by transition_assemble is
                                               (1e-5,1e-5,1e-5,0,1e-5); we keep that effect
                                               without the out-of-bounds store. */
        out[0] = 1e-5; out[6] = 1e-5; out[12] = 1e-5; out[24] = 1e-5;
        return 0;
    case ORC_MODEL_USER:
        out[0] = g_par[1]; out[3] = g_par[2];
        return 0;
    }
    return 1;
}

/* ---- stage cost ------------------------------------------------------------ */
static int stage_cb(double t, const double *x, const double *u, double *out, double *grad)
{
    (void)t; (void)grad;
    const size_t d = g_dim;
    switch (g_model) {
    case ORC_MODEL_LQGND: {                 /* lqgnd.c:146-159 (== lqg2d.c:121 for d=2) */
        double g = 0.0;
        for (size_t i = 0; i < d; i++) g += x[i] * x[i];
        for (size_t i = 0; i < d / 2; i++) g += u[i] * u[i];
        *out = g;
        return 0;
    }
    case ORC_MODEL_DOUBLE_INT: *out = 1.0; return 0;          /* double_int.c:133 */
    case ORC_MODEL_DUBINS:     *out = g_par[2]; return 0;     /* dubinscar.c:93 */
    case ORC_MODEL_SKID5D: {                /* scar.c:149-150; pow(.,2) == exact square */
        double g = 1.0 + 0.02 * (x[0] * x[0]) + 0.02 * (x[1] * x[1]);
        g = g + x[3] * x[3] + x[4] * x[4];
        *out = g;
        return 0;
    }
    case ORC_MODEL_USER: *out = x[0] * x[0] + x[1] * x[1] + u[0] * u[0]; return 0;
    }
    return 1;
}

static int bound_cb(double t, const double *x, double *out)
{
    (void)t;
    switch (g_model) {
    case ORC_MODEL_LQGND:      *out = g_par[2]; return 0;
    case ORC_MODEL_DOUBLE_INT: *out = g_par[2]; return 0;
    case ORC_MODEL_DUBINS:     *out = g_par[3]; return 0;
    case ORC_MODEL_SKID5D: {                /* scar.c:165-166 */
        double g = 0.1 * (x[0] * x[0]) + 0.1 * (x[1] * x[1]);
        g = g + 0.1 * (x[3] * x[3]) + 0.1 * (x[4] * x[4]);
        *out = g;
        return 0;
    }
    case ORC_MODEL_USER:       *out = g_par[3]; return 0;
    }
    return 1;
}

static int obs_cb(const double *x, double *out)
{
    (void)x;
    switch (g_model) {
    case ORC_MODEL_LQGND:      *out = g_par[3]; return 0;
    case ORC_MODEL_DOUBLE_INT: *out = g_par[3]; return 0;
    case ORC_MODEL_DUBINS:     *out = g_par[4]; return 0;
    case ORC_MODEL_SKID5D:     *out = g_par[0]; return 0;
    case ORC_MODEL_USER:       *out = g_par[4]; return 0;
    }
    return 1;
}

orc_dyn_fn   orc_model_drift(void)     { return drift_cb; }
orc_dyn_fn   orc_model_diff(void)      { return diff_cb; }
orc_stage_fn orc_model_stage(void)     { return stage_cb; }
orc_bound_fn orc_model_boundcost(void) { return bound_cb; }
orc_obs_fn   orc_model_obscost(void)   { return obs_cb; }
void        *orc_model_diff_arg(void)  { return g_par; }
