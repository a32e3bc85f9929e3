/* models.h -- host callbacks for the BASELINE.json problem definitions.
 * TEST INFRASTRUCTURE ONLY.  These play the role of the "user code" in the
 * reference's examples: the same functions are handed, as callbacks, both to
 * the oracle port and to the reference objects in oracle/_ref.            */
#ifndef C3SC_ORACLE_MODELS_H
#define C3SC_ORACLE_MODELS_H
#include <stddef.h>
#include "c3sc_oracle.h"
#ifdef __cplusplus
extern "C" {
#endif
/* ids are the values of enum c3sc_model in include/c3sc_b200.h */
enum { ORC_MODEL_LQGND = 1, ORC_MODEL_DOUBLE_INT = 2, ORC_MODEL_DUBINS = 3, ORC_MODEL_SKID5D = 4,
       ORC_MODEL_USER = 5 /* examples/user_model_vdp.cuh: the host callbacks of the example user model */ };

/* Select the model the callbacks below evaluate (they read file-static
 * state exactly like the reference examples read their static `dim`).
 * params may be NULL (example defaults).  Returns 0, or 1 if unknown.   */
int orc_model_select(int model, size_t dx, const double *params, size_t nparams);
int orc_model_dims(int model, size_t dx, size_t *du, size_t *dw);
orc_dyn_fn   orc_model_drift(void);
orc_dyn_fn   orc_model_diff(void);
orc_stage_fn orc_model_stage(void);
orc_bound_fn orc_model_boundcost(void);
orc_obs_fn   orc_model_obscost(void);
void        *orc_model_diff_arg(void);
#ifdef __cplusplus
}
#endif
#endif
