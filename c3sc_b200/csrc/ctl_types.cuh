// ctl_types.cuh -- argument structs of stage 2 (control_kernel.cuh), shared with the node kernel of stage 1, which calls the
// fused walk across translation units.
#pragma once
#include "dev_types.h"

namespace c3sc {

constexpr int CT_NT = 256;
constexpr int CT_NUDMAX = 6;       // control dimensions of a grid-structured table (3^6 = 729 candidates)
constexpr int CT_NGMAX = 32;       // candidate groups (distinct normaliser shares) handled by the grouped walk

struct CtlArgs {
    DevProblem P;
    int F;                    // fibers of this chunk
    const int *dim_vary;      // [F]
    const int *fixed_ind;     // [F*dx]
    int ldo;
    long long NS;             // F*ldo
    const double *cst;        // [(2dx+1)*NS] slot-major neighbour values
    const signed char *flag;  // [NS]
    const int *act;           // [NS] non-absorbed node ids
    const int *act_count;
    int parts_log2;           // candidate chunks per node = 1 << parts_log2 (<= 32)
    // separable models, FAST: candidates regrouped by their share of the normaliser (DevProblem::gtab)
    int ng;                   // number of groups, 0 = plain table walk
    int gstart[CT_NGMAX + 1]; // first grouped position of every group
    double gA[CT_NGMAX];      // the group's normaliser share
    // ... and, when the control table is a full {lo, 0, hi}^NUD grid in C order (last control fastest), its
    // per-control description for the shared-prefix walk (k_control_grid): candidate (k_0..k_{NUD-1}) adds
    // gWlo[m]*cost_left(ud m) if k_m = 0, nothing if k_m = 1, gWhi[m]*cost_right if k_m = 2; its normaliser
    // share and h2*stage_u depend on the number of non-zero controls only (gAg, gHg).
    int grid_on;
    double gWlo[CT_NUDMAX], gWhi[CT_NUDMAX], gAg[CT_NUDMAX + 1], gHg[CT_NUDMAX + 1];
    double *value;            // outputs, any may be NULL
    int *argmin;
    double *rows;
    const double *rows_in;    // k_pi_eval
    // fused all-gather: values also go to every peer's gathered buffer (peer-mapped memory over NVLink)
    double *vpeer[C3SC_MAXPEERS];
    int npeer;
    long long peer_off;       // element offset of this chunk inside a gathered buffer
};

struct FusedCta {
    const double *reg;         // reg[slot*RN + g*njp + (j - jb)], slot 2dx = the node's own value
    int RN, njp;               // RN = FT_FBMAX * njp doubles per slot
    int nf, jb, je, k, nmax;   // fibers of the group, node range (jb == 0, je == ngrid[k]: the along-fiber neighbours are the
                               // own values of other nodes of the same fiber), varying dimension, row stride of sAbs
    const int *sFid;           // [nf] fiber ids relative to the chunk
    const signed char *sAbs;   // [g*nmax + j] flags 0 / 1 / -1
    int pi_eval, arg;          // policy evaluation against c.rows_in; argmin / rows wanted
};

// one dispatcher per model family (defined in inst_lqg_lo.cu / inst_lqg_hi.cu / inst_misc.cu, called from k_ft_nodes:
// relocatable device code).  family: 0 = LQG dx <= 6, 1 = LQG dx >= 8, 2 = the other models
__device__ void fused_walk_lqg_lo(int dx, const CtlArgs &c, const FusedCta &w);
__device__ void fused_walk_lqg_hi(int dx, const CtlArgs &c, const FusedCta &w);
__device__ void fused_walk_misc(int model, int dx, const CtlArgs &c, const FusedCta &w);

}  // namespace c3sc
