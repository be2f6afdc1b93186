/* CUSTOMPROP functions of the reference's three example drivers, as C host callbacks with the
 * kfsp_propensity_fn signature (include/kfsp.h) -- what a C/Python host passes to
 * kfsp_model_set_custom_propensity() where the Fortran host sets MODEL%CUSTOMPROP =>.
 *
 *   toggle_propensity     examples/toggle.f90:60-74
 *   PROPENSITY            examples/repressilator.f90:50-69
 *   goutsias_propensity   examples/transcr6d.f90:63-89  (reaction 9 in INTEGER arithmetic)
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no fused multiply-add: each operation rounds once,
 * as in the Fortran expressions).  `reaction` is 1-based, `state` is species-fastest int32. */
#include <math.h>
#include <stdint.h>

double kfsp_example_toggle_propensity(const int32_t* state, int32_t reaction, const double* p, void* ctx) {
    (void)ctx;
    switch (reaction) {
    case 1: return p[0] + p[1] / (1.0 + pow((double)state[1], 1.5));
    case 2: return p[2] * state[0];
    case 3: return p[3] + p[4] / (1.0 + pow((double)state[0], 3.5));
    case 4: return p[5] * state[1];
    }
    return 0.0;
}

double kfsp_example_repressilator_propensity(const int32_t* state, int32_t reaction, const double* p, void* ctx) {
    (void)ctx;
    switch (reaction) {
    case 1: return p[0] / (1.0 + p[1] * pow((double)state[1], 6.0));
    case 2: return p[2] * state[0];
    case 3: return p[0] / (1.0 + p[1] * pow((double)state[2], 6.0));
    case 4: return p[2] * state[1];
    case 5: return p[0] / (1.0 + p[1] * pow((double)state[0], 6.0));
    case 6: return p[2] * state[2];
    }
    return 0.0;
}

/* species order M, D, RNA, DNA, DNA.D, DNA.2D  (examples/transcr6d.f90:14) */
double kfsp_example_goutsias_propensity(const int32_t* state, int32_t reaction, const double* p, void* ctx) {
    enum { M = 0, D = 1, RNA = 2, DNA = 3, DNAD = 4, DNA2D = 5 };
    (void)ctx;
    switch (reaction) {
    case 1: return p[0] * state[RNA];
    case 2: return p[1] * state[M];
    case 3: return p[2] * state[DNAD];
    case 4: return p[3] * state[RNA];
    case 5: return p[4] * state[DNA] * state[D];
    case 6: return p[5] * state[DNAD];
    case 7: return p[6] * state[DNAD] * state[D];
    case 8: return p[7] * state[DNA2D];
    case 9: return p[8] * (double)(state[M] * (state[M] - 1) / 2);
    case 10: return p[9] * state[D];
    }
    return 0.0;
}
