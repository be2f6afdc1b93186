/* The call sequence of the reference-side binding, statement by statement, from a compiled C host.
 *
 * krylovfspssa_b200/fortran/kfsp_c_binding.f90 (the drop-in body of CME_SOLVE, KrylovSolver.f90:7-36) cannot be compiled in
 * this image (no Fortran compiler).  This program is its C twin: the same entry points in the same order with the same
 * argument conventions --
 *   * STOICHIOMETRY(NSPECIES,NREACTIONS) handed over column-major as it lies in memory,
 *   * propensities either as the parser's byte code (MODEL%GET_BYTECODE) or as a CUSTOMPROP trampoline that finds its model
 *     through a module-level pointer (ctx = NULL),
 *   * FSP_OUT%STATE(NSPECIES,1:MAX_SIZE) passed as BOTH states_in and states_out (the reference's FSP_OUT is in/out),
 *   * FSP_IN%VECTOR / FSP_OUT%VECTOR as separate MAX_SIZE arrays, n_in = FSP_OUT%SIZE.
 * It includes include/kfsp.h as C (not C++), so the header's C-cleanliness is checked by the compiler too.
 *
 *   shim_sequence <bytecode|custom> <model.input> <t> <fsptol> <exptol> <max_size> <seed> <out.bin> x0_1 .. x0_S  p_1 .. p_P
 * writes: int64 n, int32 iflag, int32 nstep, int32 states[n*S], double w[n].
 * Test infrastructure (tests/test_gpu_shim_sequence.py); links libkfsp.so only. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kfsp.h"

/* ---- the Fortran side's MODEL object, as far as the shim reads it */
typedef struct {
    int32_t nspecies, nreactions, nparameters;
    int32_t* stoichiometry; /* (nspecies, nreactions) column-major */
    double* parameter_val;
    kfsp_model parser; /* stands for PROPPARSER(R): source of the byte code GET_BYTECODE returns */
    double (*customprop)(const int* state, int nspecies, int reaction, const double* parameters);
} host_model;

static const host_model* kfsp_active_model = NULL; /* KFSP_ACTIVE_MODEL of the shim */

/* the "Fortran" CUSTOMPROP: assumed-shape arguments, default INTEGER state -- examples/toggle.f90:55-69 with the byte-code
 * model's own expressions so that both modes of this program must print identical bits */
static double toggle_customprop(const int* state, int nspecies, int reaction, const double* p) {
    extern double pow(double, double);
    (void)nspecies;
    switch (reaction) {
    case 1: return p[0] + p[1] / (2.0 + 0.2 * pow((double)state[1], 2.0));
    case 2: return p[2] * (double)state[0];
    case 3: return p[3] + p[4] / (1.0 + 0.5 * pow((double)state[0], 1.5));
    case 4: return p[5] * (double)state[1];
    }
    return 0.0;
}

/* KFSP_CUSTOMPROP_TRAMPOLINE of the shim */
static double customprop_trampoline(const int32_t* state, int32_t reaction, const double* params, void* ctx) {
    int st[64];                 /* automatic array ST(NSPECIES) of the shim */
    (void)ctx;
    (void)params;
    for (int s = 0; s < kfsp_active_model->nspecies; ++s) st[s] = (int)state[s];
    return kfsp_active_model->customprop(st, kfsp_active_model->nspecies, (int)reaction, kfsp_active_model->parameter_val);
}

#define CHECK(call)                                                          \
    do {                                                                     \
        int rc_ = (call);                                                    \
        if (rc_ != 0) {                                                      \
            fprintf(stderr, "%s -> %d (%s)\n", #call, rc_, kfsp_status_string(rc_)); \
            return 2;                                                        \
        }                                                                    \
    } while (0)

/* CME_SOLVE(MODEL, T, FSP_IN, FSP_OUT, FSPTOL, EXP_TOL, VERBOSITY) as the shim writes it */
static int cme_solve(const host_model* model, double t, const double* fsp_in_vector, int64_t* fsp_out_size, int32_t* fsp_out_state,
                     double* fsp_out_vector, int64_t fsp_out_max_size, double fsptol, double exp_tol, int64_t seed, kfsp_stats* stats) {
    kfsp_options opts;
    kfsp_handle h;
    kfsp_model m;
    int32_t code[1024];
    double immed[256];
    int64_t n_out = 0;
    CHECK(kfsp_default_options(&opts));
    opts.max_states = fsp_out_max_size;
    opts.seed = seed;
    CHECK(kfsp_model_create(model->nspecies, model->nreactions, model->nparameters, &m));
    CHECK(kfsp_model_set_stoichiometry(m, model->stoichiometry));
    CHECK(kfsp_model_reset_parameters(m, model->parameter_val, model->nparameters));
    if (model->customprop) {
        kfsp_active_model = model;
        CHECK(kfsp_model_set_custom_propensity(m, customprop_trampoline, NULL));
    } else {
        for (int r = 1; r <= model->nreactions; ++r) {
            int32_t ncode = 1024, nimmed = 256;
            CHECK(kfsp_model_get_propensity_bytecode(model->parser, r, code, &ncode, immed, &nimmed)); /* MODEL%GET_BYTECODE */
            CHECK(kfsp_model_set_propensity_bytecode(m, r, code, ncode, immed, nimmed));
        }
    }
    CHECK(kfsp_create(&opts, &h));
    CHECK(kfsp_set_model(h, m));
    {
        const int rc = kfsp_solve(h, t, *fsp_out_size, fsp_out_state, fsp_in_vector, fsptol, exp_tol, 0, &n_out, fsp_out_state, fsp_out_vector,
                                  fsp_out_max_size, stats);
        if (rc < 0) {
            fprintf(stderr, "KFSP_SOLVE FAILED, STATUS = %d (%s)\n", rc, kfsp_status_string(rc));
            return 2;
        }
    }
    *fsp_out_size = n_out;
    CHECK(kfsp_destroy(h));
    CHECK(kfsp_model_free(m));
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 9) {
        fprintf(stderr, "usage: shim_sequence <bytecode|custom> model.input t fsptol exptol max_size seed out.bin x0.. params..\n");
        return 1;
    }
    const int custom = strcmp(argv[1], "custom") == 0;
    host_model model;
    memset(&model, 0, sizeof model);
    if (kfsp_model_load(argv[2], &model.parser) != 0) { fprintf(stderr, "cannot load %s\n", argv[2]); return 1; }
    kfsp_model_dims(model.parser, &model.nspecies, &model.nreactions, &model.nparameters);
    const int S = model.nspecies, R = model.nreactions, P = model.nparameters;
    if (argc != 9 + S + P) { fprintf(stderr, "expected %d initial counts and %d parameters\n", S, P); return 1; }
    const double t = atof(argv[3]), fsptol = atof(argv[4]), exptol = atof(argv[5]);
    const int64_t max_size = atoll(argv[6]), seed = atoll(argv[7]);
    model.stoichiometry = (int32_t*)malloc(sizeof(int32_t) * S * R);
    model.parameter_val = (double*)malloc(sizeof(double) * (P > 0 ? P : 1));
    kfsp_model_get_stoichiometry(model.parser, model.stoichiometry);
    for (int i = 0; i < P; ++i) model.parameter_val[i] = atof(argv[9 + S + i]);
    model.customprop = custom ? toggle_customprop : NULL;

    /* CALL FSP_IN%CREATE(MODEL, MAX_SIZE); CALL FSP%CREATE(...); FSP%SIZE = 1; FSP%STATE(:,1) = X0; FSP_IN%VECTOR(1) = 1 */
    int32_t* fsp_state = (int32_t*)calloc((size_t)max_size * S, sizeof(int32_t));
    double* fsp_in_vector = (double*)calloc((size_t)max_size, sizeof(double));
    double* fsp_vector = (double*)calloc((size_t)max_size, sizeof(double));
    int64_t fsp_size = 1;
    for (int s = 0; s < S; ++s) fsp_state[s] = atoi(argv[9 + s]);
    fsp_in_vector[0] = 1.0;
    kfsp_stats stats;
    memset(&stats, 0, sizeof stats);
    const int rc = cme_solve(&model, t, fsp_in_vector, &fsp_size, fsp_state, fsp_vector, max_size, fsptol, exptol, seed, &stats);
    if (rc != 0) return rc;

    FILE* f = fopen(argv[8], "wb");
    if (!f) return 1;
    const int32_t iflag = stats.iflag, nstep = stats.nstep;
    fwrite(&fsp_size, sizeof fsp_size, 1, f);
    fwrite(&iflag, sizeof iflag, 1, f);
    fwrite(&nstep, sizeof nstep, 1, f);
    fwrite(fsp_state, sizeof(int32_t), (size_t)fsp_size * S, f);
    fwrite(fsp_vector, sizeof(double), (size_t)fsp_size, f);
    fclose(f);
    printf("shim_sequence %s: N=%lld steps=%d iflag=%d\n", argv[1], (long long)fsp_size, nstep, iflag);
    kfsp_model_free(model.parser);
    free(model.stoichiometry); free(model.parameter_val); free(fsp_state); free(fsp_in_vector); free(fsp_vector);
    return 0;
}
