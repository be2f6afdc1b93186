/* kfsp.h -- C ABI of libkfsp.so, the B200 (sm_100a) implementation of the Krylov-FSP-SSA
 * time-stepping hot path of voduchuy/KrylovFspSsa.
 *
 * The reference has no FFI; its narrowest usable seam is the module procedure
 *   DGEXPV_FSP(MODEL,T,V,FSP,W,FSPTOL,KRYTOL,ITRACE,IFLAG)   src/fsp/KrylovSolver.f90:40
 * reached through the public
 *   CME_SOLVE(MODEL,T,FSP_IN,FSP_OUT,FSPTOL,EXP_TOL,VERBOSITY) src/fsp/KrylovSolver.f90:7
 * (SURVEY.md 8b).  kfsp_solve() replaces the body of CME_SOLVE; the other entry points
 * expose the pieces DGEXPV_FSP is made of, one per reference routine, so that each can be
 * parity-tested on its own.  INTEGRATION.md shows the ISO_C_BINDING interface block a
 * maintainer of the Fortran host adds (fortran/kfsp_c_binding.f90).
 *
 * Conventions
 *   - plain pointers and sizes only; every array is caller-owned HOST memory unless the
 *     name says _device; the library copies to/from the GPU itself;
 *   - STATE arrays are int32, species fastest: state[s + S*i]  == FSP%STATE(s+1,i+1)
 *     (src/state_space/StateSpace.f90:22);
 *   - STOICHIOMETRY is int32, species fastest: stoich[s + S*k] == MODEL%STOICHIOMETRY(s+1,k+1)
 *     (src/model/ModelModule.f90:24-25);
 *   - reactions and state indices are 1-based in every exported array, exactly as the
 *     Fortran host sees them (ADJ: >0 index, 0 not in the projection, -1 illegal);
 *   - every function returns a kfsp_status; nothing aborts the process (the reference STOPs).
 *   - there is no CPU fallback: without a CUDA device kfsp_create() fails with
 *     KFSP_ERR_NO_DEVICE.
 */
#ifndef KFSP_H
#define KFSP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kfsp_handle_s* kfsp_handle;   /* device solver + device state space   */
typedef struct kfsp_model_s* kfsp_model;     /* host-side CME_MODEL mirror            */

typedef enum {
    KFSP_OK = 0,
    KFSP_IFLAG_MXSTEP = 1,            /* IFLAG=1: MXSTEP reached        (KrylovSolver.f90:551)      */
    KFSP_IFLAG_TOLERANCE = 2,         /* IFLAG=2: tolerance too high    (KrylovSolver.f90:396)      */
    KFSP_ERR_BAD_SIZES = -3,          /* IFLAG=-1..-3                   (KrylovSolver.f90:142-149)  */
    KFSP_ERR_NULL_H = -4,             /* 'null H in input of DGPADM'    (dgpadm.f:254)              */
    KFSP_ERR_SINGULAR = -5,           /* 'Problem in DGESV'             (dgpadm.f:316)              */
    KFSP_ERR_OVERFLOW = -10,          /* 'FSP SIZE EXCEEDS MEMORY LIMIT' (StateSpace.f90:388-391)   */
    KFSP_ERR_BAD_STATE = -11,         /* negative / over-limit / duplicate initial state            */
    KFSP_ERR_MOLECULE_LIMIT = -12,    /* a count reached MAXNUMBERMOLECULES (key aliasing in the reference) */
    KFSP_ERR_NO_DEVICE = -20,
    KFSP_ERR_CUDA = -21,
    KFSP_ERR_ARG = -22,
    KFSP_ERR_NO_MODEL = -23,
    KFSP_ERR_PARSE = -24,
    KFSP_ERR_IO = -25,
    KFSP_ERR_UNSUPPORTED = -26,
    KFSP_ERR_OUT_TOO_SMALL = -27,
    KFSP_ERR_NCCL = -28,
    KFSP_ERR_SSA_RUNAWAY = -29
} kfsp_status;

/* Every tunable the reference hard-codes, with the reference's value as default. */
typedef struct {
    int32_t m_max;             /* 100      KrylovSolver.f90:47                                   */
    int32_t m_min;             /* 10       KrylovSolver.f90:47                                   */
    int32_t ideg;              /* 6        KrylovSolver.f90:82  (only 6 is implemented)          */
    int32_t n_init_onestep;    /* 5        KrylovSolver.f90:132                                  */
    int32_t fsp_reject_limit;  /* 5        KrylovSolver.f90:466                                  */
    int32_t mxstep;            /* 0        KrylovSolver.f90:77                                   */
    int32_t mxreject;          /* 0        KrylovSolver.f90:79                                   */
    int32_t enable_drop;       /* 1        DROP_STATES on (0 = fixed state set)                  */
    int32_t enable_expand;     /* 1        SSA_EXTENDER + ONESTEP_EXTENDER on                    */
    int32_t max_molecules;     /* 10000    StateSpace.f90:11                                     */
    int32_t device;            /* CUDA device ordinal; -1 = current device                       */
    int32_t spmv_variant;      /* 0 explicit gather-ELL (reference data model); 1 matrix-free on a lattice:
                                  the state set must be a full box in natural order (first species fastest),
                                  every propensity must read at most one species, the state set is fixed;
                                  2 index-only: any (adaptive, irregular, partitioned) state set; OFFDIAG's gather
                                  form is never stored, a_k(x - nu_k) is recomputed from the row's integer state
                                  (4R+4S+24 instead of 12R+24 bytes per state and SpMV); every propensity must be
                                  sums/products of sub-expressions that read one species each (mass action, Hill
                                  terms): kfsp_set_model returns KFSP_ERR_UNSUPPORTED otherwise                */
    int64_t max_states;        /* 6291469  NMAX, StateSpace.f90:10                               */
    double delta;              /* 1.2      KrylovSolver.f90:85                                   */
    double gamma;              /* 0.9      KrylovSolver.f90:87                                   */
    double break_tol;          /* 1e-7     KrylovSolver.f90:173                                  */
    double drop_tol0;          /* 1e-8     StateSpace.f90:416                                    */
    double drop_deriv_tol;     /* 1e-8     StateSpace.f90:491                                    */
    double drop_fraction;      /* 0.1      StateSpace.f90:497                                    */
    uint64_t seed;             /* seed of the per-trajectory Philox streams of SSA_EXTENDER      */
} kfsp_options;

/* The counters DGEXPV_FSP computes and then discards (KrylovSolver.f90:554-573). */
typedef struct {
    int32_t nmult, nexph, nscale, nstep, nreject, ibrkflag, mbrkdwn, iflag;
    double step_min, step_max, x_error, s_error, tbrkdwn, t_now, hump, beta_ratio;
    int64_t n_expand, n_drop;
    int64_t n_final;           /* FSP%SIZE on return                                             */
    int64_t n_max;             /* largest FSP%SIZE seen                                          */
    int64_t kernel_launches;   /* CUDA kernels launched by this solve                            */
    double device_seconds;     /* CUDA-event time of the solve (device work, after the H2D copy) */
    double spmv_seconds;       /* CUDA-event time inside generator SpMV launches (if profiled)   */
    double wall_seconds;       /* host wall clock of the whole call                              */
    int64_t spmv_launches;     /* generator SpMV launches timed in spmv_seconds                  */
} kfsp_stats;

/* One row per pass of the time-step loop (label 100), for decision-trace parity. */
typedef struct {
    double t_now, t_step, t_new, wsum, err_loc, beta;
    int32_t m, n_step, n_after, flags, nmult, nexph;   /* flags: 1 expanded, 2 dropped, 4 FSP 5-reject path, 8 happy breakdown */
} kfsp_trace_row;

const char* kfsp_version(void);
const char* kfsp_status_string(int status);
int kfsp_default_options(kfsp_options* opts);

/* ---- host-side model: replaces nothing in the Fortran host (it keeps src/model and
 *      src/parser); this mirror exists so that hosts without the Fortran toolchain
 *      (the C++/Python drivers in this repo) can read the same `.input` files. ---------- */
/* CME_MODEL%CREATE  src/model/ModelModule.f90:46-57 */
int kfsp_model_create(int32_t nspecies, int32_t nreactions, int32_t nparameters, kfsp_model* out);
/* CME_MODEL%LOAD    src/model/ModelModule.f90:59-161 (keywords matched case-insensitively) */
int kfsp_model_load(const char* path, kfsp_model* out);
int kfsp_model_free(kfsp_model m);
int kfsp_model_dims(kfsp_model m, int32_t* nspecies, int32_t* nreactions, int32_t* nparameters);
int kfsp_model_get_stoichiometry(kfsp_model m, int32_t* stoich /* S*R */);
int kfsp_model_set_stoichiometry(kfsp_model m, const int32_t* stoich /* S*R */);
int kfsp_model_species_name(kfsp_model m, int32_t i, char* buf, int32_t buflen);
int kfsp_model_parameter_name(kfsp_model m, int32_t i, char* buf, int32_t buflen);
/* CME_MODEL%RESET_PARAMETERS  src/model/ModelModule.f90:201-217 */
int kfsp_model_reset_parameters(kfsp_model m, const double* pval, int32_t n);
/* EQUATIONPARSER(string, vars): compile one propensity string (src/parser/FortranParser.f90:135,533) */
int kfsp_model_set_propensity_string(kfsp_model m, int32_t reaction /*1-based*/, const char* expr);
/* The PRIVATE PROPPARSER(R) byte code, for hosts that compiled it themselves
 * (opcodes of FortranParser.f90:52-73; variables = species then parameters). */
int kfsp_model_set_propensity_bytecode(kfsp_model m, int32_t reaction, const int32_t* code, int32_t ncode,
                                       const double* immed, int32_t nimmed);
int kfsp_model_get_propensity_bytecode(kfsp_model m, int32_t reaction, int32_t* code, int32_t* ncode /*in: cap*/,
                                       double* immed, int32_t* nimmed /*in: cap*/);
/* CUSTOMPROP  src/model/ModelModule.f90:6-12,31: host callback, reaction 1-based.  */
typedef double (*kfsp_propensity_fn)(const int32_t* state, int32_t reaction, const double* params, void* ctx);
int kfsp_model_set_custom_propensity(kfsp_model m, kfsp_propensity_fn fn, void* ctx);
/* MODEL%PROPENSITY(STATE, REACTION)  src/model/ModelModule.f90:163-199, evaluated on the host */
int kfsp_model_propensity(kfsp_model m, const int32_t* state, int32_t reaction, double* out);
/* The same value computed through the factored form the index-only SpMV (spmv_variant = 2) uses: sub-expressions that read one
   species each, combined by + - *.  KFSP_ERR_UNSUPPORTED if the propensity has no such form.  nterms / nops (optional): size of
   the factored form.  Host only (no GPU).  Reference: MODEL%PROPENSITY, ModelModule.f90:163-199. */
int kfsp_model_propensity_factored(kfsp_model m, const int32_t* state, int32_t reaction, double* out, int32_t* nterms, int32_t* nops);
/* Structure of a CUSTOMPROP callback found by probing (host only, no GPU): species_out[k] (R entries) = the 0-based species
   reaction k+1 reads (0 for a constant), or -2 if it reads several.  *single_out = 1: every reaction reads at most one species;
   2: the rest are bilinear mass action, a = (c * x_a) * x_b in either operand order; 0: neither.  For 1 and 2 the tables built
   over 0..max_molecules / the bilinear form agree with the callback bit for bit on 16384 pseudo-random states and kfsp_set_model
   serves the model from the device (no host round trips; lattice, index-only and multi-GPU variants included); for 0 the callback
   stays a host function.  Reference: the CUSTOMPROP models of examples/toggle.f90:55-69 (single species) and
   examples/transcr6d.f90:63-89 (reactions 5 and 7 bilinear). */
int kfsp_model_custom_structure(kfsp_model m, int32_t max_molecules, int32_t* species_out, int32_t* single_out);

/* ---- device solver handle ------------------------------------------------------------ */
int kfsp_create(const kfsp_options* opts, kfsp_handle* out);
int kfsp_destroy(kfsp_handle h);
/* Ship MODEL (sizes, stoichiometry, parameter values, propensity byte code) to the device. */
int kfsp_set_model(kfsp_handle h, kfsp_model m);

/* CME_SOLVE / DGEXPV_FSP  src/fsp/KrylovSolver.f90:7-36, 40-573.
 * states_in/p_in are FSP_OUT%STATE(:,1:n_in) and FSP_IN%VECTOR(1:n_in) (zero padded beyond).
 * On return *n_out = FSP%SIZE, states_out (S*n_out) and p_out (n_out) hold FSP%STATE and W.
 * max_out is the caller's capacity in states.  stats may be NULL.
 * states_out may be the same array as states_in (the reference's FSP_OUT is in/out); when the state set is fixed
 * (enable_expand = enable_drop = n_init_onestep = 0) an aliased list is left as it is instead of being copied back.
 * On a partitioned handle (kfsp_dist_init, all ranks call together): states_in / p_in are the GLOBAL list and vector
 * on every rank; the call returns THIS rank's rows, *n_out = hi - lo of kfsp_dist_info, in states_out / p_out. */
int kfsp_solve(kfsp_handle h, double t, int64_t n_in, const int32_t* states_in, const double* p_in,
               double fsp_tol, double kry_tol, int32_t verbosity,
               int64_t* n_out, int32_t* states_out, double* p_out, int64_t max_out, kfsp_stats* stats);
/* Same solve with the state space and the vector already resident on the device
 * (kfsp_fsp_init + kfsp_fsp_set_vector done by the caller); results stay on the device
 * and are read with kfsp_fsp_get.  This is the HBM-resident entry bench.py times. */
int kfsp_solve_resident(kfsp_handle h, double t, double fsp_tol, double kry_tol, int32_t verbosity, kfsp_stats* stats);
int kfsp_trace_length(kfsp_handle h, int64_t* n);
int kfsp_trace_get(kfsp_handle h, kfsp_trace_row* rows, int64_t cap);

/* ---- the state-space routines DGEXPV_FSP calls, one entry point each ------------------ */
/* MATRIX_STARTER   src/state_space/StateSpace.f90:248-345 on FSP%STATE(:,1:n) */
int kfsp_fsp_init(kfsp_handle h, int64_t n, const int32_t* states);
/* spmv_variant = 1 only: the projection is the lattice [0,bounds[0]) x ... x [0,bounds[S-1]) in natural order
 * (index = sum_s x_s * prod_{r<s} bounds[r]); no state list is shipped, ADJ/OFFDIAG/DIAG are never materialised
 * (FMATVEC recomputes them from the integer state, KrylovSolver.f90:577-607 + StateSpace.f90:303-327).
 * kfsp_fsp_init / kfsp_solve accept such a set as an explicit state list too and verify it on the device. */
int kfsp_fsp_init_box(kfsp_handle h, const int32_t* bounds /* S */);
/* ONESTEP_EXTENDER src/state_space/StateSpace.f90:347-396 */
int kfsp_fsp_onestep(kfsp_handle h);
/* SSA_EXTENDER     src/state_space/StateSpace.f90:550-630 (one Philox sub-stream per trajectory) */
int kfsp_fsp_ssa(kfsp_handle h, double timestep);
/* DROP_STATES      src/state_space/StateSpace.f90:431-548 on FSP%VECTOR; *dropped = 1 if compacted */
int kfsp_fsp_drop(kfsp_handle h, double dsum, int32_t* dropped, double* droptol, int64_t* drop_count);
int kfsp_fsp_size(kfsp_handle h, int64_t* n);
/* FSP%VECTOR(1:n) = v (rest zero) */
int kfsp_fsp_set_vector(kfsp_handle h, const double* v, int64_t n);
/* Export in the reference's column form: ADJ(R,N), OFFDIAG(R,N), DIAG(N), reaction fastest
 * (StateSpace.f90:13-17).  Any pointer may be NULL. */
int kfsp_fsp_get(kfsp_handle h, int32_t* states, int32_t* adj, double* offdiag, double* diag, double* vector);
/* FSP%INDEX(X) (StateSpace.f90:116-134) and FSP%PROBABILITY(X) (:96-114) for n query states */
int kfsp_fsp_index(kfsp_handle h, int64_t n, const int32_t* states, int32_t* index_out);
int kfsp_fsp_probability(kfsp_handle h, int64_t n, const int32_t* states, double* p_out);

/* ---- the numerical kernels ------------------------------------------------------------- */
/* FMATVEC  src/fsp/KrylovSolver.f90:577-607: y = A x on the current FSP matrix (host buffers) */
int kfsp_matvec(kfsp_handle h, const double* x, double* y);
/* Same with device pointers; `reps` back-to-back launches timed with CUDA events on the
 * solver's stream; *seconds_per_launch may be NULL. */
int kfsp_matvec_device(kfsp_handle h, const double* x_device, double* y_device, int32_t reps, double* seconds_per_launch);
/* One Arnoldi/IOP-2 sweep of m columns from v (KrylovSolver.f90:223-266): returns the
 * (m+2)x(m+2) column-major H with H(m+2,m+1)=1, ||A v_{m+1}|| and the happy-breakdown column (0 if none). */
int kfsp_arnoldi(kfsp_handle h, const double* v, int32_t m, double* H_out, double* avnorm, int32_t* breakdown,
                 double* seconds);
/* DGPADMnorm  src/expokit/dgpadm.f:171-339 on one CTA: out = exp(t*H(1:m,1:m)), ld(out)=m */
int kfsp_expm(kfsp_handle h, int32_t m, double t, const double* H, int32_t ldh, double* out, int32_t* ns, double* hnorm);
/* W = beta*V(:,1:mx)*e, clamp, 1-norm (KrylovSolver.f90:444-450) exposed for tests: V is n x mx column-major.
 * colscale (mx entries, NULL = all 1): the device keeps the basis un-normalised, column j of V is multiplied by
 * colscale[j] = 1/HJ1J on load (the DSCAL of KrylovSolver.f90:258, never run as a pass).  wssq (optional): sum w^2. */
int kfsp_combine(kfsp_handle h, int64_t n, int32_t mx, double beta, const double* V, const double* e, const double* colscale,
                 double* w, double* wsum, double* wssq);

/* ---- multi-GPU (one process per GPU; rows of the state space are block-partitioned) ---- */
#define KFSP_NCCL_ID_BYTES 128
int kfsp_dist_unique_id(uint8_t id[KFSP_NCCL_ID_BYTES]);
/* Call before kfsp_fsp_init.  Afterwards kfsp_fsp_init takes the GLOBAL state list on every rank and keeps
 * rows [lo,hi); kfsp_fsp_set_vector takes the global vector; kfsp_fsp_get returns this rank's rows.
 * Partitioned state sets are fixed: create the handle with enable_expand = enable_drop = n_init_onestep = 0. */
int kfsp_dist_init(kfsp_handle h, int32_t rank, int32_t nranks, const uint8_t id[KFSP_NCCL_ID_BYTES]);
/* host-side partition arithmetic (no GPU needed): block partition of n rows */
int kfsp_dist_partition(int64_t n, int32_t nranks, int32_t rank, int64_t* lo, int64_t* hi);
int kfsp_dist_owner(int64_t n, int32_t nranks, int64_t row, int32_t* owner);
/* host-side arithmetic of the lattice variant (spmv_variant = 1; no GPU needed): rank's slab [zlo,zhi) of the slowest
 * species, and which SpMV kernel a lattice model gets (*kind: 0 generic, 1 / 2 two-species stencil kernel with the
 * reaction order of toggle_model.input / toggle_test_model.input; *table_mask bit k: reaction k's table runs over
 * the slowest species). */
int kfsp_lattice_partition(int32_t nz, int32_t nranks, int32_t rank, int32_t* zlo, int32_t* zhi);
int kfsp_lattice_kernel(int32_t S, int32_t R, const int32_t* stoich /* S*R */, const int32_t* table_species /* R */,
                        int32_t* kind, int32_t* table_mask);
/* host-side arithmetic of the replicated layout used for ADAPTIVE state sets on several GPUs (no GPU needed): rows [lo,hi) of
   the Krylov loop computed by `rank` when the set has n rows; whole = 1 while n < min_rows (KFSP_REPL_MIN_ROWS, default 2^22):
   every rank then computes every row and nothing is exchanged (lo = 0, hi = n).  No reference counterpart (serial code). */
int kfsp_repl_partition(int64_t n, int32_t nranks, int32_t rank, int64_t min_rows, int64_t* lo, int64_t* hi, int32_t* whole);
int kfsp_dist_info(kfsp_handle h, int64_t* lo, int64_t* hi, int64_t* n_halo, int64_t* n_send, int64_t* halo_bytes, int64_t* reductions);
/* How the device path evaluates the propensities of the current model (MODEL%PROPENSITY, ModelModule.f90:163-199):
   n_tabulated      programs that read one species and hold a transcendental operation: tabulated by the host over the count
   n_host_evaluated reactions evaluated by the host itself: CUSTOMPROP callbacks, and byte code with a transcendental operation
                    on several species (the CUDA math library may differ from the host libm by ulps there; KFSP_DEVICE_MATH=1
                    keeps them on the device)
   n_device_libm    such programs that DO run on the CUDA math library (only with KFSP_DEVICE_MATH=1): results may then differ
                    from a host evaluation in the last bits, and with them SSA picks and pruning decisions
   factored         1 if the index-only SpMV's factored tables are in use (spmv_variant = 2) */
int kfsp_model_info(kfsp_handle h, int32_t* n_tabulated, int32_t* n_host_evaluated, int32_t* n_device_libm, int32_t* factored);
/* Peer-memory path: the reduction exchange fused into the tail of every reducing kernel (no reference counterpart; the
   reference is serial).  exchanges = fused exchanges since the last reset; mean_us / max_us = time between posting this rank's
   double-double partial to the peers and holding every rank's partial, i.e. NVLink latency plus the wait for the slowest rank. */
int kfsp_dist_exchange_stats(kfsp_handle h, int64_t* exchanges, double* mean_us, double* max_us, int32_t reset);

/* device memory helpers for hosts without their own CUDA runtime (bench, tests) */
int kfsp_device_alloc(kfsp_handle h, int64_t bytes, void** ptr);
int kfsp_device_free(kfsp_handle h, void* ptr);
int kfsp_device_upload(kfsp_handle h, void* dst_device, const void* src_host, int64_t bytes);
int kfsp_device_download(kfsp_handle h, void* dst_host, const void* src_device, int64_t bytes);
int kfsp_device_vector(kfsp_handle h, double** fsp_vector_device);
int kfsp_flush_l2(kfsp_handle h);
/* CUDA events on the solver's stream around the launches of the time-stepping loop of the next solves (no synchronisation
 * inside the solve).  level 0: off; 1: one event pair around each Arnoldi sweep (class KFSP_PROF_SWEEP: nothing is recorded
 * between the sweep's launches); 2: one pair around every launch (per-class table).  The generator SpMV's share is reported
 * in kfsp_stats.spmv_seconds / spmv_launches, everything by class through kfsp_profile_get. */
int kfsp_set_profiling(kfsp_handle h, int32_t level);
/* Host waits of this handle sleep on a blocking CUDA event instead of spinning in the driver (default off: lowest latency).
   For many handles solving concurrently on one GPU from more host threads than cores -- parameter sweeps of small models,
   krylovfspssa_b200/sweep.py.  No reference counterpart (the reference is a serial CPU code). */
int kfsp_set_blocking_sync(kfsp_handle h, int32_t on);
/* FSP%VECTOR(1:n) = src (device pointer), rest zero: device-to-device reset between benchmark steps */
int kfsp_fsp_set_vector_device(kfsp_handle h, const double* src_device, int64_t n);
int kfsp_launch_count(kfsp_handle h, int64_t* n);
/* Device time of the last solve by kernel class (kfsp_set_profiling on: CUDA events on the solver's stream around every
 * launch of the time-stepping loop).  SPMV_* and FIN_* are the generator SpMV (FMATVEC, KrylovSolver.f90:577-607) in its
 * variants: plain; one Arnoldi column (FMATVEC + the inner products that give H(J-1,J) and H(J,J), :240-245); the extra
 * product with its norm (:261-263); FIN_*: the same two with the previous column's two DAXPYs + DNRM2 (:243-247) fused
 * into the load stage.  AXPY_NRM is that finalisation as a launch of its own (explicit-matrix path).
 * bytes_per_state (optional): algorithmic bytes per state each launch had to move, summed over the class's launches. */
enum {
    KFSP_PROF_SPMV_PLAIN = 0, KFSP_PROF_SPMV_DOT = 1, KFSP_PROF_SPMV_NRM = 2, KFSP_PROF_SPMV_FIN_DOT = 3, KFSP_PROF_SPMV_FIN_NRM = 4,
    KFSP_PROF_AXPY_DOT = 5, KFSP_PROF_AXPY_NRM = 6, KFSP_PROF_COMBINE = 7, KFSP_PROF_SCALE_COPY = 8, KFSP_PROF_EXPM = 9,
    KFSP_PROF_SWEEP = 10,      /* profiling level 1: one event pair around each Arnoldi sweep (all its launches together) */
    KFSP_PROF_CLASSES = 12
};
int kfsp_profile_get(kfsp_handle h, double seconds[KFSP_PROF_CLASSES], int64_t launches[KFSP_PROF_CLASSES],
                     int64_t bytes_per_state[KFSP_PROF_CLASSES]);
/* Generator-SpMV launches of the last solve by kind: [0] plain FMATVEC, [1] fused with the first IOP DDOT
 * (KrylovSolver.f90:240-243), [2] fused with the norm of the extra product (AVNORM, :261-263), [3] how many of [1]+[2]
 * also finalised the previous Arnoldi column in their load stage (DAXPY + DNRM2 of :244-247; lattice stencil kernel). */
int kfsp_spmv_launch_counts(kfsp_handle h, int64_t out[4]);
/* Host wall clock of the last solve by phase: [0] Arnoldi sweep + Pade, [1] basis combination + norms,
 * [2] SSA_EXTENDER, [3] DROP_STATES, [4] ONESTEP_EXTENDER, [5] host propensity callbacks (CUSTOMPROP; included in 2 and 4), [6] SSA side-cache rounds and [7] host propensity evaluations (counts) (each phase ends in a stream synchronisation). */
int kfsp_phase_seconds(kfsp_handle h, double out[8]);

#ifdef __cplusplus
}
#endif
#endif /* KFSP_H */
