// libkfsp.so: C ABI (include/kfsp.h) over the device engine.  sm_100a only; there is no
// CPU path behind these entry points.
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>

#include "controller.h"
#include "engine.cuh"
#include "model_host.h"

#ifdef KFSP_WITH_NCCL
#include <nccl.h>
#endif

using namespace kfsp;

struct kfsp_model_s {
    HostModel m;
};
struct kfsp_handle_s {
    Engine e;
};

namespace {

// Controller backend: every call enqueues kernels on the engine's stream; only the calls that
// return scalars synchronise.
struct GpuBackend {
    Engine& e;
    explicit GpuBackend(Engine& en) : e(en) {}
    int64_t size() { return (e.dist.nranks > 1 && !e.dist.repl) ? e.dist.n_global : e.n; }
    // phase timers: every timed call ends in a stream synchronisation, so host wall clock is device time + latency
    struct Tick { double& acc; double t0; explicit Tick(double& a) : acc(a), t0(wall_now()) {} ~Tick() { acc += wall_now() - t0; } };
    int onestep() { Tick t(e.phase_s[4]); int st = e.fsp_onestep(); if (st == KFSP_OK) st = e.sync(); return st; }
    int ssa(double ts) { Tick t(e.phase_s[2]); int st = e.fsp_ssa(ts); if (st == KFSP_OK) st = e.sync(); return st; }
    int drop(double dsum, int* dropped) {
        Tick t(e.phase_s[3]);
        int32_t d = 0;
        int st = e.fsp_drop(dsum, &d, nullptr, nullptr);
        *dropped = d;
        return st;
    }
    int norms(double* wsum, double* wssq) {
        Tick t(e.phase_s[1]);
        k_norms<<<e.wave_grid((const void*)k_norms, e.kn()), VEC_THREADS, 0, e.stream>>>(e.kn(), e.d_w + e.kr0(), e.next_rd(), e.d_ctl);
        KFSP_TRY(e.check_launch());
        { EpiArgs en = Engine::epi_none(); en.kind = RK_NORMS; KFSP_TRY(e.dist_finalize(en, 2)); }
        KFSP_TRY(e.read_ctl());
        *wsum = e.h_ctl->scal[SC_WSUM];
        *wssq = std::sqrt(e.h_ctl->scal[SC_WSSQ]);
        return KFSP_OK;
    }
    int begin_step(double inv_beta) {
        KFSP_CUDA(cudaMemsetAsync(e.d_H, 0, sizeof(double) * e.LDH * e.LDH, e.stream));
        k_reset_ctl<<<1, 128, 0, e.stream>>>(e.d_ctl);
        KFSP_TRY(e.check_launch());
        KFSP_TRY(e.prof_begin(KFSP_PROF_SCALE_COPY, 16));
        KFSP_TRY(e.launch_pdl(k_scale_copy, e.wave_grid((const void*)k_scale_copy, e.kn()), VEC_THREADS, 0, e.kn(), inv_beta, (const double*)e.d_w + e.kr0(),
                              e.d_V + e.kr0()));
        KFSP_TRY(e.dist_barrier());        // neighbours gather column 0 straight from this GPU's HBM
        return e.prof_end();
    }
    int arnoldi(int jold, int m) { Tick t(e.phase_s[0]); return e.arnoldi(jold, m); }
    int expm(int mx_ok, double t_ok, int use_brk, double t_brk, int set_one, StepScalars* out) {
        Tick t(e.phase_s[0]);
        int st = e.expm_step(mx_ok, t_ok, use_brk, t_brk, set_one);
        if (st != KFSP_OK) return st;
        out->ns = e.h_res->ns; out->brk = use_brk ? e.h_res->brk : 0; out->mx = e.h_res->mx;
        out->hnorm = e.h_res->hnorm; out->avnorm = e.h_res->avnorm; out->e = e.h_res->e;
        return KFSP_OK;
    }
    int clear_h(int row0, int col0) {
        k_set_entry<<<1, 1, 0, e.stream>>>(e.d_H + (size_t)col0 * e.LDH + row0, 0.0);
        return e.check_launch();
    }
    int combine(int mx, double beta, double* wsum, double* wssq) {
        Tick t(e.phase_s[1]);
        KFSP_TRY(e.prof_begin(KFSP_PROF_COMBINE, 8 * (mx + 1)));
        KFSP_TRY(e.launch_pdl(k_combine, e.wave_grid((const void*)k_combine, e.kn()), VEC_THREADS, 0, e.kn(), e.ld, mx, beta, (const double*)e.d_V + e.kr0(),
                              (const double*)e.d_res->e, e.d_w + e.kr0(), e.next_rd(), e.d_ctl));
        { EpiArgs en = Engine::epi_none(); en.kind = RK_NORMS; KFSP_TRY(e.dist_finalize(en, 2)); }
        KFSP_TRY(e.prof_end());
        KFSP_TRY(e.read_ctl());
        *wsum = e.h_ctl->scal[SC_WSUM];
        *wssq = std::sqrt(e.h_ctl->scal[SC_WSSQ]);
        return KFSP_OK;
    }
    int restore_w(double beta, double* wssq) {
        k_scale_copy_nrm<<<e.grid_for(e.kn()), VEC_THREADS, 0, e.stream>>>(e.kn(), beta, e.d_V + e.kr0(), e.d_w + e.kr0(), e.next_rd(), e.d_ctl);
        KFSP_TRY(e.check_launch());
        { EpiArgs en = Engine::epi_none(); en.kind = RK_NORMS; KFSP_TRY(e.dist_finalize(en, 2)); }
        KFSP_TRY(e.read_ctl());
        *wssq = std::sqrt(e.h_ctl->scal[SC_WSSQ]);
        return KFSP_OK;
    }
};

}  // namespace

int Engine::solve(double T, double fsptol, double krytol, int itrace, kfsp_stats* stats) {
    if (n < 1) return KFSP_ERR_BAD_SIZES;
    KFSP_CUDA(cudaSetDevice(device));
    KFSP_TRY(ensure_basis());
    const double w0 = wall_now();
    const int64_t l0 = launches;
    for (double& v : phase_s) v = 0.0;
    host_prop_rounds = host_prop_evals = 0;
    spmv_by_mode[0] = spmv_by_mode[1] = spmv_by_mode[2] = 0;
    spmv_fused = 0;
    spmv_seconds = 0.0;
    spmv_timed = 0;
    ev_used = 0;
    for (int i = 0; i < KFSP_PROF_CLASSES; ++i) { prof_sec[i] = 0.0; prof_cnt[i] = 0; prof_bps[i] = 0; }
    cudaEvent_t e0, e1;
    KFSP_CUDA(cudaEventCreate(&e0));
    KFSP_CUDA(cudaEventCreate(&e1));
    KFSP_CUDA(cudaEventRecord(e0, stream));
    kfsp_stats local;
    std::memset(&local, 0, sizeof local);
    GpuBackend be(*this);
    Controller ctl(opt);
    int st = ctl.run(be, R, T, fsptol, krytol, itrace, &local, trace);
    if (dist.repl && gather_rows(d_w) != KFSP_OK && st == KFSP_OK) st = KFSP_ERR_NCCL;     // every rank ends with the whole vector
    cudaEventRecord(e1, stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    local.device_seconds = 1e-3 * ms;
    local.wall_seconds = wall_now() - w0;
    local.kernel_launches = launches - l0;
    collect_profile();
    local.spmv_seconds = spmv_seconds;
    local.spmv_launches = spmv_timed;
    local.n_final = n;
    if (stats) *stats = local;
    return st;
}

extern "C" {

const char* kfsp_version(void) { return "krylovfspssa_b200 0.1 (sm_100a)"; }

const char* kfsp_status_string(int s) {
    switch (s) {
    case KFSP_OK: return "ok";
    case KFSP_IFLAG_MXSTEP: return "maximum number of steps reached (IFLAG=1)";
    case KFSP_IFLAG_TOLERANCE: return "the requested tolerance is too high (IFLAG=2)";
    case KFSP_ERR_BAD_SIZES: return "bad sizes (in input of DGEXPV)";
    case KFSP_ERR_NULL_H: return "null H in input of DGPADM";
    case KFSP_ERR_SINGULAR: return "problem in DGESV (within DGPADM)";
    case KFSP_ERR_OVERFLOW: return "overflow error: FSP size exceeds memory limit";
    case KFSP_ERR_BAD_STATE: return "initial state negative, above the molecule limit, or duplicated";
    case KFSP_ERR_MOLECULE_LIMIT: return "a molecule count exceeded MAXNUMBERMOLECULES";
    case KFSP_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
    case KFSP_ERR_CUDA: return "CUDA runtime error";
    case KFSP_ERR_ARG: return "invalid argument";
    case KFSP_ERR_NO_MODEL: return "no model / propensities set";
    case KFSP_ERR_PARSE: return "syntax error in model input";
    case KFSP_ERR_IO: return "cannot open file";
    case KFSP_ERR_UNSUPPORTED: return "not supported by the device path";
    case KFSP_ERR_OUT_TOO_SMALL: return "output buffers too small";
    case KFSP_ERR_NCCL: return "multi-GPU exchange failed (NCCL error or a peer stopped responding)";
    case KFSP_ERR_SSA_RUNAWAY: return "SSA trajectory exceeded the jump limit";
    }
    return "unknown status";
}

int kfsp_default_options(kfsp_options* o) {
    if (!o) return KFSP_ERR_ARG;
    std::memset(o, 0, sizeof *o);
    o->m_max = 100; o->m_min = 10; o->ideg = 6; o->n_init_onestep = 5; o->fsp_reject_limit = 5;
    o->mxstep = 0; o->mxreject = 0; o->enable_drop = 1; o->enable_expand = 1; o->max_molecules = 10000;
    o->device = -1; o->spmv_variant = 0; o->max_states = 6291469;
    o->delta = 1.2; o->gamma = 0.9; o->break_tol = 1.0e-7; o->drop_tol0 = 1.0e-8; o->drop_deriv_tol = 1.0e-8;
    o->drop_fraction = 0.1; o->seed = 12345;
    return KFSP_OK;
}

// ------------------------------------------------------------------ model (host)
int kfsp_model_create(int32_t S, int32_t R, int32_t P, kfsp_model* out) {
    if (!out || S < 1 || R < 1 || P < 0) return KFSP_ERR_ARG;
    kfsp_model_s* m = new (std::nothrow) kfsp_model_s();
    if (!m) return KFSP_ERR_ARG;
    m->m.S = S; m->m.R = R; m->m.P = P;
    m->m.stoich.assign((size_t)S * R, 0);
    m->m.params.assign(P, 0.0);
    m->m.programs.assign(R, Program());
    m->m.propensity_strings.assign(R, "");
    for (int i = 0; i < S; ++i) m->m.species.push_back("X" + std::to_string(i + 1));
    for (int i = 0; i < P; ++i) m->m.parameters.push_back("p" + std::to_string(i + 1));
    *out = m;
    return KFSP_OK;
}
int kfsp_model_load(const char* path, kfsp_model* out) {
    if (!path || !out) return KFSP_ERR_ARG;
    kfsp_model_s* m = new (std::nothrow) kfsp_model_s();
    if (!m) return KFSP_ERR_ARG;
    std::string err;
    if (!load_model_file(path, m->m, err)) {
        std::fprintf(stderr, "libkfsp: %s\n", err.c_str());
        const bool io = err.rfind("ERROR OPENING FILE", 0) == 0;
        delete m;
        return io ? KFSP_ERR_IO : KFSP_ERR_PARSE;
    }
    *out = m;
    return KFSP_OK;
}
int kfsp_model_free(kfsp_model m) { delete m; return KFSP_OK; }
int kfsp_model_dims(kfsp_model m, int32_t* S, int32_t* R, int32_t* P) {
    if (!m) return KFSP_ERR_ARG;
    if (S) *S = m->m.S;
    if (R) *R = m->m.R;
    if (P) *P = m->m.P;
    return KFSP_OK;
}
int kfsp_model_get_stoichiometry(kfsp_model m, int32_t* st) {
    if (!m || !st) return KFSP_ERR_ARG;
    std::memcpy(st, m->m.stoich.data(), sizeof(int32_t) * m->m.stoich.size());
    return KFSP_OK;
}
int kfsp_model_set_stoichiometry(kfsp_model m, const int32_t* st) {
    if (!m || !st) return KFSP_ERR_ARG;
    m->m.stoich.assign(st, st + (size_t)m->m.S * m->m.R);
    return KFSP_OK;
}
static int copy_name(const std::string& s, char* buf, int32_t len) {
    if (!buf || len < 1) return KFSP_ERR_ARG;
    std::strncpy(buf, s.c_str(), len - 1);
    buf[len - 1] = '\0';
    return KFSP_OK;
}
int kfsp_model_species_name(kfsp_model m, int32_t i, char* buf, int32_t len) {
    if (!m || i < 0 || i >= (int)m->m.species.size()) return KFSP_ERR_ARG;
    return copy_name(m->m.species[i], buf, len);
}
int kfsp_model_parameter_name(kfsp_model m, int32_t i, char* buf, int32_t len) {
    if (!m || i < 0 || i >= (int)m->m.parameters.size()) return KFSP_ERR_ARG;
    return copy_name(m->m.parameters[i], buf, len);
}
int kfsp_model_reset_parameters(kfsp_model m, const double* p, int32_t cnt) {
    if (!m || (!p && cnt > 0) || cnt < m->m.P) return KFSP_ERR_ARG;
    m->m.params.assign(p, p + m->m.P);
    return KFSP_OK;
}
int kfsp_model_set_propensity_string(kfsp_model m, int32_t reaction, const char* expr) {
    if (!m || !expr || reaction < 1 || reaction > m->m.R) return KFSP_ERR_ARG;
    std::vector<std::string> vars = m->m.species;
    vars.insert(vars.end(), m->m.parameters.begin(), m->m.parameters.end());
    std::string err;
    Program p;
    if (!compile_expression(expr, vars, p, err)) {
        std::fprintf(stderr, "libkfsp: *** Error in syntax of function string: %s\n", err.c_str());
        return KFSP_ERR_PARSE;
    }
    m->m.programs[reaction - 1] = p;
    m->m.propensity_strings[reaction - 1] = expr;
    return KFSP_OK;
}
int kfsp_model_set_propensity_bytecode(kfsp_model m, int32_t reaction, const int32_t* code, int32_t ncode, const double* immed, int32_t nimmed) {
    if (!m || !code || ncode < 1 || nimmed < 0 || reaction < 1 || reaction > m->m.R) return KFSP_ERR_ARG;
    Program p;
    p.code.assign(code, code + ncode);
    if (nimmed) p.immed.assign(immed, immed + nimmed);
    p.stack_depth = program_stack_depth(p, m->m.S + m->m.P);
    if (p.stack_depth < 0) return KFSP_ERR_PARSE;
    m->m.programs[reaction - 1] = p;
    return KFSP_OK;
}
int kfsp_model_get_propensity_bytecode(kfsp_model m, int32_t reaction, int32_t* code, int32_t* ncode, double* immed, int32_t* nimmed) {
    if (!m || !ncode || !nimmed || reaction < 1 || reaction > m->m.R) return KFSP_ERR_ARG;
    const Program& p = m->m.programs[reaction - 1];
    const int32_t cc = *ncode, ci = *nimmed;
    *ncode = (int32_t)p.code.size();
    *nimmed = (int32_t)p.immed.size();
    if (cc < *ncode || ci < *nimmed) return KFSP_ERR_OUT_TOO_SMALL;
    if (code) std::memcpy(code, p.code.data(), sizeof(int32_t) * p.code.size());
    if (immed && !p.immed.empty()) std::memcpy(immed, p.immed.data(), sizeof(double) * p.immed.size());
    return KFSP_OK;
}
int kfsp_model_set_custom_propensity(kfsp_model m, kfsp_propensity_fn fn, void* ctx) {
    if (!m) return KFSP_ERR_ARG;
    m->m.custom = fn;
    m->m.custom_ctx = ctx;
    return KFSP_OK;
}
int kfsp_model_propensity(kfsp_model m, const int32_t* state, int32_t reaction, double* out) {
    if (!m || !state || !out || reaction < 1 || reaction > m->m.R) return KFSP_ERR_ARG;
    if (!m->m.custom && m->m.programs[reaction - 1].empty()) return KFSP_ERR_NO_MODEL;
    *out = m->m.propensity(state, reaction);
    return KFSP_OK;
}

// Host-side view of the factored form the index-only SpMV uses (spmv_variant = 2; model_host.h: factor_program): the
// propensity evaluated THROUGH its single-species terms, exactly as the device combines the tabulated terms.  No GPU needed.
int kfsp_model_propensity_factored(kfsp_model m, const int32_t* state, int32_t reaction, double* out, int32_t* nterms, int32_t* nops) {
    if (!m || !state || !out || reaction < 1 || reaction > m->m.R) return KFSP_ERR_ARG;
    if (m->m.custom) return KFSP_ERR_UNSUPPORTED;
    const Program& prog = m->m.programs[reaction - 1];
    if (prog.empty()) return KFSP_ERR_NO_MODEL;
    Factored fp;
    if (!factor_program(prog, m->m.S, fp)) return KFSP_ERR_UNSUPPORTED;
    if (fp.terms.empty() || (int)fp.terms.size() > FAC_MAX_TERMS || (int)fp.ops.size() > FAC_MAX_OPS) return KFSP_ERR_UNSUPPORTED;   // device structure
    if (nterms) *nterms = (int32_t)fp.terms.size();
    if (nops) *nops = (int32_t)fp.ops.size();
    const int S = m->m.S, P = m->m.P;
    std::vector<double> val((size_t)S + P, 0.0), stack;
    for (int i = 0; i < P; ++i) val[S + i] = m->m.params[i];
    const bool multi = fp.terms.size() > 1;
    for (int32_t op : fp.ops) {
        if (op >= 0) {
            const FactoredTerm& t = fp.terms[op];
            for (int s2 = 0; s2 < S; ++s2) val[s2] = 0.0;
            if (t.species >= 0) val[t.species] = (double)state[t.species];     // what the table holds at this count
            bool ab = false;
            stack.push_back(evaluate_program_checked(t.prog, val.data(), &ab));
            if (ab && multi) return KFSP_ERR_UNSUPPORTED;
        } else if (op == -cNeg) {
            stack.back() = -stack.back();
        } else {
            const double b = stack.back(); stack.pop_back();
            double& a = stack.back();
            a = op == -cAdd ? a + b : op == -cSub ? a - b : a * b;
        }
    }
    *out = stack.empty() ? 0.0 : stack[0];
    return KFSP_OK;
}

int kfsp_model_custom_structure(kfsp_model m, int32_t max_molecules, int32_t* species_out, int32_t* single_out) {
    if (!m || !species_out || !single_out || max_molecules < 1) return KFSP_ERR_ARG;
    if (!m->m.custom) return KFSP_ERR_NO_MODEL;
    CustomProbe pr;
    const bool ok = probe_custom(m->m, max_molecules, pr);
    *single_out = ok ? (pr.all_single ? 1 : 2) : 0;
    for (int k = 0; k < m->m.R; ++k) species_out[k] = pr.species[k];
    return KFSP_OK;
}

// ------------------------------------------------------------------ handle
int kfsp_create(const kfsp_options* opts, kfsp_handle* out) {
    if (!out) return KFSP_ERR_ARG;
    kfsp_options o;
    if (opts) o = *opts; else kfsp_default_options(&o);
    kfsp_handle_s* h = new (std::nothrow) kfsp_handle_s();
    if (!h) return KFSP_ERR_ARG;
    int st = h->e.init(&o);
    if (st != KFSP_OK) { delete h; return st; }
    *out = h;
    return KFSP_OK;
}
int kfsp_destroy(kfsp_handle h) {
    if (!h) return KFSP_OK;
    h->e.destroy();
    delete h;
    return KFSP_OK;
}
int kfsp_set_model(kfsp_handle h, kfsp_model m) {
    if (!h || !m) return KFSP_ERR_ARG;
    return h->e.set_model(m->m);
}

int kfsp_fsp_init(kfsp_handle h, int64_t n, const int32_t* states) {
    if (!h || !states) return KFSP_ERR_ARG;
    cudaSetDevice(h->e.device);
    return h->e.fsp_init(n, states);
}
int kfsp_fsp_init_box(kfsp_handle h, const int32_t* bounds) {
    if (!h || !bounds) return KFSP_ERR_ARG;
    cudaSetDevice(h->e.device);
    return h->e.fsp_init_box(bounds);
}
int kfsp_fsp_onestep(kfsp_handle h) {
    if (!h) return KFSP_ERR_ARG;
    cudaSetDevice(h->e.device);
    int st = h->e.fsp_onestep();
    if (st == KFSP_OK) st = h->e.sync();
    return st;
}
int kfsp_fsp_ssa(kfsp_handle h, double timestep) {
    if (!h) return KFSP_ERR_ARG;
    cudaSetDevice(h->e.device);
    int st = h->e.fsp_ssa(timestep);
    if (st == KFSP_OK) st = h->e.sync();
    return st;
}
int kfsp_fsp_drop(kfsp_handle h, double dsum, int32_t* dropped, double* droptol, int64_t* drop_count) {
    if (!h || !dropped) return KFSP_ERR_ARG;
    cudaSetDevice(h->e.device);
    int st = h->e.fsp_drop(dsum, dropped, droptol, drop_count);
    if (st == KFSP_OK) st = h->e.sync();
    return st;
}
int kfsp_fsp_size(kfsp_handle h, int64_t* n) {
    if (!h || !n) return KFSP_ERR_ARG;
    *n = h->e.n;
    return KFSP_OK;
}
int kfsp_fsp_set_vector(kfsp_handle h, const double* v, int64_t cnt) {
    if (!h || (!v && cnt > 0) || cnt < 0) return KFSP_ERR_ARG;
    Engine& e = h->e;
    if (e.ld == 0) return KFSP_ERR_BAD_SIZES;
    cudaSetDevice(e.device);
    if (e.dist.nranks > 1 && !e.dist.repl) {                       // v is the GLOBAL vector: keep rows [lo, hi)
        if (cnt > e.dist.n_global) return KFSP_ERR_BAD_SIZES;
        KFSP_CUDA(cudaMemsetAsync(e.d_w, 0, sizeof(double) * e.ld, e.stream));
        const int64_t a = std::min<int64_t>(cnt, e.dist.lo), b = std::min<int64_t>(cnt, e.dist.hi);
        if (b > a) KFSP_CUDA(cudaMemcpyAsync(e.d_w, v + a, sizeof(double) * (b - a), cudaMemcpyHostToDevice, e.stream));
        return e.sync();
    }
    if (cnt > e.ld) return KFSP_ERR_BAD_SIZES;
    KFSP_CUDA(cudaMemsetAsync(e.d_w, 0, sizeof(double) * e.ld, e.stream));
    if (cnt) KFSP_CUDA(cudaMemcpyAsync(e.d_w, v, sizeof(double) * cnt, cudaMemcpyHostToDevice, e.stream));
    return e.sync();
}
int kfsp_fsp_get(kfsp_handle h, int32_t* states, int32_t* adj, double* offdiag, double* diag, double* vector) {
    if (!h) return KFSP_ERR_ARG;
    Engine& e = h->e;
    if (e.n < 1) return KFSP_ERR_BAD_SIZES;
    cudaSetDevice(e.device);
    const int64_t n = e.n;
    if (e.box) {                                                     // lattice: the column form is computed, not stored
        if (states) {
            KFSP_TRY(e.box_states());
            KFSP_CUDA(cudaMemcpyAsync(states, e.d_states, sizeof(int32_t) * n * e.S, cudaMemcpyDeviceToHost, e.stream));
        }
        if (vector) KFSP_CUDA(cudaMemcpyAsync(vector, e.d_w, sizeof(double) * n, cudaMemcpyDeviceToHost, e.stream));
        if (adj || offdiag || diag) {
            const size_t a = Engine::align_up(sizeof(int32_t) * n * e.R), b = Engine::align_up(sizeof(double) * n * e.R);
            KFSP_TRY(e.ensure_scratch(a + b + Engine::align_up(sizeof(double) * n)));
            int32_t* dadj = (int32_t*)e.d_scratch;
            double* doff = (double*)(e.d_scratch + a);
            double* ddiag = (double*)(e.d_scratch + a + b);
            k_box_export<<<e.grid_for(n), VEC_THREADS, 0, e.stream>>>(e.lat, e.dist.lo, n, adj ? dadj : nullptr, offdiag ? doff : nullptr,
                                                                     diag ? ddiag : nullptr);
            KFSP_TRY(e.check_launch());
            if (adj) KFSP_CUDA(cudaMemcpyAsync(adj, dadj, sizeof(int32_t) * n * e.R, cudaMemcpyDeviceToHost, e.stream));
            if (offdiag) KFSP_CUDA(cudaMemcpyAsync(offdiag, doff, sizeof(double) * n * e.R, cudaMemcpyDeviceToHost, e.stream));
            if (diag) KFSP_CUDA(cudaMemcpyAsync(diag, ddiag, sizeof(double) * n, cudaMemcpyDeviceToHost, e.stream));
        }
        return e.sync();
    }
    const bool part = e.dist.nranks > 1 && !e.dist.repl;             // memory-scaled partition: this rank's rows only
    if (part && adj) return KFSP_ERR_UNSUPPORTED;                    // the column form is not kept for partitioned sets
    const int64_t row0 = part ? e.dist.lo : 0;
    if (states) KFSP_CUDA(cudaMemcpyAsync(states, e.d_states + row0 * e.S, sizeof(int32_t) * n * e.S, cudaMemcpyDeviceToHost, e.stream));
    if (diag) KFSP_CUDA(cudaMemcpyAsync(diag, e.d_diag, sizeof(double) * n, cudaMemcpyDeviceToHost, e.stream));
    if (vector) KFSP_CUDA(cudaMemcpyAsync(vector, e.d_w, sizeof(double) * n, cudaMemcpyDeviceToHost, e.stream));
    if (adj || offdiag) {
        const size_t a = Engine::align_up(sizeof(int32_t) * n * e.R);
        KFSP_TRY(e.ensure_scratch(a + Engine::align_up(sizeof(double) * n * e.R)));
        int32_t* dadj = (int32_t*)e.d_scratch;
        double* doff = (double*)(e.d_scratch + a);
        k_export_adj<<<e.grid_for(n * e.R), VEC_THREADS, 0, e.stream>>>(e.view(), adj ? dadj : nullptr, offdiag ? doff : nullptr);
        KFSP_TRY(e.check_launch());
        if (adj) KFSP_CUDA(cudaMemcpyAsync(adj, dadj, sizeof(int32_t) * n * e.R, cudaMemcpyDeviceToHost, e.stream));
        if (offdiag) KFSP_CUDA(cudaMemcpyAsync(offdiag, doff, sizeof(double) * n * e.R, cudaMemcpyDeviceToHost, e.stream));
    }
    return e.sync();
}
static int lookup_common(kfsp_handle h, int64_t nq, const int32_t* states, int32_t* idx, double* p) {
    if (!h || !states || nq < 1) return KFSP_ERR_ARG;
    Engine& e = h->e;
    if (e.n < 1) return KFSP_ERR_BAD_SIZES;
    cudaSetDevice(e.device);
    const size_t a = Engine::align_up(sizeof(int32_t) * nq * e.S), b = Engine::align_up(sizeof(int32_t) * nq);
    KFSP_TRY(e.ensure_scratch(a + b + Engine::align_up(sizeof(double) * nq)));
    int32_t* dq = (int32_t*)e.d_scratch;
    int32_t* di = (int32_t*)(e.d_scratch + a);
    double* dp = (double*)(e.d_scratch + a + b);
    KFSP_CUDA(cudaMemcpyAsync(dq, states, sizeof(int32_t) * nq * e.S, cudaMemcpyHostToDevice, e.stream));
    if (e.box)
        k_box_lookup<<<e.grid_for(nq), VEC_THREADS, 0, e.stream>>>(e.lat, dq, nq, idx ? di : nullptr, e.d_w, e.dist.lo, e.n, p ? dp : nullptr);
    else
        k_lookup_states<<<e.grid_for(nq), VEC_THREADS, 0, e.stream>>>(e.view(), dq, nq, idx ? di : nullptr, e.d_w, p ? dp : nullptr);
    KFSP_TRY(e.check_launch());
    if (idx) KFSP_CUDA(cudaMemcpyAsync(idx, di, sizeof(int32_t) * nq, cudaMemcpyDeviceToHost, e.stream));
    if (p) KFSP_CUDA(cudaMemcpyAsync(p, dp, sizeof(double) * nq, cudaMemcpyDeviceToHost, e.stream));
    return e.sync();
}
int kfsp_fsp_index(kfsp_handle h, int64_t nq, const int32_t* states, int32_t* index_out) {
    if (!index_out) return KFSP_ERR_ARG;
    return lookup_common(h, nq, states, index_out, nullptr);
}
int kfsp_fsp_probability(kfsp_handle h, int64_t nq, const int32_t* states, double* p_out) {
    if (!p_out) return KFSP_ERR_ARG;
    return lookup_common(h, nq, states, nullptr, p_out);
}

// ------------------------------------------------------------------ solve
int kfsp_solve_resident(kfsp_handle h, double t, double fsp_tol, double kry_tol, int32_t verbosity, kfsp_stats* stats) {
    if (!h) return KFSP_ERR_ARG;
    return h->e.solve(t, fsp_tol, kry_tol, verbosity, stats);
}
int kfsp_solve(kfsp_handle h, double t, int64_t n_in, const int32_t* states_in, const double* p_in, double fsp_tol, double kry_tol,
               int32_t verbosity, int64_t* n_out, int32_t* states_out, double* p_out, int64_t max_out, kfsp_stats* stats) {
    if (!h || !states_in || !p_in || !n_out || n_in < 1) return KFSP_ERR_ARG;
    Engine& e = h->e;
    const double w0 = wall_now();
    if (verbosity) std::printf(" CALLING DGEXPV_FSP\n");                              // KrylovSolver.f90:32
    cudaSetDevice(e.device);
    static const bool dbg = std::getenv("KFSP_DEBUG_E2E") != nullptr;       // phase stamps of this call on stderr
    KFSP_TRY(e.fsp_init(n_in, states_in, /*defer_check=*/true));
    const double w1 = dbg ? (e.sync(), wall_now()) : 0.0;
    // Partitioned handle (kfsp_dist_init): states_in / p_in are the GLOBAL list and vector on every rank; this rank keeps
    // rows [lo, hi) and returns them (n_out = hi - lo).  fsp_init has zeroed W.
    // An ADAPTIVE handle on several GPUs keeps the whole state space on every rank and returns the whole result on every rank.
    const bool part = e.dist.nranks > 1 && !e.dist.repl;
    const int64_t lo = part ? e.dist.lo : 0;
    if (n_in > lo) {
        const int64_t cnt = std::min<int64_t>(n_in - lo, e.n);
        KFSP_CUDA(cudaMemcpyAsync(e.d_w, p_in + lo, sizeof(double) * cnt, cudaMemcpyHostToDevice, e.stream));
    }
    KFSP_TRY(e.box_check_overlap());                       // lattice variant: the state list is uploaded and verified beside the solve
    const double w2 = dbg ? wall_now() : 0.0;
    int st = e.solve(t, fsp_tol, kry_tol, verbosity, stats);
    const double w3 = dbg ? wall_now() : 0.0;
    {
        const int chk = e.box_check_finish();               // ... and a list that is not the lattice voids the result
        if (chk != KFSP_OK) { *n_out = 0; return chk; }
    }
    const double w4 = dbg ? wall_now() : 0.0;
    *n_out = e.n;
    if (st == KFSP_OK || st == KFSP_IFLAG_MXSTEP) {
        if (e.n > max_out) return KFSP_ERR_OUT_TOO_SMALL;
        // FSP_OUT supplies the initial states and receives the final ones (KrylovSolver.f90:7-36): when the caller passes the
        // same array for both and the state set cannot have changed, the list is already what it would receive
        const bool fixed_set = !e.opt.enable_expand && !e.opt.enable_drop && e.opt.n_init_onestep == 0;
        const bool in_place = states_out == states_in && fixed_set;
        if (states_out && !in_place) {
            if (e.box) KFSP_TRY(e.box_states());
            const int64_t row0 = (part && !e.box) ? lo : 0;          // the explicit partitioned variant keeps the global list
            KFSP_CUDA(cudaMemcpyAsync(states_out, e.d_states + row0 * e.S, sizeof(int32_t) * e.n * e.S, cudaMemcpyDeviceToHost, e.stream));
        }
        if (p_out) KFSP_CUDA(cudaMemcpyAsync(p_out, e.d_w, sizeof(double) * e.n, cudaMemcpyDeviceToHost, e.stream));
        KFSP_TRY(e.sync());
    }
    if (stats) stats->wall_seconds = wall_now() - w0;
    if (dbg)
        std::fprintf(stderr, "kfsp_solve rank %d: init %.2f ms, enqueue uploads %.2f, solve %.2f (device %.2f), list check %.2f, download %.2f, total %.2f\n",
                     e.dist.rank, 1e3 * (w1 - w0), 1e3 * (w2 - w1), 1e3 * (w3 - w2), stats ? 1e3 * stats->device_seconds : 0.0, 1e3 * (w4 - w3),
                     1e3 * (wall_now() - w4), 1e3 * (wall_now() - w0));
    return st;
}
int kfsp_trace_length(kfsp_handle h, int64_t* n) {
    if (!h || !n) return KFSP_ERR_ARG;
    *n = (int64_t)h->e.trace.size();
    return KFSP_OK;
}
int kfsp_trace_get(kfsp_handle h, kfsp_trace_row* rows, int64_t cap) {
    if (!h || !rows) return KFSP_ERR_ARG;
    const int64_t cnt = std::min<int64_t>(cap, (int64_t)h->e.trace.size());
    if (cnt > 0) std::memcpy(rows, h->e.trace.data(), sizeof(kfsp_trace_row) * cnt);
    return KFSP_OK;
}

// ------------------------------------------------------------------ kernels, individually
int kfsp_matvec(kfsp_handle h, const double* x, double* y) {
    if (!h || !x || !y) return KFSP_ERR_ARG;
    Engine& e = h->e;
    if (e.n < 1) return KFSP_ERR_BAD_SIZES;
    cudaSetDevice(e.device);
    const size_t a = Engine::align_up(sizeof(double) * e.n);
    KFSP_TRY(e.ensure_scratch(2 * a));
    double* dx = (double*)e.d_scratch;
    double* dy = (double*)(e.d_scratch + a);
    KFSP_CUDA(cudaMemcpyAsync(dx, x, sizeof(double) * e.n, cudaMemcpyHostToDevice, e.stream));
    const bool was = e.dist.suspended;
    if (e.dist.repl) e.dist.suspended = true;                      // adaptive multi-GPU handle: whole vector in and out on every rank
    const int st = e.spmv<0>(dx, dy);
    e.dist.suspended = was;
    KFSP_TRY(st);
    KFSP_CUDA(cudaMemcpyAsync(y, dy, sizeof(double) * e.n, cudaMemcpyDeviceToHost, e.stream));
    return e.sync();
}
int kfsp_matvec_device(kfsp_handle h, const double* xd, double* yd, int32_t reps, double* seconds) {
    if (!h || !xd || !yd || reps < 1) return KFSP_ERR_ARG;
    Engine& e = h->e;
    if (e.n < 1) return KFSP_ERR_BAD_SIZES;
    cudaSetDevice(e.device);
    KFSP_CUDA(cudaEventRecord(e.ev_a, e.stream));
    const bool was = e.dist.suspended;
    if (e.dist.repl) e.dist.suspended = true;
    int st = KFSP_OK;
    for (int r = 0; r < reps && st == KFSP_OK; ++r) st = e.spmv<0>(xd, yd);
    e.dist.suspended = was;
    KFSP_TRY(st);
    KFSP_CUDA(cudaEventRecord(e.ev_b, e.stream));
    KFSP_CUDA(cudaEventSynchronize(e.ev_b));
    float ms = 0.f;
    KFSP_CUDA(cudaEventElapsedTime(&ms, e.ev_a, e.ev_b));
    if (seconds) *seconds = 1e-3 * ms / reps;
    return KFSP_OK;
}
int kfsp_arnoldi(kfsp_handle h, const double* v, int32_t m, double* H_out, double* avnorm, int32_t* breakdown, double* seconds) {
    if (!h || !v || !H_out || m < 1 || m > h->e.opt.m_max) return KFSP_ERR_ARG;
    Engine& e = h->e;
    if (e.n < 1) return KFSP_ERR_BAD_SIZES;
    cudaSetDevice(e.device);
    KFSP_TRY(e.ensure_basis());
    KFSP_CUDA(cudaMemcpyAsync(e.d_w, v, sizeof(double) * e.n, cudaMemcpyHostToDevice, e.stream));
    GpuBackend be(e);
    double wsum, wssq;
    KFSP_TRY(be.norms(&wsum, &wssq));
    KFSP_CUDA(cudaEventRecord(e.ev_a, e.stream));
    KFSP_TRY(be.begin_step(1.0 / wssq));
    KFSP_TRY(e.arnoldi(1, m));
    k_set_entry<<<1, 1, 0, e.stream>>>(e.d_H + (size_t)m * e.LDH + m + 1, 1.0);
    KFSP_TRY(e.check_launch());
    KFSP_CUDA(cudaEventRecord(e.ev_b, e.stream));
    KFSP_TRY(e.read_ctl());
    float ms = 0.f;
    KFSP_CUDA(cudaEventElapsedTime(&ms, e.ev_a, e.ev_b));
    if (seconds) *seconds = 1e-3 * ms;
    std::vector<double> Hh((size_t)e.LDH * e.LDH);
    KFSP_CUDA(cudaMemcpyAsync(Hh.data(), e.d_H, sizeof(double) * Hh.size(), cudaMemcpyDeviceToHost, e.stream));
    KFSP_TRY(e.sync());
    const int mh = m + 2;
    for (int j = 0; j < mh; ++j)
        for (int i = 0; i < mh; ++i) H_out[(size_t)j * mh + i] = Hh[(size_t)j * e.LDH + i];
    if (avnorm) *avnorm = e.h_ctl->scal[SC_AVNORM];
    if (breakdown) *breakdown = e.h_ctl->brk;
    return KFSP_OK;
}
int kfsp_expm(kfsp_handle h, int32_t m, double t, const double* H, int32_t ldh, double* out, int32_t* ns, double* hnorm) {
    if (!h || !H || !out || m < 1 || m > h->e.LDH || ldh < m) return KFSP_ERR_ARG;
    Engine& e = h->e;
    cudaSetDevice(e.device);
    std::vector<double> Hh((size_t)e.LDH * e.LDH, 0.0);
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) Hh[(size_t)j * e.LDH + i] = H[(size_t)j * ldh + i];
    KFSP_CUDA(cudaMemcpyAsync(e.d_H, Hh.data(), sizeof(double) * Hh.size(), cudaMemcpyHostToDevice, e.stream));
    KFSP_TRY(e.launch_expm(m, t, 0, 0.0, -1, (const SweepCtl*)nullptr, e.d_expm_full));
    KFSP_CUDA(cudaMemcpyAsync(e.h_res, e.d_res, sizeof(ExpmResult), cudaMemcpyDeviceToHost, e.stream));
    KFSP_CUDA(cudaMemcpyAsync(out, e.d_expm_full, sizeof(double) * m * m, cudaMemcpyDeviceToHost, e.stream));
    KFSP_TRY(e.sync());
    if (ns) *ns = e.h_res->ns;
    if (hnorm) *hnorm = e.h_res->hnorm;
    return e.h_res->info;
}
int kfsp_combine(kfsp_handle h, int64_t n, int32_t mx, double beta, const double* V, const double* ev, const double* colscale,
                 double* w, double* wsum, double* wssq) {
    if (!h || !V || !ev || !w || n < 1 || mx < 1 || mx > MAX_COLS) return KFSP_ERR_ARG;
    Engine& e = h->e;
    cudaSetDevice(e.device);
    const size_t a = Engine::align_up(sizeof(double) * n * mx), b = Engine::align_up(sizeof(double) * EXPM_MAXN), c = Engine::align_up(sizeof(double) * n);
    KFSP_TRY(e.ensure_scratch(a + b + c));
    double* dV = (double*)e.d_scratch;
    double* de = (double*)(e.d_scratch + a);
    double* dw = (double*)(e.d_scratch + a + b);
    KFSP_CUDA(cudaMemcpyAsync(dV, V, sizeof(double) * n * mx, cudaMemcpyHostToDevice, e.stream));
    KFSP_CUDA(cudaMemcpyAsync(de, ev, sizeof(double) * mx, cudaMemcpyHostToDevice, e.stream));
    k_reset_ctl<<<1, 128, 0, e.stream>>>(e.d_ctl);            // unit column scales unless the caller gives them
    KFSP_TRY(e.check_launch());
    if (colscale) KFSP_CUDA(cudaMemcpyAsync(e.d_ctl->colscale, colscale, sizeof(double) * mx, cudaMemcpyHostToDevice, e.stream));
    k_combine<<<e.grid_for(n), VEC_THREADS, 0, e.stream>>>(n, n, mx, beta, dV, de, dw, e.rd, e.d_ctl);
    KFSP_TRY(e.check_launch());
    KFSP_CUDA(cudaMemcpyAsync(w, dw, sizeof(double) * n, cudaMemcpyDeviceToHost, e.stream));
    KFSP_TRY(e.read_ctl());
    if (wsum) *wsum = e.h_ctl->scal[SC_WSUM];
    if (wssq) *wssq = e.h_ctl->scal[SC_WSSQ];
    k_reset_ctl<<<1, 128, 0, e.stream>>>(e.d_ctl);
    return e.check_launch();
}

// ------------------------------------------------------------------ multi-GPU plumbing
int kfsp_dist_unique_id(uint8_t id[KFSP_NCCL_ID_BYTES]) {
#ifdef KFSP_WITH_NCCL
    ncclUniqueId u;
    if (ncclGetUniqueId(&u) != ncclSuccess) return KFSP_ERR_NCCL;
    static_assert(sizeof(u) <= KFSP_NCCL_ID_BYTES, "nccl id size");
    std::memset(id, 0, KFSP_NCCL_ID_BYTES);
    std::memcpy(id, &u, sizeof u);
    return KFSP_OK;
#else
    (void)id;
    return KFSP_ERR_UNSUPPORTED;
#endif
}
int kfsp_dist_init(kfsp_handle h, int32_t rank, int32_t nranks, const uint8_t id[KFSP_NCCL_ID_BYTES]) {
    if (!h || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id)) return KFSP_ERR_ARG;
    return h->e.dist_init(rank, nranks, id);
}
int kfsp_dist_partition(int64_t n, int32_t nranks, int32_t rank, int64_t* lo, int64_t* hi) {
    if (n < 0 || nranks < 1 || rank < 0 || rank >= nranks || !lo || !hi) return KFSP_ERR_ARG;
    *lo = part_lo(n, nranks, rank);
    *hi = part_lo(n, nranks, rank + 1);
    return KFSP_OK;
}
// host-side arithmetic of the replicated layout of ADAPTIVE sets on several GPUs (Engine::repartition; no GPU needed): the rows
// `rank` computes in the Krylov loop for a state set of n rows, and whether the set is still kept whole on every rank
int kfsp_repl_partition(int64_t n, int32_t nranks, int32_t rank, int64_t min_rows, int64_t* lo, int64_t* hi, int32_t* whole) {
    if (n < 0 || nranks < 1 || rank < 0 || rank >= nranks || !lo || !hi || !whole) return KFSP_ERR_ARG;
    *whole = (nranks == 1 || n < min_rows) ? 1 : 0;
    *lo = *whole ? 0 : part_lo(n, nranks, rank);
    *hi = *whole ? n : part_lo(n, nranks, rank + 1);
    return KFSP_OK;
}
int kfsp_dist_owner(int64_t n, int32_t nranks, int64_t row, int32_t* owner) {
    if (n < nranks || nranks < 1 || row < 0 || row >= n || !owner) return KFSP_ERR_ARG;
    *owner = part_owner(n, nranks, row);
    return KFSP_OK;
}
// host-side arithmetic of the lattice variant (no GPU needed): slab of the slowest species owned by `rank`
int kfsp_lattice_partition(int32_t nz, int32_t nranks, int32_t rank, int32_t* zlo, int32_t* zhi) {
    if (nz < 1 || nranks < 1 || nz < nranks || rank < 0 || rank >= nranks || !zlo || !zhi) return KFSP_ERR_ARG;
    *zlo = (int32_t)((int64_t)nz * rank / nranks);
    *zhi = (int32_t)((int64_t)nz * (rank + 1) / nranks);
    return KFSP_OK;
}
// which SpMV kernel a lattice model gets: *kind = 0 generic k_spmv_box, 1 / 2 = k_spmv_bd2 with reaction order 0 / 1;
// *table_mask bit k = reaction k's propensity table runs over the slowest species.  stoich is species-fastest (S*R),
// table_species[k] is the one species reaction k's propensity reads.
int kfsp_lattice_kernel(int32_t S, int32_t R, const int32_t* stoich, const int32_t* table_species, int32_t* kind, int32_t* table_mask) {
    if (S < 2 || S > KFSP_MAX_SPECIES || R < 1 || R > BOX_MAX_R || !stoich || !table_species || !kind) return KFSP_ERR_ARG;
    Lattice L;
    std::memset(&L, 0, sizeof L);
    L.S = S; L.R = R;
    int32_t mask = 0;
    for (int k = 0; k < R; ++k) {
        if (table_species[k] < 0 || table_species[k] >= S) return KFSP_ERR_UNSUPPORTED;
        for (int s = 0; s < S; ++s) L.nu[k][s] = stoich[k * S + s];
        L.sp[k] = table_species[k];
        if (table_species[k] == S - 1) mask |= 1 << k;
    }
    *kind = lattice_bd2_order(L) + 1;
    if (table_mask) *table_mask = mask;
    return KFSP_OK;
}
int kfsp_dist_info(kfsp_handle h, int64_t* lo, int64_t* hi, int64_t* n_halo, int64_t* n_send, int64_t* halo_bytes, int64_t* reductions) {
    if (!h) return KFSP_ERR_ARG;
    const Dist& d = h->e.dist;
    if (lo) *lo = d.nranks > 1 ? d.lo : 0;
    if (hi) *hi = d.nranks > 1 ? d.hi : h->e.n;
    if (n_halo) *n_halo = d.n_halo;
    if (n_send) *n_send = d.n_send;
    if (halo_bytes) *halo_bytes = d.halo_bytes;
    if (reductions) *reductions = d.reductions;
    return KFSP_OK;
}

// How the propensities of the model last given to kfsp_set_model are evaluated (canonical arithmetic)
int kfsp_model_info(kfsp_handle h, int32_t* n_tabulated, int32_t* n_host_evaluated, int32_t* n_device_libm, int32_t* factored) {
    if (!h) return KFSP_ERR_ARG;
    const Engine& e = h->e;
    if (!e.have_model) return KFSP_ERR_NO_MODEL;
    if (n_tabulated) *n_tabulated = e.n_tabulated;
    if (n_host_evaluated) *n_host_evaluated = e.n_host_evaluated;
    if (n_device_libm) *n_device_libm = e.n_inexact_on_device;
    if (factored) *factored = e.idx ? 1 : 0;
    return KFSP_OK;
}
int kfsp_dist_exchange_stats(kfsp_handle h, int64_t* exchanges, double* mean_us, double* max_us, int32_t reset) {
    if (!h) return KFSP_ERR_ARG;
    Engine& e = h->e;
    unsigned long long v[4] = {0, 0, 0, 0};
    if (e.dist.d_stat) {
        cudaSetDevice(e.device);
        KFSP_CUDA(cudaMemcpyAsync(v, e.dist.d_stat, sizeof v, cudaMemcpyDeviceToHost, e.stream));
        KFSP_TRY(e.sync());
        if (reset) KFSP_CUDA(cudaMemsetAsync(e.dist.d_stat, 0, sizeof v, e.stream));
    }
    if (exchanges) *exchanges = (int64_t)v[0];
    if (mean_us) *mean_us = v[0] ? 1e-3 * (double)v[1] / (double)v[0] : 0.0;
    if (max_us) *max_us = 1e-3 * (double)v[2];
    return KFSP_OK;
}

// ------------------------------------------------------------------ device helpers
int kfsp_device_alloc(kfsp_handle h, int64_t bytes, void** ptr) {
    if (!h || !ptr || bytes < 1) return KFSP_ERR_ARG;
    cudaSetDevice(h->e.device);
    KFSP_CUDA(cudaMalloc(ptr, (size_t)bytes));
    return KFSP_OK;
}
int kfsp_device_free(kfsp_handle h, void* ptr) {
    if (!h) return KFSP_ERR_ARG;
    cudaSetDevice(h->e.device);
    KFSP_CUDA(cudaFree(ptr));
    return KFSP_OK;
}
int kfsp_device_upload(kfsp_handle h, void* dst, const void* src, int64_t bytes) {
    if (!h || !dst || !src || bytes < 0) return KFSP_ERR_ARG;
    cudaSetDevice(h->e.device);
    KFSP_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, h->e.stream));
    return h->e.sync();
}
int kfsp_device_download(kfsp_handle h, void* dst, const void* src, int64_t bytes) {
    if (!h || !dst || !src || bytes < 0) return KFSP_ERR_ARG;
    cudaSetDevice(h->e.device);
    KFSP_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, h->e.stream));
    return h->e.sync();
}
int kfsp_device_vector(kfsp_handle h, double** p) {
    if (!h || !p) return KFSP_ERR_ARG;
    *p = h->e.d_w;
    return h->e.d_w ? KFSP_OK : KFSP_ERR_BAD_SIZES;
}
int kfsp_flush_l2(kfsp_handle h) {
    if (!h) return KFSP_ERR_ARG;
    Engine& e = h->e;
    cudaSetDevice(e.device);
    const size_t bytes = 512u << 20;                       // > 126 MB of L2, written once
    if (!e.d_flush) { KFSP_CUDA(cudaMalloc(&e.d_flush, bytes)); e.flush_bytes = bytes; }
    KFSP_CUDA(cudaMemsetAsync(e.d_flush, 0, e.flush_bytes, e.stream));
    return e.sync();
}
int kfsp_set_blocking_sync(kfsp_handle h, int32_t on) {
    if (!h) return KFSP_ERR_ARG;
    h->e.blocking_sync = on != 0;
    return KFSP_OK;
}
int kfsp_set_profiling(kfsp_handle h, int32_t on) {
    if (!h) return KFSP_ERR_ARG;
    cudaSetDevice(h->e.device);
    return h->e.set_profiling(on);
}
int kfsp_fsp_set_vector_device(kfsp_handle h, const double* src, int64_t cnt) {
    if (!h || !src || cnt < 0) return KFSP_ERR_ARG;
    Engine& e = h->e;
    if (e.ld == 0 || cnt > e.ld) return KFSP_ERR_BAD_SIZES;
    cudaSetDevice(e.device);
    if (cnt < e.n) KFSP_CUDA(cudaMemsetAsync(e.d_w + cnt, 0, sizeof(double) * (e.n - cnt), e.stream));
    KFSP_CUDA(cudaMemcpyAsync(e.d_w, src, sizeof(double) * cnt, cudaMemcpyDeviceToDevice, e.stream));
    return KFSP_OK;
}
int kfsp_phase_seconds(kfsp_handle h, double out[8]) {
    if (!h || !out) return KFSP_ERR_ARG;
    for (int i = 0; i < 8; ++i) out[i] = h->e.phase_s[i];
    out[6] = (double)h->e.host_prop_rounds;
    out[7] = (double)h->e.host_prop_evals;
    return KFSP_OK;
}
int kfsp_spmv_launch_counts(kfsp_handle h, int64_t out[4]) {
    if (!h || !out) return KFSP_ERR_ARG;
    for (int i = 0; i < 3; ++i) out[i] = h->e.spmv_by_mode[i];
    out[3] = h->e.spmv_fused;
    return KFSP_OK;
}
int kfsp_profile_get(kfsp_handle h, double seconds[KFSP_PROF_CLASSES], int64_t launches[KFSP_PROF_CLASSES], int64_t bytes_per_state[KFSP_PROF_CLASSES]) {
    if (!h || !seconds || !launches) return KFSP_ERR_ARG;
    for (int i = 0; i < KFSP_PROF_CLASSES; ++i) {
        seconds[i] = h->e.prof_sec[i];
        launches[i] = h->e.prof_cnt[i];
        if (bytes_per_state) bytes_per_state[i] = h->e.prof_bps[i];
    }
    return KFSP_OK;
}
int kfsp_launch_count(kfsp_handle h, int64_t* n) {
    if (!h || !n) return KFSP_ERR_ARG;
    *n = h->e.launches;
    return KFSP_OK;
}

}  // extern "C"
