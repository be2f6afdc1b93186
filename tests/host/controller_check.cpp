// TEST INFRASTRUCTURE.  Drives the product's host-side controller template
// (krylovfspssa_b200/csrc/controller.h) with a CPU backend built from the oracle's C API,
// so that the discrete decision logic of DGEXPV_FSP can be checked against the oracle's own
// trace without a GPU.  Compiled together with oracle/kfsp_oracle.cpp by tests/test_controller_host.py.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../krylovfspssa_b200/csrc/controller.h"

extern "C" {
long ko_fsp_size(void* fp);
int ko_fsp_onestep(void* fp);
void ko_fsp_ssa(void* fp, double timestep, void* rp);
int ko_fsp_drop(void* fp, double* w, double dsum, double* droptol, long* dropcount);
void ko_fsp_matvec(void* fp, const double* x, double* y);
int ko_fsp_matrix_starter(void* fp);
int ko_dgpadm(int ideg, int m, double t, const double* H, int ldh, double* out, int* ns, double* hnorm);
void* ko_rng_create(int mode, uint64_t seed, const char* path);
void ko_rng_free(void* rp);
}

namespace {

struct CpuBackend {
    void* fsp;
    void* rng;
    int m_max;
    std::vector<double> W, V, H, E;
    int LDH;
    int brk = 0;
    double avnorm = 0.0;
    double break_tol = 1e-7;

    int64_t size() { return ko_fsp_size(fsp); }
    void fit() { W.resize((size_t)size(), 0.0); }
    int onestep() { int rc = ko_fsp_onestep(fsp); fit(); return rc; }
    int ssa(double t) { ko_fsp_ssa(fsp, t, rng); fit(); return 0; }
    int drop(double dsum, int* dropped) {
        *dropped = ko_fsp_drop(fsp, W.data(), dsum, nullptr, nullptr);
        fit();
        return 0;
    }
    // DNRM2 as the oracle restates it (netlib scaled sum of squares), so that this backend
    // reproduces the oracle's arithmetic exactly and only the controller logic is under test
    static double nrm2(const double* x, size_t n) {
        if (n < 1) return 0.0;
        if (n == 1) return std::fabs(x[0]);
        double scale = 0.0, ssq = 1.0;
        for (size_t i = 0; i < n; ++i)
            if (x[i] != 0.0) {
                const double a = std::fabs(x[i]);
                if (scale < a) { const double r = scale / a; ssq = 1.0 + ssq * r * r; scale = a; }
                else { const double r = a / scale; ssq += r * r; }
            }
        return scale * std::sqrt(ssq);
    }
    int norms(double* wsum, double* wnrm2) {
        double a = 0;
        for (double x : W) a += std::fabs(x);
        *wsum = a; *wnrm2 = nrm2(W.data(), W.size());
        return 0;
    }
    int begin_step(double inv_beta) {
        const size_t n = W.size();
        V.assign(n * (size_t)(m_max + 2), 0.0);
        for (size_t i = 0; i < n; ++i) V[i] = inv_beta * W[i];
        std::fill(H.begin(), H.end(), 0.0);
        brk = 0;
        return 0;
    }
    int arnoldi(int jold, int m) {
        const size_t n = W.size();
        for (int J = jold; J <= m; ++J) {
            double* vj = &V[(size_t)(J - 1) * n];
            double* vn = &V[(size_t)J * n];
            ko_fsp_matvec(fsp, vj, vn);
            for (int I = std::max(1, J - 1); I <= J; ++I) {
                const double* vi = &V[(size_t)(I - 1) * n];
                double h = 0;
                for (size_t i = 0; i < n; ++i) h += vi[i] * vn[i];
                if (h != 0.0) for (size_t i = 0; i < n; ++i) vn[i] += (-h) * vi[i];
                H[(size_t)(J - 1) * LDH + (I - 1)] = h;
            }
            const double hn = nrm2(vn, n);
            if (hn <= break_tol) { brk = J; return 0; }
            H[(size_t)(J - 1) * LDH + J] = hn;
            { const double inv = 1.0 / hn; for (size_t i = 0; i < n; ++i) vn[i] = inv * vn[i]; }
        }
        ko_fsp_matvec(fsp, &V[(size_t)m * n], &V[(size_t)(m + 1) * n]);
        avnorm = nrm2(&V[(size_t)(m + 1) * n], n);
        return 0;
    }
    int expm(int mx_ok, double t_ok, int use_brk, double t_brk, int set_one, kfsp::StepScalars* out) {
        if (set_one >= 0) H[(size_t)set_one * LDH + set_one + 1] = 1.0;
        int mx = mx_ok;
        double t = t_ok;
        if (use_brk && brk > 0) { mx = brk; t = t_brk; }
        E.assign((size_t)mx * mx, 0.0);
        int ns = 0;
        double hn = 0;
        int rc = ko_dgpadm(6, mx, t, H.data(), LDH, E.data(), &ns, &hn);
        if (rc) return rc;
        out->ns = ns; out->brk = use_brk ? brk : 0; out->mx = mx; out->hnorm = hn; out->avnorm = avnorm; out->e = E.data();
        return 0;
    }
    int clear_h(int row0, int col0) { H[(size_t)col0 * LDH + row0] = 0.0; return 0; }
    int combine(int mx, double beta, double* wsum, double* wnrm2) {
        const size_t n = W.size();
        for (size_t i = 0; i < n; ++i) W[i] = 0.0;
        for (int j = 0; j < mx; ++j) {
            const double t = beta * E[j];
            for (size_t i = 0; i < n; ++i) W[i] += t * V[(size_t)j * n + i];
        }
        double a = 0;
        for (size_t i = 0; i < n; ++i) { if (W[i] < 0) W[i] = 0; a += std::fabs(W[i]); }
        *wsum = a; *wnrm2 = nrm2(W.data(), n);
        return 0;
    }
    int restore_w(double beta, double* wnrm2) {
        for (size_t i = 0; i < W.size(); ++i) W[i] = beta * V[i];
        *wnrm2 = nrm2(W.data(), W.size());
        return 0;
    }
};

}  // namespace

extern "C" {

// fsp: oracle Fsp with the caller's states set (ko_fsp_set_states).  Returns the controller status.
int cc_run(void* fsp, const double* p_in, long n_in, double T, double fsptol, double krytol, uint64_t seed, int R,
           int m_max, int m_min, int n_init_onestep, int enable_drop, int enable_expand,
           kfsp_trace_row* rows, long cap, long* nrows, kfsp_stats* stats, double* w_out, long w_cap) {
    kfsp_options o;
    std::memset(&o, 0, sizeof o);
    o.m_max = m_max; o.m_min = m_min; o.ideg = 6; o.n_init_onestep = n_init_onestep; o.fsp_reject_limit = 5;
    o.enable_drop = enable_drop; o.enable_expand = enable_expand; o.max_molecules = 10000; o.max_states = 6291469;
    o.delta = 1.2; o.gamma = 0.9; o.break_tol = 1e-7; o.drop_tol0 = 1e-8; o.drop_deriv_tol = 1e-8; o.drop_fraction = 0.1;
    o.seed = seed;
    int rc = ko_fsp_matrix_starter(fsp);
    if (rc) return rc;
    CpuBackend be;
    be.fsp = fsp;
    be.rng = ko_rng_create(1, seed, nullptr);
    be.m_max = m_max;
    be.LDH = m_max + 2;
    be.H.assign((size_t)be.LDH * be.LDH, 0.0);
    be.fit();
    for (long i = 0; i < n_in && i < (long)be.W.size(); ++i) be.W[i] = p_in[i];
    std::vector<kfsp_trace_row> trace;
    kfsp::Controller ctl(o);
    rc = ctl.run(be, R, T, fsptol, krytol, 0, stats, trace);
    *nrows = (long)trace.size();
    for (long i = 0; i < *nrows && i < cap; ++i) rows[i] = trace[i];
    for (long i = 0; i < (long)be.W.size() && i < w_cap; ++i) w_out[i] = be.W[i];
    ko_rng_free(be.rng);
    return rc;
}

}
