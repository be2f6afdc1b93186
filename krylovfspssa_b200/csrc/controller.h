// The time-stepping controller of DGEXPV_FSP (src/fsp/KrylovSolver.f90:127-573) as a
// host-side, CUDA-free template over a numeric backend.  The backend owns every N-sized
// array and every dense matrix (on the GPU in libkfsp.so); the controller sees only the
// handful of scalars each phase produces, and takes the discrete decisions: accept/reject
// a step, change the Krylov dimension, shrink the step for the FSP criterion, expand or
// prune the state space.  Its arithmetic follows the reference statement by statement,
// including the quirks listed in SURVEY.md (a2, a9): 2-significant-digit rounding of step
// sizes, default-INTEGER products and a REAL (fp32) result in KRYLOV_COST, HNORM already
// containing the step size.
//
// Backend concept (every call returns a kfsp_status):
//   int64_t size();
//   int onestep();                      ONESTEP_EXTENDER
//   int ssa(double timestep);           SSA_EXTENDER
//   int drop(double dsum, int* dropped) DROP_STATES on W
//   int norms(double* wsum, double* wnrm2)          ||W||_1, ||W||_2
//   int begin_step(double inv_beta)     V(:,1) = W/BETA ; H = 0 ; clear breakdown flag
//   int arnoldi(int jold, int m)        columns jold..m of the IOP-2 sweep + the extra product
//   int expm(int mx_ok, double t_ok, int use_brk, double t_brk, int set_one, StepScalars* out)
//   int clear_h(int row0, int col0)     H(row0+1, col0+1) = 0
//   int combine(int mx, double beta, double* wsum, double* wnrm2)  W = BETA*V(:,1:mx)*e, clamp, norms
//   int restore_w(double beta, double* wnrm2)                      W = BETA*V(:,1)
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../include/kfsp.h"

namespace kfsp {

struct StepScalars {
    int ns = 0;
    int brk = 0;              // happy-breakdown column, 0 if none
    int mx = 0;
    double hnorm = 0.0;
    double avnorm = 0.0;
    const double* e = nullptr;   // first column of exp(t*H), at least mx entries (host memory)
};

namespace fortran {
inline int nint(double x) { return (int)(int32_t)(uint32_t)(uint64_t)std::llround(x); }
inline int to_int(double x) {                      // INT()/CEILING() overflow gives the x86 "integer indefinite"
    if (!(x > -2147483649.0 && x < 2147483648.0)) return INT32_MIN;
    return (int)x;
}
inline int ceiling(double x) { return to_int(std::ceil(x)); }
inline int32_t wrap(int64_t v) { return (int32_t)(uint32_t)(uint64_t)v; }
inline double powi(double x, int m) {              // real**integer as gfortran evaluates it (libgcc __powidf2)
    unsigned n = m < 0 ? 0u - (unsigned)m : (unsigned)m;
    double y = (n & 1u) ? x : 1.0;
    while (n >>= 1) {
        x *= x;
        if (n & 1u) y *= x;
    }
    return m < 0 ? 1.0 / y : y;
}
}  // namespace fortran

class Controller {
  public:
    Controller(const kfsp_options& o) : opt(o) {}

    template <class B>
    int run(B& be, int R, double T, double fsptol, double krytol_in, int itrace, kfsp_stats* st, std::vector<kfsp_trace_row>& trace) {
        using namespace fortran;
        const int M_MAX = opt.m_max, M_MIN = opt.m_min;
        const double DELTA = opt.delta, GAMMA = opt.gamma;
        trace.clear();
        int status = KFSP_OK;
        for (int i = 0; i < opt.n_init_onestep; ++i) {                                // KrylovSolver.f90:132-134
            status = be.onestep();
            if (status) return status;
        }
        if (M_MIN >= opt.max_states || M_MIN <= 0) return KFSP_ERR_BAD_SIZES;          // :145

        // machine epsilon the EXPOKIT way (:166-170)
        double p1 = 4.0 / 3.0, p2, p3, eps;
        do { p2 = p1 - 1.0; p3 = p2 + p2 + p2; eps = std::fabs(p3 - 1.0); } while (eps == 0.0);
        const double krytol = krytol_in <= eps ? std::sqrt(eps) : krytol_in;
        const double rndoff = eps * 1.0;
        const double t_out = std::fabs(T), sgn = T < 0 ? -1.0 : 1.0;
        const double sqr1 = std::sqrt(0.1);
        auto round2 = [&](double x, double bias) {                                    // 2-significant-digit rounding
            const double q = powi(10.0, nint(std::log10(x) - sqr1) - 1);
            return std::trunc(x / q + bias) * q;
        };

        double wsum = 0.0, wnrm2 = 0.0;
        status = be.norms(&wsum, &wnrm2);
        if (status) return status;
        double beta = wnrm2;
        const double vnorm = beta;
        double hump = beta;

        int m = M_MIN;
        {                                                                             // first step size (:182-187)
            const double xm = 1.0 / (double)m;
            const double q = krytol * powi((m + 1) / 2.72, m + 1) * std::sqrt(2.0 * 3.14 * (m + 1));
            t_new = (1.0 / 1.0) * std::pow(q / (4.0 * beta * 1.0), xm);
            t_new = round2(t_new, 0.55);
        }
        int64_t n_now = be.size();
        int iexpand = 0, irejectfsp = 0, imreject = 0, jold = 1, m_new = m;
        double wsum_old = 1.0;
        nnz = wrap((int64_t)(R + 1) * n_now);
        bool orderold = true, kestold = true, m_changed = false;
        double omega = 0.0, omega_old = 0.0, t_old = 0.0, order = 0.0, k_factor = 2.0;
        int m_old = -1;
        double t_now = 0.0, t_step = 0.0, err_loc = 0.0, hnorm = 0.0;
        double error = 0.0, errorold = 1.0, tau_old = 1.0, fsporder = 2.0;
        int nmult = 0, nexph = 0, nscale = 0, nstep = 0, nreject = 0, ibrkflag = 0, mbrkdwn = m;
        double tbrkdwn = 0.0, step_min = t_out, step_max = 0.0, s_error = 0.0, x_error = 0.0;
        int64_t n_expand = 0, n_drop = 0, n_max = n_now;

        while (t_now < t_out) {                                                       // label 100
            t_step = std::min(t_out - t_now, t_new);
            const int64_t n = n_now;
            m = (int)std::min<int64_t>(n - 1, m_new);
            mbrkdwn = m;
            int k1 = 2;
            nstep += 1;
            int flags = 0;
            status = be.begin_step(1.0 / beta);
            if (status) return status;
            int ireject = 0;
            bool to_404 = false;
            StepScalars sc;
            bool sweep_pending = true;
            int mx = 0;
            for (;;) {                                                                // labels 101 / 401
                if (sweep_pending) {
                    status = be.arnoldi(jold, m);                                     // :236-263
                    if (status) return status;
                    sweep_pending = false;
                    nexph += 1;
                    status = be.expm(m + 2, sgn * t_step, 1, sgn * (t_out - t_now), m, &sc);   // :266-277
                    if (status) return status;
                    if (sc.brk > 0) {                                                 // happy breakdown (:249-256)
                        nmult += sc.brk - jold + 1;
                        k1 = 0; ibrkflag = 1; mbrkdwn = sc.brk; tbrkdwn = t_now; t_step = t_out - t_now;
                        flags |= 8;
                    } else {
                        nmult += m - jold + 2;
                    }
                } else {
                    nexph += 1;
                    status = be.expm(mbrkdwn + k1, sgn * t_step, 0, 0.0, -1, &sc);      // label 401
                    if (status) return status;
                }
                mx = mbrkdwn + k1;
                nscale += sc.ns;
                hnorm = sc.hnorm;
                // label 402: local error estimate (:290-305)
                if (k1 == 0) {
                    err_loc = krytol;
                } else {
                    p1 = std::fabs(sc.e[m]) * beta;
                    p2 = std::fabs(sc.e[m + 1]) * beta * sc.avnorm;
                    if (p1 > 10.0 * p2) err_loc = p2;
                    else if (p1 > p2) err_loc = (p1 * p2) / (p1 - p2);
                    else err_loc = p1;
                }
                if (std::isnan(err_loc)) { t_step = t_step / 5.0; continue; }          // :307-310
                omega_old = omega;
                omega = err_loc / (krytol * t_step);
                if ((m == m_old) && (t_step != t_old) && (ireject >= 1)) {            // order (:316-324)
                    order = std::max(1.0, std::log(omega / omega_old) / std::log(t_step / t_old));
                    orderold = false;
                } else if (orderold || ireject == 0) {
                    order = (double)m / 4.0;
                    orderold = true;
                } else {
                    orderold = true;
                }
                if ((m != m_old) && (t_step == t_old) && (ireject >= 1)) {            // kappa (:326-334)
                    k_factor = std::max(1.1, std::pow(omega / omega_old, 1.0 / (double)(m_old - m)));
                    kestold = false;
                } else if (kestold || ireject == 0) {
                    kestold = true;
                    k_factor = 2.0;
                } else {
                    kestold = true;
                }
                t_old = t_step;
                m_old = m;
                const double t_cand = std::min(t_out - t_now, std::max(t_step / 5.0,
                                               std::min(5.0 * t_step, GAMMA * t_step * std::pow(omega, -1.0 / order))));
                if (((m == M_MAX) && (omega > DELTA)) || (imreject > 4)) {            // :339-346
                    t_new = round2(t_cand, 0.0);
                    m_changed = false;
                } else {                                                               // :348-373
                    const int cl = ceiling(std::log(omega) / std::log(k_factor));
                    const int m_opt = std::min(std::min(std::max(std::max(M_MIN, 3 * m / 4), wrap((int64_t)m + cl)), M_MAX),
                                               ceiling(4.0 * m / 3.0) + 1);
                    const float cost1 = krylov_cost(t_now, t_out, t_cand, m, (int)n, hnorm);
                    const float cost2 = krylov_cost(t_now, t_out, t_step, m_opt, (int)n, hnorm);
                    if (cost1 <= cost2) {
                        t_new = round2(t_cand, 0.0);
                        m_new = m;
                        m_changed = false;
                    } else {
                        m_new = m_opt;
                        t_new = t_step;
                        m_changed = true;
                    }
                }
                if ((k1 != 0) && (omega > DELTA) && (opt.mxreject == 0 || ireject < opt.mxreject)) {   // :375-434
                    if (!m_changed) {
                        t_step = std::min(t_out - t_now, std::max(t_step / 5.0, std::min(5.0 * t_step, t_new)));
                        t_step = round2(t_step, 0.55);
                        ireject += 1;
                        nreject += 1;
                        if ((opt.mxreject != 0) && (ireject > opt.mxreject)) return KFSP_IFLAG_TOLERANCE;
                        continue;                                                      // GO TO 401
                    }
                    nreject += 1;
                    imreject += 1;
                    status = be.clear_h(m_old + 1, m_old);       // the copy loop of :419-424 leaves H(M_OLD+2,M_OLD+1) behind
                    if (status) return status;
                    m = m_new;
                    mbrkdwn = m;
                    k1 = 2;
                    t_step = std::min(t_out - t_now, t_new);
                    jold = m_old;
                    sweep_pending = true;
                    continue;                                                          // GO TO 101
                }
                break;
            }
            imreject = 0;
            jold = 1;
            if (err_loc < 1.0e-16) t_new = std::max(t_new, 2.0 * t_step);
            mx = mbrkdwn + std::max(0, k1 - 1);
            irejectfsp = 0;
            for (;;) {                                                                // FSP criterion (:442-495)
                status = be.combine(mx, beta, &wsum, &wnrm2);
                if (status) return status;
                error = wsum_old - wsum;
                if (wsum >= (1.0 - (t_now + t_step) * fsptol / t_out)) break;
                iexpand = 1;
                irejectfsp += 1;
                if (irejectfsp >= opt.fsp_reject_limit) {
                    status = be.restore_w(beta, &wnrm2);
                    if (status) return status;
                    nstep -= 1;
                    to_404 = true;
                    flags |= 4;
                    break;
                } else if (irejectfsp == 1) {
                    fsporder = 2;
                } else {
                    fsporder = std::log(error / errorold) / std::log(t_step / tau_old) - 1.0;
                }
                const double tfsp = GAMMA * t_step * std::pow(fsptol * t_step / (error * t_out), 1.0 / fsporder);
                errorold = error;
                tau_old = t_step;
                t_step = std::min(t_out - t_now, std::max(t_step / 5.0, std::min(0.9 * t_step, tfsp)));
                t_step = round2(t_step, 0.55);
                nexph += 1;
                status = be.expm(mx, sgn * t_step, 0, 0.0, -1, &sc);                    // DGPADM on the leading mx block
                if (status) return status;
                nscale += sc.ns;
            }
            bool finished = false;
            bool space_changed = false;
            if (!to_404) {
                t_now = t_now + t_step;
                wsum_old = wsum;
                if (itrace)
                    std::printf(" TIMESTEP %d  FSP SIZE = %lld  STEP_SIZE = %.6g  NEXT_STEP = %.6g  T_NOW = %.8g  KRYLOV DIMENSION = %d  WSUM = %.12g\n",
                                nstep, (long long)be.size(), t_step, t_new, t_now, m, wsum);
                if (t_now >= t_out) {
                    finished = true;
                } else if (nstep > 1 && iexpand != 1 && opt.enable_drop) {            // :509-512
                    const double dsum = wsum - (1.0 - t_now * fsptol / t_out);
                    if (dsum > 0.0) {
                        int dropped = 0;
                        status = be.drop(dsum, &dropped);
                        if (status) return status;
                        if (dropped) { flags |= 2; n_drop += 1; space_changed = true; }
                    }
                }
            }
            if (!finished && (iexpand == 1) && (t_now < t_out)) {                     // label 404 (:516-534)
                if (nstep == 1) t_new = t_step;
                const double t_ssa = std::min(t_new, t_out - t_now);
                if (opt.enable_expand) {
                    status = be.ssa(t_ssa);
                    if (status) return status;
                    status = be.onestep();
                    if (status) return status;
                    flags |= 1;
                    n_expand += 1;
                }
                iexpand = 0;
            }
            if (finished) {
                trace.push_back(kfsp_trace_row{t_now, t_step, t_new, wsum, err_loc, beta, m, (int32_t)n, (int32_t)be.size(), flags, nmult, nexph});
                break;
            }
            n_now = be.size();
            n_max = std::max(n_max, n_now);
            nnz = wrap((int64_t)(R + 1) * n_now);                                     // :537
            if (space_changed) {
                status = be.norms(&wsum, &wnrm2);                                    // BETA = DNRM2(N_NOW, W) (:540)
                if (status) return status;
            }
            beta = wnrm2;
            hump = std::max(hump, beta);
            err_loc = std::max(err_loc, rndoff);
            step_min = std::min(step_min, t_step);
            step_max = std::max(step_max, t_step);
            s_error = s_error + err_loc;
            x_error = std::max(x_error, err_loc);
            t_new = round2(t_new, 0.55);                                              // :547-548
            trace.push_back(kfsp_trace_row{t_now, t_step, t_new, wsum, err_loc, beta, m, (int32_t)n, (int32_t)n_now, flags, nmult, nexph});
            if (!((opt.mxstep == 0) || (nstep < opt.mxstep))) { status = KFSP_IFLAG_MXSTEP; break; }
        }
        if (st) {
            st->nmult = nmult; st->nexph = nexph; st->nscale = nscale; st->nstep = nstep; st->nreject = nreject;
            st->ibrkflag = ibrkflag; st->mbrkdwn = mbrkdwn; st->iflag = status;
            st->step_min = step_min; st->step_max = step_max; st->x_error = x_error; st->s_error = s_error;
            st->tbrkdwn = tbrkdwn; st->t_now = sgn * t_now; st->hump = hump / vnorm; st->beta_ratio = beta / vnorm;
            st->n_expand = n_expand; st->n_drop = n_drop; st->n_final = be.size(); st->n_max = std::max(n_max, (int64_t)be.size());
        }
        return status;
    }

  private:
    kfsp_options opt;
    int32_t nnz = 0;
    double t_new = 0.0;

    // KRYLOV_COST (:618-639)
    float krylov_cost(double t_now, double t_out, double tau, int m, int n, double hnorm) const {
        using namespace fortran;
        const int q = 2;                                                              // QIOP (:137)
        const double nom = 25.0 / 3.0 + (double)std::max(0, 2 + to_int(std::log(tau * hnorm) / std::log(2.0)));
        const int32_t a = wrap((int64_t)(2 * (m + 1)) * nnz);
        const int32_t b = wrap((int64_t)(5 * m + 4 * q * m + 2 * q - 2 * q * q + 7) * n);
        const double inner = (double)wrap((int64_t)a + b) + 2 * nom * (m + 2) * (m + 2) * (m + 2);
        return (float)((double)nint((t_out - t_now) / tau) * inner);
    }
};

}  // namespace kfsp
