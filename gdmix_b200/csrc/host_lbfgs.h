// host_lbfgs.h -- host-side, reverse-communication L-BFGS-B (unbounded) for the FIXED-effect solve.
//
// The reference runs scipy.optimize.fmin_l_bfgs_b on every worker with func = "stream my shard, all-reduce
// value and gradient" (gdmix-trainer/src/gdmix/models/custom/fixed_effect_lr_lbfgs_model.py:635-643,
// :394-404).  Here the objective/gradient is the CUDA kernel gdmix_fe_loss_grad followed by an NCCL
// all-reduce; this class is the replicated solver state that consumes the reduced (f, g).  It is the same
// algorithm the device kernel runs per entity (L-BFGS-B 3.0 driver logic + MINPACK-2 dcsrch), written as a
// state machine so the caller owns the evaluation:
//
//     task = iterate(x, f, g)      // f, g evaluated at the x the previous call returned
//     task == kNeedFG  -> evaluate at x (updated in place) and call again
//     task == kDone    -> x holds the solution; status()/nit()/nfev() as scipy's warnflag/nit/funcalls
//
// Every rank feeds bit-identical reduced (f, g), so every rank's state stays bit-identical, exactly like the
// reference's replicated scipy instances.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <thread>
#include <vector>

namespace gdmix_host {

struct LineSearch {
    double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
    int brackt, stage;
};
enum { LS_START = 0, LS_FG = 1, LS_CONV = 2, LS_WARN = 3, LS_ERROR = 4 };

inline double max3(double a, double b, double c) { return std::fmax(std::fmax(a, b), c); }

inline void dcstep(double &stx, double &fx, double &dx, double &sty, double &fy, double &dy, double &stp, double fp,
                   double dp, int &brackt, double stpmin, double stpmax)
{
    const double sgnd = dp * (dx / std::fabs(dx));
    double theta, s, gamma, p, q, r, stpc, stpq, stpf;
    if (fp > fx) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = max3(std::fabs(theta), std::fabs(dx), std::fabs(dp));
        gamma = s * std::sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp < stx) gamma = -gamma;
        p = (gamma - dx) + theta;
        q = ((gamma - dx) + gamma) + dp;
        r = p / q;
        stpc = stx + r * (stp - stx);
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
        stpf = (std::fabs(stpc - stx) < std::fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
        brackt = 1;
    } else if (sgnd < 0.0) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = max3(std::fabs(theta), std::fabs(dx), std::fabs(dp));
        gamma = s * std::sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = ((gamma - dp) + gamma) + dx;
        r = p / q;
        stpc = stp + r * (stx - stp);
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        stpf = (std::fabs(stpc - stp) > std::fabs(stpq - stp)) ? stpc : stpq;
        brackt = 1;
    } else if (std::fabs(dp) < std::fabs(dx)) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = max3(std::fabs(theta), std::fabs(dx), std::fabs(dp));
        gamma = s * std::sqrt(std::fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = (gamma + (dx - dp)) + gamma;
        r = p / q;
        if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
        else if (stp > stx) stpc = stpmax;
        else stpc = stpmin;
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (brackt) {
            stpf = (std::fabs(stpc - stp) < std::fabs(stpq - stp)) ? stpc : stpq;
            if (stp > stx) stpf = std::fmin(stp + 0.66 * (sty - stp), stpf);
            else stpf = std::fmax(stp + 0.66 * (sty - stp), stpf);
        } else {
            stpf = (std::fabs(stpc - stp) > std::fabs(stpq - stp)) ? stpc : stpq;
            stpf = std::fmin(stpmax, stpf);
            stpf = std::fmax(stpmin, stpf);
        }
    } else {
        if (brackt) {
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
            s = max3(std::fabs(theta), std::fabs(dy), std::fabs(dp));
            gamma = s * std::sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
            if (stp > sty) gamma = -gamma;
            p = (gamma - dp) + theta;
            q = ((gamma - dp) + gamma) + dy;
            r = p / q;
            stpf = stp + r * (sty - stp);
        } else if (stp > stx) stpf = stpmax;
        else stpf = stpmin;
    }
    if (fp > fx) {
        sty = stp; fy = fp; dy = dp;
    } else {
        if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
        stx = stp; fx = fp; dx = dp;
    }
    stp = stpf;
}

inline int dcsrch(double &stp, double f, double g, double ftol, double gtol, double xtol, double stpmin,
                  double stpmax, int task, LineSearch &S)
{
    const double p5 = 0.5, p66 = 0.66, xtrapl = 1.1, xtrapu = 4.0;
    if (task == LS_START) {
        if (stp < stpmin || stp > stpmax || g >= 0.0 || stpmax < stpmin) return LS_ERROR;
        S.brackt = 0; S.stage = 1;
        S.finit = f; S.ginit = g; S.gtest = ftol * g;
        S.width = stpmax - stpmin; S.width1 = S.width / p5;
        S.stx = 0.0; S.fx = f; S.gx = g;
        S.sty = 0.0; S.fy = f; S.gy = g;
        S.stmin = 0.0; S.stmax = stp + xtrapu * stp;
        return LS_FG;
    }
    const double ftest = S.finit + stp * S.gtest;
    if (S.stage == 1 && f <= ftest && g >= 0.0) S.stage = 2;
    int out = LS_FG;
    if (S.brackt && (stp <= S.stmin || stp >= S.stmax)) out = LS_WARN;
    if (S.brackt && S.stmax - S.stmin <= xtol * S.stmax) out = LS_WARN;
    if (stp == stpmax && f <= ftest && g <= S.gtest) out = LS_WARN;
    if (stp == stpmin && (f > ftest || g >= S.gtest)) out = LS_WARN;
    if (f <= ftest && std::fabs(g) <= gtol * (-S.ginit)) out = LS_CONV;
    if (out != LS_FG) return out;
    if (S.stage == 1 && f <= S.fx && f > ftest) {
        double fm = f - stp * S.gtest, fxm = S.fx - S.stx * S.gtest, fym = S.fy - S.sty * S.gtest;
        double gm = g - S.gtest, gxm = S.gx - S.gtest, gym = S.gy - S.gtest;
        dcstep(S.stx, fxm, gxm, S.sty, fym, gym, stp, fm, gm, S.brackt, S.stmin, S.stmax);
        S.fx = fxm + S.stx * S.gtest;
        S.fy = fym + S.sty * S.gtest;
        S.gx = gxm + S.gtest;
        S.gy = gym + S.gtest;
    } else {
        dcstep(S.stx, S.fx, S.gx, S.sty, S.fy, S.gy, stp, f, g, S.brackt, S.stmin, S.stmax);
    }
    if (S.brackt) {
        if (std::fabs(S.sty - S.stx) >= p66 * S.width1) stp = S.stx + p5 * (S.sty - S.stx);
        S.width1 = S.width;
        S.width = std::fabs(S.sty - S.stx);
        S.stmin = std::fmin(S.stx, S.sty);
        S.stmax = std::fmax(S.stx, S.sty);
    } else {
        S.stmin = stp + xtrapl * (stp - S.stx);
        S.stmax = stp + xtrapu * (stp - S.stx);
    }
    stp = std::fmax(stp, stpmin);
    stp = std::fmin(stp, stpmax);
    if ((S.brackt && (stp <= S.stmin || stp >= S.stmax)) || (S.brackt && S.stmax - S.stmin <= xtol * S.stmax))
        stp = S.stx;
    return LS_FG;
}

class Lbfgs {
public:
    enum { kDone = 0, kNeedFG = 1 };

    Lbfgs(int64_t n, int m, int max_iter, int max_ls, int max_fun, double factr, double pgtol)
        : n_(n), m_(m), max_iter_(max_iter), max_ls_(max_ls), max_fun_(max_fun), factr_(factr), pgtol_(pgtol),
          g_(n), d_(n), t_(n), r_(n), q_(n), S_((size_t)n * std::max(m, 1)), Y_((size_t)n * std::max(m, 1)),
          rho_(std::max(m, 1)), alpha_(std::max(m, 1)) {}

    int iterate(double *x, double f, const double *g)
    {
        if (state_ == 2) return kDone;
        f_ = f;
        std::copy(g, g + n_, g_.begin());
        if (state_ == 0) {
            nfev_ = 1;
            if (gnorm() <= pgtol_) return finish(0);
            state_ = 1;
            begin_iteration(x);
            return advance(x);
        }
        nfev_++;
        return advance(x);
    }
    int nit() const { return iter_; }
    int nfev() const { return nfev_; }
    int status() const { return status_; }
    double f() const { return f_; }
    const double *grad() const { return g_.data(); }

private:
    // Long vectors (a 100 000-feature fixed effect streams 32 MB of history through the two-loop recursion per
    // iteration) are swept by several host threads.  Sums are formed per fixed block of kBlock entries and the block
    // sums added in block order, so the result does not depend on the number of threads -- every rank, whatever
    // its core count, keeps bit-identical solver state.  Vectors of at most kBlock entries are summed exactly as
    // before (one block).
    static constexpr int64_t kBlock = 4096, kParMin = 32768;
    static int host_threads()
    {
        static const int k = [] {
            const char *lw = getenv("LOCAL_WORLD_SIZE");
            const char *cap = getenv("GDMIX_HOST_THREADS");
            const int ranks = lw ? std::max(1, atoi(lw)) : 1;
            const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
            const int want = cap ? atoi(cap) : std::min(16, hw / ranks);
            return std::max(1, want);
        }();
        return k;
    }
    double dot(const double *a, const double *b) const
    {
        if (n_ <= kBlock) {
            double s = 0.0;
            for (int64_t i = 0; i < n_; i++) s += a[i] * b[i];
            return s;
        }
        const int64_t nb = (n_ + kBlock - 1) / kBlock;
        part_.resize((size_t)nb);
        double *part = part_.data();
        const int64_t n = n_;
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (n >= kParMin)
        for (int64_t blk = 0; blk < nb; blk++) {
            const int64_t i0 = blk * kBlock, i1 = std::min(n, i0 + kBlock);
            double s = 0.0;
            for (int64_t i = i0; i < i1; i++) s += a[i] * b[i];
            part[blk] = s;
        }
        double s = 0.0;
        for (int64_t blk = 0; blk < nb; blk++) s += part[blk];
        return s;
    }
    double gnorm() const
    {
        double s = 0.0;
        for (int64_t i = 0; i < n_; i++) s = std::fmax(s, std::fabs(g_[i]));
        return s;
    }
    // y[i] = y[i] + c * x[i]  (element-wise: any partition over threads gives the same bits)
    void axpy(double *y, const double c, const double *x) const
    {
        const int64_t n = n_;
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (n >= kParMin)
        for (int64_t i = 0; i < n; i++) y[i] += c * x[i];
    }
    int finish(int status) { status_ = status; state_ = 2; return kDone; }

    void begin_iteration(const double *x)
    {
        // two-loop recursion, H0 = I / theta
        std::copy(g_.begin(), g_.end(), q_.begin());
        for (int k = col_ - 1; k >= 0; k--) {
            const int s = (head_ + k) % m_;
            alpha_[s] = rho_[s] * dot(&S_[(size_t)s * n_], q_.data());
            axpy(q_.data(), -alpha_[s], &Y_[(size_t)s * n_]);
        }
        for (int64_t i = 0; i < n_; i++) q_[i] = q_[i] / theta_;
        for (int k = 0; k < col_; k++) {
            const int s = (head_ + k) % m_;
            const double beta = rho_[s] * dot(&Y_[(size_t)s * n_], q_.data());
            axpy(q_.data(), alpha_[s] - beta, &S_[(size_t)s * n_]);
        }
        for (int64_t i = 0; i < n_; i++) d_[i] = -q_[i];
        const double dnorm = std::sqrt(dot(d_.data(), d_.data()));
        stp_ = (iter_ == 0) ? std::fmin(1.0 / dnorm, 1e10) : 1.0;
        std::copy(x, x + n_, t_.begin());
        std::copy(g_.begin(), g_.end(), r_.begin());
        fold_ = f_;
        ifun_ = 0; iback_ = 0; lstask_ = LS_START;
    }

    int advance(double *x)
    {
        for (;;) {
            int info = 0;
            gd_ = dot(g_.data(), d_.data());
            if (ifun_ == 0) {
                gdold_ = gd_;
                if (gd_ >= 0.0) info = -4;
            }
            if (info == 0) {
                lstask_ = dcsrch(stp_, f_, gd_, 1e-3, 0.9, 0.1, 0.0, 1e10, lstask_, ls_);
                if (lstask_ == LS_CONV || lstask_ == LS_WARN) {
                    const int rc = end_iteration(x);
                    if (rc >= 0) return rc;
                    continue;  // next iteration's line search starts
                }
                if (lstask_ == LS_ERROR) info = -4;
            }
            if (info == 0) {
                ifun_++; iback_ = ifun_ - 1;
                if (iback_ < max_ls_) {
                    for (int64_t i = 0; i < n_; i++) x[i] = stp_ * d_[i] + t_[i];
                    return kNeedFG;
                }
            }
            // line search failed: restore the previous iterate
            std::copy(t_.begin(), t_.end(), x);
            std::copy(r_.begin(), r_.end(), g_.begin());
            f_ = fold_;
            if (col_ == 0) { iter_++; return finish(2); }
            col_ = 0; head_ = 0; theta_ = 1.0;
            begin_iteration(x);
        }
    }

    // returns kDone / -1 (continue with the next iteration)
    int end_iteration(double *x)
    {
        const double epsmch = 2.220446049250313e-16;
        iter_++;
        if (iter_ >= max_iter_ || nfev_ > max_fun_) return finish(1);
        if (gnorm() <= pgtol_) return finish(0);
        if ((fold_ - f_) <= epsmch * factr_ * max3(std::fabs(fold_), std::fabs(f_), 1.0)) return finish(0);
        double rr = 0.0, dr, ddum;
        for (int64_t i = 0; i < n_; i++) { r_[i] = g_[i] - r_[i]; rr += r_[i] * r_[i]; }
        if (stp_ == 1.0) { dr = gd_ - gdold_; ddum = -gdold_; }
        else {
            dr = (gd_ - gdold_) * stp_; ddum = -gdold_ * stp_;
            for (int64_t i = 0; i < n_; i++) d_[i] *= stp_;
        }
        if (!(dr <= epsmch * ddum) && m_ > 0) {
            int slot;
            if (col_ < m_) { slot = (head_ + col_) % m_; col_++; }
            else { slot = head_; head_ = (head_ + 1) % m_; }
            std::copy(d_.begin(), d_.end(), S_.begin() + (size_t)slot * n_);
            std::copy(r_.begin(), r_.end(), Y_.begin() + (size_t)slot * n_);
            rho_[slot] = 1.0 / dr;
            theta_ = rr / dr;
        }
        begin_iteration(x);
        return -1;
    }

    int64_t n_;
    int m_, max_iter_, max_ls_, max_fun_;
    double factr_, pgtol_;
    std::vector<double> g_, d_, t_, r_, q_, S_, Y_, rho_, alpha_;
    mutable std::vector<double> part_;   // block sums of dot()
    int col_ = 0, head_ = 0, iter_ = 0, nfev_ = 0, state_ = 0, status_ = 0;
    double theta_ = 1.0, f_ = 0.0, stp_ = 0.0, fold_ = 0.0, gd_ = 0.0, gdold_ = 0.0;
    int ifun_ = 0, iback_ = 0, lstask_ = LS_START;
    LineSearch ls_;
};

}  // namespace gdmix_host
