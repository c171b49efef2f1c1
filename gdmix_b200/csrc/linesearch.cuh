// linesearch.cuh -- MINPACK-2 dcsrch / dcstep (More' & Thuente), the line search inside L-BFGS-B.
#pragma once
#include "re_common.cuh"
namespace gdmix {
// ---------------------------------------------------------------------------------------
// MINPACK-2 dcsrch / dcstep (More' & Thuente), the line search inside L-BFGS-B.
// Scalar, replicated in every thread of the group.
// ---------------------------------------------------------------------------------------
struct LineSearch {
    double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
    int brackt, stage;
};
enum { LS_START = 0, LS_FG = 1, LS_CONV = 2, LS_WARN = 3, LS_ERROR = 4 };

// max of two doubles that are >= 0 (callers pass |x| or constants): their bit patterns order like the numbers, and a
// 64-bit integer max is 4 instructions where fmax(double, double) is ~10 (its NaN rules); |NaN| is the largest pattern
// and propagates, as np.max(np.abs(g)) does in the reference's stop test.
__device__ __forceinline__ double pos_max(const double a, const double b)
{
    const long long ia = __double_as_longlong(a), ib = __double_as_longlong(b);
    return __longlong_as_double(ia > ib ? ia : ib);
}
__device__ __forceinline__ double max3(double a, double b, double c) { return pos_max(pos_max(a, b), c); }   // a, b, c >= 0

__device__ inline void dcstep(double &stx, double &fx, double &dx, double &sty, double &fy, double &dy, double &stp,
                              double fp, double dp, int &brackt, double stpmin, double stpmax)
{
    const double sgnd = dp * (dx / fabs(dx));
    double theta, s, gamma, p, q, r, stpc, stpq, stpf;
    if (fp > fx) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = max3(fabs(theta), fabs(dx), fabs(dp));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp < stx) gamma = -gamma;
        p = (gamma - dx) + theta;
        q = ((gamma - dx) + gamma) + dp;
        r = p / q;
        stpc = stx + r * (stp - stx);
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
        stpf = (fabs(stpc - stx) < fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
        brackt = 1;
    } else if (sgnd < 0.0) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = max3(fabs(theta), fabs(dx), fabs(dp));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = ((gamma - dp) + gamma) + dx;
        r = p / q;
        stpc = stp + r * (stx - stp);
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
        brackt = 1;
    } else if (fabs(dp) < fabs(dx)) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = max3(fabs(theta), fabs(dx), fabs(dp));
        gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = (gamma + (dx - dp)) + gamma;
        r = p / q;
        if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
        else if (stp > stx) stpc = stpmax;
        else stpc = stpmin;
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (brackt) {
            stpf = (fabs(stpc - stp) < fabs(stpq - stp)) ? stpc : stpq;
            if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
            else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
        } else {
            stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
            stpf = fmin(stpmax, stpf);
            stpf = fmax(stpmin, stpf);
        }
    } else {
        if (brackt) {
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
            s = max3(fabs(theta), fabs(dy), fabs(dp));
            gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
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

__device__ inline int dcsrch(double &stp, double f, double g, double ftol, double gtol, double xtol, double stpmin,
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
    if (f <= ftest && fabs(g) <= gtol * (-S.ginit)) out = LS_CONV;
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
        if (fabs(S.sty - S.stx) >= p66 * S.width1) stp = S.stx + p5 * (S.sty - S.stx);
        S.width1 = S.width;
        S.width = fabs(S.sty - S.stx);
        S.stmin = fmin(S.stx, S.sty);
        S.stmax = fmax(S.stx, S.sty);
    } else {
        S.stmin = stp + xtrapl * (stp - S.stx);
        S.stmax = stp + xtrapu * (stp - S.stx);
    }
    stp = fmax(stp, stpmin);
    stp = fmin(stp, stpmax);
    if ((S.brackt && (stp <= S.stmin || stp >= S.stmax)) || (S.brackt && S.stmax - S.stmin <= xtol * S.stmax))
        stp = S.stx;
    return LS_FG;
}

}  // namespace gdmix
