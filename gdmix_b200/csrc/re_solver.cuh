// re_solver.cuh -- the random-effect hot path on sm_100a.
//
// One CTA ("entity group", G = 32..256 threads) owns one entity at a time:
//   1. stage   the entity's CSR slice (fp32 values, int32 local columns, per-sample
//              label/weight/offset) is read from HBM exactly once and laid out in shared
//              memory as a bank-skewed CSR (u16 columns) plus a bank-skewed CSC built on
//              chip by a deterministic counting sort (row-ascending inside each column);
//   2. solve   L-BFGS-B as scipy.optimize.fmin_l_bfgs_b runs it without bounds (MINPACK-2
//              dcsrch line search, skip / restart rules, pgtol + factr + maxiter stop tests)
//              entirely out of shared memory: z = X1.theta by row-threads from the CSR,
//              g = X1^T r by coefficient-threads from the CSC -- no atomics, fixed summation
//              order, fp64 throughout.  The search direction d = -H g uses the compact
//              (Byrd-Nocedal-Schnabel) form of the L-BFGS inverse Hessian with an explicitly
//              maintained R^-1: per iteration ONE batched reduction of the 2m+2 inner
//              products [S;Y]^T g, y.y, y.g, a warp-sized m x m update, and one axpy pass --
//              instead of the two-loop recursion's 2m dependent block reductions.  It is
//              the same direction algebraically (H0 = I/theta);
//   3. emit    theta (optionally thresholded), f, nit, nfev, status, SIMPLE variance.
// CTAs are persistent and pull entities from a global atomic queue, so divergent iteration
// counts between entities never idle an SM.
//
// Reference semantics being replaced (gdmix-trainer/src/gdmix/):
//   models/custom/binary_logistic_regression.py:84-131 (_loss/_gradient), :191-239 (fit),
//   :144-189 (_compute_variance SIMPLE), models/custom/scipy/job_consumers.py:36-63,
//   util/model_utils.py:4-12.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gdmix_b200.h"

namespace gdmix {

constexpr int kMaxWarps = 8;   // G <= 256
constexpr int kRedK = 4;       // values per block reduction
constexpr unsigned kFull = 0xffffffffu;

enum ReMode { kModeFit = 0, kModeLossGrad = 1 };

struct ReArgs {
    gdmix_re_batch b;
    gdmix_lr_opts o;
    const double *theta_in;  // theta0 (fit, nullable) or theta (loss_grad)
    double *theta_out;
    double *f_out;
    int32_t *nit;
    int32_t *nfev;
    int32_t *status;
    double *var_out;
    double *g_out;
    int32_t *queue;             // work counter, zeroed before launch
    unsigned char *arena;       // per-CTA global scratch for history that does not fit on chip
    unsigned long long arena_stride;
    int32_t mode;
    uint32_t smem_bytes;        // dynamic shared memory given to the kernel
};

// Byte layout of one entity's on-chip state.  Host (planning) and device (carving) share it.
struct ReLayout {
    uint32_t xa, xb, ga, gb, dv;      // fp64[p] x / trial x, g / trial g, direction
    uint32_t r;                       // fp64[n] residuals
    uint32_t y, w, off;               // fp32[n]
    uint32_t rowst, colst;            // u32[n+1], u32[d+1] skewed segment starts
    uint32_t csr_val, csc_val;        // fp32[nnz+n], fp32[nnz+d]
    uint32_t csr_col, csc_row;        // u16[nnz+n], u16[nnz+d]
    uint32_t dense;                   // fp64 small matrices / vectors of the compact L-BFGS form
    uint32_t part;                    // fp64[kMaxWarps * (2*MT+2)] per-warp partial inner products
    uint32_t fixed_bytes;             // everything above
    uint32_t hist;                    // fp64[2*m*p]: S rows then Y rows
    uint32_t total_bytes;             // fixed + history
};

// Offsets (in doubles) inside the dense block, MT = compile-time bound on m.
template <int MT>
struct Dense {
    static constexpr int rinv = 0, yy = MT * MT, d = 2 * MT * MT, p1old = d + MT, p2old = p1old + MT,
                         cu = p2old + MT, cw = cu + MT, ta = cw + MT, tb = ta + MT, tot = tb + MT,
                         count = tot + 2 * MT + 2;
};
__host__ __device__ inline uint32_t dense_doubles(uint32_t mt) { return 2 * mt * mt + 9 * mt + 2; }

__host__ __device__ inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

__host__ __device__ inline ReLayout re_layout(uint32_t n, uint32_t nnz, uint32_t d, uint32_t p, uint32_t m,
                                              uint32_t mt)
{
    ReLayout L;
    uint32_t o = 0;
    L.xa = o; o += align16(8 * p);
    L.xb = o; o += align16(8 * p);
    L.ga = o; o += align16(8 * p);
    L.gb = o; o += align16(8 * p);
    L.dv = o; o += align16(8 * p);
    L.r = o; o += align16(8 * n);
    L.y = o; o += align16(4 * n);
    L.w = o; o += align16(4 * n);
    L.off = o; o += align16(4 * n);
    L.rowst = o; o += align16(4 * (n + 1));
    L.colst = o; o += align16(4 * (d + 1));
    L.csr_val = o; o += align16(4 * (nnz + n));
    L.csc_val = o; o += align16(4 * (nnz + d));
    L.csr_col = o; o += align16(2 * (nnz + n));
    L.csc_row = o; o += align16(2 * (nnz + d));
    L.dense = o; o += align16(8 * dense_doubles(mt));
    L.part = o; o += align16(8 * kMaxWarps * (2 * mt + 2));
    L.fixed_bytes = o;
    L.hist = o; o += align16(16 * m * p);
    L.total_bytes = o;
    return L;
}

// ---------------------------------------------------------------------------------------
// block-wide reductions (deterministic: xor butterfly inside a warp, fixed order across warps;
// every thread ends up with the same bits, so all scalar solver logic can run replicated)
// ---------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void group_sync()
{
    if (G == 32) __syncwarp(); else __syncthreads();
}

template <int G, int K>
__device__ __forceinline__ void group_sum(double (&v)[K], double *red, int &flip)
{
#pragma unroll
    for (int k = 0; k < K; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(kFull, v[k], o);
    }
    if (G == 32) return;
    constexpr int W = G / 32;
    double *buf = red + flip * (kMaxWarps * kRedK);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) buf[warp * kRedK + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) {
        double s = buf[k];
#pragma unroll
        for (int w = 1; w < W; w++) s += buf[w * kRedK + k];
        v[k] = s;
    }
    flip ^= 1;
}

template <int G>
__device__ __forceinline__ double group_max(double v, double *red, int &flip)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
    if (G == 32) return v;
    constexpr int W = G / 32;
    double *buf = red + flip * (kMaxWarps * kRedK);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) buf[warp * kRedK] = v;
    __syncthreads();
    double s = buf[0];
#pragma unroll
    for (int w = 1; w < W; w++) s = fmax(s, buf[w * kRedK]);
    flip ^= 1;
    return s;
}

// ---------------------------------------------------------------------------------------
// MINPACK-2 dcsrch / dcstep (More' & Thuente), the line search inside L-BFGS-B.
// Scalar, replicated in every thread of the group.
// ---------------------------------------------------------------------------------------
struct LineSearch {
    double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
    int brackt, stage;
};
enum { LS_START = 0, LS_FG = 1, LS_CONV = 2, LS_WARN = 3, LS_ERROR = 4 };

__device__ __forceinline__ double max3(double a, double b, double c) { return fmax(fmax(a, b), c); }

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

// ---------------------------------------------------------------------------------------
// The entity's staged block and the two passes over it.
// ---------------------------------------------------------------------------------------
struct Staged {
    uint32_t n, d, p, nnz, hi;
    const float *y, *w, *off;
    const uint32_t *rowst, *colst;
    const float *csr_val, *csc_val;
    const uint16_t *csr_col, *csc_row;
    double *r;
    double inv_n, l2;
    int reg_bias;  // intercept is regularised
};

// Pass A (row threads): z = X1.xt + offset, weighted stable cross entropy, residual
// r_i = w_i (sigmoid(z_i) - y_i).  Returns per-thread partials {sum cost, sum r, sum xt_reg^2}.
template <int G>
__device__ __forceinline__ void pass_rows(const Staged &S, const double *xt, double (&part)[3])
{
    const uint32_t tid = threadIdx.x;
    const double b0 = S.hi ? xt[0] : 0.0;
    const double *xf = xt + S.hi;
    double fs = 0.0, rs = 0.0;
    for (uint32_t i = tid; i < S.n; i += G) {
        const uint32_t s = S.rowst[i], len = S.rowst[i + 1] - s - 1;
        double z0 = b0, z1 = 0.0;
        uint32_t j = 0;
        for (; j + 1 < len; j += 2) {
            z0 = fma((double)S.csr_val[s + j], xf[S.csr_col[s + j]], z0);
            z1 = fma((double)S.csr_val[s + j + 1], xf[S.csr_col[s + j + 1]], z1);
        }
        if (j < len) z0 = fma((double)S.csr_val[s + j], xf[S.csr_col[s + j]], z0);
        const double z = (z0 + z1) + (double)S.off[i];
        const double yi = (double)S.y[i], wi = (double)S.w[i];
        const double e = exp(-fabs(z));
        const double ce = fmax(z, 0.0) - z * yi + log(1.0 + e);
        fs = fma(wi, ce, fs);
        const double inv = 1.0 / (1.0 + e);
        const double sig = (z >= 0.0) ? inv : e * inv;
        const double ri = wi * (sig - yi);
        S.r[i] = ri;
        rs += ri;
    }
    double sq = 0.0;
    for (uint32_t jj = tid; jj < S.p; jj += G) {
        if (S.hi && jj == 0 && !S.reg_bias) continue;
        sq = fma(xt[jj], xt[jj], sq);
    }
    part[0] = fs; part[1] = rs; part[2] = sq;
}

// Pass B (coefficient threads, owner j = tid + k*G): g_j = (X1^T r + l2*xt_reg)_j / n.
// Returns per-thread partials of g.dv and max|g|.
template <int G>
__device__ __forceinline__ void pass_cols(const Staged &S, const double *xt, const double *dv, double *gt,
                                          double rsum, double &gd_part, double &gmax_part)
{
    const uint32_t tid = threadIdx.x;
    double gd = 0.0, gm = 0.0;
    for (uint32_t j = tid; j < S.p; j += G) {
        double gj;
        if (S.hi && j == 0) {
            gj = (rsum + (S.reg_bias ? S.l2 * xt[0] : 0.0)) * S.inv_n;
        } else {
            const uint32_t c = j - S.hi;
            const uint32_t s = S.colst[c], len = S.colst[c + 1] - s - 1;
            double a0 = 0.0, a1 = 0.0;
            uint32_t k = 0;
            for (; k + 1 < len; k += 2) {
                a0 = fma((double)S.csc_val[s + k], S.r[S.csc_row[s + k]], a0);
                a1 = fma((double)S.csc_val[s + k + 1], S.r[S.csc_row[s + k + 1]], a1);
            }
            if (k < len) a0 = fma((double)S.csc_val[s + k], S.r[S.csc_row[s + k]], a0);
            gj = ((a0 + a1) + S.l2 * xt[j]) * S.inv_n;
        }
        gt[j] = gj;
        gd = fma(gj, dv[j], gd);
        gm = fmax(gm, fabs(gj));
    }
    gd_part = gd; gmax_part = gm;
}

// f, g at xt.  All threads return identical f, gd (= g.dv) and gmax (= max|g|).
template <int G>
__device__ __forceinline__ void evaluate(const Staged &S, const double *xt, const double *dv, double *gt,
                                         double *red, int &flip, double &f, double &gd, double &gmax)
{
    double part[3];
    pass_rows<G>(S, xt, part);
    group_sum<G, 3>(part, red, flip);  // its barrier also publishes r[]
    if (G == 32) __syncwarp();
    f = (part[0] + 0.5 * S.l2 * part[2]) * S.inv_n;
    double gdp, gmp;
    pass_cols<G>(S, xt, dv, gt, part[1], gdp, gmp);
    double v[1] = {gdp};
    group_sum<G, 1>(v, red, flip);
    gd = v[0];
    gmax = group_max<G>(gmp, red, flip);
}

// ---------------------------------------------------------------------------------------
// Compact L-BFGS direction.  Ring of m physical slots (MT = compile-time bound on m); all small
// matrices are indexed by PHYSICAL slot and hold zeros in rows/columns of empty slots, so the
// m x m products need no masks.  With R_ij = s_i.y_j (i not newer than j), D = diag(s_i.y_i),
// gamma = 1/theta:
//     H g = gamma g + S u - gamma Y w,   w = R^-1 S^T g,   u = R^-T ((D + gamma Y^T Y) w - gamma Y^T g)
// R^-1 is kept explicitly: appending a pair adds the column -R^-1 (S^T y_new) / (s_new.y_new) and the
// diagonal entry 1/(s_new.y_new); dropping the oldest pair deletes its row and column (a trailing
// principal block of an upper-triangular inverse is the inverse of the trailing block).
// S^T y_new and Y^T y_new follow from the inner products with the new and the previous gradient.
// ---------------------------------------------------------------------------------------
struct Lbfgs {
    int col, head;        // pairs stored, physical slot of the oldest
    uint32_t valid;       // bit s set: physical slot s holds a pair
    double theta;
};

template <int G, int MT>
__device__ __forceinline__ void lbfgs_reset(Lbfgs &L, double *dense)
{
    L.col = 0; L.head = 0; L.valid = 0; L.theta = 1.0;
    for (int k = threadIdx.x; k < Dense<MT>::tot; k += G) dense[k] = 0.0;
}

// After an accepted step: g = new gradient, gold = previous gradient, dv = the direction just used.
// Optionally stores the new pair (s = stp*dv, y = g - gold), then writes the next direction into dv and
// returns gd = g.dv and dtd = dv.dv (identical in all threads).
template <int G, int MT>
__device__ __forceinline__ void lbfgs_direction(Lbfgs &L, const int m, const bool update, const double stp,
                                                const double dr, const double gd_new, const uint32_t p,
                                                const double *g, const double *gold, double *dv, double *Sh,
                                                double *Yh, double *dense, double *part, double *red, int &flip,
                                                double &gd, double &dtd)
{
    using DN = Dense<MT>;
    constexpr int W = G / 32;
    constexpr int K = 2 * MT + 2;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    int newslot = -1;
    if (update) newslot = (L.col < m) ? (L.head + L.col) % m : L.head;
    const uint32_t dotmask = update ? (L.valid & ~(1u << newslot)) : L.valid;  // surviving old pairs

    // ---- pass H1: inner products of every stored pair with g, plus y.y and y.g ------------------------
    double a1[MT], a2[MT], yy = 0.0, yg = 0.0;
#pragma unroll
    for (int s = 0; s < MT; s++) { a1[s] = 0.0; a2[s] = 0.0; }
    for (uint32_t j = tid; j < p; j += G) {
        const double gj = g[j], yj = gj - gold[j];
        yy = fma(yj, yj, yy);
        yg = fma(yj, gj, yg);
#pragma unroll
        for (int s = 0; s < MT; s++) {
            if ((dotmask >> s) & 1u) {
                a1[s] = fma(Sh[(size_t)s * p + j], gj, a1[s]);
                a2[s] = fma(Yh[(size_t)s * p + j], gj, a2[s]);
            }
        }
        if (update) {
            Sh[(size_t)newslot * p + j] = stp * dv[j];
            Yh[(size_t)newslot * p + j] = yj;
        }
    }
#pragma unroll
    for (int s = 0; s < MT; s++) {
        if ((dotmask >> s) & 1u) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a1[s] += __shfl_xor_sync(kFull, a1[s], o);
                a2[s] += __shfl_xor_sync(kFull, a2[s], o);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        yy += __shfl_xor_sync(kFull, yy, o);
        yg += __shfl_xor_sync(kFull, yg, o);
    }
    if (lane == 0) {
        double *row = part + warp * K;
#pragma unroll
        for (int s = 0; s < MT; s++) { row[s] = a1[s]; row[MT + s] = a2[s]; }
        row[2 * MT] = yy; row[2 * MT + 1] = yg;
    }
    group_sync<G>();

    // ---- warp 0: totals, pair update, the three m x m products ---------------------------------------
    if (warp == 0) {
        double *tot = dense + DN::tot;
        for (int k = lane; k < K; k += 32) {
            double t = part[k];
#pragma unroll
            for (int w2 = 1; w2 < W; w2++) t += part[w2 * K + k];
            tot[k] = t;
        }
        __syncwarp();
        const int i = lane;
        const bool in = i < MT;
        const bool old_i = in && ((dotmask >> i) & 1u);
        double p1i = old_i ? tot[i] : 0.0, p2i = old_i ? tot[MT + i] : 0.0;
        const double yyt = tot[2 * MT], ygt = tot[2 * MT + 1];
        double *Rinv = dense + DN::rinv, *YY = dense + DN::yy, *Dg = dense + DN::d;
        double *p1old = dense + DN::p1old, *p2old = dense + DN::p2old;
        double *ta = dense + DN::ta, *tb = dense + DN::tb;
        double theta = L.theta;
        if (update) {
            theta = yyt / dr;
            const double rc = old_i ? p1i - p1old[i] : 0.0;  // s_i . y_new
            const double yc = old_i ? p2i - p2old[i] : 0.0;  // y_i . y_new
            if (in) ta[i] = rc;
            __syncwarp();
            double acc = 0.0;
            if (in) {
#pragma unroll
                for (int j = 0; j < MT; j++) acc = fma(Rinv[i * MT + j], ta[j], acc);
            }
            __syncwarp();
            if (in) {
                Rinv[i * MT + newslot] = old_i ? -acc / dr : 0.0;
                YY[i * MT + newslot] = yc;
            }
            __syncwarp();
            if (in) {
                Rinv[newslot * MT + i] = (i == newslot) ? 1.0 / dr : 0.0;
                YY[newslot * MT + i] = (i == newslot) ? yyt : yc;
            }
            if (i == newslot) { Dg[i] = dr; p1i = stp * gd_new; p2i = ygt; }
            __syncwarp();
        }
        const uint32_t valid = update ? (L.valid | (1u << newslot)) : L.valid;
        const bool val_i = in && ((valid >> i) & 1u);
        const double gamma = 1.0 / theta;
        if (in) { p1old[i] = p1i; p2old[i] = p2i; ta[i] = val_i ? p1i : 0.0; }
        __syncwarp();
        double wv = 0.0;
        if (in) {
#pragma unroll
            for (int j = 0; j < MT; j++) wv = fma(Rinv[i * MT + j], ta[j], wv);
        }
        if (in) tb[i] = wv;
        __syncwarp();
        double yw = 0.0;
        if (in) {
#pragma unroll
            for (int j = 0; j < MT; j++) yw = fma(YY[i * MT + j], tb[j], yw);
        }
        const double tv = val_i ? fma(Dg[i], wv, gamma * (yw - p2i)) : 0.0;
        __syncwarp();
        if (in) ta[i] = tv;
        __syncwarp();
        double uv = 0.0;
        if (in) {
#pragma unroll
            for (int j = 0; j < MT; j++) uv = fma(Rinv[j * MT + i], ta[j], uv);
        }
        if (in) { dense[DN::cu + i] = uv; dense[DN::cw + i] = wv; }
    }
    group_sync<G>();
    if (update) {
        L.theta = dense[DN::tot + 2 * MT] / dr;
        L.valid |= (1u << newslot);
        if (L.col < m) L.col++; else L.head = (L.head + 1) % m;
    }

    // ---- pass H2: dv = -gamma g - S u + gamma Y w ------------------------------------------------------
    const double gamma = 1.0 / L.theta;
    double cu[MT], cw[MT];
#pragma unroll
    for (int s = 0; s < MT; s++) { cu[s] = dense[DN::cu + s]; cw[s] = gamma * dense[DN::cw + s]; }
    double v2[2] = {0.0, 0.0};
    const uint32_t valid = L.valid;
    for (uint32_t j = tid; j < p; j += G) {
        const double gj = g[j];
        double acc = -gamma * gj;
#pragma unroll
        for (int s = 0; s < MT; s++) {
            if ((valid >> s) & 1u) {
                acc = fma(-cu[s], Sh[(size_t)s * p + j], acc);
                acc = fma(cw[s], Yh[(size_t)s * p + j], acc);
            }
        }
        dv[j] = acc;
        v2[0] = fma(gj, acc, v2[0]);
        v2[1] = fma(acc, acc, v2[1]);
    }
    group_sum<G, 2>(v2, red, flip);
    gd = v2[0];
    dtd = v2[1];
}

// ---------------------------------------------------------------------------------------
// The kernel.
// ---------------------------------------------------------------------------------------
template <int G, int MT>
__global__ void __launch_bounds__(G) re_solver_kernel(const ReArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double red[2 * kMaxWarps * kRedK];
    __shared__ int s_entity;
    __shared__ unsigned s_bad;

    constexpr int W = G / 32;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t hi = a.o.has_intercept ? 1u : 0u;
    const int m = a.o.m;
    int flip = 0;

    for (;;) {
        group_sync<G>();  // previous entity fully emitted before its memory is reused
        if (tid == 0) { s_entity = atomicAdd(a.queue, 1); s_bad = 0; }
        group_sync<G>();
        const int64_t e = s_entity;
        if (e >= a.b.n_entities) break;

        const int64_t r0 = a.b.ent_rowptr[e], r1 = a.b.ent_rowptr[e + 1];
        const int64_t q0 = a.b.rowptr[r0], q1 = a.b.rowptr[r1];
        const int64_t t0 = a.b.theta_ptr[e];
        const int64_t n64 = r1 - r0, nnz64 = q1 - q0, p64 = a.b.theta_ptr[e + 1] - t0;
        const uint32_t n = (uint32_t)n64, nnz = (uint32_t)nnz64, p = (uint32_t)p64, d = p - hi;

        const ReLayout L = re_layout(n, nnz, d, p, (uint32_t)m, (uint32_t)MT);
        const bool ok = n64 >= 1 && n64 < 65535 && p64 >= 1 && p64 >= (int64_t)hi && (p64 - hi) < 65535 &&
                        nnz64 >= 0 && nnz64 < (1ll << 30) && L.fixed_bytes <= a.smem_bytes;
        if (!ok) {
            if (tid == 0 && a.status) a.status[e] = GDMIX_ERR_TOO_LARGE;
            continue;
        }
        double *xa = (double *)(smem + L.xa), *xb = (double *)(smem + L.xb);
        double *ga = (double *)(smem + L.ga), *gb = (double *)(smem + L.gb);
        double *dv = (double *)(smem + L.dv);
        double *rres = (double *)(smem + L.r);
        float *sy = (float *)(smem + L.y), *sw = (float *)(smem + L.w), *soff = (float *)(smem + L.off);
        uint32_t *rowst = (uint32_t *)(smem + L.rowst), *colst = (uint32_t *)(smem + L.colst);
        float *csr_val = (float *)(smem + L.csr_val), *csc_val = (float *)(smem + L.csc_val);
        uint16_t *csr_col = (uint16_t *)(smem + L.csr_col), *csc_row = (uint16_t *)(smem + L.csc_row);
        double *dense = (double *)(smem + L.dense), *part = (double *)(smem + L.part);
        double *hist = (L.total_bytes <= a.smem_bytes)
                           ? (double *)(smem + L.hist)
                           : (double *)(a.arena + (unsigned long long)blockIdx.x * a.arena_stride);
        double *Sh = hist, *Yh = hist + (size_t)m * p;

        // ---- stage: per-sample scalars, skewed CSR (thread per row) ----------------------
        for (uint32_t i = tid; i < n; i += G) {
            const int64_t gi = r0 + i;
            sy[i] = a.b.label[gi];
            sw[i] = a.b.weight ? a.b.weight[gi] : 1.0f;
            soff[i] = a.b.offset ? a.b.offset[gi] : 0.0f;
            const int64_t gs = a.b.rowptr[gi], ge = a.b.rowptr[gi + 1];
            uint32_t dst = (uint32_t)(gs - q0) + i;  // +i: one pad word per row skews the banks
            rowst[i] = dst;
            unsigned bad = 0;
            for (int64_t q = gs; q < ge; q++, dst++) {
                const int32_t c = a.b.col[q];
                bad |= ((uint32_t)c >= d);
                csr_val[dst] = a.b.val[q];
                csr_col[dst] = (uint16_t)c;
            }
            csr_val[dst] = 0.0f;
            csr_col[dst] = 0;
            if (bad) atomicOr(&s_bad, 1u);
        }
        if (tid == 0) rowst[n] = nnz + n;

        // ---- build the CSC on chip: per-warp column counts -> scan -> ordered fill ---------
        uint32_t *cntw = (uint32_t *)xa;  // W*d counters alias the (not yet used) solver vectors
        for (uint32_t k = tid; k < W * d; k += G) cntw[k] = 0;
        group_sync<G>();
        if (s_bad) {
            if (tid == 0 && a.status) a.status[e] = GDMIX_ERR_INVALID;
            continue;
        }
        const uint32_t chunk = (n + W - 1) / W;
        const uint32_t rbeg = min(n, warp * chunk), rend = min(n, rbeg + chunk);
        for (uint32_t i = rbeg; i < rend; i++) {
            const uint32_t s = rowst[i], len = rowst[i + 1] - s - 1;
            for (uint32_t j = lane; j < len; j += 32) atomicAdd(&cntw[warp * d + csr_col[s + j]], 1u);
        }
        group_sync<G>();
        {
            // exclusive scan over columns; thread t owns the contiguous columns [t*ipt, (t+1)*ipt)
            const uint32_t ipt = (d + G - 1) / G;
            const uint32_t cbeg = min(d, tid * ipt), cend = min(d, cbeg + ipt);
            uint32_t mine = 0;
            for (uint32_t c = cbeg; c < cend; c++)
                for (int w2 = 0; w2 < W; w2++) mine += cntw[w2 * d + c];
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(kFull, incl, o);
                if (lane >= (uint32_t)o) incl += t;
            }
            uint32_t *wtot = (uint32_t *)red;  // W partial totals (red is free here)
            if (G > 32) {
                if (lane == 31) wtot[warp] = incl;
                __syncthreads();
            }
            uint32_t base = incl - mine;
            if (G > 32)
                for (uint32_t w2 = 0; w2 < warp; w2++) base += wtot[w2];
            for (uint32_t c = cbeg; c < cend; c++) {
                uint32_t pos = base + c;  // +c: one pad word per column
                colst[c] = pos;
                for (int w2 = 0; w2 < W; w2++) {
                    const uint32_t cnt = cntw[w2 * d + c];
                    cntw[w2 * d + c] = pos;  // becomes warp w2's write cursor for column c
                    pos += cnt;
                }
                base = pos - c;
            }
            if (tid == 0) colst[d] = nnz + d;
        }
        group_sync<G>();
        for (uint32_t i = rbeg; i < rend; i++) {
            const uint32_t s = rowst[i], len = rowst[i + 1] - s - 1;
            for (uint32_t j0 = 0; j0 < len; j0 += 32) {
                const uint32_t j = j0 + lane;
                const bool act = j < len;
                const uint32_t c = act ? (uint32_t)csr_col[s + j] : (0x10000u + lane);
                const float v = act ? csr_val[s + j] : 0.0f;
                const unsigned grp = __match_any_sync(kFull, c);  // duplicates inside one row
                const uint32_t rank = __popc(grp & ((1u << lane) - 1u));
                uint32_t cur = 0;
                if (act) {
                    cur = cntw[warp * d + c];
                    csc_row[cur + rank] = (uint16_t)i;
                    csc_val[cur + rank] = v;
                }
                __syncwarp();
                if (act && rank == (uint32_t)__popc(grp) - 1u) cntw[warp * d + c] = cur + rank + 1u;
                __syncwarp();
            }
        }
        group_sync<G>();

        Staged S;
        S.n = n; S.d = d; S.p = p; S.nnz = nnz; S.hi = hi;
        S.y = sy; S.w = sw; S.off = soff; S.rowst = rowst; S.colst = colst;
        S.csr_val = csr_val; S.csc_val = csc_val; S.csr_col = csr_col; S.csc_row = csc_row;
        S.r = rres; S.inv_n = 1.0 / (double)n; S.l2 = a.o.l2; S.reg_bias = a.o.regularize_bias;

        double *x = xa, *xt = xb, *g = ga, *gt = gb;
        for (uint32_t j = tid; j < p; j += G) {
            x[j] = a.theta_in ? a.theta_in[t0 + j] : 0.0;
            dv[j] = 0.0;
        }
        Lbfgs lb;
        lbfgs_reset<G, MT>(lb, dense);
        group_sync<G>();

        double f, gd, gmax;
        evaluate<G>(S, x, dv, g, red, flip, f, gd, gmax);
        int nfev = 1, iter = 0, status = GDMIX_SOLVE_CONVERGED;

        if (a.mode == kModeLossGrad) {
            for (uint32_t j = tid; j < p; j += G) a.g_out[t0 + j] = g[j];
            if (tid == 0) a.f_out[e] = f;
            continue;
        }

        // ---- L-BFGS-B, unbounded ---------------------------------------------------------
        const double epsmch = 2.220446049250313e-16;
        const double ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmx = 1e10;
        double dtd;
        bool done = gmax <= a.o.pgtol;
        if (!done) {
            // steepest descent start: dv = -g
            double v1[1] = {0.0};
            for (uint32_t j = tid; j < p; j += G) {
                const double gj = g[j];
                dv[j] = -gj;
                v1[0] = fma(gj, gj, v1[0]);
            }
            group_sum<G, 1>(v1, red, flip);
            dtd = v1[0];
            gd = -v1[0];
        }

        while (!done) {
            // line search (lnsrlb + dcsrch) along dv; gd = g.dv, dtd = dv.dv on entry
            double stp = (iter == 0) ? fmin(1.0 / sqrt(dtd), stpmx) : 1.0;
            const double fold = f, gdold = gd;
            int ifun = 0, iback = 0, info = 0, task = LS_START;
            LineSearch ls;
            double gmax_t = gmax;
            if (gd >= 0.0) info = -4;
            while (info == 0) {
                task = dcsrch(stp, f, gd, ftol, gtol, xtol, 0.0, stpmx, task, ls);
                if (task == LS_CONV || task == LS_WARN) break;
                if (task == LS_ERROR) { info = -4; break; }
                ifun++; iback = ifun - 1;
                if (iback >= a.o.max_ls) break;
                for (uint32_t j = tid; j < p; j += G) xt[j] = fma(stp, dv[j], x[j]);
                group_sync<G>();
                evaluate<G>(S, xt, dv, gt, red, flip, f, gd, gmax_t);
                nfev++;
            }
            if (info != 0 || iback >= a.o.max_ls) {
                f = fold;  // x, g still hold the previous iterate
                if (lb.col == 0) { status = GDMIX_SOLVE_ABNORMAL; iter++; break; }
                // refresh the memory and restart from steepest descent
                group_sync<G>();
                lbfgs_reset<G, MT>(lb, dense);
                double v2[1] = {0.0};
                for (uint32_t j = tid; j < p; j += G) {
                    const double gj = g[j];
                    dv[j] = -gj;
                    v2[0] = fma(gj, gj, v2[0]);
                }
                group_sum<G, 1>(v2, red, flip);
                dtd = v2[0];
                gd = -v2[0];
                continue;
            }
            iter++;
            // accept: (x, g) <-> (xt, gt); gt now holds the previous gradient
            { double *t = x; x = xt; xt = t; t = g; g = gt; gt = t; }
            gmax = gmax_t;

            if (iter >= a.o.max_iter || nfev > a.o.max_fun) { status = GDMIX_SOLVE_MAXITER; break; }
            if (gmax <= a.o.pgtol) break;
            if ((fold - f) <= epsmch * a.o.factr * max3(fabs(fold), fabs(f), 1.0)) break;

            // curvature pair (L-BFGS-B's skip rule) and the next direction
            double dr, ddum;
            if (stp == 1.0) { dr = gd - gdold; ddum = -gdold; }
            else { dr = (gd - gdold) * stp; ddum = -gdold * stp; }
            const bool update = (m > 0) && !(dr <= epsmch * ddum);
            lbfgs_direction<G, MT>(lb, m, update, stp, dr, gd, p, g, gt, dv, Sh, Yh, dense, part, red, flip, gd,
                                   dtd);
        }

        // ---- emit ------------------------------------------------------------------------
        const double thr = a.o.sparsity_threshold;
        for (uint32_t j = tid; j < p; j += G) {
            const double xj = x[j];
            a.theta_out[t0 + j] = (thr > 0.0 && fabs(xj) <= thr) ? 0.0 : xj;
        }
        if (tid == 0) {
            if (a.f_out) a.f_out[e] = f;
            if (a.nit) a.nit[e] = iter;
            if (a.nfev) a.nfev[e] = nfev;
            if (a.status) a.status[e] = status;
        }
        if (a.var_out && a.o.variance_mode == GDMIX_VARIANCE_SIMPLE) {
            // var_j = 1 / (sum_i x_ij^2 rho_i (1-rho_i) w_i + l2 [j regularised] + 1e-12)
            const double b0 = hi ? x[0] : 0.0;
            for (uint32_t i = tid; i < n; i += G) {
                const uint32_t s = rowst[i], len = rowst[i + 1] - s - 1;
                double z = b0;
                for (uint32_t j = 0; j < len; j++) z = fma((double)csr_val[s + j], x[hi + csr_col[s + j]], z);
                z += (double)soff[i];
                const double rho = 1.0 / (1.0 + exp(-z));
                rres[i] = rho * (1.0 - rho) * (double)sw[i];
            }
            double dsum[1] = {0.0};
            for (uint32_t i = tid; i < n; i += G) dsum[0] += rres[i];
            group_sum<G, 1>(dsum, red, flip);
            if (G == 32) __syncwarp();
            for (uint32_t j = tid; j < p; j += G) {
                double h;
                if (hi && j == 0) {
                    h = dsum[0] + (a.o.regularize_bias ? a.o.l2 : 0.0);
                } else {
                    const uint32_t c = j - hi, s = colst[c], len = colst[c + 1] - s - 1;
                    h = 0.0;
                    for (uint32_t k = 0; k < len; k++) {
                        const double v = (double)csc_val[s + k];
                        h = fma(v, v * rres[csc_row[s + k]], h);
                    }
                    h += a.o.l2;
                }
                a.var_out[t0 + j] = 1.0 / (h + 1.0e-12);
            }
        }
    }
}

}  // namespace gdmix
