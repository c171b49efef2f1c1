// re_solver.cuh -- the random-effect hot path on sm_100a.
//
// One CTA ("entity group", G = 32..256 threads) owns one entity at a time:
//   1. stage   the entity's CSR slice (fp32 values, int32 local columns, per-sample
//              label/weight/offset) is read from HBM exactly once and laid out in shared
//              memory as a bank-skewed CSR (u16 columns) plus a bank-skewed CSC built on
//              chip by a deterministic counting sort (row-ascending inside each column);
//   2. solve   L-BFGS-B as scipy.optimize.fmin_l_bfgs_b runs it without bounds
//              (two-loop direction with H0 = I/theta, MINPACK-2 dcsrch line search,
//              skip / restart rules, pgtol + factr + maxiter stop tests) entirely out of
//              shared memory: z = X1.theta by row-threads from the CSR, g = X1^T r by
//              coefficient-threads from the CSC -- no atomics, fixed summation order,
//              fp64 throughout;
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
    uint32_t fixed_bytes;             // everything above
    uint32_t hist;                    // fp64[2*m*p]: S rows then Y rows
    uint32_t total_bytes;             // fixed + history
};

__host__ __device__ inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

__host__ __device__ inline ReLayout re_layout(uint32_t n, uint32_t nnz, uint32_t d, uint32_t p, uint32_t m)
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
// The kernel.
// ---------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(G) re_solver_kernel(const ReArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double red[2 * kMaxWarps * kRedK];
    __shared__ double s_rho[GDMIX_MAX_M], s_alpha[GDMIX_MAX_M];
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

        const ReLayout L = re_layout(n, nnz, d, p, (uint32_t)m);
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
        int col = 0, head = 0;
        double theta = 1.0;
        bool done = gmax <= a.o.pgtol;

        while (!done) {
            // direction: two-loop recursion in dv
            for (uint32_t j = tid; j < p; j += G) dv[j] = g[j];
            for (int k = col - 1; k >= 0; k--) {
                const int s = (head + k) % m;
                const double *Ss = Sh + (size_t)s * p, *Ys = Yh + (size_t)s * p;
                double v[1] = {0.0};
                for (uint32_t j = tid; j < p; j += G) v[0] = fma(Ss[j], dv[j], v[0]);
                group_sum<G, 1>(v, red, flip);
                const double al = s_rho[s] * v[0];
                if (tid == 0) s_alpha[s] = al;
                for (uint32_t j = tid; j < p; j += G) dv[j] = fma(-al, Ys[j], dv[j]);
            }
            for (uint32_t j = tid; j < p; j += G) dv[j] = dv[j] / theta;
            group_sync<G>();  // s_alpha visible
            for (int k = 0; k < col; k++) {
                const int s = (head + k) % m;
                const double *Ss = Sh + (size_t)s * p, *Ys = Yh + (size_t)s * p;
                double v[1] = {0.0};
                for (uint32_t j = tid; j < p; j += G) v[0] = fma(Ys[j], dv[j], v[0]);
                group_sum<G, 1>(v, red, flip);
                const double c2 = s_alpha[s] - s_rho[s] * v[0];
                for (uint32_t j = tid; j < p; j += G) dv[j] = fma(Ss[j], c2, dv[j]);
            }
            double v2[2] = {0.0, 0.0};
            for (uint32_t j = tid; j < p; j += G) {
                const double dj = -dv[j];
                dv[j] = dj;
                v2[0] = fma(dj, dj, v2[0]);
                v2[1] = fma(g[j], dj, v2[1]);
            }
            group_sum<G, 2>(v2, red, flip);
            const double dnorm = sqrt(v2[0]);
            gd = v2[1];

            // line search (lnsrlb + dcsrch)
            double stp = (iter == 0) ? fmin(1.0 / dnorm, stpmx) : 1.0;
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
                if (col == 0) { status = GDMIX_SOLVE_ABNORMAL; iter++; break; }
                col = 0; head = 0; theta = 1.0;
                continue;
            }
            iter++;
            // accept: (x, g) <-> (xt, gt); the old iterate stays reachable for y = g - g_old
            { double *t = x; x = xt; xt = t; t = g; g = gt; gt = t; }
            gmax = gmax_t;

            if (iter >= a.o.max_iter || nfev > a.o.max_fun) { status = GDMIX_SOLVE_MAXITER; break; }
            if (gmax <= a.o.pgtol) break;
            if ((fold - f) <= epsmch * a.o.factr * max3(fabs(fold), fabs(f), 1.0)) break;

            // curvature pair
            double vr[1] = {0.0};
            for (uint32_t j = tid; j < p; j += G) {
                const double yj = g[j] - gt[j];
                gt[j] = yj;
                vr[0] = fma(yj, yj, vr[0]);
            }
            group_sum<G, 1>(vr, red, flip);
            const double rr = vr[0];
            double dr, ddum;
            if (stp == 1.0) { dr = gd - gdold; ddum = -gdold; }
            else { dr = (gd - gdold) * stp; ddum = -gdold * stp; }
            if (dr <= epsmch * ddum || m == 0) continue;  // skip the update
            int slot;
            if (col < m) { slot = (head + col) % m; col++; }
            else { slot = head; head = (head + 1) % m; }
            double *Ss = Sh + (size_t)slot * p, *Ys = Yh + (size_t)slot * p;
            for (uint32_t j = tid; j < p; j += G) {
                Ss[j] = stp * dv[j];
                Ys[j] = gt[j];
            }
            if (tid == 0) s_rho[slot] = 1.0 / dr;
            theta = rr / dr;
            group_sync<G>();  // s_rho visible
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
