// re_common.cuh -- shared declarations of the random-effect hot path on sm_100a (kernel in re_solver.cuh).
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
    const int32_t *todo;        // optional: solve entities todo[0 .. *todo_count) instead of 0 .. n_entities
    const int32_t *todo_count;
    int32_t *defer_list;        // optional: entities this kernel cannot hold on chip go here instead of failing
    int32_t *defer_count;
    int32_t *giant_list;        // optional: of those, entities with at least giant_rows samples go here -- they are
    int32_t *giant_count;       //   solved by a CLUSTER of CTAs that splits the samples (re_kernel.cuh)
    int32_t giant_rows;
    unsigned char *arena;       // per-CTA global scratch for history that does not fit on chip
    unsigned long long arena_stride;
    int32_t mode;
    int32_t hist_global;        // keep the (S, Y) history in the global arena even if it would fit on chip
    uint32_t smem_bytes;        // dynamic shared memory given to the kernel
    // L2 sweep (gdmix_re_fit_sweep): n_l2 models per entity from ONE staged copy of its block.  Model j uses
    // l2_sweep[j] and writes theta_out + j * sweep_coef_stride, {f_out, nit, nfev, status} + j * n_entities.
    // n_l2 == 0: the single model of o.l2.
    double l2_sweep[GDMIX_MAX_SWEEP];
    int32_t n_l2;
    int64_t sweep_coef_stride;
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

// ---------------------------------------------------------------------------------------------------------
// The three transcendental terms of one logistic sample, t = exp(-|z|), log(1 + t) and 1 / (1 + t), written for the
// solver's rows pass: one sample per thread, nothing else to overlap with, so what counts is instruction count and
// the length of the dependent chain -- the library's exp / log / division are ~235 instructions with internal
// branches; this is ~80, branch-free, with the polynomials in Estrin form and the two reciprocals side by side.
// Accuracy (tests/test_re_gpu_parity.py::test_logistic_terms_accuracy): t, log(1 + t), 1 / (1 + t) within 2 ulp of
// the library's results for |z| <= 708; beyond that t is e^-708 (3e-308) instead of a denormal / zero.
//   exp(-a) = 2^-k e^x, k = rint(a log2 e), x = k ln2 - a in [-ln2/2, ln2/2], e^x = 1 + (x + x^2 R(x)), Taylor to x^13
//   log(u), u = 1 + t in (1, 2]: m = u or u/2 (above sqrt 2), f = m - 1, s = f / (2 + f),
//            log m = 2 s + 2 s^3 T(s^2), T = 1/3 + s^2/5 + ... + s^20/23, |s| <= 0.1716
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double rcp_newton(const double a)   // a normal and away from the range limits
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    double e = fma(-a, x, 1.0);
    e = fma(e, e, e);
    x = fma(x, e, x);
    e = fma(-a, x, 1.0);
    return fma(x, e, x);
}

__device__ __forceinline__ void logistic_terms(const double z, double &t, double &log1pt, double &inv1pt)
{
    const double a = fmin(fabs(z), 708.0);
    const double magic = 6755399441055744.0;   // 1.5 * 2^52: adding it rounds to an integer in the low word
    const double kd = fma(a, 1.4426950408889634074, magic);
    const int k = __double2loint(kd);
    const double kf = kd - magic;
    double x = fma(kf, 6.93147180369123816490e-01, -a);
    x = fma(kf, 1.90821492927058770002e-10, x);
    {
        const double x2 = x * x, x4 = x2 * x2, x8 = x4 * x4;
        const double p0 = fma(x, 1.0 / 6.0, 0.5);
        const double p1 = fma(x, 1.0 / 120.0, 1.0 / 24.0);
        const double p2 = fma(x, 1.0 / 5040.0, 1.0 / 720.0);
        const double p3 = fma(x, 1.0 / 362880.0, 1.0 / 40320.0);
        const double p4 = fma(x, 1.0 / 39916800.0, 1.0 / 3628800.0);
        const double p5 = fma(x, 1.0 / 6227020800.0, 1.0 / 479001600.0);
        const double q0 = fma(x2, p1, p0), q1 = fma(x2, p3, p2), q2 = fma(x2, p5, p4);
        const double R = fma(x8, q2, fma(x4, q1, q0));
        const double ex = 1.0 + fma(x2, R, x);
        t = ex * __hiloint2double((1023 - k) << 20, 0);
    }
    const double u = 1.0 + t;
    const bool big = u > 1.41421356237309514547;
    const double m = big ? 0.5 * u : u;
    const double f = m - 1.0, d = m + 1.0;
    const double rd = rcp_newton(d);
    inv1pt = rcp_newton(u);
    double s = f * rd;
    s = fma(fma(-d, s, f), rd, s);
    {
        const double s2 = s * s, s4 = s2 * s2, s8 = s4 * s4;
        const double p0 = fma(s2, 1.0 / 5.0, 1.0 / 3.0);
        const double p1 = fma(s2, 1.0 / 9.0, 1.0 / 7.0);
        const double p2 = fma(s2, 1.0 / 13.0, 1.0 / 11.0);
        const double p3 = fma(s2, 1.0 / 17.0, 1.0 / 15.0);
        const double p4 = fma(s2, 1.0 / 21.0, 1.0 / 19.0);
        const double q0 = fma(s4, p1, p0), q1 = fma(s4, p3, p2), q2 = fma(s4, 1.0 / 23.0, p4);
        const double s16 = s8 * s8;
        const double T = fma(s16, q2, fma(s8, q1, q0));
        const double two_s = s + s;
        const double l = fma(two_s * s2, T, two_s);
        log1pt = big ? l + 6.93147180559945286227e-01 : l;
    }
}

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
    // the callers use the reduction as the group's barrier as well (shared-memory writes before it are read after it):
    // with one warp per group the shuffles converge the lanes but order no memory -- __syncwarp does
    if (G == 32) { __syncwarp(); return; }
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
    if (G == 32) { __syncwarp(); return v; }
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

// One barrier for a sum and a max together.
template <int G>
__device__ __forceinline__ void group_sum_max(double &sum, double &mx, double *red, int &flip)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(kFull, sum, o);
        mx = fmax(mx, __shfl_xor_sync(kFull, mx, o));
    }
    if (G == 32) { __syncwarp(); return; }
    constexpr int W = G / 32;
    double *buf = red + flip * (kMaxWarps * kRedK);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { buf[warp * kRedK] = sum; buf[warp * kRedK + 1] = mx; }
    __syncthreads();
    double s = buf[0], m2 = buf[1];
#pragma unroll
    for (int w = 1; w < W; w++) { s += buf[w * kRedK]; m2 = fmax(m2, buf[w * kRedK + 1]); }
    sum = s; mx = m2;
    flip ^= 1;
}

}  // namespace gdmix
