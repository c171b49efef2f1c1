// fe_lbfgs.cuh -- device-resident, replicated L-BFGS-B (unbounded) for the FIXED-effect solve.
//
// The reference runs scipy.optimize.fmin_l_bfgs_b on every worker around "stream my shard, all-reduce value and
// gradient" (gdmix-trainer/src/gdmix/models/custom/fixed_effect_lr_lbfgs_model.py:635-643, :394-404).  host_lbfgs.h
// is that solver as a host state machine; this file is the same state machine with every vector (x, g, d, the
// previous iterate and gradient, the m curvature pairs) resident in HBM, so that an objective evaluation never
// leaves the device: kernels -> all-reduce of fg on the same stream -> the launches below -> next evaluation.
// Only a 64-byte status record crosses to the host per evaluation.
//
// One call of enqueue() consumes the all-reduced fg = [f | g] at the current x and either writes the next trial
// point into x, or finishes.  The scalar decisions (dcsrch, stop tests, pair acceptance) are taken by single-thread
// kernels between the vector kernels; every vector kernel reads the decision flags from the state record, so the
// launch sequence is the same for every evaluation (capturable in a CUDA graph, no host round trip inside).
//
// Inner products: per block of kLbBlock entries a fixed-order tree (4 strided partial sums per thread, xor
// butterfly per warp, warps in order), block sums added in block order by every consumer.  The order does not
// depend on the grid or on the rank, so replicated ranks keep bit-identical state, like the reference's replicated
// scipy instances fed the same reduced (f, g).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "linesearch.cuh"

namespace gdmix {

constexpr int kLbThreads = 256;
constexpr int kLbBlock = 1024;      // entries per CTA (4 per thread)
constexpr int kLbMaxM = 32;

struct FeLbState {
    // configuration
    int64_t n;
    int32_t m, max_iter, max_ls, max_fun;
    double factr, pgtol;
    // solver state (host_lbfgs.h: Lbfgs)
    int32_t col, head, iter, nfev, state, status;   // state: 0 before the first evaluation, 1 running, 2 done
    int32_t ifun, iback, lstask, pending_begin;
    double theta, f, stp, fold, gd, gdold;
    LineSearch ls;
    double rho[kLbMaxM], alpha[kLbMaxM];
    // decisions of the current enqueue() (read by the vector kernels)
    int32_t do_pair, do_restore, do_begin, do_newx, do_store, store_slot;
    int32_t pad0, pad1;
};

// what the host polls (pinned memory)
struct FeLbStatus {
    int32_t task;      // 0 done, 1 evaluate at x and call again, 2 call again without evaluating (restart)
    int32_t nit, nfev, status;
    double f;
};

struct FeLbBuffers {
    FeLbState *st;
    double *x, *fg;           // caller's: x[n], fg[1 + n] (value, gradient)
    double *d, *t, *r, *q;    // direction, previous iterate, previous gradient, two-loop work vector
    double *S, *Y;            // [m][n]
    double *part;             // [4][nb] block partials (ping-pong pairs)
    int32_t nb;
};

__device__ __forceinline__ double lb_block_sum(double v, double *sh)
{
    // v: this thread's partial; fixed-order tree: xor butterfly inside the warp, warps in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kLbThreads / 32; k++) s += sh[k];
    return s;
}

__device__ __forceinline__ double lb_block_max(double v, double *sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kLbThreads / 32; k++) s = fmax(s, sh[k]);
    return s;
}

// sum of the nb block partials in block order (every thread of every CTA gets the same bits)
__device__ __forceinline__ double lb_total(const double *part, const int nb)
{
    double s = 0.0;
    for (int b = 0; b < nb; b++) s += part[b];
    return s;
}

// ---- 1: g.d and max|g| partials at the point just evaluated -------------------------------------------------
__global__ void __launch_bounds__(kLbThreads) lb_dots_kernel(const FeLbBuffers B)
{
    __shared__ double sh[kLbThreads / 32];
    const FeLbState &S = *B.st;
    if (S.state == 2) return;
    const double *g = B.fg + 1;
    const int64_t base = (int64_t)blockIdx.x * kLbBlock;
    double s = 0.0, mx = 0.0;
#pragma unroll
    for (int k = 0; k < kLbBlock / kLbThreads; k++) {
        const int64_t i = base + threadIdx.x + k * kLbThreads;
        if (i < S.n) {
            const double gi = g[i];
            mx = fmax(mx, fabs(gi));
            if (S.state == 1) s += gi * B.d[i];
        }
    }
    s = lb_block_sum(s, sh);
    mx = lb_block_max(mx, sh);
    if (threadIdx.x == 0) { B.part[blockIdx.x] = s; B.part[B.nb + blockIdx.x] = mx; }
}

// line search failed: back to the previous iterate (host_lbfgs.h: advance(), the tail of the loop)
__device__ inline void lb_fail(FeLbState &S)
{
    S.do_restore = 1;
    S.f = S.fold;
    if (S.col == 0) { S.iter++; S.status = 2; S.state = 2; return; }
    S.col = 0; S.head = 0; S.theta = 1.0;
    S.do_begin = 1;
}

// ---- 2: line-search decision at the evaluated point ----------------------------------------------------------
__global__ void lb_decide1_kernel(const FeLbBuffers B)
{
    if (threadIdx.x != 0) return;
    FeLbState &S = *B.st;
    S.do_pair = S.do_restore = S.do_begin = S.do_newx = S.do_store = 0;
    if (S.state == 2) return;
    if (S.pending_begin) {          // a restart decided at the end of the previous call: no new evaluation to consume
        S.pending_begin = 0;
        S.do_begin = 1;
        return;
    }
    double gnorm = 0.0;
    for (int b = 0; b < B.nb; b++) gnorm = fmax(gnorm, B.part[B.nb + b]);
    S.f = B.fg[0];
    if (S.state == 0) {
        S.nfev = 1;
        if (gnorm <= S.pgtol) { S.status = 0; S.state = 2; return; }
        S.state = 1;
        S.do_begin = 1;
        return;
    }
    S.nfev++;
    S.gd = lb_total(B.part, B.nb);
    S.lstask = dcsrch(S.stp, S.f, S.gd, 1e-3, 0.9, 0.1, 0.0, 1e10, S.lstask, S.ls);
    if (S.lstask == LS_CONV || S.lstask == LS_WARN) {
        // end_iteration()
        const double epsmch = 2.220446049250313e-16;
        S.iter++;
        if (S.iter >= S.max_iter || S.nfev > S.max_fun) { S.status = 1; S.state = 2; return; }
        if (gnorm <= S.pgtol) { S.status = 0; S.state = 2; return; }
        if ((S.fold - S.f) <= epsmch * S.factr * max3(fabs(S.fold), fabs(S.f), 1.0)) { S.status = 0; S.state = 2; return; }
        S.do_pair = 1;
        S.do_begin = 1;
        return;
    }
    if (S.lstask != LS_ERROR) {
        S.ifun++; S.iback = S.ifun - 1;
        if (S.iback < S.max_ls) { S.do_newx = 1; return; }
    }
    lb_fail(S);
}

// ---- 3: y = g - g_prev (into r), |y|^2 partials, s = stp * d (into d); or the restore after a failed search ----
__global__ void __launch_bounds__(kLbThreads) lb_pair_kernel(const FeLbBuffers B)
{
    __shared__ double sh[kLbThreads / 32];
    const FeLbState &S = *B.st;
    if (!S.do_pair && !S.do_restore) return;
    double *g = B.fg + 1;
    const int64_t base = (int64_t)blockIdx.x * kLbBlock;
    double rr = 0.0;
#pragma unroll
    for (int k = 0; k < kLbBlock / kLbThreads; k++) {
        const int64_t i = base + threadIdx.x + k * kLbThreads;
        if (i < S.n) {
            if (S.do_restore) {
                B.x[i] = B.t[i];
                g[i] = B.r[i];
            } else {
                const double y = g[i] - B.r[i];
                B.r[i] = y;
                rr += y * y;
                if (S.stp != 1.0) B.d[i] *= S.stp;
            }
        }
    }
    if (S.do_pair) {
        rr = lb_block_sum(rr, sh);
        if (threadIdx.x == 0) B.part[2 * B.nb + blockIdx.x] = rr;
    }
}

// ---- 4: accept / skip the curvature pair ------------------------------------------------------------------------
__global__ void lb_decide2_kernel(const FeLbBuffers B)
{
    if (threadIdx.x != 0) return;
    FeLbState &S = *B.st;
    if (!S.do_pair) return;
    const double epsmch = 2.220446049250313e-16;
    const double rr = lb_total(B.part + 2 * B.nb, B.nb);
    double dr, ddum;
    if (S.stp == 1.0) { dr = S.gd - S.gdold; ddum = -S.gdold; }
    else { dr = (S.gd - S.gdold) * S.stp; ddum = -S.gdold * S.stp; }
    if (!(dr <= epsmch * ddum) && S.m > 0) {
        int slot;
        if (S.col < S.m) { slot = (S.head + S.col) % S.m; S.col++; }
        else { slot = S.head; S.head = (S.head + 1) % S.m; }
        S.rho[slot] = 1.0 / dr;
        S.theta = rr / dr;
        S.do_store = 1;
        S.store_slot = slot;
    }
}

// ---- 5: store the pair; q = g; first inner product of the backward loop -----------------------------------------
__global__ void __launch_bounds__(kLbThreads) lb_store_kernel(const FeLbBuffers B)
{
    __shared__ double sh[kLbThreads / 32];
    const FeLbState &S = *B.st;
    if (!S.do_begin) return;
    const double *g = B.fg + 1;
    const int64_t base = (int64_t)blockIdx.x * kLbBlock;
    const int s_first = S.col > 0 ? (S.head + S.col - 1) % S.m : 0;
    const double *Sv = B.S + (size_t)s_first * S.n;
    double dot = 0.0;
#pragma unroll
    for (int k = 0; k < kLbBlock / kLbThreads; k++) {
        const int64_t i = base + threadIdx.x + k * kLbThreads;
        if (i < S.n) {
            if (S.do_store) {
                B.S[(size_t)S.store_slot * S.n + i] = B.d[i];
                B.Y[(size_t)S.store_slot * S.n + i] = B.r[i];
            }
            const double gi = g[i];
            B.q[i] = gi;
            if (S.col > 0) dot += (S.do_store && s_first == S.store_slot ? B.d[i] : Sv[i]) * gi;
        }
    }
    if (S.col > 0) {
        dot = lb_block_sum(dot, sh);
        if (threadIdx.x == 0) B.part[blockIdx.x] = dot;     // ping buffer 0
    }
}

// ---- 6: backward loop, step j (k = col - 1 - j): alpha, q -= alpha * Y_s, next inner product -------------------
__global__ void __launch_bounds__(kLbThreads) lb_loop1_kernel(const FeLbBuffers B, const int j)
{
    __shared__ double sh[kLbThreads / 32];
    FeLbState &S = *B.st;
    if (!S.do_begin || j >= S.col) return;
    const int k = S.col - 1 - j;
    const int s = (S.head + k) % S.m;
    const double *pin = B.part + (size_t)(j & 1) * B.nb;
    double *pout = B.part + (size_t)((j + 1) & 1) * B.nb;
    const double alpha = S.rho[s] * lb_total(pin, B.nb);
    const bool more = k > 0;
    const int s_next = more ? (S.head + k - 1) % S.m : 0;
    const double *Yv = B.Y + (size_t)s * S.n, *Sn = B.S + (size_t)s_next * S.n;
    const int64_t base = (int64_t)blockIdx.x * kLbBlock;
    double dot = 0.0;
#pragma unroll
    for (int kk = 0; kk < kLbBlock / kLbThreads; kk++) {
        const int64_t i = base + threadIdx.x + kk * kLbThreads;
        if (i < S.n) {
            const double qi = B.q[i] + (-alpha) * Yv[i];
            B.q[i] = qi;
            if (more) dot += Sn[i] * qi;
        }
    }
    if (more) {
        dot = lb_block_sum(dot, sh);
        if (threadIdx.x == 0) pout[blockIdx.x] = dot;
    }
    // alpha is needed again by the forward loop; written after every CTA's read of rho/part is past (different words)
    if (blockIdx.x == 0 && threadIdx.x == 0) S.alpha[s] = alpha;
}

// ---- 7: q /= theta; first inner product of the forward loop (Y_head . q) --------------------------------------
__global__ void __launch_bounds__(kLbThreads) lb_scale_kernel(const FeLbBuffers B)
{
    __shared__ double sh[kLbThreads / 32];
    const FeLbState &S = *B.st;
    if (!S.do_begin) return;
    const int s0 = S.head % (S.m > 0 ? S.m : 1);
    const double *Yv = B.Y + (size_t)s0 * S.n;
    const int64_t base = (int64_t)blockIdx.x * kLbBlock;
    double dot = 0.0;
#pragma unroll
    for (int k = 0; k < kLbBlock / kLbThreads; k++) {
        const int64_t i = base + threadIdx.x + k * kLbThreads;
        if (i < S.n) {
            const double qi = B.q[i] / S.theta;
            B.q[i] = qi;
            if (S.col > 0) dot += Yv[i] * qi;
        }
    }
    if (S.col > 0) {
        dot = lb_block_sum(dot, sh);
        if (threadIdx.x == 0) B.part[blockIdx.x] = dot;     // ping buffer 0
    }
}

// ---- 8: forward loop, step k: beta, q += (alpha - beta) * S_s, next inner product -------------------------------
__global__ void __launch_bounds__(kLbThreads) lb_loop2_kernel(const FeLbBuffers B, const int k)
{
    __shared__ double sh[kLbThreads / 32];
    const FeLbState &S = *B.st;
    if (!S.do_begin || k >= S.col) return;
    const int s = (S.head + k) % S.m;
    const double *pin = B.part + (size_t)(k & 1) * B.nb;
    double *pout = B.part + (size_t)((k + 1) & 1) * B.nb;
    const double beta = S.rho[s] * lb_total(pin, B.nb);
    const double c = S.alpha[s] - beta;
    const bool more = k + 1 < S.col;
    const int s_next = more ? (S.head + k + 1) % S.m : 0;
    const double *Sv = B.S + (size_t)s * S.n, *Yn = B.Y + (size_t)s_next * S.n;
    const int64_t base = (int64_t)blockIdx.x * kLbBlock;
    double dot = 0.0;
#pragma unroll
    for (int kk = 0; kk < kLbBlock / kLbThreads; kk++) {
        const int64_t i = base + threadIdx.x + kk * kLbThreads;
        if (i < S.n) {
            const double qi = B.q[i] + c * Sv[i];
            B.q[i] = qi;
            if (more) dot += Yn[i] * qi;
        }
    }
    if (more) {
        dot = lb_block_sum(dot, sh);
        if (threadIdx.x == 0) pout[blockIdx.x] = dot;
    }
}

// ---- 9: d = -q; |d|^2 and g.d partials; remember the iterate and its gradient ----------------------------------
__global__ void __launch_bounds__(kLbThreads) lb_direction_kernel(const FeLbBuffers B)
{
    __shared__ double sh[kLbThreads / 32];
    const FeLbState &S = *B.st;
    if (!S.do_begin) return;
    const double *g = B.fg + 1;
    const int64_t base = (int64_t)blockIdx.x * kLbBlock;
    double dd = 0.0, gd = 0.0;
#pragma unroll
    for (int k = 0; k < kLbBlock / kLbThreads; k++) {
        const int64_t i = base + threadIdx.x + k * kLbThreads;
        if (i < S.n) {
            const double di = -B.q[i], gi = g[i];
            B.d[i] = di;
            dd += di * di;
            gd += gi * di;
            B.t[i] = B.x[i];
            B.r[i] = gi;
        }
    }
    dd = lb_block_sum(dd, sh);
    gd = lb_block_sum(gd, sh);
    if (threadIdx.x == 0) { B.part[2 * B.nb + blockIdx.x] = dd; B.part[3 * B.nb + blockIdx.x] = gd; }
}

// ---- 10: first trial step of the new iteration -------------------------------------------------------------------
__global__ void lb_decide3_kernel(const FeLbBuffers B, FeLbStatus *out)
{
    if (threadIdx.x != 0) return;
    FeLbState &S = *B.st;
    if (S.do_begin && S.state == 1) {
        const double dnorm = sqrt(lb_total(B.part + 2 * B.nb, B.nb));
        S.stp = (S.iter == 0) ? fmin(1.0 / dnorm, 1e10) : 1.0;
        S.fold = S.f;
        S.ifun = 0; S.iback = 0; S.lstask = LS_START;
        S.gd = lb_total(B.part + 3 * B.nb, B.nb);
        S.gdold = S.gd;
        bool ok = !(S.gd >= 0.0);
        if (ok) {
            S.lstask = dcsrch(S.stp, S.f, S.gd, 1e-3, 0.9, 0.1, 0.0, 1e10, S.lstask, S.ls);
            ok = S.lstask == LS_FG;
        }
        if (ok) {
            S.ifun = 1; S.iback = 0;
            if (S.iback < S.max_ls) S.do_newx = 1; else ok = false;
        }
        if (!ok) {
            // not a descent direction / the search cannot start: x is still the previous iterate, nothing to restore
            if (S.col == 0) { S.iter++; S.status = 2; S.state = 2; }
            else { S.col = 0; S.head = 0; S.theta = 1.0; S.pending_begin = 1; }
        }
    }
    out->task = S.state == 2 ? 0 : (S.pending_begin ? 2 : 1);
    out->nit = S.iter; out->nfev = S.nfev; out->status = S.status; out->f = S.f;
}

// ---- 11: the next trial point -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kLbThreads) lb_newx_kernel(const FeLbBuffers B)
{
    const FeLbState &S = *B.st;
    if (!S.do_newx) return;
    const int64_t base = (int64_t)blockIdx.x * kLbBlock;
#pragma unroll
    for (int k = 0; k < kLbBlock / kLbThreads; k++) {
        const int64_t i = base + threadIdx.x + k * kLbThreads;
        if (i < S.n) B.x[i] = S.stp * B.d[i] + B.t[i];
    }
}

}  // namespace gdmix
