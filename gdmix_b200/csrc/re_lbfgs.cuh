// re_lbfgs.cuh -- compact-form L-BFGS direction (see banner below).
#pragma once
#include "re_common.cuh"
namespace gdmix {
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

// The warp-sized part of the update, run by ONE full warp after dense[tot..] holds the 2*MT+2 totals
// [S^T g | Y^T g | y.y | y.g]: appends / replaces the pair, refreshes R^-1, Y^T Y, D and leaves the
// coefficient vectors u (dense[cu..]) and w (dense[cw..]) of H g = gamma g + S u - gamma Y w, and the new
// scaling theta and gamma = 1 / theta in tot[2 MT], tot[2 MT + 1] (over the two totals it no longer needs) for
// every thread of the group to pick up.  Lane i < MT returns its u_i, w_i (0 for empty slots).
//
// This is the one serial stretch of an iteration (the other warps of the group wait for it), so it is laid out
// for latency: three dependent m x m products, not four -- the new column of R^-1, -R^-1 (S^T y_new) / dr,
// and w = R_new^-1 p1 come out of ONE sweep over the old R^-1 (w_i = sum_{j != new} R^-1_ij p1_j + R^-1_i,new
// p1_new) -- and no division on the dependent chain (1 / dr and gamma are formed beside the first sweep).
template <int MT>
__device__ __forceinline__ void lbfgs_small_update(const Lbfgs &L, const bool update, const int newslot,
                                                   const uint32_t dotmask, const double stp, const double dr,
                                                   const double gd_new, double *dense, double &uv_out,
                                                   double &wv_out, double &theta_out, double &gamma_out)
{
    using DN = Dense<MT>;
    const uint32_t lane = threadIdx.x & 31;
    double *tot = dense + DN::tot;
    const int i = lane;
    const bool in = i < MT;
    const bool old_i = in && ((dotmask >> i) & 1u);
    const bool new_i = update && i == newslot;
    double p1i = old_i ? tot[i] : 0.0, p2i = old_i ? tot[MT + i] : 0.0;
    const double yyt = tot[2 * MT], ygt = tot[2 * MT + 1];
    double *Rinv = dense + DN::rinv, *YY = dense + DN::yy, *Dg = dense + DN::d;
    double *p1old = dense + DN::p1old, *p2old = dense + DN::p2old;
    double *ta = dense + DN::ta, *tb = dense + DN::tb, *cwv = dense + DN::cw;
    const double inv_dr = update ? 1.0 / dr : 0.0;
    const double theta = update ? yyt * inv_dr : L.theta;
    const double gamma = 1.0 / theta;
    const double rc = (update && old_i) ? p1i - p1old[i] : 0.0;  // s_i . y_new
    const double yc = (update && old_i) ? p2i - p2old[i] : 0.0;  // y_i . y_new
    const double p1new = stp * gd_new;                           // s_new . g_new
    if (new_i) { p1i = p1new; p2i = ygt; }
    const uint32_t valid = update ? (L.valid | (1u << newslot)) : L.valid;
    const bool val_i = in && ((valid >> i) & 1u);
    if (in) {
        ta[i] = rc;                                   // zero for the new slot and for empty ones
        tb[i] = (val_i && !new_i) ? p1i : 0.0;
        p1old[i] = p1i; p2old[i] = p2i;
    }
    __syncwarp();
    double acc = 0.0, wvp = 0.0;
    if (in) {
#pragma unroll
        for (int j = 0; j < MT; j++) {
            const double r = Rinv[i * MT + j];
            acc = fma(r, ta[j], acc);
            wvp = fma(r, tb[j], wvp);
        }
    }
    double wv = val_i ? wvp : 0.0;
    __syncwarp();                                     // every lane is done reading the old R^-1
    if (update) {
        const double cnew = old_i ? -acc * inv_dr : 0.0;          // new column of R^-1 (old rows)
        if (in) {
            wv = new_i ? inv_dr * p1new : (old_i ? fma(cnew, p1new, wvp) : 0.0);
            Rinv[i * MT + newslot] = new_i ? inv_dr : cnew;
            Rinv[newslot * MT + i] = new_i ? inv_dr : 0.0;
            YY[i * MT + newslot] = new_i ? yyt : yc;
            YY[newslot * MT + i] = new_i ? yyt : yc;
        }
        if (new_i) Dg[i] = dr;
    }
    if (in) cwv[i] = wv;
    __syncwarp();
    double yw = 0.0;
    if (in) {
#pragma unroll
        for (int j = 0; j < MT; j++) yw = fma(YY[i * MT + j], cwv[j], yw);
    }
    const double tv = val_i ? fma(Dg[i], wv, gamma * (yw - p2i)) : 0.0;
    if (in) ta[i] = tv;
    __syncwarp();
    double uv = 0.0;
    if (in) {
#pragma unroll
        for (int j = 0; j < MT; j++) uv = fma(Rinv[j * MT + i], ta[j], uv);
    }
    if (in) dense[DN::cu + i] = uv;
    if (lane == 0) { tot[2 * MT] = theta; tot[2 * MT + 1] = gamma; }
    uv_out = uv; wv_out = wv; theta_out = theta; gamma_out = gamma;
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
    if (update) { newslot = (L.col < m) ? L.head + L.col : L.head; if (newslot >= m) newslot -= m; }   // head, col < m
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
        double uv, wv, th, ga;
        lbfgs_small_update<MT>(L, update, newslot, dotmask, stp, dr, gd_new, dense, uv, wv, th, ga);
    }
    group_sync<G>();
    L.theta = dense[DN::tot + 2 * MT];
    if (update) {
        L.valid |= (1u << newslot);
        if (L.col < m) L.col++; else L.head = (L.head + 1 == m) ? 0 : L.head + 1;
    }

    // ---- pass H2: dv = -gamma g - S u + gamma Y w ------------------------------------------------------
    const double gamma = dense[DN::tot + 2 * MT + 1];
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

}  // namespace gdmix
