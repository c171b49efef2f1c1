// re_passes.cuh -- staging of one entity into shared memory and the two passes over it.
//
// On-chip form of the entity's sample block X (n rows, d local features, fp32 values):
//   CSR  row i  -> pairs [rowst[i]/2, rowst[i+1]/2) of (float2 values, ushort2 local columns)
//   CSC  col c  -> pairs [colst[c]/2, colst[c+1]/2) of (float2 values, ushort2 rows), rows ascending
// Every segment starts on an even element and is zero-padded to even length, so one 64-bit and one
// 32-bit shared load fetch two non-zeros.  Bank conflicts between the lanes of a warp (which walk
// different segments) are avoided without padding by rotating each team's starting pair.
// A row (column) is walked by a team of TR (TC) adjacent lanes, TR = largest power of two with
// TR * n <= G: when the entity has fewer rows than the CTA has threads the idle threads shorten
// the per-row dependent chain instead.
//
// z = X1.theta + offset, the stable cross entropy and the residual follow
// binary_logistic_regression.py:84-110 (_loss) and :121-131 (_gradient) of the reference.
#pragma once
#include <cooperative_groups.h>
#include "re_common.cuh"

namespace gdmix {

struct Staged {
    uint32_t n, d, p, nnz, hi;
    uint32_t tr, tc;    // lanes per row / per column (powers of two, <= 32)
    uint32_t trs, tcs;  // log2 of the above
    const float *y, *w, *off;
    const uint32_t *rowst, *colst;
    const float2 *csr_val, *csc_val;
    const ushort2 *csr_col, *csc_row;
    double *r;
    double inv_n, l2;
    int reg_bias;  // intercept is regularised
};

__device__ __forceinline__ uint32_t team_shift(uint32_t threads, uint32_t units)
{
    uint32_t s = 0;
    while (s < 5 && (2u << s) * units <= threads) s++;
    return s;
}

// Reads the entity's CSR slice from HBM once and builds both on-chip forms.  `scratch` is W*d u32 of
// shared memory that is free during staging.  Returns false (uniformly) on an out-of-range column.
template <int G>
__device__ __forceinline__ bool stage_entity(const ReArgs &a, const int64_t r0, const int64_t q0, const uint32_t n,
                                             const uint32_t d, float *sy, float *sw, float *soff, uint32_t *rowst,
                                             uint32_t *colst, float *csr_val, uint16_t *csr_col, float *csc_val,
                                             uint16_t *csc_row, uint32_t *scratch, uint32_t *wtot, unsigned *s_bad)
{
    constexpr int W = G / 32;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- row starts: exclusive scan of the even-padded row lengths (thread t owns a contiguous row range)
    {
        const uint32_t ipt = (n + G - 1) / G;
        const uint32_t ibeg = min(n, tid * ipt), iend = min(n, ibeg + ipt);
        uint32_t mine = 0;
        for (uint32_t i = ibeg; i < iend; i++) {
            const uint32_t len = (uint32_t)(a.b.rowptr[r0 + i + 1] - a.b.rowptr[r0 + i]);
            mine += (len + 1u) & ~1u;
        }
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        if (G > 32) {
            if (lane == 31) wtot[warp] = incl;
            __syncthreads();
        }
        uint32_t base = incl - mine;
        if (G > 32)
            for (uint32_t w2 = 0; w2 < warp; w2++) base += wtot[w2];
        for (uint32_t i = ibeg; i < iend; i++) {
            rowst[i] = base;
            const uint32_t len = (uint32_t)(a.b.rowptr[r0 + i + 1] - a.b.rowptr[r0 + i]);
            base += (len + 1u) & ~1u;
        }
        if (iend == n && ibeg < n) rowst[n] = base;
        if (n == 0 && tid == 0) rowst[0] = 0;
    }
    for (uint32_t k = tid; k < W * d; k += G) scratch[k] = 0;
    group_sync<G>();

    // ---- per-sample scalars and the CSR copy (thread per row; each thread streams its own row) -------------
    for (uint32_t i = tid; i < n; i += G) {
        const int64_t gi = r0 + i;
        sy[i] = a.b.label[gi];
        sw[i] = a.b.weight ? a.b.weight[gi] : 1.0f;
        soff[i] = a.b.offset ? a.b.offset[gi] : 0.0f;
        const int64_t gs = a.b.rowptr[gi], ge = a.b.rowptr[gi + 1];
        uint32_t dst = rowst[i];
        unsigned bad = 0;
        for (int64_t q = gs; q < ge; q++, dst++) {
            const int32_t c = a.b.col[q];
            bad |= ((uint32_t)c >= d);
            csr_val[dst] = a.b.val[q];
            csr_col[dst] = (uint16_t)c;
        }
        if ((ge - gs) & 1) { csr_val[dst] = 0.0f; csr_col[dst] = 0; }
        if (bad) atomicOr(s_bad, 1u);
    }
    group_sync<G>();
    if (*s_bad) return false;

    // ---- CSC: per-warp column counts -> scan of even-padded totals -> row-ordered fill ------------------
    const uint32_t chunk = (n + W - 1) / W;
    const uint32_t rbeg = min(n, warp * chunk), rend = min(n, rbeg + chunk);
    for (uint32_t i = rbeg; i < rend; i++) {
        const uint32_t s = rowst[i];
        const uint32_t len = (uint32_t)(a.b.rowptr[r0 + i + 1] - a.b.rowptr[r0 + i]);
        for (uint32_t j = lane; j < len; j += 32) atomicAdd(&scratch[warp * d + csr_col[s + j]], 1u);
    }
    group_sync<G>();
    {
        const uint32_t ipt = (d + G - 1) / G;
        const uint32_t cbeg = min(d, tid * ipt), cend = min(d, cbeg + ipt);
        uint32_t mine = 0;
        for (uint32_t c = cbeg; c < cend; c++) {
            uint32_t cnt = 0;
            for (int w2 = 0; w2 < W; w2++) cnt += scratch[w2 * d + c];
            mine += (cnt + 1u) & ~1u;
        }
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        if (G > 32) {
            if (lane == 31) wtot[warp] = incl;
            __syncthreads();
        }
        uint32_t base = incl - mine;
        if (G > 32)
            for (uint32_t w2 = 0; w2 < warp; w2++) base += wtot[w2];
        for (uint32_t c = cbeg; c < cend; c++) {
            uint32_t pos = base;
            colst[c] = pos;
            for (int w2 = 0; w2 < W; w2++) {
                const uint32_t cnt = scratch[w2 * d + c];
                scratch[w2 * d + c] = pos;  // becomes warp w2's write cursor for column c
                pos += cnt;
            }
            if ((pos - base) & 1u) { csc_val[pos] = 0.0f; csc_row[pos] = 0; pos++; }
            base = pos;
        }
        if (cend == d && cbeg < d) colst[d] = base;
        if (d == 0 && tid == 0) colst[0] = 0;
    }
    group_sync<G>();
    for (uint32_t i = rbeg; i < rend; i++) {
        const uint32_t s = rowst[i];
        const uint32_t len = (uint32_t)(a.b.rowptr[r0 + i + 1] - a.b.rowptr[r0 + i]);
        for (uint32_t j0 = 0; j0 < len; j0 += 32) {
            const uint32_t j = j0 + lane;
            const bool act = j < len;
            const uint32_t c = act ? (uint32_t)csr_col[s + j] : (0x10000u + lane);
            const float v = act ? csr_val[s + j] : 0.0f;
            const unsigned grp = __match_any_sync(kFull, c);  // duplicate columns inside one row
            const uint32_t rank = __popc(grp & ((1u << lane) - 1u));
            uint32_t cur = 0;
            if (act) {
                cur = scratch[warp * d + c];
                csc_row[cur + rank] = (uint16_t)i;
                csc_val[cur + rank] = v;
            }
            __syncwarp();
            if (act && rank == (uint32_t)__popc(grp) - 1u) scratch[warp * d + c] = cur + rank + 1u;
            __syncwarp();
        }
    }
    group_sync<G>();
    return true;
}

// One segment (row or column) walked by a team of T lanes: sum over its pairs of val * vec[idx].
// `rot` staggers the starting pair between teams so that the lanes of a warp hit different banks.
__device__ __forceinline__ double team_dot(const float2 *val, const ushort2 *idx, const uint32_t pair0,
                                           const uint32_t npairs, const uint32_t T, const uint32_t t,
                                           const uint32_t rot, const double *vec)
{
    const uint32_t niter = (npairs + T - 1) / T;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (niter == 0) return 0.0;
    uint32_t it = (rot < niter) ? rot : rot % niter;
    // four pairs (eight non-zeros) per trip: all index/value loads are issued before the gathers, all
    // gathers before the FMAs, so one trip costs about one shared-memory round trip, not eight
    for (uint32_t k = 0; k < niter; k += 4) {
        const uint32_t nv = min(4u, niter - k);
        float2 v[4];
        ushort2 c[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t ij = it + j;
            if (ij >= niter) ij -= niter;
            const uint32_t pi = t + T * ij;
            const bool ok = (uint32_t)j < nv && pi < npairs;
            const uint32_t at = pair0 + (ok ? pi : 0u);
            v[j] = val[at];
            c[j] = idx[at];
            if (!ok) { v[j].x = 0.0f; v[j].y = 0.0f; }
        }
        double x[8];
#pragma unroll
        for (int j = 0; j < 4; j++) { x[2 * j] = vec[c[j].x]; x[2 * j + 1] = vec[c[j].y]; }
        s0 = fma((double)v[0].x, x[0], s0);
        s1 = fma((double)v[0].y, x[1], s1);
        s2 = fma((double)v[1].x, x[2], s2);
        s3 = fma((double)v[1].y, x[3], s3);
        s0 = fma((double)v[2].x, x[4], s0);
        s1 = fma((double)v[2].y, x[5], s1);
        s2 = fma((double)v[3].x, x[6], s2);
        s3 = fma((double)v[3].y, x[7], s3);
        it += 4;
        if (it >= niter) it -= niter;
    }
    return (s0 + s1) + (s2 + s3);
}

// Pass A (row teams): z = X1.xt + offset, weighted stable cross entropy, residual
// r_i = w_i (sigmoid(z_i) - y_i).  Per-thread partials {sum cost, sum r, sum xt_reg^2}.
template <int G>
__device__ __forceinline__ void pass_rows(const Staged &S, const double *xt, double (&part)[3])
{
    const uint32_t tid = threadIdx.x;
    const uint32_t T = S.tr, t = tid & (T - 1), team = tid >> S.trs, teams = G >> S.trs;
    const double b0 = S.hi ? xt[0] : 0.0;
    const double *xf = xt + S.hi;
    double fs = 0.0, rs = 0.0;
    for (uint32_t base = 0; base < S.n; base += teams) {
        const uint32_t i = base + team;
        const bool act = i < S.n;
        double z = 0.0;
        if (act) {
            const uint32_t s = S.rowst[i];
            z = team_dot(S.csr_val, S.csr_col, s >> 1, (S.rowst[i + 1] - s) >> 1, T, t, team & 31u, xf);
        }
        for (uint32_t o = T >> 1; o > 0; o >>= 1) z += __shfl_xor_sync(kFull, z, o);
        if (act && t == 0) {
            z = (z + b0) + (double)S.off[i];
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
    }
    double sq = 0.0;
    for (uint32_t jj = tid; jj < S.p; jj += G) {
        if (S.hi && jj == 0 && !S.reg_bias) continue;
        sq = fma(xt[jj], xt[jj], sq);
    }
    part[0] = fs; part[1] = rs; part[2] = sq;
}

// Pass B (column teams): g_j = (X1^T r + l2 * xt_reg)_j / n; the intercept (sum of r) is handled by the last
// thread.  Per-thread partials of g.dv and max|g|.
template <int G>
__device__ __forceinline__ void pass_cols(const Staged &S, const double *xt, const double *dv, double *gt,
                                          const double rsum, double &gd_part, double &gmax_part)
{
    const uint32_t tid = threadIdx.x;
    const uint32_t T = S.tc, t = tid & (T - 1), team = tid >> S.tcs, teams = G >> S.tcs;
    double gd = 0.0, gm = 0.0;
    for (uint32_t base = 0; base < S.d; base += teams) {
        const uint32_t c = base + team;
        const bool act = c < S.d;
        double acc = 0.0;
        if (act) {
            const uint32_t s = S.colst[c];
            acc = team_dot(S.csc_val, S.csc_row, s >> 1, (S.colst[c + 1] - s) >> 1, T, t, team & 31u, S.r);
        }
        for (uint32_t o = T >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
        if (act && t == 0) {
            const uint32_t j = c + S.hi;
            const double gj = (acc + S.l2 * xt[j]) * S.inv_n;
            gt[j] = gj;
            gd = fma(gj, dv[j], gd);
            gm = fmax(gm, fabs(gj));
        }
    }
    if (S.hi && tid == G - 1) {
        const double gj = (rsum + (S.reg_bias ? S.l2 * xt[0] : 0.0)) * S.inv_n;
        gt[0] = gj;
        gd = fma(gj, dv[0], gd);
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
    group_sum_max<G>(gdp, gmp, red, flip);
    gd = gdp;
    gmax = gmp;
}

// ---------------------------------------------------------------------------------------------------------
// Entities too large to stage: X stays in global memory (L2 / HBM) and is swept twice per evaluation.
// A warp owns 32 consecutive rows at a time: sweep 1 gives every lane the z of "its" row (lanes walk one row
// at a time, coalesced), the loss / residual arithmetic then runs lane-parallel, and sweep 2 folds r_i * x_i
// into the warp's PRIVATE copy of X^T r in shared memory, row after row -- so the sum has a fixed order (only a
// column repeated inside one row makes two lanes meet, hence the atomic).  The copies are added in warp order.
// ---------------------------------------------------------------------------------------------------------
struct BigEntity {
    int64_t r0;
    uint32_t n, d, p, hi;
    uint32_t ts;      // log2 lanes per row ("team"): 32 >> ts rows are walked per warp step
    double *gw;       // [W * (32 >> ts)][d] private copies of X^T r, one per team
    double inv_n, l2;
    int reg_bias;
};

// `copies` = private gradient copies per warp (1 .. 32): one per team of lanes that walks a row of its own
__host__ __device__ inline uint32_t big_layout_bytes(uint32_t p, uint32_t d, uint32_t W, uint32_t mt, uint32_t copies = 1)
{
    return 5u * align16(8 * p) + align16(8 * W * copies * d) + align16(8 * dense_doubles(mt)) +
           align16(8 * kMaxWarps * (2 * mt + 2));
}

// When the kernel runs as a thread-block cluster (entities with hundreds of thousands of samples, see
// re_solver_kernel), the row blocks are dealt round-robin over the cluster's CTAs: each CTA ends up with ITS
// part of X^T r and of the loss in its own shared memory, the cluster synchronises, and every CTA adds the C parts
// in rank order through distributed shared memory -- so all CTAs hold bit-identical f and g and run the same
// (replicated) solver steps without any further exchange, exactly as the ranks of the fixed-effect solve do.
template <int G>
__device__ __forceinline__ void evaluate_big(const ReArgs &a, const BigEntity &B, const double *xt, const double *dv,
                                             double *gt, double *red, int &flip, unsigned *s_bad, double *cl_part,
                                             double &f, double &gd, double &gmax)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t C = cluster.num_blocks(), crank = cluster.block_rank();
    // A warp owns 32 consecutive rows at a time and walks them 32 >> ts at a step, a team of 1 << ts lanes per
    // row (ts from the entity's mean non-zeros per row: 8-wide rows go four to a step and a step's loads are one
    // coalesced line).  Sweep 1: team sums of z by butterfly, handed to the lane whose index equals the row's
    // position in the block; the loss / residual arithmetic runs on 32 rows in 32 lanes; sweep 2 folds r_i x_i
    // into the TEAM's private copy of X^T r, row after row, so every copy has a fixed summation order (only a
    // column repeated inside one row makes two lanes meet, hence the atomic).  The row pointers of a block are
    // loaded once, one per lane.  Copies are added in (warp, team) order.
    constexpr uint32_t W = G / 32;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ts = B.ts, T = 1u << ts, RS = 32u >> ts;
    const uint32_t t = lane & (T - 1u), q = lane >> ts;
    {
        double *mine = B.gw + (size_t)warp * RS * B.d;
        for (uint32_t j = lane; j < RS * B.d; j += 32) mine[j] = 0.0;
    }
    double *gw = B.gw + ((size_t)warp * RS + q) * B.d;
    __syncwarp();
    const double b0 = B.hi ? xt[0] : 0.0;
    const double *xf = xt + B.hi;
    double fs = 0.0, rs = 0.0;
    unsigned bad = 0;
    const uint32_t nblk = (B.n + 31u) >> 5;
    for (uint32_t blk = crank * W + warp; blk < nblk; blk += W * C) {
        const uint32_t base = blk << 5, cnt = min(32u, B.n - base);
        int64_t rp0 = 0, rp1 = 0;
        if (lane < cnt) { rp0 = a.b.rowptr[B.r0 + base + lane]; rp1 = a.b.rowptr[B.r0 + base + lane + 1]; }
        const uint32_t nsteps = (cnt + RS - 1u) >> (5u - ts);
        double myz = 0.0;
        // four steps at a time: their first loads are all in flight before any is consumed (a step is one
        // dependent global round trip otherwise, and this kernel has few warps to hide it behind)
        for (uint32_t s0 = 0; s0 < nsteps; s0 += 4) {
            int64_t kk[4], ee[4];
            uint32_t cc[4];
            float vv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t row = ((s0 + u) * RS + q) & 31u;
                const int64_t qs = __shfl_sync(kFull, rp0, row), qe = __shfl_sync(kFull, rp1, row);
                const bool on = (s0 + u) < nsteps && (s0 + u) * RS + q < cnt;
                kk[u] = qs + t; ee[u] = on ? qe : qs;   // off: an empty range
                cc[u] = 0; vv[u] = 0.0f;
                if (kk[u] < ee[u]) { cc[u] = (uint32_t)a.b.col[kk[u]]; vv[u] = a.b.val[kk[u]]; }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                double z = 0.0;
                if (kk[u] < ee[u]) {
                    if (cc[u] < B.d) z = (double)vv[u] * xf[cc[u]]; else bad = 1;
                    for (int64_t k = kk[u] + T; k < ee[u]; k += T) {   // rows longer than the team
                        const uint32_t c = (uint32_t)a.b.col[k];
                        if (c < B.d) z = fma((double)a.b.val[k], xf[c], z); else bad = 1;
                    }
                }
                for (uint32_t m2 = T >> 1; m2 > 0; m2 >>= 1) z += __shfl_xor_sync(kFull, z, m2);
                const double v = __shfl_sync(kFull, z, (lane & (RS - 1u)) << ts);
                if ((lane >> (5u - ts)) == s0 + u) myz = v;   // lane j takes row j = s * RS + (j mod RS)
            }
        }
        double ri = 0.0;
        if (lane < cnt) {
            const int64_t gi = B.r0 + base + lane;
            const double z = (myz + b0) + (a.b.offset ? (double)a.b.offset[gi] : 0.0);
            const double yi = (double)a.b.label[gi], wi = a.b.weight ? (double)a.b.weight[gi] : 1.0;
            const double e = exp(-fabs(z));
            const double ce = fmax(z, 0.0) - z * yi + log(1.0 + e);
            fs = fma(wi, ce, fs);
            const double inv = 1.0 / (1.0 + e);
            const double sig = (z >= 0.0) ? inv : e * inv;
            ri = wi * (sig - yi);
            rs += ri;
        }
        for (uint32_t s0 = 0; s0 < nsteps; s0 += 4) {
            int64_t kk[4], ee[4];
            uint32_t cc[4];
            float vv[4];
            double rr[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t row = ((s0 + u) * RS + q) & 31u;
                rr[u] = __shfl_sync(kFull, ri, row);
                const int64_t qs = __shfl_sync(kFull, rp0, row), qe = __shfl_sync(kFull, rp1, row);
                const bool on = (s0 + u) < nsteps && (s0 + u) * RS + q < cnt;
                kk[u] = qs + t; ee[u] = on ? qe : qs;
                cc[u] = 0; vv[u] = 0.0f;
                if (kk[u] < ee[u]) { cc[u] = (uint32_t)a.b.col[kk[u]]; vv[u] = a.b.val[kk[u]]; }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {   // rows enter the team's copy one after the other: fixed order
                if (kk[u] < ee[u]) {
                    if (cc[u] < B.d) atomicAdd(&gw[cc[u]], (double)vv[u] * rr[u]);
                    for (int64_t k = kk[u] + T; k < ee[u]; k += T) {
                        const uint32_t c = (uint32_t)a.b.col[k];
                        if (c < B.d) atomicAdd(&gw[c], (double)a.b.val[k] * rr[u]);
                    }
                }
                __syncwarp();
            }
        }
    }
    if (bad) atomicOr(s_bad, 1u);
    double sq = 0.0;
    for (uint32_t jj = tid; jj < B.p; jj += G) {
        if (B.hi && jj == 0 && !B.reg_bias) continue;
        sq = fma(xt[jj], xt[jj], sq);
    }
    double part[3] = {fs, rs, sq};
    group_sum<G, 3>(part, red, flip);  // its barrier also completes every team's copy
    if (G == 32) __syncwarp();
    if (C > 1) {
        // this CTA's part of X^T r goes where copy 0 was (thread c reads column c of every copy, then writes it),
        // its loss / residual sums and its bad-column flag beside it
        for (uint32_t c = tid; c < B.d; c += G) {
            double acc = 0.0;
            for (uint32_t w2 = 0; w2 < W * RS; w2++) acc += B.gw[(size_t)w2 * B.d + c];
            B.gw[c] = acc;
        }
        if (tid == 0) { cl_part[0] = part[0]; cl_part[1] = part[1]; cl_part[2] = *s_bad ? 1.0 : 0.0; }
        cluster.sync();
        double t0 = 0.0, t1 = 0.0, t2 = 0.0;
        for (uint32_t r = 0; r < C; r++) {
            const double *q = cluster.map_shared_rank(cl_part, r);
            t0 += q[0]; t1 += q[1]; t2 += q[2];
        }
        part[0] = t0; part[1] = t1;
        if (t2 != 0.0 && tid == 0) *s_bad = 1u;
    }
    f = (part[0] + 0.5 * B.l2 * part[2]) * B.inv_n;
    double gdp = 0.0, gmp = 0.0;
    for (uint32_t c = tid; c < B.d; c += G) {
        double acc = 0.0;
        if (C > 1) {
            for (uint32_t r = 0; r < C; r++) acc += cluster.map_shared_rank(B.gw, r)[c];
        } else {
            for (uint32_t w2 = 0; w2 < W * RS; w2++) acc += B.gw[(size_t)w2 * B.d + c];
        }
        const uint32_t j = c + B.hi;
        const double gj = (acc + B.l2 * xt[j]) * B.inv_n;
        gt[j] = gj;
        gdp = fma(gj, dv[j], gdp);
        gmp = fmax(gmp, fabs(gj));
    }
    if (B.hi && tid == G - 1) {
        const double gj = (part[1] + (B.reg_bias ? B.l2 * xt[0] : 0.0)) * B.inv_n;
        gt[0] = gj;
        gdp = fma(gj, dv[0], gdp);
        gmp = fmax(gmp, fabs(gj));
    }
    group_sum_max<G>(gdp, gmp, red, flip);
    gd = gdp;
    gmax = gmp;
    if (C > 1) cluster.sync();   // nobody rewrites its part while another CTA may still be reading it
}

// SIMPLE variance for the same entities: var_j = 1 / (sum_i x_ij^2 rho_i (1 - rho_i) w_i + l2 [j regularised] + 1e-12)
template <int G>
__device__ __forceinline__ void variance_simple_big(const ReArgs &a, const BigEntity &B, const double *x,
                                                    double *var_out, double *red, int &flip)
{
    constexpr uint32_t W = G / 32;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *gw = B.gw + (size_t)warp * B.d;
    for (uint32_t j = lane; j < B.d; j += 32) gw[j] = 0.0;
    __syncwarp();
    const double b0 = B.hi ? x[0] : 0.0;
    const double *xf = x + B.hi;
    double dsum[1] = {0.0};
    for (uint32_t i = warp; i < B.n; i += W) {
        const int64_t gi = B.r0 + i;
        const int64_t qs = a.b.rowptr[gi], qe = a.b.rowptr[gi + 1];
        double z = 0.0;
        for (int64_t k = qs + lane; k < qe; k += 32) z = fma((double)a.b.val[k], xf[a.b.col[k]], z);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(kFull, z, o);
        z = (z + b0) + (a.b.offset ? (double)a.b.offset[gi] : 0.0);
        const double rho = 1.0 / (1.0 + exp(-z));
        const double di = rho * (1.0 - rho) * (a.b.weight ? (double)a.b.weight[gi] : 1.0);
        for (int64_t k = qs + lane; k < qe; k += 32) {
            const double v = (double)a.b.val[k];
            atomicAdd(&gw[a.b.col[k]], v * v * di);
        }
        if (lane == 0) dsum[0] += di;
        __syncwarp();
    }
    group_sum<G, 1>(dsum, red, flip);
    if (G == 32) __syncwarp();
    for (uint32_t c = tid; c < B.d; c += G) {
        double h = 0.0;
        for (uint32_t w2 = 0; w2 < W; w2++) h += B.gw[(size_t)w2 * B.d + c];
        var_out[B.hi + c] = 1.0 / ((h + B.l2) + 1.0e-12);
    }
    if (B.hi && tid == 0) var_out[0] = 1.0 / ((dsum[0] + (B.reg_bias ? B.l2 : 0.0)) + 1.0e-12);
}

}  // namespace gdmix
