// re_small.cuh -- the random-effect solve for SMALL entities: one warp per entity.
//
// Same per-entity semantics as re_kernel.cuh / re_fast.cuh (BinaryLogisticRegressionTrainer.fit -> scipy
// fmin_l_bfgs_b, gdmix-trainer/src/gdmix/models/custom/binary_logistic_regression.py:84-131, :191-239;
// TrainingJobConsumer.__call__, models/custom/scipy/job_consumers.py:36-63; threshold_coefficients,
// util/model_utils.py:4-12), for the regime where an entity is a few hundred non-zeros (the per-user stage of
// BASELINE.json configs[3]: 32 samples x 64 features x 8 non-zeros).  There the CTA-per-entity kernels spend most
// of their time fetching instructions: the once-per-entity code (sliced-ELL sorts, staging, emit) is most of what such
// an entity executes, and a dozen one-warp CTAs per SM at different places of a 190 KB kernel do not fit the
// instruction caches (ncu: 4.4 no_instruction stalls per issue).  This kernel is a plain restatement sized to stay
// in them:
//   * a warp owns an entity; W warps per CTA, each with its own slice of shared memory; warps take entities from the
//     batch's work counter one by one, so differing iteration counts never idle a warp;
//   * the entity's CSR slice is read from HBM once (coalesced) into shared memory -- fp32 values, 8-bit local columns --
//     and transposed there into a CSC copy by a deterministic counting sort (rows in order, lanes ranked by
//     __match_any inside a row), so that g = X1^T r is a gather per feature with a fixed order (ascending rows, the
//     order the reference's COO accumulation has);
//   * coefficient vectors live in registers (feature j in lane j % 32, slot j / 32), the (S, Y) history in shared
//     memory; inner products are xor-butterflies, so every lane holds identical scalars and the scalar solver logic
//     (MINPACK-2 dcsrch, stop tests, skip / restart rules) runs replicated without divergence;
//   * the direction is the textbook two-loop recursion with H0 = I / theta.
// Entities that do not fit the slice (rows, non-zeros or coefficients above the launch's capacities) are appended to a
// list that the next kernel of the cascade drains.  fp64 throughout, no atomics on data, bitwise reproducible.
#pragma once
#include "linesearch.cuh"
#include "re_common.cuh"

namespace gdmix {

struct SmallShape {
    uint32_t cap_rows, cap_nnz, cap_coef;   // capacities of one warp's slice
    uint32_t m;                             // history pairs the slice holds
    // byte offsets inside a warp's slice
    uint32_t o_val, o_cval, o_y, o_w, o_off, o_r, o_xt, o_g, o_S, o_Y, o_rho, o_alpha, o_rowptr, o_colptr, o_cnt, o_col, o_crow;
    uint32_t bytes;                         // slice size (multiple of 16)
};

__host__ __device__ inline uint32_t small_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline SmallShape small_shape(uint32_t cap_rows, uint32_t cap_nnz, uint32_t cap_coef, uint32_t m)
{
    SmallShape s;
    s.cap_rows = cap_rows; s.cap_nnz = cap_nnz; s.cap_coef = cap_coef; s.m = m;
    uint32_t o = 0;
    s.o_r = o; o += 8 * cap_rows;
    s.o_xt = o; o += 8 * cap_coef;
    s.o_g = o; o += 8 * cap_coef;
    s.o_S = o; o += 8 * m * cap_coef;
    s.o_Y = o; o += 8 * m * cap_coef;
    s.o_rho = o; o += 8 * (m ? m : 1);
    s.o_alpha = o; o += 8 * (m ? m : 1);
    s.o_val = o; o += 4 * cap_nnz;
    s.o_cval = o; o += 4 * cap_nnz;
    s.o_y = o; o += 4 * cap_rows;
    s.o_w = o; o += 4 * cap_rows;
    s.o_off = o; o += 4 * cap_rows;
    s.o_cnt = o; o += 4 * cap_coef;          // u32 counters / cursors of the transposition
    s.o_rowptr = o; o += small_up(2 * (cap_rows + 1), 4);
    s.o_colptr = o; o += small_up(2 * (cap_coef + 1), 4);
    s.o_col = o; o += small_up(cap_nnz, 4);
    s.o_crow = o; o += small_up(cap_nnz, 4);
    s.bytes = small_up(o, 16);
    return s;
}

struct SmallArgs {
    ReArgs a;
    SmallShape S;
    int32_t warps;      // per CTA
};

__device__ __forceinline__ double small_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double small_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// the line search as ONE copy of code (the other kernels inline it)
__device__ __noinline__ int small_dcsrch(double &stp, double f, double g, int task, LineSearch &S)
{
    return dcsrch(stp, f, g, 1e-3, 0.9, 0.1, 0.0, 1e10, task, S);
}

// one entity's staged block as the evaluation sees it
struct SmallView {
    const float *val, *cval, *ys, *ws, *offs;
    double *r_s, *xt, *gs;
    const uint16_t *rowptr, *colptr;
    const uint8_t *col, *crow;
    uint32_t n, p, hi;
    int32_t regularize_bias;
    double l2;
};

// Objective and gradient at the point in V.xt -> f (returned, the same in every lane) and V.gs.  One copy of this code
// (it holds the fp64 exp / log1p / division sequences): the kernel's point is a small instruction footprint.
__device__ __noinline__ double small_evaluate(const SmallView &V, const uint32_t lane)
{
    const uint32_t n = V.n, p = V.p, hi = V.hi;
    const double inv_n = 1.0 / (double)n;
    double sq = 0.0;
    for (uint32_t j = lane; j < p; j += 32)
        if (!(hi && !V.regularize_bias && j == 0)) sq = fma(V.xt[j], V.xt[j], sq);
    const double b0 = hi ? V.xt[0] : 0.0;
    double cost = 0.0, rs = 0.0;
    for (uint32_t i = lane; i < n; i += 32) {
        double z = b0;
        const uint32_t b = V.rowptr[i], en = V.rowptr[i + 1];
        for (uint32_t k = b; k < en; k++) z = fma((double)V.val[k], V.xt[hi + V.col[k]], z);
        z += (double)V.offs[i];
        const double yi = (double)V.ys[i], wi = (double)V.ws[i];
        const double ex = exp(-fabs(z));
        cost = fma(wi, fmax(z, 0.0) - z * yi + log1p(ex), cost);
        const double inv = 1.0 / (1.0 + ex);
        const double ri = wi * ((z >= 0.0 ? inv : ex * inv) - yi);
        V.r_s[i] = ri;
        rs += ri;
    }
    cost = small_sum(cost);
    rs = small_sum(rs);
    sq = small_sum(sq);
    __syncwarp();
    for (uint32_t j = lane; j < p; j += 32) {
        double g;
        if (hi && j == 0) {
            g = rs;
            if (V.regularize_bias) g += V.l2 * V.xt[j];
        } else {
            const uint32_t c = j - hi;
            g = 0.0;
            for (uint32_t k = V.colptr[c]; k < V.colptr[c + 1]; k++) g = fma((double)V.cval[k], V.r_s[V.crow[k]], g);
            g += V.l2 * V.xt[j];
        }
        V.gs[j] = g * inv_n;
    }
    __syncwarp();
    return inv_n * (cost + 0.5 * V.l2 * sq);
}

// SL = coefficient slots per lane (p <= 32 * SL)
template <int SL>
__global__ void __launch_bounds__(256) re_small_kernel(const SmallArgs sa)
{
    extern __shared__ __align__(16) unsigned char smem_all[];
    const ReArgs &a = sa.a;
    const SmallShape &L = sa.S;
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned char *sm = smem_all + (size_t)wib * L.bytes;
    double *r_s = (double *)(sm + L.o_r), *xt = (double *)(sm + L.o_xt), *gs = (double *)(sm + L.o_g);
    double *Sh = (double *)(sm + L.o_S), *Yh = (double *)(sm + L.o_Y);
    double *rho = (double *)(sm + L.o_rho), *alpha = (double *)(sm + L.o_alpha);
    float *val = (float *)(sm + L.o_val), *cval = (float *)(sm + L.o_cval);
    float *ys = (float *)(sm + L.o_y), *ws = (float *)(sm + L.o_w), *offs = (float *)(sm + L.o_off);
    uint32_t *cnt = (uint32_t *)(sm + L.o_cnt);
    uint16_t *rowptr = (uint16_t *)(sm + L.o_rowptr), *colptr = (uint16_t *)(sm + L.o_colptr);
    uint8_t *col = (uint8_t *)(sm + L.o_col), *crow = (uint8_t *)(sm + L.o_crow);

    const uint32_t hi = a.o.has_intercept ? 1u : 0u;
    const int m = a.o.m;
    const double epsmch = 2.220446049250313e-16;

    for (;;) {
        __syncwarp();
        int64_t e = 0;
        if (lane == 0) e = atomicAdd(a.queue, 1);
        e = __shfl_sync(kFull, e, 0);
        if (a.todo) {
            if (e >= (int64_t)*a.todo_count) break;
            e = a.todo[e];
        }
        if (e >= a.b.n_entities) break;
        const int64_t r0 = a.b.ent_rowptr[e], r1 = a.b.ent_rowptr[e + 1];
        const int64_t q0 = a.b.rowptr[r0], q1 = a.b.rowptr[r1];
        const int64_t t0 = a.b.theta_ptr[e];
        const int64_t n64 = r1 - r0, nnz64 = q1 - q0, p64 = a.b.theta_ptr[e + 1] - t0;
        if (!(n64 >= 1 && p64 >= 1 && p64 >= (int64_t)hi && nnz64 >= 0)) {
            if (lane == 0 && a.status) a.status[e] = GDMIX_ERR_INVALID;
            continue;
        }
        if (n64 > (int64_t)L.cap_rows || nnz64 > (int64_t)L.cap_nnz || p64 > (int64_t)L.cap_coef || p64 > 32 * SL ||
            p64 - hi > 256 || n64 > 256) {
            if (lane == 0) a.defer_list[atomicAdd(a.defer_count, 1)] = (int32_t)e;   // the next kernel of the cascade takes it
            continue;
        }
        const uint32_t n = (uint32_t)n64, nnz = (uint32_t)nnz64, p = (uint32_t)p64, d = p - hi;

        // ---- stage: CSR slice, per-sample columns ----------------------------------------------------------------
        bool bad = false;
        for (uint32_t k = lane; k < nnz; k += 32) {
            const int32_t c = a.b.col[q0 + k];
            bad |= c < 0 || (uint32_t)c >= d;
            col[k] = (uint8_t)c;
            val[k] = a.b.val[q0 + k];
        }
        for (uint32_t i = lane; i <= n; i += 32) rowptr[i] = (uint16_t)(a.b.rowptr[r0 + i] - q0);
        for (uint32_t i = lane; i < n; i += 32) {
            ys[i] = a.b.label[r0 + i];
            ws[i] = a.b.weight ? a.b.weight[r0 + i] : 1.0f;
            offs[i] = a.b.offset ? a.b.offset[r0 + i] : 0.0f;
        }
        for (uint32_t j = lane; j < d; j += 32) cnt[j] = 0u;
        if (__any_sync(kFull, bad)) {
            if (lane == 0 && a.status) a.status[e] = GDMIX_ERR_INVALID;
            continue;
        }
        __syncwarp();
        // ---- transpose (CSC): rows in order; inside a row, lanes that hit the same column are ranked by lane ---------
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t b = rowptr[i], en = rowptr[i + 1];
            for (uint32_t k0 = b; k0 < en; k0 += 32) {
                const uint32_t k = k0 + lane;
                const bool on = k < en;
                const uint32_t c = on ? col[k] : 0xffffffffu - lane;
                const unsigned peers = __match_any_sync(kFull, c);
                if (on && (peers & ((1u << lane) - 1u)) == 0u) cnt[c] += __popc(peers);
                __syncwarp();
            }
        }
        {
            // exclusive scan of the column counts -> colptr; cnt becomes the write cursor
            uint32_t carry = 0;
            for (uint32_t j0 = 0; j0 < d; j0 += 32) {
                const uint32_t j = j0 + lane;
                const uint32_t mine = j < d ? cnt[j] : 0u;
                uint32_t incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t u = __shfl_up_sync(kFull, incl, o);
                    if (lane >= (uint32_t)o) incl += u;
                }
                if (j < d) { colptr[j] = (uint16_t)(carry + incl - mine); cnt[j] = carry + incl - mine; }
                carry += __shfl_sync(kFull, incl, 31);
            }
            if (lane == 0) colptr[d] = (uint16_t)carry;
        }
        __syncwarp();
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t b = rowptr[i], en = rowptr[i + 1];
            for (uint32_t k0 = b; k0 < en; k0 += 32) {
                const uint32_t k = k0 + lane;
                const bool on = k < en;
                const uint32_t c = on ? col[k] : 0xffffffffu - lane;
                const unsigned peers = __match_any_sync(kFull, c);
                const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
                if (on) {
                    const uint32_t pos = cnt[c] + rank;
                    cval[pos] = val[k];
                    crow[pos] = (uint8_t)i;
                }
                __syncwarp();
                if (on && rank == 0) cnt[c] += __popc(peers);
                __syncwarp();
            }
        }
        __syncwarp();

        const SmallView V{val, cval, ys, ws, offs, r_s, xt, gs, rowptr, colptr, col, crow, n, p, hi, a.o.regularize_bias, a.o.l2};
        // x (registers) -> shared memory, the evaluation, g back into registers
        auto evaluate = [&](const double (&xv)[SL], double (&gv)[SL]) -> double {
#pragma unroll
            for (int s = 0; s < SL; s++) { const uint32_t j = lane + 32u * s; if (j < p) xt[j] = xv[s]; }
            __syncwarp();
            const double fv = small_evaluate(V, lane);
#pragma unroll
            for (int s = 0; s < SL; s++) { const uint32_t j = lane + 32u * s; gv[s] = j < p ? gs[j] : 0.0; }
            return fv;
        };
        auto dot = [&](const double (&u)[SL], const double (&v)[SL]) -> double {
            double s0 = 0.0;
#pragma unroll
            for (int s = 0; s < SL; s++) s0 = fma(u[s], v[s], s0);   // slots past p hold zeros
            return small_sum(s0);
        };
        auto gnorm = [&](const double (&v)[SL]) -> double {
            double s0 = 0.0;
#pragma unroll
            for (int s = 0; s < SL; s++) s0 = fmax(s0, fabs(v[s]));
            return small_max(s0);
        };

        // ---- L-BFGS-B without bounds, as scipy.optimize.fmin_l_bfgs_b runs it ----
        double x[SL], g[SL], dd[SL], t[SL], rr_[SL], q[SL];
#pragma unroll
        for (int s = 0; s < SL; s++) {
            const uint32_t j = lane + 32u * s;
            x[s] = (j < p && a.theta_in) ? a.theta_in[t0 + j] : 0.0;
            g[s] = dd[s] = t[s] = rr_[s] = q[s] = 0.0;
        }
        int colh = 0, head = 0, iter = 0, nfev = 1, status = 0;
        double theta = 1.0;
        double f = evaluate(x, g);
        bool done = gnorm(g) <= a.o.pgtol;
        while (!done) {
            // direction: two-loop recursion, H0 = I / theta
#pragma unroll
            for (int s = 0; s < SL; s++) q[s] = g[s];
            for (int k = colh - 1; k >= 0; k--) {
                const int sl = (head + k) % m;
                double sv[SL];
#pragma unroll
                for (int s = 0; s < SL; s++) { const uint32_t j = lane + 32u * s; sv[s] = j < p ? Sh[(size_t)sl * p + j] : 0.0; }
                const double al = rho[sl] * dot(sv, q);
                if (lane == 0) alpha[sl] = al;
#pragma unroll
                for (int s = 0; s < SL; s++) { const uint32_t j = lane + 32u * s; if (j < p) q[s] -= al * Yh[(size_t)sl * p + j]; }
            }
            __syncwarp();
#pragma unroll
            for (int s = 0; s < SL; s++) q[s] = q[s] / theta;
            for (int k = 0; k < colh; k++) {
                const int sl = (head + k) % m;
                double yv[SL];
#pragma unroll
                for (int s = 0; s < SL; s++) { const uint32_t j = lane + 32u * s; yv[s] = j < p ? Yh[(size_t)sl * p + j] : 0.0; }
                const double beta = rho[sl] * dot(yv, q);
                const double cf = alpha[sl] - beta;
#pragma unroll
                for (int s = 0; s < SL; s++) { const uint32_t j = lane + 32u * s; if (j < p) q[s] += Sh[(size_t)sl * p + j] * cf; }
            }
#pragma unroll
            for (int s = 0; s < SL; s++) { dd[s] = -q[s]; t[s] = x[s]; rr_[s] = g[s]; }
            // line search (lnsrlb + dcsrch)
            const double dnorm = sqrt(dot(dd, dd));
            double stp = (iter == 0) ? fmin(1.0 / dnorm, 1e10) : 1.0;
            const double fold = f;
            double gd = 0.0, gdold = 0.0;
            int ifun = 0, iback = 0, info = 0, lstask = LS_START;
            LineSearch ls;
            for (;;) {
                gd = dot(g, dd);
                if (ifun == 0) {
                    gdold = gd;
                    if (gd >= 0.0) { info = -4; break; }
                }
                lstask = small_dcsrch(stp, f, gd, lstask, ls);
                if (lstask == LS_CONV || lstask == LS_WARN) break;
                if (lstask == LS_ERROR) { info = -4; break; }
                ifun++; iback = ifun - 1;
                if (iback >= a.o.max_ls) break;
#pragma unroll
                for (int s = 0; s < SL; s++) x[s] = stp * dd[s] + t[s];
                f = evaluate(x, g);
                nfev++;
            }
            if (info != 0 || iback >= a.o.max_ls) {
                // restore the previous iterate; restart from steepest descent, or give up when the memory is empty
#pragma unroll
                for (int s = 0; s < SL; s++) { x[s] = t[s]; g[s] = rr_[s]; }
                f = fold;
                if (colh == 0) { status = 2; iter++; break; }
                colh = 0; head = 0; theta = 1.0;
                continue;
            }
            iter++;
            if (iter >= a.o.max_iter || nfev > a.o.max_fun) { status = 1; break; }
            if (gnorm(g) <= a.o.pgtol) break;
            if ((fold - f) <= epsmch * a.o.factr * max3(fabs(fold), fabs(f), 1.0)) break;
            // curvature pair: s = stp * d, y = g - g_old
            double rr = 0.0;
#pragma unroll
            for (int s = 0; s < SL; s++) { rr_[s] = g[s] - rr_[s]; rr = fma(rr_[s], rr_[s], rr); }
            rr = small_sum(rr);
            double dr, ddum;
            if (stp == 1.0) { dr = gd - gdold; ddum = -gdold; }
            else {
                dr = (gd - gdold) * stp; ddum = -gdold * stp;
#pragma unroll
                for (int s = 0; s < SL; s++) dd[s] *= stp;
            }
            if (dr <= epsmch * ddum) continue;   // skip the update
            if (m > 0) {
                int slot;
                if (colh < m) { slot = (head + colh) % m; colh++; }
                else { slot = head; head = (head + 1) % m; }
#pragma unroll
                for (int s = 0; s < SL; s++) {
                    const uint32_t j = lane + 32u * s;
                    if (j < p) { Sh[(size_t)slot * p + j] = dd[s]; Yh[(size_t)slot * p + j] = rr_[s]; }
                }
                if (lane == 0) rho[slot] = 1.0 / dr;
                theta = rr / dr;
                __syncwarp();
            }
        }
        // ---- emit ------------------------------------------------------------------------------------------------------
        const double thr = a.o.sparsity_threshold;
#pragma unroll
        for (int s = 0; s < SL; s++) {
            const uint32_t j = lane + 32u * s;
            if (j < p) a.theta_out[t0 + j] = (thr > 0.0 && fabs(x[s]) <= thr) ? 0.0 : x[s];
        }
        if (lane == 0) {
            if (a.f_out) a.f_out[e] = f;
            if (a.nit) a.nit[e] = iter;
            if (a.nfev) a.nfev[e] = nfev;
            if (a.status) a.status[e] = status;
        }
    }
}

}  // namespace gdmix
