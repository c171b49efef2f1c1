// re_kernel.cuh -- the random-effect hot-path kernel on sm_100a.
//
// One CTA ("entity group", G = 32..256 threads) owns one entity at a time:
//   1. stage   (re_passes.cuh) the entity's CSR slice -- fp32 values, int32 local columns, per-sample
//              label/weight/offset -- is read from HBM exactly once and laid out in shared memory as
//              pair-packed CSR plus a CSC built on chip by a deterministic counting sort;
//   2. solve   L-BFGS-B as scipy.optimize.fmin_l_bfgs_b runs it without bounds: MINPACK-2 dcsrch line
//              search (linesearch.cuh), skip / restart rules, pgtol + factr + maxiter stop tests, compact
//              L-BFGS direction (re_lbfgs.cuh) -- entirely out of shared memory, fp64 throughout, no
//              atomics, fixed summation order (bitwise reproducible run to run);
//   3. emit    theta (optionally thresholded), f, nit, nfev, status, SIMPLE variance.
// CTAs are persistent and pull entities from a global atomic queue, so divergent iteration counts
// between entities never idle an SM.
//
// Reference semantics being replaced (gdmix-trainer/src/gdmix/):
//   models/custom/binary_logistic_regression.py:84-131 (_loss/_gradient), :191-239 (fit),
//   :144-189 (_compute_variance SIMPLE), models/custom/scipy/job_consumers.py:36-63,
//   util/model_utils.py:4-12.
#pragma once
#include "linesearch.cuh"
#include "re_common.cuh"
#include "re_lbfgs.cuh"
#include "re_passes.cuh"

namespace gdmix {

// BIG = false: the entity is staged in shared memory (re_passes.cuh, stage_entity).  BIG = true: the entity
// does not fit on chip; X stays in global memory and every evaluation sweeps it (evaluate_big), the solver
// vectors live in shared memory and the (S, Y) history in the per-CTA global arena.  Same solver either way.
template <int G, int MT, bool BIG = false>
__global__ void __launch_bounds__(G) re_solver_kernel(const ReArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double red[2 * kMaxWarps * kRedK];
    __shared__ int s_entity;
    __shared__ unsigned s_bad;
    __shared__ double cl_part[4];   // BIG, cluster launch: this CTA's loss / residual sums for the other CTAs to read

    // BIG only: launched as a thread-block cluster, all CTAs of a cluster solve ONE entity together -- the samples
    // are split over them in evaluate_big, everything else is replicated (see there).  C == 1 otherwise.
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t C = BIG ? cluster.num_blocks() : 1u, crank = BIG ? cluster.block_rank() : 0u;

    const uint32_t tid = threadIdx.x;
    const uint32_t hi = a.o.has_intercept ? 1u : 0u;
    const int m = a.o.m;
    int flip = 0;

    for (;;) {
        int64_t e;
        if (BIG && C > 1) {
            cluster.sync();   // previous entity done in every CTA of the cluster (and its index read by all)
            if (tid == 0) { if (crank == 0) s_entity = atomicAdd(a.queue, 1); s_bad = 0; }
            cluster.sync();
            e = *cluster.map_shared_rank(&s_entity, 0);
            if (a.todo ? (e >= (int64_t)*a.todo_count) : (e >= a.b.n_entities)) {
                cluster.sync();   // no CTA leaves while another may still be reading its shared memory
                break;
            }
        } else {
            group_sync<G>();  // previous entity fully emitted before its memory is reused
            if (tid == 0) { s_entity = atomicAdd(a.queue, 1); s_bad = 0; }
            group_sync<G>();
            e = s_entity;
        }
        if (a.todo) {
            if (e >= (int64_t)*a.todo_count) break;
            e = a.todo[e];
        }
        if (e >= a.b.n_entities) break;

        const int64_t r0 = a.b.ent_rowptr[e], r1 = a.b.ent_rowptr[e + 1];
        const int64_t q0 = a.b.rowptr[r0], q1 = a.b.rowptr[r1];
        const int64_t t0 = a.b.theta_ptr[e];
        const int64_t n64 = r1 - r0, nnz64 = q1 - q0, p64 = a.b.theta_ptr[e + 1] - t0;
        const uint32_t n = (uint32_t)n64, nnz = (uint32_t)nnz64, p = (uint32_t)p64, d = p - hi;

        const bool shape_ok = n64 >= 1 && p64 >= 1 && p64 >= (int64_t)hi && nnz64 >= 0 && n64 < (1ll << 31) &&
                              p64 < (1ll << 31);
        if (!shape_ok) {
            if (tid == 0 && a.status) a.status[e] = GDMIX_ERR_INVALID;
            continue;
        }
        double *xa, *xb, *ga, *gb, *dv, *dense, *part, *Sh, *Yh;
        double *rres = nullptr;
        float *sw = nullptr, *soff = nullptr, *csr_val = nullptr, *csc_val = nullptr;
        uint32_t *rowst = nullptr, *colst = nullptr;
        uint16_t *csr_col = nullptr, *csc_row = nullptr;
        Staged S;
        BigEntity B;
        if constexpr (!BIG) {
            const ReLayout L = re_layout(n, nnz, d, p, (uint32_t)m, (uint32_t)MT);
            const bool ok = n64 < 65535 && (p64 - hi) < 65535 && nnz64 < (1ll << 30) && L.fixed_bytes <= a.smem_bytes;
            if (!ok) {
                // too large to stage: hand it to the kernel that leaves X in global memory
                if (tid == 0) {
                    if (a.giant_list && n64 >= (int64_t)a.giant_rows) a.giant_list[atomicAdd(a.giant_count, 1)] = (int32_t)e;
                    else if (a.defer_list) a.defer_list[atomicAdd(a.defer_count, 1)] = (int32_t)e;
                    else if (a.status) a.status[e] = GDMIX_ERR_TOO_LARGE;
                }
                continue;
            }
            xa = (double *)(smem + L.xa); xb = (double *)(smem + L.xb);
            ga = (double *)(smem + L.ga); gb = (double *)(smem + L.gb);
            dv = (double *)(smem + L.dv);
            rres = (double *)(smem + L.r);
            float *sy = (float *)(smem + L.y);
            sw = (float *)(smem + L.w); soff = (float *)(smem + L.off);
            rowst = (uint32_t *)(smem + L.rowst); colst = (uint32_t *)(smem + L.colst);
            csr_val = (float *)(smem + L.csr_val); csc_val = (float *)(smem + L.csc_val);
            csr_col = (uint16_t *)(smem + L.csr_col); csc_row = (uint16_t *)(smem + L.csc_row);
            dense = (double *)(smem + L.dense); part = (double *)(smem + L.part);
            double *hist = (L.total_bytes <= a.smem_bytes && !a.hist_global)
                               ? (double *)(smem + L.hist)
                               : (double *)(a.arena + (unsigned long long)blockIdx.x * a.arena_stride);
            Sh = hist; Yh = hist + (size_t)m * p;

            // W*d staging counters alias the solver vectors, which are initialised afterwards
            if (!stage_entity<G>(a, r0, q0, n, d, sy, sw, soff, rowst, colst, csr_val, csr_col, csc_val, csc_row,
                                 (uint32_t *)xa, (uint32_t *)red, &s_bad)) {
                if (tid == 0 && a.status) a.status[e] = GDMIX_ERR_INVALID;
                continue;
            }
            S.n = n; S.d = d; S.p = p; S.nnz = nnz; S.hi = hi;
            S.trs = team_shift(G, n); S.tr = 1u << S.trs;
            S.tcs = team_shift(G, d); S.tc = 1u << S.tcs;
            S.y = sy; S.w = sw; S.off = soff; S.rowst = rowst; S.colst = colst;
            S.csr_val = (const float2 *)csr_val; S.csc_val = (const float2 *)csc_val;
            S.csr_col = (const ushort2 *)csr_col; S.csc_row = (const ushort2 *)csc_row;
            S.r = rres; S.inv_n = 1.0 / (double)n; S.l2 = a.o.l2; S.reg_bias = a.o.regularize_bias;
        } else {
            constexpr uint32_t W = G / 32;
            if (big_layout_bytes(p, d, W, (uint32_t)MT) > a.smem_bytes) {
                if (tid == 0 && a.status) a.status[e] = GDMIX_ERR_TOO_LARGE;  // thousands of local features AND unstageable
                continue;
            }
            uint32_t o = 0;
            xa = (double *)(smem + o); o += align16(8 * p);
            xb = (double *)(smem + o); o += align16(8 * p);
            ga = (double *)(smem + o); o += align16(8 * p);
            gb = (double *)(smem + o); o += align16(8 * p);
            dv = (double *)(smem + o); o += align16(8 * p);
            // lanes per row from the entity's mean row length; as many private gradient copies per warp as the
            // rows walked per step, if the shared memory given to the kernel holds them (else wider teams)
            uint32_t ts = 0;
            {
                const uint32_t avg = (uint32_t)((nnz64 + n64 - 1) / n64);
                while (ts < 5u && (1u << ts) < avg) ts++;
                while (ts < 5u && big_layout_bytes(p, d, W, (uint32_t)MT, 32u >> ts) > a.smem_bytes) ts++;
            }
            B.ts = ts;
            B.gw = (double *)(smem + o); o += align16(8 * W * (32u >> ts) * d);
            dense = (double *)(smem + o); o += align16(8 * dense_doubles((uint32_t)MT));
            part = (double *)(smem + o);
            double *hist = (double *)(a.arena + (unsigned long long)blockIdx.x * a.arena_stride);
            Sh = hist; Yh = hist + (size_t)m * p;
            B.r0 = r0; B.n = n; B.d = d; B.p = p; B.hi = hi;
            B.inv_n = 1.0 / (double)n; B.l2 = a.o.l2; B.reg_bias = a.o.regularize_bias;
        }
        auto eval = [&](const double *xt_, const double *dv_, double *gt_, double &f_, double &gd_, double &gm_) {
            if constexpr (BIG) evaluate_big<G>(a, B, xt_, dv_, gt_, red, flip, &s_bad, cl_part, f_, gd_, gm_);
            else evaluate<G>(S, xt_, dv_, gt_, red, flip, f_, gd_, gm_);
        };

        double *x = xa, *xt = xb, *g = ga, *gt = gb;
        for (uint32_t j = tid; j < p; j += G) {
            x[j] = a.theta_in ? a.theta_in[t0 + j] : 0.0;
            dv[j] = 0.0;
        }
        Lbfgs lb;
        lbfgs_reset<G, MT>(lb, dense);
        group_sync<G>();

        double f, gd, gmax;
        eval(x, dv, g, f, gd, gmax);
        int nfev = 1, iter = 0, status = GDMIX_SOLVE_CONVERGED;
        if constexpr (BIG) {
            if (s_bad) {  // a column index outside the entity's feature range (found by the first sweep)
                if (tid == 0 && a.status) a.status[e] = GDMIX_ERR_INVALID;
                continue;
            }
        }

        if (a.mode == kModeLossGrad) {
            if (crank == 0) {
                for (uint32_t j = tid; j < p; j += G) a.g_out[t0 + j] = g[j];
                if (tid == 0) a.f_out[e] = f;
            }
            continue;
        }

        // ---- L-BFGS-B, unbounded -----------------------------------------------------------------------
        const double epsmch = 2.220446049250313e-16;
        const double ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmx = 1e10;
        double dtd = 0.0;
        bool done = gmax <= a.o.pgtol;
        if (!done) {
            // steepest-descent start: dv = -g
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
                eval(xt, dv, gt, f, gd, gmax_t);
                nfev++;
            }
            if (info != 0 || iback >= a.o.max_ls) {
                f = fold;  // x, g still hold the previous iterate
                if (lb.col == 0) { status = GDMIX_SOLVE_ABNORMAL; iter++; break; }
                // refresh the memory and restart from steepest descent
                group_sync<G>();
                lbfgs_reset<G, MT>(lb, dense);
                double v1[1] = {0.0};
                for (uint32_t j = tid; j < p; j += G) {
                    const double gj = g[j];
                    dv[j] = -gj;
                    v1[0] = fma(gj, gj, v1[0]);
                }
                group_sum<G, 1>(v1, red, flip);
                dtd = v1[0];
                gd = -v1[0];
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

        // ---- emit ---------------------------------------------------------------------------------------
        if (crank != 0) continue;   // every CTA of a cluster holds the same answer: rank 0 writes it
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
        if constexpr (BIG) {
            if (a.var_out && a.o.variance_mode == GDMIX_VARIANCE_SIMPLE) {
                group_sync<G>();
                variance_simple_big<G>(a, B, x, a.var_out + t0, red, flip);
            }
        } else if (a.var_out && a.o.variance_mode == GDMIX_VARIANCE_SIMPLE) {
            // var_j = 1 / (sum_i x_ij^2 rho_i (1-rho_i) w_i + l2 [j regularised] + 1e-12)
            // (binary_logistic_regression.py:171-177), evaluated at the un-thresholded optimum
            const double b0 = hi ? x[0] : 0.0;
            for (uint32_t i = tid; i < n; i += G) {
                const uint32_t s = rowst[i], len = rowst[i + 1] - s;
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
                    const uint32_t c = j - hi, s = colst[c], len = colst[c + 1] - s;
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
