// re_variance.cuh -- FULL coefficient variance of the random-effect models: diag((X1^T D X1 + (l2 + 1e-12) I
// - l2 e0 e0^T [intercept unregularised])^-1), D = diag(w rho (1 - rho)), at the un-thresholded optimum
// (binary_logistic_regression.py:144-189, mode FULL).  Runs after the solver kernels: one CTA per entity builds
// the dense p x p Hessian (in shared memory when it fits, else in an L2-resident slice of the workspace),
// inverts it in place by Gauss-Jordan elimination (the matrix is symmetric positive definite: no pivoting) and
// emits the diagonal.  Rows are folded in one after the other with a barrier in between, so every entry is summed
// in row order: bitwise reproducible.  The solver kernels leave theta un-thresholded in this mode; the threshold
// (util/model_utils.py:4-12) is applied here once the entity's variance is out.
#pragma once
#include "re_common.cuh"

namespace gdmix {

struct VarArgs {
    gdmix_re_batch b;
    gdmix_lr_opts o;
    double *theta;           // in: un-thresholded optimum; out: thresholded
    double *var_out;
    const int32_t *status;   // entities the solver rejected are skipped
    int32_t *queue;
    double *scratch;         // per-CTA: p_max^2 doubles (when the matrix does not fit on chip)
    unsigned long long scratch_stride;  // in doubles
    uint32_t smem_matrix_doubles;       // capacity of the on-chip matrix
    uint32_t max_coef;
};

__global__ void __launch_bounds__(256) re_variance_full_kernel(const VarArgs a)
{
    extern __shared__ __align__(16) unsigned char vsm[];
    __shared__ int s_entity;
    const uint32_t tid = threadIdx.x, G = blockDim.x;
    const uint32_t hi = a.o.has_intercept ? 1u : 0u;
    double *rowk = (double *)vsm;                    // [max_coef]
    double *colk = rowk + a.max_coef;                // [max_coef]
    double *msm = colk + a.max_coef;                 // [smem_matrix_doubles]

    for (;;) {
        __syncthreads();
        if (tid == 0) s_entity = atomicAdd(a.queue, 1);
        __syncthreads();
        const int64_t e = s_entity;
        if (e >= a.b.n_entities) break;
        if (a.status && a.status[e] < 0) continue;
        const int64_t r0 = a.b.ent_rowptr[e], r1 = a.b.ent_rowptr[e + 1];
        const int64_t t0 = a.b.theta_ptr[e];
        const uint32_t n = (uint32_t)(r1 - r0), p = (uint32_t)(a.b.theta_ptr[e + 1] - t0);
        double *A = ((size_t)p * p <= a.smem_matrix_doubles)
                        ? msm : a.scratch + (unsigned long long)blockIdx.x * a.scratch_stride;
        const double *th = a.theta + t0;
        for (uint32_t k = tid; k < p * p; k += G) A[k] = 0.0;
        __syncthreads();
        // ---- H = sum_i d_i x_i x_i^T, one row at a time (x_i includes the leading 1 of the intercept) ----------
        for (uint32_t i = 0; i < n; i++) {
            const int64_t qs = a.b.rowptr[r0 + i], qe = a.b.rowptr[r0 + i + 1];
            const uint32_t len = (uint32_t)(qe - qs), m = len + hi;
            double z = hi ? th[0] : 0.0;
            for (int64_t q = qs; q < qe; q++) z = fma((double)a.b.val[q], th[hi + a.b.col[q]], z);
            z += a.b.offset ? (double)a.b.offset[r0 + i] : 0.0;
            const double rho = 1.0 / (1.0 + exp(-z));
            const double di = rho * (1.0 - rho) * (a.b.weight ? (double)a.b.weight[r0 + i] : 1.0);
            for (uint32_t pr = tid; pr < m * m; pr += G) {
                const uint32_t ia = pr / m, ib = pr - ia * m;
                const uint32_t ca = (hi && ia == 0) ? 0u : hi + (uint32_t)a.b.col[qs + ia - hi];
                const uint32_t cb = (hi && ib == 0) ? 0u : hi + (uint32_t)a.b.col[qs + ib - hi];
                const double va = (hi && ia == 0) ? 1.0 : (double)a.b.val[qs + ia - hi];
                const double vb = (hi && ib == 0) ? 1.0 : (double)a.b.val[qs + ib - hi];
                atomicAdd(&A[(size_t)ca * p + cb], di * va * vb);  // atomic only for a column repeated inside the row
            }
            __syncthreads();
        }
        for (uint32_t j = tid; j < p; j += G) {
            double add = a.o.l2 + 1.0e-12;
            if (hi && j == 0 && !a.o.regularize_bias) add -= a.o.l2;
            A[(size_t)j * p + j] += add;
        }
        __syncthreads();
        // ---- in-place Gauss-Jordan inversion -----------------------------------------------------------------
        for (uint32_t k = 0; k < p; k++) {
            const double inv = 1.0 / A[(size_t)k * p + k];
            for (uint32_t j = tid; j < p; j += G) {
                rowk[j] = A[(size_t)k * p + j] * inv;
                colk[j] = A[(size_t)j * p + k];
            }
            __syncthreads();
            for (uint32_t idx = tid; idx < p * p; idx += G) {
                const uint32_t i = idx / p, j = idx - i * p;
                double v;
                if (i == k) v = (j == k) ? inv : rowk[j];
                else if (j == k) v = -colk[i] * inv;
                else v = fma(-colk[i], rowk[j], A[idx]);
                A[idx] = v;
            }
            __syncthreads();
        }
        for (uint32_t j = tid; j < p; j += G) a.var_out[t0 + j] = A[(size_t)j * p + j];
        // the coefficients are thresholded only now (the variance is taken at the un-thresholded optimum)
        const double thr = a.o.sparsity_threshold;
        if (thr > 0.0)
            for (uint32_t j = tid; j < p; j += G) {
                const double xj = a.theta[t0 + j];
                if (fabs(xj) <= thr) a.theta[t0 + j] = 0.0;
            }
    }
}

}  // namespace gdmix
