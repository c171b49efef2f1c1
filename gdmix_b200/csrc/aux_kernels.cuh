// aux_kernels.cuh -- scoring and fixed-effect streaming kernels (single pass, HBM bound).
//
// Reference semantics (gdmix-trainer/src/gdmix/):
//   re_score_kernel     models/custom/scipy/job_consumers.py:138-152 (InferenceJobConsumer) +
//                       models/custom/binary_logistic_regression.py:241-262 (predict_proba, logits)
//   fe_loss_grad_kernel models/custom/fixed_effect_lr_lbfgs_model.py:309-381 (_train_model_fn, one worker's
//                       partial sum before the all-reduce; intercept LAST; no 1/n)
//   fe_score_kernel     fixed_effect_lr_lbfgs_model.py:214-270 (_scoring_fn logits)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gdmix_b200.h"
#include "re_common.cuh"

namespace gdmix {

// One thread per sample; the sample's entity is found by bisection of ent_rowptr (L2-resident), so an entity with
// a million samples is scored by as many threads as one with ten (a warp per entity left the few huge entities of a
// Zipf-distributed key as the kernel's tail).  theta is read through L1/L2, X once from HBM; per-sample arithmetic
// and its order are those of the per-entity walk.
__global__ void __launch_bounds__(256) re_score_kernel(const gdmix_re_batch b, const int hi, const double *theta,
                                                       const uint8_t *has_model, float *logit, float *logit_pc)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t first = b.ent_rowptr[0], last = first + b.n_rows;
    for (int64_t i = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < last; i += nth) {
        // e = the entity with ent_rowptr[e] <= i < ent_rowptr[e + 1]
        int64_t lo = 0, hi_e = b.n_entities;   // invariant: ent_rowptr[lo] <= i < ent_rowptr[hi_e]
        while (hi_e - lo > 1) {
            const int64_t mid = (lo + hi_e) >> 1;
            if (b.ent_rowptr[mid] <= i) lo = mid; else hi_e = mid;
        }
        const int64_t e = lo;
        const bool model = theta != nullptr && (has_model == nullptr || has_model[e] != 0);
        const double offs = b.offset ? (double)b.offset[i] : 0.0;
        double z;
        if (model) {
            const double *th = theta + b.theta_ptr[e];
            z = hi ? th[0] : 0.0;
            const int64_t qs = b.rowptr[i], qe = b.rowptr[i + 1];
            for (int64_t q = qs; q < qe; q++) z = fma((double)b.val[q], th[hi + b.col[q]], z);
            z = z + offs;
        } else {
            z = offs;
        }
        logit[i] = (float)z;
        logit_pc[i] = (float)(z - offs);
    }
}

// logistic_terms (re_common.cuh) next to the library functions it replaces in the fast kernel: out[6 i ..] =
// {t, log1p, inv} of logistic_terms, then exp(-|z|), log(1 + exp(-|z|)), 1 / (1 + exp(-|z|)) by the library
__global__ void __launch_bounds__(256) logistic_selftest_kernel(const double *z, const int64_t n, double *out)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) {
        double t, l, v;
        logistic_terms(z[i], t, l, v);
        const double ez = exp(-fabs(z[i]));
        out[6 * i] = t; out[6 * i + 1] = l; out[6 * i + 2] = v;
        out[6 * i + 3] = ez; out[6 * i + 4] = log(1.0 + ez); out[6 * i + 5] = 1.0 / (1.0 + ez);
    }
}

// 16-bit local column indices (as they crossed PCIe) -> the int32 the kernels index with
__global__ void __launch_bounds__(256) widen_u16_kernel(const uint16_t *in, int32_t *out, const int64_t n)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) out[i] = (int32_t)in[i];
}

// 16-bit row lengths (as they crossed PCIe) -> the u32 the scan kernels take
__global__ void __launch_bounds__(256) widen_len16_kernel(const uint16_t *in, uint32_t *out, const int64_t n)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) out[i] = (uint32_t)in[i];
}

// labels that crossed PCIe as bits: out[i] = bit (bit0 + i) of `bits`
__global__ void __launch_bounds__(256) label_from_bits_kernel(const uint8_t *bits, const int bit0, float *out, const int64_t n)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) {
        const int64_t b = (int64_t)bit0 + i;
        out[i] = (float)((bits[b >> 3] >> (b & 7)) & 1u);
    }
}

// 8-bit local column indices (entities with at most 256 local features): 16 per thread and trip
__global__ void __launch_bounds__(256) widen_u8_kernel(const uint8_t *in, int32_t *out, const int64_t n)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t n16 = n >> 4;   // in / out are 256-byte aligned (chunk_image)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += nth) {
        const uint4 v = ((const uint4 *)in)[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        int4 *o = (int4 *)out + 4 * i;
#pragma unroll
        for (int q = 0; q < 4; q++)
            o[q] = make_int4(w[q] & 0xffu, (w[q] >> 8) & 0xffu, (w[q] >> 16) & 0xffu, w[q] >> 24);
    }
    for (int64_t i = (n16 << 4) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) out[i] = (int32_t)in[i];
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Fixed-effect partial objective and gradient over this rank's rows.  fg[0] = value,
// fg[1 + j] = gradient of coefficient j (intercept at j = D).  Thread per row; the scatter
// into the (L2-resident, 8*(D+2) byte) accumulator uses fp64 RED atomics.
__global__ void __launch_bounds__(256) fe_loss_grad_kernel(const gdmix_fe_rows R, const gdmix_lr_opts o,
                                                           const double *x, double *fg)
{
    const int hi = o.has_intercept ? 1 : 0;
    const int64_t D = R.n_features;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    const double b0 = hi ? x[D] : 0.0;
    double value = 0.0, dz_sum = 0.0;
    for (int64_t i = tid; i < R.n_rows; i += nth) {
        const int64_t qs = R.rowptr[i], qe = R.rowptr[i + 1];
        double z = 0.0;
        for (int64_t q = qs; q < qe; q++) z = fma((double)R.val[q], x[R.col[q]], z);
        z += R.offset ? (double)R.offset[i] : 0.0;
        z += b0;
        const double yi = (double)R.label[i], wi = R.weight ? (double)R.weight[i] : 1.0;
        double dz;
        if (R.linear_regression) {
            const double e = yi - z;
            value = fma(wi * e, e, value);
            dz = -2.0 * wi * e;
        } else {
            const double ex = exp(-fabs(z));
            value = fma(wi, fmax(z, 0.0) - z * yi + log1p(ex), value);
            const double inv = 1.0 / (1.0 + ex);
            dz = wi * ((z >= 0.0 ? inv : ex * inv) - yi);
        }
        for (int64_t q = qs; q < qe; q++) atomicAdd(&fg[1 + R.col[q]], (double)R.val[q] * dz);
        dz_sum += dz;
    }
    // L2 term (the reference adds l2 * l2_loss(x_reg) / num_workers on every worker)
    const int64_t preg = (hi && !o.regularize_bias) ? D : D + hi;
    const double nw = (double)(R.num_workers > 0 ? R.num_workers : 1);
    for (int64_t j = tid; j < preg; j += nth) {
        const double xj = x[j];
        value = fma(0.5 * o.l2 / nw * xj, xj, value);
        atomicAdd(&fg[1 + j], o.l2 * xj / nw);
    }
    value = warp_sum(value);
    dz_sum = warp_sum(dz_sum);
    __shared__ double sv[8], sd[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sv[warp] = value; sd[warp] = dz_sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0, dsum = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { v += sv[w]; dsum += sd[w]; }
        atomicAdd(&fg[0], v);
        if (hi) atomicAdd(&fg[1 + D], dsum);
    }
}

// Hessian of the fixed-effect logistic loss over this rank's rows, X1^T diag(w rho (1 - rho)) X1 with the
// intercept column LAST (fixed_effect_lr_lbfgs_model.py:271-296).  full == 0: h[P] gets the diagonal;
// full != 0: h[P*P] row-major gets the whole matrix (small P only: every row adds nnz^2 entries).
__global__ void __launch_bounds__(256) fe_hessian_kernel(const gdmix_fe_rows R, const int hi, const double *x,
                                                         const int full, double *h)
{
    const int64_t D = R.n_features, P = D + hi;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    const double b0 = hi ? x[D] : 0.0;
    double dsum = 0.0;
    for (int64_t i = tid; i < R.n_rows; i += nth) {
        const int64_t qs = R.rowptr[i], qe = R.rowptr[i + 1];
        double z = 0.0;
        for (int64_t q = qs; q < qe; q++) z = fma((double)R.val[q], x[R.col[q]], z);
        z += b0;
        z += R.offset ? (double)R.offset[i] : 0.0;
        const double rho = 1.0 / (1.0 + exp(-z));
        const double di = rho * (1.0 - rho) * (R.weight ? (double)R.weight[i] : 1.0);
        if (!full) {
            for (int64_t q = qs; q < qe; q++) {
                const double v = (double)R.val[q];
                atomicAdd(&h[R.col[q]], v * v * di);
            }
            dsum += di;
        } else {
            for (int64_t q = qs; q < qe; q++) {
                const double vd = (double)R.val[q] * di;
                const int64_t r = R.col[q];
                for (int64_t q2 = qs; q2 < qe; q2++) atomicAdd(&h[r * P + R.col[q2]], vd * (double)R.val[q2]);
                if (hi) { atomicAdd(&h[r * P + D], vd); atomicAdd(&h[D * P + r], vd); }
            }
            if (hi) atomicAdd(&h[D * P + D], di);
        }
    }
    if (!full && hi) {
        dsum = warp_sum(dsum);
        if ((threadIdx.x & 31) == 0) atomicAdd(&h[D], dsum);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Fixed-effect scoring pass (gdmix_fe_score): the staged row walk.  (The objective / gradient over a planned shard
// lives in fe_tile.cuh.)
// ---------------------------------------------------------------------------------------------------------
constexpr int kFeRowsThreads = 512;         // fe_rows_kernel CTA: sixteen warps, one CTA per SM
constexpr uint32_t kFeStageCap = 1024;      // non-zeros one warp stages at a time (4 KB values + 4 KB columns)
constexpr uint32_t kFeHeadMax = 8192;       // leading coefficients of x kept in shared memory (64 KB)

__host__ __device__ inline uint32_t fe_rows_smem_bytes(const uint32_t head)
{
    return 8u * head + (kFeRowsThreads / 32) * kFeStageCap * 8u;
}

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// The row's z goes out as the two fp32 logits of _scoring_fn (fixed_effect_lr_lbfgs_model.py:214-270).
__global__ void __launch_bounds__(kFeRowsThreads, 1) fe_score_rows_kernel(const gdmix_fe_rows R, const gdmix_lr_opts o,
                                                                          const double *x, const uint32_t head,
                                                                          float *logit, float *logit_pc)
{
    // A warp owns 32 consecutive rows at a time.  Their non-zeros are one contiguous range of the CSR arrays:
    // the warp copies it into shared memory asynchronously (cp.async, 16 bytes per lane and instruction when
    // the range is aligned: the whole 8 KB is in flight at once and no register waits for it), then every LANE
    // walks its own row out of shared memory -- one FMA per lane and instruction, no cross-lane reduction (a
    // warp-per-row walk spends ~40 instructions per FMA on butterflies for 32-wide rows).  Each lane starts its
    // walk `lane` elements into its row and wraps around, which spreads equal-length rows over all banks.
    // The x gather is what bounds this kernel: 32 lanes x 32 different 128-byte lines cost the L1 one cycle or
    // two per line.  So the first `head` coefficients of x live in shared memory (the host side numbers features
    // by falling frequency, so these are the hot ones) and only the tail is gathered through L1 / L2.
    // Blocks with more non-zeros than the stage holds are taken in runs of as many rows as fit; a single row longer
    // than the stage is summed by the whole warp straight from global memory.  Every sum has a fixed order.
    extern __shared__ __align__(16) unsigned char fe_smem[];
    const int hi = o.has_intercept ? 1 : 0;
    const int64_t D = R.n_features;
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double *xs = (double *)fe_smem;
    float *sval = (float *)(fe_smem + 8u * head) + wib * (2u * kFeStageCap);
    int32_t *scol = (int32_t *)(sval + kFeStageCap);
    for (uint32_t j = threadIdx.x; j < head; j += kFeRowsThreads) xs[j] = x[j];
    __syncthreads();
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double b0 = hi ? x[D] : 0.0;
    const int64_t nblocks = (R.n_rows + 31) >> 5;
    // row pointers of the warp's NEXT block are fetched while the current one is summed
    int64_t nqs = 0, nqe = 0;
    if (warp < nblocks && (warp << 5) + lane < R.n_rows) { nqs = R.rowptr[(warp << 5) + lane]; nqe = R.rowptr[(warp << 5) + lane + 1]; }
    for (int64_t blk = warp; blk < nblocks; blk += nwarps) {
        const int64_t base = blk << 5;
        const uint32_t nrows = (uint32_t)min((int64_t)32, R.n_rows - base);
        const int64_t qs = nqs, qe = nqe;
        {
            const int64_t nb = (blk + nwarps) << 5;
            nqs = 0; nqe = 0;
            if (blk + nwarps < nblocks && nb + lane < R.n_rows) { nqs = R.rowptr[nb + lane]; nqe = R.rowptr[nb + lane + 1]; }
        }
        double myz = 0.0;
        uint32_t done = 0;
        while (done < nrows) {
            const int64_t q0 = __shfl_sync(0xffffffffu, qs, done);
            const unsigned fit = __ballot_sync(0xffffffffu, lane >= done && lane < nrows && (qe - q0) <= (int64_t)kFeStageCap);
            const uint32_t nfit = __popc(fit);   // row ends ascend: the rows that fit are done .. done + nfit - 1
            if (nfit == 0) {
                const int64_t e0 = __shfl_sync(0xffffffffu, qe, done);
                double z = 0.0;
                for (int64_t k = q0 + lane; k < e0; k += 32) z = fma((double)R.val[k], __ldg(x + R.col[k]), z);
                z = warp_sum(z);
                if (lane == done) myz = z;
                done += 1;
                continue;
            }
            const uint32_t cnt = (uint32_t)(__shfl_sync(0xffffffffu, qe, done + nfit - 1) - q0);
            {
                const float *gv = R.val + q0;
                const int32_t *gc = R.col + q0;
                if ((((uintptr_t)gv | (uintptr_t)gc) & 15u) == 0) {   // both ranges start 16-byte aligned
                    const uint32_t n4 = cnt >> 2;
                    for (uint32_t f = lane; f < n4; f += 32) {
                        cp_async16(sval + 4 * f, gv + 4 * f);
                        cp_async16(scol + 4 * f, gc + 4 * f);
                    }
                    for (uint32_t f = (n4 << 2) + lane; f < cnt; f += 32) { cp_async4(sval + f, gv + f); cp_async4(scol + f, gc + f); }
                } else {
                    for (uint32_t f = lane; f < cnt; f += 32) { cp_async4(sval + f, gv + f); cp_async4(scol + f, gc + f); }
                }
                cp_async_wait_all();
            }
            __syncwarp();
            if (lane >= done && lane < done + nfit) {
                const uint32_t s0 = (uint32_t)(qs - q0), len = (uint32_t)(qe - qs);
                uint32_t pos = len ? lane % len : 0u;
                double z = 0.0;
#pragma unroll 4
                for (uint32_t s = 0; s < len; s++) {
                    const uint32_t c = (uint32_t)scol[s0 + pos];
                    const double xv = (c < head) ? xs[c] : __ldg(x + c);
                    z = fma((double)sval[s0 + pos], xv, z);
                    pos = (pos + 1u == len) ? 0u : pos + 1u;
                }
                myz = z;
            }
            __syncwarp();
            done += nfit;
        }
        const int64_t i = base + lane;
        if (i < R.n_rows) {
            const double z = myz + b0, offs = R.offset ? (double)R.offset[i] : 0.0;
            logit_pc[i] = (float)z;
            logit[i] = (float)(z + offs);
        }
    }
}

}  // namespace gdmix
