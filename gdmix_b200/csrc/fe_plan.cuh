// fe_plan.cuh -- set-up kernels of the fixed-effect objective: per-feature non-zero counts, column renumbering.
// (The reference has no counterpart: TF streams the rows in file order every evaluation,
// fixed_effect_lr_lbfgs_model.py:309-381.  Here the shard is laid out once per training run.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gdmix {

// counts[c] += number of non-zeros with column c.  Popular features own a large share of the non-zeros (one feature
// of a Zipf-distributed bag is in most rows), so lanes of a warp that hit the same column are counted once:
// __match_any_sync groups them and the group's lowest lane adds the group size.
__global__ void __launch_bounds__(256) fe_count_columns_kernel(const int32_t *col, const int64_t nnz, const int64_t D,
                                                               unsigned long long *counts, int32_t *bad)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t start = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    for (int64_t base = start - lane; base < nnz; base += nth) {
        const int64_t i = base + lane;
        const bool live = i < nnz;
        const int32_t c = live ? col[i] : -1;
        const bool ok = live && c >= 0 && (int64_t)c < D;
        if (live && !ok) *bad = 1;
        const unsigned act = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const unsigned peers = __match_any_sync(act, c);
            if ((int)(__ffs(peers) - 1) == lane) atomicAdd(&counts[c], (unsigned long long)__popc(peers));
        }
    }
}

// out[i] = map[in[i]]
__global__ void __launch_bounds__(256) remap_i32_kernel(const int32_t *in, const int32_t *map, const int64_t n,
                                                        int32_t *out)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) out[i] = map[in[i]];
}

}  // namespace gdmix
