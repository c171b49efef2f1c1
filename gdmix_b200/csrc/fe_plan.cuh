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


// =====================================================================================================================
// Plan of the tiled objective (fe_tile.cuh): built once per training run from the shard's CSR (columns already in
// falling-frequency rank order), entirely on the device.
//
//   z side (row-major, for z = X x):  every row's non-zeros split, order kept, into a HOT part (rank < hz: the
//     coefficient sits in shared memory; fp32 value + 16-bit rank = 6 B) and a COLD part (value + 32-bit rank).
//     Both parts are packed per block of 32 rows, each block padded to a multiple of 8 (hot) / 4 (cold) entries, so
//     that a warp stages its block with 16-byte copies.  A row's hot part is padded with zero entries to whole QUADS
//     (a lane fetches four values with one 16-byte and four ranks with one 8-byte shared-memory load):
//     zh_len[row] = its quads, zc_len[row] = its cold entries.
//   g side (column-major, for g = X^T dz):  non-zeros of rank < hg sorted by (tile of `tile_rows` rows, column, row)
//     as fp32 value + 16-bit row inside the tile + 16-bit rank (8 B), each tile padded to a multiple of 256 entries
//     with zeros that extend its last run; the rest sorted by (L2 tile of `l2_tile_rows` rows, column, row) as value +
//     32-bit row, with a dense table of run starts per (L2 tile, cold column).
// One stable radix sort of all non-zeros by [class | tile | column] produces both g-side orders.
// =====================================================================================================================

__global__ void __launch_bounds__(256) fe_expand_rows_kernel(const int64_t *rowptr, const int64_t n_rows, uint32_t *row_of)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += nth) {
        const int64_t b = rowptr[i], e = rowptr[i + 1];
        for (int64_t q = b; q < e; q++) row_of[q - rowptr[0]] = (uint32_t)i;
    }
}

// hot / cold length of every row; err |= 1 when a row has more than 65535 hot or cold entries
__global__ void __launch_bounds__(256) fe_zcount_kernel(const int64_t *rowptr, const int32_t *col, const int64_t n_rows,
                                                        const int32_t hz, uint16_t *zh_len, uint16_t *zc_len, int32_t *err)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += nth) {
        const int64_t b = rowptr[i], e = rowptr[i + 1];
        uint32_t h = 0, c = 0;
        for (int64_t q = b; q < e; q++) {
            if (col[q] < hz) h++; else c++;
        }
        const uint32_t hq = (h + 3u) >> 2;        // the hot part is stored in quads (padded with zero entries)
        if (hq > 65535u || c > 65535u) { *err = 1; c = min(c, 65535u); }
        zh_len[i] = (uint16_t)min(hq, 65535u);
        zc_len[i] = (uint16_t)c;
    }
}

// entries of every 32-row block (len[row] * unit each), rounded up to a multiple of `align` (a power of two)
__global__ void __launch_bounds__(256) fe_zblock_kernel(const uint16_t *len, const int64_t n_rows, const int64_t nblocks,
                                                        const uint32_t unit, const uint32_t align, uint32_t *blk_cnt)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nblocks; b += nth) {
        uint32_t s = 0;
        for (int k = 0; k < 32; k++) {
            const int64_t i = b * 32 + k;
            if (i < n_rows) s += unit * (uint32_t)len[i];
        }
        blk_cnt[b] = (s + align - 1u) & ~(align - 1u);
    }
}

// a warp per 32-row block, a lane per row: the row's entries go to the hot / cold arrays in their original order
__global__ void __launch_bounds__(256) fe_zscatter_kernel(const int64_t *rowptr, const int32_t *col, const float *val,
                                                          const int64_t n_rows, const int32_t hz, const uint16_t *zh_len,
                                                          const uint16_t *zc_len, const int64_t *zh_blk,
                                                          const int64_t *zc_blk, float *zh_val, uint16_t *zh_col,
                                                          float *zc_val, int32_t *zc_col)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nblocks = (n_rows + 31) >> 5;
    for (int64_t blk = warp; blk < nblocks; blk += nwarps) {
        const int64_t i = blk * 32 + lane;
        const uint32_t len = i < n_rows ? 4u * (uint32_t)zh_len[i] : 0u, clen = i < n_rows ? zc_len[i] : 0u;
        uint32_t incl = len, cincl = clen;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o), cu = __shfl_up_sync(0xffffffffu, cincl, o);
            if (lane >= o) { incl += u; cincl += cu; }
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31), ctotal = __shfl_sync(0xffffffffu, cincl, 31);
        const int64_t q0 = zh_blk[blk], cq0 = zc_blk[blk];
        const uint32_t padded = (uint32_t)(zh_blk[blk + 1] - q0), cpadded = (uint32_t)(zc_blk[blk + 1] - cq0);
        if (i < n_rows) {
            int64_t h = q0 + (incl - len), c = cq0 + (cincl - clen);
            for (int64_t q = rowptr[i]; q < rowptr[i + 1]; q++) {
                const int32_t cc = col[q];
                if (cc < hz) { zh_val[h] = val[q]; zh_col[h] = (uint16_t)cc; h++; }
                else { zc_val[c] = val[q]; zc_col[c] = cc; c++; }
            }
            for (const int64_t hend = q0 + incl; h < hend; h++) { zh_val[h] = 0.0f; zh_col[h] = 0; }   // the row's last quad
        }
        if (total + lane < padded) { zh_val[q0 + total + lane] = 0.0f; zh_col[q0 + total + lane] = 0; }
        if (ctotal + lane < cpadded) { zc_val[cq0 + ctotal + lane] = 0.0f; zc_col[cq0 + ctotal + lane] = 0; }
    }
}

// sort key of every non-zero: hot = tile * hg + rank; cold = classbit | (l2 tile * n_cold + rank - hg)
__global__ void __launch_bounds__(256) fe_gkeys_kernel(const int32_t *col, const uint32_t *row_of, const int64_t nnz,
                                                       const int32_t hg, const int32_t tile_rows, const int64_t l2_tile_rows,
                                                       const int64_t n_cold, const uint64_t classbit, uint64_t *keys)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += nth) {
        const int32_t c = col[q];
        const uint64_t r = row_of[q];
        keys[q] = c < hg ? (r / (uint32_t)tile_rows) * (uint64_t)hg + (uint64_t)c
                         : classbit | ((r / (uint64_t)l2_tile_rows) * (uint64_t)n_cold + (uint64_t)(c - hg));
    }
}

// out[i] = first index whose key is >= q0 + i * stride  (i < count)
__global__ void __launch_bounds__(256) fe_lower_bound_kernel(const uint64_t *keys, const int64_t n, const uint64_t q0,
                                                             const uint64_t stride, const int64_t count, const int64_t minus,
                                                             int64_t *out)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += nth) {
        const uint64_t want = q0 + (uint64_t)i * stride;
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (keys[mid] < want) lo = mid + 1; else hi = mid;
        }
        out[i] = lo - minus;
    }
}

// entries of every g tile rounded up to a multiple of 256
__global__ void __launch_bounds__(256) fe_gtile_count_kernel(const int64_t *tile_begin, const int64_t n_tiles, uint32_t *cnt)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles; t += nth)
        cnt[t] = ((uint32_t)(tile_begin[t + 1] - tile_begin[t]) + 255u) & ~255u;
}

__global__ void __launch_bounds__(256) fe_ghot_gather_kernel(const uint64_t *keys, const uint32_t *perm, const int64_t n_hot,
                                                             const int32_t hg, const int32_t tile_rows,
                                                             const int64_t *tile_begin, const int64_t *tile_off,
                                                             const float *val, const uint32_t *row_of, float *gh_val,
                                                             uint16_t *gh_row, uint16_t *gh_col)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_hot; i += nth) {
        const uint64_t key = keys[i];
        const uint64_t t = key / (uint64_t)hg;
        const uint32_t c = (uint32_t)(key - t * (uint64_t)hg);
        const uint32_t src = perm[i];
        const int64_t dst = tile_off[t] + (i - tile_begin[t]);
        gh_val[dst] = val[src];
        gh_row[dst] = (uint16_t)(row_of[src] - (uint32_t)t * (uint32_t)tile_rows);
        gh_col[dst] = (uint16_t)c;
    }
}

// padding of every tile: zero values that extend the tile's last run (they add exact zeros to it)
__global__ void __launch_bounds__(256) fe_ghot_pad_kernel(const int64_t n_tiles, const int64_t *tile_begin,
                                                          const int64_t *tile_off, float *gh_val, uint16_t *gh_row,
                                                          uint16_t *gh_col)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = warp; t < n_tiles; t += nwarps) {
        const int64_t cnt = tile_begin[t + 1] - tile_begin[t];
        if (cnt == 0) continue;
        const int64_t b = tile_off[t] + cnt, e = tile_off[t + 1];
        const uint16_t last = gh_col[b - 1];
        for (int64_t q = b + lane; q < e; q += 32) { gh_val[q] = 0.0f; gh_row[q] = 0; gh_col[q] = last; }
    }
}

__global__ void __launch_bounds__(256) fe_gcold_gather_kernel(const uint32_t *perm, const int64_t n_cold, const float *val,
                                                              const uint32_t *row_of, float *gc_val, uint32_t *gc_row)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cold; i += nth) {
        const uint32_t src = perm[i];
        gc_val[i] = val[src];
        gc_row[i] = row_of[src];
    }
}

}  // namespace gdmix
