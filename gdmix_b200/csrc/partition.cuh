// partition.cuh -- the data movement either side of the hot path, on the device (HBM-bound integer work):
//
//   * stable LSD radix sort of (u64 key, u32 payload) pairs, 8 bits per pass, 4096-key tiles;
//   * group-by-key on top of it: row permutation that brings equal keys together (original order kept inside
//     a group), segment pointers, the distinct keys -- what Spark's groupBy(entity) does in DataPartitioner
//     (gdmix-data/src/main/scala/com/linkedin/gdmix/data/DataPartitioner.scala:296-379) without leaving HBM;
//   * gathering a CSR by a row permutation (the regrouped sample block of a random-effect stage);
//   * entity -> partition map for integer ids (abs(String.valueOf(id).hashCode) % n, PartitionUtils.scala:31-37);
//   * area under the ROC curve with ties (Evaluator.scala:29-45 -> MLlib BinaryClassificationMetrics): sort by
//     score, then a rank-sum over tie groups.
// Every kernel reads and writes each element once per pass with coalesced accesses; nothing here is a GEMM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gdmix {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 keys per CTA
constexpr int kRadix = 256;

// ---- radix sort -----------------------------------------------------------------------------------------
// hist[tile][digit]
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint64_t *keys, const int64_t n, const int shift,
                                                                  uint32_t *hist)
{
    __shared__ uint32_t h[kRadix];
    const int tid = threadIdx.x;
    h[tid] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll 4
    for (int i = 0; i < kSortItems; i++) {
        const int64_t idx = base + (int64_t)i * kSortThreads + tid;
        if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(int64_t)blockIdx.x * kRadix + tid] = h[tid];
}

// One CTA per digit: exclusive prefix over tiles of hist[.][digit] (in place) and the digit's total.
__global__ void __launch_bounds__(256) radix_scan_tiles_kernel(uint32_t *hist, const int64_t ntiles, uint64_t *digit_total)
{
    __shared__ uint64_t wsum[8];
    __shared__ uint64_t carry_s;
    const int d = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int64_t t0 = 0; t0 < ntiles; t0 += 256) {
        const int64_t t = t0 + tid;
        const uint64_t mine = (t < ntiles) ? hist[t * kRadix + d] : 0;
        uint64_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint64_t before = carry_s;
        for (int w = 0; w < warp; w++) before += wsum[w];
        // 32-bit tile offsets would overflow beyond 4G keys of one digit; keep them relative to the digit's start
        if (t < ntiles) hist[t * kRadix + d] = (uint32_t)(before + incl - mine);
        __syncthreads();
        if (tid == 255) carry_s = before + incl;
        __syncthreads();
    }
    if (tid == 0) digit_total[d] = carry_s;
}

// exclusive scan of the 256 digit totals (one warp)
__global__ void radix_scan_digits_kernel(const uint64_t *digit_total, uint64_t *digit_base)
{
    const int lane = threadIdx.x;
    uint64_t carry = 0;
    for (int c = 0; c < kRadix; c += 32) {
        const uint64_t mine = digit_total[c + lane];
        uint64_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        digit_base[c + lane] = carry + incl - mine;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// Stable scatter of one tile.  A warp owns 512 consecutive keys and walks them 32 at a time, in order; equal
// digits inside a chunk are ranked by lane (match_any), chunks and warps by running counters.
__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const uint64_t *kin, const uint32_t *vin, uint64_t *kout,
                                                                     uint32_t *vout, const int64_t n, const int shift,
                                                                     const uint32_t *tile_offs, const uint64_t *digit_base)
{
    __shared__ uint64_t wbase[kSortThreads / 32][kRadix];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < (kSortThreads / 32) * kRadix; k += kSortThreads) (&wbase[0][0])[k] = 0;
    __syncthreads();
    const int64_t wbeg = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * (kSortTile / (kSortThreads / 32));
    constexpr int kChunks = kSortTile / (kSortThreads / 32) / 32;  // 16
    for (int c = 0; c < kChunks; c++) {
        const int64_t idx = wbeg + c * 32 + lane;
        const unsigned d = (idx < n) ? (unsigned)((kin[idx] >> shift) & 255u) : (256u + lane);
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (idx < n && (peers & ((1u << lane) - 1u)) == 0) wbase[warp][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {
        // thread = digit: turn per-warp counts into per-warp global start positions
        uint64_t run = digit_base[tid] + tile_offs[(int64_t)blockIdx.x * kRadix + tid];
        for (int w = 0; w < kSortThreads / 32; w++) {
            const uint64_t cnt = wbase[w][tid];
            wbase[w][tid] = run;
            run += cnt;
        }
    }
    __syncthreads();
    for (int c = 0; c < kChunks; c++) {
        const int64_t idx = wbeg + c * 32 + lane;
        const bool act = idx < n;
        const uint64_t key = act ? kin[idx] : 0;
        const unsigned d = act ? (unsigned)((key >> shift) & 255u) : (256u + lane);
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned rank = __popc(peers & ((1u << lane) - 1u));
        if (act) {
            const uint64_t pos = wbase[warp][d] + rank;
            kout[pos] = key;
            vout[pos] = vin ? vin[idx] : (uint32_t)idx;
        }
        __syncwarp();
        if (act && rank == 0) wbase[warp][d] += __popc(peers);
        __syncwarp();
    }
}

// ---- segments of a sorted key array ---------------------------------------------------------------------
// tile_count[tile] = number of group heads in the tile
__global__ void __launch_bounds__(kSortThreads) heads_count_kernel(const uint64_t *keys, const int64_t n, uint32_t *tile_count)
{
    __shared__ uint32_t ws[kSortThreads / 32];
    const int tid = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
    uint32_t c = 0;
    for (int i = 0; i < kSortItems; i++) {
        const int64_t idx = base + (int64_t)i * kSortThreads + tid;
        if (idx < n && (idx == 0 || keys[idx] != keys[idx - 1])) c++;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((tid & 31) == 0) ws[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kSortThreads / 32; w++) t += ws[w];
        tile_count[blockIdx.x] = t;
    }
}

// single CTA: exclusive scan of per-tile counts into 64-bit offsets; total to *out_total
__global__ void __launch_bounds__(256) tiles_exclusive_scan_kernel(const uint32_t *tile_count, const int64_t ntiles,
                                                                   int64_t *tile_off, int64_t *out_total)
{
    __shared__ uint64_t wsum[8];
    __shared__ uint64_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int64_t t0 = 0; t0 < ntiles; t0 += 256) {
        const int64_t t = t0 + tid;
        const uint64_t mine = (t < ntiles) ? tile_count[t] : 0;
        uint64_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint64_t before = carry_s;
        for (int w = 0; w < warp; w++) before += wsum[w];
        if (t < ntiles) tile_off[t] = (int64_t)(before + incl - mine);
        __syncthreads();
        if (tid == 255) carry_s = before + incl;
        __syncthreads();
    }
    if (tid == 0) *out_total = (int64_t)carry_s;
}

// heads -> seg_ptr[g] = index of the g-th group's first element, seg_key[g] = its key; seg_ptr[G] = n
__global__ void __launch_bounds__(kSortThreads) heads_write_kernel(const uint64_t *keys, const int64_t n, const int64_t *tile_off,
                                                                   int64_t *seg_ptr, uint64_t *seg_key)
{
    __shared__ uint32_t ws[kSortThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // a thread owns 16 CONSECUTIVE keys here so that heads keep their order
    const int64_t base = (int64_t)blockIdx.x * kSortTile + (int64_t)tid * kSortItems;
    uint32_t flags = 0, c = 0;
    for (int i = 0; i < kSortItems; i++) {
        const int64_t idx = base + i;
        if (idx < n && (idx == 0 || keys[idx] != keys[idx - 1])) { flags |= 1u << i; c++; }
    }
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    int64_t g = tile_off[blockIdx.x] + (incl - c);
    for (int w = 0; w < warp; w++) g += ws[w];
    for (int i = 0; i < kSortItems; i++) {
        if ((flags >> i) & 1u) {
            seg_ptr[g] = base + i;
            if (seg_key) seg_key[g] = keys[base + i];
            g++;
        }
    }
    if (blockIdx.x == gridDim.x - 1 && tid == kSortThreads - 1) {
        // the very last thread knows the total
        seg_ptr[g] = n;
    }
}

// ---- regrouping a CSR by a row permutation ----------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_row_len_kernel(const int64_t *rowptr, const uint32_t *perm, const int64_t n,
                                                             uint32_t *len_out)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) {
        const uint32_t r = perm[i];
        len_out[i] = (uint32_t)(rowptr[r + 1] - rowptr[r]);
    }
}

// 64-bit exclusive scan of u32 lengths: tile sums, then tiles_exclusive_scan_kernel, then this
__global__ void __launch_bounds__(kSortThreads) tile_sum_kernel(const uint32_t *len, const int64_t n, uint32_t *tile_sum)
{
    __shared__ uint32_t ws[kSortThreads / 32];
    const int tid = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
    uint32_t c = 0;
    for (int i = 0; i < kSortItems; i++) {
        const int64_t idx = base + (int64_t)i * kSortThreads + tid;
        if (idx < n) c += len[idx];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((tid & 31) == 0) ws[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kSortThreads / 32; w++) t += ws[w];
        tile_sum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kSortThreads) rowptr_from_len_kernel(const uint32_t *len, const int64_t n,
                                                                       const int64_t *tile_off, int64_t *rowptr_out)
{
    __shared__ uint32_t ws[kSortThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t base = (int64_t)blockIdx.x * kSortTile + (int64_t)tid * kSortItems;
    uint32_t l[kSortItems], c = 0;
    for (int i = 0; i < kSortItems; i++) { l[i] = (base + i < n) ? len[base + i] : 0u; c += l[i]; }
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    int64_t run = tile_off[blockIdx.x] + (incl - c);
    for (int w = 0; w < warp; w++) run += ws[w];
    for (int i = 0; i < kSortItems; i++) {
        if (base + i < n) rowptr_out[base + i] = run;
        run += l[i];
    }
    if (base <= n - 1 && n - 1 < base + kSortItems) rowptr_out[n] = run;
}

// a warp copies one row's non-zeros (coalesced on both sides)
__global__ void __launch_bounds__(256) gather_rows_kernel(const int64_t *rowptr_in, const int32_t *col_in, const float *val_in,
                                                          const uint32_t *perm, const int64_t n, const int64_t *rowptr_out,
                                                          int32_t *col_out, float *val_out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nwarps) {
        const int64_t src = rowptr_in[perm[i]], len = rowptr_in[perm[i] + 1] - src, dst = rowptr_out[i];
        for (int64_t k = lane; k < len; k += 32) {
            col_out[dst + k] = col_in[src + k];
            val_out[dst + k] = val_in[src + k];
        }
    }
}

__global__ void __launch_bounds__(256) gather_f32_kernel(const float *in, const uint32_t *perm, const int64_t n, float *out)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) out[i] = in[perm[i]];
}

// ---- entity -> partition for integer ids -------------------------------------------------------------------
__global__ void __launch_bounds__(256) partition_i64_kernel(const int64_t *ids, const int64_t n, const int32_t num_partitions,
                                                            int32_t *partition_out)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) {
        // String.valueOf(long) then String.hashCode: h = 31 h + char, most significant digit first
        const int64_t v = ids[i];
        uint64_t mag = v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v;
        char digits[20];
        int nd = 0;
        do { digits[nd++] = (char)('0' + (int)(mag % 10)); mag /= 10; } while (mag);
        uint32_t h = 0;
        if (v < 0) h = (uint32_t)'-';
        for (int k = nd - 1; k >= 0; k--) h = 31u * h + (uint32_t)digits[k];
        const int32_t hs = (int32_t)h;
        const int32_t a = (hs == INT32_MIN) ? hs : (hs < 0 ? -hs : hs);
        partition_out[i] = a % num_partitions;
    }
}

// ---- entity-local feature indexing by presence bitmaps ------------------------------------------------------------
// What prepare_jobs derives per entity with np.unique(cols, return_inverse=True) (job_consumers.py:243): the sorted
// distinct global feature ids of an entity and, for every non-zero, the rank of its feature among them.  With a
// feature bag of at most a few thousand ids an entity's set is a bitmap of W32 = ceil(D / 32) words:
//   mark    thread per sample: entity by bisection of ent_rowptr, atomicOr of the sample's feature bits
//   count   thread per (entity, word): popcount -> prefix inside the entity (a lane walks its entity's words)
//   index   thread per sample again: local = prefix[word] + popc(bits below), written over the global id
//   list    thread per entity: its set bits in ascending order -> uniq_global[uniq_ptr[e] ..]
// Integer work only: the result is the one the (entity, feature) pair sort gives, bit for bit.
__device__ __forceinline__ int64_t entity_of_row(const int64_t *ent_rowptr, const int64_t n_entities, const int64_t i)
{
    int64_t lo = 0, hi = n_entities;   // ent_rowptr[lo] <= i < ent_rowptr[hi]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (ent_rowptr[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) bitmap_mark_kernel(const int64_t *ent_rowptr, const int64_t n_entities, const int64_t *rowptr,
                                                          const int32_t *gcol, const int64_t n_rows, const int32_t D, const int32_t W32,
                                                          uint32_t *bitmap, unsigned *bad)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += nth) {
        const int64_t e = entity_of_row(ent_rowptr, n_entities, i);
        uint32_t *bm = bitmap + e * W32;
        for (int64_t q = rowptr[i]; q < rowptr[i + 1]; q++) {
            const uint32_t c = (uint32_t)gcol[q];
            if (c >= (uint32_t)D) { *bad = 1u; continue; }
            // look first: an entity with millions of samples would otherwise hammer two words with atomics
            // (a stale read only costs a redundant atomic)
            const uint32_t bit = 1u << (c & 31u);
            if (!(__ldcg(&bm[c >> 5]) & bit)) atomicOr(&bm[c >> 5], bit);
        }
    }
}

__global__ void __launch_bounds__(256) bitmap_count_kernel(const uint32_t *bitmap, const int64_t n_entities, const int32_t W32,
                                                           uint32_t *word_prefix, int64_t *d_e)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_entities; e += nth) {
        uint32_t run = 0;
        for (int32_t w = 0; w < W32; w++) {
            word_prefix[e * W32 + w] = run;
            run += __popc(bitmap[e * W32 + w]);
        }
        d_e[e] = run;
    }
}

__global__ void __launch_bounds__(256) bitmap_index_kernel(const int64_t *ent_rowptr, const int64_t n_entities, const int64_t *rowptr,
                                                           const int32_t *gcol, const int64_t n_rows, const int32_t W32,
                                                           const uint32_t *bitmap, const uint32_t *word_prefix, int32_t *local_col)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += nth) {
        const int64_t e = entity_of_row(ent_rowptr, n_entities, i);
        const uint32_t *bm = bitmap + e * W32, *wp = word_prefix + e * W32;
        for (int64_t q = rowptr[i]; q < rowptr[i + 1]; q++) {
            const uint32_t c = (uint32_t)gcol[q], w = c >> 5;
            local_col[q] = (int32_t)(wp[w] + __popc(bm[w] & ((1u << (c & 31u)) - 1u)));
        }
    }
}

__global__ void __launch_bounds__(256) bitmap_list_kernel(const uint32_t *bitmap, const int64_t n_entities, const int32_t W32,
                                                          const int64_t *uniq_ptr, int64_t *uniq_global)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_entities; e += nth) {
        int64_t at = uniq_ptr[e];
        for (int32_t w = 0; w < W32; w++) {
            uint32_t bits = bitmap[e * W32 + w];
            while (bits) {
                const int b = __ffs(bits) - 1;
                uniq_global[at++] = (int64_t)w * 32 + b;
                bits &= bits - 1u;
            }
        }
    }
}

// ---- AUC ------------------------------------------------------------------------------------------------
// scores -> sortable keys (ascending), payload = 1 for a NEGATIVE (label <= 0), 0 for a positive
__global__ void __launch_bounds__(256) auc_keys_kernel(const float *score, const float *label, const int64_t n, uint64_t *keys,
                                                       uint32_t *payload)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) {
        float s = score[i];
        if (s == 0.0f) s = 0.0f;  // -0.0 and +0.0 tie
        uint32_t b = __float_as_uint(s);
        b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
        keys[i] = b;
        payload[i] = label[i] > 0.0f ? 0u : 1u;
    }
}

// Over the tie groups of the sorted scores: 2U = sum_g pos_g * (2 neg_below_g + neg_g) (Evaluator.scala:29-45 counts
// a tie as half a concordant pair).  With negx[i] = negatives among the first i sorted elements (exclusive scan,
// negx[n] = all negatives) a group [b, e) has neg_below = negx[b], neg_g = negx[e] - negx[b], so its term is
// pos_g * (negx[b] + negx[e]) -- one thread per group, no serial walk.  Per-thread sums in double (integers up to
// 2^53 are exact), CTA partials in a fixed order.
__global__ void __launch_bounds__(256) auc_groups_kernel(const int64_t *seg_ptr, const int64_t ngroups, const int64_t *negx,
                                                         double *cta_u2)
{
    __shared__ double sh[256];
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    double u2 = 0.0;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += nth) {
        const int64_t b = seg_ptr[g], e = seg_ptr[g + 1];
        const int64_t nb = negx[b], ne = negx[e];
        const int64_t pos = (e - b) - (ne - nb);
        u2 += (double)pos * (double)(nb + ne);
    }
    sh[threadIdx.x] = u2;
    __syncthreads();
    for (int s2 = 128; s2 > 0; s2 >>= 1) {
        if ((int)threadIdx.x < s2) sh[threadIdx.x] += sh[threadIdx.x + s2];
        __syncthreads();
    }
    if (threadIdx.x == 0) cta_u2[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(256) auc_finish_kernel(const double *cta_u2, const int n_cta, const int64_t *negx,
                                                         const int64_t n, double *out /* auc, positives, negatives */)
{
    __shared__ double sh[256];
    double u2 = 0.0;
    for (int i = threadIdx.x; i < n_cta; i += 256) u2 += cta_u2[i];
    sh[threadIdx.x] = u2;
    __syncthreads();
    for (int s2 = 128; s2 > 0; s2 >>= 1) {
        if ((int)threadIdx.x < s2) sh[threadIdx.x] += sh[threadIdx.x + s2];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double N = (double)negx[n], P = (double)(n - negx[n]);
        out[0] = (P > 0 && N > 0) ? 0.5 * sh[0] / (P * N) : 0.0;
        out[1] = P;
        out[2] = N;
    }
}

// ---- DataPartitioner's bounds and OffsetUpdater's join ------------------------------------------------------
// Group id of every row (DataPartitioner.getGroupId, gdmix-data/.../data/DataPartitioner.scala:335-379): rows are
// given grouped by entity (perm / seg_ptr of gdmix_group_by_key).  count = rows of the entity;
// groups = upper > 0 ? count / upper + 1 : 1;  id = pmod(uid, groups);  lower > 0 and count < lower: id = -1.
// 0 = active data, anything else passive.  group_id is written in INPUT row order.
__global__ void __launch_bounds__(256) group_ids_kernel(const int64_t *seg_ptr, const int64_t n_groups, const uint32_t *perm,
                                                        const int64_t *uid, const int64_t n, const int32_t lower,
                                                        const int32_t upper, int32_t *group_id)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += nth) {
        int64_t lo = 0, hi = n_groups;   // seg_ptr[lo] <= k < seg_ptr[hi]
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (seg_ptr[mid] <= k) lo = mid; else hi = mid;
        }
        const int64_t count = seg_ptr[lo + 1] - seg_ptr[lo];
        const uint32_t row = perm[k];
        int32_t id;
        if (lower > 0 && count < (int64_t)lower) {
            id = -1;
        } else {
            const int64_t groups = upper > 0 ? count / (int64_t)upper + 1 : 1;
            int64_t r = uid[row] % groups;      // pmod: the sign of the divisor
            if (r < 0) r += groups;
            id = (int32_t)r;
        }
        group_id[row] = id;
    }
}

// Offset of every data row from the previous coordinate's scores, joined by uid (OffsetUpdater.updateOffset,
// gdmix-data/.../data/OffsetUpdater.scala:105-129): offset = float(predictionScore) [- predictionScorePerCoordinate];
// matched[i] = 0 marks rows the inner join drops.  The scores' uids come sorted (gdmix_sort_pairs_u64 over the
// uids reinterpreted as u64) with the permutation that sorted them; of equal uids the first in file order wins.
__global__ void __launch_bounds__(256) offset_join_kernel(const int64_t *uid, const int64_t n, const uint64_t *skey,
                                                          const uint32_t *sperm, const int64_t m, const float *score,
                                                          const float *per_coordinate, float *offset_out, uint8_t *matched)
{
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) {
        const uint64_t want = (uint64_t)uid[i];
        int64_t lo = 0, hi = m;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (skey[mid] < want) lo = mid + 1; else hi = mid;
        }
        const bool hit = lo < m && skey[lo] == want;
        float o = 0.0f;
        if (hit) {
            const uint32_t j = sperm[lo];
            o = score[j];
            if (per_coordinate) o = o - per_coordinate[j];
        }
        offset_out[i] = o;
        matched[i] = hit ? 1 : 0;
    }
}

}  // namespace gdmix
