// fe_tile.cuh -- the fixed-effect objective / gradient over a planned shard (fe_plan.cuh), no atomics.
//
// Reference semantics: one worker's partial sum of _train_model_fn, gdmix-trainer/src/gdmix/models/custom/
// fixed_effect_lr_lbfgs_model.py:309-381 (intercept LAST, no 1/n, l2 / num_workers per worker).
//
//   fe_z_kernel      z = X x, loss, dz per row.  A warp owns 32 consecutive rows: their HOT entries (fp32 value + 16-bit
//                    rank, 6 B) are one contiguous, 16-byte aligned range that the warp copies to shared memory with
//                    cp.async; every lane then walks its own row against the first `hz` coefficients of x, which live
//                    in shared memory.  The COLD entries (rank >= hz) are read straight from global memory and their
//                    coefficients gathered through L2 while the hot copy is in flight.
//   fe_g_kernel      g = X^T dz for the ranks < hg.  A CTA owns a tile of `tile_rows` rows at a time: the tile's dz sits
//                    in shared memory next to the CTA's hg fp64 accumulators; the tile's entries, sorted by column, are
//                    streamed with 16-byte loads (8 per lane) and reduced by key: runs inside a lane, then a segmented
//                    scan over the lanes, a carry from step to step, and a head / tail slot pair per chunk of steps that
//                    warp 0 adds in chunk order after the tile.  Every run has one owner, so no atomics and a fixed order.
//   fe_gcold_kernel  the ranks >= hg: a warp per (L2 tile of rows, column) run, dz gathered through L2.
//   fe_finish2_kernel  per feature: the CTAs' accumulators (hot) or the L2 tiles' run sums (cold) in fixed order, + l2;
//                    CTA 0: objective value and the intercept's gradient from the z pass's per-CTA partials.
// Every sum has a fixed order: the objective is bitwise reproducible run to run.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "aux_kernels.cuh"

namespace gdmix {

struct FeTilePlan {
    int64_t n_rows, n_features;
    int32_t hz, hg, tile_rows, pad0;
    int64_t l2_tile_rows;
    // z side
    int64_t n_blocks;
    const int64_t *zh_blk;      // [n_blocks + 1] first hot entry of the 32-row block (multiple of 8)
    const uint16_t *zh_len;     // [n_rows] hot QUADS of the row (entries / 4, zero-padded)
    const float *zh_val;
    const uint16_t *zh_col;
    const int64_t *zc_blk;      // [n_blocks + 1] first cold entry of the 32-row block (multiple of 4)
    const uint16_t *zc_len;     // [n_rows]
    const float *zc_val;
    const int32_t *zc_col;
    // g side, ranks < hg
    int64_t n_tiles;
    const int64_t *gt_ptr;      // [n_tiles + 1] first entry of the tile (multiple of 256)
    const float *gh_val;
    const uint16_t *gh_row;
    const uint16_t *gh_col;
    // g side, ranks >= hg
    int64_t n_l2_tiles, n_cold;
    const int64_t *gc_run;      // [n_l2_tiles * n_cold + 1]
    const float *gc_val;
    const uint32_t *gc_row;
    // scratch
    double *dz;                 // [n_rows]
    double *block_part;         // [2 * z_grid]
    double *acc_part;           // [g_grid][hg]
    double *cold_part;          // [n_l2_tiles][n_cold]
    int32_t z_grid, g_grid;
};

constexpr int kFeZThreads = 512;
constexpr uint32_t kFeZStage = 1024;       // hot entries one warp stages at a time (4 KB values + 2 KB ranks)
constexpr uint32_t kFeZColdStage = 256;    // cold entries one warp stages at a time (1 KB values + 1 KB ranks)
constexpr int kFeZColdRegs = 8;            // cold coefficients a lane gathers ahead of its hot walk
constexpr int kFeGThreads = 512;
constexpr int kFeGStep = 256;              // entries per warp step (8 per lane)

// dz of a tile in shared memory: row r at r + r / 16.  A frequent feature's consecutive entries are consecutive rows, and
// a lane owns eight consecutive entries: without the skew the lanes of a half-warp read rows 8 apart -- two banks.
__host__ __device__ inline uint32_t fe_g_skew(const uint32_t r) { return r + (r >> 4); }
__host__ __device__ inline uint32_t fe_g_tile_doubles(const uint32_t tile_rows) { return (fe_g_skew(tile_rows) + 2u) & ~1u; }

__host__ __device__ inline uint32_t fe_z_smem_bytes(const uint32_t hz)
{
    return ((8u * hz + 15u) & ~15u) + (kFeZThreads / 32) * (kFeZStage * 6u + kFeZColdStage * 8u);
}
__host__ __device__ inline uint32_t fe_g_smem_bytes(const uint32_t hg, const uint32_t tile_rows)
{
    // accumulators + two dz tiles (the next one is copied under the current one's pass); a tile is stored with one
    // double of padding after every 16 (fe_g_skew), rounded to 16 bytes
    return ((8u * hg + 15u) & ~15u) + 2u * fe_g_tile_doubles(tile_rows) * 8u;
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}

// A tile's entries are cut into chunks of whole steps that the CTA's warps take from a counter (the column-sorted
// stream is cheap where one feature owns many entries and dear where every entry is its own run, so equal shares
// would leave most warps waiting).  A chunk is reduced left to right on its own; the run that opens it and the run
// it leaves open go to the chunk's slot pair (head, tail) instead of the accumulators.  After the tile warp 0 walks the
// slots in chunk order, 32 at a time: empty slots squeezed out, equal columns -- one run that spans chunks -- summed left
// to right by a segmented scan, one add per distinct column, the open run carried into the next 32.  Which warp reduced
// a chunk does not matter: every sum has the same order in every run of the program.
constexpr int kFeGMaxChunks = 64;

__device__ __forceinline__ void g_stitch(double *acc, int32_t *slot_col, double *slot_sum, const int n_slots,
                                         const uint32_t lane)
{
    int32_t ccol = -1;      // the run left open by the previous round
    double csum = 0.0;
    for (int base = 0; base < n_slots; base += 32) {
        const int idx = base + (int)lane;
        int32_t col = idx < n_slots ? slot_col[idx] : -1;
        double sum = idx < n_slots ? slot_sum[idx] : 0.0;
        const unsigned V = __ballot_sync(0xffffffffu, col >= 0);
        const int n = __popc(V);
        if (n == 0) continue;
        const int src = lane < (uint32_t)n ? (int)__fns(V, 0, lane + 1) : 0;   // lane j takes the j-th live slot
        col = __shfl_sync(0xffffffffu, col, src);
        sum = __shfl_sync(0xffffffffu, sum, src);
        if (lane >= (uint32_t)n) col = -2 - (int32_t)lane;                     // distinct dummies
        if (lane == 0) {
            if (col == ccol) sum = csum + sum;                 // the carried run goes on
            else if (ccol >= 0) acc[ccol] += csum;             // it ended with the previous round
        }
        const int32_t prev = __shfl_up_sync(0xffffffffu, col, 1);
        const unsigned F = __ballot_sync(0xffffffffu, lane == 0 || col != prev);
        double T = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double u = __shfl_up_sync(0xffffffffu, T, d);
            if (lane >= (uint32_t)d && ((F >> (lane - d + 1)) & ((1u << d) - 1u)) == 0u) T += u;
        }
        const bool last = lane + 1 < (uint32_t)n && ((F >> (lane + 1)) & 1u);   // a run that ends inside the round
        if (last) acc[col] += T;
        ccol = __shfl_sync(0xffffffffu, col, n - 1);
        csum = __shfl_sync(0xffffffffu, T, n - 1);
        __syncwarp();
    }
    if (lane == 0 && ccol >= 0) acc[ccol] += csum;
    __syncwarp();
}

struct ZMeta {
    int64_t q0, cq0;
    uint32_t padded, cpadded, len, clen;
    float y, w, off;
};

__device__ __forceinline__ void z_meta(ZMeta &m, const gdmix_fe_rows &R, const FeTilePlan &P, const int64_t blk,
                                       const uint32_t lane)
{
    const int64_t i = (blk << 5) + lane;
    const bool live = blk < P.n_blocks && i < R.n_rows;
    m.q0 = m.cq0 = 0; m.padded = m.cpadded = m.len = m.clen = 0; m.y = 0.0f; m.w = 1.0f; m.off = 0.0f;
    if (blk < P.n_blocks) {
        m.q0 = P.zh_blk[blk]; m.padded = (uint32_t)(P.zh_blk[blk + 1] - m.q0);
        m.cq0 = P.zc_blk[blk]; m.cpadded = (uint32_t)(P.zc_blk[blk + 1] - m.cq0);
    }
    if (live) {
        m.len = P.zh_len[i]; m.clen = P.zc_len[i];
        m.y = R.label[i];
        if (R.weight) m.w = R.weight[i];
        if (R.offset) m.off = R.offset[i];
    }
}

__global__ void __launch_bounds__(kFeZThreads, 1) fe_z_kernel(const gdmix_fe_rows R, const gdmix_lr_opts o,
                                                               const FeTilePlan P, const double *__restrict__ x)
{
    extern __shared__ __align__(16) unsigned char fe_smem[];
    __shared__ double sv[kFeZThreads / 32], sd[kFeZThreads / 32];
    const int hi = o.has_intercept ? 1 : 0;
    const int64_t D = R.n_features;
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double *xs = (double *)fe_smem;
    unsigned char *stage = fe_smem + ((8u * (uint32_t)P.hz + 15u) & ~15u) + wib * (kFeZStage * 6u + kFeZColdStage * 8u);
    float *sval = (float *)stage;
    uint16_t *scol = (uint16_t *)(sval + kFeZStage);
    float *cval = (float *)(stage + kFeZStage * 6u);
    int32_t *ccol = (int32_t *)(cval + kFeZColdStage);
    for (uint32_t j = threadIdx.x; j < (uint32_t)P.hz; j += kFeZThreads) xs[j] = x[j];
    __syncthreads();
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double b0 = hi ? x[D] : 0.0;
    double value = 0.0, dz_sum = 0.0;
    ZMeta m, mn;
    z_meta(m, R, P, warp, lane);
    for (int64_t blk = warp; blk < P.n_blocks; blk += nwarps) {
        const int64_t i = (blk << 5) + lane;
        const bool live = i < R.n_rows;
        // the block's hot and cold entries are two contiguous, 16-byte aligned ranges: one asynchronous copy each
        const bool staged = m.padded <= kFeZStage, cstaged = m.cpadded <= kFeZColdStage;
        if (staged) {
            const float *gv = P.zh_val + m.q0;
            const uint16_t *gc = P.zh_col + m.q0;
            for (uint32_t f = lane; f < (m.padded >> 2); f += 32) cp_async16(sval + 4 * f, gv + 4 * f);
            for (uint32_t f = lane; f < (m.padded >> 3); f += 32) cp_async16(scol + 8 * f, gc + 8 * f);
        }
        if (cstaged) {
            const float *gv = P.zc_val + m.cq0;
            const int32_t *gc = P.zc_col + m.cq0;
            for (uint32_t f = lane; f < (m.cpadded >> 2); f += 32) { cp_async16(cval + 4 * f, gv + 4 * f); cp_async16(ccol + 4 * f, gc + 4 * f); }
        }
        cp_async_commit();
        z_meta(mn, R, P, blk + nwarps, lane);      // the next block's row facts travel under this block's copy
        // this lane's first hot / cold entry inside the block
        uint32_t incl = 4u * m.len, cincl = m.clen;      // hot lengths are in quads
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, s), cu = __shfl_up_sync(0xffffffffu, cincl, s);
            if (lane >= (uint32_t)s) { incl += u; cincl += cu; }
        }
        const uint32_t nq = m.len, s0 = incl - 4u * nq, cs0 = cincl - m.clen, clen = m.clen;
        cp_async_wait_all();
        __syncwarp();
        // cold coefficients come through L2: the gathers are issued now and used after the hot walk
        double xc[kFeZColdRegs];
        float vc[kFeZColdRegs];
        if (cstaged) {
#pragma unroll
            for (int j = 0; j < kFeZColdRegs; j++) {
                const bool on = (uint32_t)j < clen;
                vc[j] = on ? cval[cs0 + j] : 0.0f;
                xc[j] = on ? __ldg(x + ccol[cs0 + j]) : 0.0;
            }
        }
        double z = 0.0, z1 = 0.0;
        // hot walk, a quad per step: four values by one 16-byte load, four ranks by one 8-byte load, four gathers of x, two
        // chains of fused multiply-adds.  The lane starts at the quad that falls on its own 16-byte bank group (when its
        // row reaches that far) and wraps around, so that lanes of rows of any length read different banks.
        uint32_t q = (lane - (s0 >> 2)) & 7u;
        if (q >= nq) q = 0u;
        if (staged) {
#pragma unroll 2
            for (uint32_t s = 0; s < nq; s++) {
                const float4 v = *(const float4 *)(sval + s0 + 4u * q);
                const uint2 c = *(const uint2 *)(scol + s0 + 4u * q);
                z = fma((double)v.x, xs[c.x & 0xffffu], z);
                z1 = fma((double)v.y, xs[c.x >> 16], z1);
                z = fma((double)v.z, xs[c.y & 0xffffu], z);
                z1 = fma((double)v.w, xs[c.y >> 16], z1);
                q = (q + 1u == nq) ? 0u : q + 1u;
            }
        } else {
            // a block with more hot entries than the stage holds (very long rows): straight from global memory, same order
            for (uint32_t s = 0; s < nq; s++) {
                const float4 v = *(const float4 *)(P.zh_val + m.q0 + s0 + 4u * q);
                const uint2 c = *(const uint2 *)(P.zh_col + m.q0 + s0 + 4u * q);
                z = fma((double)v.x, xs[c.x & 0xffffu], z);
                z1 = fma((double)v.y, xs[c.x >> 16], z1);
                z = fma((double)v.z, xs[c.y & 0xffffu], z);
                z1 = fma((double)v.w, xs[c.y >> 16], z1);
                q = (q + 1u == nq) ? 0u : q + 1u;
            }
        }
        z += z1;
        if (cstaged) {
#pragma unroll
            for (int j = 0; j < kFeZColdRegs; j++) z = fma((double)vc[j], xc[j], z);
            for (uint32_t j = kFeZColdRegs; j < clen; j++) z = fma((double)cval[cs0 + j], __ldg(x + ccol[cs0 + j]), z);
        } else {
            for (uint32_t j = 0; j < clen; j++) z = fma((double)P.zc_val[m.cq0 + cs0 + j], __ldg(x + P.zc_col[m.cq0 + cs0 + j]), z);
        }
        __syncwarp();
        if (live) {
            z += (double)m.off;
            z += b0;
            const double yi = (double)m.y, wi = (double)m.w;
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
            P.dz[i] = dz;
            dz_sum += dz;
        }
        m = mn;
    }
    value = warp_sum(value);
    dz_sum = warp_sum(dz_sum);
    if (lane == 0) { sv[wib] = value; sd[wib] = dz_sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0, dsum = 0.0;
        for (int w = 0; w < kFeZThreads / 32; w++) { v += sv[w]; dsum += sd[w]; }
        P.block_part[2 * blockIdx.x] = v;
        P.block_part[2 * blockIdx.x + 1] = dsum;
    }
}

struct GStep { float4 v0, v1; uint4 r, c; };

__device__ __forceinline__ void g_load(GStep &s, const FeTilePlan &P, const int64_t base)
{
    const float4 *pv = (const float4 *)(P.gh_val + base);
    s.v0 = __ldg(pv);
    s.v1 = __ldg(pv + 1);
    s.r = __ldg((const uint4 *)(P.gh_row + base));
    s.c = __ldg((const uint4 *)(P.gh_col + base));
}

// asynchronous copy of tile `tile`'s dz into `dst` (whole 16-byte pieces; the tail, if any, by plain loads)
__device__ __forceinline__ void g_copy_dz(double *dst, const FeTilePlan &P, const int64_t tile)
{
    if (tile >= P.n_tiles) return;
    const int64_t row0 = tile * (int64_t)P.tile_rows;
    const int32_t nr = (int32_t)min((int64_t)P.tile_rows, P.n_rows - row0);
    const double *src = P.dz + row0;
    // 8-byte asynchronous copies: the skewed destination breaks 16-byte pieces every 16 rows
    for (int32_t j = threadIdx.x; j < nr; j += kFeGThreads) cp_async8(dst + fe_g_skew((uint32_t)j), src + j);
}

__global__ void __launch_bounds__(kFeGThreads, 1) fe_g_kernel(const FeTilePlan P)
{
    extern __shared__ __align__(16) unsigned char fe_smem[];
    __shared__ double slot_sum[2 * kFeGMaxChunks];
    __shared__ int32_t slot_col[2 * kFeGMaxChunks];
    __shared__ int32_t next_chunk;
    double *acc = (double *)fe_smem;
    double *dz_buf = (double *)(fe_smem + ((8u * (uint32_t)P.hg + 15u) & ~15u));
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = kFeGThreads / 32;
    for (int32_t j = threadIdx.x; j < P.hg; j += kFeGThreads) acc[j] = 0.0;
    g_copy_dz(dz_buf, P, blockIdx.x);
    cp_async_commit();
    int64_t gb = 0, ge = 0;
    if ((int64_t)blockIdx.x < P.n_tiles) { gb = P.gt_ptr[blockIdx.x]; ge = P.gt_ptr[blockIdx.x + 1]; }
    uint32_t parity = 0;
    int prev_slots = 0;
    for (int64_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, parity ^= 1u) {
        const double *dzs = dz_buf + (size_t)parity * fe_g_tile_doubles((uint32_t)P.tile_rows);
        const int64_t b = gb;
        const int64_t nsteps = (ge - gb) / kFeGStep;
        const int64_t cs = (nsteps + kFeGMaxChunks - 1) / kFeGMaxChunks;       // steps per chunk (>= 1 when nsteps > 0)
        const int n_chunks = cs > 0 ? (int)((nsteps + cs - 1) / cs) : 0;
        // this warp's first chunk is chunk w; its first step's entries are global loads that need not wait for the barriers
        GStep cur, nxt;
        int chunk = (int)w;
        if (chunk < n_chunks) g_load(cur, P, b + (int64_t)chunk * cs * kFeGStep + lane * 8);
        const int64_t tn = tile + gridDim.x;
        if (tn < P.n_tiles) { gb = P.gt_ptr[tn]; ge = P.gt_ptr[tn + 1]; }
        cp_async_wait_all();
        __syncthreads();   // the previous tile's runs are all in acc or in the slots; this tile's dz is in place
        if (w == 0) {
            g_stitch(acc, slot_col, slot_sum, prev_slots, lane);
            if (lane == 0) next_chunk = NW;
        }
        g_copy_dz(dz_buf + (size_t)(parity ^ 1u) * fe_g_tile_doubles((uint32_t)P.tile_rows), P, tn);   // under this tile's pass
        cp_async_commit();
        __syncthreads();
        prev_slots = 2 * n_chunks;
        // the warp's steps as one stream: while a step is reduced the next one -- the chunk's next step, or the first step
        // of the chunk the warp takes next from the counter -- is already on its way
        int64_t s = (int64_t)chunk * cs, s_end = min(nsteps, s + cs);
        int32_t carry_col = -1, head_c = -1;
        double carry_sum = 0.0, head_s = 0.0;
        bool carry_head = true, first = true;
        while (chunk < n_chunks) {
            int next_chunk_id = chunk;
            int64_t ns = s + 1;
            if (ns >= s_end) {
                if (lane == 0) next_chunk_id = atomicAdd(&next_chunk, 1);
                next_chunk_id = __shfl_sync(0xffffffffu, next_chunk_id, 0);
                ns = (int64_t)next_chunk_id * cs;
            }
            if (next_chunk_id < n_chunks) g_load(nxt, P, b + ns * kFeGStep + lane * 8);
            {
                const float v[8] = {cur.v0.x, cur.v0.y, cur.v0.z, cur.v0.w, cur.v1.x, cur.v1.y, cur.v1.z, cur.v1.w};
                const uint32_t rw[4] = {cur.r.x, cur.r.y, cur.r.z, cur.r.w}, cw[4] = {cur.c.x, cur.c.y, cur.c.z, cur.c.w};
                // the eight products first (independent shared-memory gathers), then the runs
                double p[8];
#pragma unroll
                for (int j = 0; j < 8; j++) p[j] = (double)v[j] * dzs[fe_g_skew((rw[j >> 1] >> ((j & 1) * 16)) & 0xffffu)];
                // runs inside the lane: the first one may continue the previous lane's, the last one stays open
                const int32_t first_col = (int32_t)(cw[0] & 0xffffu);
                int32_t run_col = first_col;
                double run = p[0];
                double head = 0.0;
                int k = 0;
#pragma unroll
                for (int j = 1; j < 8; j++) {
                    const int32_t cj = (int32_t)((cw[j >> 1] >> ((j & 1) * 16)) & 0xffffu);
                    if (cj == run_col) {
                        run += p[j];
                    } else {
                        if (k == 0) head = run; else acc[run_col] += run;   // a run closed on both sides inside the lane
                        k++;
                        run_col = cj;
                        run = p[j];
                    }
                }
                if (first) carry_col = __shfl_sync(0xffffffffu, first_col, 0);   // an empty run opens the chunk
                first = false;
                int32_t prev_last = __shfl_up_sync(0xffffffffu, run_col, 1);
                if (lane == 0) prev_last = carry_col;
                const bool closure = first_col != prev_last;
                const unsigned F = __ballot_sync(0xffffffffu, k >= 1 || closure);
                // segmented inclusive scan of the lanes' open runs; a flag starts a new run
                double T = run;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const double u = __shfl_up_sync(0xffffffffu, T, d);
                    if (lane >= (uint32_t)d && ((F >> (lane - d + 1)) & ((1u << d) - 1u)) == 0u) T += u;
                }
                if ((F & ((2u << lane) - 1u)) == 0u) T += carry_sum;   // still the run carried in from the previous step
                double prevT = __shfl_up_sync(0xffffffffu, T, 1);
                if (lane == 0) prevT = carry_sum;
                const bool prev_is_head = carry_head && (F & ((1u << lane) - 1u)) == 0u;
                if (closure) {
                    // the run that ended with the previous lane is complete
                    if (prev_is_head) { head_c = prev_last; head_s = prevT; }
                    else acc[prev_last] += prevT;
                    if (k >= 1) acc[first_col] += head;
                } else if (k >= 1) {
                    if (prev_is_head) { head_c = first_col; head_s = prevT + head; }
                    else acc[first_col] += prevT + head;
                }
                carry_col = __shfl_sync(0xffffffffu, run_col, 31);
                carry_sum = __shfl_sync(0xffffffffu, T, 31);
                carry_head = carry_head && F == 0u;
            }
            if (s + 1 >= s_end) {
                // the chunk's slot pair: the lane that closed the opening run holds it (at most one does)
                const unsigned H = __ballot_sync(0xffffffffu, head_c >= 0);
                if (H) {
                    const int src = __ffs(H) - 1;
                    head_c = __shfl_sync(0xffffffffu, head_c, src);
                    head_s = __shfl_sync(0xffffffffu, head_s, src);
                }
                if (lane == 0) {
                    if (carry_head) { slot_col[2 * chunk] = carry_col; slot_sum[2 * chunk] = carry_sum; slot_col[2 * chunk + 1] = -1; }
                    else {
                        slot_col[2 * chunk] = head_c; slot_sum[2 * chunk] = head_s;
                        slot_col[2 * chunk + 1] = carry_col; slot_sum[2 * chunk + 1] = carry_sum;
                    }
                }
                chunk = next_chunk_id;
                s = ns;
                s_end = min(nsteps, s + cs);
                carry_col = -1; head_c = -1; carry_sum = 0.0; head_s = 0.0; carry_head = true; first = true;
            } else {
                s = ns;
            }
            cur = nxt;
        }
    }
    cp_async_wait_all();
    __syncthreads();
    if (w == 0) g_stitch(acc, slot_col, slot_sum, prev_slots, lane);
    __syncthreads();
    double *out = P.acc_part + (size_t)blockIdx.x * P.hg;
    for (int32_t j = threadIdx.x; j < P.hg; j += kFeGThreads) out[j] = acc[j];
}

// a warp per (L2 tile, cold column) run of tile `t`: four gathers of dz in flight per lane
__global__ void __launch_bounds__(256) fe_gcold_kernel(const FeTilePlan P, const int64_t t)
{
    const uint32_t lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp; c < P.n_cold; c += nwarps) {
        const int64_t it = t * P.n_cold + c;
        const int64_t b = P.gc_run[it], e = P.gc_run[it + 1];
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int64_t q = b + lane;
        for (; q + 96 < e; q += 128) {
            const double a0 = (double)P.gc_val[q] * P.dz[P.gc_row[q]];
            const double a1 = (double)P.gc_val[q + 32] * P.dz[P.gc_row[q + 32]];
            const double a2 = (double)P.gc_val[q + 64] * P.dz[P.gc_row[q + 64]];
            const double a3 = (double)P.gc_val[q + 96] * P.dz[P.gc_row[q + 96]];
            s0 += a0; s1 += a1; s2 += a2; s3 += a3;
        }
        for (; q < e; q += 32) s0 = fma((double)P.gc_val[q], P.dz[P.gc_row[q]], s0);
        const double s = warp_sum((s0 + s1) + (s2 + s3));
        if (lane == 0) P.cold_part[it] = s;
    }
}

__global__ void __launch_bounds__(256) fe_finish2_kernel(const gdmix_fe_rows R, const gdmix_lr_opts o, const FeTilePlan P,
                                                         const double *x, double *fg)
{
    const int hi = o.has_intercept ? 1 : 0;
    const int64_t D = R.n_features;
    const double l2w = o.l2 / (double)(R.num_workers > 0 ? R.num_workers : 1);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = tid; j < D; j += nth) {
        double s = 0.0;
        if (j < P.hg) {
            for (int32_t c = 0; c < P.g_grid; c++) s += P.acc_part[(size_t)c * P.hg + j];
        } else {
            for (int64_t t = 0; t < P.n_l2_tiles; t++) s += P.cold_part[t * P.n_cold + (j - P.hg)];
        }
        fg[1 + j] = s + l2w * x[j];   // features are always regularised (the intercept is handled apart)
    }
    if (blockIdx.x != 0) return;
    __shared__ double sh[3][256];
    double v = 0.0, dsum = 0.0, sq = 0.0;
    for (int32_t b = threadIdx.x; b < P.z_grid; b += 256) { v += P.block_part[2 * b]; dsum += P.block_part[2 * b + 1]; }
    const int64_t preg = (hi && !o.regularize_bias) ? D : D + hi;
    for (int64_t j = threadIdx.x; j < preg; j += 256) sq = fma(x[j], x[j], sq);
    sh[0][threadIdx.x] = v; sh[1][threadIdx.x] = dsum; sh[2][threadIdx.x] = sq;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s)
            for (int k = 0; k < 3; k++) sh[k][threadIdx.x] += sh[k][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        fg[0] = sh[0][0] + 0.5 * l2w * sh[2][0];
        if (hi) fg[1 + D] = sh[1][0] + (o.regularize_bias ? l2w * x[D] : 0.0);
    }
}

}  // namespace gdmix
