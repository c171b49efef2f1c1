// re_fast.cuh -- the random-effect solver for the common case (m <= 10, at most 2 local features per thread).
//
// Same algorithm and the same per-entity semantics as re_kernel.cuh (see its banner for the reference
// call sites); what differs is where things live and how the two sparse passes walk them:
//
//   * X on chip in sliced-ELL form.  Rows (for z = X1.theta) and columns (for g = X1^T r) are each sorted by
//     length (stable counting sort on chip), cut into slabs of 32 segments, and a slab is stored step-major:
//     step k holds, for each of its 32 lanes, four consecutive non-zeros of that lane's segment
//     ([step][lane][4] values fp32, and the same shape of u16 byte offsets into the gathered vector).  One
//     LDS.128 + one LDS.64 per lane fetch four non-zeros with no bank conflicts and no per-element index
//     arithmetic; lanes of a warp run in lockstep for exactly the slab's step count, and sorting keeps the
//     padding to the last partial quad of each segment.  A step is 33 lane-quads wide, not 32: the phantom
//     quad skews consecutive steps by four banks so that staging's row-wise writes do not collide either.
//   * L-BFGS history (S, Y: 2 m p doubles) in REGISTERS: feature c belongs to thread c mod G, slot c / G
//     (EPT = 1 or 2 slots).  The two dense passes of the compact update (2m+2 inner products with the new
//     gradient; the direction as a combination of S, Y columns) then touch no memory at all.  The intercept's
//     component of every history vector lives in lane s of warp 0 (slot s) and joins in the warp-sized m x m step.
//   * The 22 inner products are reduced across a warp by transposing butterflies, 12 values at a time
//     (18 shuffles per dozen instead of 60).
//   * Staging reads the entity's CSR slice from HBM with row-contiguous (coalesced) loads.
//
// Entities whose sliced form does not fit the shared memory planned for the batch are not solved here: their
// index goes to a deferral list that the general kernel (re_kernel.cuh) drains right afterwards.
#pragma once
#include "linesearch.cuh"
#include "re_common.cuh"
#include "re_lbfgs.cuh"

namespace gdmix {

constexpr int kFastMT = 10;
constexpr uint32_t kKeys = 64;                    // segment lengths are sorted exactly below this, clipped above
constexpr uint32_t kStepQuads = 33;               // lane-quads per slab step (32 + 1 phantom for the bank skew)
constexpr uint32_t kStepElems = 4 * kStepQuads;   // 132 non-zero slots per step
constexpr uint32_t kStepBytes = 6 * kStepElems;   // fp32 value + u16 offset
constexpr uint32_t kFastMaxRows = 8191;           // u16 byte offsets into fp64 vectors
constexpr uint32_t kFastPartK = 24;

using DNF = Dense<kFastMT>;

#ifdef GDMIX_FAST_TIMING
// kernel experiments only: cycles of thread 0 per phase, summed over entities
// [0] pop+stage [1] B1->B2 rows [2] B2->B3 cols [3] B3->B4 dots [4] B4->mxm [5] mxm [6] mxm->B5 [7] B5->B1 dir [8] rest [9] iterations
__device__ unsigned long long g_fast_cycles[12];
#define FT_MARK(k) do { if (tid == 0) { const long long t_ = clock64(); ft_acc[k] += t_ - ft_acc[11]; ft_acc[11] = t_; } } while (0)
#else
#define FT_MARK(k) do { } while (0)
#endif
constexpr uint32_t kFastDense = DNF::count + 2;   // + direction / trial value of the intercept

// Byte offsets of one CTA's dynamic shared memory.  N = max rows, D = max local features of the batch.
struct FastLayout {
    uint32_t xt, gnew, r, xcur, gold, dense, icept, part;    // solve block
    uint32_t rowoff, ccnt, keys, rowpos, colpos, seglen;     // staging scratch, aliases the solve block
    uint32_t rowperm, colperm, sy, sw, soff, rbase, cbase;   // live through the solve
    uint32_t sell_val, sell_idx;
    uint32_t cap_steps;
    uint32_t total_bytes;
    uint32_t max_n, max_d;   // rows / local features the arrays above are sized for
};

__host__ __device__ inline uint32_t fast_fixed_bytes(uint32_t N, uint32_t D, uint32_t W, FastLayout *out)
{
    FastLayout L;
    uint32_t o = 0;
    L.xt = o; o += align16(8 * D);
    L.gnew = o; o += align16(8 * D);
    L.r = o; o += align16(8 * N);
    L.xcur = o; o += align16(8 * D);       // current iterate (features)
    L.gold = o; o += align16(8 * D);       // its gradient
    L.dense = o; o += align16(8 * kFastDense);
    L.icept = o; o += align16(8 * 2 * kFastMT);   // intercept components of the stored pairs: S0[m], Y0[m]
    L.part = o; o += align16(8 * W * kFastPartK);
    const uint32_t solve_end = o;
    o = 0;
    L.rowoff = o; o += align16(4 * (N + 1));
    L.ccnt = o; o += align16(4 * W * D);
    L.keys = o; o += align16(4 * 2 * kKeys);
    L.rowpos = o; o += align16(2 * N);
    L.colpos = o; o += align16(2 * D);
    L.seglen = o; o += align16(2 * D);
    o = o > solve_end ? o : solve_end;
    const uint32_t N32 = (N + 31u) & ~31u, D32 = (D + 31u) & ~31u;
    L.rowperm = o; o += align16(2 * N32);
    L.colperm = o; o += align16(2 * D32);
    L.sy = o; o += align16(4 * N32);
    L.sw = o; o += align16(4 * N32);
    L.soff = o; o += align16(4 * N32);
    L.rbase = o; o += align16(4 * (N32 / 32 + 2));
    L.cbase = o; o += align16(4 * (D32 / 32 + 2));
    L.sell_val = o;
    L.sell_idx = 0; L.cap_steps = 0; L.total_bytes = o;
    L.max_n = N; L.max_d = D;
    if (out) *out = L;
    return o;
}

__host__ __device__ inline FastLayout fast_layout(uint32_t N, uint32_t D, uint32_t W, uint32_t cap_steps)
{
    FastLayout L;
    uint32_t o = fast_fixed_bytes(N, D, W, &L);
    L.cap_steps = cap_steps;
    L.sell_val = o; o += align16(4 * kStepElems * cap_steps);
    L.sell_idx = o; o += align16(2 * kStepElems * cap_steps);
    L.total_bytes = o;
    return L;
}

struct FastArgs {
    ReArgs a;
    FastLayout L;
    int32_t *defer_list;   // entities this kernel could not hold on chip
    int32_t *defer_count;
};

// ---------------------------------------------------------------------------------------------------------
// 12 per-lane partial sums -> totals.  Afterwards v[0..2] of a lane hold the totals of the values
// 6*b4 + 3*b3 + {0,1,2} (b_k = bit k of the lane index).  Fixed order: bitwise reproducible.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_reduce12(double (&v)[12], const uint32_t lane)
{
    {
        const bool up = lane & 16u;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const double keep = up ? v[i + 6] : v[i], send = up ? v[i] : v[i + 6];
            v[i] = keep + __shfl_xor_sync(kFull, send, 16);
        }
    }
    {
        const bool up = lane & 8u;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double keep = up ? v[i + 3] : v[i], send = up ? v[i] : v[i + 3];
            v[i] = keep + __shfl_xor_sync(kFull, send, 8);
        }
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        v[i] += __shfl_xor_sync(kFull, v[i], 4);
        v[i] += __shfl_xor_sync(kFull, v[i], 2);
        v[i] += __shfl_xor_sync(kFull, v[i], 1);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Slab sort: stable counting sort of `nseg` segments by descending clipped length (warp 0), then the step
// count of every slab of 32 and their prefix sums.  pos[i] = sorted position of segment i, perm = inverse,
// base[b] = first step of slab b (base[nslab] = end), starting at base0.  All threads call it.
// ---------------------------------------------------------------------------------------------------------
template <int G, class LenFn>
__device__ __forceinline__ void slab_sort(const LenFn len, const uint32_t nseg, uint16_t *pos, uint16_t *perm,
                                          uint32_t *base, const uint32_t base0, uint32_t *keys)
{
    constexpr uint32_t W = G / 32;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t k = tid; k < 2 * kKeys; k += G) keys[k] = 0;
    group_sync<G>();
    for (uint32_t i = tid; i < nseg; i += G) atomicAdd(&keys[min(len(i), kKeys - 1u)], 1u);
    group_sync<G>();
    if (warp == 0) {
        // lane l owns keys kKeys-1-2l and kKeys-2-2l; starts are counted from the longest key down
        const uint32_t k0 = kKeys - 1u - 2u * lane, k1 = k0 - 1u;
        const uint32_t c0 = keys[k0], c1 = keys[k1];
        uint32_t incl = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        const uint32_t excl = incl - (c0 + c1);
        keys[kKeys + k0] = excl;
        keys[kKeys + k1] = excl + c0;
        __syncwarp();
        for (uint32_t b0 = 0; b0 < nseg; b0 += 32) {
            const uint32_t i = b0 + lane;
            const bool act = i < nseg;
            const uint32_t key = act ? min(len(i), kKeys - 1u) : (kKeys + lane);
            const unsigned grp = __match_any_sync(kFull, key);
            const uint32_t rank = __popc(grp & ((1u << lane) - 1u));
            uint32_t cur = 0;
            if (act) {
                cur = keys[kKeys + key];
                pos[i] = (uint16_t)(cur + rank);
                perm[cur + rank] = (uint16_t)i;
            }
            __syncwarp();
            if (act && rank == (uint32_t)__popc(grp) - 1u) keys[kKeys + key] = cur + rank + 1u;
            __syncwarp();
        }
    }
    group_sync<G>();
    const uint32_t nslab = (nseg + 31u) >> 5;
    for (uint32_t b = warp; b < nslab; b += W) {
        const uint32_t i = b * 32u + lane;
        uint32_t l = (i < nseg) ? len(perm[i]) : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) l = max(l, __shfl_xor_sync(kFull, l, o));
        if (lane == 0) base[b] = (l + 3u) >> 2;
    }
    group_sync<G>();
    if (warp == 0) {
        uint32_t carry = base0;
        for (uint32_t b0 = 0; b0 < nslab; b0 += 32) {
            const uint32_t b = b0 + lane;
            const uint32_t mine = (b < nslab) ? base[b] : 0u;
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(kFull, incl, o);
                if (lane >= (uint32_t)o) incl += t;
            }
            if (b < nslab) base[b] = carry + incl - mine;
            carry += __shfl_sync(kFull, incl, 31);
        }
        if (lane == 0) base[nslab] = carry;
    }
    group_sync<G>();
}

// L2 prefetch hints for the CSR slice of a later entity (one warp; nobody waits for it).
__device__ __forceinline__ void prefetch_entity_l2(const gdmix_re_batch &b, const int64_t e2, const uint32_t lane)
{
    if (e2 >= b.n_entities) return;
    const int64_t r0 = b.ent_rowptr[e2], r1 = b.ent_rowptr[e2 + 1];
    const int64_t q0 = b.rowptr[r0], q1 = b.rowptr[r1];
    auto hint = [&](const char *p0, const char *p1) {
        for (const char *p = (const char *)((uintptr_t)p0 & ~(uintptr_t)127) + 128 * (size_t)lane; p < p1; p += 128 * 32)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    };
    hint((const char *)(b.col + q0), (const char *)(b.col + q1));
    hint((const char *)(b.val + q0), (const char *)(b.val + q1));
    hint((const char *)(b.rowptr + r0), (const char *)(b.rowptr + r1 + 1));
    hint((const char *)(b.label + r0), (const char *)(b.label + r1));
    if (b.offset) hint((const char *)(b.offset + r0), (const char *)(b.offset + r1));
    if (b.weight) hint((const char *)(b.weight + r0), (const char *)(b.weight + r1));
}

__device__ __forceinline__ double lds_f64(const uint32_t shared_addr)
{
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(shared_addr));
    return v;
}

// Shared-state-space accessors for the serial stretch below: one 32-bit base register and immediate offsets instead
// of a generic pointer whose window base the compiler re-derives (S2UR SR_CgaCtaId ...) in every basic block.
__device__ __forceinline__ double lds_f64v(const uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ double2 lds_v2f64v(const uint32_t a)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64v(const uint32_t a, const double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
// 1 / a as the compiler's own division forms it in the normal range (seed + three fused steps), without its slow
// path inline: out of range (never on this path in practice) the library division is called instead.
__device__ __forceinline__ double rcp_f64(const double a)
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    double e = fma(-a, x, 1.0);
    e = fma(e, e, e);
    x = fma(x, e, x);
    e = fma(-a, x, 1.0);
    x = fma(x, e, x);
    if (!(fabs(a) > 1e-290 && fabs(a) < 1e290)) x = 1.0 / a;
    return x;
}

// lbfgs_small_update (re_lbfgs.cuh: same algebra, same order of every sum) for MT = kFastMT as warp 0 of the fast
// kernel runs it.  The stretch is serial -- the other warps of the CTA wait at B5 -- and one warp alone is bound by
// instruction count x latency, so it is written for instruction count: totals arrive in registers (lane l < 2 MT + 2
// holds entry l of [S^T g | Y^T g | y.y | y.g]) and move by shuffles, lanes >= MT are clones of lane MT - 1 (they
// compute and store the same values to the same places, behind the same warp barriers: no `lane < MT` branches), rows
// are read with 16-byte loads.
// D = shared address of dense[0].
__device__ __forceinline__ void fast_small_update(const Lbfgs &L, const bool update, const int newslot,
                                                  const uint32_t dotmask, const double stp, const double dr,
                                                  const double gd_new, const double t, const uint32_t D)
{
    constexpr int MT = kFastMT;
    using DN = Dense<MT>;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t i = lane < (uint32_t)MT ? lane : (uint32_t)MT - 1u;
    double p1i = __shfl_sync(kFull, t, i), p2i = __shfl_sync(kFull, t, MT + i);
    const double yyt = __shfl_sync(kFull, t, 2 * MT), ygt = __shfl_sync(kFull, t, 2 * MT + 1);
    const bool old_i = (dotmask >> i) & 1u;
    const bool new_i = update && (int)i == newslot;
    const uint32_t valid = update ? (L.valid | (1u << newslot)) : L.valid;
    const bool val_i = (valid >> i) & 1u;
    const uint32_t rinv_i = D + 8u * (DN::rinv + i * MT), yy_i = D + 8u * (DN::yy + i * MT), vec_i = D + 8u * i;
    const double inv_dr = update ? rcp_f64(dr) : 0.0;
    const double theta = update ? yyt * inv_dr : L.theta;
    const double gamma = rcp_f64(theta);
    p1i = old_i ? p1i : 0.0;
    p2i = old_i ? p2i : 0.0;
    const double rc = (update && old_i) ? p1i - lds_f64v(vec_i + 8u * DN::p1old) : 0.0;   // s_i . y_new
    const double yc = (update && old_i) ? p2i - lds_f64v(vec_i + 8u * DN::p2old) : 0.0;   // y_i . y_new
    const double p1new = stp * gd_new;                                                   // s_new . g_new
    if (new_i) { p1i = p1new; p2i = ygt; }
    __syncwarp();     // every lane (the clones of lane MT - 1 too) has loaded p1old / p2old before anyone overwrites them
    sts_f64v(vec_i + 8u * DN::ta, rc);
    sts_f64v(vec_i + 8u * DN::tb, (val_i && !new_i) ? p1i : 0.0);
    sts_f64v(vec_i + 8u * DN::p1old, p1i);
    sts_f64v(vec_i + 8u * DN::p2old, p2i);
    __syncwarp();
    double acc = 0.0, wvp = 0.0;
#pragma unroll
    for (int j = 0; j < MT; j += 2) {
        const double2 r = lds_v2f64v(rinv_i + 8u * j);
        const double2 a2 = lds_v2f64v(D + 8u * (DN::ta + j)), b2 = lds_v2f64v(D + 8u * (DN::tb + j));
        acc = fma(r.x, a2.x, acc); wvp = fma(r.x, b2.x, wvp);
        acc = fma(r.y, a2.y, acc); wvp = fma(r.y, b2.y, wvp);
    }
    double wv = val_i ? wvp : 0.0;
    __syncwarp();                                     // every lane is done reading the old R^-1
    if (update) {
        const double cnew = old_i ? -acc * inv_dr : 0.0;          // new column of R^-1 (old rows)
        wv = new_i ? inv_dr * p1new : (old_i ? fma(cnew, p1new, wvp) : 0.0);
        const uint32_t col_new = 8u * (uint32_t)newslot, row_new = 8u * MT * (uint32_t)newslot + 8u * i;
        sts_f64v(rinv_i + col_new, new_i ? inv_dr : cnew);
        sts_f64v(D + 8u * DN::rinv + row_new, new_i ? inv_dr : 0.0);
        sts_f64v(yy_i + col_new, new_i ? yyt : yc);
        sts_f64v(D + 8u * DN::yy + row_new, new_i ? yyt : yc);
        if (new_i) sts_f64v(vec_i + 8u * DN::d, dr);
    }
    sts_f64v(vec_i + 8u * DN::cw, wv);
    __syncwarp();
    double yw = 0.0;
#pragma unroll
    for (int j = 0; j < MT; j += 2) {
        const double2 y2 = lds_v2f64v(yy_i + 8u * j), c2 = lds_v2f64v(D + 8u * (DN::cw + j));
        yw = fma(y2.x, c2.x, yw);
        yw = fma(y2.y, c2.y, yw);
    }
    const double tv = val_i ? fma(lds_f64v(vec_i + 8u * DN::d), wv, gamma * (yw - p2i)) : 0.0;
    sts_f64v(vec_i + 8u * DN::ta, tv);
    __syncwarp();
    double uv = 0.0;
#pragma unroll
    for (int j = 0; j < MT; j += 2) {
        const double2 a2 = lds_v2f64v(D + 8u * (DN::ta + j));
        uv = fma(lds_f64v(D + 8u * (DN::rinv + j * MT) + 8u * i), a2.x, uv);
        uv = fma(lds_f64v(D + 8u * (DN::rinv + (j + 1) * MT) + 8u * i), a2.y, uv);
    }
    sts_f64v(vec_i + 8u * DN::cu, uv);
    if (lane == 0) { sts_f64v(D + 8u * (DN::tot + 2 * MT), theta); sts_f64v(D + 8u * (DN::tot + 2 * MT + 1), gamma); }
}

// One lane's share of a slab: sum over `nsteps` quads of val * vec[...].  pv / pi point at this lane's quad of
// the slab's first step.  The index stream holds ABSOLUTE 16-bit shared-memory addresses of the gathered fp64
// entries (the gathered vectors sit in the first 64 KB of the CTA's window), so a gather is one LDS with no
// address arithmetic.
__device__ __forceinline__ double sell_dot(const float4 *pv, const uint2 *pi, const uint32_t nsteps)
{
    // Software pipelined: the value / address quads of step k+1 are in flight while step k's four gathers and
    // FMAs run (index load -> gather -> FMA is a chain of two shared-memory round trips otherwise); two steps per
    // trip with the two register sets alternating, so nothing is copied from "next" to "current".
    // Accumulation order per chain: steps ascending, as the data is laid out.
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (nsteps == 0) return 0.0;
    auto step = [&](const float4 &v, const uint2 &c) {
        const double x0 = lds_f64(c.x & 0xffffu), x1 = lds_f64(c.x >> 16);
        const double x2 = lds_f64(c.y & 0xffffu), x3 = lds_f64(c.y >> 16);
        s0 = fma((double)v.x, x0, s0);
        s1 = fma((double)v.y, x1, s1);
        s2 = fma((double)v.z, x2, s2);
        s3 = fma((double)v.w, x3, s3);
    };
    float4 va = pv[0], vb;
    uint2 ca = pi[0], cb;
    uint32_t k = 1;
    for (; k + 1 < nsteps; k += 2) {
        vb = pv[k * kStepQuads];
        cb = pi[k * kStepQuads];
        step(va, ca);
        va = pv[(k + 1) * kStepQuads];
        ca = pi[(k + 1) * kStepQuads];
        step(vb, cb);
    }
    if (k < nsteps) {
        vb = pv[k * kStepQuads];
        cb = pi[k * kStepQuads];
        step(va, ca);
        step(vb, cb);
    } else {
        step(va, ca);
    }
    return (s0 + s1) + (s2 + s3);
}

// The staged entity as the passes see it: five small integers; every array is addressed as smem + an offset
// of the launch-constant FastLayout (kernel parameter space), so no pointer stays live across the solve.
struct FastDims {
    uint32_t n, d, hi;
};

// ---------------------------------------------------------------------------------------------------------
// Staging.  Returns 0 = staged, 1 = invalid input (column out of range), 2 = does not fit (defer).
// ---------------------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ int fast_stage(const ReArgs &a, const FastLayout &L, unsigned char *smem,
                                          const int64_t r0, const int64_t q0, const uint32_t n, const uint32_t d,
                                          unsigned *s_flag)
{
    constexpr uint32_t W = G / 32;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *rowoff = (uint32_t *)(smem + L.rowoff), *ccnt = (uint32_t *)(smem + L.ccnt);
    uint32_t *keys = (uint32_t *)(smem + L.keys);
    uint16_t *rowpos = (uint16_t *)(smem + L.rowpos), *colpos = (uint16_t *)(smem + L.colpos);
    uint16_t *seglen = (uint16_t *)(smem + L.seglen);
    uint16_t *rowperm = (uint16_t *)(smem + L.rowperm), *colperm = (uint16_t *)(smem + L.colperm);
    uint32_t *rbase = (uint32_t *)(smem + L.rbase), *cbase = (uint32_t *)(smem + L.cbase);
    float *sval = (float *)(smem + L.sell_val);
    uint16_t *sidx = (uint16_t *)(smem + L.sell_idx);
    float *sy = (float *)(smem + L.sy), *sw = (float *)(smem + L.sw), *soff = (float *)(smem + L.soff);
    // 16-bit absolute shared addresses of xt[0] and r[0] (the planner keeps both vectors below 64 KB)
    const uint32_t xt_abs = (uint32_t)__cvta_generic_to_shared(smem + L.xt);
    const uint32_t r_abs = (uint32_t)__cvta_generic_to_shared(smem + L.r);

    for (uint32_t i = tid; i <= n; i += G) rowoff[i] = (uint32_t)(a.b.rowptr[r0 + i] - q0);
    for (uint32_t k = tid; k < W * d; k += G) ccnt[k] = 0;
    group_sync<G>();

    // ---- rows: sort, slab bases, zero fill, fill (coalesced: a warp streams one row at a time) -------------
    slab_sort<G>([&](uint32_t i) { return rowoff[i + 1] - rowoff[i]; }, n, rowpos, rowperm, rbase, 0u, keys);
    const uint32_t nrslab = (n + 31u) >> 5;
    const uint32_t total_r = rbase[nrslab];
    if (total_r > L.cap_steps) return 2;
    {
        uint4 *zv = (uint4 *)sval;
        const uint32_t nv = total_r * (kStepElems / 4);
        for (uint32_t k = tid; k < nv; k += G) zv[k] = make_uint4(0, 0, 0, 0);
        uint4 *zi = (uint4 *)sidx;
        const uint32_t ni = (total_r * kStepElems * 2 + 15) / 16;
        const uint32_t fill = xt_abs * 0x10001u;  // padding slots gather xt[0] (times a zero value)
        for (uint32_t k = tid; k < ni; k += G) zi[k] = make_uint4(fill, fill, fill, fill);
    }
    for (uint32_t i = tid; i < n; i += G) {
        const uint32_t sp = rowpos[i];
        sy[sp] = a.b.label[r0 + i];
        sw[sp] = a.b.weight ? a.b.weight[r0 + i] : 1.0f;
        soff[sp] = a.b.offset ? a.b.offset[r0 + i] : 0.0f;
    }
    group_sync<G>();
    const uint32_t chunk = (n + W - 1) / W;
    const uint32_t rbeg = min(n, warp * chunk), rend = min(n, rbeg + chunk);
    unsigned bad = 0;
    auto put = [&](const uint32_t at, const uint32_t j, const uint32_t c, const float v) {
        if (c < d) {
            const uint32_t dst = at + (j >> 2) * kStepElems + (j & 3u);
            sval[dst] = v;
            sidx[dst] = (uint16_t)(xt_abs + c * 8u);
            atomicAdd(&ccnt[warp * d + c], 1u);
        } else {
            bad = 1;
        }
    };
    // eight rows per trip: all their loads are in flight before the first is consumed (one warp streams one
    // row at a time, so without this every row would cost a full HBM round trip)
    for (uint32_t i0 = rbeg; i0 < rend; i0 += 8) {
        uint32_t cc[8], ss[8], ll[8];
        float vv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t i = i0 + u;
            ss[u] = 0; ll[u] = 0; cc[u] = 0; vv[u] = 0.0f;
            if (i < rend) {
                ss[u] = rowoff[i];
                ll[u] = rowoff[i + 1] - ss[u];
                if (lane < ll[u]) {
                    cc[u] = (uint32_t)a.b.col[q0 + ss[u] + lane];
                    vv[u] = a.b.val[q0 + ss[u] + lane];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (ll[u] == 0) continue;
            const uint32_t sp = rowpos[i0 + u];
            const uint32_t at = rbase[sp >> 5] * kStepElems + (sp & 31u) * 4u;
            if (lane < ll[u]) put(at, lane, cc[u], vv[u]);
            for (uint32_t j = lane + 32; j < ll[u]; j += 32)
                put(at, j, (uint32_t)a.b.col[q0 + ss[u] + j], a.b.val[q0 + ss[u] + j]);
        }
    }
    if (bad) atomicOr(s_flag, 1u);
    group_sync<G>();
    if (*s_flag) return 1;

    // ---- columns: lengths and per-warp cursors, sort, slab bases, zero fill, row-ordered fill ------------
    for (uint32_t c = tid; c < d; c += G) {
        uint32_t run = 0;
#pragma unroll
        for (uint32_t w2 = 0; w2 < W; w2++) {
            const uint32_t t = ccnt[w2 * d + c];
            ccnt[w2 * d + c] = run;
            run += t;
        }
        seglen[c] = (uint16_t)run;
    }
    group_sync<G>();
    slab_sort<G>([&](uint32_t c) { return (uint32_t)seglen[c]; }, d, colpos, colperm, cbase, total_r, keys);
    const uint32_t ncslab = (d + 31u) >> 5;
    const uint32_t total = cbase[ncslab];
    if (total > L.cap_steps) return 2;
    {
        uint4 *zv = (uint4 *)(sval + (size_t)total_r * kStepElems);
        const uint32_t nv = (total - total_r) * (kStepElems / 4);
        for (uint32_t k = tid; k < nv; k += G) zv[k] = make_uint4(0, 0, 0, 0);
        // u16 region: kStepElems * 2 = 264 B per step, 8-byte granular
        uint2 *zi = (uint2 *)(sidx + (size_t)total_r * kStepElems);
        const uint32_t ni = (total - total_r) * (kStepElems / 4);
        const uint32_t fill = r_abs * 0x10001u;  // padding slots gather r[0] (times a zero value)
        for (uint32_t k = tid; k < ni; k += G) zi[k] = make_uint2(fill, fill);
    }
    group_sync<G>();
    // second sweep out of the sliced rows just built (no global traffic): scatter every non-zero to its
    // column's next free slot.  A warp walks its rows in ascending order, so a column's entries end up in row order.
    for (uint32_t i = rbeg; i < rend; i++) {
        const uint32_t len = rowoff[i + 1] - rowoff[i];
        const uint32_t rp = rowpos[i];
        const uint32_t rat = rbase[rp >> 5] * kStepElems + (rp & 31u) * 4u;
        for (uint32_t j0 = 0; j0 < len; j0 += 32) {
            const uint32_t j = j0 + lane;
            const bool act = j < len;
            const uint32_t src = rat + (j >> 2) * kStepElems + (j & 3u);
            const uint32_t c = act ? (((uint32_t)sidx[src] - xt_abs) >> 3) : (0x10000u + lane);
            const float v = act ? sval[src] : 0.0f;
            // columns strictly ascending inside the chunk (the usual case) => no column occurs twice
            const uint32_t prev = __shfl_up_sync(kFull, c, 1);
            const bool uniq = __all_sync(kFull, lane == 0 || c > prev);
            unsigned grp = 1u << lane;
            if (!uniq) grp = __match_any_sync(kFull, c);
            const uint32_t rank = __popc(grp & ((1u << lane) - 1u));
            uint32_t cur = 0;
            if (act) {
                cur = ccnt[warp * d + c];
                const uint32_t e = cur + rank, sp = colpos[c];
                const uint32_t dst = (cbase[sp >> 5] + (e >> 2)) * kStepElems + (sp & 31u) * 4u + (e & 3u);
                sval[dst] = v;
                sidx[dst] = (uint16_t)(r_abs + i * 8u);
            }
            __syncwarp();
            if (act && rank == (uint32_t)__popc(grp) - 1u) ccnt[warp * d + c] = cur + rank + 1u;
            __syncwarp();
        }
    }
    group_sync<G>();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Cold half of the line search.  The first trial step is accepted almost always; only when it is not does the
// More'-Thuente state exist at all, and then it lives in shared memory (two slots used alternately: every
// thread computes and writes the same values into the slot nobody is reading).
// ---------------------------------------------------------------------------------------------------------
struct LsOut {
    double stp;
    int task;
};

__device__ __noinline__ LsOut ls_continue(LineSearch *slots, const int read_slot, const bool first,
                                          const double stp_first, const double fold, const double gdold,
                                          const double stp_in, const double f, const double g)
{
    const double ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmx = 1e10;
    LineSearch S;
    if (first) {
        double s1 = stp_first;  // dcsrch's START call: derives the state from the values at step 0
        dcsrch(s1, fold, gdold, ftol, gtol, xtol, 0.0, stpmx, LS_START, S);
    } else {
        S = slots[read_slot];
    }
    LsOut o;
    o.stp = stp_in;
    o.task = dcsrch(o.stp, f, g, ftol, gtol, xtol, 0.0, stpmx, LS_FG, S);
    slots[read_slot ^ 1] = S;
    return o;
}

// Writes the features of the trial point x + stp d into smem xt (each thread its own slots).
template <int G, int EPT>
__device__ __forceinline__ void fast_trial(const FastArgs &fa, unsigned char *smem, const FastDims &E,
                                           const double stp, const double (&dd)[EPT])
{
    double *xt = (double *)(smem + fa.L.xt);
    const double *xcur = (const double *)(smem + fa.L.xcur);
#pragma unroll
    for (int k = 0; k < EPT; k++) {
        const uint32_t c = threadIdx.x + (uint32_t)k * G;
        if (c < E.d) xt[c] = fma(stp, dd[k], xcur[c]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// The kernel.  Per L-BFGS iteration, in the common case (first trial step accepted), five block barriers:
//   B1  trial point xt published (and the per-warp partials of g.d for the direction just built)
//       rows: z = X1.xt + offset, cross entropy, residual r            -> B2 (r, and sums of cost / r / xt^2)
//       cols: g_new = (X1^T r + l2 xt) / n into smem                   -> B3
//       each thread takes its own slots of g_new back into registers and, SPECULATING that the step will be
//       accepted, forms its share of the 2m+2 inner products of the compact update together with g_new.d and
//       max|g_new|                                                     -> B4 (per-warp partials)
//       all threads: Wolfe test on the totals; accept; stop tests; store the new pair in registers;
//       warp 0: the m x m step (re_lbfgs.cuh)                          -> B5 (coefficients u, w)
//       next direction d = -gamma g - S u + gamma Y w from registers, next trial point x + d -> B1
// ---------------------------------------------------------------------------------------------------------
template <int G, int EPT>
// Registers: one feature slot per thread keeps 40 history registers and fits 128 per thread (512 threads per SM
// more than the two-slot variant's 168, which is what lets the typical-shape launch of a ragged batch run four
// 128-thread CTAs per SM).
__global__ void __launch_bounds__(G, (EPT == 1) ? 512 / G : ((G <= 128) ? 384 / G : 1)) re_fast_kernel(const FastArgs fa)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double red[2 * kMaxWarps * kRedK];
    __shared__ double wred[2 * kMaxWarps];   // per-warp partials: [w] g.d after H2, [kMaxWarps + w] max|g_new|
    __shared__ LineSearch ls_slots[2];
    __shared__ int s_entity;
    __shared__ unsigned s_flag;

    constexpr int MT = kFastMT;
    constexpr uint32_t W = G / 32;
    const ReArgs &a = fa.a;
    const FastLayout &L = fa.L;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int flip = 0;
#ifdef GDMIX_FAST_TIMING
    __shared__ long long ft_acc[12];     // in shared memory: the kernel has no registers to spare
    if (tid == 0) { for (int k = 0; k < 11; k++) ft_acc[k] = 0; ft_acc[11] = clock64(); }
#endif

    for (;;) {
        group_sync<G>();  // previous entity fully emitted before its memory is reused
        FT_MARK(8);
        if (tid == 0) {
            // plain launch: entities 0 .. n_entities; second tier: the list the typical-shape launch deferred
            int idx = atomicAdd(a.queue, 1);
            if (a.todo) idx = (idx < *a.todo_count) ? a.todo[idx] : 0x7fffffff;
            s_entity = idx;
            s_flag = 0;
        }
        group_sync<G>();
        const int e = s_entity;
        if ((int64_t)e >= a.b.n_entities) break;

        FastDims E;
        E.hi = a.o.has_intercept ? 1u : 0u;
        int st = 2;
        {
            const int64_t r0 = a.b.ent_rowptr[e], r1 = a.b.ent_rowptr[e + 1];
            const int64_t q0 = a.b.rowptr[r0], q1 = a.b.rowptr[r1];
            const int64_t n64 = r1 - r0, nnz64 = q1 - q0, p64 = a.b.theta_ptr[e + 1] - a.b.theta_ptr[e];
            E.n = (uint32_t)n64; E.d = (uint32_t)p64 - E.hi;
            const bool shape_ok = n64 >= 1 && p64 >= 1 && p64 >= (int64_t)E.hi && nnz64 >= 0 && nnz64 < (1ll << 30);
            if (!shape_ok) {
                if (tid == 0 && a.status) a.status[e] = GDMIX_ERR_TOO_LARGE;
                continue;
            }
            if (n64 <= (int64_t)L.max_n && E.d <= L.max_d && (int64_t)E.d <= (int64_t)G * EPT && a.o.m <= MT)
                st = fast_stage<G>(a, L, smem, r0, q0, E.n, E.d, &s_flag);
        }
        if (st == 1) {
            if (tid == 0 && a.status) a.status[e] = GDMIX_ERR_INVALID;
            continue;
        }
        if (st == 2) {
            if (tid == 0) fa.defer_list[atomicAdd(fa.defer_count, 1)] = (int32_t)e;
            continue;
        }
        FT_MARK(0);
        double *dense = (double *)(smem + L.dense);
        double *part = (double *)(smem + L.part);
        if (W == 1) prefetch_entity_l2(a.b, (int64_t)e + gridDim.x, lane);

        // One solve per regularisation weight out of the block just staged (gdmix_re_fit_sweep; a plain fit is
        // the sweep of length one over o.l2).
        const int n_models = a.n_l2 > 0 ? a.n_l2 : 1;
        for (int li = 0; li < n_models; li++) {
        if (li) group_sync<G>();   // the previous model fully emitted before the solve block is rewritten

        // ---- solver state: this thread's slots of x, g, direction and of every history vector -----------
        // (the iterate and its gradient live in smem xcur / gold, each thread touching only its own slots;
        //  the intercept's components of the stored pairs in smem icept[], touched by warp 0 only)
        double gn[EPT], dd[EPT], Sh[MT][EPT], Yh[MT][EPT];
        double x0 = 0.0, g0 = 0.0, gn0 = 0.0, d0 = 0.0;
        double *xcur = (double *)(smem + L.xcur), *gold = (double *)(smem + L.gold);
        double *icept = (double *)(smem + L.icept);
        {
            const int64_t t0 = a.b.theta_ptr[e];
#pragma unroll
            for (int k = 0; k < EPT; k++) {
                const uint32_t c = tid + (uint32_t)k * G;
                if (c < E.d) {
                    xcur[c] = a.theta_in ? a.theta_in[t0 + E.hi + c] : 0.0;
                    gold[c] = 0.0;
                }
                gn[k] = 0.0; dd[k] = 0.0;
#pragma unroll
                for (int s = 0; s < MT; s++) { Sh[s][k] = 0.0; Yh[s][k] = 0.0; }
            }
            if (E.hi && a.theta_in) x0 = a.theta_in[t0];
            if (tid < 2u * MT) icept[tid] = 0.0;
        }
        Lbfgs lb;
        lbfgs_reset<G, MT>(lb, dense);

        const double epsmch = 2.220446049250313e-16;
        const double ftol = 1e-3, gtol = 0.9, stpmx = 1e10;
        const double l2 = a.n_l2 > 0 ? a.l2_sweep[li] : a.o.l2, inv_n = 1.0 / (double)E.n;
        enum { PH_INIT = 0, PH_FIRST = 1, PH_MORE = 2 };
        int phase = PH_INIT, nfev = 0, iter = 0, status = GDMIX_SOLVE_CONVERGED, ifun = 0, ls_slot = 0;
        double f = 0.0, gd = 0.0, fold = 0.0, gdold = 0.0, stp = 0.0;
        bool gd_pending = false, emitted = false;
        fast_trial<G, EPT>(fa, smem, E, 0.0, dd);

        // A search along the steepest-descent direction: the first iteration, or a restart after a failed search.
        auto steepest_search = [&]() {
            double v1[1] = {0.0};
#pragma unroll
            for (int k = 0; k < EPT; k++) {
                const uint32_t c = tid + (uint32_t)k * G;
                const double gk = (c < E.d) ? gold[c] : 0.0;
                dd[k] = -gk;
                v1[0] = fma(gk, gk, v1[0]);
            }
            d0 = -g0;
            if (tid == 0) v1[0] = fma(g0, g0, v1[0]);
            group_sum<G, 1>(v1, red, flip);
            gd = -v1[0];
            gd_pending = false;
            stp = (iter == 0) ? fmin(1.0 / sqrt(v1[0]), stpmx) : 1.0;
            fast_trial<G, EPT>(fa, smem, E, stp, dd);
            phase = PH_FIRST;
        };

        for (;;) {
            group_sync<G>();                                                    // ---- B1
            FT_MARK(7);
            if (gd_pending) {
                double t = wred[0];
#pragma unroll
                for (uint32_t w2 = 1; w2 < W; w2++) t += wred[w2];
                gd = t;
                gd_pending = false;
            }
            bool failed = false;
            if (phase == PH_FIRST) {
                // a fresh search starts here: dcsrch's START checks
                fold = f; gdold = gd; ifun = 1;
                if (gd >= 0.0 || stp < 0.0 || stp > stpmx || a.o.max_ls <= 0) failed = true;
            }
            if (!failed) {
                // ================= evaluate f, g at xt ====================================================
                double ft, gdt, gmt;
                {
                    const double xt0 = fma(stp, d0, x0);
                    double p3[3] = {0.0, 0.0, 0.0};   // sums of cost, r, xt_reg^2
#pragma unroll
                    for (int k = 0; k < EPT; k++) {
                        const uint32_t c = tid + (uint32_t)k * G;
                        if (c < E.d) { const double t = ((const double *)(smem + L.xt))[c]; p3[2] = fma(t, t, p3[2]); }
                    }
                    if (E.hi && a.o.regularize_bias && tid == 0) p3[2] = fma(xt0, xt0, p3[2]);
                    const uint32_t *rbase = (const uint32_t *)(smem + L.rbase);
                    for (uint32_t b = warp; b < ((E.n + 31u) >> 5); b += W) {
                        const uint32_t sb = rbase[b];
                        double z = sell_dot((const float4 *)(smem + L.sell_val) + sb * kStepQuads + lane,
                                            (const uint2 *)(smem + L.sell_idx) + sb * kStepQuads + lane,
                                            rbase[b + 1] - sb);
                        const uint32_t sp = b * 32u + lane;
                        if (sp < E.n) {
                            z = (z + (E.hi ? xt0 : 0.0)) + (double)((const float *)(smem + L.soff))[sp];
                            const double yi = (double)((const float *)(smem + L.sy))[sp];
                            const double wi = (double)((const float *)(smem + L.sw))[sp];
                            double ez, l1p, inv;
                            logistic_terms(z, ez, l1p, inv);     // exp(-|z|), log(1 + ez), 1 / (1 + ez)
                            const double ce = fmax(z, 0.0) - z * yi + l1p;
                            p3[0] = fma(wi, ce, p3[0]);
                            const double sig = (z >= 0.0) ? inv : ez * inv;
                            const double ri = wi * (sig - yi);
                            ((double *)(smem + L.r))[((const uint16_t *)(smem + L.rowperm))[sp]] = ri;
                            p3[1] += ri;
                        }
                    }
                    group_sum<G, 3>(p3, red, flip);                             // ---- B2 (publishes r[])
                    FT_MARK(1);
                    if (G == 32) __syncwarp();
                    ft = (p3[0] + 0.5 * l2 * p3[2]) * inv_n;
                    gn0 = E.hi ? (p3[1] + (a.o.regularize_bias ? l2 * xt0 : 0.0)) * inv_n : 0.0;
                }
                {
                    const uint32_t *cbase = (const uint32_t *)(smem + L.cbase);
                    for (uint32_t b = warp; b < ((E.d + 31u) >> 5); b += W) {
                        const uint32_t sb = cbase[b];
                        const double acc = sell_dot((const float4 *)(smem + L.sell_val) + sb * kStepQuads + lane,
                                                    (const uint2 *)(smem + L.sell_idx) + sb * kStepQuads + lane,
                                                    cbase[b + 1] - sb);
                        const uint32_t sp = b * 32u + lane;
                        if (sp < E.d) {
                            const uint32_t c = ((const uint16_t *)(smem + L.colperm))[sp];
                            ((double *)(smem + L.gnew))[c] = (acc + l2 * ((const double *)(smem + L.xt))[c]) * inv_n;
                        }
                    }
                }
                group_sync<G>();                                                // ---- B3 (publishes g_new)
                FT_MARK(2);
                // own slots of g_new; speculative inner products with the stored pairs (12 at a time)
                double gm = fabs(gn0);
#pragma unroll
                for (int k = 0; k < EPT; k++) {
                    const uint32_t c = tid + (uint32_t)k * G;
                    gn[k] = (c < E.d) ? ((const double *)(smem + L.gnew))[c] : 0.0;
                    gm = pos_max(gm, fabs(gn[k]));
                }
                {
                    double v[12];
#pragma unroll
                    for (int s = 0; s < 12; s++) v[s] = 0.0;
#pragma unroll
                    for (int k = 0; k < EPT; k++) {
                        const uint32_t c = tid + (uint32_t)k * G;
                        const double gj = gn[k], yj = gj - ((c < E.d) ? gold[c] : 0.0);
                        v[MT] = fma(yj, yj, v[MT]);
                        v[MT + 1] = fma(yj, gj, v[MT + 1]);
#pragma unroll
                        for (int s = 0; s < MT; s++) v[s] = fma(Sh[s][k], gj, v[s]);
                    }
                    warp_reduce12(v, lane);
                    if ((lane & 7u) == 0) {
                        const uint32_t at = warp * kFastPartK + 6u * ((lane >> 4) & 1u) + 3u * ((lane >> 3) & 1u);
                        part[at] = v[0]; part[at + 1] = v[1]; part[at + 2] = v[2];
                    }
                }
                {
                    double v[12];
#pragma unroll
                    for (int s = 0; s < 12; s++) v[s] = 0.0;
#pragma unroll
                    for (int k = 0; k < EPT; k++) {
                        const double gj = gn[k];
                        v[MT] = fma(gj, dd[k], v[MT]);   // g_new . d
#pragma unroll
                        for (int s = 0; s < MT; s++) v[s] = fma(Yh[s][k], gj, v[s]);
                    }
                    if (tid == 0) v[MT] = fma(gn0, d0, v[MT]);
                    warp_reduce12(v, lane);
                    if ((lane & 7u) == 0) {
                        const uint32_t at = warp * kFastPartK + 12u + 6u * ((lane >> 4) & 1u) + 3u * ((lane >> 3) & 1u);
                        part[at] = v[0]; part[at + 1] = v[1]; part[at + 2] = v[2];
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) gm = pos_max(gm, __shfl_xor_sync(kFull, gm, o));
                if (lane == 0) wred[kMaxWarps + warp] = gm;
                group_sync<G>();                                                // ---- B4
                FT_MARK(3);
                gdt = part[2 * MT + 2];
                gmt = wred[kMaxWarps];
#pragma unroll
                for (uint32_t w2 = 1; w2 < W; w2++) {
                    gdt += part[w2 * kFastPartK + 2 * MT + 2];
                    gmt = pos_max(gmt, wred[kMaxWarps + w2]);
                }
                nfev++;
                // ================= what the evaluation was for ===========================================
                if (phase == PH_INIT) {
                    f = ft;
#pragma unroll
                    for (int k = 0; k < EPT; k++) {
                        const uint32_t c = tid + (uint32_t)k * G;
                        if (c < E.d) gold[c] = gn[k];
                    }
                    g0 = gn0;
                    if (a.mode == kModeLossGrad) {
                        const int64_t t0 = a.b.theta_ptr[e];
#pragma unroll
                        for (int k = 0; k < EPT; k++) {
                            const uint32_t c = tid + (uint32_t)k * G;
                            if (c < E.d) a.g_out[t0 + E.hi + c] = gn[k];
                        }
                        if (tid == 0) {
                            if (E.hi) a.g_out[t0] = g0;
                            a.f_out[e] = f;
                        }
                        emitted = true;
                        break;
                    }
                    if (gmt <= a.o.pgtol) break;
                    steepest_search();
                    continue;
                }
                bool accept = false;
                if (phase == PH_FIRST)   // dcsrch's convergence test, on the state its START call would have set
                    accept = (ft <= fold + stp * (ftol * gdold)) && (fabs(gdt) <= gtol * (-gdold));
                if (!accept) {
                    const LsOut o = ls_continue(ls_slots, ls_slot, phase == PH_FIRST, stp, fold, gdold, stp, ft, gdt);
                    ls_slot ^= 1;
                    if (o.task == LS_CONV || o.task == LS_WARN) {
                        accept = true;
                    } else if (o.task == LS_ERROR || ifun >= a.o.max_ls) {
                        failed = true;
                    } else {
                        ifun++;
                        stp = o.stp;
                        fast_trial<G, EPT>(fa, smem, E, stp, dd);
                        phase = PH_MORE;
                        continue;
                    }
                }
                if (accept) {
                    iter++;
                    f = ft; gd = gdt;
                    // the accepted point is the trial point; gold still holds the old gradient
#pragma unroll
                    for (int k = 0; k < EPT; k++) {
                        const uint32_t c = tid + (uint32_t)k * G;
                        if (c < E.d) xcur[c] = ((const double *)(smem + L.xt))[c];
                    }
                    x0 = fma(stp, d0, x0);
                    if (iter >= a.o.max_iter || nfev > a.o.max_fun) { status = GDMIX_SOLVE_MAXITER; break; }
                    if (gmt <= a.o.pgtol) break;
                    if ((fold - f) <= epsmch * a.o.factr * max3(fabs(fold), fabs(f), 1.0)) break;

                    // ---- curvature pair (L-BFGS-B's skip rule) and the next direction -----------------------
                    double dr, ddum;
                    if (stp == 1.0) { dr = gd - gdold; ddum = -gdold; }
                    else { dr = (gd - gdold) * stp; ddum = -gdold * stp; }
                    const int m = a.o.m;
                    const bool update = (m > 0) && !(dr <= epsmch * ddum);
                    int newslot = -1;
                    if (update) { newslot = (lb.col < m) ? lb.head + lb.col : lb.head; if (newslot >= m) newslot -= m; }   // head, col < m
                    const uint32_t dotmask = update ? (lb.valid & ~(1u << newslot)) : lb.valid;
                    // the new pair takes its slot: s = stp d, y = g_new - g_old
                    // (newslot = -1 without an update: no slot matches.  Predicated moves IN PLACE: written as
                    //  `if (s == newslot) Sh[s][k] = sj` the compiler kept two copies of the history registers and
                    //  moved one onto the other twice per iteration -- 190 instructions for these 4 EPT stores.)
                    {
                        double sj[EPT], yj[EPT];
#pragma unroll
                        for (int k = 0; k < EPT; k++) {
                            const uint32_t c = tid + (uint32_t)k * G;
                            sj[k] = stp * dd[k];
                            yj[k] = gn[k] - ((c < E.d) ? gold[c] : 0.0);
                        }
#pragma unroll
                        for (int s = 0; s < MT; s++) {
#pragma unroll
                            for (int k = 0; k < EPT; k++)
                                asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %4, %5;\n\t@p mov.f64 %0, %2;\n\t@p mov.f64 %1, %3;\n\t}"
                                    : "+d"(Sh[s][k]), "+d"(Yh[s][k])
                                    : "d"(sj[k]), "d"(yj[k]), "r"(newslot), "r"(s));
                        }
                    }
                    FT_MARK(4);
                    if (warp == 0) {
                        const double y0 = gn0 - g0;
                        // part row: [S^T g (10) | y.y | y.g | Y^T g (10) | g.d | -]; lane l < 2 MT + 2 takes entry l of
                        // [S^T g | Y^T g | y.y | y.g], summed over the warps, plus the intercept's share
                        double t = 0.0;
                        if (lane < 2 * MT + 2) {
                            const uint32_t src = (lane < (uint32_t)MT) ? lane : (lane < 2u * MT) ? lane + 2u : lane - MT;
                            t = part[src];
#pragma unroll
                            for (uint32_t w2 = 1; w2 < W; w2++) t += part[w2 * kFastPartK + src];
                            if (E.hi) {
                                // icept[] = [S0 (m) | Y0 (m)], same order as the totals
                                if (lane < 2u * MT) t = fma(icept[lane], gn0, t);
                                else if (lane == 2u * MT) t = fma(y0, y0, t);
                                else t = fma(y0, gn0, t);
                            }
                        }
                        fast_small_update(lb, update, newslot, dotmask, stp, dr, gd, t,
                                          (uint32_t)__cvta_generic_to_shared(dense));
                        // the intercept's components of the new pair (its direction component is formed by every
                        // thread for itself in H2 below: that keeps it off this serial stretch)
                        if (update && (int)lane == newslot) { icept[lane] = stp * d0; icept[MT + lane] = y0; }
                        FT_MARK(5);
#ifdef GDMIX_FAST_TIMING
                        if (tid == 0) ft_acc[9] += 1;
#endif
                    } else if (warp == 1 && iter == 1) {
                        // idle while warp 0 works: pull the entity a grid-width ahead in the queue towards L2
                        prefetch_entity_l2(a.b, (int64_t)e + gridDim.x, lane);
                    }
                    group_sync<G>();                                            // ---- B5
                    FT_MARK(6);
                    lb.theta = dense[DNF::tot + 2 * MT];   // warp 0 left theta and gamma = 1 / theta there
                    if (update) {
                        lb.valid |= (1u << newslot);
                        if (lb.col < m) lb.col++; else lb.head = (lb.head + 1 == m) ? 0 : lb.head + 1;
                    }
                    // H2: d = -gamma g - S u + gamma Y w, the next trial point x + d and the partials of g.d
                    {
                        const double gamma = dense[DNF::tot + 2 * MT + 1];
                        double acc[EPT];
#pragma unroll
                        for (int k = 0; k < EPT; k++) acc[k] = -gamma * gn[k];
                        double a0 = -gamma * gn0;   // the intercept is one more coordinate: S0 / Y0 come from icept[]
#pragma unroll
                        for (int s = 0; s < MT; s++) {
                            const double cu = dense[DNF::cu + s], cw = gamma * dense[DNF::cw + s];
#pragma unroll
                            for (int k = 0; k < EPT; k++) {
                                acc[k] = fma(-cu, Sh[s][k], acc[k]);
                                acc[k] = fma(cw, Yh[s][k], acc[k]);
                            }
                            a0 = fma(-cu, icept[s], a0);
                            a0 = fma(cw, icept[MT + s], a0);
                        }
                        d0 = E.hi ? a0 : 0.0;
                        double gdp = 0.0;
#pragma unroll
                        for (int k = 0; k < EPT; k++) {
                            const uint32_t c = tid + (uint32_t)k * G;
                            dd[k] = acc[k];
                            if (c < E.d) gold[c] = gn[k];
                            gdp = fma(gn[k], acc[k], gdp);
                        }
                        g0 = gn0;
                        if (tid == 0) gdp = fma(g0, d0, gdp);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) gdp += __shfl_xor_sync(kFull, gdp, o);
                        if (lane == 0) wred[warp] = gdp;
                        gd_pending = true;
                        stp = 1.0;
                        fast_trial<G, EPT>(fa, smem, E, stp, dd);
                        phase = PH_FIRST;
                    }
                    continue;
                }
            }
            // ---- the search along d failed (or could not start): x, g, f are those of the last iterate -------
            f = fold;
            if (lb.col == 0) { status = GDMIX_SOLVE_ABNORMAL; iter++; break; }
            group_sync<G>();
            lbfgs_reset<G, MT>(lb, dense);
#pragma unroll
            for (int k = 0; k < EPT; k++) {
#pragma unroll
                for (int s = 0; s < MT; s++) { Sh[s][k] = 0.0; Yh[s][k] = 0.0; }
            }
            if (tid < 2u * MT) icept[tid] = 0.0;
            steepest_search();
        }
        if (emitted) continue;

        // ---- emit ---------------------------------------------------------------------------------------
        const int64_t t0 = a.b.theta_ptr[e] + (int64_t)li * a.sweep_coef_stride;
        const int64_t eo = (int64_t)e + (int64_t)li * a.b.n_entities;
        const double thr = a.o.sparsity_threshold;
#pragma unroll
        for (int k = 0; k < EPT; k++) {
            const uint32_t c = tid + (uint32_t)k * G;
            if (c < E.d) { const double xv = xcur[c]; a.theta_out[t0 + E.hi + c] = (thr > 0.0 && fabs(xv) <= thr) ? 0.0 : xv; }
        }
        if (tid == 0) {
            if (E.hi) a.theta_out[t0] = (thr > 0.0 && fabs(x0) <= thr) ? 0.0 : x0;
            if (a.f_out) a.f_out[eo] = f;
            if (a.nit) a.nit[eo] = iter;
            if (a.nfev) a.nfev[eo] = nfev;
            if (a.status) a.status[eo] = status;
        }
        if (a.var_out && a.o.variance_mode == GDMIX_VARIANCE_SIMPLE) {
            // var_j = 1 / (sum_i x_ij^2 rho_i (1-rho_i) w_i + l2 [j regularised] + 1e-12)
            // (binary_logistic_regression.py:171-177), at the un-thresholded optimum
            group_sync<G>();
            double *xt = (double *)(smem + L.xt), *rr = (double *)(smem + L.r);
#pragma unroll
            for (int k = 0; k < EPT; k++) {
                const uint32_t c = tid + (uint32_t)k * G;
                if (c < E.d) xt[c] = xcur[c];
            }
            group_sync<G>();
            const uint32_t *rbase = (const uint32_t *)(smem + L.rbase), *cbase = (const uint32_t *)(smem + L.cbase);
            double dsum[1] = {0.0};
            for (uint32_t b = warp; b < ((E.n + 31u) >> 5); b += W) {
                const uint32_t sb = rbase[b];
                double z = sell_dot((const float4 *)(smem + L.sell_val) + sb * kStepQuads + lane,
                                    (const uint2 *)(smem + L.sell_idx) + sb * kStepQuads + lane, rbase[b + 1] - sb);
                const uint32_t sp = b * 32u + lane;
                if (sp < E.n) {
                    z = (z + (E.hi ? x0 : 0.0)) + (double)((const float *)(smem + L.soff))[sp];
                    const double rho = 1.0 / (1.0 + exp(-z));
                    const double di = rho * (1.0 - rho) * (double)((const float *)(smem + L.sw))[sp];
                    rr[((const uint16_t *)(smem + L.rowperm))[sp]] = di;
                    dsum[0] += di;
                }
            }
            group_sum<G, 1>(dsum, red, flip);
            if (G == 32) __syncwarp();
            for (uint32_t b = warp; b < ((E.d + 31u) >> 5); b += W) {
                const uint32_t sb = cbase[b], ns = cbase[b + 1] - sb;
                const float4 *pv = (const float4 *)(smem + L.sell_val) + sb * kStepQuads + lane;
                const uint2 *pi = (const uint2 *)(smem + L.sell_idx) + sb * kStepQuads + lane;
                double h = 0.0;
                for (uint32_t k = 0; k < ns; k++) {
                    const float4 va = pv[k * kStepQuads];
                    const uint2 ca = pi[k * kStepQuads];
                    const double a0 = (double)va.x, a1 = (double)va.y, a2 = (double)va.z, a3 = (double)va.w;
                    h = fma(a0, a0 * lds_f64(ca.x & 0xffffu), h);
                    h = fma(a1, a1 * lds_f64(ca.x >> 16), h);
                    h = fma(a2, a2 * lds_f64(ca.y & 0xffffu), h);
                    h = fma(a3, a3 * lds_f64(ca.y >> 16), h);
                }
                const uint32_t sp = b * 32u + lane;
                if (sp < E.d)
                    a.var_out[t0 + E.hi + ((const uint16_t *)(smem + L.colperm))[sp]] = 1.0 / ((h + l2) + 1.0e-12);
            }
            if (E.hi && tid == 0)
                a.var_out[t0] = 1.0 / ((dsum[0] + (a.o.regularize_bias ? l2 : 0.0)) + 1.0e-12);
        }
        }   // models of the sweep
    }
#ifdef GDMIX_FAST_TIMING
    if (tid == 0)
        for (int k = 0; k < 11; k++) atomicAdd(&g_fast_cycles[k], (unsigned long long)ft_acc[k]);
#endif
}

}  // namespace gdmix
