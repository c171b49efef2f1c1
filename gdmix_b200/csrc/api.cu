// api.cu -- the extern "C" boundary declared in include/gdmix_b200.h: argument checks, launch
// planning (threads per entity, shared memory per CTA, persistent grid), the host-buffer
// pipeline (pinned staging, chunked H2D -> solve -> D2H on two streams) and the partition map.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/gdmix_b200.h"
#include "aux_kernels.cuh"
#include "host_lbfgs.h"
#include "fe_lbfgs.cuh"
#include "fe_plan.cuh"
#include "fe_tile.cuh"
#include "seqex_parser.h"
#include "seqex_writer.h"
#include "avro_writer.h"
#include "re_fast.cuh"
#include "re_kernel.cuh"
#include "re_small.cuh"
#include "re_variance.cuh"
#include "partition.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
thread_local int32_t g_last_plan[16] = {0};
thread_local int32_t g_last_small[8] = {0};

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                               \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return fail(GDMIX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                                   \
    } while (0)

struct DeviceInfo {
    int sm_count = 0, smem_optin = 0, cc = 0;
    bool ok = false;
};

int device_info(DeviceInfo &d)
{
    static std::mutex mu;
    static DeviceInfo cached[64];
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 64 && cached[dev].ok) { d = cached[dev]; return GDMIX_OK; }
    int major = 0, minor = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    d.cc = major * 10 + minor;
    d.ok = true;
    if (dev < 64) cached[dev] = d;
    return GDMIX_OK;
}

// Launch plan for one batch shape.
struct RePlan {
    int G = 128;                 // threads per entity (= CTA size)
    int MT = 10;                 // compile-time bound on the number of curvature pairs (10 or 32)
    uint32_t smem = 0;           // dynamic shared memory per CTA
    int ctas_per_sm = 1;
    int grid = 1;
    int hist_global = 0;         // history lives in the global arena (L2) instead of shared memory
    unsigned long long arena_stride = 0;  // per-CTA history spill in global memory
    // fast kernel (re_fast.cuh): entities that fit its sliced-ELL staging; the rest is deferred to the
    // general kernel above through a list in the workspace
    int fast = 0;
    int fG = 128, fEPT = 2;
    int fctas_per_sm = 1;
    int fgrid = 1;
    gdmix::FastLayout fL;
    // ragged batches: a first fast launch planned for the TYPICAL entity (2.5 x the mean sample count) at the
    // residency that shape allows; what it defers takes the launch above (planned for the largest entity)
    int fast0 = 0;
    int f0G = 128, f0EPT = 1, f0ctas_per_sm = 1, f0grid = 1;
    gdmix::FastLayout f0L;
    size_t off_defer0 = 0;
    size_t off_defer = 0, off_arena = 0;
    // entities that cannot be staged at all: X stays in global memory (re_solver_kernel<256, MT, true>)
    int bgrid = 1;
    uint32_t bsmem = 0;
    unsigned long long barena_stride = 0;
    size_t off_defer_b = 0, off_barena = 0;
    // ... and the ones with so many samples that one CTA would be the launch's tail: a cluster of kGiantCluster CTAs each
    int ggrid = 0, giant_rows = 0;
    size_t off_giant = 0, off_garena = 0;
    // small entities: one warp per entity (re_small.cuh), ahead of everything else; what does not fit its slices is
    // deferred to the launches above through a list
    int small = 0, sSL = 1, swarps = 8, sctas_per_sm = 1, sgrid = 1;
    gdmix::SmallShape sS;
    size_t off_defer_s = 0;
    // FULL variance pass (re_variance.cuh)
    int vgrid = 0;
    uint32_t vsmem = 0, vsmem_matrix_doubles = 0;
    unsigned long long vscratch_stride = 0;  // doubles per CTA in the workspace (0: the matrix fits on chip)
    size_t off_vscratch = 0;
    size_t workspace = 0;
};

constexpr size_t kQueueBytes = 256;
constexpr int kGiantCluster = 8;          // CTAs per cluster of the giant-entity launch (portable maximum); 16 where the device takes it
constexpr int kGiantRowsDefault = 32768;  // samples from which an unstageable entity is solved by a cluster
constexpr uint32_t kStaticSmem = 1024;  // upper bound on the kernels' static shared memory

int choose_group(const gdmix_re_batch *b, const gdmix_lr_opts *o)
{
    if (o->threads_per_entity == 32 || o->threads_per_entity == 64 || o->threads_per_entity == 128 ||
        o->threads_per_entity == 256)
        return o->threads_per_entity;
    // Rows and coefficients are the two thread-parallel axes; one thread per unit of the larger
    // one, clamped to [32, 256].
    const int span = std::max(b->max_rows, b->max_coef);
    if (span <= 40) return 32;
    if (span <= 96) return 64;
    if (span <= 512) return 128;
    return 256;
}

// ---- fast path planning ----------------------------------------------------------------------------
template <int G, int EPT>
int fast_regs(int &regs)
{
    cudaFuncAttributes at;
    CUDA_TRY(cudaFuncGetAttributes(&at, gdmix::re_fast_kernel<G, EPT>));
    regs = at.numRegs;
    return GDMIX_OK;
}

int fast_regs_of(int G, int EPT, int &regs)
{
    switch (G * 4 + EPT) {
    case 32 * 4 + 1: return fast_regs<32, 1>(regs);
    case 32 * 4 + 2: return fast_regs<32, 2>(regs);
    case 64 * 4 + 1: return fast_regs<64, 1>(regs);
    case 64 * 4 + 2: return fast_regs<64, 2>(regs);
    case 128 * 4 + 1: return fast_regs<128, 1>(regs);
    case 128 * 4 + 2: return fast_regs<128, 2>(regs);
    case 256 * 4 + 1: return fast_regs<256, 1>(regs);
    default: return fast_regs<256, 2>(regs);
    }
}

// Decides whether the batch goes through re_fast_kernel and with what geometry.  Shared memory per CTA is
// whatever the chosen residency leaves, so the sliced-ELL capacity (cap_steps) is as large as it can be;
// residency is the largest for which an entity of the batch's maximal shape with evenly spread non-zeros
// fits.  Entities that still do not fit are deferred to the general kernel one by one.
struct FastShapePlan {
    int ok = 0, G = 128, EPT = 1, ctas_per_sm = 1, grid = 1;
    gdmix::FastLayout L;
};

// Geometry of one re_fast_kernel launch for entities of up to N_full rows / D_full features / nnz_full non-zeros.
// `must_fit`: give up (ok = 0) unless an entity of that shape with evenly spread non-zeros fits at some residency;
// otherwise plan for two CTAs per SM and let whatever does not fit defer.
int plan_fast_shape(const gdmix_re_batch *b, const gdmix_lr_opts *o, const DeviceInfo &dev, uint32_t N_full,
                    uint32_t D_full, int64_t nnz_full, bool typical, FastShapePlan &fp)
{
    fp.ok = 0;
    const char *env_ctas = getenv("GDMIX_FAST_CTAS");
    const char *env_cap = getenv("GDMIX_FAST_CAP_STEPS");  // test hook: shrink the sliced-ELL capacity to force deferrals
    const uint32_t D = std::min(D_full, 512u), N = std::min(N_full, 4096u);
    bool clamped = D < D_full || N < N_full;
    int G;
    if (o->threads_per_entity == 32 || o->threads_per_entity == 64 || o->threads_per_entity == 128 ||
        o->threads_per_entity == 256) {
        G = o->threads_per_entity;
    } else {
        const uint32_t want = std::max((D + 1) / 2, (std::min(N, 512u) + 1) / 2);
        G = 32;
        while ((uint32_t)G < want && G < 256) G <<= 1;
    }
    if (D > 2u * (uint32_t)G) return GDMIX_OK;
    const int EPT = (D > (uint32_t)G) ? 2 : 1;
    int regs = 0;
    int rc = fast_regs_of(G, EPT, regs);
    if (rc) return rc;
    const int regs_alloc = (regs + 7) & ~7;
    const int k_hw = std::max(1, std::min({65536 / (G * regs_alloc), 2048 / G, 32}));
    const uint32_t W = (uint32_t)G / 32;
    gdmix::FastLayout fl;
    const uint32_t fixed = gdmix::fast_fixed_bytes(N, D, W, &fl);
    // the sliced index stream holds 16-bit absolute shared addresses into xt[] and r[]
    if (kStaticSmem + fl.r + 8u * N > 65536u) return GDMIX_OK;
    const uint32_t nrslab = (N + 31) / 32, ncslab = (D + 31) / 32;
    const uint32_t ar = (uint32_t)((nnz_full + (int64_t)N_full - 1) / N_full);
    const uint32_t ac = D_full ? (uint32_t)((nnz_full + (int64_t)D_full - 1) / D_full) : 0;
    // columns are never evenly filled: one more step per column slab for the typical-shape plan
    const uint32_t est = nrslab * ((ar + 3) / 4 + 1) + ncslab * ((ac + 3) / 4 + (typical ? 2 : 1));
    auto cap_for = [&](int k) {
        const int64_t budget = (int64_t)(228 * 1024) / k - 1024 - (int64_t)kStaticSmem - 64;
        return (std::min<int64_t>(budget, (int64_t)dev.smem_optin - kStaticSmem - 64) - (int64_t)fixed) /
               (int64_t)gdmix::kStepBytes;
    };
    int k = env_ctas ? std::max(1, std::min(atoi(env_ctas), k_hw)) : k_hw;
    int64_t cap = 0;
    bool found = false;
    if (!clamped) {
        for (; k >= 1; k--) {
            cap = cap_for(k);
            if (cap >= (int64_t)est) { found = true; break; }
            if (env_ctas) break;
        }
    }
    if (!found && !typical) {
        // the largest entity is not one for this kernel: plan for two CTAs per SM and let the big ones defer
        k = env_ctas ? std::max(1, std::min(atoi(env_ctas), k_hw)) : std::min(2, k_hw);
        cap = cap_for(k);
        found = cap >= 8;
    }
    if (!found) return GDMIX_OK;
    if (env_cap) cap = std::max<int64_t>(1, std::min<int64_t>(cap, atoi(env_cap)));
    fp.ok = 1;
    fp.G = G; fp.EPT = EPT; fp.ctas_per_sm = k;
    fp.L = gdmix::fast_layout(N, D, W, (uint32_t)cap);
    const int64_t want = (int64_t)dev.sm_count * k;
    fp.grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, b->n_entities));
    return GDMIX_OK;
}

// Decides whether the batch goes through re_fast_kernel and with what geometry.  Shared memory per CTA is
// whatever the chosen residency leaves, so the sliced-ELL capacity (cap_steps) is as large as it can be;
// residency is the largest for which an entity of the batch's maximal shape with evenly spread non-zeros
// fits.  Entities that still do not fit are deferred to the general kernel one by one.
// A ragged batch (largest entity well above the typical one) gets TWO launches: the first planned for 2.5 x the
// mean sample count -- the residency the many small entities deserve -- deferring the rest to the second.
int plan_fast(const gdmix_re_batch *b, const gdmix_lr_opts *o, const DeviceInfo &dev, RePlan &pl)
{
    pl.fast = 0; pl.fast0 = 0;
    const char *env_path = getenv("GDMIX_RE_PATH");
    const char *env_tiers = getenv("GDMIX_FAST_TIERS");    // test hook: "1" disables the typical-shape launch
    if (env_path && (strcmp(env_path, "generic") == 0 || strcmp(env_path, "big") == 0)) return GDMIX_OK;
    const uint32_t hi = o->has_intercept ? 1u : 0u;
    if (o->m > gdmix::kFastMT) return GDMIX_OK;
    const uint32_t D_full = (uint32_t)b->max_coef - hi, N_full = (uint32_t)b->max_rows;
    FastShapePlan big;
    int rc = plan_fast_shape(b, o, dev, N_full, D_full, b->max_nnz, false, big);
    if (rc) return rc;
    if (big.ok) {
        pl.fast = 1;
        pl.fG = big.G; pl.fEPT = big.EPT; pl.fctas_per_sm = big.ctas_per_sm; pl.fgrid = big.grid; pl.fL = big.L;
        if (b->n_rows > 0 && b->nnz > 0 && b->n_entities > 0 && !(env_tiers && atoi(env_tiers) == 1) &&
            o->threads_per_entity == 0) {
            const int64_t mean_rows = (b->n_rows + b->n_entities - 1) / b->n_entities;
            const char *env_f = getenv("GDMIX_TYPICAL_HALVES");   // tuning hook: typical shape = this many halves of the mean
            const int64_t halves = env_f ? std::max(2, atoi(env_f)) : 5;
            const uint32_t N_typ = (uint32_t)std::min<int64_t>(N_full, ((halves * mean_rows + 1) / 2 + 31) & ~(int64_t)31);
            const int64_t nnz_per_row = (b->nnz + b->n_rows - 1) / b->n_rows;
            if ((int64_t)N_typ * 3 <= (int64_t)N_full * 2) {
                FastShapePlan typ;
                rc = plan_fast_shape(b, o, dev, N_typ, D_full, (int64_t)N_typ * nnz_per_row, true, typ);
                if (rc) return rc;
                // worth a launch of its own when it puts more entities on an SM at a time
                if (typ.ok && typ.ctas_per_sm > big.ctas_per_sm) {
                    pl.fast0 = 1;
                    pl.f0G = typ.G; pl.f0EPT = typ.EPT; pl.f0ctas_per_sm = typ.ctas_per_sm; pl.f0grid = typ.grid;
                    pl.f0L = typ.L;
                }
            }
        }
        return GDMIX_OK;
    }
    if (env_path && strcmp(env_path, "fast") == 0)
        return fail(GDMIX_ERR_TOO_LARGE, "GDMIX_RE_PATH=fast but the batch shape (%u rows, %d nnz, %u features) "
                    "does not fit the fast kernel", N_full, b->max_nnz, D_full);
    return GDMIX_OK;
}

template <int SL>
int small_regs(int &regs)
{
    cudaFuncAttributes at;
    CUDA_TRY(cudaFuncGetAttributes(&at, gdmix::re_small_kernel<SL>));
    regs = at.numRegs;
    return GDMIX_OK;
}

// The warp-per-entity tier (re_small.cuh).  Slices are sized for twice the mean entity (all of it when the batch is
// uniform); whatever is larger defers to the CTA-per-entity kernels.  MEASURED (B200, per-user shape of configs[3],
// 32 samples x 64 features x 8 non-zeros, 4 M entities): 11.6 M entities/s against 12.6 M for re_fast_kernel<32,2> --
// the plain CSR / CSC walk executes 32 K warp instructions per entity and still misses the instruction cache
// (profiles/r2_ncu_full_re_small_v1.txt), so the planner does NOT pick it; GDMIX_RE_PATH=small selects it (parity
// tests run through it as a fifth path, and it is the shortest restatement of the solver on the device).
int plan_small(const gdmix_re_batch *b, const gdmix_lr_opts *o, const DeviceInfo &dev, RePlan &pl)
{
    pl.small = 0;
    const char *env_path = getenv("GDMIX_RE_PATH");
    const bool forced = env_path && strcmp(env_path, "small") == 0;
    if (!forced) return GDMIX_OK;
    if (o->m > 10 || b->n_entities <= 0 || b->n_rows <= 0) return GDMIX_OK;
    const int64_t mean_rows = (b->n_rows + b->n_entities - 1) / b->n_entities;
    const int64_t mean_nnz = (b->nnz + b->n_entities - 1) / b->n_entities;
    const uint32_t cap_coef = (uint32_t)std::min(b->max_coef, 96);
    const uint32_t cap_rows = (uint32_t)std::min<int64_t>(std::min<int64_t>(b->max_rows, 256), std::max<int64_t>(32, 2 * mean_rows));
    const uint32_t cap_nnz = (uint32_t)std::min<int64_t>(std::min<int64_t>(b->max_nnz, 2048), std::max<int64_t>(128, 2 * mean_nnz));
    pl.sSL = (int)((cap_coef + 31) / 32);
    pl.sS = gdmix::small_shape(std::max(cap_rows, 1u), std::max(cap_nnz, 4u), std::max(cap_coef, 1u), (uint32_t)o->m);
    int regs = 0;
    int rc = pl.sSL == 1 ? small_regs<1>(regs) : pl.sSL == 2 ? small_regs<2>(regs) : small_regs<3>(regs);
    if (rc) return rc;
    const int regs_alloc = (regs + 7) & ~7;
    const int by_regs = 65536 / (32 * regs_alloc), by_smem = (int)((228u * 1024u - 2048u) / pl.sS.bytes);
    const int warps_per_sm = std::min({by_regs, by_smem, 48});
    if (warps_per_sm < 1 || pl.sS.bytes > (uint32_t)dev.smem_optin - 1024u) return GDMIX_OK;
    int W = 8;
    while (W > 1 && (warps_per_sm / W < 1 || (size_t)W * pl.sS.bytes + 1024 > (size_t)dev.smem_optin)) W >>= 1;
    // as many whole CTAs as fit; prefer the CTA size that wastes the fewest warps
    int best_w = W, best_total = (warps_per_sm / W) * W;
    for (int w = W; w >= 1; w >>= 1) {
        const int total = (warps_per_sm / w) * w;
        if (total > best_total) { best_total = total; best_w = w; }
    }
    pl.swarps = best_w;
    pl.sctas_per_sm = std::max(1, warps_per_sm / best_w);
    const int64_t want = (int64_t)dev.sm_count * pl.sctas_per_sm;
    pl.sgrid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (b->n_entities + best_w - 1) / best_w));
    pl.small = 1;
    return GDMIX_OK;
}

int plan_re(const gdmix_re_batch *b, const gdmix_lr_opts *o, const DeviceInfo &dev, RePlan &pl)
{
    if (b->max_rows <= 0 || b->max_coef <= 0 || b->max_nnz < 0)
        return fail(GDMIX_ERR_INVALID, "gdmix_re_batch.max_rows/max_nnz/max_coef must be set (got %d/%d/%d)",
                    b->max_rows, b->max_nnz, b->max_coef);
    if (o->m < 0 || o->m > GDMIX_MAX_M) return fail(GDMIX_ERR_INVALID, "m = %d outside [0, %d]", o->m, GDMIX_MAX_M);
    const uint32_t hi = o->has_intercept ? 1u : 0u;
    if ((uint32_t)b->max_coef < hi) return fail(GDMIX_ERR_INVALID, "max_coef < has_intercept");
    pl.MT = (o->m <= 10) ? 10 : 32;
    const uint32_t budget = (uint32_t)dev.smem_optin - kStaticSmem;
    // The staged kernels are planned for the batch's largest entity when that fits on chip.  When it does not
    // (or its rows / features overflow the 16-bit on-chip indices), they are planned for half an SM's shared
    // memory and whatever entity does not fit is deferred, one by one, to the kernel that leaves X in global
    // memory -- one huge entity must not cost the other million their residency.
    gdmix::ReLayout L = gdmix::re_layout((uint32_t)std::min(b->max_rows, 65534), (uint32_t)b->max_nnz,
                                         (uint32_t)std::min<int64_t>(b->max_coef - (int)hi, 65534),
                                         (uint32_t)std::min(b->max_coef, 65534), (uint32_t)o->m, (uint32_t)pl.MT);
    const bool max_fits = b->max_rows < 65535 && b->max_coef - (int)hi < 65535 && L.fixed_bytes <= budget;
    if (!max_fits) {
        L.fixed_bytes = std::min<uint32_t>(budget, (228u * 1024u) / 2 - 1024u - kStaticSmem);
        L.total_bytes = 0xffffffffu;  // history never on chip in this case
    }
    {
        const char *env_path = getenv("GDMIX_RE_PATH");  // test hook: "big" sends every entity to the global-X kernel
        if (env_path && strcmp(env_path, "big") == 0) { L.fixed_bytes = 0; L.total_bytes = 0xffffffffu; }
    }
    pl.G = choose_group(b, o);
    // Where the (S, Y) history lives: on chip when that does not cost residency, else in a per-CTA global
    // arena that stays L2-resident (444 CTAs x 41 KB at the C1 shape).  GDMIX_HIST_GLOBAL=0/1 overrides.
    static const char *env_hist = getenv("GDMIX_HIST_GLOBAL");
    auto ctas_for = [&](uint32_t smem) {
        const int by_smem = (int)((228u * 1024u) / (smem + kStaticSmem + 1024u));
        return std::max(1, std::min({by_smem, 2048 / pl.G, 32, 65536 / (pl.G * 168)}));
    };
    pl.G = choose_group(b, o);
    bool hist_on_chip = L.total_bytes <= budget;
    if (hist_on_chip && ctas_for(L.fixed_bytes) > ctas_for(L.total_bytes)) hist_on_chip = false;
    if (env_hist) hist_on_chip = (L.total_bytes <= budget) && atoi(env_hist) == 0;
    pl.hist_global = hist_on_chip ? 0 : 1;
    pl.smem = hist_on_chip ? L.total_bytes : L.fixed_bytes;
    pl.arena_stride = hist_on_chip ? 0ull : (unsigned long long)gdmix::align16(16u * o->m * b->max_coef);
    pl.ctas_per_sm = ctas_for(pl.smem);
    const int64_t want = (int64_t)dev.sm_count * pl.ctas_per_sm;
    pl.grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, b->n_entities));
    int rc = plan_fast(b, o, dev, pl);
    if (rc) return rc;
    rc = plan_small(b, o, dev, pl);
    if (rc) return rc;
    const size_t list_bytes = ((size_t)b->n_entities * 4 + 255) & ~(size_t)255;
    pl.off_defer_s = kQueueBytes;
    pl.off_defer0 = pl.off_defer_s + list_bytes;
    pl.off_defer = pl.off_defer0 + (pl.fast ? list_bytes : 0);   // reserved whether or not the typical-shape launch is planned
    pl.off_defer_b = pl.off_defer + (pl.fast ? list_bytes : 0);
    pl.off_arena = pl.off_defer_b + list_bytes;
    pl.off_barena = pl.off_arena + (size_t)pl.arena_stride * (size_t)want;
    // global-X kernel: 256 threads, vectors + per-warp gradient copies on chip, history in its own arena
    {
        uint32_t need = gdmix::big_layout_bytes((uint32_t)b->max_coef, (uint32_t)b->max_coef - hi, 8u, (uint32_t)pl.MT);
        // room for several private gradient copies per warp (short rows are walked several to a warp step, one
        // copy per team of lanes) as long as four CTAs still fit an SM
        for (uint32_t copies = 32; copies > 1; copies >>= 1) {
            const uint32_t with = gdmix::big_layout_bytes((uint32_t)b->max_coef, (uint32_t)b->max_coef - hi, 8u,
                                                          (uint32_t)pl.MT, copies);
            if (with <= 48u * 1024u) { need = std::max(need, with); break; }
        }
        pl.bsmem = std::min(need, budget);   // entities with more coefficients than fit get GDMIX_ERR_TOO_LARGE
        const int per_sm = std::max(1, std::min(4, (int)((228u * 1024u) / (pl.bsmem + kStaticSmem + 1024u))));
        pl.bgrid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)dev.sm_count * per_sm, b->n_entities));
        pl.barena_stride = (unsigned long long)gdmix::align16(16u * (uint32_t)o->m * (uint32_t)b->max_coef);
    }
    pl.workspace = pl.off_barena + (size_t)pl.barena_stride * (size_t)pl.bgrid;
    {
        // giant entities: two clusters' worth of CTAs per eight SMs, same shared memory and arena stride
        const char *env_giant = getenv("GDMIX_GIANT_ROWS");   // test hook: 1 sends every unstageable entity there
        pl.giant_rows = env_giant ? std::max(1, atoi(env_giant)) : kGiantRowsDefault;
        pl.ggrid = 16 * std::max(1, dev.sm_count * 2 / 16);   // a multiple of both cluster sizes
        pl.off_giant = (pl.workspace + 255) & ~(size_t)255;
        pl.off_garena = pl.off_giant + list_bytes;
        pl.workspace = pl.off_garena + (size_t)pl.barena_stride * (size_t)pl.ggrid;
    }
    if (o->variance_mode == GDMIX_VARIANCE_FULL) {
        const size_t P = (size_t)b->max_coef;
        const size_t vec_bytes = 2 * 8 * P;
        const size_t avail = (size_t)dev.smem_optin - 1024;
        if (vec_bytes > avail) return fail(GDMIX_ERR_TOO_LARGE, "%zu coefficients are too many for the variance pass", P);
        pl.vsmem_matrix_doubles = (vec_bytes + 8 * P * P <= avail) ? (uint32_t)(P * P) : 0u;
        pl.vsmem = (uint32_t)(vec_bytes + 8 * (size_t)pl.vsmem_matrix_doubles);
        pl.vscratch_stride = pl.vsmem_matrix_doubles ? 0ull : (unsigned long long)(P * P);
        const int per_sm = pl.vsmem_matrix_doubles ? std::max(1, (int)((228u * 1024u) / (pl.vsmem + 2048u))) : 2;
        pl.vgrid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)dev.sm_count * std::min(per_sm, 8), b->n_entities));
        pl.off_vscratch = (pl.workspace + 255) & ~(size_t)255;
        pl.workspace = pl.off_vscratch + 8 * (size_t)pl.vscratch_stride * (size_t)pl.vgrid;
    }
    return GDMIX_OK;
}

template <int G, int MT>
int launch_re_t(const gdmix::ReArgs &args, const RePlan &pl, cudaStream_t st)
{
    static std::atomic<uint32_t> configured{0};
    if (configured.load() < pl.smem) {
        CUDA_TRY(cudaFuncSetAttribute(gdmix::re_solver_kernel<G, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      227 * 1024 - (int)kStaticSmem));
        configured.store(227 * 1024);
    }
    gdmix::re_solver_kernel<G, MT><<<pl.grid, G, pl.smem, st>>>(args);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}


template <int G, int EPT>
int launch_fast_t(const gdmix::FastArgs &fa, int grid, cudaStream_t st)
{
    static std::atomic<int> configured{0};
    if (!configured.load()) {
        CUDA_TRY(cudaFuncSetAttribute(gdmix::re_fast_kernel<G, EPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      227 * 1024 - (int)kStaticSmem));
        configured.store(1);
    }
    gdmix::re_fast_kernel<G, EPT><<<grid, G, fa.L.total_bytes, st>>>(fa);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}

// The same kernel as a thread-block cluster: kGiantCluster CTAs share one entity (re_kernel.cuh).
template <int MT>
int launch_giant_t(const gdmix::ReArgs &args, const RePlan &pl, cudaStream_t st)
{
    static std::atomic<int> configured{0};
    if (!configured.load()) {
        CUDA_TRY(cudaFuncSetAttribute(gdmix::re_solver_kernel<256, MT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      227 * 1024 - (int)kStaticSmem));
        configured.store(1);
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)pl.ggrid, 1, 1);   // a multiple of 16
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = pl.bsmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kGiantCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // sixteen CTAs per entity (non-portable cluster size) where a GPC holds such clusters with this launch's shared
    // memory; asked per launch, because the shared memory per CTA follows the batch
    static std::atomic<int> nonportable{-1};
    if (nonportable.load() < 0) {
        const bool ok = cudaFuncSetAttribute(gdmix::re_solver_kernel<256, MT, true>,
                                             cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
        cudaGetLastError();
        nonportable.store(ok ? 1 : 0);
    }
    if (nonportable.load() == 1) {
        attr[0].val.clusterDim.x = 16;
        int nclusters = 0;
        const bool fits = cudaOccupancyMaxActiveClusters(&nclusters, gdmix::re_solver_kernel<256, MT, true>, &cfg) ==
                              cudaSuccess && nclusters >= 4;
        cudaGetLastError();
        if (fits && cudaLaunchKernelEx(&cfg, gdmix::re_solver_kernel<256, MT, true>, args) == cudaSuccess) {
            g_launches++;
            return GDMIX_OK;
        }
        cudaGetLastError();   // fall back to the portable size
        attr[0].val.clusterDim.x = kGiantCluster;
    }
    CUDA_TRY(cudaLaunchKernelEx(&cfg, gdmix::re_solver_kernel<256, MT, true>, args));
    g_launches++;
    return GDMIX_OK;
}

template <int MT>
int launch_big_t(const gdmix::ReArgs &args, const RePlan &pl, cudaStream_t st)
{
    static std::atomic<int> configured{0};
    if (!configured.load()) {
        CUDA_TRY(cudaFuncSetAttribute(gdmix::re_solver_kernel<256, MT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      227 * 1024 - (int)kStaticSmem));
        configured.store(1);
    }
    gdmix::re_solver_kernel<256, MT, true><<<pl.bgrid, 256, pl.bsmem, st>>>(args);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}

int launch_fast(const gdmix::FastArgs &fa, int G, int EPT, int grid, cudaStream_t st)
{
    switch (G * 4 + EPT) {
    case 32 * 4 + 1: return launch_fast_t<32, 1>(fa, grid, st);
    case 32 * 4 + 2: return launch_fast_t<32, 2>(fa, grid, st);
    case 64 * 4 + 1: return launch_fast_t<64, 1>(fa, grid, st);
    case 64 * 4 + 2: return launch_fast_t<64, 2>(fa, grid, st);
    case 128 * 4 + 1: return launch_fast_t<128, 1>(fa, grid, st);
    case 128 * 4 + 2: return launch_fast_t<128, 2>(fa, grid, st);
    case 256 * 4 + 1: return launch_fast_t<256, 1>(fa, grid, st);
    default: return launch_fast_t<256, 2>(fa, grid, st);
    }
}

template <int SL>
int launch_small_t(const gdmix::SmallArgs &sa, int grid, cudaStream_t st)
{
    static std::atomic<int> configured{0};
    if (!configured.load()) {
        CUDA_TRY(cudaFuncSetAttribute(gdmix::re_small_kernel<SL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      227 * 1024 - (int)kStaticSmem));
        configured.store(1);
    }
    gdmix::re_small_kernel<SL><<<grid, 32 * sa.warps, (size_t)sa.warps * sa.S.bytes, st>>>(sa);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}

int launch_re(const gdmix_re_batch *b, const gdmix_lr_opts *o, int mode, const double *theta_in, double *theta_out,
              double *f_out, int32_t *nit, int32_t *nfev, int32_t *status, double *var_out, double *g_out,
              void *workspace, size_t workspace_bytes, cudaStream_t st, const double *l2_values = nullptr,
              int n_l2 = 0, int64_t coef_stride = 0)
{
    if (!b || !o) return fail(GDMIX_ERR_INVALID, "null batch/opts");
    if (b->n_entities < 0) return fail(GDMIX_ERR_INVALID, "n_entities < 0");
    if (b->n_entities == 0) return GDMIX_OK;
    if (!b->ent_rowptr || !b->rowptr || !b->label || !b->theta_ptr || (b->nnz > 0 && (!b->col || !b->val)))
        return fail(GDMIX_ERR_INVALID, "null array in gdmix_re_batch (device entry points need the int32 col)");
    DeviceInfo dev;
    int rc = device_info(dev);
    if (rc) return rc;
    if (dev.cc < 100) return fail(GDMIX_ERR_NO_DEVICE, "device compute capability %d < 100 (sm_100a build)", dev.cc);
    RePlan pl;
    rc = plan_re(b, o, dev, pl);
    if (rc) return rc;
    if (!workspace || workspace_bytes < pl.workspace)
        return fail(GDMIX_ERR_WORKSPACE, "workspace %zu B < required %zu B", workspace_bytes, pl.workspace);

    const bool full_var = mode == gdmix::kModeFit && var_out && o->variance_mode == GDMIX_VARIANCE_FULL;
    if (full_var && !status)
        return fail(GDMIX_ERR_INVALID, "variance_mode FULL needs the status array (rejected entities are skipped)");
    gdmix::ReArgs a;
    memset(&a, 0, sizeof(a));
    a.b = *b; a.o = *o;
    if (full_var) a.o.sparsity_threshold = 0.0;  // thresholded by the variance pass, after it has used theta
    a.theta_in = theta_in; a.theta_out = theta_out; a.f_out = f_out; a.nit = nit; a.nfev = nfev; a.status = status;
    a.var_out = var_out; a.g_out = g_out;
    a.queue = (int32_t *)workspace;
    a.arena = (unsigned char *)workspace + pl.off_arena;
    a.arena_stride = pl.arena_stride;
    a.mode = mode;
    a.hist_global = pl.hist_global;
    a.smem_bytes = pl.smem;
    a.n_l2 = n_l2;
    a.sweep_coef_stride = coef_stride;
    for (int j = 0; j < n_l2; j++) a.l2_sweep[j] = l2_values[j];
    CUDA_TRY(cudaMemsetAsync(workspace, 0, kQueueBytes, st));
    g_last_plan[0] = pl.fast; g_last_plan[1] = pl.fast ? pl.fG : pl.G; g_last_plan[2] = pl.fast ? pl.fEPT : 0;
    g_last_plan[3] = pl.fast ? pl.fctas_per_sm : pl.ctas_per_sm;
    g_last_plan[4] = pl.fast ? (int32_t)pl.fL.cap_steps : 0;
    g_last_plan[5] = pl.fast ? (int32_t)pl.fL.total_bytes : (int32_t)pl.smem;
    g_last_plan[6] = pl.fast ? pl.fgrid : pl.grid;
    g_last_plan[7] = pl.hist_global;
    g_last_plan[8] = pl.fast0; g_last_plan[9] = pl.f0G; g_last_plan[10] = pl.f0EPT; g_last_plan[11] = pl.f0ctas_per_sm;
    g_last_plan[12] = pl.fast0 ? (int32_t)pl.f0L.cap_steps : 0; g_last_plan[13] = pl.fast0 ? (int32_t)pl.f0L.total_bytes : 0;
    g_last_plan[14] = pl.fast0 ? (int32_t)pl.f0L.max_n : 0; g_last_plan[15] = 0;
    // tier 0: a warp per entity for batches of small entities (fit of one model, no variance output); the entities it
    // defers are what every later launch works on.  Counters: [10] its work counter, [11] its list's length.
    const bool use_small = pl.small && mode == gdmix::kModeFit && n_l2 == 0 && !var_out && theta_out;
    memset(g_last_small, 0, sizeof(g_last_small));
    if (use_small) {
        gdmix::SmallArgs sa;
        sa.a = a;
        sa.a.queue = (int32_t *)workspace + 10;
        sa.a.defer_list = (int32_t *)((unsigned char *)workspace + pl.off_defer_s);
        sa.a.defer_count = (int32_t *)workspace + 11;
        sa.S = pl.sS;
        sa.warps = pl.swarps;
        rc = pl.sSL == 1 ? launch_small_t<1>(sa, pl.sgrid, st) : pl.sSL == 2 ? launch_small_t<2>(sa, pl.sgrid, st)
                                                                              : launch_small_t<3>(sa, pl.sgrid, st);
        if (rc) return rc;
        a.todo = sa.a.defer_list;
        a.todo_count = sa.a.defer_count;
        g_last_small[0] = 1; g_last_small[1] = pl.swarps; g_last_small[2] = pl.sSL; g_last_small[3] = pl.sctas_per_sm;
        g_last_small[4] = (int32_t)pl.sS.cap_rows; g_last_small[5] = (int32_t)pl.sS.cap_nnz;
        g_last_small[6] = (int32_t)(pl.swarps * pl.sS.bytes); g_last_small[7] = pl.sgrid;
    }
    if (pl.fast) {
        // fast kernel first; what it defers (entities whose sliced form does not fit) is drained by the
        // general kernel from the list, with its own work counter.  Work counters in the workspace:
        // [0] first launch, [1] general kernel, [2] list A length, [3] variance, [4] global-X kernel,
        // [5] list B length, [6] second fast launch, [7] list A0 length.
        gdmix::FastArgs fa;
        fa.a = a;
        if (pl.fast0) {
            // ragged batch: typical-shape launch over everything, deferring to list A0 ...
            fa.L = pl.f0L;
            fa.defer_list = (int32_t *)((unsigned char *)workspace + pl.off_defer0);
            fa.defer_count = (int32_t *)workspace + 7;
            rc = launch_fast(fa, pl.f0G, pl.f0EPT, pl.f0grid, st);
            if (rc) return rc;
            // ... which the launch planned for the largest entity drains
            fa.a.queue = (int32_t *)workspace + 6;
            fa.a.todo = fa.defer_list;
            fa.a.todo_count = fa.defer_count;
        }
        fa.L = pl.fL;
        fa.defer_list = (int32_t *)((unsigned char *)workspace + pl.off_defer);
        fa.defer_count = (int32_t *)workspace + 2;
        rc = launch_fast(fa, pl.fG, pl.fEPT, pl.fgrid, st);
        if (rc) return rc;
        a.queue = (int32_t *)workspace + 1;
        a.todo = fa.defer_list;
        a.todo_count = fa.defer_count;
    }
    // whatever the staged kernels cannot hold goes to list B
    a.defer_list = (int32_t *)((unsigned char *)workspace + pl.off_defer_b);
    a.defer_count = (int32_t *)workspace + 5;
    // ... the ones with very many samples to list C ([8] its queue, [9] its length)
    a.giant_list = (int32_t *)((unsigned char *)workspace + pl.off_giant);
    a.giant_count = (int32_t *)workspace + 9;
    a.giant_rows = pl.giant_rows;
    // The general kernels solve one model per launch: a sweep runs them once per weight over the same lists
    // (only what the fast kernel deferred, when there is one), with their work counters rewound in between.
    const int n_models = n_l2 > 0 ? n_l2 : 1;
    const gdmix::ReArgs a_first = a;
    for (int j = 0; j < n_models; j++) {
    a = a_first;
    a.n_l2 = 0;
    if (n_l2 > 0) {
        a.o.l2 = l2_values[j];
        a.theta_out = theta_out + (int64_t)j * coef_stride;
        if (f_out) a.f_out = f_out + (int64_t)j * b->n_entities;
        if (nit) a.nit = nit + (int64_t)j * b->n_entities;
        if (nfev) a.nfev = nfev + (int64_t)j * b->n_entities;
        if (status) a.status = status + (int64_t)j * b->n_entities;
        if (j > 0) {
            // counters: [0] / [1] general queue, [4] list-B queue, [5] list-B length ([2] = fast kernel's list length stays)
            CUDA_TRY(cudaMemsetAsync((int32_t *)workspace + (pl.fast ? 1 : 0), 0, 4, st));
            CUDA_TRY(cudaMemsetAsync((int32_t *)workspace + 4, 0, 8, st));
            CUDA_TRY(cudaMemsetAsync((int32_t *)workspace + 8, 0, 8, st));
        }
    }
    if (pl.MT == 10) {
        switch (pl.G) {
        case 32: rc = launch_re_t<32, 10>(a, pl, st); break;
        case 64: rc = launch_re_t<64, 10>(a, pl, st); break;
        case 128: rc = launch_re_t<128, 10>(a, pl, st); break;
        default: rc = launch_re_t<256, 10>(a, pl, st); break;
        }
    } else {
        switch (pl.G) {
        case 32: rc = launch_re_t<32, 32>(a, pl, st); break;
        case 64: rc = launch_re_t<64, 32>(a, pl, st); break;
        case 128: rc = launch_re_t<128, 32>(a, pl, st); break;
        default: rc = launch_re_t<256, 32>(a, pl, st); break;
        }
    }
    if (rc) return rc;
    {
        // list B: X stays in global memory
        gdmix::ReArgs g2 = a;
        g2.queue = (int32_t *)workspace + 4;
        g2.todo = a.defer_list;
        g2.todo_count = a.defer_count;
        g2.defer_list = nullptr; g2.defer_count = nullptr;
        g2.giant_list = nullptr; g2.giant_count = nullptr;
        g2.arena = (unsigned char *)workspace + pl.off_barena;
        g2.arena_stride = pl.barena_stride;
        g2.hist_global = 1;
        g2.smem_bytes = pl.bsmem;
        rc = (pl.MT == 10) ? launch_big_t<10>(g2, pl, st) : launch_big_t<32>(g2, pl, st);
        if (rc) return rc;
        // list C: a cluster of CTAs per entity
        gdmix::ReArgs g3 = g2;
        g3.queue = (int32_t *)workspace + 8;
        g3.todo = a.giant_list;
        g3.todo_count = a.giant_count;
        g3.arena = (unsigned char *)workspace + pl.off_garena;
        rc = (pl.MT == 10) ? launch_giant_t<10>(g3, pl, st) : launch_giant_t<32>(g3, pl, st);
    }
    if (rc) return rc;
    }   // models of the sweep
    if (rc || !full_var) return rc;
    // FULL variance at the un-thresholded optimum, then the threshold (re_variance.cuh)
    gdmix::VarArgs v;
    memset(&v, 0, sizeof(v));
    v.b = *b; v.o = *o;
    v.theta = theta_out; v.var_out = var_out; v.status = status;
    v.queue = (int32_t *)workspace + 3;
    v.scratch = (double *)((unsigned char *)workspace + pl.off_vscratch);
    v.scratch_stride = pl.vscratch_stride;
    v.smem_matrix_doubles = pl.vsmem_matrix_doubles;
    v.max_coef = (uint32_t)b->max_coef;
    static std::atomic<int> vconfigured{0};
    if (!vconfigured.load()) {
        CUDA_TRY(cudaFuncSetAttribute(gdmix::re_variance_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      227 * 1024 - 1024));
        vconfigured.store(1);
    }
    gdmix::re_variance_full_kernel<<<pl.vgrid, 256, pl.vsmem, st>>>(v);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}

// ---- host-buffer pipeline ----------------------------------------------------------------------
struct Slot {
    cudaStream_t st = nullptr;
    void *dev = nullptr; size_t dev_bytes = 0;
    void *pin = nullptr; size_t pin_bytes = 0;
    void *ws = nullptr; size_t ws_bytes = 0;
};
struct HostCtx {
    std::mutex mu;
    Slot slot[3];
    int device = -1;
} g_host;

int ensure(Slot &s, size_t dev_bytes, size_t pin_bytes, size_t ws_bytes)
{
    if (!s.st) CUDA_TRY(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
    if (s.dev_bytes < dev_bytes) {
        if (s.dev) cudaFree(s.dev);
        s.dev = nullptr; s.dev_bytes = 0;
        CUDA_TRY(cudaMalloc(&s.dev, dev_bytes + dev_bytes / 4));
        s.dev_bytes = dev_bytes + dev_bytes / 4;
    }
    if (s.pin_bytes < pin_bytes) {
        if (s.pin) cudaFreeHost(s.pin);
        s.pin = nullptr; s.pin_bytes = 0;
        CUDA_TRY(cudaMallocHost(&s.pin, pin_bytes + pin_bytes / 4));
        s.pin_bytes = pin_bytes + pin_bytes / 4;
    }
    if (s.ws_bytes < ws_bytes) {
        if (s.ws) cudaFree(s.ws);
        s.ws = nullptr; s.ws_bytes = 0;
        CUDA_TRY(cudaMalloc(&s.ws, ws_bytes));
        s.ws_bytes = ws_bytes;
    }
    return GDMIX_OK;
}

inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

// Carves a chunk's device image: inputs first, outputs after.  Returns total bytes.
struct ChunkImage {
    size_t ent_rowptr, rowptr, col, col16, val, label, weight, offset, theta_ptr, theta0;  // inputs
    size_t len16, len32, scan_ws, label_bits;   // narrow row lengths / label bits and what rebuilds them
    size_t in_bytes;
    size_t theta, f, nit, nfev, status, var;  // outputs
    size_t total;
};

ChunkImage chunk_image(int64_t ne, int64_t nr, int64_t nz, int64_t nt, bool has_w, bool has_off, bool warm,
                       bool want_var, int narrow_col = 0 /* bytes per index crossing PCIe: 0 (= 4), 2 or 1 */,
                       bool narrow_rows = false, bool bit_labels = false)
{
    ChunkImage c;
    size_t o = 0;
    c.ent_rowptr = o; o += up256(8 * (ne + 1));
    c.rowptr = o; o += up256(8 * (nr + 1));
    c.theta_ptr = o; o += up256(8 * (ne + 1));
    c.col = o; o += up256(4 * nz);
    c.col16 = o; o += narrow_col ? up256(narrow_col * nz) : 0;
    c.val = o; o += up256(4 * nz);
    c.label = o; o += up256(4 * nr);
    c.weight = o; o += has_w ? up256(4 * nr) : 0;
    c.offset = o; o += has_off ? up256(4 * nr) : 0;
    c.theta0 = o; o += warm ? up256(8 * nt) : 0;
    c.len16 = o; o += narrow_rows ? up256(2 * nr) : 0;
    c.len32 = o; o += narrow_rows ? up256(4 * nr) : 0;
    // scan workspace: u32 tile sums + int64 tile offsets + total, tiles of 4096 rows
    c.scan_ws = o; o += narrow_rows ? up256(4 * (nr / 4096 + 2)) + up256(8 * (nr / 4096 + 3)) + 256 : 0;
    c.label_bits = o; o += bit_labels ? up256(nr / 8 + 2) : 0;
    c.in_bytes = o;
    c.theta = o; o += up256(8 * nt);
    c.f = o; o += up256(8 * ne);
    c.nit = o; o += up256(4 * ne);
    c.nfev = o; o += up256(4 * ne);
    c.status = o; o += up256(4 * ne);
    c.var = o; o += want_var ? up256(8 * nt) : 0;
    c.total = o;
    return c;
}

}  // namespace

extern "C" {

const char *gdmix_last_error(void) { return g_err; }
const char *gdmix_version(void) { return "gdmix_b200 0.1 (sm_100a)"; }
int64_t gdmix_launch_count(void) { return g_launches.load(); }

int gdmix_device_info(int32_t *sm_count, int32_t *smem_per_block_optin, int32_t *cc)
{
    DeviceInfo d;
    int rc = device_info(d);
    if (rc) return rc;
    if (sm_count) *sm_count = d.sm_count;
    if (smem_per_block_optin) *smem_per_block_optin = d.smem_optin;
    if (cc) *cc = d.cc;
    return GDMIX_OK;
}

void gdmix_re_last_plan(int32_t *out8)
{
    if (out8) memcpy(out8, g_last_plan, 8 * sizeof(int32_t));
}

void gdmix_re_last_plan_typical(int32_t *out8)
{
    if (out8) memcpy(out8, g_last_plan + 8, 8 * sizeof(int32_t));
}

void gdmix_re_last_plan_small(int32_t *out8)
{
    if (out8) memcpy(out8, g_last_small, 8 * sizeof(int32_t));
}

int gdmix_re_workspace_size(const gdmix_re_batch *batch, const gdmix_lr_opts *opts, size_t *bytes)
{
    if (!batch || !opts || !bytes) return fail(GDMIX_ERR_INVALID, "null argument");
    DeviceInfo dev;
    int rc = device_info(dev);
    if (rc) return rc;
    RePlan pl;
    rc = plan_re(batch, opts, dev, pl);
    if (rc) return rc;
    *bytes = pl.workspace;
    return GDMIX_OK;
}

int gdmix_re_loss_grad(const gdmix_re_batch *batch, const gdmix_lr_opts *opts, const double *theta, double *f,
                       double *g, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!theta || !f || !g) return fail(GDMIX_ERR_INVALID, "null theta/f/g");
    return launch_re(batch, opts, gdmix::kModeLossGrad, theta, nullptr, f, nullptr, nullptr, nullptr, nullptr, g,
                     workspace, workspace_bytes, (cudaStream_t)stream);
}

int gdmix_re_fit(const gdmix_re_batch *batch, const gdmix_lr_opts *opts, const double *theta0, double *theta_out,
                 double *f_out, int32_t *nit, int32_t *nfev, int32_t *status, double *var_out, void *workspace,
                 size_t workspace_bytes, void *stream)
{
    if (!theta_out) return fail(GDMIX_ERR_INVALID, "null theta_out");
    if (var_out && opts && opts->variance_mode == GDMIX_VARIANCE_NONE)
        return fail(GDMIX_ERR_INVALID, "var_out given but variance_mode is NONE");
    return launch_re(batch, opts, gdmix::kModeFit, theta0, theta_out, f_out, nit, nfev, status, var_out, nullptr,
                     workspace, workspace_bytes, (cudaStream_t)stream);
}

int gdmix_re_fit_sweep(const gdmix_re_batch *batch, const gdmix_lr_opts *opts, const double *l2_values, int32_t n_l2,
                       const double *theta0, double *theta_out, int64_t coef_stride, double *f_out, int32_t *nit,
                       int32_t *nfev, int32_t *status, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!theta_out || !l2_values) return fail(GDMIX_ERR_INVALID, "null theta_out / l2_values");
    if (n_l2 < 1 || n_l2 > GDMIX_MAX_SWEEP)
        return fail(GDMIX_ERR_INVALID, "n_l2 = %d outside [1, %d]", n_l2, GDMIX_MAX_SWEEP);
    if (coef_stride < 0) return fail(GDMIX_ERR_INVALID, "coef_stride < 0");
    for (int j = 0; j < n_l2; j++)
        if (!(l2_values[j] >= 0.0)) return fail(GDMIX_ERR_INVALID, "l2_values[%d] is negative or NaN", j);
    return launch_re(batch, opts, gdmix::kModeFit, theta0, theta_out, f_out, nit, nfev, status, nullptr, nullptr,
                     workspace, workspace_bytes, (cudaStream_t)stream, l2_values, n_l2, coef_stride);
}

int gdmix_re_score(const gdmix_re_batch *b, const gdmix_lr_opts *o, const double *theta, const uint8_t *has_model,
                   float *logit, float *logit_pc, void *stream)
{
    if (!b || !o || !logit || !logit_pc) return fail(GDMIX_ERR_INVALID, "null argument");
    if (b->n_entities <= 0) return GDMIX_OK;
    DeviceInfo dev;
    int rc = device_info(dev);
    if (rc) return rc;
    if (b->n_rows <= 0) return GDMIX_OK;
    const int grid = (int)std::min<int64_t>((b->n_rows + 255) / 256, (int64_t)dev.sm_count * 8);
    gdmix::re_score_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*b, o->has_intercept ? 1 : 0, theta, has_model,
                                                                  logit, logit_pc);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}

int gdmix_fe_loss_grad(const gdmix_fe_rows *rows, const gdmix_lr_opts *o, const double *x, double *fg, void *stream)
{
    if (!rows || !o || !x || !fg) return fail(GDMIX_ERR_INVALID, "null argument");
    DeviceInfo dev;
    int rc = device_info(dev);
    if (rc) return rc;
    const size_t len = (size_t)rows->n_features + (o->has_intercept ? 1 : 0) + 1;
    CUDA_TRY(cudaMemsetAsync(fg, 0, 8 * len, (cudaStream_t)stream));
    const int64_t work = std::max<int64_t>(rows->n_rows, rows->n_features + 1);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, (int64_t)dev.sm_count * 8));
    gdmix::fe_loss_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*rows, *o, x, fg);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}

namespace {
void fe_rows_geometry(const gdmix_fe_rows *rows, const DeviceInfo &dev, int &grid, int &team_shift)
{
    team_shift = 0;   // (unused since the rows kernel stages 32 rows per warp in shared memory)
    // a warp owns 32 rows at a time, 8 warps per CTA, one CTA per SM (x head + stages fill its shared memory)
    constexpr int T = gdmix::kFeRowsThreads;
    grid = (int)std::max<int64_t>(1, std::min<int64_t>((rows->n_rows + T - 1) / T, (int64_t)dev.sm_count));
}
}  // namespace

int gdmix_fe_column_counts(const int32_t *col, int64_t nnz, int64_t n_features, int64_t *counts, void *stream)
{
    if ((!col && nnz > 0) || !counts || nnz < 0 || n_features <= 0) return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_fe_column_counts");
    DeviceInfo dev;
    int rc = device_info(dev);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *bad = nullptr;
    CUDA_TRY(cudaMallocAsync((void **)&bad, 4, st));
    CUDA_TRY(cudaMemsetAsync(bad, 0, 4, st));
    CUDA_TRY(cudaMemsetAsync(counts, 0, 8 * (size_t)n_features, st));
    if (nnz > 0) {
        const int grid = (int)std::min<int64_t>((nnz + 255) / 256, (int64_t)dev.sm_count * 16);
        gdmix::fe_count_columns_kernel<<<grid, 256, 0, st>>>(col, nnz, n_features, (unsigned long long *)counts, bad);
        g_launches++;
    }
    int32_t hbad = 0;
    CUDA_TRY(cudaMemcpyAsync(&hbad, bad, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaFreeAsync(bad, st));
    if (hbad) return fail(GDMIX_ERR_INVALID, "feature index outside [0, %lld)", (long long)n_features);
    return GDMIX_OK;
}

int gdmix_remap_i32(const int32_t *in, const int32_t *map, int64_t n, int32_t *out, void *stream)
{
    if (n < 0 || (n > 0 && (!in || !map || !out))) return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_remap_i32");
    if (n == 0) return GDMIX_OK;
    DeviceInfo dev;
    int rc = device_info(dev);
    if (rc) return rc;
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)dev.sm_count * 16);
    gdmix::remap_i32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, map, n, out);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}

int gdmix_fe_hessian(const gdmix_fe_rows *rows, const gdmix_lr_opts *o, const double *x, int32_t mode, double *h,
                     void *stream)
{
    if (!rows || !o || !x || !h) return fail(GDMIX_ERR_INVALID, "null argument");
    if (mode != GDMIX_VARIANCE_SIMPLE && mode != GDMIX_VARIANCE_FULL)
        return fail(GDMIX_ERR_INVALID, "mode must be GDMIX_VARIANCE_SIMPLE or GDMIX_VARIANCE_FULL");
    if (rows->linear_regression) return fail(GDMIX_ERR_INVALID, "variance is defined for logistic regression only");
    DeviceInfo dev;
    int rc = device_info(dev);
    if (rc) return rc;
    const size_t P = (size_t)rows->n_features + (o->has_intercept ? 1 : 0);
    const int full = mode == GDMIX_VARIANCE_FULL;
    if (full && P > 16384) return fail(GDMIX_ERR_TOO_LARGE, "FULL variance needs a dense %zu x %zu matrix", P, P);
    CUDA_TRY(cudaMemsetAsync(h, 0, 8 * (full ? P * P : P), (cudaStream_t)stream));
    if (rows->n_rows > 0) {
        const int grid = (int)std::min<int64_t>((rows->n_rows + 255) / 256, (int64_t)dev.sm_count * 8);
        gdmix::fe_hessian_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*rows, o->has_intercept ? 1 : 0, x, full, h);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return GDMIX_OK;
}

int gdmix_fe_score(const gdmix_fe_rows *rows, const gdmix_lr_opts *o, const double *x, float *logit, float *logit_pc,
                   void *stream)
{
    if (!rows || !o || !x || !logit || !logit_pc) return fail(GDMIX_ERR_INVALID, "null argument");
    if (rows->n_rows <= 0) return GDMIX_OK;
    DeviceInfo dev;
    int rc = device_info(dev);
    if (rc) return rc;
    // the rows pass of the objective (rows staged per warp, leading coefficients of x in shared memory), writing logits
    int grid = 1, team_shift = 0;
    fe_rows_geometry(rows, dev, grid, team_shift);
    const uint32_t head = (uint32_t)std::min<int64_t>(rows->n_features, gdmix::kFeHeadMax);
    static std::atomic<int> configured{0};
    if (!configured.load()) {
        CUDA_TRY(cudaFuncSetAttribute(gdmix::fe_score_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)gdmix::fe_rows_smem_bytes(gdmix::kFeHeadMax)));
        configured.store(1);
    }
    gdmix::fe_score_rows_kernel<<<grid, gdmix::kFeRowsThreads, gdmix::fe_rows_smem_bytes(head), (cudaStream_t)stream>>>(
        *rows, *o, x, head, logit, logit_pc);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}

// ---- host-buffer entry points --------------------------------------------------------------------
int exclusive_scan_u32(const uint32_t *len, int64_t n, int64_t *out, void *ws, cudaStream_t st);   // defined below


int gdmix_re_fit_host(const gdmix_re_batch *hb, const gdmix_lr_opts *o, const double *theta0, double *theta_out,
                      double *f_out, int32_t *nit, int32_t *nfev, int32_t *status, double *var_out,
                      int64_t chunk_entities)
{
    if (!hb || !o || !theta_out) return fail(GDMIX_ERR_INVALID, "null argument");
    const int64_t E = hb->n_entities;
    if (E <= 0) return GDMIX_OK;
    if (!hb->ent_rowptr || !hb->rowptr || (!hb->label && !hb->label_bits) || !hb->theta_ptr)
        return fail(GDMIX_ERR_INVALID, "null array in gdmix_re_batch");
    if (hb->nnz > 0 && !hb->col && !hb->col16 && !hb->col8)
        return fail(GDMIX_ERR_INVALID, "gdmix_re_batch needs col, col16 or col8");
    const int narrow = hb->col8 ? 1 : hb->col16 ? 2 : 0;
    const bool want_var = var_out != nullptr;
    if (want_var && o->variance_mode != GDMIX_VARIANCE_SIMPLE && o->variance_mode != GDMIX_VARIANCE_FULL)
        return fail(GDMIX_ERR_INVALID, "var_out needs variance_mode SIMPLE or FULL");
    std::vector<int32_t> own_status;
    if (!status) { own_status.resize((size_t)E); status = own_status.data(); }
    std::lock_guard<std::mutex> lk(g_host.mu);

    // chunk boundaries: ~chunk_entities entities, or ~512 MB of input, whichever is smaller
    // enough chunks to keep three slots busy (H2D of chunk k+1 and D2H of chunk k-1 under the solve of chunk k),
    // each still large enough to fill the GPU
    if (chunk_entities <= 0) chunk_entities = std::max<int64_t>(2048, std::min<int64_t>(16384, (E + 11) / 12));
    std::vector<int64_t> cuts;
    cuts.push_back(0);
    {
        const size_t cap = (size_t)512 << 20;
        int64_t e = 0;
        while (e < E) {
            int64_t e1 = std::min(E, e + chunk_entities);
            auto bytes = [&](int64_t a, int64_t b2) {
                const int64_t r0 = hb->ent_rowptr[a], r1 = hb->ent_rowptr[b2];
                return (size_t)(8 * (hb->rowptr[r1] - hb->rowptr[r0]) + 20 * (r1 - r0));
            };
            while (e1 > e + 1 && bytes(e, e1) > cap) e1 = e + (e1 - e) / 2;
            cuts.push_back(e1);
            e = e1;
        }
    }
    const size_t nchunks = cuts.size() - 1;

    // Each chunk is copied straight out of the caller's arrays (pinned memory makes these copies truly
    // asynchronous; pageable memory still works, staged by the driver).  The pointer tables keep their
    // absolute values: the device-side array pointers are shifted instead, so nothing is rewritten.
    for (size_t ci = 0; ci < nchunks; ci++) {
        Slot &s = g_host.slot[ci % 3];
        if (s.st) CUDA_TRY(cudaStreamSynchronize(s.st));  // the slot's previous chunk is fully drained
        const int64_t e0 = cuts[ci], e1 = cuts[ci + 1], ne = e1 - e0;
        const int64_t r0 = hb->ent_rowptr[e0], r1 = hb->ent_rowptr[e1], nr = r1 - r0;
        const int64_t q0 = hb->rowptr[r0], q1 = hb->rowptr[r1], nz = q1 - q0;
        const int64_t t0 = hb->theta_ptr[e0], t1 = hb->theta_ptr[e1], nt = t1 - t0;
        gdmix_re_batch db = *hb;
        db.n_entities = ne; db.n_rows = nr; db.nnz = nz;
        int64_t mr = 0, mz = 0, mc = 0;
        for (int64_t e = e0; e < e1; e++) {
            const int64_t a = hb->ent_rowptr[e], b2 = hb->ent_rowptr[e + 1];
            mr = std::max<int64_t>(mr, b2 - a);
            mz = std::max<int64_t>(mz, hb->rowptr[b2] - hb->rowptr[a]);
            mc = std::max<int64_t>(mc, hb->theta_ptr[e + 1] - hb->theta_ptr[e]);
        }
        if (mr >= (1ll << 31) || mc >= (1ll << 31) || mz >= (1ll << 31))
            return fail(GDMIX_ERR_TOO_LARGE, "an entity has %lld rows / %lld nnz / %lld coefficients",
                        (long long)mr, (long long)mz, (long long)mc);
        db.max_rows = (int32_t)mr; db.max_nnz = (int32_t)mz; db.max_coef = (int32_t)mc;
        size_t ws_bytes = 0;
        int rc = gdmix_re_workspace_size(&db, o, &ws_bytes);
        if (rc) return rc;
        const bool narrow_rows = hb->row_len16 != nullptr, bit_labels = hb->label_bits != nullptr;
        const ChunkImage img = chunk_image(ne, nr, nz, nt, hb->weight != nullptr, hb->offset != nullptr,
                                           theta0 != nullptr, want_var, narrow, narrow_rows, bit_labels);
        rc = ensure(s, img.total, 0, ws_bytes);
        if (rc) return rc;
        char *dv = (char *)s.dev;
        const cudaMemcpyKind H2D = cudaMemcpyHostToDevice, D2H = cudaMemcpyDeviceToHost;
        // Every copy of the chunk first, then the small kernels that expand what crossed narrow: a kernel in between
        // would hold the copies behind it on this stream until it finds an SM free of the previous chunk's solve.
        CUDA_TRY(cudaMemcpyAsync(dv + img.ent_rowptr, hb->ent_rowptr + e0, 8 * (ne + 1), H2D, s.st));
        if (narrow_rows) CUDA_TRY(cudaMemcpyAsync(dv + img.len16, hb->row_len16 + r0, 2 * nr, H2D, s.st));
        else CUDA_TRY(cudaMemcpyAsync(dv + img.rowptr, hb->rowptr + r0, 8 * (nr + 1), H2D, s.st));
        CUDA_TRY(cudaMemcpyAsync(dv + img.theta_ptr, hb->theta_ptr + e0, 8 * (ne + 1), H2D, s.st));
        if (nz) {
            if (narrow == 1) CUDA_TRY(cudaMemcpyAsync(dv + img.col16, hb->col8 + q0, nz, H2D, s.st));
            else if (narrow == 2) CUDA_TRY(cudaMemcpyAsync(dv + img.col16, hb->col16 + q0, 2 * nz, H2D, s.st));
            else CUDA_TRY(cudaMemcpyAsync(dv + img.col, hb->col + q0, 4 * nz, H2D, s.st));
            CUDA_TRY(cudaMemcpyAsync(dv + img.val, hb->val + q0, 4 * nz, H2D, s.st));
        }
        const int64_t lb0 = r0 >> 3, lb1 = (r1 + 7) >> 3;
        if (bit_labels) CUDA_TRY(cudaMemcpyAsync(dv + img.label_bits, hb->label_bits + lb0, (size_t)(lb1 - lb0), H2D, s.st));
        else CUDA_TRY(cudaMemcpyAsync(dv + img.label, hb->label + r0, 4 * nr, H2D, s.st));
        if (hb->weight) CUDA_TRY(cudaMemcpyAsync(dv + img.weight, hb->weight + r0, 4 * nr, H2D, s.st));
        if (hb->offset) CUDA_TRY(cudaMemcpyAsync(dv + img.offset, hb->offset + r0, 4 * nr, H2D, s.st));
        if (theta0) CUDA_TRY(cudaMemcpyAsync(dv + img.theta0, theta0 + t0, 8 * nt, H2D, s.st));
        if (narrow_rows) {
            // 2 bytes per row crossed PCIe; the chunk's row pointers (counted from its own first non-zero) come from a scan
            gdmix::widen_len16_kernel<<<(int)std::min<int64_t>((nr + 255) / 256, 148 * 8), 256, 0, s.st>>>(
                (const uint16_t *)(dv + img.len16), (uint32_t *)(dv + img.len32), nr);
            g_launches++;
            rc = exclusive_scan_u32((const uint32_t *)(dv + img.len32), nr, (int64_t *)(dv + img.rowptr), dv + img.scan_ws, s.st);
            if (rc) return rc;
        }
        if (nz && narrow == 1) {
            gdmix::widen_u8_kernel<<<(int)std::min<int64_t>((nz / 16 + 255) / 256 + 1, 148 * 8), 256, 0, s.st>>>(
                (const uint8_t *)(dv + img.col16), (int32_t *)(dv + img.col), nz);
            g_launches++;
        } else if (nz && narrow == 2) {
            gdmix::widen_u16_kernel<<<(int)std::min<int64_t>((nz + 255) / 256, 148 * 8), 256, 0, s.st>>>(
                (const uint16_t *)(dv + img.col16), (int32_t *)(dv + img.col), nz);
            g_launches++;
        }
        if (bit_labels) {
            gdmix::label_from_bits_kernel<<<(int)std::min<int64_t>((nr + 255) / 256, 148 * 8), 256, 0, s.st>>>(
                (const uint8_t *)(dv + img.label_bits), (int)(r0 & 7), (float *)(dv + img.label), nr);
            g_launches++;
        }
        db.ent_rowptr = (const int64_t *)(dv + img.ent_rowptr);            // indexed by local entity
        db.theta_ptr = (const int64_t *)(dv + img.theta_ptr);
        db.rowptr = (const int64_t *)(dv + img.rowptr) - r0;               // indexed by absolute row
        db.label = (const float *)(dv + img.label) - r0;
        db.weight = hb->weight ? (const float *)(dv + img.weight) - r0 : nullptr;
        db.offset = hb->offset ? (const float *)(dv + img.offset) - r0 : nullptr;
        db.col = (const int32_t *)(dv + img.col) - (narrow_rows ? 0 : q0);  // indexed by the row pointers' non-zero numbers:
        db.val = (const float *)(dv + img.val) - (narrow_rows ? 0 : q0);    // absolute, or the chunk's own after a scan
        db.row_len16 = nullptr; db.label_bits = nullptr;
        rc = gdmix_re_fit(&db, o, theta0 ? (const double *)(dv + img.theta0) - t0 : nullptr,
                          (double *)(dv + img.theta) - t0, (double *)(dv + img.f), (int32_t *)(dv + img.nit),
                          (int32_t *)(dv + img.nfev), (int32_t *)(dv + img.status),
                          want_var ? (double *)(dv + img.var) - t0 : nullptr, s.ws, s.ws_bytes, s.st);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(theta_out + t0, dv + img.theta, 8 * nt, D2H, s.st));
        if (f_out) CUDA_TRY(cudaMemcpyAsync(f_out + e0, dv + img.f, 8 * ne, D2H, s.st));
        if (nit) CUDA_TRY(cudaMemcpyAsync(nit + e0, dv + img.nit, 4 * ne, D2H, s.st));
        if (nfev) CUDA_TRY(cudaMemcpyAsync(nfev + e0, dv + img.nfev, 4 * ne, D2H, s.st));
        CUDA_TRY(cudaMemcpyAsync(status + e0, dv + img.status, 4 * ne, D2H, s.st));
        if (want_var) CUDA_TRY(cudaMemcpyAsync(var_out + t0, dv + img.var, 8 * nt, D2H, s.st));
    }
    for (Slot &s : g_host.slot)
        if (s.st) CUDA_TRY(cudaStreamSynchronize(s.st));
    for (int64_t e = 0; e < E; e++)
        if (status[e] < 0)
            return fail(status[e], "entity %lld rejected by the device path (status %d)", (long long)e, status[e]);
    return GDMIX_OK;
}

int gdmix_host_register(void *ptr, size_t bytes)
{
    if (!ptr || !bytes) return fail(GDMIX_ERR_INVALID, "null buffer");
    CUDA_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return GDMIX_OK;
}

int gdmix_host_unregister(void *ptr)
{
    if (!ptr) return fail(GDMIX_ERR_INVALID, "null buffer");
    CUDA_TRY(cudaHostUnregister(ptr));
    return GDMIX_OK;
}

int gdmix_selftest_logistic(const double *z, int64_t n, double *out, void *stream)
{
    if (n < 0 || (n > 0 && (!z || !out))) return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_selftest_logistic");
    if (n == 0) return GDMIX_OK;
    gdmix::logistic_selftest_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(z, n, out);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}

int gdmix_pinned_alloc(size_t bytes, void **out)
{
    if (!out || !bytes) return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_pinned_alloc");
    *out = nullptr;
    CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return GDMIX_OK;
}

int gdmix_pinned_free(void *ptr)
{
    if (!ptr) return GDMIX_OK;
    CUDA_TRY(cudaFreeHost(ptr));
    return GDMIX_OK;
}

int gdmix_narrow_columns(const int32_t *col, int64_t n, int32_t width, void *out)
{
    if (n < 0 || (n > 0 && (!col || !out)) || (width != 1 && width != 2)) return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_narrow_columns");
    int bad = 0;
    if (width == 1) {
        uint8_t *o8 = (uint8_t *)out;
#pragma omp parallel for schedule(static) reduction(| : bad)
        for (int64_t i = 0; i < n; i++) { bad |= (col[i] < 0 || col[i] > 255); o8[i] = (uint8_t)col[i]; }
    } else {
        uint16_t *o16 = (uint16_t *)out;
#pragma omp parallel for schedule(static) reduction(| : bad)
        for (int64_t i = 0; i < n; i++) { bad |= (col[i] < 0 || col[i] > 65535); o16[i] = (uint16_t)col[i]; }
    }
    if (bad) return fail(GDMIX_ERR_INVALID, "column index does not fit %d byte(s)", (int)width);
    return GDMIX_OK;
}

int gdmix_re_score_host(const gdmix_re_batch *hb, const gdmix_lr_opts *o, const double *theta,
                        const uint8_t *has_model, float *logit, float *logit_pc)
{
    if (!hb || !o || !logit || !logit_pc) return fail(GDMIX_ERR_INVALID, "null argument");
    const int64_t E = hb->n_entities;
    if (E <= 0) return GDMIX_OK;
    std::lock_guard<std::mutex> lk(g_host.mu);
    const int64_t nr_all = hb->ent_rowptr[E], nz_all = hb->rowptr[nr_all];
    if (nz_all > 0 && ((!hb->col && !hb->col16 && !hb->col8) || !hb->val))
        return fail(GDMIX_ERR_INVALID, "gdmix_re_score_host needs col, col16 or col8, and val");
    const int narrow = hb->col8 ? 1 : hb->col16 ? 2 : 0;     // as gdmix_re_fit_host: narrow indices cross PCIe, widened here
    // Chunks of about 512 MB of input, alternating between two slots (the upload of a chunk runs under the scoring and
    // the download of the one before): a partition that trained in chunks scores in chunks.  The kernel indexes with the
    // batch's absolute row / non-zero / coefficient numbers, so a chunk's device arrays are handed to it shifted back
    // by the chunk's first row / non-zero / coefficient.
    const char *env_chunk = getenv("GDMIX_SCORE_CHUNK_BYTES");    // test hook
    const int64_t budget = env_chunk ? std::max<int64_t>(1, atoll(env_chunk)) : (512ll << 20);
    int64_t e0 = 0;
    int k = 0;
    while (e0 < E) {
        const int64_t r0 = hb->ent_rowptr[e0], q0 = hb->rowptr[r0], t0 = hb->theta_ptr[e0];
        int64_t e1 = e0 + 1;
        {
            // largest e1 with bytes(e0 .. e1) <= budget (at least one entity): bisection over the cumulative arrays
            auto bytes_to = [&](int64_t e) {
                const int64_t r = hb->ent_rowptr[e], q = hb->rowptr[r];
                return 8 * (q - q0) + 16 * (r - r0) + 8 * (hb->theta_ptr[e] - t0) + 17 * (e - e0);
            };
            int64_t lo = e0 + 1, hi2 = E;
            while (lo < hi2) {
                const int64_t mid = (lo + hi2 + 1) >> 1;
                if (bytes_to(mid) <= budget) lo = mid; else hi2 = mid - 1;
            }
            e1 = lo;
        }
        const int64_t Ec = e1 - e0, r1 = hb->ent_rowptr[e1], q1 = hb->rowptr[r1], t1 = hb->theta_ptr[e1];
        const int64_t nr = r1 - r0, nz = q1 - q0, nt = t1 - t0;
        size_t o_ent = 0, o_row = up256(8 * (Ec + 1)), o_tp = o_row + up256(8 * (nr + 1));
        size_t o_col = o_tp + up256(8 * (Ec + 1)), o_val = o_col + up256(4 * nz), o_off = o_val + up256(4 * nz);
        size_t o_th = o_off + up256(4 * nr), o_hm = o_th + up256(8 * nt), o_in = o_hm + up256(Ec);
        size_t o_lg = o_in, o_pc = o_lg + up256(4 * nr), o_nc = o_pc + up256(4 * nr);
        size_t total = o_nc + (narrow ? up256((size_t)narrow * nz) : 0);
        Slot &s = g_host.slot[k & 1];
        if (s.st) CUDA_TRY(cudaStreamSynchronize(s.st));     // the chunk that used this slot two turns ago is home
        int rc = ensure(s, total, 0, 0);
        if (rc) return rc;
        char *dv = (char *)s.dev;
        CUDA_TRY(cudaMemcpyAsync(dv + o_ent, hb->ent_rowptr + e0, 8 * (Ec + 1), cudaMemcpyHostToDevice, s.st));
        CUDA_TRY(cudaMemcpyAsync(dv + o_row, hb->rowptr + r0, 8 * (nr + 1), cudaMemcpyHostToDevice, s.st));
        CUDA_TRY(cudaMemcpyAsync(dv + o_tp, hb->theta_ptr + e0, 8 * (Ec + 1), cudaMemcpyHostToDevice, s.st));
        if (nz) {
            if (narrow == 1) CUDA_TRY(cudaMemcpyAsync(dv + o_nc, hb->col8 + q0, nz, cudaMemcpyHostToDevice, s.st));
            else if (narrow == 2) CUDA_TRY(cudaMemcpyAsync(dv + o_nc, hb->col16 + q0, 2 * nz, cudaMemcpyHostToDevice, s.st));
            else CUDA_TRY(cudaMemcpyAsync(dv + o_col, hb->col + q0, 4 * nz, cudaMemcpyHostToDevice, s.st));
            CUDA_TRY(cudaMemcpyAsync(dv + o_val, hb->val + q0, 4 * nz, cudaMemcpyHostToDevice, s.st));
        }
        if (hb->offset) CUDA_TRY(cudaMemcpyAsync(dv + o_off, hb->offset + r0, 4 * nr, cudaMemcpyHostToDevice, s.st));
        if (theta) CUDA_TRY(cudaMemcpyAsync(dv + o_th, theta + t0, 8 * nt, cudaMemcpyHostToDevice, s.st));
        if (has_model) CUDA_TRY(cudaMemcpyAsync(dv + o_hm, has_model + e0, Ec, cudaMemcpyHostToDevice, s.st));
        if (nz && narrow == 1) {           // after every copy of the chunk (see gdmix_re_fit_host)
            gdmix::widen_u8_kernel<<<(int)std::min<int64_t>((nz / 16 + 255) / 256 + 1, 148 * 8), 256, 0, s.st>>>(
                (const uint8_t *)(dv + o_nc), (int32_t *)(dv + o_col), nz);
            g_launches++;
        } else if (nz && narrow == 2) {
            gdmix::widen_u16_kernel<<<(int)std::min<int64_t>((nz + 255) / 256, 148 * 8), 256, 0, s.st>>>(
                (const uint16_t *)(dv + o_nc), (int32_t *)(dv + o_col), nz);
            g_launches++;
        }
        gdmix_re_batch db = *hb;
        db.n_entities = Ec;
        db.ent_rowptr = (const int64_t *)(dv + o_ent);
        db.rowptr = (const int64_t *)(dv + o_row) - r0;
        db.theta_ptr = (const int64_t *)(dv + o_tp);
        db.col = (const int32_t *)(dv + o_col) - q0;
        db.val = (const float *)(dv + o_val) - q0;
        db.label = nullptr; db.weight = nullptr;
        db.n_rows = nr; db.nnz = nz;
        db.offset = hb->offset ? (const float *)(dv + o_off) - r0 : nullptr;
        rc = gdmix_re_score(&db, o, theta ? (const double *)(dv + o_th) - t0 : nullptr,
                            has_model ? (const uint8_t *)(dv + o_hm) : nullptr, (float *)(dv + o_lg) - r0,
                            (float *)(dv + o_pc) - r0, s.st);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(logit + r0, dv + o_lg, 4 * nr, cudaMemcpyDeviceToHost, s.st));
        CUDA_TRY(cudaMemcpyAsync(logit_pc + r0, dv + o_pc, 4 * nr, cudaMemcpyDeviceToHost, s.st));
        e0 = e1;
        k++;
    }
    for (int j = 0; j < 2; j++)
        if (g_host.slot[j].st) CUDA_TRY(cudaStreamSynchronize(g_host.slot[j].st));
    return GDMIX_OK;
}

#ifdef GDMIX_FAST_TIMING
GDMIX_API int gdmix_debug_fast_cycles(unsigned long long *out, int reset)
{
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpyFromSymbol(out, gdmix::g_fast_cycles, 12 * sizeof(unsigned long long)));
    if (reset) {
        unsigned long long z[12] = {0};
        CUDA_TRY(cudaMemcpyToSymbol(gdmix::g_fast_cycles, z, sizeof(z)));
    }
    return GDMIX_OK;
}
#endif

// ---- partitioner / evaluator entry points (partition.cuh) ------------------------------------------------
namespace {
inline size_t up256z(size_t x) { return (x + 255) & ~(size_t)255; }
inline int64_t sort_tiles(int64_t n) { return (n + gdmix::kSortTile - 1) / gdmix::kSortTile; }

struct SortScratch {
    uint64_t *ktmp; uint32_t *vtmp; uint32_t *hist; uint64_t *digit_total, *digit_base;
    size_t bytes;
};
SortScratch carve_sort(void *ws, int64_t n)
{
    SortScratch c;
    const int64_t nt = sort_tiles(n);
    size_t o = 0;
    char *b = (char *)ws;
    c.ktmp = (uint64_t *)(b + o); o += up256z(8 * (size_t)n);
    c.vtmp = (uint32_t *)(b + o); o += up256z(4 * (size_t)n);
    c.hist = (uint32_t *)(b + o); o += up256z(4 * (size_t)nt * gdmix::kRadix);
    c.digit_total = (uint64_t *)(b + o); o += up256z(8 * gdmix::kRadix);
    c.digit_base = (uint64_t *)(b + o); o += up256z(8 * gdmix::kRadix);
    c.bytes = o;
    return c;
}

int sort_pairs(const uint64_t *keys_in, const uint32_t *vals_in, int64_t n, int key_bits, uint64_t *keys_out,
               uint32_t *vals_out, void *ws, cudaStream_t st)
{
    const SortScratch c = carve_sort(ws, n);
    const int passes = std::max(1, (key_bits + 7) / 8);
    const int grid = (int)sort_tiles(n);
    const uint64_t *kin = keys_in;
    const uint32_t *vin = vals_in;
    for (int p = 0; p < passes; p++) {
        // the last pass must land in the output pair
        const bool to_out = ((passes - 1 - p) % 2) == 0;
        uint64_t *ko = to_out ? keys_out : c.ktmp;
        uint32_t *vo = to_out ? vals_out : c.vtmp;
        gdmix::radix_hist_kernel<<<grid, gdmix::kSortThreads, 0, st>>>(kin, n, 8 * p, c.hist);
        gdmix::radix_scan_tiles_kernel<<<gdmix::kRadix, 256, 0, st>>>(c.hist, grid, c.digit_total);
        gdmix::radix_scan_digits_kernel<<<1, 32, 0, st>>>(c.digit_total, c.digit_base);
        gdmix::radix_scatter_kernel<<<grid, gdmix::kSortThreads, 0, st>>>(kin, vin, ko, vo, n, 8 * p, c.hist, c.digit_base);
        g_launches += 4;
        kin = ko; vin = vo;
    }
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}
}  // namespace

// ---- tiled fixed-effect objective: plan handle (fe_plan.cuh) + evaluation (fe_tile.cuh) ------------------------------
struct gdmix_fe_tile_plan {
    gdmix::FeTilePlan P{};
    std::vector<void *> owned;
    int64_t nnz = 0, n_hot_z = 0, n_cold_z = 0, n_hot_g = 0, n_cold_g = 0, bytes = 0;
};

namespace {
int plan_alloc_bytes(gdmix_fe_tile_plan *h, void **out, size_t bytes, bool keep)
{
    void *p = nullptr;
    bytes = std::max<size_t>(bytes, 256);
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(GDMIX_ERR_CUDA, "fixed-effect plan: cannot allocate %zu bytes of device memory", bytes);
    }
    *out = p;
    if (keep) { h->owned.push_back(p); h->bytes += (int64_t)bytes; }
    return GDMIX_OK;
}
#define plan_alloc(h, out, count, keep) plan_alloc_bytes((h), (void **)(out), sizeof(**(out)) * (size_t)(count), (keep))

// out[0..n] = exclusive scan of len[0..n) (64-bit), ws: >= 4 * tiles + 8 * tiles + 8 bytes
int exclusive_scan_u32(const uint32_t *len, int64_t n, int64_t *out, void *ws, cudaStream_t st)
{
    if (n <= 0) { CUDA_TRY(cudaMemsetAsync(out, 0, 8, st)); return GDMIX_OK; }
    const int64_t nt = sort_tiles(n);
    uint32_t *tile_sum = (uint32_t *)ws;
    int64_t *tile_off = (int64_t *)((char *)ws + up256z(4 * (size_t)nt));
    int64_t *total = tile_off + nt;
    gdmix::tile_sum_kernel<<<(int)nt, gdmix::kSortThreads, 0, st>>>(len, n, tile_sum);
    gdmix::tiles_exclusive_scan_kernel<<<1, 256, 0, st>>>(tile_sum, nt, tile_off, total);
    gdmix::rowptr_from_len_kernel<<<(int)nt, gdmix::kSortThreads, 0, st>>>(len, n, tile_off, out);
    g_launches += 3;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}
size_t scan_ws_bytes(int64_t n)
{
    const size_t nt = (size_t)sort_tiles(std::max<int64_t>(n, 1));
    return up256z(4 * nt) + up256z(8 * nt + 8) + 256;
}
int bits_for(uint64_t v)
{
    int b = 1;
    while (b < 64 && (v >> b)) b++;
    return b;
}
}  // namespace

void gdmix_fe_tile_plan_destroy(gdmix_fe_tile_plan *h)
{
    if (!h) return;
    for (void *p : h->owned) cudaFree(p);
    delete h;
}

gdmix_fe_tile_plan *gdmix_fe_tile_plan_create(const gdmix_fe_rows *rows, int32_t hz, int32_t hg, int32_t tile_rows,
                                              int64_t l2_tile_rows, void *stream)
{
    if (!rows || rows->n_rows < 0 || rows->nnz < 0 || rows->n_features <= 0 || !rows->rowptr ||
        (rows->nnz > 0 && (!rows->col || !rows->val))) {
        fail(GDMIX_ERR_INVALID, "bad rows in gdmix_fe_tile_plan_create");
        return nullptr;
    }
    if (rows->nnz >= (1ll << 32) - 4096 || rows->n_rows >= (1ll << 32) - 1) {
        fail(GDMIX_ERR_TOO_LARGE, "a planned shard holds fewer than 2^32 rows and non-zeros (split the shard)");
        return nullptr;
    }
    DeviceInfo dev;
    if (device_info(dev)) return nullptr;
    const int64_t D = rows->n_features, n = rows->n_rows, nnz = rows->nnz;
    // defaults: as many coefficients / accumulators in shared memory as the two kernels can hold
    const int64_t smem = dev.smem_optin;
    const int64_t hz_max = (smem - 1024 - (gdmix::kFeZThreads / 32) * (gdmix::kFeZStage * 6 + gdmix::kFeZColdStage * 8)) / 8;
    if (tile_rows <= 0) tile_rows = 8192;
    if (tile_rows > 65536) { fail(GDMIX_ERR_INVALID, "tile_rows must be at most 65536 (16-bit rows inside a tile)"); return nullptr; }
    const int64_t hg_max = std::min<int64_t>((smem - 4096 - 16 * (int64_t)gdmix::fe_g_tile_doubles((uint32_t)tile_rows)) / 8, 65536);
    if (hz <= 0) hz = (int32_t)std::min<int64_t>(12288, hz_max);
    if (hg <= 0) hg = (int32_t)std::min<int64_t>(24064, hg_max);
    hz = (int32_t)std::min<int64_t>(std::min<int64_t>(hz, D), std::min<int64_t>(hz_max, 65536));
    hg = (int32_t)std::min<int64_t>(std::min<int64_t>(hg, D), hg_max);
    if (hz < 0 || hg < 1) {
        fail(GDMIX_ERR_INVALID, "tile_rows = %d leaves no shared memory for the gradient accumulators (16 bytes per row of a tile, %lld bytes per CTA)",
             tile_rows, (long long)smem);
        return nullptr;
    }
    if (l2_tile_rows <= 0) l2_tile_rows = 4 << 20;
    cudaStream_t st = (cudaStream_t)stream;
    gdmix_fe_tile_plan *h = new (std::nothrow) gdmix_fe_tile_plan();
    if (!h) { fail(GDMIX_ERR_INVALID, "out of host memory"); return nullptr; }
    gdmix::FeTilePlan &P = h->P;
    P.n_rows = n; P.n_features = D; P.hz = hz; P.hg = hg; P.tile_rows = tile_rows; P.l2_tile_rows = l2_tile_rows;
    P.n_blocks = (n + 31) / 32;
    P.n_tiles = (n + tile_rows - 1) / tile_rows;
    P.n_l2_tiles = std::max<int64_t>(1, (n + l2_tile_rows - 1) / l2_tile_rows);
    P.n_cold = D - hg;
    P.z_grid = (int32_t)std::max<int64_t>(1, std::min<int64_t>((n + gdmix::kFeZThreads - 1) / gdmix::kFeZThreads, dev.sm_count));
    P.g_grid = (int32_t)std::min<int64_t>(P.n_tiles, dev.sm_count);
    h->nnz = nnz;
    std::vector<void *> temps;
    auto cleanup = [&](bool ok) {
        cudaStreamSynchronize(st);
        for (void *p : temps) cudaFree(p);
        if (!ok) { gdmix_fe_tile_plan_destroy(h); h = nullptr; }
    };
#define PLAN_TRY(expr) do { if ((expr) != GDMIX_OK) { cleanup(false); return nullptr; } } while (0)
#define PLAN_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { fail(GDMIX_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); cleanup(false); return nullptr; } } while (0)
    auto temp_bytes = [&](void **out, size_t bytes) {
        int rc = plan_alloc_bytes(h, out, bytes, false);
        if (rc == GDMIX_OK) temps.push_back(*out);
        return rc;
    };
#define temp(out, count) temp_bytes((void **)(out), sizeof(**(out)) * (size_t)(count))
    const int gsm = dev.sm_count * 16;
    auto grid_for = [&](int64_t items, int per) { return (int)std::max<int64_t>(1, std::min<int64_t>((items + per - 1) / per, gsm)); };
    // ---- scratch of the evaluation -------------------------------------------------------------------------------
    double *dz, *block_part, *acc_part, *cold_part;
    PLAN_TRY(plan_alloc(h, &dz, (size_t)n, true));
    PLAN_TRY(plan_alloc(h, &block_part, 2 * (size_t)P.z_grid, true));
    PLAN_TRY(plan_alloc(h, &acc_part, (size_t)std::max(P.g_grid, 1) * hg, true));
    PLAN_TRY(plan_alloc(h, &cold_part, (size_t)(P.n_l2_tiles * std::max<int64_t>(P.n_cold, 1)), true));
    P.dz = dz; P.block_part = block_part; P.acc_part = acc_part; P.cold_part = cold_part;
    // ---- z side ----------------------------------------------------------------------------------------------------
    uint16_t *zh_len, *zc_len; uint32_t *blk_cnt, *cblk_cnt; int64_t *zh_blk, *zc_blk; int32_t *err; void *scan_ws;
    PLAN_TRY(plan_alloc(h, &zh_len, (size_t)n, true));
    PLAN_TRY(plan_alloc(h, &zh_blk, (size_t)P.n_blocks + 1, true));
    PLAN_TRY(plan_alloc(h, &zc_len, (size_t)n, true));
    PLAN_TRY(plan_alloc(h, &zc_blk, (size_t)P.n_blocks + 1, true));
    PLAN_TRY(temp(&blk_cnt, (size_t)P.n_blocks));
    PLAN_TRY(temp(&cblk_cnt, (size_t)P.n_blocks));
    PLAN_TRY(temp(&err, 1));
    { char *w; PLAN_TRY(temp(&w, scan_ws_bytes(std::max<int64_t>(n, P.n_tiles)))); scan_ws = w; }
    PLAN_CUDA(cudaMemsetAsync(err, 0, 4, st));
    int64_t tot_hot_z = 0, tot_cold_z = 0;
    if (n > 0) {
        gdmix::fe_zcount_kernel<<<grid_for(n, 256), 256, 0, st>>>(rows->rowptr, rows->col, n, hz, zh_len, zc_len, err);
        gdmix::fe_zblock_kernel<<<grid_for(P.n_blocks, 256), 256, 0, st>>>(zh_len, n, P.n_blocks, 4u, 8u, blk_cnt);
        gdmix::fe_zblock_kernel<<<grid_for(P.n_blocks, 256), 256, 0, st>>>(zc_len, n, P.n_blocks, 1u, 4u, cblk_cnt);
        g_launches += 3;
    }
    PLAN_TRY(exclusive_scan_u32(blk_cnt, P.n_blocks, zh_blk, scan_ws, st));
    PLAN_TRY(exclusive_scan_u32(cblk_cnt, P.n_blocks, zc_blk, scan_ws, st));
    {
        int32_t herr = 0;
        PLAN_CUDA(cudaMemcpyAsync(&tot_hot_z, zh_blk + P.n_blocks, 8, cudaMemcpyDeviceToHost, st));
        PLAN_CUDA(cudaMemcpyAsync(&tot_cold_z, zc_blk + P.n_blocks, 8, cudaMemcpyDeviceToHost, st));
        PLAN_CUDA(cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, st));
        PLAN_CUDA(cudaStreamSynchronize(st));
        if (herr) { fail(GDMIX_ERR_TOO_LARGE, "a row has more than 262140 non-zeros among (or 65535 outside) the %d most frequent features", hz); cleanup(false); return nullptr; }
    }
    float *zh_val, *zc_val; uint16_t *zh_col; int32_t *zc_col;
    PLAN_TRY(plan_alloc(h, &zh_val, (size_t)tot_hot_z + 8, true));
    PLAN_TRY(plan_alloc(h, &zh_col, (size_t)tot_hot_z + 8, true));
    PLAN_TRY(plan_alloc(h, &zc_val, (size_t)tot_cold_z + 4, true));
    PLAN_TRY(plan_alloc(h, &zc_col, (size_t)tot_cold_z + 4, true));
    if (n > 0) {
        gdmix::fe_zscatter_kernel<<<grid_for(P.n_blocks, 8), 256, 0, st>>>(rows->rowptr, rows->col, rows->val, n, hz, zh_len,
                                                                          zc_len, zh_blk, zc_blk, zh_val, zh_col, zc_val, zc_col);
        g_launches++;
    }
    P.zh_blk = zh_blk; P.zh_len = zh_len; P.zh_val = zh_val; P.zh_col = zh_col;
    P.zc_blk = zc_blk; P.zc_len = zc_len; P.zc_val = zc_val; P.zc_col = zc_col;
    h->n_hot_z = tot_hot_z; h->n_cold_z = tot_cold_z;
    // ---- g side: one stable sort of all non-zeros by [class | tile | column] -------------------------------------------
    const uint64_t hot_span = (uint64_t)std::max<int64_t>(P.n_tiles, 1) * (uint64_t)hg;
    const uint64_t cold_span = (uint64_t)P.n_l2_tiles * (uint64_t)std::max<int64_t>(P.n_cold, 1);
    const int class_shift = bits_for(std::max(hot_span, cold_span));
    const uint64_t classbit = 1ull << class_shift;
    int64_t *tile_begin, *gt_ptr, *gc_run;
    PLAN_TRY(plan_alloc(h, &gt_ptr, (size_t)P.n_tiles + 1, true));
    PLAN_TRY(plan_alloc(h, &gc_run, (size_t)(P.n_l2_tiles * P.n_cold) + 1, true));
    PLAN_TRY(temp(&tile_begin, (size_t)P.n_tiles + 2));
    int64_t n_hot_g = 0, tot_hot_g = 0;
    uint32_t *row_of = nullptr, *perm = nullptr; uint64_t *keys = nullptr, *keys_sorted = nullptr;
    if (nnz > 0) {
        void *sort_ws;
        const size_t sort_bytes = carve_sort(nullptr, nnz).bytes + 256;   // exactly what the radix passes use
        PLAN_TRY(temp(&row_of, (size_t)nnz));
        PLAN_TRY(temp(&keys, (size_t)nnz));
        PLAN_TRY(temp(&keys_sorted, (size_t)nnz));
        PLAN_TRY(temp(&perm, (size_t)nnz));
        { char *w; PLAN_TRY(temp(&w, sort_bytes)); sort_ws = w; }
        gdmix::fe_expand_rows_kernel<<<grid_for(n, 256), 256, 0, st>>>(rows->rowptr, n, row_of);
        gdmix::fe_gkeys_kernel<<<grid_for(nnz, 256), 256, 0, st>>>(rows->col, row_of, nnz, hg, tile_rows, l2_tile_rows,
                                                                   std::max<int64_t>(P.n_cold, 1), classbit, keys);
        g_launches += 2;
        PLAN_TRY(sort_pairs(keys, nullptr, nnz, class_shift + 1, keys_sorted, perm, sort_ws, st));
    }
    // tile_begin[t] = first sorted entry of tile t (t = n_tiles: the number of hot entries)
    gdmix::fe_lower_bound_kernel<<<grid_for(P.n_tiles + 1, 256), 256, 0, st>>>(keys_sorted, nnz, 0, (uint64_t)hg, P.n_tiles + 1, 0,
                                                                              tile_begin);
    g_launches++;
    PLAN_CUDA(cudaMemcpyAsync(&n_hot_g, tile_begin + P.n_tiles, 8, cudaMemcpyDeviceToHost, st));
    {
        uint32_t *tcnt;
        PLAN_TRY(temp(&tcnt, (size_t)std::max<int64_t>(P.n_tiles, 1)));
        if (P.n_tiles > 0) {
            gdmix::fe_gtile_count_kernel<<<grid_for(P.n_tiles, 256), 256, 0, st>>>(tile_begin, P.n_tiles, tcnt);
            g_launches++;
        }
        PLAN_TRY(exclusive_scan_u32(tcnt, P.n_tiles, gt_ptr, scan_ws, st));
        PLAN_CUDA(cudaMemcpyAsync(&tot_hot_g, gt_ptr + P.n_tiles, 8, cudaMemcpyDeviceToHost, st));
        PLAN_CUDA(cudaStreamSynchronize(st));
    }
    const int64_t n_cold_g = nnz - n_hot_g;
    float *gh_val, *gc_val; uint16_t *gh_row, *gh_col; uint32_t *gc_row;
    PLAN_TRY(plan_alloc(h, &gh_val, (size_t)tot_hot_g + 8, true));
    PLAN_TRY(plan_alloc(h, &gh_row, (size_t)tot_hot_g + 8, true));
    PLAN_TRY(plan_alloc(h, &gh_col, (size_t)tot_hot_g + 8, true));
    PLAN_TRY(plan_alloc(h, &gc_val, (size_t)n_cold_g + 4, true));
    PLAN_TRY(plan_alloc(h, &gc_row, (size_t)n_cold_g + 4, true));
    if (n_hot_g > 0) {
        gdmix::fe_ghot_gather_kernel<<<grid_for(n_hot_g, 256), 256, 0, st>>>(keys_sorted, perm, n_hot_g, hg, tile_rows, tile_begin,
                                                                            gt_ptr, rows->val, row_of, gh_val, gh_row, gh_col);
        gdmix::fe_ghot_pad_kernel<<<grid_for(P.n_tiles, 8), 256, 0, st>>>(P.n_tiles, tile_begin, gt_ptr, gh_val, gh_row, gh_col);
        g_launches += 2;
    }
    if (P.n_cold > 0) {
        gdmix::fe_lower_bound_kernel<<<grid_for(P.n_l2_tiles * P.n_cold + 1, 256), 256, 0, st>>>(
            keys_sorted, nnz, classbit, 1, P.n_l2_tiles * P.n_cold + 1, n_hot_g, gc_run);
        g_launches++;
        if (n_cold_g > 0) {
            gdmix::fe_gcold_gather_kernel<<<grid_for(n_cold_g, 256), 256, 0, st>>>(perm + n_hot_g, n_cold_g, rows->val, row_of,
                                                                                  gc_val, gc_row);
            g_launches++;
        }
    } else {
        PLAN_CUDA(cudaMemsetAsync(gc_run, 0, 8, st));
    }
    P.gt_ptr = gt_ptr; P.gh_val = gh_val; P.gh_row = gh_row; P.gh_col = gh_col;
    P.gc_run = gc_run; P.gc_val = gc_val; P.gc_row = gc_row;
    h->n_hot_g = n_hot_g; h->n_cold_g = n_cold_g;
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { fail(GDMIX_ERR_CUDA, "fixed-effect plan kernels: %s", cudaGetErrorString(e)); cleanup(false); return nullptr; }
    }
    cleanup(true);
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { fail(GDMIX_ERR_CUDA, "fixed-effect plan: %s", cudaGetErrorString(e)); gdmix_fe_tile_plan_destroy(h); return nullptr; }
    }
#undef PLAN_TRY
#undef PLAN_CUDA
#undef temp
    return h;
}

int gdmix_fe_tile_plan_info(const gdmix_fe_tile_plan *h, int64_t *out8)
{
    if (!h || !out8) return fail(GDMIX_ERR_INVALID, "null argument");
    out8[0] = h->P.hz; out8[1] = h->P.hg; out8[2] = h->P.tile_rows; out8[3] = h->P.n_tiles;
    out8[4] = h->n_cold_z; out8[5] = h->n_cold_g; out8[6] = h->bytes; out8[7] = h->P.n_l2_tiles;
    return GDMIX_OK;
}

int gdmix_fe_loss_grad_tiled(const gdmix_fe_rows *rows, const gdmix_fe_tile_plan *h, const gdmix_lr_opts *o, const double *x,
                             double *fg, void *stream)
{
    if (!rows || !h || !o || !x || !fg) return fail(GDMIX_ERR_INVALID, "null argument");
    const gdmix::FeTilePlan &P = h->P;
    if (rows->n_rows != P.n_rows || rows->n_features != P.n_features || rows->nnz != h->nnz)
        return fail(GDMIX_ERR_INVALID, "the plan was built for another shard");
    if (P.n_rows > 0 && !rows->label) return fail(GDMIX_ERR_INVALID, "null label");
    DeviceInfo dev;
    int rc = device_info(dev);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    static std::atomic<int> configured{0};
    if (!configured.load()) {
        CUDA_TRY(cudaFuncSetAttribute(gdmix::fe_z_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dev.smem_optin - 1024));
        CUDA_TRY(cudaFuncSetAttribute(gdmix::fe_g_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dev.smem_optin - 4096));
        configured.store(1);
    }
    gdmix::fe_z_kernel<<<P.z_grid, gdmix::kFeZThreads, gdmix::fe_z_smem_bytes((uint32_t)P.hz), st>>>(*rows, *o, P, x);
    g_launches++;
    if (P.g_grid > 0) {
        gdmix::fe_g_kernel<<<P.g_grid, gdmix::kFeGThreads, gdmix::fe_g_smem_bytes((uint32_t)P.hg, (uint32_t)P.tile_rows), st>>>(P);
        g_launches++;
    }
    if (P.n_cold > 0) {
        const int cgrid = (int)std::max<int64_t>(1, std::min<int64_t>((P.n_cold + 7) / 8, (int64_t)dev.sm_count * 16));
        for (int64_t t = 0; t < P.n_l2_tiles; t++) {
            gdmix::fe_gcold_kernel<<<cgrid, 256, 0, st>>>(P, t);
            g_launches++;
        }
    }
    const int fgrid = (int)std::max<int64_t>(1, std::min<int64_t>((P.n_features + 255) / 256, (int64_t)dev.sm_count * 4));
    gdmix::fe_finish2_kernel<<<fgrid, 256, 0, st>>>(*rows, *o, P, x, fg);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}


int gdmix_partition_workspace_size(int64_t n, size_t *bytes)
{
    if (n < 0 || !bytes) return fail(GDMIX_ERR_INVALID, "bad argument");
    const size_t nt = (size_t)sort_tiles(std::max<int64_t>(n, 1));
    *bytes = 64 * (size_t)std::max<int64_t>(n, 1) + 16 * nt * gdmix::kRadix + (1u << 16);
    return GDMIX_OK;
}

int gdmix_sort_pairs_u64(const uint64_t *keys_in, int64_t n, int32_t key_bits, uint64_t *keys_out, uint32_t *perm_out,
                         void *ws, size_t ws_bytes, void *stream)
{
    if (n < 0 || key_bits < 1 || key_bits > 64 || (n > 0 && (!keys_in || !keys_out || !perm_out || !ws)))
        return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_sort_pairs_u64");
    if (n >= (1ll << 32)) return fail(GDMIX_ERR_TOO_LARGE, "at most 2^32 - 1 keys per call");
    if (n == 0) return GDMIX_OK;
    size_t need = 0;
    gdmix_partition_workspace_size(n, &need);
    if (ws_bytes < need) return fail(GDMIX_ERR_WORKSPACE, "workspace %zu B < required %zu B", ws_bytes, need);
    return sort_pairs(keys_in, nullptr, n, key_bits, keys_out, perm_out, ws, (cudaStream_t)stream);
}

int gdmix_group_by_key(const uint64_t *keys, int64_t n, int32_t key_bits, uint64_t *keys_sorted, uint32_t *perm,
                       int64_t *seg_ptr, uint64_t *seg_key, int64_t *n_groups_dev, void *ws, size_t ws_bytes, void *stream)
{
    if (!n_groups_dev || !seg_ptr) return fail(GDMIX_ERR_INVALID, "null output");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { CUDA_TRY(cudaMemsetAsync(n_groups_dev, 0, 8, st)); CUDA_TRY(cudaMemsetAsync(seg_ptr, 0, 8, st)); return GDMIX_OK; }
    int rc = gdmix_sort_pairs_u64(keys, n, key_bits, keys_sorted, perm, ws, ws_bytes, stream);
    if (rc) return rc;
    // segment heads: scratch after the sort's own
    const SortScratch c = carve_sort(ws, n);
    const int64_t nt = sort_tiles(n);
    uint32_t *tile_count = (uint32_t *)((char *)ws + c.bytes);
    int64_t *tile_off = (int64_t *)((char *)ws + c.bytes + up256z(4 * (size_t)nt));
    gdmix::heads_count_kernel<<<(int)nt, gdmix::kSortThreads, 0, st>>>(keys_sorted, n, tile_count);
    gdmix::tiles_exclusive_scan_kernel<<<1, 256, 0, st>>>(tile_count, nt, tile_off, n_groups_dev);
    gdmix::heads_write_kernel<<<(int)nt, gdmix::kSortThreads, 0, st>>>(keys_sorted, n, tile_off, seg_ptr, seg_key);
    g_launches += 3;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}

int gdmix_csr_gather_rows(const int64_t *rowptr_in, const int32_t *col_in, const float *val_in, const uint32_t *perm,
                          int64_t n_rows, int64_t *rowptr_out, int32_t *col_out, float *val_out, void *ws,
                          size_t ws_bytes, void *stream)
{
    if (n_rows < 0 || !rowptr_out) return fail(GDMIX_ERR_INVALID, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rows == 0) { CUDA_TRY(cudaMemsetAsync(rowptr_out, 0, 8, st)); return GDMIX_OK; }
    if (!rowptr_in || !perm || !ws) return fail(GDMIX_ERR_INVALID, "null argument");
    size_t need = 0;
    gdmix_partition_workspace_size(n_rows, &need);
    if (ws_bytes < need) return fail(GDMIX_ERR_WORKSPACE, "workspace %zu B < required %zu B", ws_bytes, need);
    const int64_t nt = sort_tiles(n_rows);
    uint32_t *len = (uint32_t *)ws;
    uint32_t *tile_sum = (uint32_t *)((char *)ws + up256z(4 * (size_t)n_rows));
    int64_t *tile_off = (int64_t *)((char *)tile_sum + up256z(4 * (size_t)nt));
    int64_t *total = tile_off + nt;
    DeviceInfo dev;
    int rc = device_info(dev);
    if (rc) return rc;
    const int g1 = (int)std::min<int64_t>((n_rows + 255) / 256, (int64_t)dev.sm_count * 8);
    gdmix::gather_row_len_kernel<<<g1, 256, 0, st>>>(rowptr_in, perm, n_rows, len);
    gdmix::tile_sum_kernel<<<(int)nt, gdmix::kSortThreads, 0, st>>>(len, n_rows, tile_sum);
    gdmix::tiles_exclusive_scan_kernel<<<1, 256, 0, st>>>(tile_sum, nt, tile_off, total);
    gdmix::rowptr_from_len_kernel<<<(int)nt, gdmix::kSortThreads, 0, st>>>(len, n_rows, tile_off, rowptr_out);
    if (col_in && val_in && col_out && val_out) {
        const int g2 = (int)std::min<int64_t>((n_rows + 7) / 8, (int64_t)dev.sm_count * 16);
        gdmix::gather_rows_kernel<<<g2, 256, 0, st>>>(rowptr_in, col_in, val_in, perm, n_rows, rowptr_out, col_out, val_out);
        g_launches++;
    }
    g_launches += 4;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}

int gdmix_gather_f32(const float *in, const uint32_t *perm, int64_t n, float *out, void *stream)
{
    if (n < 0 || (n > 0 && (!in || !perm || !out))) return fail(GDMIX_ERR_INVALID, "bad argument");
    if (n == 0) return GDMIX_OK;
    gdmix::gather_f32_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(in, perm, n, out);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}

int gdmix_group_ids(const int64_t *seg_ptr, int64_t n_groups, const uint32_t *perm, const int64_t *uid, int64_t n,
                    int32_t lower_bound, int32_t upper_bound, int32_t *group_id, void *stream)
{
    if (n < 0 || n_groups < 0 || (n > 0 && (!seg_ptr || !perm || !uid || !group_id || n_groups < 1)))
        return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_group_ids");
    if (n == 0) return GDMIX_OK;
    DeviceInfo dev;
    int rc = device_info(dev);
    if (rc) return rc;
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)dev.sm_count * 16);
    gdmix::group_ids_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(seg_ptr, n_groups, perm, uid, n, lower_bound, upper_bound,
                                                                    group_id);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}

int gdmix_offset_join(const int64_t *uid, int64_t n, const uint64_t *score_uid_sorted, const uint32_t *score_perm, int64_t m,
                      const float *score, const float *per_coordinate, float *offset_out, uint8_t *matched, void *stream)
{
    if (n < 0 || m < 0 || (n > 0 && (!uid || !offset_out || !matched)) || (m > 0 && (!score_uid_sorted || !score_perm || !score)))
        return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_offset_join");
    if (n == 0) return GDMIX_OK;
    DeviceInfo dev;
    int rc = device_info(dev);
    if (rc) return rc;
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)dev.sm_count * 16);
    gdmix::offset_join_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(uid, n, score_uid_sorted, score_perm, m, score,
                                                                      per_coordinate, offset_out, matched);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}

int gdmix_local_index_host(const int64_t *ent_rowptr, const int64_t *rowptr, const int64_t *gcol, int64_t n_entities,
                           int32_t *local_col, int64_t *d_e, int64_t *scratch, const int64_t *uniq_ptr, int64_t *uniq_global)
{
    if (n_entities < 0 || !ent_rowptr || !rowptr || !d_e || !scratch) return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_local_index_host");
    if (!uniq_global) {
        if (!local_col && n_entities > 0) return fail(GDMIX_ERR_INVALID, "null local_col");
        if (gdmix_host::local_index_pass1(ent_rowptr, rowptr, gcol, n_entities, local_col, d_e, scratch) != 0)
            return fail(GDMIX_ERR_INVALID, "negative feature index");
        return GDMIX_OK;
    }
    if (!uniq_ptr) return fail(GDMIX_ERR_INVALID, "null uniq_ptr");
    gdmix_host::local_index_pass2(ent_rowptr, rowptr, n_entities, d_e, uniq_ptr, scratch, uniq_global);
    return GDMIX_OK;
}

int gdmix_seqex_encode(const gdmix_seqex_spec *spec, int64_t n_entities, const int64_t *ent_rows, const int64_t *entity_int,
                       const char *id_chars, const int64_t *id_ptr, const int64_t *row_len, const int64_t *gcol,
                       const float *val, const int64_t *uid, const float *label, int32_t label_as_int, const float *offset,
                       const float *weight, uint8_t *out, int64_t capacity, int64_t *written)
{
    if (!spec || n_entities < 0 || !written || (n_entities > 0 && !ent_rows))
        return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_seqex_encode");
    if (spec->entity && !entity_int && !(id_chars && id_ptr)) return fail(GDMIX_ERR_INVALID, "entity column without ids");
    if (spec->uid && !uid && n_entities > 0) return fail(GDMIX_ERR_INVALID, "uid column without values");
    if (spec->bag_indices && n_entities > 0 && (!row_len || (!gcol && !val))) return fail(GDMIX_ERR_INVALID, "feature bag without arrays");
    gdmix_host::SeqexColumns c;
    c.entity = spec->entity; c.uid = spec->uid; c.label = spec->label; c.offset = spec->offset; c.weight = spec->weight;
    c.bag_indices = spec->bag_indices; c.bag_values = spec->bag_values;
    c.n_entities = n_entities; c.ent_rows = ent_rows; c.entity_int = entity_int; c.id_chars = id_chars; c.id_ptr = id_ptr;
    c.row_len = row_len; c.gcol = gcol; c.val = val; c.uid_v = uid; c.label_v = label; c.label_as_int = label_as_int;
    c.offset_v = offset; c.weight_v = weight;
    try {
        if (gdmix_host::seqex_encode(c, out, capacity, written) != 0)
            return fail(GDMIX_ERR_WORKSPACE, "gdmix_seqex_encode: capacity %lld < %lld bytes", (long long)capacity, (long long)*written);
    } catch (...) {
        return fail(GDMIX_ERR_INVALID, "out of host memory in gdmix_seqex_encode");
    }
    return GDMIX_OK;
}

int gdmix_partition_ids_i64(const int64_t *ids, int64_t n, int32_t num_partitions, int32_t *partition_out, void *stream)
{
    if (n < 0 || num_partitions <= 0 || (n > 0 && (!ids || !partition_out))) return fail(GDMIX_ERR_INVALID, "bad argument");
    if (n == 0) return GDMIX_OK;
    gdmix::partition_i64_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(
        ids, n, num_partitions, partition_out);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}

int gdmix_local_index_mark(const int64_t *ent_rowptr, int64_t n_entities, const int64_t *rowptr, const int32_t *gcol,
                           int64_t n_rows, int32_t num_features, uint32_t *bitmap, uint32_t *word_prefix, int64_t *d_e,
                           void *stream)
{
    if (!ent_rowptr || !rowptr || !bitmap || !word_prefix || !d_e || n_entities < 0 || n_rows < 0 || num_features <= 0)
        return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_local_index_mark");
    if (n_entities == 0) return GDMIX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int32_t W32 = (num_features + 31) / 32;
    // bitmap has one word more than the entities need: the out-of-range flag
    unsigned *bad = bitmap + (size_t)n_entities * W32;
    CUDA_TRY(cudaMemsetAsync(bitmap, 0, ((size_t)n_entities * W32 + 1) * 4, st));
    if (n_rows > 0) {
        const int g = (int)std::min<int64_t>((n_rows + 255) / 256, 148 * 16);
        gdmix::bitmap_mark_kernel<<<g, 256, 0, st>>>(ent_rowptr, n_entities, rowptr, gcol, n_rows, num_features, W32, bitmap, bad);
        g_launches++;
    }
    unsigned bad_h = 0;
    CUDA_TRY(cudaMemcpyAsync(&bad_h, bad, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (bad_h) return fail(GDMIX_ERR_INVALID, "a feature id is outside [0, num_features)");
    const int g2 = (int)std::min<int64_t>((n_entities + 255) / 256, 148 * 16);
    gdmix::bitmap_count_kernel<<<g2, 256, 0, st>>>(bitmap, n_entities, W32, word_prefix, d_e);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}

int gdmix_local_index_apply(const int64_t *ent_rowptr, int64_t n_entities, const int64_t *rowptr, const int32_t *gcol,
                            int64_t n_rows, int32_t num_features, const uint32_t *bitmap, const uint32_t *word_prefix,
                            const int64_t *uniq_ptr, int32_t *local_col, int64_t *uniq_global, void *stream)
{
    if (!ent_rowptr || !rowptr || !bitmap || !word_prefix || !uniq_ptr || !uniq_global || num_features <= 0)
        return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_local_index_apply");
    if (n_entities <= 0) return GDMIX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int32_t W32 = (num_features + 31) / 32;
    if (n_rows > 0) {
        const int g = (int)std::min<int64_t>((n_rows + 255) / 256, 148 * 16);
        gdmix::bitmap_index_kernel<<<g, 256, 0, st>>>(ent_rowptr, n_entities, rowptr, gcol, n_rows, W32, bitmap, word_prefix, local_col);
    }
    const int g2 = (int)std::min<int64_t>((n_entities + 255) / 256, 148 * 16);
    gdmix::bitmap_list_kernel<<<g2, 256, 0, st>>>(bitmap, n_entities, W32, uniq_ptr, uniq_global);
    g_launches += 2;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}

int gdmix_auc(const float *score, const float *label, int64_t n, double *out3, void *ws, size_t ws_bytes, void *stream)
{
    if (n <= 0 || !score || !label || !out3 || !ws) return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_auc");
    if (n >= (1ll << 32)) return fail(GDMIX_ERR_TOO_LARGE, "at most 2^32 - 1 scores per call");
    size_t need = 0;
    gdmix_partition_workspace_size(n, &need);
    if (ws_bytes < need) return fail(GDMIX_ERR_WORKSPACE, "workspace %zu B < required %zu B", ws_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    const SortScratch c = carve_sort(ws, n);
    const int64_t nt = sort_tiles(n);
    char *b = (char *)ws + c.bytes;
    size_t o = 0;
    uint64_t *keys = (uint64_t *)(b + o); o += up256z(8 * (size_t)n);
    uint32_t *lab = (uint32_t *)(b + o); o += up256z(4 * (size_t)n);
    uint64_t *keys_s = (uint64_t *)(b + o); o += up256z(8 * (size_t)n);
    uint32_t *lab_s = (uint32_t *)(b + o); o += up256z(4 * (size_t)n);
    int64_t *seg_ptr = (int64_t *)(b + o); o += up256z(8 * (size_t)(n + 1));
    unsigned long long *gpos = (unsigned long long *)(b + o); o += up256z(8 * (size_t)n);
    unsigned long long *gneg = (unsigned long long *)(b + o); o += up256z(8 * (size_t)n);
    uint32_t *tile_count = (uint32_t *)(b + o); o += up256z(4 * (size_t)nt);
    int64_t *tile_off = (int64_t *)(b + o); o += up256z(8 * (size_t)nt);
    int64_t *ngroups = (int64_t *)(b + o); o += 256;
    const int g1 = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    gdmix::auc_keys_kernel<<<g1, 256, 0, st>>>(score, label, n, keys, lab);
    int rc = sort_pairs(keys, lab, n, 32, keys_s, lab_s, ws, st);
    if (rc) return rc;
    gdmix::heads_count_kernel<<<(int)nt, gdmix::kSortThreads, 0, st>>>(keys_s, n, tile_count);
    gdmix::tiles_exclusive_scan_kernel<<<1, 256, 0, st>>>(tile_count, nt, tile_off, ngroups);
    gdmix::heads_write_kernel<<<(int)nt, gdmix::kSortThreads, 0, st>>>(keys_s, n, tile_off, seg_ptr, nullptr);
    // the number of groups is only known on the device: size the group kernels for n, they stop at seg_ptr's end
    int64_t ng_host = 0;
    CUDA_TRY(cudaMemcpyAsync(&ng_host, ngroups, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    // negx = exclusive scan of the sorted negative flags (n + 1 entries; spans the two per-group arrays of the
    // workspace layout, which are contiguous), with the tile arrays reused now that the heads are written
    int64_t *negx = (int64_t *)gpos;
    (void)gneg;
    gdmix::tile_sum_kernel<<<(int)nt, gdmix::kSortThreads, 0, st>>>(lab_s, n, tile_count);
    gdmix::tiles_exclusive_scan_kernel<<<1, 256, 0, st>>>(tile_count, nt, tile_off, ngroups + 1);
    gdmix::rowptr_from_len_kernel<<<(int)nt, gdmix::kSortThreads, 0, st>>>(lab_s, n, tile_off, negx);
    const int n_cta = (int)std::max<int64_t>(1, std::min<int64_t>((ng_host + 255) / 256, 148 * 8));
    double *cta_u2 = (double *)keys;   // the unsorted keys are dead after the sort
    gdmix::auc_groups_kernel<<<n_cta, 256, 0, st>>>(seg_ptr, ng_host, negx, cta_u2);
    gdmix::auc_finish_kernel<<<1, 256, 0, st>>>(cta_u2, n_cta, negx, n, out3);
    g_launches += 3;
    g_launches += 6;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}

int gdmix_seqex_count(const uint8_t *file_image, int64_t len, const gdmix_seqex_spec *spec, gdmix_seqex_sizes *sizes)
{
    if (!spec || !sizes || (len > 0 && !file_image) || len < 0 || !spec->entity || !spec->uid || !spec->bag_indices ||
        !spec->bag_values)
        return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_seqex_count");
    std::string err;
    gdmix_host::SeqexReader r(*spec, err);
    if (!r.run(file_image, len, *sizes, gdmix_host::SeqexOut())) return fail(GDMIX_ERR_INVALID, "%s", err.c_str());
    return GDMIX_OK;
}

int gdmix_seqex_fill(const uint8_t *file_image, int64_t len, const gdmix_seqex_spec *spec, int64_t *ent_rows,
                     int64_t *row_len, int64_t *gcol, float *val, int64_t *uid, float *label, float *offset,
                     float *weight, char *id_chars, int64_t *id_ptr, int64_t *index_range)
{
    if (!spec || (len > 0 && !file_image) || len < 0 || !ent_rows || !row_len || !gcol || !val || !uid || !label ||
        !offset || !weight || !id_chars || !id_ptr)
        return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_seqex_fill");
    std::string err;
    gdmix_host::SeqexReader r(*spec, err);
    gdmix_host::SeqexOut o;
    o.ent_rows = ent_rows; o.row_len = row_len; o.gcol = gcol; o.val = val; o.uid = uid; o.label = label;
    o.offset = offset; o.weight = weight; o.id_chars = id_chars; o.id_ptr = id_ptr;
    gdmix_seqex_sizes sz;
    if (!r.run(file_image, len, sz, o)) return fail(GDMIX_ERR_INVALID, "%s", err.c_str());
    if (index_range) { index_range[0] = sz.min_index; index_range[1] = sz.max_index; }
    return GDMIX_OK;
}

int gdmix_seqex_fill_local(const uint8_t *file_image, int64_t len, const gdmix_seqex_spec *spec, int64_t *ent_rows,
                           int64_t *row_len, uint16_t *local16, int64_t *d_e, int64_t *uniq_scratch, float *val,
                           int64_t *uid, float *label, float *offset, float *weight, char *id_chars, int64_t *id_ptr,
                           int64_t *index_range)
{
    if (!spec || (len > 0 && !file_image) || len < 0 || !ent_rows || !row_len || !local16 || !d_e || !uniq_scratch ||
        !val || !uid || !label || !offset || !weight || !id_chars || !id_ptr)
        return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_seqex_fill_local");
    std::string err;
    gdmix_host::SeqexReader r(*spec, err);
    gdmix_host::SeqexOut o;
    o.ent_rows = ent_rows; o.row_len = row_len; o.val = val; o.uid = uid; o.label = label;
    o.offset = offset; o.weight = weight; o.id_chars = id_chars; o.id_ptr = id_ptr;
    o.local16 = local16; o.d_e = d_e; o.uniq_scratch = uniq_scratch;
    gdmix_seqex_sizes sz;
    if (!r.run(file_image, len, sz, o))
        return fail(err.find("65535 distinct") != std::string::npos ? GDMIX_ERR_TOO_LARGE : GDMIX_ERR_INVALID, "%s", err.c_str());
    if (index_range) { index_range[0] = sz.min_index; index_range[1] = sz.max_index; }
    return GDMIX_OK;
}

int gdmix_example_count(const uint8_t *file_image, int64_t len, const gdmix_seqex_spec *spec, gdmix_seqex_sizes *sizes)
{
    if (!spec || !sizes || (len > 0 && !file_image) || len < 0 || (!spec->bag_indices != !spec->bag_values))
        return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_example_count");
    std::string err;
    gdmix_host::ExampleReader r(*spec, err);
    if (!r.run(file_image, len, *sizes, gdmix_host::ExampleOut())) return fail(GDMIX_ERR_INVALID, "%s", err.c_str());
    return GDMIX_OK;
}

int gdmix_example_fill(const uint8_t *file_image, int64_t len, const gdmix_seqex_spec *spec, int64_t *row_len,
                       int32_t *col, float *val, int64_t *uid, float *label, float *offset, float *weight)
{
    if (!spec || (len > 0 && !file_image) || len < 0 || !row_len || !col || !val || !uid || !label || !offset || !weight)
        return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_example_fill");
    std::string err;
    gdmix_host::ExampleReader r(*spec, err);
    gdmix_host::ExampleOut o;
    o.row_len = row_len; o.col = col; o.val = val; o.uid = uid; o.label = label; o.offset = offset; o.weight = weight;
    gdmix_seqex_sizes sz;
    if (!r.run(file_image, len, sz, o)) return fail(GDMIX_ERR_INVALID, "%s", err.c_str());
    return GDMIX_OK;
}

int gdmix_avro_score_blocks(const int64_t *uid, const float *score, const float *label, const float *weight,
                            const float *per_coordinate, int64_t n, int32_t records_per_block, const uint8_t *sync16,
                            uint8_t *out, int64_t capacity, int64_t *written)
{
    if (n < 0 || records_per_block <= 0 || !sync16 || !written || (n > 0 && (!uid || !score || !out)))
        return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_avro_score_blocks");
    if (capacity < gdmix_host::avro_score_blocks_bound(n, records_per_block))
        return fail(GDMIX_ERR_WORKSPACE, "output buffer %lld B < bound %lld B", (long long)capacity,
                    (long long)gdmix_host::avro_score_blocks_bound(n, records_per_block));
    *written = gdmix_host::avro_score_blocks(uid, score, label, weight, per_coordinate, n, records_per_block, sync16, out);
    return GDMIX_OK;
}

static int model_blocks_impl(const gdmix_model_table *t, int32_t records_per_block, const uint8_t *sync16, uint8_t *out,
                             int64_t capacity, int64_t *written, uint8_t **alloc_out)
{
    if (!t || !written || records_per_block <= 0 || !sync16 || t->n_models < 0 ||
        (t->n_models > 0 && (!t->id_chars || !t->id_ptr || !t->model_class || !t->coef || !t->coef_ptr || !t->intercept_name)))
        return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_avro_model_blocks");
    gdmix_host::ModelTable m;
    m.n_models = t->n_models; m.id_chars = t->id_chars; m.id_ptr = t->id_ptr; m.model_class = t->model_class;
    m.coef = t->coef; m.var = t->var; m.coef_ptr = t->coef_ptr; m.feat_idx = t->feat_idx;
    m.has_intercept = t->has_intercept; m.threshold = t->threshold; m.intercept_name = t->intercept_name;
    m.name_chars = t->name_chars; m.name_ptr = t->name_ptr; m.term_chars = t->term_chars; m.term_ptr = t->term_ptr;
    m.n_features = t->n_features;
    std::vector<int64_t> start, body;
    const int64_t need = gdmix_host::avro_model_sizes(m, records_per_block, start, body);
    if (need < 0) return fail(GDMIX_ERR_INVALID, "model table is inconsistent (coefficient slices / feature indices)");
    if (alloc_out) {
        // one sizing pass, one writing pass, into a buffer of the library's (gdmix_buffer_free)
        uint8_t *buf = (uint8_t *)malloc((size_t)std::max<int64_t>(need, 1));
        if (!buf) return fail(GDMIX_ERR_INVALID, "out of host memory for %lld bytes of model records", (long long)need);
        gdmix_host::avro_model_write(m, records_per_block, sync16, start, body, buf);
        *alloc_out = buf;
        *written = need;
        return GDMIX_OK;
    }
    if (!out) { *written = need; return GDMIX_OK; }
    if (capacity < need) return fail(GDMIX_ERR_WORKSPACE, "output buffer %lld B < required %lld B", (long long)capacity, (long long)need);
    gdmix_host::avro_model_write(m, records_per_block, sync16, start, body, out);
    *written = need;
    return GDMIX_OK;
}

int gdmix_avro_model_blocks(const gdmix_model_table *t, int32_t records_per_block, const uint8_t *sync16, uint8_t *out,
                            int64_t capacity, int64_t *written)
{
    return model_blocks_impl(t, records_per_block, sync16, out, capacity, written, nullptr);
}

int gdmix_avro_model_blocks_alloc(const gdmix_model_table *t, int32_t records_per_block, const uint8_t *sync16,
                                  uint8_t **out, int64_t *written)
{
    if (!out) return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_avro_model_blocks_alloc");
    *out = nullptr;
    return model_blocks_impl(t, records_per_block, sync16, nullptr, 0, written, out);
}

void gdmix_buffer_free(void *ptr) { free(ptr); }

struct gdmix_feature_map { gdmix_host::FeatureMap fm; };

gdmix_feature_map *gdmix_feature_map_create(const char *name_chars, const int64_t *name_ptr, const char *term_chars,
                                            const int64_t *term_ptr, int64_t n_features, const char *intercept_name)
{
    if (n_features < 0 || !intercept_name || (n_features > 0 && (!name_chars || !name_ptr || !term_chars || !term_ptr))) {
        fail(GDMIX_ERR_INVALID, "bad argument to gdmix_feature_map_create");
        return nullptr;
    }
    gdmix_feature_map *h = new gdmix_feature_map;
    h->fm.intercept = intercept_name;
    h->fm.index.reserve((size_t)n_features * 2);
    std::string key;
    for (int64_t g = 0; g < n_features; g++) {
        key.assign(name_chars + name_ptr[g], (size_t)(name_ptr[g + 1] - name_ptr[g]));
        key.push_back('\x01');
        key.append(term_chars + term_ptr[g], (size_t)(term_ptr[g + 1] - term_ptr[g]));
        h->fm.index[key] = g;   // a repeated (name, term) keeps its LAST row, as the reference's dict does
    }
    return h;
}

void gdmix_feature_map_destroy(gdmix_feature_map *h) { delete h; }

int gdmix_avro_model_decode(const gdmix_feature_map *h, const uint8_t *block, int64_t len, int64_t n_records,
                            int64_t *n_means, int64_t *id_bytes, char *id_chars, int64_t *id_ptr, int64_t *mean_ptr,
                            int64_t *mean_feat, double *mean_val, double *var_val, uint8_t *has_var)
{
    if (!h || len < 0 || n_records < 0 || (len > 0 && !block) || !n_means || !id_bytes)
        return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_avro_model_decode");
    std::string err;
    gdmix_host::ModelDecoder d(h->fm, err);
    gdmix_host::ModelDecodeSizes sz;
    gdmix_host::ModelDecodeOut o;
    if (id_chars) {
        if (!id_ptr || !mean_ptr || !mean_feat || !mean_val || !var_val || !has_var)
            return fail(GDMIX_ERR_INVALID, "gdmix_avro_model_decode: all output arrays or none");
        o.id_chars = id_chars; o.id_ptr = id_ptr; o.mean_ptr = mean_ptr; o.mean_feat = mean_feat; o.mean_val = mean_val;
        o.var_val = var_val; o.has_var = has_var;
    }
    if (!d.run(block, len, n_records, sz, o)) return fail(GDMIX_ERR_INVALID, "%s", err.c_str());
    *n_means = sz.n_means; *id_bytes = sz.id_bytes;
    return GDMIX_OK;
}

void gdmix_host_release(void)
{
    std::lock_guard<std::mutex> lk(g_host.mu);
    for (Slot &s : g_host.slot) {
        if (s.dev) cudaFree(s.dev);
        if (s.pin) cudaFreeHost(s.pin);
        if (s.ws) cudaFree(s.ws);
        if (s.st) cudaStreamDestroy(s.st);
        s = Slot();
    }
}

struct gdmix_lbfgs {
    gdmix_host::Lbfgs impl;
    gdmix_lbfgs(int64_t n, const gdmix_lr_opts *o)
        : impl(n, o->m, o->max_iter, o->max_ls, o->max_fun, o->factr, o->pgtol) {}
};

gdmix_lbfgs *gdmix_lbfgs_create(int64_t n, const gdmix_lr_opts *opts)
{
    if (n <= 0 || !opts || opts->m < 0) { fail(GDMIX_ERR_INVALID, "bad argument to gdmix_lbfgs_create"); return nullptr; }
    try {
        return new gdmix_lbfgs(n, opts);
    } catch (...) {
        fail(GDMIX_ERR_INVALID, "out of host memory in gdmix_lbfgs_create");
        return nullptr;
    }
}

int gdmix_lbfgs_iterate(gdmix_lbfgs *h, double *x, double f, const double *g)
{
    if (!h || !x || !g) return fail(GDMIX_ERR_INVALID, "null argument to gdmix_lbfgs_iterate");
    return h->impl.iterate(x, f, g);
}

int gdmix_lbfgs_info(const gdmix_lbfgs *h, int32_t *nit, int32_t *nfev, int32_t *status, double *f)
{
    if (!h) return fail(GDMIX_ERR_INVALID, "null handle");
    if (nit) *nit = h->impl.nit();
    if (nfev) *nfev = h->impl.nfev();
    if (status) *status = h->impl.status();
    if (f) *f = h->impl.f();
    return GDMIX_OK;
}

void gdmix_lbfgs_destroy(gdmix_lbfgs *h) { delete h; }

// ---- device-resident replicated L-BFGS of the fixed-effect solve (fe_lbfgs.cuh) ----------------------------------
namespace {
// Status records the host polls: one page of pinned memory for the life of the process (cudaFreeHost of a small
// pinned block costs tens to hundreds of milliseconds next to large pinned staging buffers; a solve must not pay it).
constexpr int kStatusSlots = 128;
std::mutex g_status_mu;
gdmix::FeLbStatus *g_status_page = nullptr;
bool g_status_used[kStatusSlots] = {false};
gdmix::FeLbStatus *status_slot_acquire()
{
    std::lock_guard<std::mutex> lk(g_status_mu);
    if (!g_status_page && cudaMallocHost((void **)&g_status_page, sizeof(gdmix::FeLbStatus) * kStatusSlots) != cudaSuccess) {
        cudaGetLastError();
        g_status_page = nullptr;
        return nullptr;
    }
    for (int i = 0; i < kStatusSlots; i++)
        if (!g_status_used[i]) { g_status_used[i] = true; return g_status_page + i; }
    return nullptr;
}
void status_slot_release(gdmix::FeLbStatus *p)
{
    std::lock_guard<std::mutex> lk(g_status_mu);
    if (p && g_status_page && p >= g_status_page && p < g_status_page + kStatusSlots) g_status_used[p - g_status_page] = false;
}
}  // namespace

struct gdmix_fe_lbfgs {
    gdmix::FeLbBuffers B{};
    gdmix::FeLbState init{};
    gdmix::FeLbStatus *status_dev = nullptr, *status_host = nullptr;
    void *arena = nullptr;
    int64_t n = 0;
    int m = 0;
};

gdmix_fe_lbfgs *gdmix_fe_lbfgs_create(int64_t n, const gdmix_lr_opts *o, double *x_dev, double *fg_dev)
{
    if (n <= 0 || !o || o->m < 0 || o->m > gdmix::kLbMaxM || !x_dev || !fg_dev) {
        fail(GDMIX_ERR_INVALID, "bad argument to gdmix_fe_lbfgs_create (n > 0, 0 <= m <= %d, device x and fg)", gdmix::kLbMaxM);
        return nullptr;
    }
    gdmix_fe_lbfgs *h = new (std::nothrow) gdmix_fe_lbfgs();
    if (!h) { fail(GDMIX_ERR_INVALID, "out of host memory"); return nullptr; }
    h->n = n; h->m = o->m;
    const int64_t nb = (n + gdmix::kLbBlock - 1) / gdmix::kLbBlock;
    const int64_t mm = std::max(o->m, 1);
    // one arena: state | status | d t r q | S Y | partials
    const size_t off_state = 0, off_status = 1024, off_vec = 2048;
    const size_t doubles = (size_t)(4 * n + 2 * mm * n + 4 * nb);
    const size_t bytes = off_vec + 8 * doubles;
    h->status_host = status_slot_acquire();
    if (!h->status_host || cudaMalloc(&h->arena, bytes) != cudaSuccess) {
        cudaGetLastError();
        fail(GDMIX_ERR_CUDA, "gdmix_fe_lbfgs_create: cannot allocate %zu bytes of device memory (or no status slot)", bytes);
        status_slot_release(h->status_host);
        delete h;
        return nullptr;
    }
    char *a = (char *)h->arena;
    h->B.st = (gdmix::FeLbState *)(a + off_state);
    h->status_dev = (gdmix::FeLbStatus *)(a + off_status);
    double *v = (double *)(a + off_vec);
    h->B.x = x_dev; h->B.fg = fg_dev;
    h->B.d = v; h->B.t = v + n; h->B.r = v + 2 * n; h->B.q = v + 3 * n;
    h->B.S = v + 4 * n; h->B.Y = h->B.S + mm * n;
    h->B.part = h->B.Y + mm * n;
    h->B.nb = (int32_t)nb;
    memset(&h->init, 0, sizeof(h->init));
    h->init.n = n; h->init.m = o->m; h->init.max_iter = o->max_iter; h->init.max_ls = o->max_ls;
    h->init.max_fun = o->max_fun; h->init.factr = o->factr; h->init.pgtol = o->pgtol; h->init.theta = 1.0;
    return h;
}

int gdmix_fe_lbfgs_reset(gdmix_fe_lbfgs *h, void *stream)
{
    if (!h) return fail(GDMIX_ERR_INVALID, "null handle");
    CUDA_TRY(cudaMemcpyAsync(h->B.st, &h->init, sizeof(h->init), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return GDMIX_OK;
}

int gdmix_fe_lbfgs_step(gdmix_fe_lbfgs *h, void *stream)
{
    if (!h) return fail(GDMIX_ERR_INVALID, "null handle");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = h->B.nb, T = gdmix::kLbThreads;
    gdmix::lb_dots_kernel<<<nb, T, 0, st>>>(h->B);
    gdmix::lb_decide1_kernel<<<1, 32, 0, st>>>(h->B);
    gdmix::lb_pair_kernel<<<nb, T, 0, st>>>(h->B);
    gdmix::lb_decide2_kernel<<<1, 32, 0, st>>>(h->B);
    gdmix::lb_store_kernel<<<nb, T, 0, st>>>(h->B);
    for (int j = 0; j < h->m; j++) gdmix::lb_loop1_kernel<<<nb, T, 0, st>>>(h->B, j);
    gdmix::lb_scale_kernel<<<nb, T, 0, st>>>(h->B);
    for (int k = 0; k < h->m; k++) gdmix::lb_loop2_kernel<<<nb, T, 0, st>>>(h->B, k);
    gdmix::lb_direction_kernel<<<nb, T, 0, st>>>(h->B);
    gdmix::lb_decide3_kernel<<<1, 32, 0, st>>>(h->B, h->status_dev);
    gdmix::lb_newx_kernel<<<nb, T, 0, st>>>(h->B);
    g_launches += 9 + 2 * h->m;
    CUDA_TRY(cudaGetLastError());
    return GDMIX_OK;
}

int gdmix_fe_lbfgs_poll(gdmix_fe_lbfgs *h, void *stream, int32_t *task, int32_t *nit, int32_t *nfev, int32_t *status,
                        double *f)
{
    if (!h) return fail(GDMIX_ERR_INVALID, "null handle");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(h->status_host, h->status_dev, sizeof(gdmix::FeLbStatus), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (task) *task = h->status_host->task;
    if (nit) *nit = h->status_host->nit;
    if (nfev) *nfev = h->status_host->nfev;
    if (status) *status = h->status_host->status;
    if (f) *f = h->status_host->f;
    return GDMIX_OK;
}

void gdmix_fe_lbfgs_destroy(gdmix_fe_lbfgs *h)
{
    if (!h) return;
    if (h->arena) cudaFree(h->arena);
    status_slot_release(h->status_host);
    delete h;
}

int gdmix_partition_ids(const uint16_t *units, const int64_t *id_ptr, int64_t n_ids, int32_t num_partitions,
                        int32_t *hash_out, int32_t *partition_out)
{
    if (!id_ptr || n_ids < 0 || num_partitions <= 0 || (!hash_out && !partition_out))
        return fail(GDMIX_ERR_INVALID, "bad argument to gdmix_partition_ids");
    for (int64_t e = 0; e < n_ids; e++) {
        uint32_t h = 0;
        for (int64_t k = id_ptr[e]; k < id_ptr[e + 1]; k++) h = 31u * h + (uint32_t)units[k];
        const int32_t hs = (int32_t)h;
        if (hash_out) hash_out[e] = hs;
        if (partition_out) {
            const int32_t a = (hs == INT32_MIN) ? hs : (hs < 0 ? -hs : hs);  // Math.abs(Int.MinValue) < 0
            partition_out[e] = a % num_partitions;                          // sign of the dividend, like the JVM
        }
    }
    return GDMIX_OK;
}

}  // extern "C"
