// Definitions of the per-curve host entry points (kernel launches).  Included by inst_*.cu only.
#pragma once
#include "engine.cuh"
#include "coop.cuh"

namespace b200 {

template <class C> struct CurveTraits;
// ACC_SM: default accumulate kernel (0 registers only, 1 = X, Y of the accumulator and the cp.async-staged point in shared-memory
// slots, 2 = the whole accumulator in slots) with ACC_SM_BLOCKS1 / ACC_SM_BLOCKS2 blocks of 128 threads per SM.  Measured
// (profiles/r2_experiments.md): BLS12-377 G1 n = 2^20 8.27 -> 7.95 ms (1: 128 registers, 4 blocks), 7.94 (2: 96 registers, 5
// blocks), 8.08 (2 with 6 blocks: spills); BW6-761 n = 2^22 150.0 -> 143.6 ms (1: 252 registers, no spills, 2 blocks), 147.0 (2);
// BLS12-377 G2 n = 2^20 29.2 -> 26.6 ms (1), 27.4 (2).
// ACC_SM_PIPE: the same choice inside the batch pipeline (3 = mode 1 with one block fewer per SM: with four blocks the sort and
// tail kernels of the neighbouring MSMs find no room and a pipelined 2^20 MSM takes 8.32 ms instead of 7.19; BW6-761: 42.6 ms with
// mode 1 against 44.4 with the register kernel).
// REDUCE_THREAD_MIN: from this many bucket-reduce segments on, one thread per segment (k_bucket_reduce_thread) instead of one quad:
// BW6-761 n = 2^22 143.7 -> 137.7 ms, 2^20 45.6 -> 43.9; BLS12-377 G1 n = 2^24 102.2 -> 100.5, 2^22 25.9 -> 25.6 (quads stay for the
// 8704 segments of n = 2^20, a latency problem); BLS12-377 G2 n = 2^22 88.4 -> 85.0 ms, but 2^20 26.6 -> 28.0; BW6-761 2^18 14.2 -> 16.0.
// REDUCE_THREAD_BLOCKS: resident blocks per SM of that kernel's register-capped build, used when it saves a round of blocks (<= 1 %:
// profiles/r2_experiments.md); 1: no capped build.
// AFFINE: the curve also has the experimental batched-affine accumulate kernels (compiled only with B200_WITH_CROSSCHECKS);
// SHARED_MUL: the accumulate kernel multiplies through one out-of-line product body; COOP_COMBINE: the Horner combine runs
// on four warps with one limb per lane (coop.cuh) -- 2.5x faster for the 24-limb field and for Fq2, no faster for the
// 12-limb field (0.9 against 0.85 ms) where it only costs the neighbouring MSMs of a pipelined batch more issue slots
#ifdef B200_WITH_CROSSCHECKS
constexpr bool B200_AFFINE_BUILD = true;
#else
constexpr bool B200_AFFINE_BUILD = false;
#endif
template <> struct CurveTraits<G1_377> { static constexpr int ACC_THREADS = 128, ACC_MIN_BLOCKS = 3, RED_THREADS = 128, AFT_THREADS = 128, AFT_MIN_BLOCKS = 2, ACC_SM = 1, ACC_SM_PIPE = 3, ACC_SM_BLOCKS1 = 4, ACC_SM_BLOCKS2 = 5; static constexpr bool ACC_SM_BUILD = true; static constexpr long REDUCE_THREAD_MIN = 24576; static constexpr int REDUCE_THREAD_BLOCKS = 8; static constexpr bool AFFINE = B200_AFFINE_BUILD, SHARED_MUL = true, COOP_COMBINE = false, AFFTREE = B200_AFFINE_BUILD, AFT_PREFETCH = true; };
template <> struct CurveTraits<G2_377> { static constexpr int ACC_THREADS = 128, ACC_MIN_BLOCKS = 1, RED_THREADS = 64, AFT_THREADS = 128, AFT_MIN_BLOCKS = 1, ACC_SM = 1, ACC_SM_PIPE = 1, ACC_SM_BLOCKS1 = 2, ACC_SM_BLOCKS2 = 3; static constexpr long REDUCE_THREAD_MIN = 24576; static constexpr int REDUCE_THREAD_BLOCKS = 1; static constexpr bool ACC_SM_BUILD = true; static constexpr bool AFFINE = false, SHARED_MUL = false, COOP_COMBINE = true, AFFTREE = false, AFT_PREFETCH = false; };
template <> struct CurveTraits<G_761> { static constexpr int ACC_THREADS = 128, ACC_MIN_BLOCKS = 1, RED_THREADS = 64, AFT_THREADS = 128, AFT_MIN_BLOCKS = 1, ACC_SM = 1, ACC_SM_PIPE = 1, ACC_SM_BLOCKS1 = 2, ACC_SM_BLOCKS2 = 3; static constexpr long REDUCE_THREAD_MIN = 16384; static constexpr int REDUCE_THREAD_BLOCKS = 5; static constexpr bool ACC_SM_BUILD = true; static constexpr bool AFFINE = false, SHARED_MUL = false, COOP_COMBINE = true, AFFTREE = B200_AFFINE_BUILD, AFT_PREFETCH = false; };

// Window plan: minimise (madds + bucket-reduce work) in field-multiplication units while
// keeping enough buckets in flight to fill 148 SMs.
template <class C>
static MsmPlan make_plan(size_t n);

// window plan for n pairs of which only n_eff scalars reach the buckets (the others are 0 or 1): the width follows n_eff
template <class C>
static MsmPlan make_plan_eff(size_t n, size_t n_eff) {
    MsmPlan p = make_plan<C>(std::max<size_t>(std::min(n, n_eff), 256));
    p.n = (uint32_t)n;
    return p;
}

template <class C>
static MsmPlan make_plan(size_t n) {
    MsmPlan best{};
    double best_cost = 1e300;
    for (int c = 4; c <= 20; c++) {
        int windows = (C::SCALAR_BITS + 1 + c - 1) / c;
        double nb = (double)(1u << (c - 1));
        // 10.5 ~ mixed add, 2*14 running-sum adds + ~20 for the segment scalar-mul / tree share
        double cost = windows * ((double)n * 10.5 + nb * 48.0);
        // a top window holding only a few scalar bits piles all n points into a handful of buckets: those
        // go through the block-per-bucket paths at a fraction of the machine (measured: BW6-761, n = 2^20,
        // c = 15 -> 3 top bits: 57 ms against 45 / 47 ms for c = 14 / 16)
        int top_bits = C::SCALAR_BITS + 1 - (windows - 1) * c;
        if (top_bits <= 6 && top_bits < c - 4 && n >= 4096) cost += (double)n * 10.5 * 3.0;
        // starve penalty: fewer bucket-threads than the machine holds means idle SMs
        double threads = windows * nb;
        if (threads < 148.0 * 384.0) cost *= 1.0 + 0.35 * (148.0 * 384.0 / threads - 1.0);
        if (cost < best_cost) {
            best_cost = cost;
            best.c = c;
            best.windows = windows;
            best.nb = 1u << (c - 1);
        }
    }
    if (const char *env = getenv("B200_MSM_C")) {                  // experiments: force the window width
        int c = atoi(env);
        if (c >= 2 && c <= 22) {
            best.c = c;
            best.windows = (C::SCALAR_BITS + 1 + c - 1) / c;
            best.nb = 1u << (c - 1);
        }
    }
    best.n = (uint32_t)n;
    best.seg_len = best.nb >= 8192 ? 32 : best.nb >= 4096 ? 16 : (best.nb >= 64 ? 8 : (int)std::min<uint32_t>(best.nb, 4u));
    if (const char *env = getenv("B200_MSM_SEG")) {                // experiments: force the reduce segment length
        int v = atoi(env);
        if (v >= 1 && (uint32_t)v <= best.nb && (best.nb % v) == 0) best.seg_len = v;
    }
    best.segs = best.nb / best.seg_len;
    // block-summed buckets: a few times the mean population of a bucket (uniform digits)
    // (8x: the top window of a 253-bit scalar at c = 15 holds 13 bits, its buckets are 4x the mean by design)
    best.big = (uint32_t)std::min<size_t>(SIZE_BINS - 1, std::max<size_t>(64, 8 * (n / best.nb) + 64));
    return best;
}

// arkworks records (device memory) -> native packed images
template <class C>
int pack_bases(const void *src_dev, size_t stride, size_t n, void *dst, cudaStream_t st) {
    using F = typename C::F;
    if (n == 0) return B200_OK;
    constexpr size_t XY = sizeof(AffineMem<F>);
    if (stride % 4 || stride < XY) return fail(B200_ERR_ARG, "bad base stride %zu", stride);
    constexpr int TH = 128;
    k_pack_bases<F, TH><<<ceil_div(n, TH), TH, 0, st>>>(reinterpret_cast<const uint32_t *>(src_dev), (uint32_t)n,
                                                        (uint32_t)(stride / 4), stride > XY ? 1 : 0,
                                                        reinterpret_cast<AffineMem<F> *>(dst));
    LAUNCH_CHECK();
    return B200_OK;
}

// ---- one MSM = three stages over a workspace set ----------------------------------------------------
// sort (digit histogram, scan, scatter, population order) -> accumulate (buckets) -> tail (bucket
// reduce, window sums, Horner).  msm_native runs them back to back on the caller's stream;
// msm_batch software-pipelines consecutive MSMs over three internal streams and two workspace sets.
// batched-affine accumulation (k_bucket_accumulate_affine): EXPERIMENTAL, off unless B200_MSM_AFFINE=1 --
// parity-green but only level with the XYZZ kernel (6.50 against 6.52 ms at n = 2^20); per-curve switch,
// scratch within a byte budget
constexpr size_t AFFINE_SCRATCH_BUDGET = (size_t)12 << 30;
template <class C>
static bool msm_use_affine(const MsmPlan &p, size_t n, size_t *rec_a, size_t *rec_b) {
    using F = typename C::F;
    static const int mode = getenv("B200_MSM_AFFINE") ? atoi(getenv("B200_MSM_AFFINE")) : 0;
    size_t total = (size_t)p.windows * p.nb, entries = n * (size_t)p.windows;
    *rec_a = (entries + total) / 2 + 2;
    *rec_b = (entries + 3 * total) / 4 + 2;
    if (!CurveTraits<C>::AFFINE || !mode) return false;
    return (*rec_a + *rec_b) * sizeof(AffineMem<F>) <= AFFINE_SCRATCH_BUDGET && n / p.nb >= 8;
}

// ---- batched-affine tree (msm_afftree.cuh): how many levels, how many windows per pass ------------------------------
// EXPERIMENTAL, compiled only with B200_WITH_CROSSCHECKS and off unless B200_MSM_AFFTREE = k (k levels): parity-green
// (the 70 MSM parity tests pass with the tree on, both fields), MEASURED SLOWER on B200 (profiles/r2_experiments.md):
// BLS12-377 G1 n = 2^20 8.28 ms without, 8.76 / 9.40 / 9.88 / 10.3 ms with 2 / 3 / 4 / 5 levels; BW6-761 47.2 -> 47.4 / 48.0.
// levels: the tree halves a bucket `levels` times, the XYZZ finish adds what is left (mean population / 2^levels points).
// Windows are taken in groups so that the two level buffers stay within AFT_SCRATCH_BUDGET.
constexpr size_t AFT_SCRATCH_BUDGET = (size_t)6 << 30;
struct AftPlan {
    int levels = 0, group = 0;                      // windows per pass
    size_t slots_a = 0, slots_b = 0;                // records of the odd / even level buffers
};
template <class C>
static AftPlan aft_plan(const MsmPlan &p, size_t n) {
    using F = typename C::F;
    AftPlan a;
    if (!CurveTraits<C>::AFFTREE) return a;
    static const int forced = getenv("B200_MSM_AFFTREE") ? atoi(getenv("B200_MSM_AFFTREE")) : 0;
    if (forced <= 0 || n / p.nb < 8) return a;
    const int levels = std::min(forced, 10);
    int group = p.windows;
    for (;;) {
        const size_t entries = n * (size_t)group, nbk = (size_t)group * p.nb;
        a.slots_a = aft_slots(entries, nbk, 1);
        a.slots_b = levels > 1 ? aft_slots(entries, nbk, 2) : 0;
        if ((a.slots_a + a.slots_b) * sizeof(AffineMem<F>) <= AFT_SCRATCH_BUDGET || group == 1) break;
        group = (group + 1) / 2;
    }
    if (n * (size_t)group >= ((size_t)1 << 31)) return AftPlan{};   // slot indices are 31 bits
    a.levels = levels;
    a.group = group;
    return a;
}

template <class C>
static int msm_reserve(MsmWs &W, const MsmPlan &p, size_t n) {
    using F = typename C::F;
    {
        const AftPlan a = aft_plan<C>(p, n);
        int rc0;
        if (a.levels && ((rc0 = W.aff_a.reserve(a.slots_a * sizeof(AffineMem<F>))) || (rc0 = W.aff_b.reserve(a.slots_b * sizeof(AffineMem<F>)))))
            return rc0;
    }
    {
        size_t ra, rb;
        int rc0;
        if (msm_use_affine<C>(p, n, &ra, &rb) &&
            ((rc0 = W.aff_a.reserve(ra * sizeof(AffineMem<F>))) || (rc0 = W.aff_b.reserve(rb * sizeof(AffineMem<F>)))))
            return rc0;
    }
    size_t total = (size_t)p.windows * p.nb;
    uint32_t tiles = (uint32_t)ceil_div(total, SCAN_TILE);
    size_t max_huge = std::min<size_t>(total, n * (size_t)p.windows / HUGE_BUCKET + 1);
    int rc;
    if ((rc = W.counts.reserve(total * 4)) || (rc = W.offsets.reserve((total + 1) * 4)) ||
        (rc = W.cursor.reserve(total * 4)) || (rc = W.tile_sums.reserve((size_t)tiles * 4)) ||
        (rc = W.bins.reserve(4 * SIZE_BINS * 4)) || (rc = W.order.reserve(total * 4)) ||
        (rc = W.sorted.reserve(n * (size_t)p.windows * 4)) || (rc = W.buckets.reserve(total * sizeof(XYZZMem<F>))) ||
        (rc = W.partials.reserve(((size_t)p.windows * p.segs + ONES_PARTS + ONES_GROUPS) * sizeof(XYZZMem<F>))) ||
        (rc = W.window_sums.reserve((size_t)(p.windows + 2) * sizeof(XYZZMem<F>))) ||
        (rc = W.ones.reserve((n + 1) * 4)) || (rc = W.huge_slices.reserve(max_huge * HUGE_SLICES * sizeof(XYZZMem<F>))))
        return rc;
    return B200_OK;
}

template <class C>
static int msm_stage_sort(Engine &E, MsmWs &W, const MsmPlan &p, const void *d_scalars, size_t n, cudaStream_t st, int split = 0) {
    NvtxRange range("msm.sort");
    size_t total = (size_t)p.windows * p.nb;
    uint32_t tiles = (uint32_t)ceil_div(total, SCAN_TILE);
    const uint32_t *sc = reinterpret_cast<const uint32_t *>(d_scalars);
    uint32_t *counts = W.counts.as<uint32_t>(), *offsets = W.offsets.as<uint32_t>(), *cursor = W.cursor.as<uint32_t>();
    uint32_t *bins = W.bins.as<uint32_t>();
    CUDA_TRY(cudaMemsetAsync(counts, 0, total * 4, st));
    CUDA_TRY(cudaMemsetAsync(bins, 0, 4 * SIZE_BINS * 4, st));
    CUDA_TRY(cudaMemsetAsync(W.ones.p, 0, 4, st));
    int nblk = ceil_div(n, 256);
    k_digit_hist<C::SCALAR_WORDS><<<nblk, 256, 0, st>>>(sc, p, counts);
    LAUNCH_CHECK();
    k_scan_tile_sums<<<tiles, SCAN_THREADS, 0, st>>>(counts, (uint32_t)total, W.tile_sums.as<uint32_t>());
    LAUNCH_CHECK();
    k_scan_tiles<<<1, SCAN_THREADS, 0, st>>>(W.tile_sums.as<uint32_t>(), tiles);
    LAUNCH_CHECK();
    k_scan_apply<<<tiles, SCAN_THREADS, 0, st>>>(counts, (uint32_t)total, W.tile_sums.as<uint32_t>(), offsets, cursor);
    LAUNCH_CHECK();
    k_digit_scatter<C::SCALAR_WORDS><<<nblk, 256, 0, st>>>(sc, p, cursor, W.sorted.as<uint32_t>(), W.ones.as<uint32_t>());
    LAUNCH_CHECK();
    // population order, separately for the buckets of windows [split, windows) -- group A, first in `order` -- and of
    // windows [0, split) -- group B (split = 0: one group)
    const size_t total_b = (size_t)split * p.nb, total_a = total - total_b;
    for (int g = 0; g < 2; g++) {
        const size_t cnt = g ? total_b : total_a, first = g ? 0 : total_b;     // bucket range of the group
        if (cnt == 0) continue;
        uint32_t *gb = bins + (size_t)g * 2 * SIZE_BINS;
        k_size_hist<<<std::min(ceil_div(cnt, 256), E.sm_count * 4), 256, 0, st>>>(counts + first, (uint32_t)cnt, p.big, gb);
        LAUNCH_CHECK();
        k_size_scan<<<1, SIZE_BINS, 0, st>>>(gb, gb + SIZE_BINS);
        LAUNCH_CHECK();
        k_size_scatter<<<ceil_div(cnt, 256), 256, 0, st>>>(counts + first, (uint32_t)cnt, p.big, gb + SIZE_BINS,
                                                           W.order.as<uint32_t>() + (g ? total_a : 0), (uint32_t)first);
        LAUNCH_CHECK();
    }
    return B200_OK;
}

// W: the sort buffers of this input (chunk); B: the workspace holding the buckets and the unit-scalar partials (B is W
// for a whole-input MSM; the host-pointer MSM sorts chunk by chunk and resumes ONE bucket set: resume = not the first chunk)
// split / group: the sort stage ordered the buckets of windows [split, windows) (group 0) and [0, split) (group 1)
// separately; group = -1 takes everything in one launch (split must be 0), 0 / 1 one group (unit scalars ride with group 0)
template <class C>
static int msm_stage_accumulate(Engine &E, MsmWs &W, const MsmPlan &p, const void *d_bases, size_t n, cudaStream_t st,
                                MsmWs *Bp = nullptr, int resume = 0, int split = 0, int group = -1, bool pipelined = false) {
    using F = typename C::F;
    using T = CurveTraits<C>;
    NvtxRange range("msm.accumulate");
    MsmWs &B = Bp ? *Bp : W;
    const size_t all = (size_t)p.windows * p.nb, total_b = (size_t)split * p.nb, total_a = all - total_b;
    const size_t total = group < 0 ? all : (group ? total_b : total_a);
    const size_t pts = group < 0 ? n * (size_t)p.windows : n * (size_t)(group ? split : p.windows - split);   // sorted entries of the group
    const uint32_t *order = W.order.as<uint32_t>() + (group == 1 ? total_a : 0);
    const AffineMem<F> *bases = reinterpret_cast<const AffineMem<F> *>(d_bases);
    uint32_t *offsets = W.offsets.as<uint32_t>(), *bins = W.bins.as<uint32_t>() + (group == 1 ? 2 * SIZE_BINS : 0);
    if (total == 0) return B200_OK;
    bool prof = group < 0 && E.profile && E.prof_used < Engine::PROF_SLOTS;    // window groups: timed by msm_split_run
    if (prof) CUDA_TRY(cudaEventRecord(E.prof_ev[2 * E.prof_used], st));
    size_t ra, rb;
    bool affine = false;
    if constexpr (T::AFFINE) affine = !resume && msm_use_affine<C>(p, n, &ra, &rb);
    if constexpr (T::AFFINE) {
        if (affine) {
            static const int variant = getenv("B200_AFFINE_VARIANT") ? atoi(getenv("B200_AFFINE_VARIANT")) : 0;
            AffineMem<F> *sa = W.aff_a.as<AffineMem<F>>(), *sb = W.aff_b.as<AffineMem<F>>();
            const uint32_t *so = W.sorted.as<uint32_t>(), *ord = order;
            XYZZMem<F> *bk = B.buckets.as<XYZZMem<F>>();
#define B200_AFF_LAUNCH(TH, MB, BB, CC, BI)                                                                          \
    k_bucket_accumulate_affine<F, TH, MB, BB, CC, BI><<<ceil_div(total, TH), TH, 0, st>>>(bases, so, offsets, ord,      \
                                                                                          (uint32_t)total, p.big, sa, sb, bk)
            switch (variant) {                                     // profiles/r1_experiments.md has the measurements
            case 1: B200_AFF_LAUNCH(128, 3, 32, 4, true); break;   // one inversion per block and round (scans + barriers)
            case 2: B200_AFF_LAUNCH(128, 3, 32, 4, false); break;  // per-thread inversion, tree down to 4 points
            default: B200_AFF_LAUNCH(128, 3, 32, 8, false); break; // per-thread inversion, tree down to 8 points
            }
#undef B200_AFF_LAUNCH
        }
    }
    // 12-limb curves: the accumulate kernel with ONE out-of-line product body (12 KB of code instead of 117 KB)
    // is 1.8 % faster under the pipelined batch (7.48 against 7.61 ms); B200_MSM_SHAREDMUL=0 selects the inlined one
    static const bool shared_mul = !(getenv("B200_MSM_SHAREDMUL") && !atoi(getenv("B200_MSM_SHAREDMUL")));
    AftPlan aft{};
    if constexpr (T::AFFTREE) {
        if (group < 0 && !affine) aft = aft_plan<C>(p, n);
    }
    if constexpr (T::AFFTREE) if (aft.levels) {
        // batched-affine tree, window group by window group (the level buffers are reused), then the XYZZ finish
        const uint32_t *so = W.sorted.as<uint32_t>();
        AffineMem<F> *buf[2] = {W.aff_a.as<AffineMem<F>>(), W.aff_b.as<AffineMem<F>>()};
        for (int w0 = 0; w0 < p.windows; w0 += aft.group) {
            const int gw = std::min(aft.group, p.windows - w0);
            const uint32_t b0 = (uint32_t)w0 * p.nb, nbk = (uint32_t)gw * p.nb;
            const size_t entries = n * (size_t)gw;
            for (int L = 0; L < aft.levels; L++) {
                const size_t slots = aft_slots(entries, nbk, L + 1);
                const unsigned blocks = (unsigned)ceil_div(ceil_div(slots, (size_t)AFT_B), (size_t)T::AFT_THREADS);
                k_affine_level<F, T::AFT_THREADS, T::AFT_MIN_BLOCKS, T::AFT_PREFETCH><<<blocks, T::AFT_THREADS, 0, st>>>(
                    bases, so, offsets, b0, nbk, p.big, L, L ? buf[(L - 1) & 1] : nullptr, buf[L & 1], (uint32_t)slots);
                LAUNCH_CHECK();
            }
            k_bucket_finish<F, T::ACC_THREADS, T::ACC_MIN_BLOCKS><<<ceil_div(total, T::ACC_THREADS), T::ACC_THREADS, 0, st>>>(
                buf[(aft.levels - 1) & 1], offsets, order, (uint32_t)total, b0, nbk, p.big, aft.levels, resume, B.buckets.as<XYZZMem<F>>());
            LAUNCH_CHECK();
        }
    }
    // Accumulator (partly) in shared memory, more blocks per SM (msm.cuh k_bucket_accumulate_sm).  B200_MSM_ACC_SM overrides
    // the per-curve default: 0 register kernel; 1 X, Y and the staged point in slots; 2 the whole accumulator in slots
    static const int acc_sm_env = getenv("B200_MSM_ACC_SM") ? atoi(getenv("B200_MSM_ACC_SM")) : -1;
    // the batch pipeline (msm_batch) keeps sort / tail kernels of the neighbouring MSMs on the SMs beside this one
    static const int acc_sm_pipe_env = getenv("B200_MSM_ACC_SM_PIPE") ? atoi(getenv("B200_MSM_ACC_SM_PIPE")) : -1;
    const int acc_sm = pipelined ? (acc_sm_pipe_env >= 0 ? acc_sm_pipe_env : T::ACC_SM_PIPE) : (acc_sm_env >= 0 ? acc_sm_env : T::ACC_SM);
    bool acc_sm_done = false;
    if constexpr (T::ACC_SM_BUILD) {
        if (!aft.levels && !affine && acc_sm) {
            using SL = AccSlots<F, 128>;
            using PP = F;
            const unsigned grid = (unsigned)ceil_div(total, 128);
            auto go = [&](auto kern, int slots) -> int {
                CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SL::bytes(slots)));
                kern<<<grid, 128, SL::bytes(slots), st>>>(bases, W.sorted.as<uint32_t>(), offsets, order, (uint32_t)total, p.big, resume,
                                                          B.buckets.as<XYZZMem<F>>());
                return B200_OK;
            };
            const int rc2 = acc_sm == 1   ? go(k_bucket_accumulate_sm<PP, 128, T::ACC_SM_BLOCKS1, false>, 4)
                            : acc_sm == 3 ? go(k_bucket_accumulate_sm<PP, 128, T::ACC_SM_BLOCKS1 - 1, false>, 4)
                                          : go(k_bucket_accumulate_sm<PP, 128, T::ACC_SM_BLOCKS2, true>, 6);
            if (rc2) return rc2;
            acc_sm_done = true;
        }
    }
    if (aft.levels || acc_sm_done) {
    } else if (!affine && shared_mul && T::SHARED_MUL)
        k_bucket_accumulate_shared<F, T::ACC_THREADS, T::ACC_MIN_BLOCKS>
            <<<ceil_div(total, T::ACC_THREADS), T::ACC_THREADS, 0, st>>>(
                bases, W.sorted.as<uint32_t>(), offsets, order, (uint32_t)total, p.big, resume,
                B.buckets.as<XYZZMem<F>>());
    else if (!affine)
        k_bucket_accumulate<F, T::ACC_THREADS, T::ACC_MIN_BLOCKS>
            <<<ceil_div(total, T::ACC_THREADS), T::ACC_THREADS, 0, st>>>(
                bases, W.sorted.as<uint32_t>(), offsets, order, (uint32_t)total, p.big, resume,
                B.buckets.as<XYZZMem<F>>());
    LAUNCH_CHECK();
    if (prof) {
        CUDA_TRY(cudaEventRecord(E.prof_ev[2 * E.prof_used + 1], st));
        E.prof_used++;
        E.prof_units += n;
    }
    // over-populated buckets (skewed scalars) and unit scalars: bounded extra launches
    constexpr int BT = T::RED_THREADS;                             // smem: BT XYZZ images (<= 24.5 KB)
    uint32_t max_big = (uint32_t)std::min<size_t>(total, pts / p.big + 1);
    uint32_t max_huge = (uint32_t)std::min<size_t>(total, pts / HUGE_BUCKET + 1);
    k_big_buckets<F, BT><<<std::min<uint32_t>(max_big, (uint32_t)E.sm_count * 16), BT, BT * sizeof(XYZZMem<F>), st>>>(
        bases, W.sorted.as<uint32_t>(), offsets, order, bins, p.big, resume, B.buckets.as<XYZZMem<F>>());
    LAUNCH_CHECK();
    k_huge_buckets<F, BT><<<dim3(max_huge, HUGE_SLICES), BT, BT * sizeof(XYZZMem<F>), st>>>(
        bases, W.sorted.as<uint32_t>(), offsets, order, bins, p.big, W.huge_slices.as<XYZZMem<F>>());
    LAUNCH_CHECK();
    k_huge_finish<F, BT><<<ceil_div((size_t)max_huge * 4, BT), BT, 0, st>>>(
        W.huge_slices.as<XYZZMem<F>>(), offsets, order, bins, p.big, max_huge, resume, B.buckets.as<XYZZMem<F>>());
    LAUNCH_CHECK();
    if (group <= 0) {
        k_ones_accumulate<F, BT><<<ONES_PARTS / BT, BT, 0, st>>>(bases, W.ones.as<uint32_t>(), resume,
                                                                 B.partials.as<XYZZMem<F>>() + (size_t)p.windows * p.segs);
        LAUNCH_CHECK();
    }
    return B200_OK;
}

// bucket reduce + window sums of windows [w_lo, w_hi) (with_ones: also the unit-scalar partials -> window_sums[windows])
template <class C>
static int msm_stage_reduce(MsmWs &W, const MsmPlan &p, int w_lo, int w_hi, bool with_ones, cudaStream_t st) {
    using F = typename C::F;
    using T = CurveTraits<C>;
    uint32_t red_threads = (uint32_t)(w_hi - w_lo) * p.segs * 4;   // one quad per segment
    if (red_threads) {
        // quads (latency) below REDUCE_THREAD_MIN segments, one thread per segment (throughput) from there on;
        // B200_REDUCE_THREAD_MIN overrides the per-curve limit (0: always threads)
        static const long env_min = getenv("B200_REDUCE_THREAD_MIN") ? atol(getenv("B200_REDUCE_THREAD_MIN")) : -1;
        const long thread_min = env_min >= 0 ? env_min : T::REDUCE_THREAD_MIN;
        const uint32_t segments = red_threads / 4;
        if ((long)segments >= thread_min) {
            constexpr int TT = 64;
            // register cap (REDUCE_THREAD_BLOCKS resident blocks per SM) when it saves a round of blocks; B200_REDUCE_THREAD_CAP = 0 / 1
            // forces it off / on (read at every call: tools compare both forms inside one process)
            const char *cap_env = getenv("B200_REDUCE_THREAD_CAP");
            bool cap = false;
            if constexpr (T::REDUCE_THREAD_BLOCKS > 1) {
                static const int uncapped_blocks = [] {
                    int b = 1;
                    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_bucket_reduce_thread<F, TT>, TT, 0) != cudaSuccess) b = 1;
                    return std::max(1, b);
                }();
                static const int sm_count = [] {
                    int d = 0, n = 148;
                    if (cudaGetDevice(&d) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess) n = 148;
                    return std::max(1, n);
                }();
                const size_t blocks = ceil_div(segments, TT), sms = (size_t)sm_count;
                const size_t rounds_now = ceil_div(blocks, sms * (size_t)uncapped_blocks);
                const size_t rounds_cap = ceil_div(blocks, sms * (size_t)T::REDUCE_THREAD_BLOCKS);
                cap = cap_env ? atoi(cap_env) != 0 : rounds_cap < rounds_now;
                if (cap)
                    k_bucket_reduce_thread<F, TT, T::REDUCE_THREAD_BLOCKS><<<ceil_div(segments, TT), TT, 0, st>>>(
                        W.buckets.as<XYZZMem<F>>(), p, w_lo, w_hi, W.partials.as<XYZZMem<F>>());
            }
            if (!cap)
                k_bucket_reduce_thread<F, TT><<<ceil_div(segments, TT), TT, 0, st>>>(W.buckets.as<XYZZMem<F>>(), p, w_lo, w_hi,
                                                                                 W.partials.as<XYZZMem<F>>());
        } else {
            k_bucket_reduce<F, T::RED_THREADS><<<ceil_div(red_threads, T::RED_THREADS), T::RED_THREADS, 0, st>>>(
                W.buckets.as<XYZZMem<F>>(), p, w_lo, w_hi, W.partials.as<XYZZMem<F>>());
        }
        LAUNCH_CHECK();
    }
    constexpr int WS_THREADS = 256;                                // 64 quads per window
    size_t ws_smem = (WS_THREADS / 4) * sizeof(XYZZMem<F>);
    if (with_ones) {
        XYZZMem<F> *parts = W.partials.as<XYZZMem<F>>() + (size_t)p.windows * p.segs;
        k_ones_fold<F, 128><<<ONES_GROUPS, 128, (128 / 4) * sizeof(XYZZMem<F>), st>>>(parts, parts + ONES_PARTS);
        LAUNCH_CHECK();
    }
    k_window_sum<F, WS_THREADS><<<(w_hi - w_lo) + (with_ones ? 1 : 0), WS_THREADS, ws_smem, st>>>(
        W.partials.as<XYZZMem<F>>(), p, w_lo, w_hi, W.window_sums.as<XYZZMem<F>>());
    LAUNCH_CHECK();
    return B200_OK;
}

template <class C>
static int msm_stage_tail(MsmWs &W, const MsmPlan &p, void *d_out, cudaStream_t st) {
    using F = typename C::F;
    NvtxRange range("msm.tail");
    int rc = msm_stage_reduce<C>(W, p, 0, p.windows, true, st);
    if (rc) return rc;
    // Horner over the windows: one quad (12-limb field) or four warps sharing every point operation (coop.cuh);
    // B200_COMBINE = quad | coop overrides the per-curve choice (experiments)
    static const char *force = getenv("B200_COMBINE");
    const bool coop = force ? force[0] == 'c' : CurveTraits<C>::COOP_COMBINE;
    if (!coop)
        k_window_combine<F><<<1, 32, 0, st>>>(W.window_sums.as<XYZZMem<F>>(), p, reinterpret_cast<JacobianMem<F> *>(d_out));
    else
        k_window_combine_coop<F><<<1, COOP_THREADS, 0, st>>>(W.window_sums.as<XYZZMem<F>>(), 0, p.windows, p.c, 0,
                                                             W.window_sums.as<XYZZMem<F>>() + p.windows, nullptr,
                                                             reinterpret_cast<JacobianMem<F> *>(d_out), nullptr);
    LAUNCH_CHECK();
    return B200_OK;
}

// ---- window groups: the tail of the high windows hides under the accumulation of the low ones ---------------
// The tail (bucket reduce, window sums, Horner) is latency-bound and leaves the multiplier pipe idle; a single MSM
// used to pay it in full after the accumulation (1.6 of 8.3 ms at n = 2^20, 14 of 47 ms for BW6-761).  The windows
// are therefore cut in two groups: group A = windows [split, W) is bucketed first, its tail -- including the
// split * c doublings that carry its Horner value down to weight 2^0 -- runs on the high-priority tail stream while
// group B = windows [0, split) is still being bucketed; only B's short tail (split windows, (split - 1) * c doublings)
// and one addition remain exposed.  Both groups' accumulate kernels are in flight together (A launched first), so B's
// blocks fill the SMs as A's drain.
template <class C>
static int msm_split_of(const MsmPlan &p, size_t n) {
    // MEASURED NEGATIVE on B200 (profiles/r2_experiments.md): BLS12-377 G1 n = 2^20 8.44 ms unsplit, 8.53 / 8.58 / 8.76 / 9.10 /
    // 9.41 ms at split = 2 / 3 / 4 / 6 / 8; BW6-761 47.4 -> 47.5 .. 48.9 ms.  The exposed remainder is not the Horner chain any
    // more (coop.cuh) but the bucket reduce, whose duration does not shrink with the number of windows (it is paced by the
    // two warps per scheduler it puts on the machine), while the second accumulate launch and the displaced blocks cost more
    // than the hidden part saves.  Off by default; B200_MSM_SPLIT = k forces windows [0, k) into the second group.
    static const int forced = getenv("B200_MSM_SPLIT") ? atoi(getenv("B200_MSM_SPLIT")) : 0;
    (void)n;
    return forced > 0 ? std::min(forced, p.windows - 1) : 0;
}

// sort buffers in W (already sorted with the same `split`), buckets in B; everything ordered after `st`'s prior work and
// joined back into it.  resume: the buckets hold the sums of earlier chunks.
template <class C>
static int msm_split_run(Engine &E, MsmWs &W, MsmWs &B, const MsmPlan &p, int split, const void *d_bases, size_t n, int resume,
                         void *d_out, cudaStream_t st) {
    using F = typename C::F;
    int rc;
    if (split == 0) {
        if ((rc = msm_stage_accumulate<C>(E, W, p, d_bases, n, st, &B, resume))) return rc;
        return msm_stage_tail<C>(B, p, d_out, st);
    }
    cudaStream_t s_a = E.split_stream, s_t = E.pipe_stream[2];
    const bool prof = E.profile && E.prof_used < Engine::PROF_SLOTS;
    if (prof) CUDA_TRY(cudaEventRecord(E.prof_ev[2 * E.prof_used], st));
    CUDA_TRY(cudaEventRecord(E.ev_split[0], st));
    CUDA_TRY(cudaStreamWaitEvent(s_a, E.ev_split[0], 0));
    if ((rc = msm_stage_accumulate<C>(E, W, p, d_bases, n, s_a, &B, resume, split, 0))) return rc;     // group A first
    CUDA_TRY(cudaEventRecord(E.ev_split[1], s_a));
    if ((rc = msm_stage_accumulate<C>(E, W, p, d_bases, n, st, &B, resume, split, 1))) return rc;      // group B beside it
    if (prof) {
        CUDA_TRY(cudaStreamWaitEvent(st, E.ev_split[1], 0));       // the timed span covers both groups' kernels
        CUDA_TRY(cudaEventRecord(E.prof_ev[2 * E.prof_used + 1], st));
        E.prof_used++;
        E.prof_units += n;
    }
    // tail of group A on the tail stream: value carried down by split * c doublings, unit scalars added -> XYZZ in window_sums[windows + 1]
    XYZZMem<F> *ws = B.window_sums.as<XYZZMem<F>>();
    CUDA_TRY(cudaStreamWaitEvent(s_t, E.ev_split[1], 0));
    if ((rc = msm_stage_reduce<C>(B, p, split, p.windows, true, s_t))) return rc;
    k_window_combine_coop<F><<<1, COOP_THREADS, 0, s_t>>>(ws, split, p.windows, p.c, split * p.c, ws + p.windows, nullptr, nullptr,
                                                          ws + p.windows + 1);
    LAUNCH_CHECK();
    CUDA_TRY(cudaEventRecord(E.ev_split[2], s_t));
    // tail of group B, then the sum of the two
    if ((rc = msm_stage_reduce<C>(B, p, 0, split, false, st))) return rc;
    CUDA_TRY(cudaStreamWaitEvent(st, E.ev_split[2], 0));
    k_window_combine_coop<F><<<1, COOP_THREADS, 0, st>>>(ws, 0, split, p.c, 0, ws + p.windows + 1, nullptr,
                                                         reinterpret_cast<JacobianMem<F> *>(d_out), nullptr);
    LAUNCH_CHECK();
    return B200_OK;
}

// d_bases: native packed images (pack_bases output); d_out: arkworks GroupProjective image
template <class C>
int msm_native(Engine &E, const void *d_bases, const void *d_scalars, size_t n, void *d_out, cudaStream_t st, size_t n_eff) {
    using F = typename C::F;
    if (n == 0) {
        k_sum_jacobian<F><<<1, SUM_THREADS, 0, st>>>(nullptr, 0, reinterpret_cast<JacobianMem<F> *>(d_out));
        LAUNCH_CHECK();
        return B200_OK;
    }
    if (n > (size_t)1 << 26) return fail(B200_ERR_ARG, "n = %zu exceeds the 2^26 per-call limit", n);
    MsmPlan p = n_eff ? make_plan_eff<C>(n, n_eff) : make_plan<C>(n);
    MsmWs &W = E.ws[0];
    int rc;
    if ((rc = msm_reserve<C>(W, p, n))) return rc;
    ENGINE_ORDER(st);
    const int split = msm_split_of<C>(p, n);
    if ((rc = msm_stage_sort<C>(E, W, p, d_scalars, n, st, split))) return rc;
    if ((rc = msm_split_run<C>(E, W, W, p, split, d_bases, n, 0, d_out, st))) return rc;
    ENGINE_MARK(st);
    return B200_OK;
}

// `count` independent MSMs (native packed bases), software-pipelined: the sort of job i+1 and the tail
// of job i-1 run on their own streams beside the accumulation of job i (workspace sets alternate).
// Everything is ordered after `st`'s prior work and joined back into `st` before returning.
// ready (optional): ready[i] is recorded (on any stream) once job i's inputs are in place; the
// sort of job i waits for it.  Used by the host-pointer API to overlap H2D copies with compute.
template <class C>
int msm_batch(Engine &E, const b200_msm_job *jobs, size_t count, cudaStream_t st, const cudaEvent_t *ready, const size_t *n_eff) {
    using F = typename C::F;
    int rc;
    std::vector<MsmPlan> plans(count);
    for (size_t i = 0; i < count; i++) {
        if (jobs[i].n > (size_t)1 << 26) return fail(B200_ERR_ARG, "n = %zu exceeds the 2^26 per-call limit", jobs[i].n);
        if (!jobs[i].d_out_jacobian || (jobs[i].n && (!jobs[i].d_bases_packed || !jobs[i].d_scalars)))
            return fail(B200_ERR_ARG, "null pointer in job %zu", i);
        if (jobs[i].n == 0) continue;
        plans[i] = n_eff ? make_plan_eff<C>(jobs[i].n, n_eff[i]) : make_plan<C>(jobs[i].n);
    }
    // workspace sets rotate over the non-empty jobs.  THREE sets: the tail of job k (bucket reduce, window sums, Horner: 1.6 ms
    // alone, several times that while the next accumulation owns the SMs) may lag two accumulations behind before job k + 3
    // needs its buckets; with two sets every accumulation waited ~1 ms for the tail of the job two before it
    // (B200_PIPE_SETS=2 restores that)
    static const int sets = getenv("B200_PIPE_SETS") ? std::min(std::max(atoi(getenv("B200_PIPE_SETS")), 2), Engine::PIPE_SETS_MAX) : 3;
    size_t live = 0;
    for (size_t i = 0; i < count; i++)
        if (jobs[i].n && (rc = msm_reserve<C>(E.ws[live++ % sets], plans[i], jobs[i].n))) return rc;
    ENGINE_ORDER(st);
    cudaStream_t s_sort = E.pipe_stream[0], s_tail = E.pipe_stream[2];
    // consecutive accumulations alternate between TWO streams (they touch different workspace sets): the blocks of job k + 1
    // fill the SMs as job k's last blocks drain instead of waiting for the whole kernel to end (B200_PIPE_ACC_STREAMS=1: one stream)
    static const bool two_acc = !(getenv("B200_PIPE_ACC_STREAMS") && atoi(getenv("B200_PIPE_ACC_STREAMS")) == 1);
    cudaStream_t s_acc2[2] = {E.pipe_stream[1], two_acc ? E.split_stream : E.pipe_stream[1]};
    CUDA_TRY(cudaEventRecord(E.ev_fork, st));
    for (cudaStream_t s : {s_sort, s_acc2[0], s_acc2[1], s_tail}) CUDA_TRY(cudaStreamWaitEvent(s, E.ev_fork, 0));
    size_t k = 0;
    for (size_t i = 0; i < count; i++) {
        if (jobs[i].n == 0) {
            k_sum_jacobian<F><<<1, SUM_THREADS, 0, s_tail>>>(nullptr, 0, reinterpret_cast<JacobianMem<F> *>(jobs[i].d_out_jacobian));
            LAUNCH_CHECK();
            continue;
        }
        const int w = (int)(k % sets);
        MsmWs &W = E.ws[w];
        cudaStream_t s_acc = s_acc2[k & 1];
        if (ready) CUDA_TRY(cudaStreamWaitEvent(s_sort, ready[i], 0));
        if (k >= (size_t)sets) CUDA_TRY(cudaStreamWaitEvent(s_sort, E.ev_acc[w], 0));      // sort buffers of the set's previous job consumed
        if ((rc = msm_stage_sort<C>(E, W, plans[i], jobs[i].d_scalars, jobs[i].n, s_sort))) return rc;
        CUDA_TRY(cudaEventRecord(E.ev_sorted[w], s_sort));
        CUDA_TRY(cudaStreamWaitEvent(s_acc, E.ev_sorted[w], 0));
        if (k >= (size_t)sets) CUDA_TRY(cudaStreamWaitEvent(s_acc, E.ev_tail[w], 0));      // buckets of the set's previous job consumed
        if ((rc = msm_stage_accumulate<C>(E, W, plans[i], jobs[i].d_bases_packed, jobs[i].n, s_acc, nullptr, 0, 0, -1, true))) return rc;
        CUDA_TRY(cudaEventRecord(E.ev_acc[w], s_acc));
        CUDA_TRY(cudaStreamWaitEvent(s_tail, E.ev_acc[w], 0));
        if ((rc = msm_stage_tail<C>(W, plans[i], jobs[i].d_out_jacobian, s_tail))) return rc;
        CUDA_TRY(cudaEventRecord(E.ev_tail[w], s_tail));
        k++;
    }
    CUDA_TRY(cudaEventRecord(E.ev_join, s_tail));                  // the tail stream is last in every chain
    CUDA_TRY(cudaStreamWaitEvent(st, E.ev_join, 0));
    ENGINE_MARK(st);
    return B200_OK;
}

// ---- one MSM fed chunk by chunk (the host-pointer path) -------------------------------------------------
// The input crosses PCIe in chunks; chunk k is sorted by window digit and bucketed as soon as it has landed,
// into ONE bucket set that the following chunks resume, and the tail runs once at the end: copies hide behind
// the bucket accumulation of earlier chunks without paying one tail per chunk.  Window width follows the whole
// input.  Sort buffers alternate between the two workspace sets (the sort of chunk k + 1 runs beside the
// accumulation of chunk k); buckets and unit-scalar partials live in set 0.
//   begin(total)  ->  add(d_bases, d_scalars, cnt, ready) per chunk  ->  finish(d_out) joins everything into `st`
template <class C>
int msm_chunks_begin(Engine &E, size_t n_total, size_t chunk_max, cudaStream_t st) {
    if (n_total > (size_t)1 << 26) return fail(B200_ERR_ARG, "n = %zu exceeds the 2^26 per-call limit", n_total);
    E.chunk_plan = make_plan<C>(n_total);
    E.chunk_index = 0;
    E.chunk_done = false;
    MsmPlan p = E.chunk_plan;
    int rc;
    // bucket-side buffers sized by the plan (set 0), sort-side buffers by the largest chunk (both sets)
    if ((rc = msm_reserve<C>(E.ws[0], p, chunk_max)) || (rc = msm_reserve<C>(E.ws[1], p, chunk_max))) return rc;
    cudaStream_t s_sort = E.pipe_stream[0], s_acc = E.pipe_stream[1];
    CUDA_TRY(cudaEventRecord(E.ev_fork, st));
    for (cudaStream_t s : {s_sort, s_acc}) CUDA_TRY(cudaStreamWaitEvent(s, E.ev_fork, 0));
    return B200_OK;
}

// last: the final chunk -- its accumulation runs in two window groups and the tail of the high group hides under the
// low group's accumulation (msm_split_run); msm_chunks_finish then only joins
template <class C>
int msm_chunks_add(Engine &E, const void *d_bases_packed, const void *d_scalars, size_t cnt, cudaEvent_t ready, int last,
                   void *d_out) {
    if (cnt == 0 && !last) return B200_OK;
    const size_t k = E.chunk_index++;
    const int w = (int)(k & 1);
    MsmPlan p = E.chunk_plan;
    p.n = (uint32_t)cnt;
    p.big = (uint32_t)std::min<size_t>(SIZE_BINS - 1, std::max<size_t>(64, 8 * (cnt / p.nb) + 64));
    cudaStream_t s_sort = E.pipe_stream[0], s_acc = E.pipe_stream[1];
    int rc;
    if (ready) CUDA_TRY(cudaStreamWaitEvent(s_sort, ready, 0));
    if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(s_sort, E.ev_acc[w], 0));          // sort buffers of chunk k - 2 consumed
    const int split = last ? msm_split_of<C>(p, E.chunk_plan.n) : 0;
    if ((rc = msm_stage_sort<C>(E, E.ws[w], p, d_scalars, cnt, s_sort, split))) return rc;
    CUDA_TRY(cudaEventRecord(E.ev_sorted[w], s_sort));
    CUDA_TRY(cudaStreamWaitEvent(s_acc, E.ev_sorted[w], 0));
    if (last) {
        if ((rc = msm_split_run<C>(E, E.ws[w], E.ws[0], p, split, d_bases_packed, cnt, k > 0, d_out, s_acc))) return rc;
        E.chunk_done = true;
    } else if ((rc = msm_stage_accumulate<C>(E, E.ws[w], p, d_bases_packed, cnt, s_acc, &E.ws[0], k > 0))) {
        return rc;
    }
    CUDA_TRY(cudaEventRecord(E.ev_acc[w], s_acc));
    return B200_OK;
}

template <class C>
int msm_chunks_finish(Engine &E, void *d_out, cudaStream_t st) {
    using F = typename C::F;
    cudaStream_t s_acc = E.pipe_stream[1];
    if (E.chunk_done) {
        // the last chunk ran its own tail
    } else if (E.chunk_index == 0) {
        k_sum_jacobian<F><<<1, SUM_THREADS, 0, s_acc>>>(nullptr, 0, reinterpret_cast<JacobianMem<F> *>(d_out));
        LAUNCH_CHECK();
    } else {
        int rc = msm_stage_tail<C>(E.ws[0], E.chunk_plan, d_out, s_acc);
        if (rc) return rc;
    }
    CUDA_TRY(cudaEventRecord(E.ev_join, s_acc));
    CUDA_TRY(cudaStreamWaitEvent(st, E.ev_join, 0));
    ENGINE_MARK(st);
    return B200_OK;
}

// d_bases: arkworks-radix records in device memory at `stride` bytes
template <class C>
int msm_device(Engine &E, const void *d_bases, size_t stride, const void *d_scalars, size_t n, void *d_out,
               cudaStream_t st) {
    using F = typename C::F;
    ENGINE_ORDER(st);
    // packed records are already the engine's layout (same Montgomery radix as arkworks)
    if (stride == sizeof(AffineMem<F>)) return msm_native<C>(E, d_bases, d_scalars, n, d_out, st);
    int rc = E.native_bases.reserve(n * sizeof(AffineMem<F>));
    if (rc) return rc;
    if ((rc = pack_bases<C>(d_bases, stride, n, E.native_bases.p, st))) return rc;
    return msm_native<C>(E, E.native_bases.p, d_scalars, n, d_out, st);
}

// how many of the n scalars are neither 0 nor 1 (synchronises `st`: call it before queueing the work it plans)
template <class C>
int scalar_census(Engine &E, const void *d_scalars, size_t n, size_t *n_eff, cudaStream_t st) {
    *n_eff = n;
    if (n == 0) return B200_OK;
    int rc = E.census.reserve(16);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(E.census.p, 0, 16, st));
    k_scalar_census<C::SCALAR_WORDS><<<std::min(ceil_div(n, 256), E.sm_count * 8), 256, 0, st>>>(
        reinterpret_cast<const uint32_t *>(d_scalars), (uint32_t)n, E.census.as<uint32_t>());
    LAUNCH_CHECK();
    uint32_t h[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(h, E.census.p, 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *n_eff = h[2];
    return B200_OK;
}

template <class C>
int sum_jacobian(const void *pts, size_t count, void *out, cudaStream_t st) {
    using F = typename C::F;
    k_sum_jacobian<F><<<1, SUM_THREADS, 0, st>>>(reinterpret_cast<const JacobianMem<F> *>(pts), (uint32_t)count,
                                       reinterpret_cast<JacobianMem<F> *>(out));
    LAUNCH_CHECK();
    return B200_OK;
}

// out[b] = sum_{i < count} pts[i * batch + b] for b < batch
template <class C>
int sum_jacobian_batch(const void *pts, size_t count, size_t batch, void *out, cudaStream_t st) {
    using F = typename C::F;
    if (batch == 0) return B200_OK;
    k_sum_jacobian<F><<<(unsigned)batch, SUM_THREADS, 0, st>>>(reinterpret_cast<const JacobianMem<F> *>(pts), (uint32_t)count,
                                                     reinterpret_cast<JacobianMem<F> *>(out), (uint32_t)batch);
    LAUNCH_CHECK();
    return B200_OK;
}

template <class C>
int fixed_base_mul(Engine &E, const void *base, const void *scalars, size_t n, void *out, cudaStream_t st) {
    using F = typename C::F;
    if (n == 0) return B200_OK;
    int rc = E.ws[0].buckets.reserve(n * sizeof(XYZZMem<F>));
    if (rc) return rc;
    constexpr int TH = 64;
    k_fixed_base_mul<F, C::SCALAR_WORDS, TH><<<ceil_div(n, TH), TH, 0, st>>>(
        reinterpret_cast<const AffineMem<F> *>(base), reinterpret_cast<const uint32_t *>(scalars), (uint32_t)n,
        E.ws[0].buckets.as<XYZZMem<F>>());
    LAUNCH_CHECK();
    k_xyzz_to_affine<F, TH><<<ceil_div(n, TH), TH, 0, st>>>(E.ws[0].buckets.as<XYZZMem<F>>(), (uint32_t)n,
                                                            reinterpret_cast<AffineMem<F> *>(out));
    LAUNCH_CHECK();
    return B200_OK;
}

template <class C>
int point_runs(Engine &E, const void *base, const void *scalars, size_t runs, size_t run, void *out, cudaStream_t st) {
    using F = typename C::F;
    const size_t n = runs * run;
    if (n == 0) return B200_OK;
    int rc = E.ws[0].buckets.reserve(n * sizeof(XYZZMem<F>));
    if (rc) return rc;
    constexpr int TH = 64;
    k_point_runs<F, C::SCALAR_WORDS, TH><<<ceil_div(runs, TH), TH, 0, st>>>(
        reinterpret_cast<const AffineMem<F> *>(base), reinterpret_cast<const uint32_t *>(scalars), (uint32_t)runs, (uint32_t)run,
        E.ws[0].buckets.as<XYZZMem<F>>());
    LAUNCH_CHECK();
    k_xyzz_to_affine<F, TH><<<ceil_div(n, TH), TH, 0, st>>>(E.ws[0].buckets.as<XYZZMem<F>>(), (uint32_t)n,
                                                            reinterpret_cast<AffineMem<F> *>(out));
    LAUNCH_CHECK();
    return B200_OK;
}

template <class C>
int batch_to_affine(const void *jac, size_t n, void *out, cudaStream_t st) {
    using F = typename C::F;
    if (n == 0) return B200_OK;
    constexpr int TH = 64, BATCH = 8;
    k_jacobian_to_affine<F, TH, BATCH><<<ceil_div(ceil_div(n, BATCH), TH), TH, 0, st>>>(
        reinterpret_cast<const JacobianMem<F> *>(jac), (uint32_t)n, reinterpret_cast<AffineMem<F> *>(out));
    LAUNCH_CHECK();
    return B200_OK;
}

template <class C>
int plan_query(size_t n, int *c, int *w, uint32_t *nb) {
    MsmPlan p = make_plan<C>(n);
    if (c) *c = p.c;
    if (w) *w = p.windows;
    if (nb) *nb = p.nb;
    return B200_OK;
}


template <class C>
int field_op(int op, const void *a, const void *b, size_t n, void *out, cudaStream_t st) {
    using F = typename C::F;
    if (n == 0) return B200_OK;
    using M = typename F::Mem;
    if (op >= 8) {                                                 // the warp-cooperative routines of coop.cuh, one warp per element
        k_coop_field_op<F><<<ceil_div(n * 32, COOP_THREADS), COOP_THREADS, 0, st>>>(
            op - 8, reinterpret_cast<const M *>(a), reinterpret_cast<const M *>(b), (uint32_t)n, reinterpret_cast<M *>(out));
        LAUNCH_CHECK();
        return B200_OK;
    }
    k_field_op<F><<<ceil_div(n, 64), 64, 0, st>>>(op, reinterpret_cast<const M *>(a), reinterpret_cast<const M *>(b),
                                                  (uint32_t)n, reinterpret_cast<M *>(out));
    LAUNCH_CHECK();
    return B200_OK;
}

#define B200_INSTANTIATE(C)                                                                                       \
    template int msm_device<C>(Engine &, const void *, size_t, const void *, size_t, void *, cudaStream_t);       \
    template int pack_bases<C>(const void *, size_t, size_t, void *, cudaStream_t);                               \
    template int msm_native<C>(Engine &, const void *, const void *, size_t, void *, cudaStream_t, size_t);       \
    template int msm_batch<C>(Engine &, const b200_msm_job *, size_t, cudaStream_t, const cudaEvent_t *, const size_t *); \
    template int scalar_census<C>(Engine &, const void *, size_t, size_t *, cudaStream_t);                         \
    template int sum_jacobian<C>(const void *, size_t, void *, cudaStream_t);                                     \
    template int sum_jacobian_batch<C>(const void *, size_t, size_t, void *, cudaStream_t);                       \
    template int fixed_base_mul<C>(Engine &, const void *, const void *, size_t, void *, cudaStream_t);           \
    template int msm_chunks_begin<C>(Engine &, size_t, size_t, cudaStream_t);                                     \
    template int msm_chunks_add<C>(Engine &, const void *, const void *, size_t, cudaEvent_t, int, void *);       \
    template int msm_chunks_finish<C>(Engine &, void *, cudaStream_t);                                            \
    template int point_runs<C>(Engine &, const void *, const void *, size_t, size_t, void *, cudaStream_t);       \
    template int batch_to_affine<C>(const void *, size_t, void *, cudaStream_t);                                  \
    template int plan_query<C>(size_t, int *, int *, uint32_t *);                                                 \
    template int field_op<C>(int, const void *, const void *, size_t, void *, cudaStream_t);

}  // namespace b200
