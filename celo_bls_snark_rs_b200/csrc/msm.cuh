// Pippenger bucket MSM for sm_100a -- kernels.
//
// Device replacement for ark-ec 0.1.0 VariableBaseMSM::multi_scalar_mul as called at
//   crates/bls-crypto/src/bls/signature.rs:85   (BLS12-377 G1)
//   crates/bls-crypto/src/bls/public.rs:61      (BLS12-377 G2)
//   crates/epoch-snark/src/api/prover.rs:78,112 (Groth16 MSMs, BW6-761 / BLS12-377)
// The result is the same group element; the schedule is B200-first and deliberately
// not arkworks' (which parallelises over ~17 windows only):
//
//   1. k_digit_hist      signed-digit recoding of every scalar (c-bit windows, digits in
//                        [-2^(c-1), 2^(c-1)]) + per-(window, bucket) histogram
//   2. k_scan_*          exclusive scan of the histogram -> bucket offsets
//   3. k_digit_scatter   counting-sort scatter of (point index | sign) by bucket
//   4. k_size_*          buckets ordered by population so the lanes of a warp run equally long
//   5. k_bucket_accumulate  one thread per bucket, XYZZ accumulator in registers, affine
//                        bases gathered from the (L2-resident) packed base array with
//                        128-bit loads, next point prefetched while the current one is added
//   6. k_bucket_reduce   segment running sums (sum_b b*B_b) -> one partial per segment
//   7. k_window_sum      per-window tree sum of the partials
//   8. k_window_combine  Horner over the windows (c doublings each), Jacobian out
#pragma once
#include "ec.cuh"

namespace b200 {

struct G1_377 {
    using F = Fq377;
    static constexpr int SCALAR_WORDS = 8;          // 32-bit words per scalar in memory (BigInteger256)
    static constexpr int SCALAR_BITS = 253;
};
struct G2_377 {
    using F = Fp2<Fq377>;
    static constexpr int SCALAR_WORDS = 8;
    static constexpr int SCALAR_BITS = 253;
};
struct G_761 {                                      // BW6-761 G1 and G2 (both over Fq, a = 0)
    using F = Fq761;
    static constexpr int SCALAR_WORDS = 12;         // BigInteger384
    static constexpr int SCALAR_BITS = 377;
};

struct MsmPlan {
    uint32_t n;
    int c;                 // window bits
    int windows;           // ceil((SCALAR_BITS + 1) / c)
    uint32_t nb;           // buckets per window = 2^(c-1)
    int seg_len;           // buckets per reduce segment
    uint32_t segs;         // segments per window = nb / seg_len
    uint32_t big;          // buckets holding >= big points are summed by a whole block (<= SIZE_BINS - 1)
};

// signed window digit of the scalar at s (global memory, canonical little-endian words)
template <int SW>
B200_DEV int signed_digit(const uint32_t *__restrict__ s, int w, int c, int &carry) {
    int start = w * c, k = start >> 5, off = start & 31;
    uint64_t lo = k < SW ? __ldg(s + k) : 0u;
    uint64_t hi = k + 1 < SW ? __ldg(s + k + 1) : 0u;
    int d = (int)((uint32_t)(((hi << 32) | lo) >> off) & ((1u << c) - 1u)) + carry;
    carry = d > (1 << (c - 1));
    return carry ? d - (1 << c) : d;
}

// scalar == 1: arkworks adds such bases once, outside the buckets; here
// they go to a separate list summed by k_ones_accumulate -- Groth16 witnesses are mostly 0/1
// (crates/epoch-snark/src/api/prover.rs:78), which would otherwise pile half the input into one bucket.
template <int SW>
B200_DEV bool scalar_is_one(const uint32_t *__restrict__ s) {
    uint32_t o = __ldg(s) ^ 1u;
#pragma unroll
    for (int k = 1; k < SW; k++) o |= __ldg(s + k);
    return o == 0;
}

template <int SW>
__global__ void __launch_bounds__(256) k_digit_hist(const uint32_t *__restrict__ scalars, MsmPlan p,
                                                    uint32_t *__restrict__ counts) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const uint32_t *s = scalars + (size_t)i * SW;
    if (scalar_is_one<SW>(s)) return;
    int carry = 0;
    for (int w = 0; w < p.windows; w++) {
        int d = signed_digit<SW>(s, w, p.c, carry);
        if (d) atomicAdd(&counts[(size_t)w * p.nb + (uint32_t)(abs(d) - 1)], 1u);
    }
}

template <int SW>
__global__ void __launch_bounds__(256) k_digit_scatter(const uint32_t *__restrict__ scalars, MsmPlan p,
                                                       uint32_t *__restrict__ cursor, uint32_t *__restrict__ sorted,
                                                       uint32_t *__restrict__ ones /*[0] = count, [1..] = indices*/) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const uint32_t *s = scalars + (size_t)i * SW;
    if (scalar_is_one<SW>(s)) {
        ones[1 + atomicAdd(ones, 1u)] = i;
        return;
    }
    int carry = 0;
    for (int w = 0; w < p.windows; w++) {
        int d = signed_digit<SW>(s, w, p.c, carry);
        if (d) {
            uint32_t pos = atomicAdd(&cursor[(size_t)w * p.nb + (uint32_t)(abs(d) - 1)], 1u);
            sorted[pos] = i | (d < 0 ? 0x80000000u : 0u);
        }
    }
}

// ---- exclusive scan over `total` counters: 3 phases, SCAN_TILE elements per block --------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_PER_THREAD = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;

B200_DEV uint32_t block_exclusive_scan(uint32_t v, uint32_t *smem /*[33]*/, uint32_t &block_total) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t ws = lane < nwarps ? smem[lane] : 0u;
        uint32_t wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        smem[lane] = wi - ws;                       // exclusive warp offsets
        if (lane == 31) smem[32] = wi;
    }
    __syncthreads();
    uint32_t r = smem[warp] + incl - v;
    block_total = smem[32];
    __syncthreads();
    return r;
}

static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const uint32_t *__restrict__ counts, uint32_t total,
                                                                  uint32_t *__restrict__ tile_sums) {
    __shared__ uint32_t sm[33];
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD, s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) s += base + k < total ? counts[base + k] : 0u;
    uint32_t tot;
    block_exclusive_scan(s, sm, tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// single block: in-place exclusive scan of the tile sums (any count)
static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(uint32_t *__restrict__ tile_sums, uint32_t tiles) {
    __shared__ uint32_t sm[33];
    uint32_t running = 0;
    for (uint32_t b = 0; b < tiles; b += SCAN_THREADS) {
        uint32_t i = b + threadIdx.x;
        uint32_t v = i < tiles ? tile_sums[i] : 0u, tot;
        uint32_t e = block_exclusive_scan(v, sm, tot);
        if (i < tiles) tile_sums[i] = running + e;
        running += tot;
    }
}

static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t *__restrict__ counts, uint32_t total,
                                                              const uint32_t *__restrict__ tile_sums,
                                                              uint32_t *__restrict__ offsets /*[total + 1]*/,
                                                              uint32_t *__restrict__ cursor /*[total]*/) {
    __shared__ uint32_t sm[33];
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
    uint32_t v[SCAN_PER_THREAD], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        v[k] = base + k < total ? counts[base + k] : 0u;
        s += v[k];
    }
    uint32_t tot;
    uint32_t e = block_exclusive_scan(s, sm, tot) + tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        if (base + k < total) {
            offsets[base + k] = e;
            cursor[base + k] = e;
        }
        e += v[k];
        if (base + k + 1 == total) offsets[total] = e;
    }
}

// ---- order buckets by population (descending) so warps are uniformly loaded ---------------
constexpr int SIZE_BINS = 1024;

static __global__ void __launch_bounds__(256) k_size_hist(const uint32_t *__restrict__ counts, uint32_t total, uint32_t big,
                                                   uint32_t *__restrict__ bin_counts) {
    __shared__ uint32_t h[SIZE_BINS];
    for (int i = threadIdx.x; i < SIZE_BINS; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
        atomicAdd(&h[min(counts[i], big)], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < SIZE_BINS; i += blockDim.x)
        if (h[i]) atomicAdd(&bin_counts[i], h[i]);
}

// one block of SIZE_BINS threads: bin_cursor[b] = number of buckets in strictly larger bins
static __global__ void __launch_bounds__(SIZE_BINS) k_size_scan(const uint32_t *__restrict__ bin_counts,
                                                         uint32_t *__restrict__ bin_cursor) {
    __shared__ uint32_t sm[33];
    int rev = SIZE_BINS - 1 - threadIdx.x;          // thread 0 owns the largest bin
    uint32_t tot;
    uint32_t e = block_exclusive_scan(bin_counts[rev], sm, tot);
    bin_cursor[rev] = e;
}

static __global__ void __launch_bounds__(256) k_size_scatter(const uint32_t *__restrict__ counts, uint32_t total, uint32_t big,
                                                      uint32_t *__restrict__ bin_cursor, uint32_t *__restrict__ order,
                                                      uint32_t id_base) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    uint32_t bin = min(counts[i], big);
    // warp-aggregated cursor bump: one atomic per distinct bin per warp
    uint32_t peers = __match_any_sync(__activemask(), bin);
    int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&bin_cursor[bin], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    order[base + __popc(peers & ((1u << lane) - 1u))] = id_base + i;     // counts points at the group's first bucket
}

// Buckets holding at least MsmPlan::big points (the top population bin of k_size_*; a few times the
// mean population) are summed by a whole block instead of one thread, and those holding at least
// HUGE_BUCKET points by HUGE_SLICES blocks: skewed scalar sets (many equal small scalars in Groth16
// witnesses, the short top window of any scalar size) stay bounded.
constexpr uint32_t HUGE_BUCKET = 8192;
constexpr uint32_t HUGE_SLICES = 32;

// ---- bucket accumulation: the dominant kernel ---------------------------------------------
// One thread per bucket.  bases are native-radix packed affine images (k_pack_bases output);
// the next image is prefetched (still packed: 24 / 48 registers) while the current point is added.
template <class F, int THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
k_bucket_accumulate(const AffineMem<F> *__restrict__ bases, const uint32_t *__restrict__ sorted,
                    const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ order, uint32_t total_buckets,
                    uint32_t big, int resume, XYZZMem<F> *__restrict__ buckets) {
    uint32_t t = blockIdx.x * THREADS + threadIdx.x;
    if (t >= total_buckets) return;
    uint32_t id = order[t];
    uint32_t k = offsets[id], end = offsets[id + 1];
    if (end - k >= big) return;                      // left to k_big_buckets (one block per bucket)
    // resume: the buckets already hold the sums of earlier input chunks (host-pointer MSM, one chunk per H2D copy)
    if (resume && k == end) return;
    XYZZ<F> acc = resume ? XYZZ<F>::load(buckets[id]) : XYZZ<F>::inf();
    if (k < end) {
        // software pipeline, two deep on the indices: the gather of point k + 1 needs sorted[k + 1], which was loaded one
        // iteration earlier -- an index load and the gather that depends on it never wait for each other in one iteration
        // (ncu before: 0.62 long-scoreboard stalls per issue, the warp parked on the index before it could even start the add)
        uint32_t e = __ldg(sorted + k);
        uint32_t e_next = k + 1 < end ? __ldg(sorted + k + 1) : 0u;
        AffineMem<F> img = ldg_mem(bases + (e & 0x7fffffffu));
        for (;;) {
            ++k;
            AffineMem<F> img_next;
            const bool more = k < end;
            const uint32_t e_after = k + 1 < end ? __ldg(sorted + k + 1) : 0u;
            if (more) img_next = ldg_mem(bases + (e_next & 0x7fffffffu));
            Affine<F> pt = Affine<F>::load(img);
            if (!pt.is_inf()) acc.madd(pt.x, pt.y.cneg(e >> 31));
            if (!more) break;
            e = e_next;
            e_next = e_after;
            img = img_next;
        }
    }
    buckets[id] = acc.store();
}

// ---- bucket accumulation with batched affine additions ------------------------------------------
// Same job as k_bucket_accumulate (one thread per bucket), different arithmetic: the points of a bucket
// are summed as a pairwise tree of AFFINE additions, lambda = (y2 - y1) / (x2 - x1), whose inversions
// are shared by the whole block with Montgomery's trick -- 6 field products per addition (1 prefix
// product, 2 to peel the shared inverse, lambda, lambda^2, y3) against 10 for the XYZZ mixed addition.
// A level halves the point count; a thread handles its pairs in rounds of up to B, every round ends in
// ONE field inversion per block (thread 0, binary Euclid) behind a warp-shuffle product scan.  Level
// outputs ping-pong between two scratch arrays (per-bucket regions, sized from the bucket offsets);
// once a bucket is down to CUT points the rest is a short XYZZ chain.  Exceptional pairs are exact:
// P + P takes the tangent (denominator 2y), P + (-P) and infinite operands use denominator 1.
// The affine kernel's warps are spread over different phases (prefix products, inversion, peel-off,
// final chain), so its instruction working set is the whole kernel: with the 600-instruction Montgomery
// product inlined at ~20 sites it thrashed the instruction cache (ncu: 2.1 "no instruction" stalls per
// issue).  Everything in this kernel therefore multiplies through ONE out-of-line body.
template <class F> struct Shared {
    B200_DEV static F mul(const F &a, const F &b) { return a * b; }                 // Fp2: already out of line
    B200_DEV static F sqr(const F &a) { return a.sqr(); }
};
template <class P> struct Shared<Fp<P>> {
    B200_DEV static Fp<P> mul(const Fp<P> &a, const Fp<P> &b) { return Fp<P>::mul_outline(a, b); }
    B200_DEV static Fp<P> sqr(const Fp<P> &a) { return Fp<P>::sqr_outline(a); }      // dedicated squaring (fp.cuh)
};
// this += (px, py), XYZZ mixed addition built on the shared product (see XYZZ::madd for the formulas)
template <class F>
__device__ __noinline__ XYZZ<F> xyzz_madd_shared(XYZZ<F> a, F px, F py) {
    using S = Shared<F>;
    if (a.is_inf()) return {px, py, F::one(), F::one()};
    F p = S::mul(px, a.zz) - a.x;
    F r = S::mul(py, a.zzz) - a.y;
    if (p.is_zero()) return r.is_zero() ? XYZZ<F>::dbl_affine(px, py) : XYZZ<F>::inf();
    F pp = S::sqr(p);
    F ppp = S::mul(p, pp);
    F q = S::mul(a.x, pp);
    XYZZ<F> o;
    o.x = S::sqr(r) - ppp - q.dbl();
    o.y = S::mul(r, q - o.x) - S::mul(a.y, ppp);
    o.zz = S::mul(a.zz, pp);
    o.zzz = S::mul(a.zzz, ppp);
    return o;
}

template <class F>
B200_DEV F aff_denominator(const Affine<F> &a, const Affine<F> &b) {
    if (a.is_inf() || b.is_inf()) return F::one();
    F d = b.x - a.x;
    if (!d.is_zero()) return d;
    F s2 = a.y + b.y;                                // equal x: 2y for P + P, 0 for P + (-P)
    return s2.is_zero() ? F::one() : s2;
}
template <class F>
B200_DEV Affine<F> aff_add_with_inverse(const Affine<F> &a, const Affine<F> &b, const F &inv) {
    if (a.is_inf()) return b;
    if (b.is_inf()) return a;
    F dx = b.x - a.x, num;
    if (dx.is_zero()) {
        if ((a.y + b.y).is_zero()) return {F::zero(), F::zero()};
        F xx = Shared<F>::sqr(a.x);
        num = xx.dbl() + xx;                         // tangent: 3 x^2 / (2 y)
    } else {
        num = b.y - a.y;
    }
    F lam = Shared<F>::mul(num, inv);
    F x3 = Shared<F>::sqr(lam) - a.x - b.x;
    return {x3, Shared<F>::mul(lam, a.x - x3) - a.y};
}

// every thread passes the product `run` of its own denominators (non-zero) and gets 1 / run back:
// inclusive prefix and suffix product scans inside each warp (shuffles), the warp totals through shared
// memory, one inversion of the block total.  sm: THREADS / 32 + 1 field images.
template <class F> struct FieldInv;
template <class F, int THREADS>
B200_DEV F block_inverse(const F &run, typename F::Mem *sm) {
    constexpr int NW = THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    F incl = run, sufx = run;
#pragma unroll 1
    for (int o = 1; o < 32; o <<= 1) {
        F up = incl.shfl(0xffffffffu, lane >= o ? lane - o : lane);
        F dn = sufx.shfl(0xffffffffu, lane + o < 32 ? lane + o : lane);
        if (lane >= o) incl = Shared<F>::mul(incl, up);
        if (lane + o < 32) sufx = Shared<F>::mul(sufx, dn);
    }
    if (lane == 31) sm[warp] = incl.store();
    __syncthreads();
    F total = F::one(), others = F::one();
#pragma unroll 1
    for (int v = 0; v < NW; v++) {
        F wv = F::load(sm[v]);
        total = Shared<F>::mul(total, wv);
        if (v != warp) others = Shared<F>::mul(others, wv);
    }
    __syncthreads();
    if (threadIdx.x == 0) sm[NW] = FieldInv<F>::inv(total).store();
    __syncthreads();
    F r = Shared<F>::mul(F::load(sm[NW]), others);
    F pe = incl.shfl(0xffffffffu, lane ? lane - 1 : 0), se = sufx.shfl(0xffffffffu, lane < 31 ? lane + 1 : 31);
    if (lane) r = Shared<F>::mul(r, pe);
    if (lane < 31) r = Shared<F>::mul(r, se);
    __syncthreads();
    return r;
}

// BLOCKINV = false: every thread inverts its own running product (32 independent safegcd inversions
// cost a warp the instructions of one; they are mostly ALU-pipe work beside the other warps' multiplier
// work) -- no scans, no block barriers, warps stay independent as in k_bucket_accumulate.
template <class F, int THREADS, int MIN_BLOCKS, int B, int CUT, bool BLOCKINV>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
k_bucket_accumulate_affine(const AffineMem<F> *__restrict__ bases, const uint32_t *__restrict__ sorted,
                           const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ order, uint32_t total_buckets,
                           uint32_t big, AffineMem<F> *scratch_a, AffineMem<F> *scratch_b, XYZZMem<F> *__restrict__ buckets) {
    __shared__ typename F::Mem sm_inv[THREADS / 32 + 1];
    const uint32_t t = blockIdx.x * THREADS + threadIdx.x;
    bool mine = t < total_buckets;
    uint32_t id = mine ? order[t] : 0u, lo = 0, cnt = 0;
    if (mine) {
        lo = offsets[id];
        cnt = offsets[id + 1] - lo;
        if (cnt >= big) {                            // left to k_big_buckets / k_huge_buckets
            mine = false;
            cnt = 0;
        }
    }
    // per-bucket regions: ceil(m / 2) records in A, ceil(m / 4) in B (see msm_reserve for the totals)
    AffineMem<F> *reg_a = scratch_a + (((size_t)lo + id) >> 1), *reg_b = scratch_b + (((size_t)lo + 3 * (size_t)id) >> 2);
    int level = 0;
    auto point = [&](int lvl, uint32_t j) -> Affine<F> {
        if (lvl == 0) {
            uint32_t e = __ldg(sorted + lo + j);
            Affine<F> p = Affine<F>::load(ldg_mem(bases + (e & 0x7fffffffu)));
            p.y = p.y.cneg(e >> 31);                  // -(0) = 0: the infinity encoding (0, 0) survives
            return p;
        }
        const AffineMem<F> *src = (lvl & 1) ? reg_a : reg_b;
        return Affine<F>::load(src[j]);
    };
    typename F::Mem prefix[B];
    while (BLOCKINV ? __syncthreads_or(cnt > (uint32_t)CUT) : cnt > (uint32_t)CUT) {
        const uint32_t pairs = cnt > (uint32_t)CUT ? cnt >> 1 : 0u;
        AffineMem<F> *dst = (level & 1) ? reg_b : reg_a;           // outputs of level L are the inputs of level L + 1
        for (uint32_t j0 = 0; BLOCKINV ? __syncthreads_or(j0 < pairs) : j0 < pairs; j0 += B) {
            const int nb = j0 < pairs ? (int)min((uint32_t)B, pairs - j0) : 0;
            F run = F::one();
#pragma unroll 1
            for (int j = 0; j < nb; j++) {
                run = Shared<F>::mul(run, aff_denominator(point(level, 2 * (j0 + j)), point(level, 2 * (j0 + j) + 1)));
                prefix[j] = run.store();
            }
            F inv = BLOCKINV ? block_inverse<F, THREADS>(run, sm_inv) : FieldInv<F>::inv(run);
#pragma unroll 1
            for (int j = nb - 1; j >= 0; j--) {
                Affine<F> p1 = point(level, 2 * (j0 + j)), p2 = point(level, 2 * (j0 + j) + 1);
                F di = j ? Shared<F>::mul(inv, F::load(prefix[j - 1])) : inv;
                inv = Shared<F>::mul(inv, aff_denominator(p1, p2));
                dst[j0 + j] = aff_add_with_inverse(p1, p2, di).store();
            }
        }
        if (pairs) {
            if (cnt & 1u) dst[pairs] = point(level, cnt - 1).store();
            cnt = (cnt + 1) >> 1;
            level++;
        }
    }
    if (!mine) return;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t j = 0; j < cnt; j++) {
        Affine<F> p = point(level, j);
        if (!p.is_inf()) acc = xyzz_madd_shared(acc, p.x, p.y);
    }
    buckets[id] = acc.store();
}

// k_bucket_accumulate with every product through the shared out-of-line body: code size 117 KB -> ~12 KB
// against ~8 % call overhead; the default for the 12-limb field (B200_MSM_SHAREDMUL=0 selects the inlined kernel)
template <class F, int THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
k_bucket_accumulate_shared(const AffineMem<F> *__restrict__ bases, const uint32_t *__restrict__ sorted,
                           const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ order, uint32_t total_buckets,
                           uint32_t big, int resume, XYZZMem<F> *__restrict__ buckets) {
    uint32_t t = blockIdx.x * THREADS + threadIdx.x;
    if (t >= total_buckets) return;
    uint32_t id = order[t];
    uint32_t k = offsets[id], end = offsets[id + 1];
    if (end - k >= big) return;
    if (resume && k == end) return;
    XYZZ<F> acc = resume ? XYZZ<F>::load(buckets[id]) : XYZZ<F>::inf();
    if (k < end) {
        // software pipeline, two deep on the indices: the gather of point k + 1 needs sorted[k + 1], which was loaded one
        // iteration earlier -- an index load and the gather that depends on it never wait for each other in one iteration
        // (ncu before: 0.62 long-scoreboard stalls per issue, the warp parked on the index before it could even start the add)
        uint32_t e = __ldg(sorted + k);
        uint32_t e_next = k + 1 < end ? __ldg(sorted + k + 1) : 0u;
        AffineMem<F> img = ldg_mem(bases + (e & 0x7fffffffu));
        for (;;) {
            ++k;
            AffineMem<F> img_next;
            const bool more = k < end;
            const uint32_t e_after = k + 1 < end ? __ldg(sorted + k + 1) : 0u;
            if (more) img_next = ldg_mem(bases + (e_next & 0x7fffffffu));
            Affine<F> pt = Affine<F>::load(img);
            if (!pt.is_inf()) acc = xyzz_madd_shared(acc, pt.x, pt.y.cneg(e >> 31));
            if (!more) break;
            e = e_next;
            e_next = e_after;
            img = img_next;
        }
    }
    buckets[id] = acc.store();
}

// ---- bucket accumulation with half the accumulator in shared memory (round 2) -----------------------------------------
// k_bucket_accumulate_shared is bound by the multiplier pipe at 85 % with THREE blocks per SM (168 registers: XYZZ
// accumulator 48, current and prefetched point image 48, the temporaries of the mixed addition); the remaining stalls are
// fixed-latency waits that only more warps can fill.  Here X and Y of the accumulator and the staged point live in
// per-thread shared-memory slots (conflict-free 128-bit layout: chunk k of slot s of thread t at ((s * CH + k) * THREADS
// + t) * 16), the next point is gathered global -> shared by cp.async (LDGSTS, no register staging) as soon as the current
// one's two products have consumed it, and ZZ / ZZZ plus at most five temporaries stay in registers: <= 128 registers,
// FOUR blocks per SM.  Same additions in the same order as the register kernel (bit-identical buckets).
B200_DEV uint4 acc_lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
B200_DEV void acc_sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" : : "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <class F, int THREADS>
struct AccSlots {
    using Mem = typename F::Mem;
    static_assert(sizeof(Mem) % 16 == 0, "field images are whole 16-byte chunks");
    static constexpr int CH = (int)(sizeof(Mem) / 16);
    static constexpr uint32_t STRIDE = THREADS * 16u, SLOT = CH * STRIDE;
    enum : uint32_t { X = 0, Y = 1, PX = 2, PY = 3 };
    static constexpr uint32_t bytes(int slots) { return (uint32_t)slots * SLOT; }
    B200_DEV static F ld(uint32_t base, uint32_t slot) {
        Mem m;
        uint4 *d = reinterpret_cast<uint4 *>(&m);
#pragma unroll
        for (int k = 0; k < CH; k++) d[k] = acc_lds128(base + slot * SLOT + k * STRIDE);
        return F::load(m);
    }
    B200_DEV static void st(uint32_t base, uint32_t slot, const F &v) {
        const Mem m = v.store();
        const uint4 *d = reinterpret_cast<const uint4 *>(&m);
#pragma unroll
        for (int k = 0; k < CH; k++) acc_sts128(base + slot * SLOT + k * STRIDE, d[k]);
    }
    // one packed affine record global -> slots PX, PY, asynchronously
    B200_DEV static void fetch_point(uint32_t base, const AffineMem<F> *g) {
        const uint4 *src = reinterpret_cast<const uint4 *>(g);
#pragma unroll
        for (int k = 0; k < 2 * CH; k++)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" : : "r"(base + (PX + k / CH) * SLOT + (k % CH) * STRIDE), "l"(src + k) : "memory");
        asm volatile("cp.async.commit_group;" : : : "memory");
    }
    B200_DEV static void wait_point() { asm volatile("cp.async.wait_group 0;" : : : "memory"); }
};

// ZS: ZZ and ZZZ live in slots too (six slots per thread), leaving only the temporaries of one addition in registers.
template <class F, int THREADS, int MIN_BLOCKS, bool ZS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
k_bucket_accumulate_sm(const AffineMem<F> *__restrict__ bases, const uint32_t *__restrict__ sorted,
                       const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ order, uint32_t total_buckets,
                       uint32_t big, int resume, XYZZMem<F> *__restrict__ buckets) {
    using S = AccSlots<F, THREADS>;
    constexpr uint32_t SZZ = 4, SZZZ = 5;            // slots of ZZ / ZZZ when ZS
    extern __shared__ uint4 acc_sm[];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(acc_sm) + threadIdx.x * 16u;
    uint32_t t = blockIdx.x * THREADS + threadIdx.x;
    if (t >= total_buckets) return;
    const uint32_t id = order[t];
    uint32_t k = offsets[id];
    const uint32_t end = offsets[id + 1];
    if (end - k >= big) return;                      // left to k_big_buckets (one block per bucket)
    if (resume && k == end) return;
    F zz_r = F::zero(), zzz_r = F::zero();           // the register copies (ZS: unused)
    bool empty = true;                               // the accumulator holds nothing yet
    auto get_zz = [&]() { return ZS ? S::ld(base, SZZ) : zz_r; };
    auto get_zzz = [&]() { return ZS ? S::ld(base, SZZZ) : zzz_r; };
    auto set_z = [&](const F &a, const F &b) {
        if (ZS) {
            S::st(base, SZZ, a);
            S::st(base, SZZZ, b);
        } else {
            zz_r = a;
            zzz_r = b;
        }
    };
    if (resume) {
        const XYZZ<F> b = XYZZ<F>::load(buckets[id]);
        S::st(base, S::X, b.x);
        S::st(base, S::Y, b.y);
        set_z(b.zz, b.zzz);
        empty = b.zz.is_zero();
    }
    if (k < end) {
        uint32_t e = __ldg(sorted + k);
        uint32_t e_next = k + 1 < end ? __ldg(sorted + k + 1) : 0u;
        S::fetch_point(base, bases + (e & 0x7fffffffu));
        for (;;) {
            ++k;
            const bool more = k < end;
            const uint32_t e_after = k + 1 < end ? __ldg(sorted + k + 1) : 0u;
            S::wait_point();
            F px = S::ld(base, S::PX), py = S::ld(base, S::PY);
            const bool pt_inf = px.is_zero() && py.is_zero();
            py = py.cneg(e >> 31);
            if (pt_inf) {
                if (more) S::fetch_point(base, bases + (e_next & 0x7fffffffu));
            } else if (empty) {                       // first point of the bucket
                S::st(base, S::X, px);
                S::st(base, S::Y, py);
                set_z(F::one(), F::one());
                empty = false;
                if (more) S::fetch_point(base, bases + (e_next & 0x7fffffffu));
            } else {
                // madd-2008-s, operands named as in XYZZ::madd
                F p = F::mul_outline(px, get_zz()) - S::ld(base, S::X);
                F r = F::mul_outline(py, get_zzz()) - S::ld(base, S::Y);
                if (p.is_zero()) {                    // same x: P + P or P + (-P) (the point's slots are still intact)
                    if (r.is_zero()) {
                        const XYZZ<F> d = XYZZ<F>::dbl_affine(px, py);
                        S::st(base, S::X, d.x);
                        S::st(base, S::Y, d.y);
                        set_z(d.zz, d.zzz);
                    } else {
                        empty = true;
                    }
                    if (more) S::fetch_point(base, bases + (e_next & 0x7fffffffu));
                } else {
                    if (more) S::fetch_point(base, bases + (e_next & 0x7fffffffu));     // lands under the eight products below
                    const F pp = F::sqr_outline(p);
                    const F ppp = F::mul_outline(p, pp);
                    const F q = F::mul_outline(S::ld(base, S::X), pp);
                    if (ZS) {
                        S::st(base, SZZ, F::mul_outline(S::ld(base, SZZ), pp));
                        S::st(base, SZZZ, F::mul_outline(S::ld(base, SZZZ), ppp));
                    } else {
                        zz_r = F::mul_outline(zz_r, pp);
                        zzz_r = F::mul_outline(zzz_r, ppp);
                    }
                    const F x3 = F::sqr_outline(r) - ppp - q.dbl();
                    S::st(base, S::X, x3);
                    const F m1 = F::mul_outline(r, q - x3);
                    S::st(base, S::Y, m1 - F::mul_outline(S::ld(base, S::Y), ppp));
                }
            }
            if (!more) break;
            e = e_next;
            e_next = e_after;
        }
    }
    XYZZ<F> acc = XYZZ<F>::inf();
    if (!empty) acc = {S::ld(base, S::X), S::ld(base, S::Y), get_zz(), get_zzz()};
    buckets[id] = acc.store();
}

// block-wide sum of XYZZ values held one per thread (smem tree); result valid in thread 0
template <class F, int THREADS>
B200_DEV XYZZ<F> block_sum(XYZZ<F> acc, XYZZMem<F> *sm) {
    sm[threadIdx.x] = acc.store();
    __syncthreads();
    for (int s = THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            acc.add(XYZZ<F>::load(sm[threadIdx.x + s]));
            sm[threadIdx.x] = acc.store();
        }
        __syncthreads();
    }
    return acc;
}

// one block per over-populated bucket: order[] is sorted by population, so the first
// bin_counts[big] entries are exactly the buckets k_bucket_accumulate skipped.  Blocks stride over
// them; buckets of HUGE_BUCKET points or more are left to k_huge_buckets.
template <class F, int THREADS>
__global__ void __launch_bounds__(THREADS) k_big_buckets(const AffineMem<F> *__restrict__ bases,
                                                         const uint32_t *__restrict__ sorted,
                                                         const uint32_t *__restrict__ offsets,
                                                         const uint32_t *__restrict__ order,
                                                         const uint32_t *__restrict__ bin_counts, uint32_t big, int resume,
                                                         XYZZMem<F> *buckets) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t nbig = bin_counts[big];
    for (uint32_t b = blockIdx.x; b < nbig; b += gridDim.x) {
        uint32_t id = order[b];
        uint32_t lo = offsets[id], hi = offsets[id + 1];
        if (hi - lo >= HUGE_BUCKET) continue;
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t k = lo + threadIdx.x; k < hi; k += THREADS) {
            uint32_t e = __ldg(sorted + k);
            Affine<F> pt = Affine<F>::load(ldg_mem(bases + (e & 0x7fffffffu)));
            if (!pt.is_inf()) acc.madd(pt.x, pt.y.cneg(e >> 31));
        }
        acc = block_sum<F, THREADS>(acc, reinterpret_cast<XYZZMem<F> *>(smem_raw));
        if (threadIdx.x == 0) {
            if (resume) acc.add(XYZZ<F>::load(buckets[id]));
            buckets[id] = acc.store();
        }
        __syncthreads();
    }
}

// block (b, y): slice y of huge bucket order[b] -> slices[b * HUGE_SLICES + y]   (b < max_huge)
template <class F, int THREADS>
__global__ void __launch_bounds__(THREADS) k_huge_buckets(const AffineMem<F> *__restrict__ bases,
                                                          const uint32_t *__restrict__ sorted,
                                                          const uint32_t *__restrict__ offsets,
                                                          const uint32_t *__restrict__ order,
                                                          const uint32_t *__restrict__ bin_counts, uint32_t big,
                                                          XYZZMem<F> *__restrict__ slices) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t b = blockIdx.x;
    if (b >= bin_counts[big]) return;
    uint32_t id = order[b];
    uint32_t lo = offsets[id], hi = offsets[id + 1];
    if (hi - lo < HUGE_BUCKET) return;
    uint32_t per = (hi - lo + HUGE_SLICES - 1) / HUGE_SLICES;
    uint32_t s_lo = lo + blockIdx.y * per, s_hi = min(hi, s_lo + per);
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t k = s_lo + threadIdx.x; k < s_hi; k += THREADS) {
        uint32_t e = __ldg(sorted + k);
        Affine<F> pt = Affine<F>::load(ldg_mem(bases + (e & 0x7fffffffu)));
        if (!pt.is_inf()) acc.madd(pt.x, pt.y.cneg(e >> 31));
    }
    acc = block_sum<F, THREADS>(acc, reinterpret_cast<XYZZMem<F> *>(smem_raw));
    if (threadIdx.x == 0) slices[(size_t)b * HUGE_SLICES + blockIdx.y] = acc.store();
}

// how many scalars are 0, 1, or neither: a Groth16 witness is mostly bits (crates/epoch-snark/src/api/prover.rs:78), and the
// window width should follow the scalars that actually reach the buckets (out[0] zeros, out[1] ones, out[2] the rest)
template <int SW>
__global__ void __launch_bounds__(256) k_scalar_census(const uint32_t *__restrict__ scalars, uint32_t n, uint32_t *__restrict__ out) {
    uint32_t zeros = 0, ones = 0, rest = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t *s = scalars + (size_t)i * SW;
        uint32_t hi = 0;
#pragma unroll
        for (int k = 1; k < SW; k++) hi |= __ldg(s + k);
        const uint32_t lo = __ldg(s);
        const bool z = (hi | lo) == 0, o = hi == 0 && lo == 1;
        zeros += z;
        ones += o;
        rest += !(z || o);
    }
    for (int off = 16; off; off >>= 1) {
        zeros += __shfl_down_sync(0xffffffffu, zeros, off);
        ones += __shfl_down_sync(0xffffffffu, ones, off);
        rest += __shfl_down_sync(0xffffffffu, rest, off);
    }
    if ((threadIdx.x & 31) == 0) {
        if (zeros) atomicAdd(out, zeros);
        if (ones) atomicAdd(out + 1, ones);
        if (rest) atomicAdd(out + 2, rest);
    }
}

// unit scalars: ONES_PARTS strided partial sums of the listed bases (weight 1, added after Horner), folded to ONES_GROUPS
// sums by k_ones_fold before the window-sum kernel adds those up (a witness holds millions of ones: 8192 parts meant
// chains of 150 dependent additions and a 8192-term sum on one block -- 8.7 ms per BW6-761 MSM, now about 1 ms)
constexpr uint32_t ONES_PARTS = 32768;
constexpr uint32_t ONES_GROUPS = 256;
template <class F, int THREADS>
__global__ void __launch_bounds__(THREADS) k_ones_accumulate(const AffineMem<F> *__restrict__ bases,
                                                             const uint32_t *__restrict__ ones, int resume,
                                                             XYZZMem<F> *parts) {
    uint32_t t = blockIdx.x * THREADS + threadIdx.x;
    if (t >= ONES_PARTS) return;
    uint32_t count = ones[0];
    XYZZ<F> acc = resume ? XYZZ<F>::load(parts[t]) : XYZZ<F>::inf();
    for (uint32_t k = t; k < count; k += ONES_PARTS) {
        Affine<F> pt = Affine<F>::load(ldg_mem(bases + __ldg(ones + 1 + k)));
        if (!pt.is_inf()) acc.madd(pt.x, pt.y);
    }
    parts[t] = acc.store();
}

// ---- bucket reduction (quad-cooperative: 4 lanes per point operation, see ec.cuh) -----------
// k * p for small k (double-and-add, MSB first)
template <class F, class QT>
B200_DEV XYZZ<F> quad_small_mul(const QT &Q, const XYZZ<F> &p, uint32_t k) {
    XYZZ<F> r = XYZZ<F>::inf();
    for (int b = 31 - __clz(k | 1u); b >= 0; b--) {
        quad_dbl(Q, r);
        if ((k >> b) & 1u) quad_add(Q, r, p);
    }
    return r;
}

// quad b: bucket order[b] = sum of its HUGE_SLICES slice sums (huge buckets only)
template <class F, int THREADS>
__global__ void __launch_bounds__(THREADS) k_huge_finish(const XYZZMem<F> *__restrict__ slices,
                                                         const uint32_t *__restrict__ offsets,
                                                         const uint32_t *__restrict__ order,
                                                         const uint32_t *__restrict__ bin_counts, uint32_t big, uint32_t max_huge,
                                                         int resume, XYZZMem<F> *buckets) {
    const Quad Q;
    uint32_t b = (blockIdx.x * THREADS + threadIdx.x) >> 2;
    if (b >= max_huge || b >= bin_counts[big]) return;
    uint32_t id = order[b];
    if (offsets[id + 1] - offsets[id] < HUGE_BUCKET) return;
    XYZZ<F> acc = resume ? XYZZ<F>::load(buckets[id]) : XYZZ<F>::inf();
    for (uint32_t y = 0; y < HUGE_SLICES; y++) quad_add(Q, acc, XYZZ<F>::load(ldg_mem(slices + (size_t)b * HUGE_SLICES + y)));
    if (Q.q == 0) buckets[id] = acc.store();
}

// quad (w, seg): partial = sum_{j < L} (seg*L + j + 1) * B[w][seg*L + j]
template <class F, int THREADS, class QT = Quad>
__global__ void __launch_bounds__(THREADS) k_bucket_reduce(const XYZZMem<F> *__restrict__ buckets, MsmPlan p, int w_lo, int w_hi,
                                                           XYZZMem<F> *__restrict__ partials) {
    const QT Q;
    uint32_t t = (blockIdx.x * THREADS + threadIdx.x) >> 2;
    uint32_t total = (uint32_t)(w_hi - w_lo) * p.segs;
    if (t >= total) return;                          // whole quads leave together (THREADS % 4 == 0)
    t += (uint32_t)w_lo * p.segs;                    // windows [w_lo, w_hi) only (window groups, see msm_split_tail)
    uint32_t w = t / p.segs, seg = t % p.segs;
    const XYZZMem<F> *b = buckets + (size_t)w * p.nb + (size_t)seg * p.seg_len;
    XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
    for (int j = p.seg_len - 1; j >= 0; j--) {
        quad_add(Q, run, XYZZ<F>::load(ldg_mem(b + j)));
        quad_add(Q, acc, run);
    }
    if (seg) quad_add(Q, acc, quad_small_mul(Q, run, seg * (uint32_t)p.seg_len));
    if (Q.q == 0) partials[t] = acc.store();
}

// The same partial, ONE THREAD per segment: the throughput form.  A quad gives a point operation the latency of 3-4 product
// rounds, which is what the 8704 segments of a 2^20 BLS12-377 MSM need; but the 86 k segments of a BW6-761 MSM at n = 2^22
// are nine waves of quads, and there the idle lanes of the short rounds, the shuffles and 7.5 KB of spills per thread cost
// more than latency buys (24.7 ms for 10 ms of multiplier time).  Here every thread walks its segment with the out-of-line
// XYZZ addition / doubling (exact in every exceptional case): no idle lanes, no shuffles.
// MIN_BLOCKS > 1 caps the registers so that more blocks are resident: every thread does the same work, so the launch takes
// ceil(blocks / resident slots) rounds and one more resident block per SM can save a whole round (curve_impl.cuh).
template <class F, int THREADS, int MIN_BLOCKS = 0>                     // 0: no cap (ptxas treats it as unspecified)
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) k_bucket_reduce_thread(const XYZZMem<F> *__restrict__ buckets, MsmPlan p, int w_lo,
                                                                              int w_hi, XYZZMem<F> *__restrict__ partials) {
    uint32_t t = blockIdx.x * THREADS + threadIdx.x;
    const uint32_t total = (uint32_t)(w_hi - w_lo) * p.segs;
    if (t >= total) return;
    t += (uint32_t)w_lo * p.segs;
    const uint32_t w = t / p.segs, seg = t % p.segs;
    const XYZZMem<F> *b = buckets + (size_t)w * p.nb + (size_t)seg * p.seg_len;
    XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
#pragma unroll 1
    for (int j = p.seg_len - 1; j >= 0; j--) {
        run.add(XYZZ<F>::load(ldg_mem(b + j)));
        launder(run);
        acc.add(run);
        launder(acc);
    }
    if (seg) {                                       // + (seg * seg_len) * run: double-and-add, MSB first
        const uint32_t k = seg * (uint32_t)p.seg_len;
        XYZZ<F> r = XYZZ<F>::inf();
#pragma unroll 1
        for (int bit = 31 - __clz(k | 1u); bit >= 0; bit--) {
            r.dbl();
            launder(r);
            if ((k >> bit) & 1u) {
                r.add(run);
                launder(r);
            }
        }
        acc.add(r);
        launder(acc);
    }
    partials[t] = acc.store();
}

// block w: window_sums[w] = sum of the window's partials (THREADS / 4 quads, tree in smem)
template <class F, int THREADS>
__global__ void __launch_bounds__(THREADS) k_window_sum(const XYZZMem<F> *__restrict__ partials, MsmPlan p, int w_lo, int w_hi,
                                                        XYZZMem<F> *__restrict__ window_sums) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    XYZZMem<F> *sm = reinterpret_cast<XYZZMem<F> *>(smem_raw);
    constexpr int QUADS = THREADS / 4;
    const Quad Q;
    const int quad = threadIdx.x >> 2;
    // blocks 0 .. (w_hi - w_lo) - 1: the segment partials of window w_lo + block; one block more: the unit-scalar partials
    const bool ones_block = blockIdx.x == (uint32_t)(w_hi - w_lo);
    const uint32_t window = ones_block ? (uint32_t)p.windows : (uint32_t)w_lo + blockIdx.x;
    const XYZZMem<F> *src = partials + (size_t)window * p.segs + (ones_block ? ONES_PARTS : 0u);   // the folded unit-scalar sums
    const uint32_t count = ones_block ? ONES_GROUPS : p.segs;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t i = quad; i < count; i += QUADS) quad_add(Q, acc, XYZZ<F>::load(ldg_mem(src + i)));
    if (Q.q == 0) sm[quad] = acc.store();
    __syncthreads();
    for (int s = QUADS / 2; s > 0; s >>= 1) {
        if (quad < s) {
            quad_add(Q, acc, XYZZ<F>::load(sm[quad + s]));
            if (Q.q == 0) sm[quad] = acc.store();
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) window_sums[window] = acc.store();
}

// block g: folded[g] = sum of parts[g * (ONES_PARTS / ONES_GROUPS) ...) (quads + shared-memory tree, as k_window_sum)
template <class F, int THREADS>
__global__ void __launch_bounds__(THREADS) k_ones_fold(const XYZZMem<F> *__restrict__ parts, XYZZMem<F> *__restrict__ folded) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    XYZZMem<F> *sm = reinterpret_cast<XYZZMem<F> *>(smem_raw);
    constexpr int QUADS = THREADS / 4;
    constexpr uint32_t PER = ONES_PARTS / ONES_GROUPS;
    const Quad Q;
    const int quad = threadIdx.x >> 2;
    const XYZZMem<F> *src = parts + (size_t)blockIdx.x * PER;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t i = quad; i < PER; i += QUADS) quad_add(Q, acc, XYZZ<F>::load(ldg_mem(src + i)));
    if (Q.q == 0) sm[quad] = acc.store();
    __syncthreads();
    for (int s = QUADS / 2; s > 0; s >>= 1) {
        if (quad < s) {
            quad_add(Q, acc, XYZZ<F>::load(sm[quad + s]));
            if (Q.q == 0) sm[quad] = acc.store();
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) folded[blockIdx.x] = acc.store();
}

// one quad: Horner over the windows, high to low; result leaves as an arkworks GroupProjective
template <class F>
__global__ void __launch_bounds__(32) k_window_combine(const XYZZMem<F> *__restrict__ window_sums, MsmPlan p,
                                                       JacobianMem<F> *__restrict__ out) {
    if (blockIdx.x || threadIdx.x >= 4) return;
    const Quad Q;
    XYZZ<F> total = XYZZ<F>::inf();
    for (int w = p.windows - 1; w >= 0; w--) {
        quad_add(Q, total, XYZZ<F>::load(ldg_mem(window_sums + w)));
        if (w)
            for (int k = 0; k < p.c; k++) quad_dbl(Q, total);
    }
    quad_add(Q, total, XYZZ<F>::load(ldg_mem(window_sums + p.windows)));      // bases with scalar == 1
    if (Q.q == 0) *out = total.to_jacobian().to_ark();
}

// ---- layout + radix conversion: arkworks GroupAffine records -> native packed images ---------
// src records: x | y [| u8 infinity | pad] at `stride` bytes (4-byte aligned); infinity -> (0, 0).
template <class F, int THREADS>
__global__ void __launch_bounds__(THREADS) k_pack_bases(const uint32_t *__restrict__ src, uint32_t n,
                                                        uint32_t stride_words, int has_flag,
                                                        AffineMem<F> *__restrict__ dst) {
    uint32_t i = blockIdx.x * THREADS + threadIdx.x;
    if (i >= n) return;
    constexpr int WORDS = sizeof(AffineMem<F>) / 4;
    const uint32_t *s = src + (size_t)i * stride_words;
    AffineMem<F> img;
    uint32_t *iw = reinterpret_cast<uint32_t *>(&img);
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < WORDS; k++) {
        iw[k] = __ldg(s + k);
        any |= iw[k];
    }
    bool inf = (has_flag && (__ldg(s + WORDS) & 0xffu)) || any == 0;
    Affine<F> pt = inf ? Affine<F>{F::zero(), F::zero()} : Affine<F>::from_ark(img);
    dst[i] = pt.store();
}

// out = sum of `count` Jacobian points in arkworks radix: the multi-GPU partial combine (a handful of points) and
// Signature::aggregate / PublicKey::aggregate over a batch (thousands).  One block: every thread sums a strided
// subset, a shared-memory tree folds the 128 partial sums (count / 128 + 7 dependent additions instead of count).
constexpr int SUM_THREADS = 128;
template <class F>
__global__ void __launch_bounds__(SUM_THREADS) k_sum_jacobian(const JacobianMem<F> *__restrict__ pts, uint32_t count,
                                                              JacobianMem<F> *__restrict__ out, uint32_t stride = 1) {
    __shared__ JacobianMem<F> part[SUM_THREADS];
    const int t = threadIdx.x;
    // block b: out[b] = sum_i pts[i * stride + b]  (one block: the plain sum; a grid of `stride` blocks: the combine of a
    // whole batch of sharded MSMs after ONE all-gather, rank-major records of `stride` partials each)
    pts += blockIdx.x;
    out += blockIdx.x;
    Jacobian<F> total = Jacobian<F>::inf();
    launder(total);
    for (uint32_t i = t; i < count; i += SUM_THREADS) {
        total.add(Jacobian<F>::from_ark(pts[(size_t)i * stride]));
        launder(total);
    }
    part[t] = total.to_ark();
    __syncthreads();
    for (int off = SUM_THREADS / 2; off >= 1; off >>= 1) {
        if (t < off && (uint32_t)(t + off) < count) {      // partial sums beyond `count` are the identity
            total.add(Jacobian<F>::from_ark(part[t + off]));
            launder(total);
            part[t] = total.to_ark();
        }
        __syncthreads();
    }
    if (t == 0) *out = total.to_ark();
}

// out[i] = scalars[i] * base (double-and-add); synthesises benchmark / test bases on device
template <class F, int SW, int THREADS>
__global__ void __launch_bounds__(THREADS) k_fixed_base_mul(const AffineMem<F> *__restrict__ base,
                                                            const uint32_t *__restrict__ scalars, uint32_t n,
                                                            XYZZMem<F> *__restrict__ out) {
    uint32_t i = blockIdx.x * THREADS + threadIdx.x;
    if (i >= n) return;
    Affine<F> g = Affine<F>::from_ark(ldg_mem(base));
    XYZZ<F> r = XYZZ<F>::inf();
    for (int w = SW - 1; w >= 0; w--) {
        uint32_t word = __ldg(scalars + (size_t)i * SW + w);
        for (int b = 31; b >= 0; b--) {
            r.dbl();
            launder(r);
            if ((word >> b) & 1u) {
                r.madd(g.x, g.y);
                launder(r);
            }
        }
    }
    out[i] = r.store();
}

template <class F> struct FieldInv;

// ---- many small MSMs in one launch ------------------------------------------------------------------------------
// batch_verify_strict hands over hundreds of batches of a few dozen signatures (the reference's own benchmark: 300 epochs
// x 20 validators, crates/bls-crypto/benches/batch_bls.rs:62-95); each is two tiny MSMs (signature.rs:85, public.rs:61)
// for which the bucket method's sort / reduce / Horner stages cost more than they save.  One block per MSM: thread i
// multiplies point i by its scalar (double-and-add over the top `bits` bits, XYZZ), a shared-memory tree adds the block
// up, thread 0 normalises.  out record of MSM b: out[b * out_stride + out_slot] (packed affine, (0, 0) = infinity).
template <class F, int SW, int THREADS>
__global__ void __launch_bounds__(THREADS) k_small_msm(const AffineMem<F> *__restrict__ bases, const uint32_t *__restrict__ scalars,
                                                       const uint32_t *__restrict__ offsets, int bits,
                                                       AffineMem<F> *__restrict__ out, uint32_t out_stride, uint32_t out_slot) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t lo = offsets[blockIdx.x], hi = offsets[blockIdx.x + 1];
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t i = lo + threadIdx.x; i < hi; i += THREADS) {
        Affine<F> g = Affine<F>::from_ark(ldg_mem(bases + i));
        if (g.is_inf()) continue;
        XYZZ<F> r = XYZZ<F>::inf();
        for (int b = bits - 1; b >= 0; b--) {
            r.dbl();
            launder(r);
            if ((__ldg(scalars + (size_t)i * SW + (b >> 5)) >> (b & 31)) & 1u) {
                r.madd(g.x, g.y);
                launder(r);
            }
        }
        acc.add(r);
        launder(acc);
    }
    acc = block_sum<F, THREADS>(acc, reinterpret_cast<XYZZMem<F> *>(smem_raw));
    if (threadIdx.x == 0) {
        Affine<F> r = {F::zero(), F::zero()};
        if (!acc.is_inf()) {
            F iv = FieldInv<F>::inv(acc.zz * acc.zzz);
            r.x = acc.x * (iv * acc.zzz);
            r.y = acc.y * (iv * acc.zz);
        }
        out[(size_t)blockIdx.x * out_stride + out_slot] = r.to_ark();
    }
}

// The same with a QUAD per point (quad-cooperative doubling / addition, ec.cuh): a doubling is 3 product rounds instead of
// 9 sequential products, so the chain of `bits` doublings that paces the launch is three times shorter -- 300 batches of
// 20 keys over Fq2: 11.8 -> ~4 ms.  THREADS / 4 points per pass; the block sum is a quad tree in shared memory.
template <class F, int SW, int THREADS>
__global__ void __launch_bounds__(THREADS) k_small_msm_quad(const AffineMem<F> *__restrict__ bases, const uint32_t *__restrict__ scalars,
                                                            const uint32_t *__restrict__ offsets, int bits,
                                                            AffineMem<F> *__restrict__ out, uint32_t out_stride, uint32_t out_slot) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    XYZZMem<F> *sm = reinterpret_cast<XYZZMem<F> *>(smem_raw);
    constexpr int QUADS = THREADS / 4;
    const QuadShared Q;
    const int quad = threadIdx.x >> 2;
    const uint32_t lo = offsets[blockIdx.x], hi = offsets[blockIdx.x + 1];
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t i = lo + quad; i < hi; i += QUADS) {
        Affine<F> g = Affine<F>::from_ark(ldg_mem(bases + i));
        if (g.is_inf()) continue;                    // uniform over the quad
        const XYZZ<F> gp = {g.x, g.y, F::one(), F::one()};
        XYZZ<F> r = XYZZ<F>::inf();
        for (int b = bits - 1; b >= 0; b--) {
            quad_dbl(Q, r);
            if ((__ldg(scalars + (size_t)i * SW + (b >> 5)) >> (b & 31)) & 1u) quad_add(Q, r, gp);
        }
        quad_add(Q, acc, r);
    }
    if (Q.q == 0) sm[quad] = acc.store();
    __syncthreads();
    for (int s = QUADS / 2; s > 0; s >>= 1) {
        if (quad < s) {
            quad_add(Q, acc, XYZZ<F>::load(sm[quad + s]));
            if (Q.q == 0) sm[quad] = acc.store();
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        Affine<F> r = {F::zero(), F::zero()};
        if (!acc.is_inf()) {
            F iv = FieldInv<F>::inv(acc.zz * acc.zzz);
            r.x = acc.x * (iv * acc.zzz);
            r.y = acc.y * (iv * acc.zz);
        }
        out[(size_t)blockIdx.x * out_stride + out_slot] = r.to_ark();
    }
}

// out[t * run + j] = (scalars[t] + j) * base for j < run: `run` consecutive multiples behind one double-and-add, so a
// synthetic base array of 2^24 distinct points costs ~1 / run of k_fixed_base_mul (bench.py, configs 4 and 5)
template <class F, int SW, int THREADS>
__global__ void __launch_bounds__(THREADS) k_point_runs(const AffineMem<F> *__restrict__ base,
                                                        const uint32_t *__restrict__ scalars, uint32_t runs, uint32_t run,
                                                        XYZZMem<F> *__restrict__ out) {
    uint32_t t = blockIdx.x * THREADS + threadIdx.x;
    if (t >= runs) return;
    Affine<F> g = Affine<F>::from_ark(ldg_mem(base));
    XYZZ<F> r = XYZZ<F>::inf();
    for (int w = SW - 1; w >= 0; w--) {
        uint32_t word = __ldg(scalars + (size_t)t * SW + w);
        for (int b = 31; b >= 0; b--) {
            r.dbl();
            launder(r);
            if ((word >> b) & 1u) {
                r.madd(g.x, g.y);
                launder(r);
            }
        }
    }
    for (uint32_t j = 0; j < run; j++) {
        out[(size_t)t * run + j] = r.store();
        r.madd(g.x, g.y);
        launder(r);
    }
}

}  // namespace b200

namespace b200 {

// ---- projective -> affine -------------------------------------------------------------------
template <class F> struct FieldInv;
template <class P> struct FieldInv<Fp<P>> {
    B200_DEV static Fp<P> inv(const Fp<P> &a) { return a.inv(); }
};
template <class B> struct FieldInv<Fp2<B>> {
    // 1/(a0 + a1 u) = (a0 - a1 u) / (a0^2 + 5 a1^2)
    B200_DEV static Fp2<B> inv(const Fp2<B> &a) {
        B n = (a.c0.sqr() + Fp2<B>::mul5(a.c1.sqr())).inv();
        return {a.c0 * n, (a.c1 * n).neg()};
    }
};

// native XYZZ images -> arkworks-radix packed affine records
template <class F, int THREADS>
__global__ void __launch_bounds__(THREADS) k_xyzz_to_affine(const XYZZMem<F> *__restrict__ in, uint32_t n,
                                                            AffineMem<F> *__restrict__ out) {
    uint32_t i = blockIdx.x * THREADS + threadIdx.x;
    if (i >= n) return;
    XYZZ<F> p = XYZZ<F>::load(ldg_mem(in + i));
    Affine<F> r = {F::zero(), F::zero()};
    if (!p.is_inf()) {
        F iv = FieldInv<F>::inv(p.zz * p.zzz);      // 1/ZZ = iv * ZZZ, 1/ZZZ = iv * ZZ
        r.x = p.x * (iv * p.zzz);
        r.y = p.y * (iv * p.zz);
    }
    out[i] = r.to_ark();
}

// One inversion per thread, shared by BATCH consecutive points (Montgomery's trick):
// the device form of batch_normalization_into_affine (signature.rs:82, public.rs:58).
// Input: arkworks GroupProjective images; output: arkworks-radix packed affine records.
template <class F, int THREADS, int BATCH>
__global__ void __launch_bounds__(THREADS) k_jacobian_to_affine(const JacobianMem<F> *__restrict__ in, uint32_t n,
                                                                AffineMem<F> *__restrict__ out) {
    uint32_t t = blockIdx.x * THREADS + threadIdx.x;
    uint32_t first = t * BATCH;
    if (first >= n) return;
    uint32_t cnt = min((uint32_t)BATCH, n - first);
    F prefix[BATCH];                                // prefix[k] = prod of non-zero z[0..k]
    F run = F::one();
    for (uint32_t k = 0; k < cnt; k++) {
        F z = F::from_ark(ldg_mem(&in[first + k].z));
        if (!z.is_zero()) run = run * z;
        launder(run);
        prefix[k] = run;
    }
    F iv = FieldInv<F>::inv(run);
    for (int k = (int)cnt - 1; k >= 0; k--) {
        Jacobian<F> p = Jacobian<F>::from_ark(ldg_mem(in + first + k));
        Affine<F> r = {F::zero(), F::zero()};
        if (!p.z.is_zero()) {
            F zi = iv;                               // 1 / z_k
            if (k) zi = zi * prefix[k - 1];
            iv = iv * p.z;
            launder(iv);
            F zi2 = zi.sqr();
            r.x = p.x * zi2;
            r.y = p.y * (zi2 * zi);
        }
        out[first + k] = r.to_ark();
    }
}

// ---- element-wise field ops (parity tests of the field layer against the CPU restatement) ----
enum FieldOp { FOP_ADD = 0, FOP_SUB = 1, FOP_MUL = 2, FOP_SQR = 3, FOP_INV = 4, FOP_NEG = 5, FOP_DBL = 6 };

template <class F>
__global__ void __launch_bounds__(64) k_field_op(int op, const typename F::Mem *__restrict__ a,
                                                 const typename F::Mem *__restrict__ b, uint32_t n,
                                                 typename F::Mem *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = F::from_ark(ldg_mem(a + i)), y = F::from_ark(ldg_mem(b + i)), r;
    switch (op) {
    case FOP_ADD: r = x + y; break;
    case FOP_SUB: r = x - y; break;
    case FOP_MUL: r = x * y; break;
    case FOP_SQR: r = Shared<F>::sqr(x); break;      // the dedicated squaring where the field has one
    case FOP_INV: r = FieldInv<F>::inv(x); break;
    case FOP_NEG: r = x.neg(); break;
    default: r = x.dbl(); break;
    }
    out[i] = r.to_ark();
}

}  // namespace b200

#include "msm_afftree.cuh"
