// Bucket accumulation as a TREE OF BATCHED AFFINE ADDITIONS over the whole input (round 2).
//
// The XYZZ accumulate kernel (msm.cuh) is at 86-93 % of what IMAD.WIDE.X itself retires (profiles/r2_experiments.md): the
// only lever left is the number of field products per point.  An affine addition with a shared inversion costs
// 5 M + 1 S (one prefix product, two to peel the shared inverse, lambda, lambda^2, y3) against 8 M + 2 S for the XYZZ
// mixed addition.  Round 1 tried this with one thread per bucket (k_bucket_accumulate_affine): a bucket of 64 points
// gives batches of 32, 16, 8, ... pairs -- ten additions per inversion -- and every thread walks its own scratch.
// Here the tree is built LEVEL BY LEVEL OVER ALL BUCKETS AT ONCE:
//
//   level 0:      the points of bucket b are sorted[offsets[b] .. offsets[b + 1])          (m = its population)
//   level L + 1:  element j of bucket b = element 2j + element 2j + 1 of level L (or a copy of the last odd one),
//                 ceil(m / 2^(L+1)) elements, stored at slot ((lo_b + (2^(L+1) - 1) b) >> (L + 1)) + j
//                 (lo_b = offsets[b] - offsets[first bucket], b counted from the first bucket: closed form, regions never overlap)
//
// and a thread takes B = 64 CONSECUTIVE OUTPUT SLOTS, whatever buckets they belong to: every inversion (Bernstein-Yang
// safegcd, ~20 k ALU instructions per warp) is shared by 64 additions at every level, outputs are contiguous, and the
// inputs of levels >= 1 are contiguous pairs.  After LEVELS levels a bucket holds ceil(m / 2^LEVELS) affine points;
// k_bucket_finish adds those (a handful) into the XYZZ bucket with the ordinary mixed addition -- one thread per bucket,
// population order, as before -- which is also where a chunk-fed MSM resumes its bucket set.
// Exceptional pairs are exact (aff_denominator / aff_add_with_inverse of msm.cuh): P + P takes the tangent, P + (-P)
// and infinite operands use denominator 1; infinity is (0, 0).
// Buckets holding >= `big` points are skipped (k_big_buckets / k_huge_buckets own them, as with the XYZZ kernel).
#pragma once

namespace b200 {

constexpr int AFT_B = 64;                           // output slots per thread = additions per inversion

// first slot of bucket b_rel at level L (lo_rel: its first sorted position relative to the group's first bucket)
B200_DEV uint32_t aft_start(uint32_t lo_rel, uint32_t b_rel, int L) {
    return (uint32_t)(((uint64_t)lo_rel + ((((uint64_t)1) << L) - 1u) * (uint64_t)b_rel) >> L);
}
inline size_t aft_slots(size_t entries, size_t nbuckets, int L) {                     // host: slots of level L (upper bound)
    return ((entries + ((((size_t)1) << L) - 1) * nbuckets) >> L) + 1;
}

// One level: reads level L (L = 0: sorted / bases; else `in`), writes level L + 1 to `out`.
// b0 / nbk: the bucket range of this launch (window group); out_slots = aft_slots(entries of the range, nbk, L + 1).
template <class F, int THREADS, int MIN_BLOCKS, bool PF>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
k_affine_level(const AffineMem<F> *__restrict__ bases, const uint32_t *__restrict__ sorted, const uint32_t *__restrict__ offsets,
               uint32_t b0, uint32_t nbk, uint32_t big, int L, const AffineMem<F> *__restrict__ in, AffineMem<F> *__restrict__ out,
               uint32_t out_slots) {
    const uint32_t s0 = (blockIdx.x * THREADS + threadIdx.x) * (uint32_t)AFT_B;
    if (s0 >= out_slots) return;
    const uint32_t lo0 = __ldg(offsets + b0);
    const AffineMem<F> *src = L == 0 ? bases : in;
    // bucket of slot s0: the last b with start_{L+1}(b) <= s0
    uint32_t b = 0;
    {
        uint32_t lo = 0, hi = nbk;                  // invariant: start(lo) <= s0 < start(hi) (start(nbk) = +inf)
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (aft_start(__ldg(offsets + b0 + mid) - lo0, mid, L + 1) <= s0) lo = mid;
            else hi = mid;
        }
        b = lo;
    }
    uint32_t off_b = __ldg(offsets + b0 + b), off_n = __ldg(offsets + b0 + b + 1);
    uint32_t start_out = aft_start(off_b - lo0, b, L + 1);
    uint32_t next_out = b + 1 < nbk ? aft_start(off_n - lo0, b + 1, L + 1) : 0xffffffffu;

    typename F::Mem pre[AFT_B];                     // pre[k] = product of the denominators of slots < k
    uint32_t ea[AFT_B], eb[AFT_B];                  // operand records: index into src (| sign << 31 at level 0)
    uint8_t kind[AFT_B];                            // 0 nothing, 1 copy, 2 addition

    auto load_pt = [&](uint32_t e) -> Affine<F> {
        Affine<F> p = Affine<F>::load(ldg_mem(src + (e & 0x7fffffffu)));
        if (L == 0) p.y = p.y.cneg(e >> 31);        // -(0) = 0: the infinity encoding survives
        return p;
    };

    // ---- describe the slots (integer work; the index loads of level 0 are independent of each other) ----
#pragma unroll 4
    for (int k = 0; k < AFT_B; k++) {
        const uint32_t s = s0 + (uint32_t)k;
        while (s >= next_out) {                     // walk to the bucket of this slot
            b++;
            off_b = off_n;
            off_n = __ldg(offsets + b0 + b + 1);
            start_out = next_out;
            next_out = b + 1 < nbk ? aft_start(off_n - lo0, b + 1, L + 1) : 0xffffffffu;
        }
        uint32_t m0 = off_n - off_b;
        if (m0 >= big) m0 = 0;                      // block-summed buckets are not ours
        const uint32_t mL = (m0 + ((1u << L) - 1u)) >> L, j = s - start_out;
        uint8_t kd = 0;
        uint32_t e0 = 0, e1 = 0;
        if (s < out_slots && 2 * j < mL) {
            const bool pair = 2 * j + 1 < mL;
            if (L == 0) {
                e0 = __ldg(sorted + off_b + 2 * j);
                if (pair) e1 = __ldg(sorted + off_b + 2 * j + 1);
            } else {
                e0 = aft_start(off_b - lo0, b, L) + 2 * j;
                e1 = e0 + 1;
            }
            kd = pair ? 2 : 1;
        }
        ea[k] = e0;
        eb[k] = e1;
        kind[k] = kd;
    }
    // ---- pass 1: denominators and prefix products; the x coordinates of slot k + 1 are in flight during product k ----
    using XM = typename F::Mem;
    F run = F::one();
    XM xa_n, xb_n;
    if (PF && kind[0] == 2) {
        xa_n = ldg_mem(&src[ea[0] & 0x7fffffffu].x);
        xb_n = ldg_mem(&src[eb[0] & 0x7fffffffu].x);
    }
#pragma unroll 1
    for (int k = 0; k < AFT_B; k++) {
        if (!PF && kind[k] == 2) {
            xa_n = ldg_mem(&src[ea[k] & 0x7fffffffu].x);
            xb_n = ldg_mem(&src[eb[k] & 0x7fffffffu].x);
        }
        const XM xa_m = xa_n, xb_m = xb_n;
        if (PF && k + 1 < AFT_B && kind[k + 1] == 2) {
            xa_n = ldg_mem(&src[ea[k + 1] & 0x7fffffffu].x);
            xb_n = ldg_mem(&src[eb[k + 1] & 0x7fffffffu].x);
        }
        pre[k] = run.store();
        if (kind[k] == 2) {
            // the denominator needs the x coordinates only, except for equal x and for x = 0, which may be the infinity encoding
            F xa = F::load(xa_m), xb = F::load(xb_m);
            F d = xb - xa;
            if (d.is_zero() || xa.is_zero() || xb.is_zero()) d = aff_denominator(load_pt(ea[k]), load_pt(eb[k]));
            run = Shared<F>::mul(run, d);
            launder(run);
        }
    }
    F inv = FieldInv<F>::inv(run);
    // ---- pass 2 (reverse): peel the inverse, add, store; the operands of slot k - 1 are in flight during slot k ----
    AffineMem<F> a_n, c_n;
    if (PF && kind[AFT_B - 1]) a_n = ldg_mem(src + (ea[AFT_B - 1] & 0x7fffffffu));
    if (PF && kind[AFT_B - 1] == 2) c_n = ldg_mem(src + (eb[AFT_B - 1] & 0x7fffffffu));
#pragma unroll 1
    for (int k = AFT_B - 1; k >= 0; k--) {
        if (!PF) {
            if (kind[k]) a_n = ldg_mem(src + (ea[k] & 0x7fffffffu));
            if (kind[k] == 2) c_n = ldg_mem(src + (eb[k] & 0x7fffffffu));
        }
        const AffineMem<F> a_m = a_n, c_m = c_n;
        if (PF && k > 0) {
            if (kind[k - 1]) a_n = ldg_mem(src + (ea[k - 1] & 0x7fffffffu));
            if (kind[k - 1] == 2) c_n = ldg_mem(src + (eb[k - 1] & 0x7fffffffu));
        }
        const uint8_t kd = kind[k];
        if (kd == 0) continue;
        Affine<F> a = Affine<F>::load(a_m);
        if (L == 0) a.y = a.y.cneg(ea[k] >> 31);
        if (kd == 2) {
            Affine<F> c = Affine<F>::load(c_m);
            if (L == 0) c.y = c.y.cneg(eb[k] >> 31);
            F di = Shared<F>::mul(inv, F::load(pre[k]));
            inv = Shared<F>::mul(inv, aff_denominator(a, c));
            launder(inv);
            a = aff_add_with_inverse(a, c, di);
        }
        out[s0 + (uint32_t)k] = a.store();
    }
}

// The rest of every bucket: its ceil(m / 2^L) affine elements of level L >= 1 are added
// into the XYZZ bucket, one thread per bucket in population order.  resume: the bucket already holds earlier chunks.
template <class F, int THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
k_bucket_finish(const AffineMem<F> *__restrict__ in, const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ order,
                uint32_t total_buckets, uint32_t b0, uint32_t nbk, uint32_t big, int L, int resume, XYZZMem<F> *__restrict__ buckets) {
    uint32_t t = blockIdx.x * THREADS + threadIdx.x;
    if (t >= total_buckets) return;
    const uint32_t id = order[t];
    if (id < b0 || id >= b0 + nbk) return;          // another window group's bucket
    const uint32_t off_b = offsets[id], m0 = offsets[id + 1] - off_b;
    if (m0 >= big) return;
    if (resume && m0 == 0) return;
    const uint32_t mL = (m0 + ((1u << L) - 1u)) >> L;
    XYZZ<F> acc = resume ? XYZZ<F>::load(buckets[id]) : XYZZ<F>::inf();
    const AffineMem<F> *p = in + aft_start(off_b - offsets[b0], id - b0, L);
    if (mL) {
        AffineMem<F> img = ldg_mem(p);
        for (uint32_t i = 0;;) {
            ++i;
            AffineMem<F> img_next;
            const bool more = i < mL;
            if (more) img_next = ldg_mem(p + i);
            Affine<F> pt = Affine<F>::load(img);
            if (!pt.is_inf()) acc = xyzz_madd_shared(acc, pt.x, pt.y);
            if (!more) break;
            img = img_next;
        }
    }
    buckets[id] = acc.store();
}

}  // namespace b200
