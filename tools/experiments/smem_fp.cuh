// EXPERIMENT, not shipped: field operands resident in SHARED MEMORY (measured on B200: 2.90 G mixed additions/s against
// 2.92 for the register kernel -- level; profiles/r2_experiments.md).
//
// Why: k_bucket_accumulate_shared (msm.cuh) keeps the XYZZ accumulator, the current and the prefetched point in
// registers (168) and sends every product through one out-of-line body.  The call ABI then moves ~50 registers per
// product, and ptxas turns about a third of those moves into IMAD.MOV.U32 -- which run on the same FMA-heavy pipe the
// kernel is bound by (tools/microbench_kara.cu: 4 cycles per IMAD.WIDE + 2 per IMAD.MOV explains the measured product
// rate to 1 %); the calls cost ~11 % of the pipe.  Here a thread's operands live in shared memory "slots" (227 KB per
// SM on sm_100a: 10 slots x 48 B x 384 threads = 180 KB), an operation names its operands by slot offset (three
// integers cross the call, no field element does), loads them with LDS.128, and stores its result with STS.128 -- the
// load/store pipe idles in this kernel.  The next bucket point is gathered global -> shared by cp.async (LDGSTS.128,
// no register staging: the 24 / 48 prefetch registers of the register kernel are gone) as soon as the current point's
// two products have consumed it, and lands under the remaining eight.
//
// Layout: slot s, 16-byte chunk k of thread t at ((s * CH + k) * THREADS + t) * 16 -- consecutive threads touch
// consecutive 16-byte words: conflict-free for 128-bit accesses.
#pragma once
#include "fp_kara.cuh"
#include "ec.cuh"

namespace b200 {

B200_DEV uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
B200_DEV void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" : : "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
B200_DEV void cp_async16(uint32_t saddr, const void *gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" : : "r"(saddr), "l"(gptr) : "memory");
}
B200_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" : : : "memory"); }
B200_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" : : : "memory"); }

template <class F, int THREADS>
struct Slots {
    static constexpr int N = F::N, CH = N / 4;
    static constexpr uint32_t STRIDE = THREADS * 16u;               // bytes between chunks of one element
    static constexpr uint32_t SLOT = CH * STRIDE;                   // bytes between slots
    static_assert(N % 4 == 0, "elements are whole 16-byte chunks");
    // slot numbers of the accumulation kernel
    enum : uint32_t { X = 0, Y = 1, ZZ = 2, ZZZ = 3, PX = 4, PY = 5, T0 = 6, T1 = 7, T2 = 8, T3 = 9, COUNT = 10 };
    static constexpr uint32_t BYTES = COUNT * SLOT;                 // per block

    B200_DEV static F ld(uint32_t base, uint32_t slot) {
        F r;
#pragma unroll
        for (int k = 0; k < CH; k++) {
            uint4 v = lds128(base + slot * SLOT + k * STRIDE);
            r.l[4 * k] = v.x;
            r.l[4 * k + 1] = v.y;
            r.l[4 * k + 2] = v.z;
            r.l[4 * k + 3] = v.w;
        }
        return r;
    }
    B200_DEV static void st(uint32_t base, uint32_t slot, const F &v) {
#pragma unroll
        for (int k = 0; k < CH; k++) sts128(base + slot * SLOT + k * STRIDE, make_uint4(v.l[4 * k], v.l[4 * k + 1], v.l[4 * k + 2], v.l[4 * k + 3]));
    }
    // element image in global memory -> slot, asynchronously
    B200_DEV static void fetch(uint32_t base, uint32_t slot, const typename F::Mem *g) {
#pragma unroll
        for (int k = 0; k < CH; k++) cp_async16(base + slot * SLOT + k * STRIDE, reinterpret_cast<const uint4 *>(g) + k);
    }
};

// The two product bodies of the kernel (one instance each per field): everything else is loads, stores and
// additions.  KARA selects the separated Karatsuba product of fp_kara.cuh (BASE = N / 2).
//   mode 0: d = a b          mode 1: d = a b - c, returns "d != 0"          mode 2: d = c - a b
template <class F, int THREADS, bool KARA>
__device__ __noinline__ uint32_t slot_mul(uint32_t base, uint32_t d, uint32_t a, uint32_t b, uint32_t c, int mode) {
    using S = Slots<F, THREADS>;
    F r;
    if constexpr (KARA) r = kara_mul_inline<F::N / 2>(S::ld(base, a), S::ld(base, b));
    else r = F::mul_inline(S::ld(base, a), S::ld(base, b));
    uint32_t nz = 1;
    if (mode == 1) {
        r = r - S::ld(base, c);
        nz = 0;
#pragma unroll
        for (int i = 0; i < F::N; i++) nz |= r.l[i];
    } else if (mode == 2) {
        r = S::ld(base, c) - r;
    }
    S::st(base, d, r);
    return nz;
}
//   mode 0: d = a^2          mode 1: d = a^2 - c1 - 2 c2
template <class F, int THREADS>
__device__ __noinline__ void slot_sqr(uint32_t base, uint32_t d, uint32_t a, uint32_t c1, uint32_t c2, int mode) {
    using S = Slots<F, THREADS>;
    F r = F::sqr_inline(S::ld(base, a));
    if (mode == 1) {
        r = r - S::ld(base, c1);
        r = r - S::ld(base, c2).dbl();
    }
    S::st(base, d, r);
}
template <class F, int THREADS>
B200_DEV void slot_sub(uint32_t base, uint32_t d, uint32_t a, uint32_t b) {
    using S = Slots<F, THREADS>;
    S::st(base, d, S::ld(base, a) - S::ld(base, b));
}

// acc (slots X, Y, ZZ, ZZZ; `inf` says the accumulator is still empty) += the finite affine point in slots PX, PY.
// madd-2008-s with the exceptional cases of XYZZ::madd.  `after_consume` runs once the point's slots are free (the
// caller issues the next gather there); it is called on every path exactly once.
template <class F, int THREADS, bool KARA, class Hook>
B200_DEV void slot_madd(uint32_t base, bool &inf, Hook after_consume) {
    using S = Slots<F, THREADS>;
    if (inf) {
        S::st(base, S::X, S::ld(base, S::PX));
        S::st(base, S::Y, S::ld(base, S::PY));
        S::st(base, S::ZZ, F::one());
        S::st(base, S::ZZZ, F::one());
        inf = false;
        after_consume();
        return;
    }
    const uint32_t p_nz = slot_mul<F, THREADS, KARA>(base, S::T0, S::PX, S::ZZ, S::X, 1);     // P = px zz - x
    const uint32_t r_nz = slot_mul<F, THREADS, KARA>(base, S::T1, S::PY, S::ZZZ, S::Y, 1);    // R = py zzz - y
    if (p_nz == 0) {                                                                         // same x: P + P or P - P
        if (r_nz == 0) {
            XYZZ<F> t = XYZZ<F>::dbl_affine(S::ld(base, S::PX), S::ld(base, S::PY));
            S::st(base, S::X, t.x);
            S::st(base, S::Y, t.y);
            S::st(base, S::ZZ, t.zz);
            S::st(base, S::ZZZ, t.zzz);
        } else {
            inf = true;
        }
        after_consume();
        return;
    }
    after_consume();
    slot_sqr<F, THREADS>(base, S::T2, S::T0, 0, 0, 0);                                       // PP
    slot_mul<F, THREADS, KARA>(base, S::T3, S::X, S::T2, 0, 0);                              // Q = x PP
    slot_mul<F, THREADS, KARA>(base, S::ZZ, S::ZZ, S::T2, 0, 0);                             // zz PP
    slot_mul<F, THREADS, KARA>(base, S::T0, S::T0, S::T2, 0, 0);                             // PPP
    slot_mul<F, THREADS, KARA>(base, S::ZZZ, S::ZZZ, S::T0, 0, 0);                           // zzz PPP
    slot_sqr<F, THREADS>(base, S::X, S::T1, S::T0, S::T3, 1);                                // X3 = R^2 - PPP - 2 Q
    slot_sub<F, THREADS>(base, S::T3, S::T3, S::X);                                          // Q - X3
    slot_mul<F, THREADS, KARA>(base, S::T2, S::T1, S::T3, 0, 0);                             // R (Q - X3)
    slot_mul<F, THREADS, KARA>(base, S::Y, S::Y, S::T0, S::T2, 2);                           // Y3 = R (Q - X3) - y PPP
}

}  // namespace b200
