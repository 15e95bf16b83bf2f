// Montgomery products with a SEPARATED operand product: t = a * b by Karatsuba over half-width
// schoolbook products, then one Montgomery reduction of the double-width t.
//
// EXPERIMENT, not shipped (measured on B200, profiles/r2_experiments.md + r2_microbench_kara.txt: 13-15 % fewer
// IMAD.WIDE, parity-green in the product / mixed-addition checks, but NO gain in products or mixed additions per
// second -- the added carry chains cost what the saved multiply-adds bought).  Kept with its host check and micro-benchmark.
//
// Why it was tried: the bucket accumulation is bound by the multiplier pipe
// (IMAD.WIDE holds it 4 cycles per warp instruction, fp.cuh), at 85-87 % busy.  The fused row-interleaved
// product of fp.cuh spends 2 N^2 wide multiply-adds per product; splitting the operand product from the
// reduction lets
//   * the operand product run as Karatsuba: 3 (N/2)^2 instead of N^2 (one level; 9 (N/4)^2 with two), the
//     extra additions go to the ALU pipe, which idles beside the multiplier;
//   * a SUM of two products be reduced once (lazy reduction): the mixed addition's
//     Y3 = R (Q - X3) - Y1 PPP = R (Q - X3) + (p - Y1) PPP pays one reduction instead of two
//     (2 p^2 + p R < R^2 / 32: the 7 spare bits of both moduli keep every bound of the fused product).
// Wide multiply-adds per XYZZ mixed addition, 12 limbs: 8 x 288 + 2 x 222 = 2748 -> 8 x 108 + 9 x 144 + 2 x 78 = 2316.
//
// Register-pair discipline (as in fp.cuh): a wide product lands on an even-aligned register pair, so
// partial products whose low limb is even accumulate in one array (E) and the odd ones in another (O,
// one limb up); the two are added once at the end.  Every chain below touches pairs (2k, 2k + 1) only.
//
// The carry-chain primitives are re-declared here (namespace kara) so that tools/experiments/kara_host_check.cpp can
// swap in a portable emulation (B200_KARA_HOST_EMU) and check the index logic against big integers on the
// CPU before a GPU trip; the product path only ever compiles the PTX forms.
#pragma once
#include <cstdint>

namespace b200 {
namespace kara {

#ifndef B200_KARA_HOST_EMU
#define KDEV __device__ __forceinline__
KDEV void add_cc(uint32_t &r, uint32_t a, uint32_t b) { asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
KDEV void addc_cc(uint32_t &r, uint32_t a, uint32_t b) { asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
KDEV void addc(uint32_t &r, uint32_t a, uint32_t b) { asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
KDEV void sub_cc(uint32_t &r, uint32_t a, uint32_t b) { asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
KDEV void subc_cc(uint32_t &r, uint32_t a, uint32_t b) { asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
KDEV void subc(uint32_t &r, uint32_t a, uint32_t b) { asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
KDEV uint32_t mul_lo(uint32_t a, uint32_t b) {
    uint32_t r;
    asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
// (lo, hi) = a * b
KDEV void wmul(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b) {
    asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
// (lo, hi) += a * b [+ CC]; carry out in CC
template <bool CIN>
KDEV void wmad(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b) {
    if (CIN) asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
    else asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (lo, hi) = a * b + (c_lo, c_hi) [+ CC]; carry out in CC
template <bool CIN>
KDEV void wmad4(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b, uint32_t c_lo, uint32_t c_hi) {
    if (CIN) asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;" : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(c_lo), "r"(c_hi));
    else asm volatile("mad.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;" : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(c_lo), "r"(c_hi));
}
// top of a chain: (lo, hi) = a * b + (c_lo, 0) [+ CC]; the high word cannot overflow, no carry out
template <bool CIN>
KDEV void wmad_top(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b, uint32_t c_lo) {
    if (CIN) asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.u32 %1, %2, %3, 0;" : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(c_lo));
    else asm volatile("mad.lo.cc.u32 %0, %2, %3, %4; madc.hi.u32 %1, %2, %3, 0;" : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(c_lo));
}
#endif  // !B200_KARA_HOST_EMU (the emulation supplies the same names)

// ---- t[0 .. 2n) = a * b, schoolbook, n even ------------------------------------------------------------------
// E[k] holds weight 2^(32 k), O[k] weight 2^(32 (k + 1)).  Row i adds a * b_i at limb i: the products whose limb
// i + j is even chain through E, the others through O.  Rows extend the arrays by one limb each: the chain that
// ends inside written limbs leaves its carry in a fresh limb, the other one writes a fresh high word that cannot
// overflow ((2^32 - 1)^2 + 2 (2^32 - 1) < 2^64).
template <int n, int I>
KDEV void school_rows(uint32_t (&E)[2 * n], uint32_t (&O)[2 * n], const uint32_t (&a)[n], const uint32_t (&b)[n]) {
    if constexpr (I < n) {
        const uint32_t bi = b[I];
        if constexpr (I & 1) {
            // O: j even, pairs (I + j - 1, I + j), all written; carry -> O[I + n - 1]
            wmad<false>(O[I - 1], O[I], a[0], bi);
#pragma unroll
            for (int j = 2; j < n; j += 2) wmad<true>(O[I + j - 1], O[I + j], a[j], bi);
            addc(O[I + n - 1], 0, 0);
            // E: j odd, pairs (I + j, I + j + 1); top pair (I + n - 1, I + n): low word written for I >= 3
            if constexpr (n == 2) {
                wmad_top<false>(E[I + 1], E[I + 2], a[1], bi, I == 1 ? 0u : E[I + 1]);
            } else {
                wmad<false>(E[I + 1], E[I + 2], a[1], bi);
#pragma unroll
                for (int j = 3; j < n - 1; j += 2) wmad<true>(E[I + j], E[I + j + 1], a[j], bi);
                wmad_top<true>(E[I + n - 1], E[I + n], a[n - 1], bi, I == 1 ? 0u : E[I + n - 1]);
            }
        } else {
            // E: j even, pairs (I + j, I + j + 1), all written; carry -> E[I + n]
            wmad<false>(E[I], E[I + 1], a[0], bi);
#pragma unroll
            for (int j = 2; j < n; j += 2) wmad<true>(E[I + j], E[I + j + 1], a[j], bi);
            addc(E[I + n], 0, 0);
            // O: j odd, pairs (I + j - 1, I + j); top pair (I + n - 2, I + n - 1): low word = last row's carry limb
            if constexpr (n == 2) {
                wmad_top<false>(O[I], O[I + 1], a[1], bi, O[I]);
            } else {
                wmad<false>(O[I], O[I + 1], a[1], bi);
#pragma unroll
                for (int j = 3; j < n - 1; j += 2) wmad<true>(O[I + j - 1], O[I + j], a[j], bi);
                wmad_top<true>(O[I + n - 2], O[I + n - 1], a[n - 1], bi, O[I + n - 2]);
            }
        }
        school_rows<n, I + 1>(E, O, a, b);
    }
}
template <int n>
KDEV void mul_school(uint32_t (&t)[2 * n], const uint32_t (&a)[n], const uint32_t (&b)[n]) {
    static_assert(n >= 2 && n % 2 == 0, "even limb counts only");
    uint32_t E[2 * n], O[2 * n];
#pragma unroll
    for (int j = 0; j < n; j += 2) {
        wmul(E[j], E[j + 1], a[j], b[0]);
        wmul(O[j], O[j + 1], a[j + 1], b[0]);
    }
    school_rows<n, 1>(E, O, a, b);
    t[0] = E[0];
    add_cc(t[1], E[1], O[0]);
#pragma unroll
    for (int k = 2; k < 2 * n - 1; k++) addc_cc(t[k], E[k], O[k - 1]);
    addc(t[2 * n - 1], E[2 * n - 1], O[2 * n - 2]);
}

// d = |x - y|, returns the all-ones mask when x < y
template <int n>
KDEV uint32_t abs_diff(uint32_t (&d)[n], const uint32_t (&x)[n], const uint32_t (&y)[n]) {
    uint32_t neg;
    sub_cc(d[0], x[0], y[0]);
#pragma unroll
    for (int i = 1; i < n; i++) subc_cc(d[i], x[i], y[i]);
    subc(neg, 0, 0);
    // two's complement negation under the mask: (d ^ neg) + (neg & 1)
    add_cc(d[0], d[0] ^ neg, neg & 1u);
#pragma unroll
    for (int i = 1; i < n - 1; i++) addc_cc(d[i], d[i] ^ neg, 0);
    addc(d[n - 1], d[n - 1] ^ neg, 0);
    return neg;
}

// ---- t = a * b: Karatsuba down to schoolbook products of BASE limbs --------------------------------------------
// a = a0 + a1 W, b = b0 + b1 W (W = 2^(32 h)): t = z0 + (z0 + z2 - (a0 - a1)(b0 - b1)) W + z2 W^2.  The subtractive
// middle term keeps every operand at h limbs (no carry limb to multiply).
template <int n, int BASE>
struct Prod {
    KDEV static void run(uint32_t (&t)[2 * n], const uint32_t (&a)[n], const uint32_t (&b)[n]) {
        if constexpr (n <= BASE || (n / 2) % 2 != 0) {
            mul_school<n>(t, a, b);
        } else {
            constexpr int h = n / 2;
            uint32_t a0[h], a1[h], b0[h], b1[h];
#pragma unroll
            for (int i = 0; i < h; i++) {
                a0[i] = a[i];
                a1[i] = a[h + i];
                b0[i] = b[i];
                b1[i] = b[h + i];
            }
            uint32_t z0[2 * h], z2[2 * h], mm[2 * h], da[h], db[h];
            Prod<h, BASE>::run(z0, a0, b0);
            Prod<h, BASE>::run(z2, a1, b1);
            const uint32_t sa = abs_diff<h>(da, a0, a1), sb = abs_diff<h>(db, b0, b1);
            Prod<h, BASE>::run(mm, da, db);
            // s = z0 + z2 (carry cs); mid = s -/+ mm: subtract when the two differences have the same sign
            uint32_t s[2 * h], cs;
            add_cc(s[0], z0[0], z2[0]);
#pragma unroll
            for (int k = 1; k < 2 * h; k++) addc_cc(s[k], z0[k], z2[k]);
            addc(cs, 0, 0);
            const uint32_t neg = ~(sa ^ sb);
            uint32_t mid[2 * h], cm, dummy;
            add_cc(dummy, neg, neg);                   // CC = (neg != 0): the + 1 of the two's complement
#pragma unroll
            for (int k = 0; k < 2 * h; k++) addc_cc(mid[k], s[k], mm[k] ^ neg);
            addc(cm, cs, 0);
            cm -= neg & 1u;                            // s - mm = s + ~mm + 1 - W^2: the carry out pays the W^2
#pragma unroll
            for (int k = 0; k < h; k++) t[k] = z0[k];
            add_cc(t[h], z0[h], mid[0]);
#pragma unroll
            for (int k = 1; k < h; k++) addc_cc(t[h + k], z0[h + k], mid[k]);
#pragma unroll
            for (int k = 0; k < h; k++) addc_cc(t[2 * h + k], z2[k], mid[h + k]);
            addc_cc(t[3 * h], z2[h], cm);
#pragma unroll
            for (int k = 1; k < h - 1; k++) addc_cc(t[3 * h + k], z2[h + k], 0);
            addc(t[4 * h - 1], z2[2 * h - 1], 0);
        }
    }
};

// ---- Montgomery reduction of a double-width value: r = t / 2^(32 N) mod p, r in [0, 2p) for t < 2 p^2 ----------
// Row i clears limb i: m = v_0 * (-p^-1), v += m p, v >>= 32; limb N + i - 1 of t enters at the top one row late (as
// the low word of the fresh pair, whose high word then cannot overflow), limb 2N - 1 in the final sum.  The even / odd
// accumulators swap roles every row exactly as in Fp::mont_row: the division by 2^32 is free.
template <class P, int I>
KDEV void redc_rows(uint32_t (&even)[P::N], uint32_t (&odd)[P::N], const uint32_t (&t)[2 * P::N], uint32_t inv) {
    constexpr int N = P::N;
    if constexpr (I < N) {
        add_cc(even[0], even[0], odd[1]);
        const uint32_t m = mul_lo(even[0], inv);
#pragma unroll
        for (int j = 0; j < N - 2; j += 2) wmad4<true>(odd[j], odd[j + 1], P::mod(j + 1), m, odd[j + 2], odd[j + 3]);
        wmad_top<true>(odd[N - 2], odd[N - 1], P::mod(N - 1), m, t[N + I - 1]);
        wmad<false>(even[0], even[1], P::mod(0), m);
#pragma unroll
        for (int j = 2; j < N; j += 2) wmad<true>(even[j], even[j + 1], P::mod(j), m);
        addc(odd[N - 1], odd[N - 1], 0);
        redc_rows<P, I + 1>(odd, even, t, inv);
    }
}
template <class P>
KDEV void redc(uint32_t (&r)[P::N], const uint32_t (&t)[2 * P::N], uint32_t inv) {
    constexpr int N = P::N;
    static_assert(N % 2 == 0, "even limb counts only");
    uint32_t A[N], B[N];
#pragma unroll
    for (int i = 0; i < N; i++) A[i] = t[i];
    const uint32_t m = mul_lo(A[0], inv);
#pragma unroll
    for (int j = 0; j < N; j += 2) wmul(B[j], B[j + 1], P::mod(j + 1), m);
    wmad<false>(A[0], A[1], P::mod(0), m);
#pragma unroll
    for (int j = 2; j < N; j += 2) wmad<true>(A[j], A[j + 1], P::mod(j), m);
    addc(B[N - 1], B[N - 1], 0);
    redc_rows<P, 1>(B, A, t, inv);
    // N rows done (N even): the last row ran with even = B, odd = A; B[0] == 0 and the value is A + (B >> 32)
    add_cc(r[0], A[0], B[1]);
#pragma unroll
    for (int k = 1; k < N - 1; k++) addc_cc(r[k], A[k], B[k + 1]);
    addc(r[N - 1], A[N - 1], t[2 * N - 1]);
}

}  // namespace kara
}  // namespace b200

#ifndef B200_KARA_HOST_EMU
#include "fp.cuh"
namespace b200 {
// r = a b / R and r = (a b + c d) / R, canonical.  The second form is the lazy reduction of a sum of two products
// (2 p^2 + p R < R^2: one conditional subtraction still lands in [0, p)).
template <int BASE, class P>
B200_DEV Fp<P> kara_mul_inline(const Fp<P> &a, const Fp<P> &b) {
    uint32_t t[2 * P::N];
    kara::Prod<P::N, BASE>::run(t, a.l, b.l);
    Fp<P> r;
    kara::redc<P>(r.l, t, c_mont_inv[P::INV_SLOT]);
    r.reduce_once();
    return r;
}
template <int BASE, class P>
B200_DEV Fp<P> kara_dot2_inline(const Fp<P> &a, const Fp<P> &b, const Fp<P> &c, const Fp<P> &d) {
    constexpr int N = P::N;
    uint32_t t[2 * N], u[2 * N];
    kara::Prod<N, BASE>::run(t, a.l, b.l);
    kara::Prod<N, BASE>::run(u, c.l, d.l);
    add_cc(t[0], t[0], u[0]);
#pragma unroll
    for (int i = 1; i < 2 * N - 1; i++) addc_cc(t[i], t[i], u[i]);
    addc(t[2 * N - 1], t[2 * N - 1], u[2 * N - 1]);
    Fp<P> r;
    kara::redc<P>(r.l, t, c_mont_inv[P::INV_SLOT]);
    r.reduce_once();
    return r;
}
template <int BASE, class P>
__device__ __noinline__ Fp<P> kara_mul_outline(Fp<P> a, Fp<P> b) { return kara_mul_inline<BASE>(a, b); }
template <int BASE, class P>
__device__ __noinline__ Fp<P> kara_dot2_outline(Fp<P> a, Fp<P> b, Fp<P> c, Fp<P> d) { return kara_dot2_inline<BASE>(a, b, c, d); }
}  // namespace b200
#endif
