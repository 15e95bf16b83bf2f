// Prime-field arithmetic for sm_100a.
//
// Device replacement for ark-ff 0.1.0 Fp384 / Fp768 as reached from
// crates/bls-crypto/src/bls/signature.rs:85 and public.rs:61.  Elements are arkworks' own
// representation both in memory and in registers: Montgomery residues (R = 2^384 / 2^768),
// canonical in [0, p), 32-bit limbs (12 for BLS12-377 Fq, 24 for BW6-761 Fq), so device
// buffers are bit-compatible with what the Rust side holds and nothing is converted at the
// C-ABI boundary.
//
// Multiplication is a row-interleaved (CIOS-style) Montgomery product with the partial
// products split into an "even" and an "odd" accumulator (each N limbs, the odd one offset by
// one limb).  A wide product a[j]*b lands on limbs (j, j+1), so products of even j chain
// through one accumulator without overlapping and products of odd j through the other.
// Dividing by 2^32 after each row is free: the two accumulators swap roles (mont_row()).
// The chains are PTX mad.lo.cc / madc.hi.cc pairs that ptxas fuses into IMAD.WIDE.U32.X.
//
// Measured on B200 (profiles/r1_microbench_*.txt, profiles/r1_field_layer_notes.md):
//   * a 32x32->64 multiply-add (IMAD.WIDE.U32, with or without carry) occupies the FMA-heavy
//     pipe for 4 cycles per warp; IMAD.HI also 4, a 32-bit IMAD 2.  2*N*N wide products per
//     Montgomery product is therefore the floor; an unsaturated 28-bit-limb variant (no carry
//     chains, 14 limbs) was built and measured: it saturates the pipe (93 %) but needs 36 %
//     more products and lost.
//   * -p^-1 mod 2^32 is read from constant memory at run time on purpose: when ptxas can see
//     that it is 0xffffffff (BLS12-377: p = 1 mod 2^32) it rewrites the reduction products as
//     IMAD.X + IMAD.HI.U32.X pairs (6 pipe cycles instead of 4); with an opaque value all
//     2*N*N products stay fused.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "params_gen.cuh"

namespace b200 {

#define B200_DEV __device__ __forceinline__

// -p^-1 mod 2^32 per field (slot 0: BLS12-377 Fq, slot 1: BW6-761 Fq, slot 2: BLS12-377 Fr); see the note above.
static __constant__ uint32_t c_mont_inv[3] = {Fq377Params::INV, Fq761Params::INV, Fr253Params::INV};

// ---- carry-chain primitives (all volatile: the CC flag links consecutive asm statements) ----
B200_DEV void add_cc(uint32_t &r, uint32_t a, uint32_t b) { asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
B200_DEV void addc_cc(uint32_t &r, uint32_t a, uint32_t b) { asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
B200_DEV void addc(uint32_t &r, uint32_t a, uint32_t b) { asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
B200_DEV void sub_cc(uint32_t &r, uint32_t a, uint32_t b) { asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
B200_DEV void subc_cc(uint32_t &r, uint32_t a, uint32_t b) { asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
B200_DEV void subc(uint32_t &r, uint32_t a, uint32_t b) { asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }

// (lo, hi) = a * b
B200_DEV void mul_wide(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b) {
    asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
// (lo, hi) += a * b, starting a chain (carry out in CC)
B200_DEV void mad_wide_cc(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (lo, hi) += a * b + CC, continuing a chain
B200_DEV void madc_wide_cc(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (lo, hi) = a * b + (c_lo, c_hi) + CC, continuing a chain
B200_DEV void madc_wide_cc(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b, uint32_t c_lo, uint32_t c_hi) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
                 : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(c_lo), "r"(c_hi));
}
// top of the shifted chain: (lo, hi) = a * b + CC ; hi cannot overflow
B200_DEV void madc_wide_top(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, 0; madc.hi.u32 %1, %2, %3, 0;" : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b));
}

// Compiler fence for large aggregates.  cicc (CUDA 12.9) was caught forwarding a stale source
// into a by-value argument: after `Fq12 r = f;` every `r = f12_sqr(r)` of a loop squared f, not
// r (the PTX re-stored f's registers into the argument slot at the loop head).  Escaping the
// object's address into an asm with a memory clobber makes the copy opaque and stops the
// forwarding.  Used after every copy-initialisation / loop-carried assignment of a large aggregate (Fq12, and the cold MSM helper kernels).
template <class T>
B200_DEV void launder(T &x) {
    asm volatile("" : : "l"(&x) : "memory");
}

template <int NW>
struct alignas(16) FpMem {                         // memory image: NW little-endian 32-bit words
    uint32_t w[NW];
};

template <class P>
struct Fp {
    static constexpr int N = P::N;
    static constexpr int NW = P::N;
    using Params = P;
    using Mem = FpMem<NW>;
    uint32_t l[N];

    B200_DEV static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = 0;
        return r;
    }
    B200_DEV static Fp one() {                     // Montgomery form of 1
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::one(i);
        return r;
    }

    // ---- memory image <-> registers.  The register form IS the arkworks form, so the "ark"
    // and the engine-internal ("load/store") flavours coincide; both names are kept so the
    // kernels say which side of the C-ABI a buffer is on.
    B200_DEV static Fp load(const Mem &m) {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = m.w[i];
        return r;
    }
    B200_DEV Mem store() const {
        Mem m;
#pragma unroll
        for (int i = 0; i < N; i++) m.w[i] = l[i];
        return m;
    }
    B200_DEV static Fp from_ark(const Mem &m) { return load(m); }
    B200_DEV Mem to_ark() const { return store(); }

    // lane exchange / selection helpers for the quad-cooperative point operations (ec.cuh)
    B200_DEV Fp shfl(unsigned mask, int src_lane) const {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = __shfl_sync(mask, l[i], src_lane);
        return r;
    }
    B200_DEV static Fp sel4(int q, const Fp &a0, const Fp &a1, const Fp &a2, const Fp &a3) {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) {
            uint32_t lo = (q & 1) ? a1.l[i] : a0.l[i];
            uint32_t hi = (q & 1) ? a3.l[i] : a2.l[i];
            r.l[i] = (q & 2) ? hi : lo;
        }
        return r;
    }

    B200_DEV bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; i++) o |= l[i];
        return o == 0;
    }
    B200_DEV bool operator==(const Fp &b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; i++) o |= l[i] ^ b.l[i];
        return o == 0;
    }

    // r in [0, 2p) -> [0, p)
    B200_DEV void reduce_once() {
        uint32_t t[N], borrow;
        sub_cc(t[0], l[0], P::mod(0));
#pragma unroll
        for (int i = 1; i < N; i++) subc_cc(t[i], l[i], P::mod(i));
        subc(borrow, 0, 0);                        // 0xffffffff iff l < p
#pragma unroll
        for (int i = 0; i < N; i++) l[i] = borrow ? l[i] : t[i];
    }

    B200_DEV friend Fp operator+(const Fp &a, const Fp &b) {
        Fp r;                                      // a + b < 2p < 2^(32N): no carry out
        add_cc(r.l[0], a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) addc_cc(r.l[i], a.l[i], b.l[i]);
        addc(r.l[N - 1], a.l[N - 1], b.l[N - 1]);
        r.reduce_once();
        return r;
    }
    B200_DEV friend Fp operator-(const Fp &a, const Fp &b) {
        Fp r;
        uint32_t borrow;
        sub_cc(r.l[0], a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N; i++) subc_cc(r.l[i], a.l[i], b.l[i]);
        subc(borrow, 0, 0);                        // all-ones iff a < b
        add_cc(r.l[0], r.l[0], P::mod(0) & borrow);
#pragma unroll
        for (int i = 1; i < N - 1; i++) addc_cc(r.l[i], r.l[i], P::mod(i) & borrow);
        addc(r.l[N - 1], r.l[N - 1], P::mod(N - 1) & borrow);
        return r;
    }
    B200_DEV Fp neg() const { return zero() - *this; }      // 0 - 0 stays 0 (no borrow, no +p)
    B200_DEV Fp cneg(bool flag) const {            // flag ? -x : x
        Fp n = neg(), r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = flag ? n.l[i] : l[i];
        return r;
    }
    B200_DEV Fp dbl() const { return *this + *this; }

    // one row of the interleaved product: (even, odd) += a * bi + m * p, then / 2^32 by role swap.
    // On entry (non-first rows) `even` is the previous row's odd accumulator and `odd` the
    // previous even one, whose limb 0 is zero: odd >> 64 re-aligns it one limb above `even`,
    // and its limb 1 is folded into even[0] with the carry entering the odd chain.
    template <bool FIRST>
    B200_DEV static void mont_row(uint32_t (&even)[N], uint32_t (&odd)[N], const uint32_t (&a)[N], uint32_t bi,
                                  uint32_t inv) {
        if (FIRST) {
#pragma unroll
            for (int j = 0; j < N; j += 2) {
                mul_wide(odd[j], odd[j + 1], a[j + 1], bi);
                mul_wide(even[j], even[j + 1], a[j], bi);
            }
        } else {
            add_cc(even[0], even[0], odd[1]);
#pragma unroll
            for (int j = 0; j < N - 2; j += 2) madc_wide_cc(odd[j], odd[j + 1], a[j + 1], bi, odd[j + 2], odd[j + 3]);
            madc_wide_top(odd[N - 2], odd[N - 1], a[N - 1], bi);
            mad_wide_cc(even[0], even[1], a[0], bi);
#pragma unroll
            for (int j = 2; j < N; j += 2) madc_wide_cc(even[j], even[j + 1], a[j], bi);
            addc(odd[N - 1], odd[N - 1], 0);
        }
        uint32_t m = even[0] * inv;
        mad_wide_cc(odd[0], odd[1], P::mod(1), m);
#pragma unroll
        for (int j = 2; j < N; j += 2) madc_wide_cc(odd[j], odd[j + 1], P::mod(j + 1), m);
        mad_wide_cc(even[0], even[1], P::mod(0), m);
#pragma unroll
        for (int j = 2; j < N; j += 2) madc_wide_cc(even[j], even[j + 1], P::mod(j), m);
        addc(odd[N - 1], odd[N - 1], 0);
    }

    B200_DEV static Fp mul_inline(const Fp &a, const Fp &b) {
        const uint32_t inv = c_mont_inv[P::INV_SLOT];
        uint32_t even[N], odd[N];
        mont_row<true>(even, odd, a.l, b.l[0], inv);
        mont_row<false>(odd, even, a.l, b.l[1], inv);
#pragma unroll
        for (int i = 2; i < N; i += 2) {
            mont_row<false>(even, odd, a.l, b.l[i], inv);
            mont_row<false>(odd, even, a.l, b.l[i + 1], inv);
        }
        // the last row left odd[0] == 0: value = even + (odd >> 32), i.e. r[k] = even[k] + odd[k + 1]
        Fp r;
        add_cc(r.l[0], even[0], odd[1]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) addc_cc(r.l[i], even[i], odd[i + 1]);
        addc(r.l[N - 1], even[N - 1], 0);
        r.reduce_once();
        return r;
    }
    // by value on purpose: pointer-taking out-of-line helpers were observed to be miscompiled
    // (the caller passed one stack slot for all three pointers); value semantics cannot alias.
    __device__ __noinline__ static Fp mul_outline(Fp a, Fp b) { return mul_inline(a, b); }
    // N = 12 (BLS12-377): fully inlined, 288 IMAD.WIDE.  N = 24 (BW6-761) is 1152 per product:
    // one shared out-of-line body keeps code size (and the instruction cache) sane.
    B200_DEV friend Fp operator*(const Fp &a, const Fp &b) {
        if constexpr (N <= 12) {
            return mul_inline(a, b);
        } else {
            return mul_outline(a, b);
        }
    }
    B200_DEV Fp sqr() const;                       // dedicated squaring, defined below

    // ---- dedicated squaring -----------------------------------------------------------------------
    // a^2 = sum_i a_i 2^(32 i) * w_i with w_i = a_i + 2 * (a >> 32 (i + 1)) << 32: row i of the
    // interleaved product multiplies a_i by the limbs j >= i of w_i only (the products with j < i are
    // the ones already counted by doubling), so N (N - 1) / 2 of the N^2 operand products disappear
    // (66 of 144 for 12 limbs; the reduction products stay): 222 wide multiply-adds instead of 288.
    // Where the general row shifts the odd accumulator inside a product chain, skipped products become
    // plain add-with-carry moves (ALU pipe, not the multiplier).  Requires a < 2^(32 N - 1): true for
    // all three moduli (377 / 761 / 253 bits), so doubling never carries out of the top limb.
    template <int ROW>
    B200_DEV static uint32_t sq_limb(const uint32_t (&a)[N], int j) {             // limb j >= ROW of w_ROW
        return j == ROW ? a[j] : (j == ROW + 1 ? a[j] << 1 : __funnelshift_l(a[j - 1], a[j], 1));
    }
    template <int ROW>
    B200_DEV static void mont_row_sq(uint32_t (&even)[N], uint32_t (&odd)[N], const uint32_t (&a)[N], uint32_t inv) {
        const uint32_t bi = a[ROW];
        if (ROW == 0) {
#pragma unroll
            for (int j = 0; j < N; j += 2) {
                mul_wide(odd[j], odd[j + 1], sq_limb<ROW>(a, j + 1), bi);
                mul_wide(even[j], even[j + 1], sq_limb<ROW>(a, j), bi);
            }
        } else {
            add_cc(even[0], even[0], odd[1]);
#pragma unroll
            for (int j = 0; j < N - 2; j += 2) {
                if (j + 1 >= ROW) {
                    madc_wide_cc(odd[j], odd[j + 1], sq_limb<ROW>(a, j + 1), bi, odd[j + 2], odd[j + 3]);
                } else {                                                          // skipped product: shift with carry
                    addc_cc(odd[j], odd[j + 2], 0);
                    addc_cc(odd[j + 1], odd[j + 3], 0);
                }
            }
            madc_wide_top(odd[N - 2], odd[N - 1], sq_limb<ROW>(a, N - 1), bi);     // N - 1 >= ROW always
            // even chain: starts at the first limb index >= ROW of its parity, nothing to add below it
            constexpr int J0 = (ROW + 1) & ~1;                                    // first even j >= ROW
            if (J0 < N) {
                mad_wide_cc(even[J0], even[J0 + 1], sq_limb<ROW>(a, J0), bi);
#pragma unroll
                for (int j = J0 + 2; j < N; j += 2) madc_wide_cc(even[j], even[j + 1], sq_limb<ROW>(a, j), bi);
                addc(odd[N - 1], odd[N - 1], 0);
            }
        }
        uint32_t m = even[0] * inv;
        mad_wide_cc(odd[0], odd[1], P::mod(1), m);
#pragma unroll
        for (int j = 2; j < N; j += 2) madc_wide_cc(odd[j], odd[j + 1], P::mod(j + 1), m);
        mad_wide_cc(even[0], even[1], P::mod(0), m);
#pragma unroll
        for (int j = 2; j < N; j += 2) madc_wide_cc(even[j], even[j + 1], P::mod(j), m);
        addc(odd[N - 1], odd[N - 1], 0);
    }
    template <int ROW>
    B200_DEV static void sq_rows(uint32_t (&even)[N], uint32_t (&odd)[N], const uint32_t (&a)[N], uint32_t inv) {
        if constexpr (ROW < N) {
            mont_row_sq<ROW>(even, odd, a, inv);
            mont_row_sq<ROW + 1>(odd, even, a, inv);
            sq_rows<ROW + 2>(even, odd, a, inv);
        }
    }
    B200_DEV static Fp sqr_inline(const Fp &a) {
        const uint32_t inv = c_mont_inv[P::INV_SLOT];
        uint32_t even[N], odd[N];
        sq_rows<0>(even, odd, a.l, inv);
        Fp r;
        add_cc(r.l[0], even[0], odd[1]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) addc_cc(r.l[i], even[i], odd[i + 1]);
        addc(r.l[N - 1], even[N - 1], 0);
        r.reduce_once();
        return r;
    }
    __device__ __noinline__ static Fp sqr_outline(Fp a) { return sqr_inline(a); }

    // ---- inversion; inv(0) = 0.  Cold path: out of line. -------------------------------------------
    // Bernstein-Yang "safegcd" division steps on signed 30-bit limbs (the modinv32 scheme): 30 division
    // steps at a time are run on the low words of (f, g) only and collected in a 2x2 integer matrix,
    // which is then applied to the full-width f, g and -- modulo p, with an exact division by 2^30 --
    // to d, e (invariants d * a = f, e * a = g mod p).  Ends when g = 0 (f = +-1, d = +-a^-1):
    // <= 26 / 52 outer iterations of ~800 instructions for the 377 / 761-bit moduli, against ~750 / 1500
    // iterations of the bit-at-a-time binary Euclid it replaced (5x fewer instructions; the batched-affine
    // bucket accumulation waits on exactly this latency) and a 570-product Fermat chain before that.
    // The input is the Montgomery residue aR, so the loop yields (aR)^-1; one Montgomery product with
    // R^3 turns that into a^-1 R.
    static constexpr int L30 = (32 * N + 29) / 30;
    static constexpr uint32_t M30 = (1u << 30) - 1u;
    __host__ __device__ static constexpr uint32_t mod30(int i) {                 // bits [30 i, 30 i + 30) of p
        const int pos = 30 * i, w = pos >> 5, sh = pos & 31;
        const uint64_t lo = w < N ? P::mod(w) : 0u, hi = w + 1 < N ? P::mod(w + 1) : 0u;
        return (uint32_t)(((hi << 32) | lo) >> sh) & M30;
    }
    B200_DEV Fp inv() const { return inv_outline(*this); }
    __device__ __noinline__ static Fp inv_outline(Fp a) {
        if (a.is_zero()) return a;
        int32_t f[L30], g[L30], d[L30], e[L30];
#pragma unroll
        for (int i = 0; i < L30; i++) {
            const int pos = 30 * i, w = pos >> 5, sh = pos & 31;
            const uint64_t lo = a.l[w < N ? w : N - 1] * (uint64_t)(w < N), hi = w + 1 < N ? a.l[w + 1] : 0u;
            g[i] = (int32_t)((uint32_t)(((hi << 32) | lo) >> sh) & M30);
            f[i] = (int32_t)mod30(i);
            d[i] = 0;
            e[i] = i == 0;
        }
        const uint32_t inv30 = (0u - c_mont_inv[P::INV_SLOT]) & M30;             // p^-1 mod 2^30
        int32_t zeta = -1;
#pragma unroll 1
        for (int iter = 0; iter < 4 * L30 + 8; iter++) {
            // 30 division steps on the low words -> transition matrix (u v; q r), scaled by 2^30
            int32_t u = 1, v = 0, q = 0, r = 1;
            {
                uint32_t fl = (uint32_t)f[0] | ((uint32_t)f[1] << 30), gl = (uint32_t)g[0] | ((uint32_t)g[1] << 30);
#pragma unroll 1
                for (int k = 0; k < 30; k++) {
                    int32_t c1 = zeta >> 31, c2 = -(int32_t)(gl & 1u);
                    uint32_t x = (fl ^ (uint32_t)c1) - (uint32_t)c1;
                    int32_t y = (u ^ c1) - c1, z = (v ^ c1) - c1;
                    gl += x & (uint32_t)c2;
                    q += y & c2;
                    r += z & c2;
                    c1 &= c2;
                    zeta = (zeta ^ c1) - 1;
                    fl += gl & (uint32_t)c1;
                    u += q & c1;
                    v += r & c1;
                    gl >>= 1;
                    u <<= 1;
                    v <<= 1;
                }
            }
            {   // (d, e) <- (u d + v e, q d + r e) / 2^30 mod p
                const int32_t sd = d[L30 - 1] >> 31, se = e[L30 - 1] >> 31;
                int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);
                long long cd = (long long)u * d[0] + (long long)v * e[0], ce = (long long)q * d[0] + (long long)r * e[0];
                md -= (int32_t)((inv30 * (uint32_t)cd + (uint32_t)md) & M30);
                me -= (int32_t)((inv30 * (uint32_t)ce + (uint32_t)me) & M30);
                cd += (long long)(int32_t)mod30(0) * md;
                ce += (long long)(int32_t)mod30(0) * me;
                cd >>= 30;
                ce >>= 30;
#pragma unroll
                for (int i = 1; i < L30; i++) {
                    cd += (long long)u * d[i] + (long long)v * e[i] + (long long)(int32_t)mod30(i) * md;
                    ce += (long long)q * d[i] + (long long)r * e[i] + (long long)(int32_t)mod30(i) * me;
                    d[i - 1] = (int32_t)((uint32_t)cd & M30);
                    e[i - 1] = (int32_t)((uint32_t)ce & M30);
                    cd >>= 30;
                    ce >>= 30;
                }
                d[L30 - 1] = (int32_t)cd;
                e[L30 - 1] = (int32_t)ce;
            }
            uint32_t g_any = 0;
            {   // (f, g) <- (u f + v g, q f + r g) / 2^30 (exact)
                long long cf = (long long)u * f[0] + (long long)v * g[0], cg = (long long)q * f[0] + (long long)r * g[0];
                cf >>= 30;
                cg >>= 30;
#pragma unroll
                for (int i = 1; i < L30; i++) {
                    cf += (long long)u * f[i] + (long long)v * g[i];
                    cg += (long long)q * f[i] + (long long)r * g[i];
                    f[i - 1] = (int32_t)((uint32_t)cf & M30);
                    g[i - 1] = (int32_t)((uint32_t)cg & M30);
                    g_any |= (uint32_t)g[i - 1];
                    cf >>= 30;
                    cg >>= 30;
                }
                f[L30 - 1] = (int32_t)cf;
                g[L30 - 1] = (int32_t)cg;
                g_any |= (uint32_t)g[L30 - 1];
            }
            if (g_any == 0) break;
        }
        // d * sign(f) in (-2p, 2p) -> [0, 2p): conditional negation, carry normalisation, += p while negative
        const int32_t fs = f[L30 - 1] >> 31;
        int32_t carry = 0;
#pragma unroll
        for (int i = 0; i < L30; i++) {
            int32_t t = ((d[i] ^ fs) - fs) + carry;
            carry = i < L30 - 1 ? t >> 30 : 0;
            d[i] = i < L30 - 1 ? (int32_t)((uint32_t)t & M30) : t;
        }
#pragma unroll 1
        for (int rep = 0; rep < 2; rep++) {
            const int32_t neg = d[L30 - 1] >> 31;
            carry = 0;
#pragma unroll
            for (int i = 0; i < L30; i++) {
                int32_t t = d[i] + (int32_t)(mod30(i) & (uint32_t)neg) + carry;
                carry = i < L30 - 1 ? t >> 30 : 0;
                d[i] = i < L30 - 1 ? (int32_t)((uint32_t)t & M30) : t;
            }
        }
        Fp x = zero();
#pragma unroll
        for (int i = 0; i < L30; i++) {
            const int pos = 30 * i, w = pos >> 5, sh = pos & 31;
            const uint64_t val = (uint64_t)(uint32_t)d[i] << sh;
            if (w < N) x.l[w] |= (uint32_t)val;
            if (w + 1 < N) x.l[w + 1] |= (uint32_t)(val >> 32);
        }
        x.reduce_once();
        Fp r3;
#pragma unroll
        for (int i = 0; i < N; i++) r3.l[i] = P::r3(i);
        return mul_outline(x, r3);                    // shared body: keeps this cold function small
    }

    // a^(p-2) (Fermat) -- the round-1 inversion, kept to cross-check inv() on the device
    __device__ __noinline__ static uint32_t pm2_word(int w) {        // runtime-indexed word of p - 2
        uint32_t r = 0;
#pragma unroll
        for (int i = 0; i < N; i++) r = (i == w) ? P::pm2(i) : r;
        return r;
    }
    B200_DEV Fp inv_fermat() const { return inv_fermat_outline(*this); }
    __device__ __noinline__ static Fp inv_fermat_outline(Fp base) {
        Fp acc = one();
#pragma unroll 1
        for (int w = 0; w < N; w++) {
            uint32_t e = pm2_word(w);
#pragma unroll 1
            for (int b = 0; b < 32; b++) {
                if ((e >> b) & 1u) acc = acc * base;
                base = base.sqr();
            }
        }
        return acc;
    }
};

// same inlining policy as operator*: 12 limbs and fewer inline, the 24-limb body shared
template <class P>
B200_DEV Fp<P> Fp<P>::sqr() const {
    if constexpr (N <= 12) {
        return sqr_inline(*this);
    } else {
        return sqr_outline(*this);
    }
}

using Fq377 = Fp<Fq377Params>;
using Fq761 = Fp<Fq761Params>;
using Fr253 = Fp<Fr253Params>;                      // BLS12-377 scalar field (NTT domain of the inner Groth16 proof)

}  // namespace b200
