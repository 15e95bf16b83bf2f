// Warp-cooperative field and point arithmetic for the latency-bound tail of the MSM (sm_100a).
//
// The Horner combine over the windows of VariableBaseMSM (crates/bls-crypto/src/bls/signature.rs:85,
// public.rs:61, crates/epoch-snark/src/api/prover.rs:78) is ONE chain of (windows - 1) * c dependent point
// doublings: 240 for BLS12-377 at n = 2^20, 368 for BW6-761.  A thread-per-element Montgomery product issues
// 2 N^2 wide multiply-adds from one warp however many lanes are alive (1152 / 4608 issue cycles for 12 / 24
// limbs), so a chain is paced by that figure.  Here ONE product is spread over the lanes of a warp instead:
//
//   * limb i of every operand lives in lane i of a group (16 lanes for the 12-limb field -- two groups per
//     warp, which is how an Fq2 element is held: c0 in the lower half, c1 in the upper half -- 32 lanes for
//     the 24-limb field);
//   * a product is three column-parallel passes (T = a b; m = T_low q mod R with q = -p^-1 mod R; (T + m p) / R):
//     lane i owns columns i and N + i, so every lane does exactly N partial products per pass; operands reach
//     it by shuffle (a broadcast of limb j and a rotation of the other operand), nothing waits on a quotient
//     digit the way row-wise (CIOS) reduction does;
//   * column sums are 96-bit; limbs are recovered with two neighbour shuffles and ONE generate / propagate
//     carry look-ahead on __ballot_sync masks ((g | p) + g) ^ (g | p) ^ g -- additions, subtractions and the
//     final conditional subtraction use the same look-ahead.
//
// tools/coop_model.py is the lane-level model of exactly these routines (tests/test_coop_model.py checks it
// against integers on the CPU); this file follows it statement by statement.
//
// A point operation then runs on a block of four warps: the (at most four) independent products of a round
// go to one warp each and meet in shared memory.  XYZZ doubling = 3 rounds, addition = 4 rounds.
#pragma once
#include "ec.cuh"

namespace b200 {

constexpr unsigned COOP_FULL = 0xffffffffu;

// acc (96 bit) += a * b
B200_DEV void mad96(uint32_t &l0, uint32_t &l1, uint32_t &l2, uint32_t a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
                 : "+r"(l0), "+r"(l1), "+r"(l2)
                 : "r"(a), "r"(b));
}

// Per-lane constants of one base field: limb g of p, of q = -p^-1 mod R and of the Montgomery form of 1.
template <class P, int GW>
struct CoopCtx {
    static constexpr int N = P::N;
    uint32_t p, q, one;
    int g;                                           // lane within the group
    int gbase;                                       // first lane of the group within the warp
    bool active;                                     // g < N
    B200_DEV static CoopCtx make() {
        CoopCtx c;
        const int lane = threadIdx.x & 31;
        c.g = lane & (GW - 1);
        c.gbase = lane & ~(GW - 1);
        c.active = c.g < N;
        c.p = c.q = c.one = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            c.p = c.g == i ? P::mod(i) : c.p;
            c.q = c.g == i ? P::ninv(i) : c.q;
            c.one = c.g == i ? P::one(i) : c.one;
        }
        return c;
    }
};

template <class P, int GW>
struct CoopOps {
    static constexpr int N = P::N;
    using Ctx = CoopCtx<P, GW>;
    static_assert(GW == 16 || GW == 32, "groups are half warps or warps");
    static_assert(N <= GW && 2 * N <= 64, "one limb per lane; the look-ahead runs on 64-bit masks");

    B200_DEV static uint32_t gballot(const Ctx &c, bool pred) {
        unsigned b = __ballot_sync(COOP_FULL, pred && c.active);
        return GW == 32 ? b : (b >> c.gbase) & 0xffffu;
    }
    // carry-in mask of a ripple: cout[i] = gen[i] | prop[i] & cin[i]
    B200_DEV static uint64_t lookahead(uint64_t gen, uint64_t prop) {
        const uint64_t x = gen | prop;
        return (x + gen) ^ x ^ gen;
    }

    // 96-bit column sums (lane g: columns g and N + g) -> limbs of the low and the high half
    B200_DEV static void normalize(const Ctx &c, uint32_t l0, uint32_t l1, uint32_t l2, uint32_t h0, uint32_t h1, uint32_t h2,
                                   uint32_t &out_lo, uint32_t &out_hi) {
        const int g = c.g, s1 = g >= 1 ? g - 1 : N - 1, s2 = g >= 2 ? g - 2 : g + N - 2;
        const uint32_t l1n = __shfl_sync(COOP_FULL, l1, s1, GW), h1n = __shfl_sync(COOP_FULL, h1, s1, GW);
        const uint32_t l2n = __shfl_sync(COOP_FULL, l2, s2, GW), h2n = __shfl_sync(COOP_FULL, h2, s2, GW);
        uint64_t y = (uint64_t)l0 + (g >= 1 ? l1n : 0u) + (g >= 2 ? l2n : 0u);
        const uint32_t lo = (uint32_t)y, klo = (uint32_t)(y >> 32);
        y = (uint64_t)h0 + (g >= 1 ? h1n : l1n) + (g >= 2 ? h2n : l2n);
        const uint32_t hi = (uint32_t)y, khi = (uint32_t)(y >> 32);
        const uint32_t klon = __shfl_sync(COOP_FULL, klo, s1, GW), khin = __shfl_sync(COOP_FULL, khi, s1, GW);
        y = (uint64_t)lo + (g >= 1 ? klon : 0u);
        const uint32_t xlo = (uint32_t)y;
        const bool glo = (y >> 32) != 0;
        y = (uint64_t)hi + (g >= 1 ? khin : klon);
        const uint32_t xhi = (uint32_t)y;
        const bool ghi = (y >> 32) != 0;
        const uint64_t gen = (uint64_t)gballot(c, glo) | ((uint64_t)gballot(c, ghi) << N);
        const uint64_t prop = (uint64_t)gballot(c, xlo == 0xffffffffu) | ((uint64_t)gballot(c, xhi == 0xffffffffu) << N);
        const uint64_t cin = lookahead(gen, prop);
        out_lo = xlo + (uint32_t)((cin >> g) & 1u);
        out_hi = xhi + (uint32_t)((cin >> (N + g)) & 1u);
    }
    // low half only (mod R): the quotient m
    B200_DEV static uint32_t normalize_low(const Ctx &c, uint32_t l0, uint32_t l1, uint32_t l2) {
        const int g = c.g, s1 = g >= 1 ? g - 1 : N - 1, s2 = g >= 2 ? g - 2 : g + N - 2;
        const uint32_t l1n = __shfl_sync(COOP_FULL, l1, s1, GW), l2n = __shfl_sync(COOP_FULL, l2, s2, GW);
        uint64_t y = (uint64_t)l0 + (g >= 1 ? l1n : 0u) + (g >= 2 ? l2n : 0u);
        const uint32_t lo = (uint32_t)y, klo = (uint32_t)(y >> 32);
        const uint32_t klon = __shfl_sync(COOP_FULL, klo, s1, GW);
        y = (uint64_t)lo + (g >= 1 ? klon : 0u);
        const uint32_t xlo = (uint32_t)y;
        const uint64_t cin = lookahead(gballot(c, (y >> 32) != 0), gballot(c, xlo == 0xffffffffu));
        return xlo + (uint32_t)((cin >> g) & 1u);
    }

    // r in [0, 2p) -> [0, p)
    B200_DEV static uint32_t cond_sub(const Ctx &c, uint32_t r) {
        const uint32_t d = r - c.p;
        const uint64_t bin = lookahead(gballot(c, r < c.p), gballot(c, r == c.p));
        const bool borrow_out = (bin >> N) & 1u;          // r < p: keep r
        const uint32_t v = borrow_out ? r : d - (uint32_t)((bin >> c.g) & 1u);
        return c.active ? v : 0u;
    }

    B200_DEV static uint32_t add(const Ctx &c, uint32_t a, uint32_t b) {
        const uint32_t s = a + b;
        const uint64_t cin = lookahead(gballot(c, s < a), gballot(c, s == 0xffffffffu));
        return cond_sub(c, s + (uint32_t)((cin >> c.g) & 1u));   // a + b < 2p < R: nothing leaves the top limb
    }
    B200_DEV static uint32_t sub(const Ctx &c, uint32_t a, uint32_t b) {
        uint32_t d = a - b;
        const uint64_t bin = lookahead(gballot(c, a < b), gballot(c, a == b));
        d -= (uint32_t)((bin >> c.g) & 1u);
        const bool negative = (bin >> N) & 1u;            // a < b: add p back (the carry out of the top is dropped)
        const uint32_t s = d + c.p;
        const uint64_t cin = lookahead(gballot(c, s < d), gballot(c, s == 0xffffffffu));
        const uint32_t v = negative ? s + (uint32_t)((cin >> c.g) & 1u) : d;
        return c.active ? v : 0u;
    }
    B200_DEV static bool is_zero(const Ctx &c, uint32_t a) { return gballot(c, a != 0u) == 0u; }

    // Montgomery product a b / R mod p, in two halves so that a SUM of products can share one reduction:
    // accumulate() adds the column sums of a b into s (pass 1), reduce() turns s into (sum) / R mod p (passes 2 and 3).
    // Columns are 96 bits wide and a lane adds N products of 64 bits per accumulate(): hundreds of terms fit; the
    // reduction needs sum < R p, i.e. K p < R for K terms -- K < 128 for both moduli (7 spare bits).
    struct Acc {
        uint32_t l0 = 0, l1 = 0, l2 = 0, h0 = 0, h1 = 0, h2 = 0;
    };
    B200_DEV static void accumulate(const Ctx &c, Acc &s, uint32_t a, uint32_t b) {
        const int g = c.g;
#pragma unroll
        for (int j = 0; j < N; j++) {
            const uint32_t aj = __shfl_sync(COOP_FULL, a, j, GW);
            int k = g - j;
            k += k < 0 ? N : 0;
            const uint32_t bk = __shfl_sync(COOP_FULL, b, k, GW);
            const bool low = j <= g;
            mad96(s.l0, s.l1, s.l2, low ? aj : 0u, bk);
            mad96(s.h0, s.h1, s.h2, low ? 0u : aj, bk);
        }
    }
    B200_DEV static uint32_t reduce(const Ctx &c, const Acc &s) {
        const int g = c.g;
        uint32_t t_lo, t_hi;
        normalize(c, s.l0, s.l1, s.l2, s.h0, s.h1, s.h2, t_lo, t_hi);
        uint32_t m0 = 0, m1 = 0, m2 = 0;
#pragma unroll
        for (int j = 0; j < N; j++) {
            const uint32_t tj = __shfl_sync(COOP_FULL, t_lo, j, GW);
            int k = g - j;
            k += k < 0 ? N : 0;
            const uint32_t qk = __shfl_sync(COOP_FULL, c.q, k, GW);
            mad96(m0, m1, m2, j <= g ? tj : 0u, qk);
        }
        const uint32_t m = normalize_low(c, m0, m1, m2);
        uint32_t l0 = t_lo, l1 = 0, l2 = 0, h0 = t_hi, h1 = 0, h2 = 0;
#pragma unroll
        for (int j = 0; j < N; j++) {
            const uint32_t mj = __shfl_sync(COOP_FULL, m, j, GW);
            int k = g - j;
            k += k < 0 ? N : 0;
            const uint32_t pk = __shfl_sync(COOP_FULL, c.p, k, GW);
            const bool low = j <= g;
            mad96(l0, l1, l2, low ? mj : 0u, pk);
            mad96(h0, h1, h2, low ? 0u : mj, pk);
        }
        uint32_t u_lo, u_hi;
        normalize(c, l0, l1, l2, h0, h1, h2, u_lo, u_hi);    // u_lo == 0 by construction
        return cond_sub(c, u_hi);
    }
    B200_DEV static uint32_t mul(const Ctx &c, uint32_t a, uint32_t b) {
        Acc s;
        accumulate(c, s, a, b);
        return reduce(c, s);
    }
};

// ---- field elements spread over a warp ----------------------------------------------------------------
// Coop<F>: one element of F per warp.  Every member function is warp-collective (all 32 lanes call it with
// the element they hold); load / store take the word image of F::Mem (F::Mem is little-endian 32-bit words).
template <class F> struct Coop;

template <class P>
struct Coop<Fp<P>> {
    static constexpr int N = P::N, GW = N <= 16 ? 16 : 32, WORDS = N;
    using Ops = CoopOps<P, GW>;
    using Ctx = CoopCtx<P, GW>;
    uint32_t v;
    // 12-limb field: both half warps hold the same element (identical work, identical results)
    B200_DEV static Coop load(const Ctx &c, const uint32_t *w) { return {c.active ? w[c.g] : 0u}; }
    B200_DEV void store(const Ctx &c, uint32_t *w) const {
        if (c.active && c.gbase == 0) w[c.g] = v;
    }
    B200_DEV static Coop zero() { return {0u}; }
    B200_DEV static Coop one(const Ctx &c) { return {c.one}; }
    B200_DEV static Coop add(const Ctx &c, Coop a, Coop b) { return {Ops::add(c, a.v, b.v)}; }
    B200_DEV static Coop sub(const Ctx &c, Coop a, Coop b) { return {Ops::sub(c, a.v, b.v)}; }
    B200_DEV static Coop dbl(const Ctx &c, Coop a) { return add(c, a, a); }
    B200_DEV static Coop mul(const Ctx &c, Coop a, Coop b) { return {Ops::mul(c, a.v, b.v)}; }
    // whole-warp vote: the result is warp-uniform for the compiler too (branches on it stay convergent)
    B200_DEV static bool is_zero(const Ctx &c, Coop a) { return __ballot_sync(COOP_FULL, c.active && a.v != 0u) == 0u; }
};

// Fq2 = Fq[u] / (u^2 + 5): c0 in lanes 0..11, c1 in lanes 16..27
template <class P>
struct Coop<Fp2<Fp<P>>> {
    static constexpr int N = P::N, GW = 16, WORDS = 2 * N;
    static_assert(N <= 16, "Fq2 over a field of at most 16 limbs");
    using Ops = CoopOps<P, GW>;
    using Ctx = CoopCtx<P, GW>;
    uint32_t v;
    B200_DEV static Coop load(const Ctx &c, const uint32_t *w) { return {c.active ? w[(c.gbase ? N : 0) + c.g] : 0u}; }
    B200_DEV void store(const Ctx &c, uint32_t *w) const {
        if (c.active) w[(c.gbase ? N : 0) + c.g] = v;
    }
    B200_DEV static Coop zero() { return {0u}; }
    B200_DEV static Coop one(const Ctx &c) { return {c.gbase ? 0u : c.one}; }
    B200_DEV static Coop add(const Ctx &c, Coop a, Coop b) { return {Ops::add(c, a.v, b.v)}; }
    B200_DEV static Coop sub(const Ctx &c, Coop a, Coop b) { return {Ops::sub(c, a.v, b.v)}; }
    B200_DEV static Coop dbl(const Ctx &c, Coop a) { return add(c, a, a); }
    // schoolbook over the two halves: (a0 b0 | a1 b1), then (a0 b1 | a1 b0); c0 = a0 b0 - 5 a1 b1, c1 = a0 b1 + a1 b0
    B200_DEV static Coop mul(const Ctx &c, Coop a, Coop b) {
        const uint32_t v01 = Ops::mul(c, a.v, b.v);
        const uint32_t bs = __shfl_xor_sync(COOP_FULL, b.v, 16);
        const uint32_t t01 = Ops::mul(c, a.v, bs);
        const uint32_t vx = __shfl_xor_sync(COOP_FULL, v01, 16), tx = __shfl_xor_sync(COOP_FULL, t01, 16);
        // both halves run both formulas (the operations are collective); each keeps its own
        const uint32_t v1 = c.gbase ? v01 : vx, v0 = c.gbase ? vx : v01;     // a1 b1, a0 b0 in every half
        uint32_t f = Ops::add(c, v1, v1);
        f = Ops::add(c, f, f);
        f = Ops::add(c, f, v1);                                            // 5 a1 b1
        const uint32_t c0 = Ops::sub(c, v0, f);
        const uint32_t c1 = Ops::add(c, t01, tx);
        return {c.gbase ? c1 : c0};
    }
    B200_DEV static bool is_zero(const Ctx &c, Coop a) { return __ballot_sync(COOP_FULL, c.active && a.v != 0u) == 0u; }
};

// ---- XYZZ point operations on a block of four warps ---------------------------------------------------
// The point and the scratch values live in shared memory as word images (XYZZMem layout: x | y | zz | zzz);
// every thread of the block calls these functions (they contain __syncthreads()).
constexpr int COOP_THREADS = 128;

template <class F>
struct alignas(16) CoopSm {
    static constexpr int W = Coop<F>::WORDS;
    uint32_t pt[4][W];                               // the running point: x, y, zz, zzz
    uint32_t in[4][W];                               // the addend (16-byte aligned: the TMA bulk copy lands here)
    uint32_t t[8][W];                                // round outputs
    unsigned long long bar;                          // mbarrier of the bulk copies
};

// ---- TMA (bulk asynchronous copy) staging of a point into shared memory ---------------------------------------
// One elected thread arms the mbarrier with the byte count and issues ONE cp.async.bulk for the whole 192 / 384-byte
// XYZZ record; the copy engine moves it while the block is still doubling, and every thread waits on the barrier's phase
// when the addend is needed.  (SASS: UBLKCP + SYNCS, profiles/r2_sass_tma_coop.txt.)
B200_DEV uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
B200_DEV void tma_bar_init(unsigned long long *bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
B200_DEV void tma_load(void *smem_dst, const void *gmem_src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier generic reads of the slot are done (barrier before)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
B200_DEV void tma_wait(unsigned long long *bar, uint32_t phase) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done)
                     : "r"(smem_addr(bar)), "r"(phase)
                     : "memory");
    }
}

template <class F>
struct CoopPoint {
    using C = Coop<F>;
    using Ctx = typename C::Ctx;
    static constexpr int W = C::WORDS;

    // pt <- 2 pt   (dbl-2008-s-1 in 3 rounds)
    B200_DEV static void dbl(const Ctx &c, CoopSm<F> &sm) {
        const int warp = threadIdx.x >> 5;
        const C x = C::load(c, sm.pt[0]), y = C::load(c, sm.pt[1]), zz = C::load(c, sm.pt[2]), zzz = C::load(c, sm.pt[3]);
        if (C::is_zero(c, zz)) return;               // infinity (uniform over the block)
        const C u = C::dbl(c, y);
        // every warp multiplies in every round (a branch on the warp index around the shuffles would make the
        // compiler guard each of them with a WARPSYNC); spare warps repeat a neighbour's product and skip the store
        {
            const C a = (warp & 1) ? x : u;
            const C r = C::mul(c, a, a);                                 // v, xx
            if (warp < 2) r.store(c, sm.t[warp]);
        }
        __syncthreads();
        const C v = C::load(c, sm.t[0]), xx = C::load(c, sm.t[1]);
        const C m = C::add(c, C::dbl(c, xx), xx);
        {
            const C a = warp == 0 ? u : warp == 1 ? x : warp == 2 ? m : v;
            const C b = warp == 0 ? v : warp == 1 ? v : warp == 2 ? m : zz;
            C::mul(c, a, b).store(c, sm.t[4 + warp]);                // w, s, mm, zz3
        }
        __syncthreads();
        const C w = C::load(c, sm.t[4]), s = C::load(c, sm.t[5]), mm = C::load(c, sm.t[6]), zz3 = C::load(c, sm.t[7]);
        const C x3 = C::sub(c, mm, C::dbl(c, s));
        {
            const C a = warp == 0 ? m : w;
            const C b = warp == 0 ? C::sub(c, s, x3) : warp == 1 ? y : zzz;
            const C r = C::mul(c, a, b);                                 // t1, t2, zzz3
            if (warp < 3) r.store(c, sm.t[warp]);
        }
        __syncthreads();
        {
            const C t1 = C::load(c, sm.t[0]), t2 = C::load(c, sm.t[1]), zzz3 = C::load(c, sm.t[2]);
            const C y3 = C::sub(c, t1, t2);
            if (warp == 0) {
                x3.store(c, sm.pt[0]);
                y3.store(c, sm.pt[1]);
                zz3.store(c, sm.pt[2]);
                zzz3.store(c, sm.pt[3]);
            }
        }
        __syncthreads();
    }

    // pt <- pt + in   (add-2008-s in 4 rounds; infinity operands, P + P and P + (-P) exact)
    B200_DEV static void add(const Ctx &c, CoopSm<F> &sm) {
        const int warp = threadIdx.x >> 5;
        const C x2 = C::load(c, sm.in[0]), y2 = C::load(c, sm.in[1]), zz2 = C::load(c, sm.in[2]), zzz2 = C::load(c, sm.in[3]);
        if (C::is_zero(c, zz2)) return;
        const C x1 = C::load(c, sm.pt[0]), y1 = C::load(c, sm.pt[1]), zz1 = C::load(c, sm.pt[2]), zzz1 = C::load(c, sm.pt[3]);
        if (C::is_zero(c, zz1)) {
            __syncthreads();                         // everyone has read pt / in
            if (warp == 0) {
                x2.store(c, sm.pt[0]);
                y2.store(c, sm.pt[1]);
                zz2.store(c, sm.pt[2]);
                zzz2.store(c, sm.pt[3]);
            }
            __syncthreads();
            return;
        }
        {
            const C a = warp == 0 ? x1 : warp == 1 ? x2 : warp == 2 ? y1 : y2;
            const C b = warp == 0 ? zz2 : warp == 1 ? zz1 : warp == 2 ? zzz2 : zzz1;
            C::mul(c, a, b).store(c, sm.t[warp]);                    // u1, u2, s1, s2
        }
        __syncthreads();
        const C u1 = C::load(c, sm.t[0]), u2 = C::load(c, sm.t[1]), s1 = C::load(c, sm.t[2]), s2 = C::load(c, sm.t[3]);
        const C p = C::sub(c, u2, u1), r = C::sub(c, s2, s1);
        if (C::is_zero(c, p)) {
            if (C::is_zero(c, r)) {
                __syncthreads();                     // every warp has read the round outputs dbl() overwrites
                dbl(c, sm);                          // P + P
            } else {
                __syncthreads();
                if (warp == 0) C::zero().store(c, sm.pt[2]);         // P + (-P): infinity is zz == 0
                __syncthreads();
            }
            return;
        }
        {
            const C a = warp == 0 ? p : warp == 1 ? r : warp == 2 ? zz1 : zzz1;
            const C b = warp == 0 ? p : warp == 1 ? r : warp == 2 ? zz2 : zzz2;
            C::mul(c, a, b).store(c, sm.t[4 + warp]);                // pp, rr, zza, zzza
        }
        __syncthreads();
        const C pp = C::load(c, sm.t[4]), rr = C::load(c, sm.t[5]), zza = C::load(c, sm.t[6]), zzza = C::load(c, sm.t[7]);
        {
            const C a = warp == 0 ? p : warp == 1 ? u1 : zza;
            const C r = C::mul(c, a, pp);                                // ppp, q, zz3
            if (warp < 3) r.store(c, sm.t[warp]);
        }
        __syncthreads();
        const C ppp = C::load(c, sm.t[0]), q = C::load(c, sm.t[1]), zz3 = C::load(c, sm.t[2]);
        const C x3 = C::sub(c, C::sub(c, rr, ppp), C::dbl(c, q));
        {
            const C a = warp == 0 ? r : warp == 1 ? s1 : zzza;
            const C b = warp == 0 ? C::sub(c, q, x3) : ppp;
            const C o = C::mul(c, a, b);                                 // t1, t2, zzz3
            if (warp < 3) o.store(c, sm.t[4 + warp]);
        }
        __syncthreads();
        {
            const C t1 = C::load(c, sm.t[4]), t2 = C::load(c, sm.t[5]), zzz3 = C::load(c, sm.t[6]);
            const C y3 = C::sub(c, t1, t2);
            if (warp == 0) {
                x3.store(c, sm.pt[0]);
                y3.store(c, sm.pt[1]);
                zz3.store(c, sm.pt[2]);
                zzz3.store(c, sm.pt[3]);
            }
        }
        __syncthreads();
    }

    // shared <- global XYZZ image (nullptr: infinity), all threads
    B200_DEV static void fetch(uint32_t (*dst)[W], const XYZZMem<F> *src) {
        const uint32_t *s = reinterpret_cast<const uint32_t *>(src);
        __syncthreads();                             // the early exits of dbl / add leave without a barrier
        for (int i = threadIdx.x; i < 4 * W; i += blockDim.x) dst[i / W][i % W] = src ? __ldg(s + i) : 0u;
        __syncthreads();
    }
    // the same through the copy engine: issue (one thread) ... later ... wait (all threads)
    B200_DEV static void fetch_issue(CoopSm<F> &sm, const XYZZMem<F> *src) {
        __syncthreads();                             // everyone is done with the previous addend
        if (threadIdx.x == 0) tma_load(sm.in, src, (uint32_t)sizeof(XYZZMem<F>), &sm.bar);
    }
    B200_DEV static void fetch_wait(CoopSm<F> &sm, uint32_t &phase) {
        tma_wait(&sm.bar, phase);
        phase ^= 1u;
    }
};

// ---- the Horner combine --------------------------------------------------------------------------------
// One block of four warps: total = sum_{w in [w_lo, w_hi)} 2^((w - w_lo) c) window_sums[w], then `shift` further
// doublings (the weight of window w_lo inside the whole MSM when the windows are combined in groups), plus the
// optional addends (XYZZ images; e.g. the sum over the unit scalars, or another group's combined value).
// out_jac != nullptr: the result leaves as an arkworks GroupProjective (X ZZ, Y ZZZ, ZZ); else as XYZZ in out_xyzz.
template <class F>
__global__ void __launch_bounds__(COOP_THREADS) k_window_combine_coop(const XYZZMem<F> *__restrict__ window_sums, int w_lo,
                                                                      int w_hi, int c_bits, int shift,
                                                                      const XYZZMem<F> *extra0, const XYZZMem<F> *extra1,
                                                                      JacobianMem<F> *out_jac, XYZZMem<F> *out_xyzz) {
    using CP = CoopPoint<F>;
    using C = Coop<F>;
    constexpr int W = C::WORDS;
    __shared__ CoopSm<F> sm;
    const typename C::Ctx c = C::Ctx::make();
    static_assert(sizeof(XYZZMem<F>) % 16 == 0 && (4 * W * 4) % 16 == 0, "bulk copies move 16-byte multiples to 16-byte aligned slots");
    if (threadIdx.x == 0) tma_bar_init(&sm.bar);
    CP::fetch(sm.pt, nullptr);                       // (contains the barrier that publishes the mbarrier)
    uint32_t phase = 0;
    // window sum w is staged by the copy engine while the c doublings that precede its addition run
    if (w_hi > w_lo) CP::fetch_issue(sm, window_sums + (w_hi - 1));
    for (int w = w_hi - 1; w >= w_lo; w--) {
        CP::fetch_wait(sm, phase);
        CP::add(c, sm);
        if (w > w_lo) CP::fetch_issue(sm, window_sums + (w - 1));
        const int doublings = w > w_lo ? c_bits : shift;
        for (int k = 0; k < doublings; k++) CP::dbl(c, sm);
    }
    if (extra0) {
        CP::fetch(sm.in, extra0);
        CP::add(c, sm);
    }
    if (extra1) {
        CP::fetch(sm.in, extra1);
        CP::add(c, sm);
    }
    const int warp = threadIdx.x >> 5;
    if (out_xyzz) {
        uint32_t *o = reinterpret_cast<uint32_t *>(out_xyzz);
        for (int i = threadIdx.x; i < 4 * W; i += blockDim.x) o[i] = sm.pt[i / W][i % W];
    }
    if (out_jac) {
        uint32_t *o = reinterpret_cast<uint32_t *>(out_jac);
        const C zz = C::load(c, sm.pt[2]);
        if (C::is_zero(c, zz)) {                     // GroupProjective::zero() = (1, 1, 0)
            if (warp == 0) {
                C::one(c).store(c, o);
                C::one(c).store(c, o + W);
                C::zero().store(c, o + 2 * W);
            }
        } else {
            const C a = (warp & 1) ? C::load(c, sm.pt[1]) : C::load(c, sm.pt[0]);
            const C b = (warp & 1) ? C::load(c, sm.pt[3]) : zz;
            const C r = C::mul(c, a, b);                                 // X ZZ, Y ZZZ
            if (warp < 2) r.store(c, o + warp * W);
            if (warp == 2) zz.store(c, o + 2 * W);
        }
    }
}

// element-wise field operations through the cooperative routines (parity tests; one warp per element)
template <class F>
__global__ void __launch_bounds__(COOP_THREADS) k_coop_field_op(int op, const typename F::Mem *__restrict__ a,
                                                                const typename F::Mem *__restrict__ b, uint32_t n,
                                                                typename F::Mem *__restrict__ out) {
    using C = Coop<F>;
    const typename C::Ctx c = C::Ctx::make();
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;                              // whole warps leave together
    const C x = C::load(c, reinterpret_cast<const uint32_t *>(a + i)), y = C::load(c, reinterpret_cast<const uint32_t *>(b + i));
    C r;
    switch (op) {
    case 0: r = C::add(c, x, y); break;
    case 1: r = C::sub(c, x, y); break;
    case 2: r = C::mul(c, x, y); break;
    case 3: r = C::mul(c, x, x); break;
    case 5: r = C::sub(c, C::zero(), x); break;
    default: r = C::dbl(c, x); break;
    }
    r.store(c, reinterpret_cast<uint32_t *>(out + i));
}

}  // namespace b200
