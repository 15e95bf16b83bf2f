// BLS12-377 optimal-ate multi-pairing for sm_100a (round-1 version: one thread per pair).
//
// Device replacement for ark-ec 0.1.0 `Bls12::product_of_pairings` as called at
//   crates/bls-crypto/src/bls/signature.rs:149  (batch_verify_hashes, N + 1 pairs)
//   crates/bls-crypto/src/bls/public.rs:102     (verify_sig, 2 pairs)
// following SURVEY.md appendix A.3: G2 line coefficients in homogeneous projective
// coordinates (doubling / addition step), D-twist line evaluation, f <- f^2 * lines over the
// bits of x = 0x8508c00000000001 below the MSB, final exponentiation by the 2016/130 chain.
// The value before and after the final exponentiation is the one arkworks computes (same loop,
// same line scaling, same exponent 3 (p^12 - 1) / r), so GT bytes can be compared, not only
// the `== 1` outcome.
//
// Structure: k_miller_loop (thread per pair, independent Miller values) -> k_fq12_product
// (one block, strided products + tree) -> k_final_exp (one thread).  A product of Miller values
// is exact, so splitting the pairs across threads (or GPUs) does not change the result.
// Tower: Fq2 = Fq[u]/(u^2 + 5), Fq6 = Fq2[v]/(v^3 - u), Fq12 = Fq6[w]/(w^2 - v).
// Round 2 replaces the per-thread tower by a warp-cooperative one.
#pragma once
#include "ec.cuh"
#include "pairing_params_gen.cuh"

namespace b200 {

using PFq = Fq377;
using PFq2 = Fp2<Fq377>;

B200_DEV PFq pfq_const(const uint32_t *w) {
    PFq r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = w[i];
    return r;
}

struct Fq6 {
    PFq2 c0, c1, c2;
    struct alignas(16) Mem {
        PFq2::Mem c0, c1, c2;
    };
};
struct Fq12 {
    Fq6 c0, c1;
    struct alignas(16) Mem {
        Fq6::Mem c0, c1;
    };
};

B200_DEV Fq6 f6_load(const Fq6::Mem &m) { return {PFq2::from_ark(m.c0), PFq2::from_ark(m.c1), PFq2::from_ark(m.c2)}; }
B200_DEV Fq6::Mem f6_store(const Fq6 &a) { return {a.c0.to_ark(), a.c1.to_ark(), a.c2.to_ark()}; }
B200_DEV Fq12 f12_load(const Fq12::Mem &m) { return {f6_load(m.c0), f6_load(m.c1)}; }
B200_DEV Fq12::Mem f12_store(const Fq12 &a) { return {f6_store(a.c0), f6_store(a.c1)}; }

B200_DEV PFq2 f2_mul_xi(const PFq2 &a) { return {PFq2::mul5(a.c1).neg(), a.c0}; }      // * u, u^2 = -5
B200_DEV PFq2 f2_scale(const PFq2 &a, const PFq &k) { return {a.c0 * k, a.c1 * k}; }
B200_DEV PFq2 f2_conj(const PFq2 &a) { return {a.c0, a.c1.neg()}; }
B200_DEV PFq2 f2_inv(const PFq2 &a) {
    PFq n = (a.c0.sqr() + PFq2::mul5(a.c1.sqr())).inv();
    return {a.c0 * n, (a.c1 * n).neg()};
}

B200_DEV Fq6 f6_zero() { return {PFq2::zero(), PFq2::zero(), PFq2::zero()}; }
B200_DEV Fq6 f6_one() { return {PFq2::one(), PFq2::zero(), PFq2::zero()}; }
B200_DEV Fq6 f6_add(const Fq6 &a, const Fq6 &b) { return {a.c0 + b.c0, a.c1 + b.c1, a.c2 + b.c2}; }
B200_DEV Fq6 f6_sub(const Fq6 &a, const Fq6 &b) { return {a.c0 - b.c0, a.c1 - b.c1, a.c2 - b.c2}; }
B200_DEV Fq6 f6_neg(const Fq6 &a) { return {a.c0.neg(), a.c1.neg(), a.c2.neg()}; }
B200_DEV Fq6 f6_mul_v(const Fq6 &a) { return {f2_mul_xi(a.c2), a.c0, a.c1}; }

// Karatsuba over Fq2: 6 products.  Out of line, arguments by value (see Fp::mul_outline).
static __device__ __noinline__ Fq6 f6_mul(Fq6 a, Fq6 b) {
    PFq2 v0 = a.c0 * b.c0, v1 = a.c1 * b.c1, v2 = a.c2 * b.c2;
    PFq2 t0 = (a.c1 + a.c2) * (b.c1 + b.c2) - v1 - v2;
    PFq2 t1 = (a.c0 + a.c1) * (b.c0 + b.c1) - v0 - v1;
    PFq2 t2 = (a.c0 + a.c2) * (b.c0 + b.c2) - v0 - v2;
    return {v0 + f2_mul_xi(t0), t1 + f2_mul_xi(v2), t2 + v1};
}
static __device__ __noinline__ Fq6 f6_inv(Fq6 a) {
    PFq2 t0 = a.c0.sqr() - f2_mul_xi(a.c1 * a.c2);
    PFq2 t1 = f2_mul_xi(a.c2.sqr()) - a.c0 * a.c1;
    PFq2 t2 = a.c1.sqr() - a.c0 * a.c2;
    PFq2 d = a.c0 * t0 + f2_mul_xi(a.c2 * t1 + a.c1 * t2);
    PFq2 di = f2_inv(d);
    return {t0 * di, t1 * di, t2 * di};
}

B200_DEV Fq12 f12_one() { return {f6_one(), f6_zero()}; }
static __device__ __noinline__ Fq12 f12_mul(Fq12 a, Fq12 b) {
    Fq6 t0 = f6_mul(a.c0, b.c0);
    Fq6 t1 = f6_mul(a.c1, b.c1);
    Fq6 m = f6_mul(f6_add(a.c0, a.c1), f6_add(b.c0, b.c1));
    return {f6_add(t0, f6_mul_v(t1)), f6_sub(f6_sub(m, t0), t1)};
}
// complex squaring: 2 Fq6 products
static __device__ __noinline__ Fq12 f12_sqr(Fq12 a) {
    Fq6 ab = f6_mul(a.c0, a.c1);
    Fq6 m = f6_mul(f6_add(a.c0, a.c1), f6_add(a.c0, f6_mul_v(a.c1)));
    return {f6_sub(f6_sub(m, ab), f6_mul_v(ab)), f6_add(ab, ab)};
}
B200_DEV Fq12 f12_conj(const Fq12 &a) { return {a.c0, f6_neg(a.c1)}; }
static __device__ __noinline__ Fq12 f12_inv(Fq12 a) {
    Fq6 d = f6_sub(f6_mul(a.c0, a.c0), f6_mul_v(f6_mul(a.c1, a.c1)));
    Fq6 di = f6_inv(d);
    return {f6_mul(a.c0, di), f6_neg(f6_mul(a.c1, di))};
}
B200_DEV bool f12_is_one(const Fq12 &a) {
    return a.c0.c0 == PFq2::one() && a.c0.c1.is_zero() && a.c0.c2.is_zero() && a.c1.c0.is_zero() && a.c1.c1.is_zero() &&
           a.c1.c2.is_zero();
}

// frob^j: conjugate the Fq2 coefficients when j is odd, scale the w^k coefficient by xi^(k (p^j - 1)/6)
B200_DEV PFq2 frob_coeff(int j, int k) {
    return {pfq_const(PAIRING_FROB[j - 1][k][0]), pfq_const(PAIRING_FROB[j - 1][k][1])};
}
static __device__ __noinline__ Fq12 f12_frob(Fq12 a, int j) {
    bool odd = j & 1;
    auto fr = [&](const PFq2 &c, int k) { return (odd ? f2_conj(c) : c) * frob_coeff(j, k); };
    // w-power basis: c0 = (w^0, w^2, w^4), c1 = (w^1, w^3, w^5)
    return {{fr(a.c0.c0, 0), fr(a.c0.c1, 2), fr(a.c0.c2, 4)}, {fr(a.c1.c0, 1), fr(a.c1.c1, 3), fr(a.c1.c2, 5)}};
}
// f^x, x = PAIRING_X (square-and-multiply, MSB first)
static __device__ __noinline__ Fq12 f12_exp_by_x(Fq12 f) {
    Fq12 r = f;
    launder(r);
#pragma unroll 1
    for (int b = 62; b >= 0; b--) {
        r = f12_sqr(r);
        launder(r);
        if ((PAIRING_X >> b) & 1ull) {
            r = f12_mul(r, f);
            launder(r);
        }
    }
    return r;
}

// ---- Miller loop ------------------------------------------------------------------------------
struct G2Hom {
    PFq2 x, y, z;
};
struct LineCoeffs {
    PFq2 c0, c1, c2;
};
struct StepOut {                                    // returned by value (no pointer out-parameters)
    G2Hom r;
    LineCoeffs l;
};

static __device__ __noinline__ StepOut g2_doubling_step(G2Hom r) {
    const PFq two_inv = pfq_const(PAIRING_TWO_INV);
    const PFq2 twist_b = {PFq::zero(), pfq_const(PAIRING_TWIST_B_C1)};
    PFq2 a = f2_scale(r.x * r.y, two_inv);
    PFq2 b = r.y.sqr();
    PFq2 c = r.z.sqr();
    PFq2 e = twist_b * (c.dbl() + c);
    PFq2 f = e.dbl() + e;
    PFq2 g = f2_scale(b + f, two_inv);
    PFq2 h = (r.y + r.z).sqr() - (b + c);
    PFq2 i = e - b;
    PFq2 j = r.x.sqr();
    PFq2 e2 = e.sqr();
    StepOut o;
    o.r.x = a * (b - f);
    o.r.y = g.sqr() - (e2.dbl() + e2);
    o.r.z = b * h;
    o.l = {h.neg(), j.dbl() + j, i};
    return o;
}
static __device__ __noinline__ StepOut g2_addition_step(G2Hom r, PFq2 qx, PFq2 qy) {
    PFq2 theta = r.y - qy * r.z;
    PFq2 lam = r.x - qx * r.z;
    PFq2 c = theta.sqr();
    PFq2 d = lam.sqr();
    PFq2 e = lam * d;
    PFq2 f = r.z * c;
    PFq2 g = r.x * d;
    PFq2 h = e + f - g.dbl();
    StepOut o;
    o.r.x = lam * h;
    o.r.y = theta * (g - h) - e * r.y;
    o.r.z = r.z * e;
    PFq2 j = theta * qx - lam * qy;
    o.l = {lam, theta.neg(), j};
    return o;
}
// D-twist line evaluation at P, then f * (c0 + (c1 + c2 v) w)   [mul_by_034]
static __device__ __noinline__ Fq12 miller_ell(Fq12 f, LineCoeffs l, PFq px, PFq py) {
    Fq12 s = {{f2_scale(l.c0, py), PFq2::zero(), PFq2::zero()}, {f2_scale(l.c1, px), l.c2, PFq2::zero()}};
    return f12_mul(f, s);
}

// out[i] = Miller value of (P_i, Q_i); pairs with an infinite member give 1 (arkworks skips them)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_miller_loop(const AffineMem<PFq> *__restrict__ g1,
                                                         const AffineMem<PFq2> *__restrict__ g2, uint32_t n,
                                                         Fq12::Mem *__restrict__ out) {
    uint32_t i = blockIdx.x * THREADS + threadIdx.x;
    if (i >= n) return;
    Affine<PFq> p = Affine<PFq>::from_ark(ldg_mem(g1 + i));
    Affine<PFq2> q = Affine<PFq2>::from_ark(ldg_mem(g2 + i));
    Fq12 f = f12_one();
    launder(f);
    if (!p.is_inf() && !q.is_inf()) {
        G2Hom r = {q.x, q.y, PFq2::one()};
        launder(r);
#pragma unroll 1
        for (int b = 62; b >= 0; b--) {
            f = f12_sqr(f);
            launder(f);
            StepOut so = g2_doubling_step(r);
            r = so.r;
            launder(r);
            f = miller_ell(f, so.l, p.x, p.y);
            launder(f);
            if ((PAIRING_X >> b) & 1ull) {
                so = g2_addition_step(r, q.x, q.y);
                r = so.r;
                launder(r);
                f = miller_ell(f, so.l, p.x, p.y);
                launder(f);
            }
        }
    }
    out[i] = f12_store(f);
}

// single block: vals[0] = prod vals[0..n)   (in place; strided products, then a tree)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_fq12_product(Fq12::Mem *__restrict__ vals, uint32_t n) {
    if (blockIdx.x) return;
    uint32_t t = threadIdx.x;
    if (t < n) {
        Fq12 acc = f12_load(vals[t]);
        launder(acc);
        for (uint32_t i = t + THREADS; i < n; i += THREADS) {
            acc = f12_mul(acc, f12_load(vals[i]));
            launder(acc);
        }
        vals[t] = f12_store(acc);
    }
    __syncthreads();
    uint32_t live = n < THREADS ? n : THREADS;
    while (live > 1) {
        uint32_t half = (live + 1) / 2;
        if (t + half < live) vals[t] = f12_store(f12_mul(f12_load(vals[t]), f12_load(vals[t + half])));
        __syncthreads();
        live = half;
    }
}

#ifdef B200_WITH_CROSSCHECKS                        // cross-check kernels are not part of the shipped library (build.py --crosschecks)
// Bls12::final_exponentiation (eprint 2016/130 table 1 chain), one thread
__global__ void k_final_exp(const Fq12::Mem *__restrict__ in, Fq12::Mem *__restrict__ out, int *__restrict__ is_one) {
    if (blockIdx.x || threadIdx.x) return;
    Fq12 f = f12_load(in[0]);
    launder(f);
    Fq12 r = f12_mul(f12_conj(f), f12_inv(f));
    launder(r);
    r = f12_mul(f12_frob(r, 2), r);                  // easy part: (p^6 - 1)(p^2 + 1)
    launder(r);
    Fq12 y0 = f12_conj(f12_sqr(r));                  // cyclotomic: conjugate == inverse
    launder(y0);
    Fq12 y5 = f12_exp_by_x(r);
    launder(y5);
    Fq12 y1 = f12_sqr(y5);
    launder(y1);
    Fq12 y3 = f12_mul(y0, y5);
    launder(y3);
    y0 = f12_exp_by_x(y3);
    launder(y0);
    Fq12 y2 = f12_exp_by_x(y0);
    launder(y2);
    Fq12 y4 = f12_mul(f12_exp_by_x(y2), y1);
    launder(y4);
    y1 = f12_exp_by_x(y4);
    launder(y1);
    y3 = f12_conj(y3);
    launder(y3);
    y1 = f12_mul(f12_mul(y1, y3), r);
    launder(y1);
    y3 = f12_conj(r);
    launder(y3);
    y0 = f12_frob(f12_mul(y0, r), 3);
    launder(y0);
    y4 = f12_frob(f12_mul(y4, y3), 1);
    launder(y4);
    y5 = f12_frob(f12_mul(y5, y2), 2);
    launder(y5);
    Fq12 res = f12_mul(f12_mul(f12_mul(y5, y0), y4), y1);
    launder(res);
    if (out) out[0] = f12_store(res);
    if (is_one) *is_one = f12_is_one(res) ? 1 : 0;
}
#endif

}  // namespace b200
