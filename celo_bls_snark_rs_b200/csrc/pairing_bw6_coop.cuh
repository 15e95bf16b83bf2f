// BW6-761 final exponentiation on warp-cooperative arithmetic (round 2): the latency form of k_bw6_final_exp.
//
// pairing_bw6.cuh gives every one of the 36 coefficient products of an F_q^6 product its own THREAD: a product round
// costs one 761-bit Montgomery product issued by one warp (1152 wide multiply-adds, ~2.4 us) plus the six sums, 5.6 us
// per F_q^6 product, 4.65 ms for the ~830 products of the final exponentiation -- the largest piece of a Groth16
// verification.  Here (coop.cuh) one field element is spread over a warp, one limb per lane, and
//
//   * WARP e of a six-warp block owns OUTPUT COEFFICIENT e of every F_q^6 product: c_e = sum_i a_i b_((e - i) mod 6),
//     wrapped terms times -4 (w^6 = -4);
//   * the six terms share ONE Montgomery reduction: CoopOps::accumulate adds the column sums of each a_i b_j into the same
//     96-bit accumulators, CoopOps::reduce runs once (6 p^2 < R p: the 7 spare bits of the modulus);
//   * every value is stored together with -4 x itself, so wrapped terms are ordinary non-negative products and the sum
//     needs no subtraction.
//
// An F_q^6 product is then 6 x 24 + 72 multiply-add steps per lane and two barriers, nothing staged per coefficient
// product.  Same chain, same exponent (R0 + q R1 over the joint sparse form), same values as k_bw6_final_exp.
#pragma once
#include "coop.cuh"
#include "pairing_bw6.cuh"

namespace b200 {

constexpr int BW6C_THREADS = 192;                    // six warps
constexpr int BW6C_W = 24;                           // words per coefficient

struct alignas(16) Bw6CoopVal {
    uint32_t c[6][BW6C_W];                           // coefficients, power basis
    uint32_t m4[6][BW6C_W];                          // -4 x coefficient
};

struct Bw6Coop {
    using C = Coop<BFq>;
    using Ops = C::Ops;
    using Ctx = C::Ctx;

    // all six warps call; O may alias A or B
    B200_DEV static void store(const Ctx &c, Bw6CoopVal &O, int e, C r) {
        C n = C::dbl(c, C::dbl(c, r));
        C r4 = C::sub(c, C::zero(), n);
        __syncthreads();                             // every read of the operands is done
        r.store(c, O.c[e]);
        r4.store(c, O.m4[e]);
        __syncthreads();
    }
    // O = A * B; exponents of B restricted to jmask (dense 0x3f, line 0x0d)
    B200_DEV static void mul(const Ctx &c, const Bw6CoopVal &A, const Bw6CoopVal &B, Bw6CoopVal &O, uint32_t jmask) {
        const int e = threadIdx.x >> 5;
        Ops::Acc s;
#pragma unroll 1
        for (int j = 0; j < 6; j++) {
            if (!((jmask >> j) & 1u)) continue;
            int i = e - j;
            const bool wrap = i < 0;
            i += wrap ? 6 : 0;
            const C a = C::load(c, A.c[i]), b = C::load(c, wrap ? B.m4[j] : B.c[j]);
            Ops::accumulate(c, s, a.v, b.v);
        }
        store(c, O, e, C{Ops::reduce(c, s)});
    }
    // O = frob^j(A) (j = 1, 2) or the conjugate (j = 3)
    B200_DEV static void frob(const Ctx &c, const Bw6CoopVal &A, Bw6CoopVal &O, int j) {
        const int e = threadIdx.x >> 5;
        C v = C::load(c, A.c[e]);
        if (j == 3) {
            if (e & 1) v = C::sub(c, C::zero(), v);
        } else {
            v = C::mul(c, v, C::load(c, BW6_GAMMA[j - 1][e]));
        }
        store(c, O, e, v);
    }
    B200_DEV static void load_global(const Ctx &c, Bw6CoopVal &O, const BImg *src) {
        const int e = threadIdx.x >> 5;
        store(c, O, e, C::load(c, reinterpret_cast<const uint32_t *>(src + e)));
    }
};

struct alignas(16) Bw6CoopFinalScratch {
    Bw6CoopVal V[16];
    uint32_t inv[BW6C_W];
    int flags[6];
};

// One block of six warps: product of `count` Miller values, final exponentiation (the chain of k_bw6_final_exp).
__global__ void __launch_bounds__(BW6C_THREADS) k_bw6_final_exp_coop(const BImg *__restrict__ in, uint32_t count, BImg *__restrict__ out,
                                                                     int *__restrict__ is_one) {
    __shared__ Bw6CoopFinalScratch S;
    using K = Bw6Coop;
    using C = K::C;
    const K::Ctx c = K::Ctx::make();
    const int t = threadIdx.x, e = t >> 5;
    enum { X = 0, CJ, N, N1, N2, M, R, ACC, T0 /* .. T0 + 7: table of the joint exponentiation */ };
    Bw6CoopVal *V = S.V;
    K::load_global(c, V[X], in);
#pragma unroll 1
    for (uint32_t k = 1; k < count; k++) {
        K::load_global(c, V[CJ], in + 6 * (size_t)k);
        K::mul(c, V[X], V[CJ], V[X], BW6_DENSE);
    }
    // f^-1: n = f conj(f) in F_q^3, m = n^q n^(q^2), n m in F_q
    K::frob(c, V[X], V[CJ], 3);
    K::mul(c, V[X], V[CJ], V[N], BW6_DENSE);
    K::frob(c, V[N], V[N1], 1);
    K::frob(c, V[N], V[N2], 2);
    K::mul(c, V[N1], V[N2], V[M], BW6_DENSE);
    K::mul(c, V[N], V[M], V[N1], BW6_DENSE);         // N1[0] = norm in F_q
    if (t == 0) {                                    // the one base-field inversion: per-thread safegcd (fp.cuh)
        BFq n;
#pragma unroll
        for (int i = 0; i < BW6C_W; i++) n.l[i] = V[N1].c[0][i];
        n = n.inv();
#pragma unroll
        for (int i = 0; i < BW6C_W; i++) S.inv[i] = n.l[i];
    }
    K::mul(c, V[CJ], V[M], V[N2], BW6_DENSE);        // conj(f) m  (the barriers inside also publish S.inv)
    K::store(c, V[N2], e, C::mul(c, C::load(c, V[N2].c[e]), C::load(c, S.inv)));      // f^-1
    // easy part: r = (conj(f) / f)^(q + 1)
    K::mul(c, V[CJ], V[N2], V[R], BW6_DENSE);
    K::frob(c, V[R], V[N], 1);
    K::mul(c, V[N], V[R], V[R], BW6_DENSE);
    // hard part: r^(R0 + q R1) = (r^-1)^(-R0) (r^q)^R1, one joint square-and-multiply over the signed digit pairs
    K::frob(c, V[R], V[T0], 3);                      // f' = r^-1 (cyclotomic: the conjugate)
    K::frob(c, V[R], V[T0 + 1], 1);                  // g = r^q
    K::mul(c, V[T0], V[T0 + 1], V[T0 + 2], BW6_DENSE);           // f' g
    K::frob(c, V[T0 + 1], V[N], 3);                  // g^-1
    K::mul(c, V[T0], V[N], V[T0 + 3], BW6_DENSE);    // f' g^-1
#pragma unroll 1
    for (int k = 0; k < 4; k++) K::frob(c, V[T0 + k], V[T0 + 4 + k], 3);          // the four inverses
    {
        const int top = BW6_HARD_JSF[BW6_HARD_JSF_LEN - 1];
        const Bw6CoopVal &src = V[T0 + ((top & 7) - 1) + ((top & 8) ? 4 : 0)];
        K::store(c, V[ACC], e, C::load(c, src.c[e]));
    }
#pragma unroll 1
    for (int b = BW6_HARD_JSF_LEN - 2; b >= 0; b--) {
        K::mul(c, V[ACC], V[ACC], V[ACC], BW6_DENSE);
        const int d = BW6_HARD_JSF[b];
        if (d) K::mul(c, V[ACC], V[T0 + ((d & 7) - 1) + ((d & 8) ? 4 : 0)], V[ACC], BW6_DENSE);
    }
    const C v = C::load(c, V[ACC].c[e]);
    if (out) v.store(c, reinterpret_cast<uint32_t *>(out + (3 * (e & 1) + (e >> 1))));
    if (is_one) {
        const bool ok = e == 0 ? C::is_zero(c, C::sub(c, v, C::one(c))) : C::is_zero(c, v);
        if ((t & 31) == 0) S.flags[e] = ok ? 1 : 0;
        __syncthreads();
        if (t == 0) *is_one = (S.flags[0] & S.flags[1] & S.flags[2] & S.flags[3] & S.flags[4] & S.flags[5]) ? 1 : 0;
    }
}

}  // namespace b200
