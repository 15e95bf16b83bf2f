// BLS12-377 multi-pairing, warp-cooperative: ONE PAIRING PER WARP.
//
// Same function as pairing.cuh (Bls12::product_of_pairings, crates/bls-crypto/src/bls/
// signature.rs:149, public.rs:102; algorithm per SURVEY.md appendix A.3) -- pairing.cuh's
// one-thread-per-pair tower is kept as the cross-check and for the single Fq12 inversion.
// A pairing is ~12 k dependent-looking field products; done by one thread it is bound by product
// latency (0.94 us each).  Here the 32 lanes of a warp share the work of one pairing:
//
//   * Fq12 is held in the POWER BASIS Fq[w]/(w^12 + 5) (w^2 = v, v^3 = u, u^2 = -5, so w^12 = -5;
//     the tower coefficient c[I][J][K] of u^K v^J w^I sits at exponent e = 6K + 2J + I).  A product
//     is a length-12 negacyclic-style convolution: 144 independent Fq products.  24 lanes take
//     6 products each (lane = (output exponent e, half h)), halves are combined by one shuffle.
//     Frobenius is a coefficient-wise product with constants gamma_j^e, conjugation negates the
//     odd exponents.
//   * G2 line steps are lists of independent Fq2 products: 4 lanes per product (schoolbook),
//     up to 8 products per round, operands and results in per-warp shared-memory slots;
//     the linear glue (add / sub / small multiples) runs 2 lanes per statement.
//
// Values (Miller value and GT element) are identical to pairing.cuh's and to arkworks': same
// formulas, same line scaling, same final-exponentiation chain.
#pragma once
#include "pairing.cuh"

namespace b200 {

using FqImg = PFq::Mem;                              // 48-byte memory image of one Fq

struct alignas(16) Fq2Slot {
    FqImg c0, c1;
};

constexpr int W_SLOTS = 24;
struct alignas(16) WarpScratch {                     // per-warp shared memory (4032 bytes)
    FqImg f[3][12];                                  // three Fq12 values in the power basis (rotating roles)
    Fq2Slot s[W_SLOTS];                              // Fq2 registers of the line computation
};

// slot numbers
enum : int { S_RX = 0, S_RY, S_RZ, S_QX, S_QY, S_PX, S_PY, S_TWOINV, S_TWISTB, S_L0, S_L1, S_L2, S_T0 /* = 12 .. 23 temps */ };

B200_DEV PFq w_ld(const FqImg &m) { return PFq::load(m); }
// Every field product of the warp kernels goes through ONE out-of-line body: with the ~600-instruction product inlined
// in each helper (f12 product, squaring, the two Fq2 product rounds, Frobenius) the warps of an SM -- all in different
// phases of their Miller loops -- ran out of instruction cache (ncu: 3.7 warps per issue stalled on "no instruction").
// (SH = false: inlined, for the single-block latency kernels, where a call per product only adds to the chain.)
template <bool SH = true>
B200_DEV PFq w_mul(const PFq &a, const PFq &b) {
    if constexpr (SH) return PFq::mul_outline(a, b);
    else return a * b;
}
B200_DEV void w_st(FqImg &m, const PFq &v) { m = v.store(); }

// tower image index m = 6I + 2J + K  <->  power-basis exponent e = 6K + 2J + I
B200_DEV int tower_to_power(int m) { return 6 * (m & 1) + 2 * ((m % 6) >> 1) + (m / 6); }

// ---- Fq12 product in the power basis -----------------------------------------------------------
// O = A * B.  NJ = number of non-zero exponents of B handled per half (6 dense, 3 for a line);
// jpack holds the exponent list: nibble 2k + h is the k-th exponent of half h.
template <int NJ>
__device__ __noinline__ void w_f12_mul(const FqImg *A, const FqImg *B, FqImg *O, uint64_t jpack, int lane) {
    PFq part = PFq::zero();
    const int e = lane % 12, h = lane / 12;
    if (lane < 24) {
        PFq accP = PFq::zero(), accN = PFq::zero();
#pragma unroll 1
        for (int k = 0; k < NJ; k++) {
            int j = (int)((jpack >> (4 * (2 * k + h))) & 0xFull);
            int i = e - j;
            bool wrap = i < 0;
            i += wrap ? 12 : 0;
            PFq prod = w_mul(w_ld(A[i]), w_ld(B[j]));
            if (wrap) accN = accN + prod;
            else accP = accP + prod;
        }
        part = accP - (accN.dbl().dbl() + accN);     // w^12 = -5
    }
    PFq other = part.shfl(0xffffffffu, (lane + 12) & 31);
    __syncwarp();                                    // all reads of A / B done before O is written (O may be A's buffer of a later call)
    if (lane < 12) w_st(O[e], part + other);
    __syncwarp();
}
// O = A^2 in the power basis.  c_e = sum over UNORDERED pairs {i, j}, i + j = e (mod 12): 2 a_i a_j (i < j)
// or a_i^2 (i = j), times -5 when i + j >= 12.  Exponent e has 7 such pairs when even, 6 when odd: the
// m-th one is (m, e - m) for m <= e / 2 (no wrap) and (e + m - e / 2, 12 + e / 2 - m) beyond (wrap).
// Lane (e, h) takes m = h, h + 2, ...: at most 4 products against 6 in the general product.
__device__ __noinline__ void w_f12_sqr(const FqImg *A, FqImg *O, int lane) {
    PFq part = PFq::zero();
    const int e = lane % 12, h = lane / 12;
    if (lane < 24) {
        PFq accP = PFq::zero(), accN = PFq::zero();
        const int half = e >> 1, count = 7 - (e & 1);
#pragma unroll 1
        for (int k = 0; k < 4; k++) {
            const int m = 2 * k + h;
            if (m < count) {
                const bool wrap = m > half;
                const int i = wrap ? e + m - half : m;
                const int j = wrap ? 12 + half - m : e - m;
                PFq prod = w_mul(w_ld(A[i]), w_ld(A[j]));
                if (i != j) prod = prod.dbl();
                if (wrap) accN = accN + prod;
                else accP = accP + prod;
            }
        }
        part = accP - (accN.dbl().dbl() + accN);     // w^12 = -5
    }
    PFq other = part.shfl(0xffffffffu, (lane + 12) & 31);
    __syncwarp();
    if (lane < 12) w_st(O[e], part + other);
    __syncwarp();
}
constexpr uint64_t JPACK_DENSE = 0xBA9876543210ull;  // half 0: 0,2,4,6,8,10   half 1: 1,3,5,7,9,11
constexpr uint64_t JPACK_LINE = 0x976310ull;         // exponents {0,1,3,6,7,9}: half 0: 0,3,7  half 1: 1,6,9

// coefficient-wise: O[e] = A[e] * gamma_j^e (Frobenius), or negate odd e (conjugation)
B200_DEV void w_f12_frob(const FqImg *A, FqImg *O, int j, int lane) {
    if (lane < 12) w_st(O[lane], w_mul(w_ld(A[lane]), pfq_const(PAIRING_GAMMA[j - 1][lane])));
    __syncwarp();
}
B200_DEV void w_f12_conj(const FqImg *A, FqImg *O, int lane) {
    if (lane < 12) {
        PFq v = w_ld(A[lane]);
        w_st(O[lane], (lane & 1) ? v.neg() : v);
    }
    __syncwarp();
}
B200_DEV void w_f12_copy(const FqImg *A, FqImg *O, int lane) {
    if (lane < 12) O[lane] = A[lane];
    __syncwarp();
}
B200_DEV void w_f12_set_one(FqImg *O, int lane) {
    if (lane < 12) w_st(O[lane], lane == 0 ? PFq::one() : PFq::zero());
    __syncwarp();
}
// global tower image <-> shared power basis
B200_DEV void w_f12_load_global(const Fq12::Mem *g, FqImg *O, int lane) {
    if (lane < 12) O[tower_to_power(lane)] = reinterpret_cast<const FqImg *>(g)[lane];
    __syncwarp();
}
B200_DEV void w_f12_store_global(const FqImg *A, Fq12::Mem *g, int lane) {
    if (lane < 12) reinterpret_cast<FqImg *>(g)[lane] = A[tower_to_power(lane)];
    __syncwarp();
}

// ---- batches of Fq2 products / linear statements on the slots ----------------------------------
// product k (lanes 4k .. 4k+3): slot[dst_k] = slot[x_k] * slot[y_k]; packs hold one byte per product
template <bool SH = true>
static __device__ __noinline__ void w_f2_mul(Fq2Slot *s, int nprod, uint64_t xpack, uint64_t ypack, uint64_t dpack, int lane) {
    const int k = lane >> 2, t = lane & 3;
    const bool live = k < nprod;
    PFq prod = PFq::zero();
    if (live) {
        const Fq2Slot &X = s[(xpack >> (8 * k)) & 0xff], &Y = s[(ypack >> (8 * k)) & 0xff];
        PFq a = w_ld((t & 1) ? X.c1 : X.c0);                       // t: 0 (c0,c0) 1 (c1,c1) 2 (c0,c1) 3 (c1,c0)
        PFq b = w_ld((t == 1 || t == 2) ? Y.c1 : Y.c0);
        prod = w_mul<SH>(a, b);
    }
    PFq other = prod.shfl(0xffffffffu, lane ^ 1);
    __syncwarp();                                                  // operands read before any slot is overwritten
    if (live) {
        Fq2Slot &D = s[(dpack >> (8 * k)) & 0xff];
        if (t == 0) w_st(D.c0, prod - (other.dbl().dbl() + other));   // a0 b0 - 5 a1 b1
        if (t == 2) w_st(D.c1, prod + other);                         // a0 b1 + a1 b0
    }
    __syncwarp();
}
// statement q (lanes 2q, 2q+1; one Fq component each): 32-bit descriptor op | dst<<8 | a<<16 | b<<24
enum : uint32_t { L_ADD = 0, L_SUB = 1, L_DBL = 2, L_TRI = 3, L_NEG = 4, L_CPY = 5 };
__host__ __device__ constexpr uint32_t LST(uint32_t op, uint32_t dst, uint32_t a, uint32_t b = 0) { return op | (dst << 8) | (a << 16) | (b << 24); }
static __device__ __noinline__ void w_f2_lin(Fq2Slot *s, int nstmt, const uint32_t *desc, int lane) {
    const int q = lane >> 1, comp = lane & 1;
    const bool live = q < nstmt;
    PFq r = PFq::zero();
    uint32_t d = live ? desc[q] : 0u;
    if (live) {
        const Fq2Slot &A = s[(d >> 16) & 0xff], &B = s[(d >> 24) & 0xff];
        PFq a = w_ld(comp ? A.c1 : A.c0), b = w_ld(comp ? B.c1 : B.c0);
        switch (d & 0xff) {
        case L_ADD: r = a + b; break;
        case L_SUB: r = a - b; break;
        case L_DBL: r = a.dbl(); break;
        case L_TRI: r = a.dbl() + a; break;
        case L_NEG: r = a.neg(); break;
        default: r = a; break;
        }
    }
    __syncwarp();
    if (live) {
        Fq2Slot &D = s[(d >> 8) & 0xff];
        w_st(comp ? D.c1 : D.c0, r);
    }
    __syncwarp();
}
__host__ __device__ constexpr uint64_t PK(int a0 = 0, int a1 = 0, int a2 = 0, int a3 = 0, int a4 = 0, int a5 = 0, int a6 = 0, int a7 = 0) {
    return (uint64_t)a0 | ((uint64_t)a1 << 8) | ((uint64_t)a2 << 16) | ((uint64_t)a3 << 24) | ((uint64_t)a4 << 32) |
           ((uint64_t)a5 << 40) | ((uint64_t)a6 << 48) | ((uint64_t)a7 << 56);
}

// temps
enum : int { T12 = 12, T13, T14, T15, T16, T17, T18, T19, T20, T21, T22, T23 };

// statement tables (device constant memory)
__device__ const uint32_t DBL_A[] = {LST(L_ADD, T12, S_RY, S_RZ)};
__device__ const uint32_t DBL_B[] = {LST(L_TRI, T18, T15)};
__device__ const uint32_t DBL_C1[] = {LST(L_TRI, T20, T19), LST(L_ADD, T21, T14, T15), LST(L_SUB, T22, T19, T14), LST(L_TRI, T23, T17)};
__device__ const uint32_t DBL_C2[] = {LST(L_ADD, T12, T14, T20), LST(L_SUB, T18, T14, T20), LST(L_SUB, T16, T16, T21)};
__device__ const uint32_t DBL_C3[] = {LST(L_NEG, T21, T16)};
__device__ const uint32_t DBL_D1[] = {LST(L_TRI, T20, T20), LST(L_CPY, S_L2, T22)};
__device__ const uint32_t DBL_D2[] = {LST(L_SUB, S_RY, T12, T20)};

// doubling_step + line scaling by P (arkworks: coeffs (-h, 3j, i); c0 *= P.y, c1 *= P.x)
template <bool SH = true>
B200_DEV void w_doubling_step(Fq2Slot *s, int lane) {
    w_f2_lin(s, 1, DBL_A, lane);                                                    // T12 = ry + rz
    w_f2_mul<SH>(s, 5, PK(S_RX, S_RY, S_RZ, T12, S_RX), PK(S_RY, S_RY, S_RZ, T12, S_RX),
             PK(T13, T14, T15, T16, T17), lane);                                    // rx ry | b | c | (ry+rz)^2 | j
    w_f2_lin(s, 1, DBL_B, lane);                                                    // T18 = 3c
    w_f2_mul<SH>(s, 2, PK(S_TWISTB, T13), PK(T18, S_TWOINV), PK(T19, T13), lane);       // e = b' 3c | a = rx ry / 2
    w_f2_lin(s, 4, DBL_C1, lane);                                                   // f | b + c | i | 3j
    w_f2_lin(s, 3, DBL_C2, lane);                                                   // b + f | b - f | h
    w_f2_lin(s, 1, DBL_C3, lane);                                                   // -h
    w_f2_mul<SH>(s, 6, PK(T12, T19, T13, T14, T21, T23), PK(S_TWOINV, T19, T18, T16, S_PY, S_PX),
             PK(T12, T20, S_RX, S_RZ, S_L0, S_L1), lane);                           // g | e^2 | rx' | rz' | l0 | l1
    w_f2_mul<SH>(s, 1, PK(T12), PK(T12), PK(T12), lane);                                // g^2
    w_f2_lin(s, 2, DBL_D1, lane);                                                   // 3 e^2 | l2 = i
    w_f2_lin(s, 1, DBL_D2, lane);                                                   // ry' = g^2 - 3 e^2
}

__device__ const uint32_t ADD_A[] = {LST(L_SUB, T14, S_RY, T12), LST(L_SUB, T15, S_RX, T13)};
__device__ const uint32_t ADD_B[] = {LST(L_ADD, T21, T18, T19), LST(L_DBL, T22, T20)};
__device__ const uint32_t ADD_C[] = {LST(L_SUB, T23, T21, T22)};
__device__ const uint32_t ADD_D[] = {LST(L_SUB, T12, T20, T23)};
__device__ const uint32_t ADD_E[] = {LST(L_SUB, S_RY, T13, T16), LST(L_SUB, S_L2, T17, T19), LST(L_NEG, T21, T14)};

// addition_step + line scaling (coeffs (lambda, -theta, j))
template <bool SH = true>
B200_DEV void w_addition_step(Fq2Slot *s, int lane) {
    w_f2_mul<SH>(s, 2, PK(S_QY, S_QX), PK(S_RZ, S_RZ), PK(T12, T13), lane);             // qy rz | qx rz
    w_f2_lin(s, 2, ADD_A, lane);                                                    // theta = T14 | lambda = T15
    w_f2_mul<SH>(s, 2, PK(T14, T15), PK(T14, T15), PK(T16, T17), lane);                 // c = theta^2 | d = lambda^2
    w_f2_mul<SH>(s, 3, PK(T15, S_RZ, S_RX), PK(T17, T16, T17), PK(T18, T19, T20), lane);  // e | f | g
    w_f2_lin(s, 2, ADD_B, lane);                                                    // e + f | 2g
    w_f2_lin(s, 1, ADD_C, lane);                                                    // h = T23
    w_f2_lin(s, 1, ADD_D, lane);                                                    // g - h = T12
    w_f2_mul<SH>(s, 6, PK(T15, T14, T18, S_RZ, T14, T15), PK(T23, T12, S_RY, T18, S_QX, S_QY),
             PK(S_RX, T13, T16, S_RZ, T17, T19), lane);   // rx' | theta (g-h) | e ry | rz' | theta qx | lambda qy
    w_f2_lin(s, 3, ADD_E, lane);                                                    // ry' | l2 = j | -theta
    w_f2_mul<SH>(s, 2, PK(T15, T21), PK(S_PY, S_PX), PK(S_L0, S_L1), lane);             // l0 = lambda py | l1 = -theta px
}

// line (l0, l1, l2) -> sparse power-basis element: exponents 0,6 <- l0 ; 1,7 <- l1 ; 3,9 <- l2
B200_DEV void w_line_to_power(const Fq2Slot *s, FqImg *B, int lane) {
    if (lane < 6) {
        const int which = lane >> 1, comp = lane & 1;
        const int e = (which == 0 ? 0 : which == 1 ? 1 : 3) + 6 * comp;
        const Fq2Slot &L = s[S_L0 + which];
        B[e] = comp ? L.c1 : L.c0;
    }
    __syncwarp();
}

// ---- kernels ---------------------------------------------------------------------------------------
constexpr int W_WARPS = 4;                           // warps (pairings) per block

#ifdef B200_WITH_CROSSCHECKS
// warp w: Miller value of pair w -> out[w] (tower image); infinite members give one
__global__ void __launch_bounds__(32 * W_WARPS) k_w_miller_loop(const AffineMem<PFq> *__restrict__ g1,
                                                                const AffineMem<PFq2> *__restrict__ g2, uint32_t n,
                                                                Fq12::Mem *__restrict__ out) {
    __shared__ WarpScratch scratch[W_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t pair = blockIdx.x * W_WARPS + wib;
    if (pair >= n) return;                           // whole warps leave together
    WarpScratch &S = scratch[wib];
    Affine<PFq> p = Affine<PFq>::from_ark(ldg_mem(g1 + pair));
    Affine<PFq2> q = Affine<PFq2>::from_ark(ldg_mem(g2 + pair));
    int fa = 0;                                      // S.f[fa] holds f
    w_f12_set_one(S.f[0], lane);
    if (!p.is_inf() && !q.is_inf()) {                // warp-uniform (every lane loaded the same pair)
        if (lane == 0) {
            S.s[S_RX] = {q.x.c0.store(), q.x.c1.store()};
            S.s[S_RY] = {q.y.c0.store(), q.y.c1.store()};
            S.s[S_RZ] = {PFq::one().store(), PFq::zero().store()};
            S.s[S_QX] = S.s[S_RX];
            S.s[S_QY] = S.s[S_RY];
            S.s[S_PX] = {p.x.store(), PFq::zero().store()};
            S.s[S_PY] = {p.y.store(), PFq::zero().store()};
            S.s[S_TWOINV] = {pfq_const(PAIRING_TWO_INV).store(), PFq::zero().store()};
            S.s[S_TWISTB] = {PFq::zero().store(), pfq_const(PAIRING_TWIST_B_C1).store()};
        }
        if (lane < 12) w_st(S.f[2][lane], PFq::zero());      // line element: untouched exponents stay zero
        __syncwarp();
#pragma unroll 1
        for (int b = 62; b >= 0; b--) {
            w_f12_mul<6>(S.f[fa], S.f[fa], S.f[fa ^ 1], JPACK_DENSE, lane);          // f^2
            fa ^= 1;
            w_doubling_step(S.s, lane);
            w_line_to_power(S.s, S.f[2], lane);
            w_f12_mul<3>(S.f[fa], S.f[2], S.f[fa ^ 1], JPACK_LINE, lane);            // f * line
            fa ^= 1;
            if ((PAIRING_X >> b) & 1ull) {
                w_addition_step(S.s, lane);
                w_line_to_power(S.s, S.f[2], lane);
                w_f12_mul<3>(S.f[fa], S.f[2], S.f[fa ^ 1], JPACK_LINE, lane);
                fa ^= 1;
            }
        }
    }
    w_f12_store_global(S.f[fa], out + pair, lane);
}
#endif

// ---- two pairs per warp, one shared f ----------------------------------------------------------------
// prod_i f_i can be accumulated in ONE Miller variable: f <- f^2 * line_0 * line_1 per bit, so the
// dense squaring (the largest item of an iteration) is paid once per two pairs, and the two G2 line
// steps run side by side in the same product rounds (<= 4 Fq2 products per pair and round, 4 lanes
// each).  Slot sets s0 / s1 hold the two pairs' line state; `two` is false when the warp has only
// one finite pair.
template <bool SH = true>
static __device__ __noinline__ void w_f2_mul2(Fq2Slot *s0, Fq2Slot *s1, bool two, int nprod, uint32_t xpack,
                                              uint32_t ypack, uint32_t dpack, int lane) {
    const int k = lane >> 2, t = lane & 3;
    const int pr = k >= nprod ? 1 : 0, kk = k - pr * nprod;
    const bool live = kk < nprod && (pr == 0 || two);
    Fq2Slot *s = pr ? s1 : s0;
    PFq prod = PFq::zero();
    if (live) {
        const Fq2Slot &X = s[(xpack >> (8 * kk)) & 0xff], &Y = s[(ypack >> (8 * kk)) & 0xff];
        PFq a = w_ld((t & 1) ? X.c1 : X.c0);                       // t: 0 (c0,c0) 1 (c1,c1) 2 (c0,c1) 3 (c1,c0)
        PFq b = w_ld((t == 1 || t == 2) ? Y.c1 : Y.c0);
        prod = w_mul<SH>(a, b);
    }
    PFq other = prod.shfl(0xffffffffu, lane ^ 1);
    __syncwarp();
    if (live) {
        Fq2Slot &D = s[(dpack >> (8 * kk)) & 0xff];
        if (t == 0) w_st(D.c0, prod - (other.dbl().dbl() + other));
        if (t == 2) w_st(D.c1, prod + other);
    }
    __syncwarp();
}
static __device__ __noinline__ void w_f2_lin2(Fq2Slot *s0, Fq2Slot *s1, bool two, int nstmt, const uint32_t *desc, int lane) {
    const int q = lane >> 1, comp = lane & 1;
    const int pr = q >= nstmt ? 1 : 0, qq = q - pr * nstmt;
    const bool live = qq < nstmt && (pr == 0 || two);
    Fq2Slot *s = pr ? s1 : s0;
    PFq r = PFq::zero();
    uint32_t d = live ? desc[qq] : 0u;
    if (live) {
        const Fq2Slot &A = s[(d >> 16) & 0xff], &B = s[(d >> 24) & 0xff];
        PFq a = w_ld(comp ? A.c1 : A.c0), b = w_ld(comp ? B.c1 : B.c0);
        switch (d & 0xff) {
        case L_ADD: r = a + b; break;
        case L_SUB: r = a - b; break;
        case L_DBL: r = a.dbl(); break;
        case L_TRI: r = a.dbl() + a; break;
        case L_NEG: r = a.neg(); break;
        default: r = a; break;
        }
    }
    __syncwarp();
    if (live) {
        Fq2Slot &D = s[(d >> 8) & 0xff];
        w_st(comp ? D.c1 : D.c0, r);
    }
    __syncwarp();
}
__host__ __device__ constexpr uint32_t PK4(int a0 = 0, int a1 = 0, int a2 = 0, int a3 = 0) {
    return (uint32_t)a0 | ((uint32_t)a1 << 8) | ((uint32_t)a2 << 16) | ((uint32_t)a3 << 24);
}
// w_doubling_step for both slot sets at once (same statements, product rounds regrouped to <= 4)
template <bool SH = true>
B200_DEV void w_doubling_step2(Fq2Slot *s0, Fq2Slot *s1, bool two, int lane) {
    w_f2_lin2(s0, s1, two, 1, DBL_A, lane);                                                   // T12 = ry + rz
    w_f2_mul2<SH>(s0, s1, two, 4, PK4(S_RX, S_RY, S_RZ, T12), PK4(S_RY, S_RY, S_RZ, T12), PK4(T13, T14, T15, T16), lane);
    w_f2_lin2(s0, s1, two, 1, DBL_B, lane);                                                   // T18 = 3c
    w_f2_mul2<SH>(s0, s1, two, 3, PK4(S_TWISTB, T13, S_RX), PK4(T18, S_TWOINV, S_RX), PK4(T19, T13, T17), lane);   // e | a | j
    w_f2_lin2(s0, s1, two, 4, DBL_C1, lane);
    w_f2_lin2(s0, s1, two, 3, DBL_C2, lane);
    w_f2_lin2(s0, s1, two, 1, DBL_C3, lane);
    w_f2_mul2<SH>(s0, s1, two, 4, PK4(T12, T19, T13, T14), PK4(S_TWOINV, T19, T18, T16), PK4(T12, T20, S_RX, S_RZ), lane);   // g | e^2 | rx' | rz'
    w_f2_mul2<SH>(s0, s1, two, 3, PK4(T12, T21, T23), PK4(T12, S_PY, S_PX), PK4(T12, S_L0, S_L1), lane);                     // g^2 | l0 | l1
    w_f2_lin2(s0, s1, two, 2, DBL_D1, lane);
    w_f2_lin2(s0, s1, two, 1, DBL_D2, lane);
}

struct alignas(16) WarpScratch2 {                    // per-warp shared memory: shared f, two line states
    FqImg f[3][12];
    Fq2Slot s[2][W_SLOTS];
};

B200_DEV void w_init_line_state(Fq2Slot *s, const Affine<PFq> &p, const Affine<PFq2> &q) {
    s[S_RX] = {q.x.c0.store(), q.x.c1.store()};
    s[S_RY] = {q.y.c0.store(), q.y.c1.store()};
    s[S_RZ] = {PFq::one().store(), PFq::zero().store()};
    s[S_QX] = s[S_RX];
    s[S_QY] = s[S_RY];
    s[S_PX] = {p.x.store(), PFq::zero().store()};
    s[S_PY] = {p.y.store(), PFq::zero().store()};
    s[S_TWOINV] = {pfq_const(PAIRING_TWO_INV).store(), PFq::zero().store()};
    s[S_TWISTB] = {PFq::zero().store(), pfq_const(PAIRING_TWIST_B_C1).store()};
}

// warp w: out[w] = Miller value of pair 2w  *  Miller value of pair 2w + 1 (tower image);
// pairs with an infinite member (and the missing partner of an odd n) contribute one
// WW warps per block: the warps are independent, so the block size only decides how evenly 2049 warps (config 2) spread
// over 148 SMs -- 513 blocks of 4 leave some SMs with 4 blocks and some with 3 (the kernel ends with the fullest SM),
// 1025 blocks of 2 put 13 or 14 warps on every SM
template <int WW>
__global__ void __launch_bounds__(32 * WW) k_w2_miller_loop(const AffineMem<PFq> *__restrict__ g1,
                                                            const AffineMem<PFq2> *__restrict__ g2, uint32_t n,
                                                            Fq12::Mem *__restrict__ out) {
    __shared__ WarpScratch2 scratch[WW];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t w = blockIdx.x * WW + wib;
    if (2 * w >= n) return;                          // whole warps leave together
    WarpScratch2 &S = scratch[wib];
    int live = 0;                                    // finite pairs placed in slot sets 0 .. live-1
    for (uint32_t pair = 2 * w; pair < min(n, 2 * w + 2); pair++) {
        Affine<PFq> p = Affine<PFq>::from_ark(ldg_mem(g1 + pair));
        Affine<PFq2> q = Affine<PFq2>::from_ark(ldg_mem(g2 + pair));
        if (p.is_inf() || q.is_inf()) continue;      // warp-uniform
        if (lane == 0) w_init_line_state(S.s[live], p, q);
        live++;
    }
    int fa = 0;
    w_f12_set_one(S.f[0], lane);
    if (live) {
        const bool two = live == 2;
        if (lane < 12) w_st(S.f[2][lane], PFq::zero());      // line element: untouched exponents stay zero
        __syncwarp();
#pragma unroll 1
        for (int b = 62; b >= 0; b--) {
            if (b != 62) {                                   // f is still one before the first lines
                w_f12_sqr(S.f[fa], S.f[fa ^ 1], lane);                               // f^2
                fa ^= 1;
            }
            w_doubling_step2(S.s[0], S.s[1], two, lane);
            for (int k = 0; k < live; k++) {
                w_line_to_power(S.s[k], S.f[2], lane);
                w_f12_mul<3>(S.f[fa], S.f[2], S.f[fa ^ 1], JPACK_LINE, lane);        // f * line
                fa ^= 1;
            }
            if ((PAIRING_X >> b) & 1ull) {
                for (int k = 0; k < live; k++) {
                    w_addition_step(S.s[k], lane);
                    w_line_to_power(S.s[k], S.f[2], lane);
                    w_f12_mul<3>(S.f[fa], S.f[2], S.f[fa ^ 1], JPACK_LINE, lane);
                    fa ^= 1;
                }
            }
        }
    }
    w_f12_store_global(S.f[fa], out + w, lane);
}

// ---- two pairs per BLOCK: the latency form of the same loop ----------------------------------------------------------
// k_w2_miller_loop gives a check (two pairs) one warp: every Fq12 operation is 3-4 field products deep per lane and the
// G2 line steps wait behind them -- 27 us per bit, 1.7 ms for the 2-pair check of verify_signature.  Here a block of four
// warps shares it: warp 3 walks the G2 points (w_doubling_step2 / w_addition_step, unchanged) ONE BIT AHEAD and leaves the
// sparse line elements in a double-buffered slot, warps 0-2 (96 threads) keep the Miller variable -- f^2 is 78 products,
// a sparse line product 72, each ONE field product deep, eight lanes per output coefficient folded by shuffles -- and the
// two groups meet once per bit.  The pipe work is the same as the warp kernel's, spread over four times the warps.  Values are identical: the same field operations on the same operands, multiplied in another order.
constexpr int B2_F_THREADS = 96, B2_THREADS = 128;
struct alignas(16) Block2Scratch {
    FqImg f[12];                                     // the Miller variable, power basis
    FqImg line[2][4][12];                            // [bit parity][pair 0 / pair 1 doubling, pair 0 / pair 1 addition][exponent]
    Fq2Slot s[2][W_SLOTS];                           // line state of the two pairs
};
B200_DEV void b2_f_bar() { asm volatile("bar.sync 1, 96;" ::: "memory"); }
__device__ const uint8_t B2_LINE_EXP[6] = {0, 1, 3, 6, 7, 9};
// Eight lanes per output exponent e (t = 8 e + m): lane m computes ONE term of c_e -- wrapped terms (i + j >= 12) already
// times -5, w^12 = -5 -- and three xor-shuffle steps add the eight up; no product ever goes through shared memory.
B200_DEV PFq b2_fold8(PFq term, bool wrap, int lane) {
    if (wrap) term = (term.dbl().dbl() + term).neg();
#pragma unroll 1
    for (int o = 4; o >= 1; o >>= 1) term = term + term.shfl(0xffffffffu, lane ^ o);
    return term;
}
// f <- f^2: the unordered pairs {i, j}, i + j = e (mod 12): 7 for even e, 6 for odd (enumeration of w_f12_sqr)
B200_DEV void b2_f12_sqr(FqImg *F, int t) {
    const int e = t >> 3, m = t & 7, half = e >> 1, count = 7 - (e & 1);
    PFq term = PFq::zero();
    bool wrap = false;
    if (m < count) {
        wrap = m > half;
        const int i = wrap ? e + m - half : m, j = wrap ? 12 + half - m : e - m;
        term = w_ld(F[i]) * w_ld(F[j]);
        if (i != j) term = term.dbl();
    }
    term = b2_fold8(term, wrap, t & 31);
    b2_f_bar();                                      // every read of f is done
    if (m == 0) w_st(F[e], term);
    b2_f_bar();
}
// f <- f * line (exponents 0, 1, 3, 6, 7, 9 of L)
B200_DEV void b2_f12_mul_line(FqImg *F, const FqImg *L, int t) {
    const int e = t >> 3, m = t & 7;
    PFq term = PFq::zero();
    bool wrap = false;
    if (m < 6) {
        const int j = B2_LINE_EXP[m];
        int i = e - j;
        wrap = i < 0;
        i += wrap ? 12 : 0;
        term = w_ld(F[i]) * w_ld(L[j]);
    }
    term = b2_fold8(term, wrap, t & 31);
    b2_f_bar();
    if (m == 0) w_st(F[e], term);
    b2_f_bar();
}

// block b: out[b] = Miller value of pair 2b  *  Miller value of pair 2b + 1 (tower image), as k_w2_miller_loop's warp b
__global__ void __launch_bounds__(B2_THREADS) k_b2_miller_loop(const AffineMem<PFq> *__restrict__ g1, const AffineMem<PFq2> *__restrict__ g2,
                                                               uint32_t n, Fq12::Mem *__restrict__ out) {
    __shared__ Block2Scratch S;
    const int t = threadIdx.x, lane = t & 31;
    const bool line_warp = t >= B2_F_THREADS;
    const uint32_t w = blockIdx.x;
    if (2 * w >= n) return;
    int live = 0;                                    // finite pairs placed in slot sets 0 .. live-1 (block-uniform)
    for (uint32_t pair = 2 * w; pair < min(n, 2 * w + 2); pair++) {
        Affine<PFq> p = Affine<PFq>::from_ark(ldg_mem(g1 + pair));
        Affine<PFq2> q = Affine<PFq2>::from_ark(ldg_mem(g2 + pair));
        if (p.is_inf() || q.is_inf()) continue;
        if (t == B2_F_THREADS) w_init_line_state(S.s[live], p, q);
        live++;
    }
    if (t < 12) w_st(S.f[t], t == 0 ? PFq::one() : PFq::zero());
    __syncthreads();
    if (live) {
        const bool two = live == 2;
        // lines of bit b: doubling lines in slots 0 / 1, addition lines (bit set) in slots 2 / 3
        auto lines_of_bit = [&](int b, int par) {
            w_doubling_step2<false>(S.s[0], S.s[1], two, lane);
            for (int k = 0; k < live; k++) w_line_to_power(S.s[k], S.line[par][k], lane);
            if ((PAIRING_X >> b) & 1ull) {
                for (int k = 0; k < live; k++) {
                    w_addition_step<false>(S.s[k], lane);
                    w_line_to_power(S.s[k], S.line[par][2 + k], lane);
                }
            }
        };
        if (line_warp) lines_of_bit(62, 0);
        __syncthreads();
        int par = 0;
#pragma unroll 1
        for (int b = 62; b >= 0; b--) {
            if (line_warp) {
                if (b > 0) lines_of_bit(b - 1, par ^ 1);
            } else {
                if (b != 62) b2_f12_sqr(S.f, t);                 // f is still one before the first lines
                for (int k = 0; k < live; k++) b2_f12_mul_line(S.f, S.line[par][k], t);
                if ((PAIRING_X >> b) & 1ull)
                    for (int k = 0; k < live; k++) b2_f12_mul_line(S.f, S.line[par][2 + k], t);
            }
            __syncthreads();
            par ^= 1;
        }
    }
    if (t < 12) reinterpret_cast<FqImg *>(out + w)[t] = S.f[tower_to_power(t)];
}

// warp w of the grid: vals[w] = prod_{i = w (mod stride)} vals[i]   (strided in-place partial products)
__global__ void __launch_bounds__(32 * W_WARPS) k_w_fq12_strided_product(Fq12::Mem *__restrict__ vals, uint32_t n,
                                                                         uint32_t stride) {
    __shared__ WarpScratch scratch[W_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t w = blockIdx.x * W_WARPS + wib;
    if (w >= stride || w >= n) return;
    WarpScratch &S = scratch[wib];
    int fa = 0;
    w_f12_load_global(vals + w, S.f[0], lane);
    for (uint32_t i = w + stride; i < n; i += stride) {
        w_f12_load_global(vals + i, S.f[2], lane);
        w_f12_mul<6>(S.f[fa], S.f[2], S.f[fa ^ 1], JPACK_DENSE, lane);
        fa ^= 1;
    }
    w_f12_store_global(S.f[fa], vals + w, lane);
}

// one warp: r = f^x (x = PAIRING_X), power basis; uses S.f[0..2] as rotating buffers, result index returned
B200_DEV int w_exp_by_x(WarpScratch &S, FqImg *base /* separate 12-entry buffer holding f */, int lane) {
    int cur = 0;
    w_f12_copy(base, S.f[0], lane);
#pragma unroll 1
    for (int b = 62; b >= 0; b--) {
        w_f12_mul<6>(S.f[cur], S.f[cur], S.f[cur ^ 1], JPACK_DENSE, lane);
        cur ^= 1;
        if ((PAIRING_X >> b) & 1ull) {
            w_f12_mul<6>(S.f[cur], base, S.f[cur ^ 1], JPACK_DENSE, lane);
            cur ^= 1;
        }
    }
    return cur;
}

#ifdef B200_WITH_CROSSCHECKS
// one warp: Bls12::final_exponentiation (2016/130 table 1 chain) of vals[0]
__global__ void __launch_bounds__(32) k_w_final_exp(const Fq12::Mem *__restrict__ in, Fq12::Mem *__restrict__ out,
                                                    int *__restrict__ is_one, Fq12::Mem *__restrict__ tmp /* 1 image */) {
    __shared__ WarpScratch S;
    __shared__ FqImg V[8][12];                       // named values of the chain, power basis
    const int lane = threadIdx.x & 31;
    enum { R = 0, Y0, Y1, Y2, Y3, Y4, Y5, X = 7 };
    // f^(p^6 - 1) = conj(f) * f^-1 : the one inversion, on lane 0 with the per-thread tower
    if (lane == 0) {
        Fq12 f = f12_load(in[0]);
        launder(f);
        Fq12 fi = f12_inv(f);
        launder(fi);
        tmp[0] = f12_store(fi);
    }
    __syncwarp();
    __threadfence_block();
    w_f12_load_global(in, V[X], lane);               // f
    w_f12_conj(V[X], V[Y0], lane);                   // conj(f)
    w_f12_load_global(tmp, V[Y1], lane);             // f^-1
    w_f12_mul<6>(V[Y0], V[Y1], V[R], JPACK_DENSE, lane);          // r = conj(f) / f
    w_f12_frob(V[R], V[Y0], 2, lane);
    w_f12_mul<6>(V[Y0], V[R], V[Y1], JPACK_DENSE, lane);
    w_f12_copy(V[Y1], V[R], lane);                   // r = frob^2(r) * r          (easy part)
    // hard part
    w_f12_mul<6>(V[R], V[R], V[X], JPACK_DENSE, lane);
    w_f12_conj(V[X], V[Y0], lane);                   // y0 = conj(r^2)
    int c = w_exp_by_x(S, V[R], lane);
    w_f12_copy(S.f[c], V[Y5], lane);                 // y5 = r^x
    w_f12_mul<6>(V[Y5], V[Y5], V[Y1], JPACK_DENSE, lane);         // y1 = y5^2
    w_f12_mul<6>(V[Y0], V[Y5], V[Y3], JPACK_DENSE, lane);         // y3 = y0 y5
    c = w_exp_by_x(S, V[Y3], lane);
    w_f12_copy(S.f[c], V[Y0], lane);                 // y0 = y3^x
    c = w_exp_by_x(S, V[Y0], lane);
    w_f12_copy(S.f[c], V[Y2], lane);                 // y2 = y0^x
    c = w_exp_by_x(S, V[Y2], lane);
    w_f12_mul<6>(S.f[c], V[Y1], V[Y4], JPACK_DENSE, lane);        // y4 = y2^x y1
    c = w_exp_by_x(S, V[Y4], lane);
    w_f12_copy(S.f[c], V[Y1], lane);                 // y1 = y4^x
    w_f12_conj(V[Y3], V[X], lane);
    w_f12_copy(V[X], V[Y3], lane);                   // y3 = conj(y3)
    w_f12_mul<6>(V[Y1], V[Y3], V[X], JPACK_DENSE, lane);
    w_f12_mul<6>(V[X], V[R], V[Y1], JPACK_DENSE, lane);           // y1 = y1 y3 r
    w_f12_conj(V[R], V[Y3], lane);                   // y3 = conj(r)
    w_f12_mul<6>(V[Y0], V[R], V[X], JPACK_DENSE, lane);
    w_f12_frob(V[X], V[Y0], 3, lane);                // y0 = frob^3(y0 r)
    w_f12_mul<6>(V[Y4], V[Y3], V[X], JPACK_DENSE, lane);
    w_f12_frob(V[X], V[Y4], 1, lane);                // y4 = frob(y4 y3)
    w_f12_mul<6>(V[Y5], V[Y2], V[X], JPACK_DENSE, lane);
    w_f12_frob(V[X], V[Y5], 2, lane);                // y5 = frob^2(y5 y2)
    w_f12_mul<6>(V[Y5], V[Y0], V[X], JPACK_DENSE, lane);
    w_f12_mul<6>(V[X], V[Y4], V[Y2], JPACK_DENSE, lane);
    w_f12_mul<6>(V[Y2], V[Y1], V[X], JPACK_DENSE, lane);          // result
    if (out) w_f12_store_global(V[X], out, lane);
    if (is_one) {
        bool ok = true;
        if (lane < 12) {
            PFq v = w_ld(V[X][lane]);
            ok = lane == 0 ? (v == PFq::one()) : v.is_zero();
        }
        unsigned all = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) *is_one = all == 0xffffffffu ? 1 : 0;
    }
}
#endif

// ---- block-cooperative final exponentiation ---------------------------------------------------------
// The final exponentiation is ONE Fq12 chain of ~350 dependent products: with one warp each product
// is 6 sequential field products per lane.  Here a block of 160 threads gives every one of the 144
// coefficient products a[i] * b[j] its own thread (one field-product latency per Fq12 product), the
// 12 output coefficients are then summed from shared memory by 24 threads (positive / wrapped halves).
constexpr int B_THREADS = 192;
struct alignas(16) BlockScratch {
    FqImg V[9][12];                                  // named values of the chain, power basis
};

// O = A * B; all B_THREADS threads call; O may alias A and / or B.  Sixteen lanes per output exponent e (t = 16 e + i,
// twelve of them busy): lane i computes a_i b_((e - i) mod 12), wrapped terms times -5 (w^12 = -5), four xor-shuffle steps
// add the group up -- one field product and four additions deep, nothing staged through shared memory.
B200_DEV void b_f12_mul(const FqImg *A, const FqImg *B, FqImg *O, int t) {
    const int e = t >> 4, i = t & 15;
    PFq term = PFq::zero();
    if (i < 12) {
        int j = e - i;
        const bool wrap = j < 0;
        j += wrap ? 12 : 0;
        term = w_ld(A[i]) * w_ld(B[j]);
        if (wrap) term = (term.dbl().dbl() + term).neg();
    }
#pragma unroll 1
    for (int o = 8; o >= 1; o >>= 1) term = term + term.shfl(0xffffffffu, (t & 31) ^ o);
    __syncthreads();                                 // every read of A / B is done
    if (i == 0) w_st(O[e], term);
    __syncthreads();
}
B200_DEV void b_f12_frob(const FqImg *A, FqImg *O, int j, int t) {
    if (t < 12) w_st(O[t], w_ld(A[t]) * pfq_const(PAIRING_GAMMA[j - 1][t]));
    __syncthreads();
}
B200_DEV void b_f12_conj(const FqImg *A, FqImg *O, int t) {
    if (t < 12) {
        PFq v = w_ld(A[t]);
        w_st(O[t], (t & 1) ? v.neg() : v);
    }
    __syncthreads();
}
B200_DEV void b_f12_copy(const FqImg *A, FqImg *O, int t) {
    if (t < 12) O[t] = A[t];
    __syncthreads();
}
// O = base^x (x = PAIRING_X), O != base
B200_DEV void b_exp_by_x(const FqImg *base, FqImg *O, int t) {
    b_f12_copy(base, O, t);
#pragma unroll 1
    for (int b = 62; b >= 0; b--) {
        b_f12_mul(O, O, O, t);
        if ((PAIRING_X >> b) & 1ull) b_f12_mul(O, base, O, t);
    }
}

// one block: Bls12::final_exponentiation (2016/130 table 1 chain) of in[0]
__global__ void __launch_bounds__(B_THREADS) k_b_final_exp(const Fq12::Mem *__restrict__ in, Fq12::Mem *__restrict__ out,
                                                           int *__restrict__ is_one) {
    __shared__ BlockScratch S;
    __shared__ Fq12::Mem inv_img;
    const int t = threadIdx.x;
    // one block per value: a grid of `count` blocks finishes `count` independent products of pairings (the per-batch
    // checks of batch_verify_strict); the single-value callers launch one block
    in += blockIdx.x;
    if (out) out += blockIdx.x;
    if (is_one) is_one += blockIdx.x;
    enum { R = 0, Y0, Y1, Y2, Y3, Y4, Y5, X, T };
    FqImg(*V)[12] = S.V;
    // f^(p^6 - 1) = conj(f) * f^-1 : the one inversion, on thread 0 with the per-thread tower
    if (t == 0) {
        Fq12 f = f12_load(in[0]);
        launder(f);
        Fq12 fi = f12_inv(f);
        launder(fi);
        inv_img = f12_store(fi);
    }
    if (t < 12) V[X][tower_to_power(t)] = reinterpret_cast<const FqImg *>(in)[t];          // f
    __syncthreads();
    if (t < 12) V[Y1][tower_to_power(t)] = reinterpret_cast<const FqImg *>(&inv_img)[t];   // f^-1
    __syncthreads();
    b_f12_conj(V[X], V[Y0], t);                      // conj(f)
    b_f12_mul(V[Y0], V[Y1], V[R], t);             // r = conj(f) / f
    b_f12_frob(V[R], V[Y0], 2, t);
    b_f12_mul(V[Y0], V[R], V[R], t);              // r = frob^2(r) * r          (easy part)
    // hard part
    b_f12_mul(V[R], V[R], V[X], t);
    b_f12_conj(V[X], V[Y0], t);                      // y0 = conj(r^2)
    b_exp_by_x(V[R], V[Y5], t);                   // y5 = r^x
    b_f12_mul(V[Y5], V[Y5], V[Y1], t);            // y1 = y5^2
    b_f12_mul(V[Y0], V[Y5], V[Y3], t);            // y3 = y0 y5
    b_exp_by_x(V[Y3], V[Y0], t);                  // y0 = y3^x
    b_exp_by_x(V[Y0], V[Y2], t);                  // y2 = y0^x
    b_exp_by_x(V[Y2], V[Y4], t);
    b_f12_mul(V[Y4], V[Y1], V[Y4], t);            // y4 = y2^x y1
    b_exp_by_x(V[Y4], V[Y1], t);                  // y1 = y4^x
    b_f12_conj(V[Y3], V[Y3], t);                     // y3 = conj(y3)
    b_f12_mul(V[Y1], V[Y3], V[Y1], t);
    b_f12_mul(V[Y1], V[R], V[Y1], t);             // y1 = y1 y3 r
    b_f12_conj(V[R], V[Y3], t);                      // y3 = conj(r)
    b_f12_mul(V[Y0], V[R], V[X], t);
    b_f12_frob(V[X], V[Y0], 3, t);                   // y0 = frob^3(y0 r)
    b_f12_mul(V[Y4], V[Y3], V[X], t);
    b_f12_frob(V[X], V[Y4], 1, t);                   // y4 = frob(y4 y3)
    b_f12_mul(V[Y5], V[Y2], V[X], t);
    b_f12_frob(V[X], V[Y5], 2, t);                   // y5 = frob^2(y5 y2)
    b_f12_mul(V[Y5], V[Y0], V[X], t);
    b_f12_mul(V[X], V[Y4], V[X], t);
    b_f12_mul(V[X], V[Y1], V[X], t);              // result
    if (out && t < 12) reinterpret_cast<FqImg *>(out)[t] = V[X][tower_to_power(t)];
    if (is_one && t < 32) {
        bool ok = true;
        if (t < 12) {
            PFq v = w_ld(V[X][t]);
            ok = t == 0 ? (v == PFq::one()) : v.is_zero();
        }
        unsigned all = __ballot_sync(0xffffffffu, ok);
        if (t == 0) *is_one = all == 0xffffffffu ? 1 : 0;
    }
}

}  // namespace b200
