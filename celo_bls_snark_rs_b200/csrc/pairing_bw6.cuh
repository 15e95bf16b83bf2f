// BW6-761 product of pairings, block-cooperative.
//
// Replaces BW6_761::product_of_pairings as reached from ark-groth16's verify_proof at
// crates/epoch-snark/src/api/verifier.rs:35 (the check e(A, B) e(g_ic, -gamma) e(C, -delta) == e(alpha, beta)).
// The pairing is the optimal ate pairing upstream computes (eprint 2020/351 algorithm 5):
//
//     e(P, Q) = ( f_{u+1,Q}(P) * f_{u^3-u^2-u,Q}(P)^q ) ^ ((q^6 - 1) / r)
//
// with this layout (the test-side restatement of the same algorithm is named in DESIGN.md section 4.5):
//   * F_q^6 in the power basis F_q[w] / (w^6 + 4): a product is 36 independent 761-bit Montgomery
//     products (one per thread) and six short sums; Frobenius is a coefficient-wise product with
//     gamma^k, conjugation (the q^3 power) negates the odd coefficients.
//   * G2 on the M-twist y^2 = x^3 + 4 over F_q; the running point T is Jacobian, a line is the three
//     coefficients of (1, w^2, w^3) after scaling by elements of F_q and by w^3 (both die in the final
//     exponentiation).
//
// This is a SINGLE-INSTANCE, latency-bound path (4 pairs per Groth16 verification): the unit of time is
// one dependent 761-bit product (1152 wide multiply-adds issued by one warp, ~2.5 us).  The schedule is
// therefore written in ROUNDS: in every round each thread of the block performs at most one product, all
// through the same out-of-line body, then a few threads do the linear glue.
//   * k_bw6_miller: one block per (pair, sub-loop).  The point chain T (3 rounds per doubling, 5 per
//     addition) is the critical path; the Miller variable f (square, line product, [line product]) lags
//     one step behind and rides in the same rounds on other threads.
//   * k_bw6_final_exp: one block; product of the Miller values, easy part with one base-field inversion
//     (norm to F_q^3, norm to F_q), hard part as f^R0 (f^q)^R1 with R0 + q R1 = c (q^2 - q + 1) / r: a joint
//     square-and-multiply over 575 signed digit pairs (244 products) instead of 1144 bits (563 products).
#pragma once
#include "ec.cuh"
#include "fp.cuh"
#include "pairing_bw6_params_gen.cuh"

namespace b200 {

using BFq = Fq761;
using BImg = BFq::Mem;                               // 96-byte memory image

constexpr int BW6_THREADS = 64;
constexpr int BW6_TBASE = 36;                        // threads 36.. take the products of the point chain

B200_DEV BFq b_ld(const BImg &m) { return BFq::load(m); }
B200_DEV void b_st(BImg &m, const BFq &v) { m = v.store(); }
B200_DEV BFq b_const(const uint32_t (&w)[24]) {
    BFq r;
#pragma unroll
    for (int i = 0; i < 24; i++) r.l[i] = w[i];
    return r;
}

// one round: every thread with a destination multiplies; the barrier publishes the products
B200_DEV void bw6_round(const BImg *a, const BImg *b, BImg *o) {
    if (o) b_st(*o, b_ld(*a) * b_ld(*b));
    __syncthreads();
}

// F[e] = sum_{i + j = e} P[6 i + j] - 4 sum_{i + j = e + 6} P[6 i + j], over the j in jmask; threads 0..5
B200_DEV void bw6_f6_reduce(const BImg *P, BImg *F, uint32_t jmask, int t) {
    if (t < 6) {
        BFq pos = BFq::zero(), neg = BFq::zero();
#pragma unroll 1
        for (int j = 0; j < 6; j++) {
            if (!((jmask >> j) & 1u)) continue;
            int i = t - j;
            const bool wrap = i < 0;
            i += wrap ? 6 : 0;
            BFq v = b_ld(P[6 * i + j]);
            if (wrap) neg = neg + v;
            else pos = pos + v;
        }
        b_st(F[t], pos - neg.dbl().dbl());           // w^6 = -4
    }
}

// operands of thread t for the product A * B restricted to the exponents of B in jmask (dense: 0x3f, line: 0x0d)
B200_DEV void bw6_f6_operands(const BImg *A, const BImg *B, BImg *P, uint32_t jmask, int t, const BImg *&a, const BImg *&b,
                              BImg *&o) {
    if (t < 36) {
        const int i = t / 6, j = t % 6;
        if ((jmask >> j) & 1u) {
            a = &A[i];
            b = &B[j];
            o = &P[6 * i + j];
        }
    }
}

// O = A * B, whole block (O may alias A or B)
B200_DEV void bw6_f6_mul(const BImg *A, const BImg *B, BImg *O, BImg *P, uint32_t jmask, int t) {
    const BImg *a = nullptr, *b = nullptr;
    BImg *o = nullptr;
    bw6_f6_operands(A, B, P, jmask, t, a, b, o);
    bw6_round(a, b, o);
    bw6_f6_reduce(P, O, jmask, t);
    __syncthreads();
}

constexpr uint32_t BW6_DENSE = 0x3fu;
constexpr uint32_t BW6_LINE = 0x0du;                 // exponents {0, 2, 3}

struct alignas(16) Bw6MillerScratch {
    BImg F[6];                                       // Miller variable, power basis
    BImg P[36];                                      // coefficient products
    BImg T[3];                                       // running point X, Y, Z (Jacobian, on the twist)
    BImg LD[2][6];                                   // doubling line of a step, embedded in 6 coefficients (1, 3, 4, 5 stay zero)
    BImg LA[2][6];                                   // addition line of a step
    BImg W[20];                                      // temporaries of the point step
    BImg NXP, YP, X2, Y2, NY2;                       // -x_P, y_P, x_Q, y_Q, -y_Q
};

// temporaries of the doubling step
enum : int { D_A = 0, D_B, D_ZZ, D_YZ, D_E, D_XB, D_Z3, D_C, D_S, D_F, D_EZZ, D_EX, D_Z3ZZ, D_X3, D_DX, D_C0, D_C8, D_Y3P, D_C2, D_C3 };
// temporaries of the addition step
enum : int { A_ZZ = 0, A_T1, A_U2, A_S2, A_H, A_R, A_HH, A_Z3, A_RR, A_RX2, A_C2, A_HHH, A_V, A_Y2Z3, A_C3, A_X3, A_VX, A_C0, A_M1, A_M2 };

// One block: Miller value of one (pair, sub-loop).  vals[blockIdx.x] receives f_{loop,Q}(P) (sub-loop 1 already
// raised to the q-th power); a pair with an infinite member yields one.
__global__ void __launch_bounds__(BW6_THREADS) k_bw6_miller(const AffineMem<BFq> *__restrict__ g1,
                                                            const AffineMem<BFq> *__restrict__ g2, uint32_t n,
                                                            BImg *__restrict__ vals /* 2 n x 6 */) {
    __shared__ Bw6MillerScratch S;
    __shared__ int s_skip;
    const int t = threadIdx.x;
    const uint32_t pair = blockIdx.x >> 1;
    const int loop = blockIdx.x & 1;
    const int len = loop ? BW6_LOOP2_LEN : BW6_LOOP1_LEN;
    BImg *out = vals + 6 * (size_t)blockIdx.x;
    if (t == 0) {
        BFq xp = b_ld(g1[pair].x), yp = b_ld(g1[pair].y), xq = b_ld(g2[pair].x), yq = b_ld(g2[pair].y);
        s_skip = (xp.is_zero() && yp.is_zero()) || (xq.is_zero() && yq.is_zero());
        b_st(S.NXP, xp.neg());
        b_st(S.YP, yp);
        b_st(S.X2, xq);
        b_st(S.Y2, yq);
        b_st(S.NY2, yq.neg());
        b_st(S.T[0], xq);
        b_st(S.T[1], yq);
        b_st(S.T[2], BFq::one());
    }
    if (t < 6) {
        b_st(S.F[t], t == 0 ? BFq::one() : BFq::zero());
        for (int k = 0; k < 2; k++) {
            b_st(S.LD[k][t], BFq::zero());
            b_st(S.LA[k][t], BFq::zero());
        }
    }
    __syncthreads();
    if (s_skip) {
        if (t < 6) out[t] = S.F[t];
        return;
    }
    const int u = t - BW6_TBASE;                     // lane of the point chain (0..), negative for the f threads
    BImg *W = S.W;
    int cur = 0;                                     // line buffers holding the previous step's lines
#pragma unroll 1
    for (int s = len - 2; s >= -1; s--) {            // s = -1: epilogue, only the lagging f work
        const bool f_work = s != len - 2;            // f processes step s + 1
        const int d_prev = f_work ? BW6_LOOP_DIGITS[loop][s + 1] : 0;
        const bool t_work = s >= 0;
        const int d = t_work ? BW6_LOOP_DIGITS[loop][s] : 0;
        const int nxt = cur ^ 1;
        // ---- round 1: f <- f^2  ||  A = X^2, B = Y^2, ZZ = Z^2, YZ = Y Z
        {
            const BImg *a = nullptr, *b = nullptr;
            BImg *o = nullptr;
            if (f_work) bw6_f6_operands(S.F, S.F, S.P, BW6_DENSE, t, a, b, o);
            if (t_work && u >= 0 && u < 4) {
                a = u == 0 ? &S.T[0] : (u == 2 ? &S.T[2] : &S.T[1]);
                b = u == 0 ? &S.T[0] : (u == 1 ? &S.T[1] : &S.T[2]);
                o = &W[D_A + u];
            }
            bw6_round(a, b, o);
            if (f_work) bw6_f6_reduce(S.P, S.F, BW6_DENSE, t);
            if (t_work) {
                if (u == 0) {
                    BFq A = b_ld(W[D_A]);
                    b_st(W[D_E], A.dbl() + A);
                } else if (u == 1) {
                    b_st(W[D_XB], b_ld(S.T[0]) + b_ld(W[D_B]));
                } else if (u == 2) {
                    b_st(W[D_Z3], b_ld(W[D_YZ]).dbl());
                }
            }
            __syncthreads();
        }
        // ---- round 2: f <- f * line_dbl  ||  C = B^2, S = (X + B)^2, F = E^2, E ZZ, E X, Z3 ZZ
        {
            const BImg *a = nullptr, *b = nullptr;
            BImg *o = nullptr;
            if (f_work) bw6_f6_operands(S.F, S.LD[cur], S.P, BW6_LINE, t, a, b, o);
            if (t_work && u >= 0 && u < 6) {
                switch (u) {
                    case 0: a = &W[D_B]; b = &W[D_B]; o = &W[D_C]; break;
                    case 1: a = &W[D_XB]; b = &W[D_XB]; o = &W[D_S]; break;
                    case 2: a = &W[D_E]; b = &W[D_E]; o = &W[D_F]; break;
                    case 3: a = &W[D_E]; b = &W[D_ZZ]; o = &W[D_EZZ]; break;
                    case 4: a = &W[D_E]; b = &S.T[0]; o = &W[D_EX]; break;
                    default: a = &W[D_Z3]; b = &W[D_ZZ]; o = &W[D_Z3ZZ]; break;
                }
            }
            bw6_round(a, b, o);
            if (f_work) bw6_f6_reduce(S.P, S.F, BW6_LINE, t);
            if (t_work) {
                if (u == 0) {                        // D = 2 (S - A - C), X3 = F - 2 D, DX = D - X3
                    BFq D = (b_ld(W[D_S]) - b_ld(W[D_A]) - b_ld(W[D_C])).dbl();
                    BFq X3 = b_ld(W[D_F]) - D.dbl();
                    b_st(W[D_X3], X3);
                    b_st(W[D_DX], D - X3);
                } else if (u == 1) {                 // c0 = E X - 2 B
                    b_st(W[D_C0], b_ld(W[D_EX]) - b_ld(W[D_B]).dbl());
                } else if (u == 2) {
                    b_st(W[D_C8], b_ld(W[D_C]).dbl().dbl().dbl());
                }
            }
            __syncthreads();
        }
        // ---- round 3: f <- f * line_add (if the previous digit was non-zero)  ||  E (D - X3), -E ZZ x_P, Z3 ZZ y_P
        {
            const BImg *a = nullptr, *b = nullptr;
            BImg *o = nullptr;
            const bool f_add = f_work && d_prev != 0;
            if (f_add) bw6_f6_operands(S.F, S.LA[cur], S.P, BW6_LINE, t, a, b, o);
            if (t_work && u >= 0 && u < 3) {
                switch (u) {
                    case 0: a = &W[D_E]; b = &W[D_DX]; o = &W[D_Y3P]; break;
                    case 1: a = &W[D_EZZ]; b = &S.NXP; o = &W[D_C2]; break;
                    default: a = &W[D_Z3ZZ]; b = &S.YP; o = &W[D_C3]; break;
                }
            }
            bw6_round(a, b, o);
            if (f_add) bw6_f6_reduce(S.P, S.F, BW6_LINE, t);
            if (t_work) {
                if (u == 0) {
                    b_st(S.T[1], b_ld(W[D_Y3P]) - b_ld(W[D_C8]));
                    S.T[0] = W[D_X3];
                    S.T[2] = W[D_Z3];
                } else if (u == 1) {
                    S.LD[nxt][0] = W[D_C0];
                    S.LD[nxt][2] = W[D_C2];
                    S.LD[nxt][3] = W[D_C3];
                }
            }
            __syncthreads();
        }
        // ---- addition step of the point chain: T <- T +- Q, five rounds
        if (d != 0) {
            const BImg *Y2 = d > 0 ? &S.Y2 : &S.NY2;
            {   // a1: ZZ = Z^2, T1 = y2 Z
                const BImg *a = nullptr, *b = nullptr;
                BImg *o = nullptr;
                if (u == 0) { a = &S.T[2]; b = &S.T[2]; o = &W[A_ZZ]; }
                if (u == 1) { a = Y2; b = &S.T[2]; o = &W[A_T1]; }
                bw6_round(a, b, o);
            }
            {   // a2: U2 = x2 ZZ, S2 = T1 ZZ;  H = U2 - X, R = S2 - Y
                const BImg *a = nullptr, *b = nullptr;
                BImg *o = nullptr;
                if (u == 0) { a = &S.X2; b = &W[A_ZZ]; o = &W[A_U2]; }
                if (u == 1) { a = &W[A_T1]; b = &W[A_ZZ]; o = &W[A_S2]; }
                bw6_round(a, b, o);
                if (u == 0) b_st(W[A_H], b_ld(W[A_U2]) - b_ld(S.T[0]));
                if (u == 1) b_st(W[A_R], b_ld(W[A_S2]) - b_ld(S.T[1]));
                __syncthreads();
            }
            {   // a3: HH = H^2, Z3 = Z H, RR = R^2, R x2, -R x_P
                const BImg *a = nullptr, *b = nullptr;
                BImg *o = nullptr;
                switch (u) {
                    case 0: a = &W[A_H]; b = &W[A_H]; o = &W[A_HH]; break;
                    case 1: a = &S.T[2]; b = &W[A_H]; o = &W[A_Z3]; break;
                    case 2: a = &W[A_R]; b = &W[A_R]; o = &W[A_RR]; break;
                    case 3: a = &W[A_R]; b = &S.X2; o = &W[A_RX2]; break;
                    case 4: a = &W[A_R]; b = &S.NXP; o = &W[A_C2]; break;
                    default: break;
                }
                bw6_round(a, b, o);
            }
            {   // a4: HHH = H HH, V = X HH, y2 Z3, Z3 y_P;  X3 = RR - HHH - 2 V, VX = V - X3, c0 = R x2 - y2 Z3
                const BImg *a = nullptr, *b = nullptr;
                BImg *o = nullptr;
                switch (u) {
                    case 0: a = &W[A_H]; b = &W[A_HH]; o = &W[A_HHH]; break;
                    case 1: a = &S.T[0]; b = &W[A_HH]; o = &W[A_V]; break;
                    case 2: a = Y2; b = &W[A_Z3]; o = &W[A_Y2Z3]; break;
                    case 3: a = &W[A_Z3]; b = &S.YP; o = &W[A_C3]; break;
                    default: break;
                }
                bw6_round(a, b, o);
                if (u == 0) {
                    BFq V = b_ld(W[A_V]);
                    BFq X3 = b_ld(W[A_RR]) - b_ld(W[A_HHH]) - V.dbl();
                    b_st(W[A_X3], X3);
                    b_st(W[A_VX], V - X3);
                } else if (u == 1) {
                    b_st(W[A_C0], b_ld(W[A_RX2]) - b_ld(W[A_Y2Z3]));
                }
                __syncthreads();
            }
            {   // a5: R (V - X3), Y HHH;  Y3 = difference
                const BImg *a = nullptr, *b = nullptr;
                BImg *o = nullptr;
                if (u == 0) { a = &W[A_R]; b = &W[A_VX]; o = &W[A_M1]; }
                if (u == 1) { a = &S.T[1]; b = &W[A_HHH]; o = &W[A_M2]; }
                bw6_round(a, b, o);
                if (u == 0) {
                    b_st(S.T[1], b_ld(W[A_M1]) - b_ld(W[A_M2]));
                    S.T[0] = W[A_X3];
                    S.T[2] = W[A_Z3];
                } else if (u == 1) {
                    S.LA[nxt][0] = W[A_C0];
                    S.LA[nxt][2] = W[A_C2];
                    S.LA[nxt][3] = W[A_C3];
                }
                __syncthreads();
            }
        }
        cur = nxt;
    }
    // sub-loop 1 enters the product as its q-th power
    if (t < 6) {
        BFq v = b_ld(S.F[t]);
        if (loop == 1) v = v * b_const(BW6_GAMMA[0][t]);
        b_st(out[t], v);
    }
}

struct alignas(16) Bw6FinalScratch {
    BImg V[16][6];                                   // named values, power basis
    BImg P[36];
};

// V[o] = frob^j(V[a]) (j = 1, 2) or the conjugate (j = 3)
B200_DEV void bw6_f6_frob(const BImg *A, BImg *O, int j, int t) {
    if (t < 6) {
        BFq v = b_ld(A[t]);
        if (j == 3) v = (t & 1) ? v.neg() : v;
        else v = v * b_const(BW6_GAMMA[j - 1][t]);
        b_st(O[t], v);
    }
    __syncthreads();
}

// One block: product of `count` Miller values, final exponentiation.  out (may be NULL) receives the result as
// arkworks' Fq6 image (c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2); is_one (may be NULL) the comparison with one.
__global__ void __launch_bounds__(BW6_THREADS) k_bw6_final_exp(const BImg *__restrict__ in, uint32_t count,
                                                               BImg *__restrict__ out, int *__restrict__ is_one) {
    __shared__ Bw6FinalScratch S;
    __shared__ BImg s_inv;
    const int t = threadIdx.x;
    enum { X = 0, C, N, N1, N2, M, R, ACC, T0 /* .. T0 + 7: table of the joint exponentiation */ };
    BImg(*V)[6] = S.V;
    BImg *P = S.P;
    if (t < 6) V[X][t] = in[t];
    __syncthreads();
#pragma unroll 1
    for (uint32_t k = 1; k < count; k++) {
        if (t < 6) V[C][t] = in[6 * (size_t)k + t];
        __syncthreads();
        bw6_f6_mul(V[X], V[C], V[X], P, BW6_DENSE, t);
    }
    // f^-1: n = f conj(f) in F_q^3, m = n^q n^(q^2), n m in F_q
    bw6_f6_frob(V[X], V[C], 3, t);
    bw6_f6_mul(V[X], V[C], V[N], P, BW6_DENSE, t);
    bw6_f6_frob(V[N], V[N1], 1, t);
    bw6_f6_frob(V[N], V[N2], 2, t);
    bw6_f6_mul(V[N1], V[N2], V[M], P, BW6_DENSE, t);
    bw6_f6_mul(V[N], V[M], V[N1], P, BW6_DENSE, t);          // N1[0] = norm in F_q
    if (t == 0) b_st(s_inv, b_ld(V[N1][0]).inv());
    bw6_f6_mul(V[C], V[M], V[N2], P, BW6_DENSE, t);          // conj(f) m  (the barrier inside also publishes s_inv)
    if (t < 6) b_st(V[N2][t], b_ld(V[N2][t]) * b_ld(s_inv)); // f^-1
    __syncthreads();
    // easy part: r = (conj(f) / f)^(q + 1)
    bw6_f6_mul(V[C], V[N2], V[R], P, BW6_DENSE, t);
    bw6_f6_frob(V[R], V[N], 1, t);
    bw6_f6_mul(V[N], V[R], V[R], P, BW6_DENSE, t);
    // hard part: r^(R0 + q R1) = (r^-1)^(-R0) (r^q)^R1, one joint square-and-multiply over the signed digit pairs.
    // After the easy part r lies in the cyclotomic subgroup: the inverse is the conjugate.
    bw6_f6_frob(V[R], V[T0], 3, t);                          // f' = r^-1
    bw6_f6_frob(V[R], V[T0 + 1], 1, t);                      // g = r^q
    bw6_f6_mul(V[T0], V[T0 + 1], V[T0 + 2], P, BW6_DENSE, t);        // f' g
    bw6_f6_frob(V[T0 + 1], V[N], 3, t);                      // g^-1
    bw6_f6_mul(V[T0], V[N], V[T0 + 3], P, BW6_DENSE, t);     // f' g^-1
#pragma unroll 1
    for (int k = 0; k < 4; k++) bw6_f6_frob(V[T0 + k], V[T0 + 4 + k], 3, t);   // the four inverses
    {
        const int top = BW6_HARD_JSF[BW6_HARD_JSF_LEN - 1];
        const BImg *src = V[T0 + ((top & 7) - 1) + ((top & 8) ? 4 : 0)];
        if (t < 6) V[ACC][t] = src[t];
        __syncthreads();
    }
#pragma unroll 1
    for (int b = BW6_HARD_JSF_LEN - 2; b >= 0; b--) {
        bw6_f6_mul(V[ACC], V[ACC], V[ACC], P, BW6_DENSE, t);
        const int c = BW6_HARD_JSF[b];
        if (c) bw6_f6_mul(V[ACC], V[T0 + ((c & 7) - 1) + ((c & 8) ? 4 : 0)], V[ACC], P, BW6_DENSE, t);
    }
    if (out && t < 6) out[3 * (t & 1) + (t >> 1)] = V[ACC][t];
    if (is_one && t < 32) {
        bool ok = true;
        if (t < 6) {
            BFq v = b_ld(V[ACC][t]);
            ok = t == 0 ? (v == BFq::one()) : v.is_zero();
        }
        unsigned all = __ballot_sync(0xffffffffu, ok);
        if (t == 0) *is_one = all == 0xffffffffu ? 1 : 0;
    }
}

}  // namespace b200
