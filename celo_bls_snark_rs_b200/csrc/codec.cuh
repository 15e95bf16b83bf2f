// Point decoding on the device: arkworks compressed encodings -> packed affine Montgomery records, with the
// checks ark-serialize 0.1.0 `GroupAffine::deserialize` makes (coordinate < modulus, x on the curve, point in the
// prime-order subgroup).  Reached in the reference from
//   crates/bls-snark-sys/src/snark/epoch_block.rs:154-196  (read_slice::<VerifyingKey / Proof>, read_pubkeys)
// i.e. the first thing the C-ABI `verify` (snark/mod.rs:23-45) does with its byte arguments.
//
// Wire format: x little-endian (Fq2: c0 | c1), two flag bits in the top of the LAST byte:
// bit 7 = "y is the larger of {y, -y}" (Fq2: compare c1 first, then c0), bit 6 = infinity.
//
// One thread per point: every step is a chain of dependent field products (a square root is an exponentiation),
// and the callers decode at most a few hundred points at a time, all in parallel.
#pragma once
#include "ec.cuh"
#include "pairing_params_gen.cuh"
#include "pairing_bw6_params_gen.cuh"

// Non-template kernels and out-of-line functions below are `static`: this header is included by more than one
// translation unit (inst_epoch_verify.cu, inst_hash.cu).
namespace b200 {

// local names: pairing.cuh (which defines CFq / CFq2 next to non-template kernels) is not included here
using CFq = Fq377;
using CFq2 = Fp2<Fq377>;
B200_DEV CFq cfq_const(const uint32_t *w) {
    CFq r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = w[i];
    return r;
}
B200_DEV CFq2 cfq2_inv(const CFq2 &a) {
    CFq n = (a.c0.sqr() + CFq2::mul5(a.c1.sqr())).inv();
    return {a.c0 * n, (a.c1 * n).neg()};
}

enum : int { DECODE_OK = 0, DECODE_INFINITY = 1, DECODE_BAD_COORD = 2, DECODE_NOT_ON_CURVE = 3, DECODE_NOT_IN_SUBGROUP = 4 };

// ---- canonical <-> Montgomery, comparisons ------------------------------------------------------------
template <class F>
B200_DEV bool fp_words_lt_modulus(const uint32_t (&w)[F::N]) {
    using P = typename F::Params;
    uint32_t t, borrow;
    sub_cc(t, w[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < F::N; i++) subc_cc(t, w[i], P::mod(i));
    subc(borrow, 0, 0);
    return borrow != 0;
}
template <class F>
B200_DEV F fp_from_canonical(const uint32_t (&w)[F::N]) {
    using P = typename F::Params;
    F a, r2;
#pragma unroll
    for (int i = 0; i < F::N; i++) {
        a.l[i] = w[i];
        r2.l[i] = P::r2(i);
    }
    return a * r2;
}
template <class F>
B200_DEV F fp_to_canonical(const F &a) {                  // limbs of the result are the canonical integer
    F one_raw = F::zero();
    one_raw.l[0] = 1;
    return a * one_raw;
}
// canonical a > (p - 1) / 2, i.e. 2 a >= p (the moduli leave the top bit of the limb array clear)
template <class F>
B200_DEV bool fp_canonical_over_half(const F &c) {
    using P = typename F::Params;
    uint32_t d[F::N], t, borrow;
    add_cc(d[0], c.l[0], c.l[0]);
#pragma unroll
    for (int i = 1; i < F::N - 1; i++) addc_cc(d[i], c.l[i], c.l[i]);
    addc(d[F::N - 1], c.l[F::N - 1], c.l[F::N - 1]);
    sub_cc(t, d[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < F::N; i++) subc_cc(t, d[i], P::mod(i));
    subc(borrow, 0, 0);
    return borrow == 0;
}

// base^e, e given as little-endian words in global / constant memory, most significant bit first
template <class F>
__device__ __noinline__ F fp_pow_words(F base, const uint32_t *exp, int bits) {
    F acc = base;                                          // top bit of the exponent is set
#pragma unroll 1
    for (int b = bits - 2; b >= 0; b--) {
        acc = acc.sqr();
        if ((exp[b >> 5] >> (b & 31)) & 1u) acc = acc * base;
    }
    return acc;
}

// ---- square roots --------------------------------------------------------------------------------------
// BW6-761 Fq: q = 3 mod 4, a^((q + 1) / 4)
B200_DEV bool fq761_sqrt(const Fq761 &a, Fq761 &out) {
    if (a.is_zero()) {
        out = a;
        return true;
    }
    Fq761 s = fp_pow_words<Fq761>(a, BW6_SQRT_EXP, BW6_SQRT_EXP_BITS);
    out = s;
    return s.sqr() == a;
}
// BLS12-377 Fq: p - 1 = 2^46 t, Tonelli-Shanks with the 2^46-th root of unity (-5)^t
static __device__ __noinline__ bool fq377_sqrt(Fq377 a, Fq377 *out) {
    if (a.is_zero()) {
        *out = a;
        return true;
    }
    const Fq377 one = Fq377::one();
    Fq377 z = fp_pow_words<Fq377>(a, FQ377_TS_EXP, FQ377_TS_EXP_BITS);          // a^((t - 1) / 2)
    Fq377 x = a * z;                                                            // a^((t + 1) / 2)
    Fq377 b = x * z;                                                            // a^t
    Fq377 g;
#pragma unroll
    for (int i = 0; i < 12; i++) g.l[i] = Fq377Params::root(i);
    int v = Fq377Params::TWO_ADICITY;
#pragma unroll 1
    for (int round = 0; round < 48; round++) {
        if (b == one) break;
        int k = 0;
        Fq377 b2 = b;
#pragma unroll 1
        while (!(b2 == one) && k < v) {
            b2 = b2.sqr();
            k++;
        }
        if (k >= v) return false;                          // a is not a square
        Fq377 w = g;
#pragma unroll 1
        for (int i = 0; i < v - k - 1; i++) w = w.sqr();
        g = w.sqr();
        x = x * w;
        b = b * g;
        v = k;
    }
    *out = x;
    return x.sqr() == a;
}
// Fq2 = Fq[u] / (u^2 + 5): norm trick (any root; the caller picks the sign)
static __device__ __noinline__ bool fq2_377_sqrt(CFq2 a, CFq2 *out) {
    if (a.is_zero()) {
        *out = a;
        return true;
    }
    CFq2 r;
    if (a.c1.is_zero()) {
        CFq s;
        if (fq377_sqrt(a.c0, &s)) {
            r = {s, CFq::zero()};
        } else {                                           // c0 is a non-residue: root = t u with -5 t^2 = c0
            CFq five_inv = CFq2::mul5(CFq::one()).inv();
            if (!fq377_sqrt((a.c0 * five_inv).neg(), &s)) return false;
            r = {CFq::zero(), s};
        }
    } else {
        CFq alpha;
        if (!fq377_sqrt(a.c0.sqr() + CFq2::mul5(a.c1.sqr()), &alpha)) return false;
        const CFq two_inv = cfq_const(PAIRING_TWO_INV);
        CFq x0;
        if (!fq377_sqrt((a.c0 + alpha) * two_inv, &x0)) {
            if (!fq377_sqrt((a.c0 - alpha) * two_inv, &x0)) return false;
        }
        r = {x0, a.c1 * x0.dbl().inv()};
    }
    *out = r;
    return r.sqr() == a;
}

// ---- decoding kernels ----------------------------------------------------------------------------------
// BW6-761: n x 96 bytes -> packed affine + status.  Records [g2_lo, g2_hi) are G2 points (y^2 = x^3 + 4), the
// others G1 (y^2 = x^3 - 1): a verifying key and a proof mix both and are decoded by one launch.
static __global__ void __launch_bounds__(64) k_bw6_decompress(const uint32_t *__restrict__ src, uint32_t n, uint32_t g2_lo, uint32_t g2_hi,
                                                       AffineMem<Fq761> *__restrict__ out, int *__restrict__ status) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool g2 = i >= g2_lo && i < g2_hi;
    uint32_t w[24];
#pragma unroll
    for (int k = 0; k < 24; k++) w[k] = src[24 * (size_t)i + k];
    const bool larger = (w[23] >> 31) & 1u, infinity = (w[23] >> 30) & 1u;
    w[23] &= 0x3fffffffu;
    AffineMem<Fq761> rec;
    rec.x = Fq761::zero().store();
    rec.y = rec.x;
    int st = DECODE_OK;
    if (infinity) {
        st = DECODE_INFINITY;
    } else if (!fp_words_lt_modulus<Fq761>(w)) {
        st = DECODE_BAD_COORD;
    } else {
        Fq761 x = fp_from_canonical<Fq761>(w);
        Fq761 one = Fq761::one();
        Fq761 rhs = x.sqr() * x + (g2 ? one.dbl().dbl() : one.neg());
        Fq761 y;
        if (!fq761_sqrt(rhs, y)) {
            st = DECODE_NOT_ON_CURVE;
        } else {
            if (fp_canonical_over_half(fp_to_canonical(y)) != larger) y = y.neg();
            rec.x = x.store();
            rec.y = y.store();
        }
    }
    out[i] = rec;
    status[i] = st;
}

// BLS12-377 G2 (Fq2 coordinates, y^2 = x^3 + (0, -1/5)): n x 96 bytes -> packed affine + status
static __global__ void __launch_bounds__(64) k_g2_377_decompress(const uint32_t *__restrict__ src, uint32_t n,
                                                          AffineMem<CFq2> *__restrict__ out, int *__restrict__ status) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w0[12], w1[12];
#pragma unroll
    for (int k = 0; k < 12; k++) {
        w0[k] = src[24 * (size_t)i + k];
        w1[k] = src[24 * (size_t)i + 12 + k];
    }
    const bool larger = (w1[11] >> 31) & 1u, infinity = (w1[11] >> 30) & 1u;
    w1[11] &= 0x3fffffffu;
    AffineMem<CFq2> rec;
    rec.x = CFq2::zero().store();
    rec.y = rec.x;
    int st = DECODE_OK;
    if (infinity) {
        st = DECODE_INFINITY;
    } else if (!fp_words_lt_modulus<CFq>(w0) || !fp_words_lt_modulus<CFq>(w1)) {
        st = DECODE_BAD_COORD;
    } else {
        CFq2 x = {fp_from_canonical<CFq>(w0), fp_from_canonical<CFq>(w1)};
        CFq2 rhs = x.sqr() * x;
        rhs.c1 = rhs.c1 + cfq_const(PAIRING_TWIST_B_C1);
        CFq2 y;
        if (!fq2_377_sqrt(rhs, &y)) {
            st = DECODE_NOT_ON_CURVE;
        } else {
            // "larger": compare c1 first, then c0 (crates/epoch-snark/src/encoding.rs:31-33)
            const bool is_larger = y.c1.is_zero() ? fp_canonical_over_half(fp_to_canonical(y.c0))
                                                  : fp_canonical_over_half(fp_to_canonical(y.c1));
            if (is_larger != larger) y = y.neg();
            rec.x = x.store();
            rec.y = y.store();
        }
    }
    out[i] = rec;
    status[i] = st;
}

// BLS12-377 G1 (y^2 = x^3 + 1): n x 48 bytes -> packed affine + status (Signature::deserialize,
// crates/bls-crypto/src/bls/signature.rs via G1Affine::deserialize)
static __global__ void __launch_bounds__(64) k_g1_377_decompress(const uint32_t *__restrict__ src, uint32_t n,
                                                                 AffineMem<CFq> *__restrict__ out, int *__restrict__ status) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[12];
#pragma unroll
    for (int k = 0; k < 12; k++) w[k] = src[12 * (size_t)i + k];
    const bool larger = (w[11] >> 31) & 1u, infinity = (w[11] >> 30) & 1u;
    w[11] &= 0x3fffffffu;
    AffineMem<CFq> rec;
    rec.x = CFq::zero().store();
    rec.y = rec.x;
    int st = DECODE_OK;
    if (infinity) {
        st = DECODE_INFINITY;
    } else if (!fp_words_lt_modulus<CFq>(w)) {
        st = DECODE_BAD_COORD;
    } else {
        CFq x = fp_from_canonical<CFq>(w), y;
        if (!fq377_sqrt(x.sqr() * x + CFq::one(), &y)) {
            st = DECODE_NOT_ON_CURVE;
        } else {
            if (fp_canonical_over_half(fp_to_canonical(y)) != larger) y = y.neg();
            rec.x = x.store();
            rec.y = y.store();
        }
    }
    out[i] = rec;
    status[i] = st;
}

// ---- encoding: GroupProjective memory images -> arkworks compressed bytes (into_affine().serialize()) ------------
// Signature::serialize / PublicKey::serialize (crates/bls-crypto/src/bls/signature.rs, public.rs:123-135).
// One thread per point: one inversion, canonical x, the "y is the larger root" bit; z = 0 -> the infinity encoding.
static __global__ void __launch_bounds__(64) k_g1_377_compress(const JacobianMem<CFq> *__restrict__ pts, uint32_t n, uint32_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const CFq X = CFq::from_ark(pts[i].x), Y = CFq::from_ark(pts[i].y), Z = CFq::from_ark(pts[i].z);
    uint32_t w[12] = {0};
    if (Z.is_zero()) {
        w[11] = 1u << 30;
    } else {
        const CFq zi = Z.inv(), zi2 = zi.sqr();
        const CFq x = fp_to_canonical(X * zi2), y = Y * zi2 * zi;
#pragma unroll
        for (int k = 0; k < 12; k++) w[k] = x.l[k];
        if (fp_canonical_over_half(fp_to_canonical(y))) w[11] |= 1u << 31;
    }
#pragma unroll
    for (int k = 0; k < 12; k++) out[12 * (size_t)i + k] = w[k];
}
static __global__ void __launch_bounds__(64) k_g2_377_compress(const JacobianMem<CFq2> *__restrict__ pts, uint32_t n, uint32_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const CFq2 X = CFq2::from_ark(pts[i].x), Y = CFq2::from_ark(pts[i].y), Z = CFq2::from_ark(pts[i].z);
    uint32_t w[24] = {0};
    if (Z.is_zero()) {
        w[23] = 1u << 30;
    } else {
        const CFq2 zi = cfq2_inv(Z), zi2 = zi.sqr();
        const CFq2 x = X * zi2, y = Y * zi2 * zi;
        const CFq c0 = fp_to_canonical(x.c0), c1 = fp_to_canonical(x.c1);
#pragma unroll
        for (int k = 0; k < 12; k++) {
            w[k] = c0.l[k];
            w[12 + k] = c1.l[k];
        }
        const bool larger = y.c1.is_zero() ? fp_canonical_over_half(fp_to_canonical(y.c0))
                                           : fp_canonical_over_half(fp_to_canonical(y.c1));
        if (larger) w[23] |= 1u << 31;
    }
#pragma unroll
    for (int k = 0; k < 24; k++) out[24 * (size_t)i + k] = w[k];
}

// r * P == O for every decoded point (is_in_correct_subgroup_assuming_on_curve); r = the words of the scalar-field
// modulus.  One thread per point, XYZZ double-and-add (exceptional cases exact, so the last addition lands on O).
template <class F, class RP>
__global__ void __launch_bounds__(64) k_subgroup_check(const AffineMem<F> *__restrict__ pts, uint32_t n, int *__restrict__ status) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || status[i] != DECODE_OK) return;
    const F px = F::load(pts[i].x), py = F::load(pts[i].y);
    XYZZ<F> acc = XYZZ<F>::inf();
    acc.madd(px, py);
#pragma unroll 1
    for (int b = RP::BITS - 2; b >= 0; b--) {
        acc.dbl();
        uint32_t word = 0;
#pragma unroll
        for (int k = 0; k < RP::N; k++) word = (k == (b >> 5)) ? RP::mod(k) : word;
        if ((word >> (b & 31)) & 1u) acc.madd(px, py);
    }
    if (!acc.is_inf()) status[i] = DECODE_NOT_IN_SUBGROUP;
}

// ---- aggregated public key of an epoch block -----------------------------------------------------------
// PublicKey::aggregate (crates/epoch-snark/src/epoch_block.rs:75-83 via encode_last_epoch_to_bits_with_aggregated_pk_cip22):
// sum of the block's G2 keys, then what encode_public_key (encoding.rs:23-47) needs of it: canonical x.c0, x.c1 and the
// "y over half" bit.  One warp: lanes sum strided subsets, a shuffle tree folds them, lane 0 normalises.
// out: 24 words of canonical x (c0 | c1), word 24 = y bit, word 25 = 1 if the sum is the point at infinity.
static __global__ void __launch_bounds__(32) k_g2_377_aggregate_emit(const AffineMem<CFq2> *__restrict__ pts, const int *__restrict__ status,
                                                              uint32_t n, uint32_t *__restrict__ out) {
    const int lane = threadIdx.x;
    XYZZ<CFq2> acc = XYZZ<CFq2>::inf();
#pragma unroll 1
    for (uint32_t i = lane; i < n; i += 32) {
        if (status[i] != DECODE_OK) continue;              // infinite keys add nothing; invalid ones fail the call on the host
        acc.madd(CFq2::load(pts[i].x), CFq2::load(pts[i].y));
    }
#pragma unroll 1
    for (int off = 16; off >= 1; off >>= 1) {
        XYZZ<CFq2> o = {acc.x.shfl(0xffffffffu, lane ^ off), acc.y.shfl(0xffffffffu, lane ^ off),
                        acc.zz.shfl(0xffffffffu, lane ^ off), acc.zzz.shfl(0xffffffffu, lane ^ off)};
        if ((lane & off) == 0) acc.add(o);                 // the partner's copy is discarded
    }
    if (lane == 0) {
        if (acc.is_inf()) {
            for (int k = 0; k < 25; k++) out[k] = 0;
            out[25] = 1;
            return;
        }
        CFq2 zi = cfq2_inv(acc.zzz);                          // x = X / ZZ = X ZZZ^-1 ZZZ / ZZ ... use ZZ^-1 = ZZZ^-2 ZZ^2
        // 1 / ZZ = ZZ^2 / ZZZ^2 (ZZ^3 = ZZZ^2): one inversion serves both coordinates
        CFq2 zz_inv = acc.zz.sqr() * zi.sqr();
        CFq2 x = acc.x * zz_inv, y = acc.y * zi;
        CFq c0 = fp_to_canonical(x.c0), c1 = fp_to_canonical(x.c1);
#pragma unroll
        for (int k = 0; k < 12; k++) {
            out[k] = c0.l[k];
            out[12 + k] = c1.l[k];
        }
        const bool over_half = y.c1.is_zero() ? fp_canonical_over_half(fp_to_canonical(y.c0))
                                              : fp_canonical_over_half(fp_to_canonical(y.c1));
        out[24] = over_half ? 1u : 0u;
        out[25] = 0;
    }
}

}  // namespace b200
