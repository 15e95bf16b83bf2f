/* CPU oracle (TEST INFRASTRUCTURE ONLY): plain-C restatement of the arkworks CPU
 * path that celo-bls-snark-rs calls for its MSMs.  See fp_tmpl.h / ec_tmpl.h for
 * the algorithm citations.  Built by oracle/Makefile into oracle/_build/libcpu_ref.so
 * and loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs ONLY.  The product never links this.
 *
 * "port", not "reference": the reference is Rust + un-vendored arkworks crates
 * and no Rust toolchain exists in this image, so the reference itself cannot be
 * compiled here ("Oracle").
 *
 * All field elements cross this ABI in arkworks' in-memory form: Montgomery
 * residues, 64-bit little-endian limbs.  Scalars are canonical integers
 * (PrimeField::into_repr(), signature.rs:83 / public.rs:59).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* ---------------- BLS12-377 Fq (6 limbs) and BW6-761 Fq (12 limbs) ------------- */
#define FP fq377
#define FP_NL 6
#include "fp_tmpl.h"

#define FP fq761
#define FP_NL 12
#include "fp_tmpl.h"

static const uint64_t MOD377[6] = {
    0x8508c00000000001ULL, 0x170b5d4430000000ULL, 0x1ef3622fba094800ULL,
    0x1a22d9f300f5138fULL, 0xc63b05c06ca1493bULL, 0x01ae3a4617c510eaULL };
static const uint64_t MOD761[12] = {
    0xf49d00000000008bULL, 0xe6913e6870000082ULL, 0x160cf8aeeaf0a437ULL, 0x98a116c25667a8f8ULL,
    0x71dcd3dc73ebff2eULL, 0x8689c8ed12f9fd90ULL, 0x03cebaff25b42304ULL, 0x707ba638e584e919ULL,
    0x528275ef8087be41ULL, 0xb926186a81d14688ULL, 0xd187c94004faff3eULL, 0x0122e824fb83ce0aULL };

static inline fq377_t fq377_one(void) { return fq377_R1; }
static inline fq761_t fq761_one(void) { return fq761_R1; }

/* ---------------- Fq2 = Fq[u]/(u^2 + 5) over BLS12-377 Fq ---------------------- */
typedef struct { fq377_t c0, c1; } fq2_377_t;
static inline fq2_377_t fq2_377_one(void) { fq2_377_t r; memset(&r, 0, sizeof r); r.c0 = fq377_R1; return r; }
static inline void fq2_377_add(fq2_377_t *r, const fq2_377_t *a, const fq2_377_t *b) {
    fq377_add(&r->c0, &a->c0, &b->c0); fq377_add(&r->c1, &a->c1, &b->c1); }
static inline void fq2_377_sub(fq2_377_t *r, const fq2_377_t *a, const fq2_377_t *b) {
    fq377_sub(&r->c0, &a->c0, &b->c0); fq377_sub(&r->c1, &a->c1, &b->c1); }
static inline void fq2_377_dbl(fq2_377_t *r, const fq2_377_t *a) { fq2_377_add(r, a, a); }
static inline void fq2_377_neg(fq2_377_t *r, const fq2_377_t *a) { fq377_neg(&r->c0, &a->c0); fq377_neg(&r->c1, &a->c1); }
static inline int fq2_377_is_zero(const fq2_377_t *a) { return fq377_is_zero(&a->c0) && fq377_is_zero(&a->c1); }
static inline int fq2_377_eq(const fq2_377_t *a, const fq2_377_t *b) { return fq377_eq(&a->c0, &b->c0) && fq377_eq(&a->c1, &b->c1); }
static inline void fq377_mul5(fq377_t *r, const fq377_t *a) {
    fq377_t t; fq377_dbl(&t, a); fq377_dbl(&t, &t); fq377_add(r, &t, a); }
static inline void fq2_377_mul(fq2_377_t *r, const fq2_377_t *a, const fq2_377_t *b) {
    /* (a0 + a1 u)(b0 + b1 u) = a0 b0 - 5 a1 b1 + (a0 b1 + a1 b0) u */
    fq377_t v0, v1, t0, t1;
    fq377_mul(&v0, &a->c0, &b->c0);
    fq377_mul(&v1, &a->c1, &b->c1);
    fq377_add(&t0, &a->c0, &a->c1);
    fq377_add(&t1, &b->c0, &b->c1);
    fq377_mul(&t0, &t0, &t1);
    fq377_sub(&t0, &t0, &v0);
    fq377_sub(&r->c1, &t0, &v1);
    fq377_mul5(&v1, &v1);
    fq377_sub(&r->c0, &v0, &v1);
}
static inline void fq2_377_sqr(fq2_377_t *r, const fq2_377_t *a) { fq2_377_mul(r, a, a); }
static void fq2_377_inv(fq2_377_t *r, const fq2_377_t *a) {
    fq377_t n, t;
    fq377_sqr(&n, &a->c0);
    fq377_sqr(&t, &a->c1);
    fq377_mul5(&t, &t);
    fq377_add(&n, &n, &t);                        /* a0^2 + 5 a1^2 */
    fq377_inv(&n, &n);
    fq377_mul(&r->c0, &a->c0, &n);
    fq377_mul(&t, &a->c1, &n);
    fq377_neg(&r->c1, &t);
}

/* ---------------- the four groups ---------------------------------------------- */
#define EC g1_377
#define F fq377
#define SC_NL 4
#define SC_BITS 253
#include "ec_tmpl.h"

#define EC g2_377
#define F fq2_377
#define SC_NL 4
#define SC_BITS 253
#include "ec_tmpl.h"

#define EC g_761                                  /* BW6-761 G1 and G2: both over Fq, a = 0 */
#define F fq761
#define SC_NL 6
#define SC_BITS 377
#include "ec_tmpl.h"

static int g_inited;
static void ensure_init(void) {
    if (g_inited) return;
    fq377_init(MOD377);
    fq761_init(MOD761);
    g_inited = 1;
}
void cpu_ref_init(void) { ensure_init(); }

/* curve ids shared with the tests */
enum { CURVE_BLS12_377_G1 = 0, CURVE_BLS12_377_G2 = 1, CURVE_BW6_761_G1 = 2, CURVE_BW6_761_G2 = 3 };

#define DISPATCH(curve, CALL)                           \
    switch (curve) {                                    \
    case CURVE_BLS12_377_G1: { CALL(g1_377) } break;    \
    case CURVE_BLS12_377_G2: { CALL(g2_377) } break;    \
    case CURVE_BW6_761_G1:                              \
    case CURVE_BW6_761_G2: { CALL(g_761) } break;       \
    default: return -1;                                 \
    }

/* MSM: bases = n records of `stride` bytes (x | y [| infinity flag byte]), scalars
 * = n x SC_NL limbs canonical, out = Jacobian (X|Y|Z).  Returns #window tasks. */
int cpu_ref_msm(int curve, const void *bases, size_t stride, const uint64_t *scalars, size_t n,
                void *out_jac, int threads) {
    ensure_init();
    int nw = 0;
#define CALL(E) E##_jac r; nw = E##_msm((const uint8_t *)bases, stride, scalars, n, &r, threads); memcpy(out_jac, &r, sizeof r);
    DISPATCH(curve, CALL)
#undef CALL
    return nw;
}

/* out_aff = x | y | u8 infinity, packed (2*coord + 1 bytes) */
int cpu_ref_jac_to_affine(int curve, const void *jac, void *out_aff) {
    ensure_init();
#define CALL(E) E##_jac p; E##_aff a; memcpy(&p, jac, sizeof p); E##_jac_to_affine(&a, &p); \
    memcpy(out_aff, &a.x, sizeof a.x); memcpy((uint8_t *)out_aff + sizeof a.x, &a.y, sizeof a.y); \
    ((uint8_t *)out_aff)[2 * sizeof a.x] = (uint8_t)a.inf;
    DISPATCH(curve, CALL)
#undef CALL
    return 0;
}

/* out_jac = scalar * base (one affine record, same layout as cpu_ref_msm) */
int cpu_ref_scalar_mul(int curve, const void *base, size_t stride, const uint64_t *scalar, void *out_jac) {
    ensure_init();
#define CALL(E) E##_aff a; E##_jac r; E##_load_affine(&a, (const uint8_t *)base, stride); \
    E##_scalar_mul(&r, &a, scalar); memcpy(out_jac, &r, sizeof r);
    DISPATCH(curve, CALL)
#undef CALL
    return 0;
}

/* out (n packed affine records x|y, stride 2*coord) = scalars[i] * base  -- test-input generator */
int cpu_ref_fixed_base_batch(int curve, const void *base, size_t stride, const uint64_t *scalars, size_t n,
                             void *out_affine_packed) {
    ensure_init();
#define CALL(E) E##_aff a, o; E##_jac r; E##_load_affine(&a, (const uint8_t *)base, stride);              \
    for (size_t i = 0; i < n; i++) {                                                                     \
        E##_scalar_mul(&r, &a, scalars + i * E##_SCNL);                                                     \
        E##_jac_to_affine(&o, &r);                                                                       \
        uint8_t *dst = (uint8_t *)out_affine_packed + i * 2 * sizeof o.x;                                \
        memcpy(dst, &o.x, sizeof o.x); memcpy(dst + sizeof o.x, &o.y, sizeof o.y);                       \
    }
    enum { g1_377_SCNL = 4, g2_377_SCNL = 4, g_761_SCNL = 6 };
    DISPATCH(curve, CALL)
#undef CALL
    return 0;
}

/* Montgomery <-> canonical for one base-prime-field element (nl = 6 or 12 limbs) */
int cpu_ref_to_mont(int nl, const uint64_t *in, uint64_t *out) {
    ensure_init();
    if (nl == 6) { fq377_t a; memcpy(&a, in, sizeof a); fq377_to_mont(&a, &a); memcpy(out, &a, sizeof a); return 0; }
    if (nl == 12) { fq761_t a; memcpy(&a, in, sizeof a); fq761_to_mont(&a, &a); memcpy(out, &a, sizeof a); return 0; }
    return -1;
}
int cpu_ref_from_mont(int nl, const uint64_t *in, uint64_t *out) {
    ensure_init();
    if (nl == 6) { fq377_t a; memcpy(&a, in, sizeof a); fq377_from_mont(&a, &a); memcpy(out, &a, sizeof a); return 0; }
    if (nl == 12) { fq761_t a; memcpy(&a, in, sizeof a); fq761_from_mont(&a, &a); memcpy(out, &a, sizeof a); return 0; }
    return -1;
}
int cpu_ref_fq_mul(int nl, const uint64_t *a_, const uint64_t *b_, uint64_t *out) {
    ensure_init();
    if (nl == 6) { fq377_t a, b; memcpy(&a, a_, sizeof a); memcpy(&b, b_, sizeof b); fq377_mul(&a, &a, &b); memcpy(out, &a, sizeof a); return 0; }
    if (nl == 12) { fq761_t a, b; memcpy(&a, a_, sizeof a); memcpy(&b, b_, sizeof b); fq761_mul(&a, &a, &b); memcpy(out, &a, sizeof a); return 0; }
    return -1;
}
