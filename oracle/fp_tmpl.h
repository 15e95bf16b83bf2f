/* CPU oracle (TEST INFRASTRUCTURE ONLY) -- prime-field "template".
 *
 * Restates ark-ff 0.1.0 Fp384 / Fp768 Montgomery arithmetic (64-bit little-endian
 * limbs, plain 64x64->128 products: the reference builds ark-ff WITHOUT the `asm`
 * feature, crates/bls-crypto/Cargo.toml:10).  ark-ff is an un-vendored git
 * dependency (Cargo.lock: arkworks-rs/algebra#8d76d181), so this follows the
 * published CIOS algorithm; correctness is pinned by tests/ against oracle.py.
 *
 * Include with:  FP (name prefix), FP_NL (limbs).  The modulus is supplied at
 * run time through FP_(init); R^2 and -p^-1 mod 2^64 are derived from it.
 */
#include <stdint.h>
#include <string.h>

#define FP_CAT_(a, b) a##_##b
#define FP_CAT(a, b) FP_CAT_(a, b)
#define FP_(name) FP_CAT(FP, name)

typedef struct { uint64_t l[FP_NL]; } FP_(t);

static uint64_t FP_(MOD)[FP_NL];
static uint64_t FP_(INV);          /* -p^-1 mod 2^64 */
static FP_(t) FP_(R1);             /* R mod p  == Montgomery one */
static FP_(t) FP_(R2);             /* R^2 mod p */

static inline int FP_(geq_mod)(const uint64_t *a) {
    for (int i = FP_NL - 1; i >= 0; i--) {
        if (a[i] > FP_(MOD)[i]) return 1;
        if (a[i] < FP_(MOD)[i]) return 0;
    }
    return 1;
}
static inline void FP_(sub_mod)(uint64_t *a) {
    unsigned __int128 br = 0;
    for (int i = 0; i < FP_NL; i++) {
        unsigned __int128 d = (unsigned __int128)a[i] - FP_(MOD)[i] - (uint64_t)br;
        a[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
}
static inline void FP_(add)(FP_(t) *r, const FP_(t) *a, const FP_(t) *b) {
    unsigned __int128 c = 0;
    for (int i = 0; i < FP_NL; i++) {
        c += (unsigned __int128)a->l[i] + b->l[i];
        r->l[i] = (uint64_t)c;
        c >>= 64;
    }
    /* moduli here leave >= 1 spare bit (377 of 384, 761 of 768): no carry-out */
    if (FP_(geq_mod)(r->l)) FP_(sub_mod)(r->l);
}
static inline void FP_(sub)(FP_(t) *r, const FP_(t) *a, const FP_(t) *b) {
    unsigned __int128 br = 0;
    uint64_t t[FP_NL];
    for (int i = 0; i < FP_NL; i++) {
        unsigned __int128 d = (unsigned __int128)a->l[i] - b->l[i] - (uint64_t)br;
        t[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
    if (br) {
        unsigned __int128 c = 0;
        for (int i = 0; i < FP_NL; i++) {
            c += (unsigned __int128)t[i] + FP_(MOD)[i];
            t[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    memcpy(r->l, t, sizeof t);
}
static inline int FP_(is_zero)(const FP_(t) *a) {
    uint64_t o = 0;
    for (int i = 0; i < FP_NL; i++) o |= a->l[i];
    return o == 0;
}
static inline int FP_(eq)(const FP_(t) *a, const FP_(t) *b) { return memcmp(a->l, b->l, sizeof a->l) == 0; }
static inline void FP_(neg)(FP_(t) *r, const FP_(t) *a) {
    if (FP_(is_zero)(a)) { *r = *a; return; }
    unsigned __int128 br = 0;
    for (int i = 0; i < FP_NL; i++) {
        unsigned __int128 d = (unsigned __int128)FP_(MOD)[i] - a->l[i] - (uint64_t)br;
        r->l[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
}
static inline void FP_(dbl)(FP_(t) *r, const FP_(t) *a) { FP_(add)(r, a, a); }

/* CIOS Montgomery product, r = a*b/R mod p */
static inline void FP_(mul)(FP_(t) *r, const FP_(t) *a, const FP_(t) *b) {
    uint64_t t[FP_NL + 2];
    memset(t, 0, sizeof t);
    for (int i = 0; i < FP_NL; i++) {
        unsigned __int128 c = 0;
        for (int j = 0; j < FP_NL; j++) {
            c += (unsigned __int128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[FP_NL];
        t[FP_NL] = (uint64_t)c;
        t[FP_NL + 1] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * FP_(INV);
        c = (unsigned __int128)m * FP_(MOD)[0] + t[0];
        c >>= 64;
        for (int j = 1; j < FP_NL; j++) {
            c += (unsigned __int128)m * FP_(MOD)[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[FP_NL];
        t[FP_NL - 1] = (uint64_t)c;
        t[FP_NL] = t[FP_NL + 1] + (uint64_t)(c >> 64);
    }
    if (t[FP_NL] || FP_(geq_mod)(t)) FP_(sub_mod)(t);
    memcpy(r->l, t, sizeof r->l);
}
static inline void FP_(sqr)(FP_(t) *r, const FP_(t) *a) { FP_(mul)(r, a, a); }

static inline void FP_(to_mont)(FP_(t) *r, const FP_(t) *a) { FP_(mul)(r, a, &FP_(R2)); }
static inline void FP_(from_mont)(FP_(t) *r, const FP_(t) *a) {
    FP_(t) one;
    memset(&one, 0, sizeof one);
    one.l[0] = 1;
    FP_(mul)(r, a, &one);
}
/* Fermat inversion a^(p-2); inv(0) = 0 */
static void FP_(inv)(FP_(t) *r, const FP_(t) *a) {
    uint64_t e[FP_NL];
    memcpy(e, FP_(MOD), sizeof e);
    e[0] -= 2;                                  /* p is odd and > 2: no borrow */
    FP_(t) acc = FP_(R1), base = *a;
    for (int i = 0; i < 64 * FP_NL; i++) {
        if ((e[i / 64] >> (i % 64)) & 1) FP_(mul)(&acc, &acc, &base);
        FP_(sqr)(&base, &base);
    }
    *r = acc;
}
static void FP_(init)(const uint64_t *modulus) {
    memcpy(FP_(MOD), modulus, sizeof FP_(MOD));
    uint64_t x = 1;                             /* Newton: x = p^-1 mod 2^64 */
    for (int i = 0; i < 6; i++) x *= 2 - modulus[0] * x;
    FP_(INV) = (uint64_t)0 - x;
    FP_(t) v;
    memset(&v, 0, sizeof v);
    v.l[0] = 1;
    for (int i = 0; i < 2 * 64 * FP_NL; i++) {  /* v = 2^i mod p */
        FP_(add)(&v, &v, &v);
        if (i == 64 * FP_NL - 1) FP_(R1) = v;
    }
    FP_(R2) = v;
}

#undef FP
#undef FP_NL
