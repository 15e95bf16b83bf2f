/* CPU oracle (TEST INFRASTRUCTURE ONLY) -- BLS12-377 optimal-ate product of pairings in C.
 *
 * Restates ark-ec 0.1.0 models::bls12 (un-vendored git dependency, Cargo.lock: arkworks-rs/algebra#8d76d181;
 * SURVEY.md appendix A.3) as reached from
 *   crates/bls-crypto/src/bls/signature.rs:149   Bls12_377::product_of_pairings (batch_verify_hashes, N + 1 pairs)
 *   crates/bls-crypto/src/bls/public.rs:102      the 2-pair check of verify_sig
 * G2Prepared (homogeneous projective doubling / addition steps, 63 + 6 line triples), the D-twist `ell` with
 * mul_by_034 (13 Fq2 products), Karatsuba Fq6 and complex Fq12 squaring as ark-ff, the Miller loop over all pairs (serial, as arkworks' is), and the final exponentiation (easy part,
 * then the eprint 2016/130 chain with exp_by_x).  Fq12 = Fq6[w]/(w^2 - v), Fq6 = Fq2[v]/(v^3 - u), Fq2 = Fq[u]/(u^2 + 5);
 * the memory image of an Fq12 is arkworks': c0.c0.c0, c0.c0.c1, c0.c1.c0, ... (12 Montgomery residues, 576 bytes).
 * Frobenius constants are computed at start-up (xi^((p - 1) / 6) by exponentiation, the p^2 and p^3 ones from it),
 * never typed in.  Pinned against oracle/oracle.py (itself checked on bilinearity and the reference's behavioural
 * tests) in tests/test_oracle_cref.py; it is the CPU baseline of bench.py's pairing sub-result and the byte-level
 * checker of the 4097-pair GPU test.
 * Included by cpu_ref.c after the field definitions. */
#include <pthread.h>

typedef struct { fq2_377_t c0, c1, c2; } fq6_t;
typedef struct { fq6_t c0, c1; } fq12_t;

static fq2_377_t PR_TWIST_B;          /* b / u = (0, -1/5) */
static fq377_t PR_TWO_INV;
static fq2_377_t PR_FROB[3][6];       /* [j - 1][k] = xi^(k (p^j - 1) / 6) */
static const uint64_t PR_X = 0x8508c00000000001ULL;
static int pr_inited;

static inline void fq2_mul_xi(fq2_377_t *r, const fq2_377_t *a) {          /* (a0 + a1 u) u = -5 a1 + a0 u */
    fq377_t t;
    fq377_mul5(&t, &a->c1);
    fq377_t a0 = a->c0;
    fq377_neg(&r->c0, &t);
    r->c1 = a0;
}
static inline void fq2_scale(fq2_377_t *r, const fq2_377_t *a, const fq377_t *k) {
    fq377_mul(&r->c0, &a->c0, k);
    fq377_mul(&r->c1, &a->c1, k);
}
static inline void fq2_conj(fq2_377_t *r, const fq2_377_t *a) { r->c0 = a->c0; fq377_neg(&r->c1, &a->c1); }

static void fq6_add(fq6_t *r, const fq6_t *a, const fq6_t *b) {
    fq2_377_add(&r->c0, &a->c0, &b->c0); fq2_377_add(&r->c1, &a->c1, &b->c1); fq2_377_add(&r->c2, &a->c2, &b->c2); }
static void fq6_sub(fq6_t *r, const fq6_t *a, const fq6_t *b) {
    fq2_377_sub(&r->c0, &a->c0, &b->c0); fq2_377_sub(&r->c1, &a->c1, &b->c1); fq2_377_sub(&r->c2, &a->c2, &b->c2); }
static void fq6_neg(fq6_t *r, const fq6_t *a) { fq2_377_neg(&r->c0, &a->c0); fq2_377_neg(&r->c1, &a->c1); fq2_377_neg(&r->c2, &a->c2); }
static void fq6_mul(fq6_t *r, const fq6_t *a, const fq6_t *b) {          /* Karatsuba, 6 Fq2 products (as ark-ff's Fp6) */
    fq2_377_t v0, v1, v2, s, t, c0, c1, c2;
    fq2_377_mul(&v0, &a->c0, &b->c0); fq2_377_mul(&v1, &a->c1, &b->c1); fq2_377_mul(&v2, &a->c2, &b->c2);
    fq2_377_add(&s, &a->c1, &a->c2); fq2_377_add(&t, &b->c1, &b->c2); fq2_377_mul(&c0, &s, &t);
    fq2_377_sub(&c0, &c0, &v1); fq2_377_sub(&c0, &c0, &v2); fq2_mul_xi(&c0, &c0); fq2_377_add(&c0, &c0, &v0);
    fq2_377_add(&s, &a->c0, &a->c1); fq2_377_add(&t, &b->c0, &b->c1); fq2_377_mul(&c1, &s, &t);
    fq2_377_sub(&c1, &c1, &v0); fq2_377_sub(&c1, &c1, &v1); fq2_mul_xi(&t, &v2); fq2_377_add(&c1, &c1, &t);
    fq2_377_add(&s, &a->c0, &a->c2); fq2_377_add(&t, &b->c0, &b->c2); fq2_377_mul(&c2, &s, &t);
    fq2_377_sub(&c2, &c2, &v0); fq2_377_sub(&c2, &c2, &v2); fq2_377_add(&c2, &c2, &v1);
    r->c0 = c0; r->c1 = c1; r->c2 = c2;
}
/* (a0, a1, a2) * (c0, c1, 0): 5 Fq2 products (Fp6::mul_by_01) */
static void fq6_mul_by_01(fq6_t *r, const fq6_t *a, const fq2_377_t *c0, const fq2_377_t *c1) {
    fq2_377_t aa, bb, t1, t2, t3, s, u;
    fq2_377_mul(&aa, &a->c0, c0);
    fq2_377_mul(&bb, &a->c1, c1);
    fq2_377_add(&s, &a->c1, &a->c2); fq2_377_mul(&t1, &s, c1); fq2_377_sub(&t1, &t1, &bb); fq2_mul_xi(&t1, &t1); fq2_377_add(&t1, &t1, &aa);
    fq2_377_add(&s, &a->c0, &a->c2); fq2_377_mul(&t3, &s, c0); fq2_377_sub(&t3, &t3, &aa); fq2_377_add(&t3, &t3, &bb);
    fq2_377_add(&s, &a->c0, &a->c1); fq2_377_add(&u, c0, c1); fq2_377_mul(&t2, &s, &u); fq2_377_sub(&t2, &t2, &aa); fq2_377_sub(&t2, &t2, &bb);
    r->c0 = t1; r->c1 = t2; r->c2 = t3;
}
static void fq6_mul_v(fq6_t *r, const fq6_t *a) {                          /* (a0 + a1 v + a2 v^2) v = xi a2 + a0 v + a1 v^2 */
    fq2_377_t t;
    fq2_mul_xi(&t, &a->c2);
    fq2_377_t a0 = a->c0, a1 = a->c1;
    r->c0 = t; r->c1 = a0; r->c2 = a1;
}
static void fq6_inv(fq6_t *r, const fq6_t *a) {
    fq2_377_t t0, t1, t2, s, d;
    fq2_377_sqr(&t0, &a->c0); fq2_377_mul(&s, &a->c1, &a->c2); fq2_mul_xi(&s, &s); fq2_377_sub(&t0, &t0, &s);
    fq2_377_sqr(&t1, &a->c2); fq2_mul_xi(&t1, &t1); fq2_377_mul(&s, &a->c0, &a->c1); fq2_377_sub(&t1, &t1, &s);
    fq2_377_sqr(&t2, &a->c1); fq2_377_mul(&s, &a->c0, &a->c2); fq2_377_sub(&t2, &t2, &s);
    fq2_377_mul(&d, &a->c2, &t1); fq2_377_mul(&s, &a->c1, &t2); fq2_377_add(&d, &d, &s); fq2_mul_xi(&d, &d);
    fq2_377_mul(&s, &a->c0, &t0); fq2_377_add(&d, &d, &s);
    fq2_377_inv(&d, &d);
    fq2_377_mul(&r->c0, &t0, &d); fq2_377_mul(&r->c1, &t1, &d); fq2_377_mul(&r->c2, &t2, &d);
}

static void fq12_one(fq12_t *r) { memset(r, 0, sizeof *r); r->c0.c0.c0 = fq377_R1; }
static void fq12_mul(fq12_t *r, const fq12_t *a, const fq12_t *b) {
    fq6_t t0, t1, s0, s1, c1;
    fq6_mul(&t0, &a->c0, &b->c0);
    fq6_mul(&t1, &a->c1, &b->c1);
    fq6_add(&s0, &a->c0, &a->c1); fq6_add(&s1, &b->c0, &b->c1);
    fq6_mul(&c1, &s0, &s1); fq6_sub(&c1, &c1, &t0); fq6_sub(&c1, &c1, &t1);
    fq6_mul_v(&t1, &t1);
    fq6_add(&r->c0, &t0, &t1);
    r->c1 = c1;
}
static void fq12_sqr(fq12_t *r, const fq12_t *a) {                        /* complex squaring, 2 Fq6 products */
    fq6_t ab, s0, s1, t;
    fq6_mul(&ab, &a->c0, &a->c1);
    fq6_add(&s0, &a->c0, &a->c1);
    fq6_mul_v(&t, &a->c1); fq6_add(&s1, &a->c0, &t);
    fq6_mul(&s0, &s0, &s1);                                                /* (a0 + a1)(a0 + v a1) */
    fq6_mul_v(&t, &ab);
    fq6_sub(&s0, &s0, &ab); fq6_sub(&r->c0, &s0, &t);
    fq6_add(&r->c1, &ab, &ab);
}
/* f * ((c0, 0, 0), (d0, d1, 0)): 13 Fq2 products (Fp12::mul_by_034) */
static void fq12_mul_by_034(fq12_t *f, const fq2_377_t *c0, const fq2_377_t *d0, const fq2_377_t *d1) {
    fq6_t a, b, e;
    fq2_377_t s;
    fq2_377_mul(&a.c0, &f->c0.c0, c0); fq2_377_mul(&a.c1, &f->c0.c1, c0); fq2_377_mul(&a.c2, &f->c0.c2, c0);
    fq6_mul_by_01(&b, &f->c1, d0, d1);
    fq2_377_add(&s, c0, d0);
    fq6_add(&e, &f->c0, &f->c1);
    fq6_mul_by_01(&e, &e, &s, d1);
    fq6_sub(&e, &e, &a); fq6_sub(&f->c1, &e, &b);
    fq6_mul_v(&b, &b);
    fq6_add(&f->c0, &a, &b);
}
static void fq12_conj(fq12_t *r, const fq12_t *a) { r->c0 = a->c0; fq6_neg(&r->c1, &a->c1); }
static void fq12_inv(fq12_t *r, const fq12_t *a) {
    fq6_t d, t;
    fq6_mul(&d, &a->c0, &a->c0); fq6_mul(&t, &a->c1, &a->c1); fq6_mul_v(&t, &t); fq6_sub(&d, &d, &t);
    fq6_inv(&d, &d);
    fq6_mul(&r->c0, &a->c0, &d);
    fq6_mul(&t, &a->c1, &d); fq6_neg(&r->c1, &t);
}
static void fq12_frob(fq12_t *r, const fq12_t *a, int j) {                 /* basis w^k: c0 = (w^0, w^2, w^4), c1 = (w^1, w^3, w^5) */
    const fq2_377_t *co = PR_FROB[j - 1];
    const fq2_377_t *src[6] = {&a->c0.c0, &a->c1.c0, &a->c0.c1, &a->c1.c1, &a->c0.c2, &a->c1.c2};
    fq2_377_t out[6];
    for (int k = 0; k < 6; k++) {
        fq2_377_t t = *src[k];
        if (j & 1) fq2_conj(&t, &t);
        fq2_377_mul(&out[k], &t, &co[k]);
    }
    r->c0.c0 = out[0]; r->c1.c0 = out[1]; r->c0.c1 = out[2]; r->c1.c1 = out[3]; r->c0.c2 = out[4]; r->c1.c2 = out[5];
}

static void fq2_pow_limbs(fq2_377_t *r, const fq2_377_t *a, const uint64_t *e, int nl) {
    fq2_377_t acc = fq2_377_one(), base = *a;
    for (int i = 0; i < 64 * nl; i++) {
        if ((e[i / 64] >> (i % 64)) & 1) fq2_377_mul(&acc, &acc, &base);
        fq2_377_sqr(&base, &base);
    }
    *r = acc;
}

static void pairing_init(void) {
    if (pr_inited) return;
    ensure_init();
    fq377_t two, five;
    fq377_add(&two, &fq377_R1, &fq377_R1);
    fq377_inv(&PR_TWO_INV, &two);
    fq377_mul5(&five, &fq377_R1);
    memset(&PR_TWIST_B, 0, sizeof PR_TWIST_B);
    fq377_inv(&five, &five);
    fq377_neg(&PR_TWIST_B.c1, &five);
    /* (p - 1) / 6 by long division on the limbs */
    uint64_t e[6];
    memcpy(e, MOD377, sizeof e);
    e[0] -= 1;
    unsigned __int128 rem = 0;
    for (int i = 5; i >= 0; i--) {
        unsigned __int128 cur = (rem << 64) | e[i];
        e[i] = (uint64_t)(cur / 6);
        rem = cur % 6;
    }
    fq2_377_t xi, g1, g2, g3, t;
    memset(&xi, 0, sizeof xi);
    xi.c1 = fq377_R1;
    fq2_pow_limbs(&g1, &xi, e, 6);                                        /* xi^((p - 1) / 6) */
    fq2_conj(&t, &g1);
    fq2_377_mul(&g2, &g1, &t);                                            /* ^(p + 1)     = xi^((p^2 - 1) / 6) */
    fq2_377_mul(&g3, &g2, &g1);                                           /* ^(p^2 + p + 1) = xi^((p^3 - 1) / 6) */
    const fq2_377_t *gs[3] = {&g1, &g2, &g3};
    for (int j = 0; j < 3; j++) {
        PR_FROB[j][0] = fq2_377_one();
        for (int k = 1; k < 6; k++) fq2_377_mul(&PR_FROB[j][k], &PR_FROB[j][k - 1], gs[j]);
    }
    pr_inited = 1;
}

typedef struct { fq2_377_t c0, c1, c2; } pr_line;
typedef struct { fq2_377_t x, y, z; } pr_g2proj;

static void pr_doubling_step(pr_g2proj *r, pr_line *l) {
    fq2_377_t a, b, c, e, f, g, h, i, j, e2, t;
    fq2_377_mul(&a, &r->x, &r->y); fq2_scale(&a, &a, &PR_TWO_INV);
    fq2_377_sqr(&b, &r->y);
    fq2_377_sqr(&c, &r->z);
    fq2_377_add(&t, &c, &c); fq2_377_add(&t, &t, &c); fq2_377_mul(&e, &PR_TWIST_B, &t);
    fq2_377_add(&f, &e, &e); fq2_377_add(&f, &f, &e);
    fq2_377_add(&g, &b, &f); fq2_scale(&g, &g, &PR_TWO_INV);
    fq2_377_add(&t, &r->y, &r->z); fq2_377_sqr(&h, &t); fq2_377_add(&t, &b, &c); fq2_377_sub(&h, &h, &t);
    fq2_377_sub(&i, &e, &b);
    fq2_377_sqr(&j, &r->x);
    fq2_377_sqr(&e2, &e);
    fq2_377_sub(&t, &b, &f); fq2_377_mul(&r->x, &a, &t);
    fq2_377_sqr(&g, &g); fq2_377_add(&t, &e2, &e2); fq2_377_add(&t, &t, &e2); fq2_377_sub(&r->y, &g, &t);
    fq2_377_mul(&r->z, &b, &h);
    fq2_377_neg(&l->c0, &h);
    fq2_377_add(&l->c1, &j, &j); fq2_377_add(&l->c1, &l->c1, &j);
    l->c2 = i;
}
static void pr_addition_step(pr_g2proj *r, const fq2_377_t *qx, const fq2_377_t *qy, pr_line *l) {
    fq2_377_t theta, lam, c, d, e, f, g, h, t, u;
    fq2_377_mul(&t, qy, &r->z); fq2_377_sub(&theta, &r->y, &t);
    fq2_377_mul(&t, qx, &r->z); fq2_377_sub(&lam, &r->x, &t);
    fq2_377_sqr(&c, &theta);
    fq2_377_sqr(&d, &lam);
    fq2_377_mul(&e, &lam, &d);
    fq2_377_mul(&f, &r->z, &c);
    fq2_377_mul(&g, &r->x, &d);
    fq2_377_add(&h, &e, &f); fq2_377_add(&t, &g, &g); fq2_377_sub(&h, &h, &t);
    fq2_377_mul(&r->x, &lam, &h);
    fq2_377_sub(&t, &g, &h); fq2_377_mul(&t, &theta, &t); fq2_377_mul(&u, &e, &r->y); fq2_377_sub(&r->y, &t, &u);
    fq2_377_mul(&r->z, &r->z, &e);
    fq2_377_mul(&t, &theta, qx); fq2_377_mul(&u, &lam, qy);
    l->c0 = lam;
    fq2_377_neg(&l->c1, &theta);
    fq2_377_sub(&l->c2, &t, &u);
}

/* f <- f * (c0 P.y + c1 P.x w^3... ): the D-twist sparse element ((c0, 0, 0), (d0, d1, 0)) */
static void pr_ell(fq12_t *f, const pr_line *l, const fq377_t *px, const fq377_t *py) {
    fq2_377_t c0, c1;
    fq2_scale(&c0, &l->c0, py);
    fq2_scale(&c1, &l->c1, px);
    fq12_mul_by_034(f, &c0, &c1, &l->c2);
}

#define PR_LINES 69
typedef struct {
    const uint8_t *g1, *g2;
    size_t stride1, stride2, lo, hi;
    fq12_t out;
} pr_job;

static int pr_is_inf(const uint8_t *rec, size_t stride, size_t coord) {
    if (stride > 2 * coord && rec[2 * coord]) return 1;
    for (size_t i = 0; i < 2 * coord; i++)
        if (rec[i]) return 0;
    return 1;                                                              /* packed records: (0, 0) */
}

/* Miller loop over pairs [lo, hi): G2 prepared per pair first (as G2Prepared::from), then one shared f */
static void *pr_miller_range(void *arg) {
    pr_job *J = (pr_job *)arg;
    size_t cap = J->hi - J->lo, m = 0;
    pr_line *lines = (pr_line *)malloc(cap ? cap * PR_LINES * sizeof(pr_line) : 1);
    fq377_t *pxy = (fq377_t *)malloc(cap ? cap * 2 * sizeof(fq377_t) : 1);
    for (size_t i = J->lo; i < J->hi; i++) {
        const uint8_t *r1 = J->g1 + i * J->stride1, *r2 = J->g2 + i * J->stride2;
        if (pr_is_inf(r1, J->stride1, 48) || pr_is_inf(r2, J->stride2, 96)) continue;   /* skipped, as arkworks does */
        memcpy(&pxy[2 * m], r1, 96);
        fq2_377_t qx, qy;
        memcpy(&qx, r2, 96);
        memcpy(&qy, r2 + 96, 96);
        pr_g2proj R;
        R.x = qx; R.y = qy; R.z = fq2_377_one();
        pr_line *L = lines + m * PR_LINES;
        int k = 0;
        for (int b = 62; b >= 0; b--) {
            pr_doubling_step(&R, &L[k++]);
            if ((PR_X >> b) & 1) pr_addition_step(&R, &qx, &qy, &L[k++]);
        }
        m++;
    }
    fq12_t f;
    fq12_one(&f);
    int idx = 0;
    for (int b = 62; b >= 0; b--) {
        fq12_sqr(&f, &f);
        for (size_t p = 0; p < m; p++) pr_ell(&f, &lines[p * PR_LINES + idx], &pxy[2 * p], &pxy[2 * p + 1]);
        idx++;
        if ((PR_X >> b) & 1) {
            for (size_t p = 0; p < m; p++) pr_ell(&f, &lines[p * PR_LINES + idx], &pxy[2 * p], &pxy[2 * p + 1]);
            idx++;
        }
    }
    J->out = f;                                                            /* x > 0: no conjugation */
    free(lines);
    free(pxy);
    return NULL;
}

static void pr_exp_by_x(fq12_t *r, const fq12_t *a) {
    fq12_t acc, base = *a;
    fq12_one(&acc);
    for (int i = 0; i < 64; i++) {
        if ((PR_X >> i) & 1) fq12_mul(&acc, &acc, &base);
        fq12_sqr(&base, &base);
    }
    *r = acc;
}

static void pr_final_exp(fq12_t *out, const fq12_t *f) {
    fq12_t f1, f2, r, y0, y1, y2, y3, y4, y5, t;
    fq12_conj(&f1, f);
    fq12_inv(&f2, f);
    fq12_mul(&r, &f1, &f2);
    fq12_frob(&t, &r, 2); fq12_mul(&r, &t, &r);                            /* easy part: (p^6 - 1)(p^2 + 1) */
    fq12_sqr(&t, &r); fq12_conj(&y0, &t);
    pr_exp_by_x(&y5, &r);
    fq12_sqr(&y1, &y5);
    fq12_mul(&y3, &y0, &y5);
    pr_exp_by_x(&y0, &y3);
    pr_exp_by_x(&y2, &y0);
    pr_exp_by_x(&t, &y2); fq12_mul(&y4, &t, &y1);
    pr_exp_by_x(&y1, &y4);
    fq12_conj(&y3, &y3);
    fq12_mul(&y1, &y1, &y3); fq12_mul(&y1, &y1, &r);
    fq12_conj(&y3, &r);
    fq12_mul(&t, &y0, &r); fq12_frob(&y0, &t, 3);
    fq12_mul(&t, &y4, &y3); fq12_frob(&y4, &t, 1);
    fq12_mul(&t, &y5, &y2); fq12_frob(&y5, &t, 2);
    fq12_mul(&t, &y5, &y0); fq12_mul(&t, &t, &y4); fq12_mul(out, &t, &y1);
}

/* product_of_pairings over n (G1Affine, G2Affine) records (arkworks layouts, `stride` bytes each; an infinity flag
 * byte follows the coordinates when stride exceeds them).  threads = 1 is arkworks' own schedule (one serial Miller
 * loop); threads > 1 cuts the pairs into ranges whose Miller values are multiplied -- the same GT element.
 * phase: 0 = everything; 1 = Miller product only (out_fq12 = Miller value); 2 = final exponentiation of *out_fq12. */
int cpu_ref_multi_pairing(const void *g1, size_t stride1, const void *g2, size_t stride2, size_t n, void *out_fq12,
                          int *out_is_one, int threads, int phase) {
    pairing_init();
    fq12_t f;
    if (phase != 2) {
        if (threads < 1) threads = 1;
        if ((size_t)threads > n) threads = n ? (int)n : 1;
        pr_job *jobs = (pr_job *)calloc((size_t)threads, sizeof(pr_job));
        pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
        for (int t = 0; t < threads; t++) {
            jobs[t].g1 = (const uint8_t *)g1; jobs[t].g2 = (const uint8_t *)g2;
            jobs[t].stride1 = stride1; jobs[t].stride2 = stride2;
            jobs[t].lo = n * (size_t)t / (size_t)threads; jobs[t].hi = n * (size_t)(t + 1) / (size_t)threads;
            if (t) pthread_create(&th[t], NULL, pr_miller_range, &jobs[t]);
        }
        pr_miller_range(&jobs[0]);
        f = jobs[0].out;
        for (int t = 1; t < threads; t++) {
            pthread_join(th[t], NULL);
            fq12_mul(&f, &f, &jobs[t].out);
        }
        free(jobs);
        free(th);
    } else {
        memcpy(&f, out_fq12, sizeof f);
    }
    if (phase != 1) pr_final_exp(&f, &f);
    if (out_fq12) memcpy(out_fq12, &f, sizeof f);
    if (out_is_one) {
        fq12_t one;
        fq12_one(&one);
        *out_is_one = memcmp(&one, &f, sizeof f) == 0;
    }
    return 0;
}
