/* CPU oracle (TEST INFRASTRUCTURE ONLY) -- short-Weierstrass (a = 0) group law and
 * VariableBaseMSM::multi_scalar_mul "template".
 *
 * Restates ark-ec 0.1.0 (un-vendored git dependency, Cargo.lock:
 * arkworks-rs/algebra#8d76d181) as reached from
 *   crates/bls-crypto/src/bls/signature.rs:85   (G1 MSM)
 *   crates/bls-crypto/src/bls/public.rs:61      (G2 MSM)
 *   crates/epoch-snark/src/api/prover.rs:78,112 (Groth16 MSMs, via ark-groth16)
 * following SURVEY.md appendix A.1 / A.2: Jacobian coordinates, dbl-2009-l,
 * madd-2007-bl, add-2007-bl, window rule c = 3 | floor(ceil_log2(n)*69/100)+2,
 * zero scalars skipped, unit scalars added once in window 0, running-sum bucket
 * reduction, high->low combine with c doublings, threads over windows only (the
 * decomposition rayon uses upstream).
 *
 * Include with: EC (prefix), F (coordinate field prefix), SC_NL (scalar limbs),
 * SC_BITS (scalar field modulus bits).
 */
#include <pthread.h>
#include <stdlib.h>

#define EC_CAT_(a, b) a##_##b
#define EC_CAT(a, b) EC_CAT_(a, b)
#define EC_(name) EC_CAT(EC, name)
#define F_(name) EC_CAT(F, name)
typedef F_(t) EC_(fe);

typedef struct { EC_(fe) x, y; int inf; } EC_(aff);
typedef struct { EC_(fe) x, y, z; } EC_(jac);       /* (X/Z^2, Y/Z^3); infinity <=> Z == 0 */

static inline void EC_(jac_zero)(EC_(jac) *p) {
    memset(p, 0, sizeof *p);
    p->x = F_(one)();
    p->y = F_(one)();
}
static inline int EC_(jac_is_zero)(const EC_(jac) *p) { return F_(is_zero)(&p->z); }

/* dbl-2009-l */
static void EC_(jac_double)(EC_(jac) *p) {
    if (EC_(jac_is_zero)(p)) return;
    EC_(fe) a, b, c, d, e, f, t;
    F_(sqr)(&a, &p->x);
    F_(sqr)(&b, &p->y);
    F_(sqr)(&c, &b);
    F_(add)(&t, &p->x, &b);
    F_(sqr)(&t, &t);
    F_(sub)(&t, &t, &a);
    F_(sub)(&t, &t, &c);
    F_(dbl)(&d, &t);
    F_(dbl)(&e, &a);
    F_(add)(&e, &e, &a);
    F_(sqr)(&f, &e);
    F_(mul)(&p->z, &p->z, &p->y);
    F_(dbl)(&p->z, &p->z);
    F_(sub)(&p->x, &f, &d);
    F_(sub)(&p->x, &p->x, &d);
    F_(sub)(&t, &d, &p->x);
    F_(mul)(&t, &t, &e);
    F_(dbl)(&c, &c);
    F_(dbl)(&c, &c);
    F_(dbl)(&c, &c);
    F_(sub)(&p->y, &t, &c);
}

/* madd-2007-bl with the arkworks special cases */
static void EC_(jac_add_mixed)(EC_(jac) *p, const EC_(aff) *q) {
    if (q->inf) return;
    if (EC_(jac_is_zero)(p)) {
        p->x = q->x;
        p->y = q->y;
        p->z = F_(one)();
        return;
    }
    EC_(fe) z1z1, u2, s2, h, hh, i, j, r, v, t;
    F_(sqr)(&z1z1, &p->z);
    F_(mul)(&u2, &q->x, &z1z1);
    F_(mul)(&s2, &q->y, &p->z);
    F_(mul)(&s2, &s2, &z1z1);
    if (F_(eq)(&p->x, &u2) && F_(eq)(&p->y, &s2)) {
        EC_(jac_double)(p);
        return;
    }
    F_(sub)(&h, &u2, &p->x);
    F_(sqr)(&hh, &h);
    F_(dbl)(&i, &hh);
    F_(dbl)(&i, &i);
    F_(mul)(&j, &h, &i);
    F_(sub)(&r, &s2, &p->y);
    F_(dbl)(&r, &r);
    F_(mul)(&v, &p->x, &i);
    F_(sqr)(&t, &r);
    F_(sub)(&t, &t, &j);
    F_(sub)(&t, &t, &v);
    F_(sub)(&t, &t, &v);                         /* X3 */
    EC_(fe) y1j;
    F_(mul)(&y1j, &p->y, &j);
    F_(dbl)(&y1j, &y1j);
    p->x = t;
    F_(sub)(&t, &v, &p->x);
    F_(mul)(&t, &t, &r);
    F_(sub)(&p->y, &t, &y1j);
    F_(add)(&t, &p->z, &h);
    F_(sqr)(&t, &t);
    F_(sub)(&t, &t, &z1z1);
    F_(sub)(&p->z, &t, &hh);
}

/* add-2007-bl with the arkworks special cases */
static void EC_(jac_add)(EC_(jac) *p, const EC_(jac) *q) {
    if (EC_(jac_is_zero)(p)) { *p = *q; return; }
    if (EC_(jac_is_zero)(q)) return;
    EC_(fe) z1z1, z2z2, u1, u2, s1, s2, h, i, j, r, v, t;
    F_(sqr)(&z1z1, &p->z);
    F_(sqr)(&z2z2, &q->z);
    F_(mul)(&u1, &p->x, &z2z2);
    F_(mul)(&u2, &q->x, &z1z1);
    F_(mul)(&s1, &p->y, &q->z);
    F_(mul)(&s1, &s1, &z2z2);
    F_(mul)(&s2, &q->y, &p->z);
    F_(mul)(&s2, &s2, &z1z1);
    if (F_(eq)(&u1, &u2) && F_(eq)(&s1, &s2)) {
        EC_(jac_double)(p);
        return;
    }
    F_(sub)(&h, &u2, &u1);
    F_(dbl)(&i, &h);
    F_(sqr)(&i, &i);
    F_(mul)(&j, &h, &i);
    F_(sub)(&r, &s2, &s1);
    F_(dbl)(&r, &r);
    F_(mul)(&v, &u1, &i);
    F_(sqr)(&t, &r);
    F_(sub)(&t, &t, &j);
    F_(sub)(&t, &t, &v);
    F_(sub)(&t, &t, &v);                         /* X3 */
    EC_(fe) s1j;
    F_(mul)(&s1j, &s1, &j);
    F_(dbl)(&s1j, &s1j);
    p->x = t;
    F_(sub)(&t, &v, &p->x);
    F_(mul)(&t, &t, &r);
    F_(sub)(&p->y, &t, &s1j);
    F_(add)(&t, &p->z, &q->z);
    F_(sqr)(&t, &t);
    F_(sub)(&t, &t, &z1z1);
    F_(sub)(&t, &t, &z2z2);
    F_(mul)(&p->z, &t, &h);
}

static void EC_(jac_to_affine)(EC_(aff) *r, const EC_(jac) *p) {
    memset(r, 0, sizeof *r);
    if (EC_(jac_is_zero)(p)) {
        r->inf = 1;
        r->y = F_(one)();                        /* arkworks zero(): (0, 1, infinity) */
        return;
    }
    EC_(fe) zi, zi2;
    F_(inv)(&zi, &p->z);
    F_(sqr)(&zi2, &zi);
    F_(mul)(&r->x, &p->x, &zi2);
    F_(mul)(&zi2, &zi2, &zi);
    F_(mul)(&r->y, &p->y, &zi2);
}

/* scalar helpers: canonical little-endian 64-bit limbs */
static inline int EC_(sc_is_zero)(const uint64_t *s) {
    uint64_t o = 0;
    for (int i = 0; i < SC_NL; i++) o |= s[i];
    return o == 0;
}
static inline int EC_(sc_is_one)(const uint64_t *s) {
    uint64_t o = s[0] ^ 1;
    for (int i = 1; i < SC_NL; i++) o |= s[i];
    return o == 0;
}
static inline uint64_t EC_(sc_window)(const uint64_t *s, int start, int c) {
    int limb = start / 64, off = start % 64;
    uint64_t v = s[limb] >> off;
    if (off + c > 64 && limb + 1 < SC_NL) v |= s[limb + 1] << (64 - off);
    return v & (((uint64_t)1 << c) - 1);
}

/* double-and-add, MSB first (ProjectiveCurve::mul) */
static void EC_(scalar_mul)(EC_(jac) *r, const EC_(aff) *p, const uint64_t *s) {
    EC_(jac_zero)(r);
    for (int i = 64 * SC_NL - 1; i >= 0; i--) {
        EC_(jac_double)(r);
        if ((s[i / 64] >> (i % 64)) & 1) EC_(jac_add_mixed)(r, p);
    }
}

static void EC_(load_affine)(EC_(aff) *a, const uint8_t *rec, size_t stride) {
    memcpy(&a->x, rec, sizeof a->x);
    memcpy(&a->y, rec + sizeof a->x, sizeof a->y);
    if (stride > 2 * sizeof a->x) a->inf = rec[2 * sizeof a->x] != 0;
    else a->inf = F_(is_zero)(&a->x) && F_(is_zero)(&a->y);
}

typedef struct {
    const uint8_t *bases;
    size_t stride;
    const uint64_t *scalars;
    size_t n;
    int c, num_windows;
    volatile int *next;                          /* work-stealing window counter */
    EC_(jac) *window_sums;
} EC_(msm_job);

static void EC_(msm_window)(const EC_(msm_job) *job, int w) {
    const int c = job->c, w_start = w * c;
    const size_t nb = ((size_t)1 << c) - 1;
    EC_(jac) res, running, *buckets = (EC_(jac) *)malloc(nb * sizeof(EC_(jac)));
    EC_(jac_zero)(&res);
    for (size_t b = 0; b < nb; b++) EC_(jac_zero)(&buckets[b]);
    for (size_t i = 0; i < job->n; i++) {
        const uint64_t *s = job->scalars + i * SC_NL;
        if (EC_(sc_is_zero)(s)) continue;
        EC_(aff) q;
        if (EC_(sc_is_one)(s)) {
            if (w_start == 0) {
                EC_(load_affine)(&q, job->bases + i * job->stride, job->stride);
                EC_(jac_add_mixed)(&res, &q);
            }
            continue;
        }
        uint64_t d = EC_(sc_window)(s, w_start, c);
        if (d) {
            EC_(load_affine)(&q, job->bases + i * job->stride, job->stride);
            EC_(jac_add_mixed)(&buckets[d - 1], &q);
        }
    }
    EC_(jac_zero)(&running);
    for (size_t b = nb; b-- > 0;) {
        EC_(jac_add)(&running, &buckets[b]);
        EC_(jac_add)(&res, &running);
    }
    free(buckets);
    job->window_sums[w] = res;
}

static void *EC_(msm_worker)(void *arg) {
    const EC_(msm_job) *job = (const EC_(msm_job) *)arg;
    for (;;) {
        int w = __sync_fetch_and_add(job->next, 1);
        if (w >= job->num_windows) break;
        EC_(msm_window)(job, w);
    }
    return NULL;
}

static int EC_(ceil_log2)(size_t n) {
    int l = 0;
    while (((size_t)1 << l) < n) l++;
    return l;
}

/* returns the number of window tasks (= useful threads), result in *out */
static int EC_(msm)(const uint8_t *bases, size_t stride, const uint64_t *scalars, size_t n,
                    EC_(jac) *out, int threads) {
    int c = n < 32 ? 3 : EC_(ceil_log2)(n) * 69 / 100 + 2;
    int nw = (SC_BITS + c - 1) / c;
    EC_(jac) *ws = (EC_(jac) *)malloc(nw * sizeof(EC_(jac)));
    volatile int next = 0;
    EC_(msm_job) job = { bases, stride, scalars, n, c, nw, &next, ws };
    if (threads > nw) threads = nw;
    if (threads < 1) threads = 1;
    pthread_t *tid = (pthread_t *)malloc(threads * sizeof(pthread_t));
    for (int t = 1; t < threads; t++) pthread_create(&tid[t], NULL, EC_(msm_worker), &job);
    EC_(msm_worker)(&job);
    for (int t = 1; t < threads; t++) pthread_join(tid[t], NULL);
    free(tid);
    EC_(jac) total;
    EC_(jac_zero)(&total);
    for (int w = nw - 1; w >= 1; w--) {
        EC_(jac_add)(&total, &ws[w]);
        for (int k = 0; k < c; k++) EC_(jac_double)(&total);
    }
    *out = ws[0];
    EC_(jac_add)(out, &total);
    free(ws);
    return nw;
}

#undef EC
#undef F
#undef SC_NL
#undef SC_BITS
