/* CPU oracle (TEST INFRASTRUCTURE ONLY) -- radix-2 transforms of the Groth16 prover in C.
 *
 * Restates ark-poly 0.1.0 Radix2EvaluationDomain::{fft, ifft, coset_fft, coset_ifft}_in_place and ark-groth16's
 * R1CStoQAP::witness_map transform chain (un-vendored git dependencies, Cargo.lock: arkworks-rs/algebra#8d76d181,
 * groth16#d8acb2b2), entered from create_proof_no_zk at crates/epoch-snark/src/api/prover.rs:78 (BW6-761: Fr =
 * BLS12-377 Fq) and :112 (BLS12-377 Fr).  Same semantics as oracle/ntt.py (natural order in and out, generator =
 * TWO_ADIC_ROOT^(2^(s - log n)), 1/n on the inverse, coset offset = multiplicative generator), which pins it in
 * tests/test_oracle_cref.py; iterative decimation-in-time after a bit reversal, butterflies of a stage split over
 * threads.  The CPU baseline of bench.py's prover sub-result and the checker that makes witness-map parity at 2^16 cheap.
 *
 * Include with NT (name prefix), NTF (field prefix), NT_NL (limbs), NT_GEN (small generator, negated if NT_GEN_NEG),
 * NT_S (two-adicity). */
#define NT_CAT_(a, b) a##_##b
#define NT_CAT(a, b) NT_CAT_(a, b)
#define NT_(name) NT_CAT(NT, name)
#define NF_(name) NT_CAT(NTF, name)
typedef NF_(t) NT_(fe);

static NT_(fe) NT_(GEN), NT_(GEN_INV), NT_(ROOT);   /* Montgomery form; ROOT = GEN^((p - 1) / 2^s) */
static int NT_(inited);

static void NT_(pow_limbs)(NT_(fe) *r, const NT_(fe) *a, const uint64_t *e, int nl) {
    NT_(fe) acc = NF_(R1), base = *a;
    for (int i = 0; i < 64 * nl; i++) {
        if ((e[i / 64] >> (i % 64)) & 1) NF_(mul)(&acc, &acc, &base);
        NF_(sqr)(&base, &base);
    }
    *r = acc;
}
static void NT_(from_small)(NT_(fe) *r, uint64_t v) {
    memset(r, 0, sizeof *r);
    r->l[0] = v;
    NF_(to_mont)(r, r);
}
static void NT_(init)(void) {
    if (NT_(inited)) return;
    ensure_init();
    NT_(from_small)(&NT_(GEN), NT_GEN);
    if (NT_GEN_NEG) NF_(neg)(&NT_(GEN), &NT_(GEN));
    NF_(inv)(&NT_(GEN_INV), &NT_(GEN));
    uint64_t e[NT_NL];
    memcpy(e, NF_(MOD), sizeof e);
    e[0] -= 1;
    for (int k = 0; k < NT_S; k++) {               /* (p - 1) >> s */
        for (int i = 0; i < NT_NL; i++) e[i] = (e[i] >> 1) | (i + 1 < NT_NL ? e[i + 1] << 63 : 0);
    }
    NT_(pow_limbs)(&NT_(ROOT), &NT_(GEN), e, NT_NL);
    NT_(inited) = 1;
}

typedef struct {
    NT_(fe) *a;
    const NT_(fe) *tw;       /* omega^i, i < n / 2 */
    size_t n;
    int tid, threads;
    pthread_barrier_t *bar;
} NT_(job);

static void *NT_(stages)(void *arg) {
    NT_(job) *J = (NT_(job) *)arg;
    const size_t n = J->n, half = n / 2;
    for (size_t len = 2; len <= n; len <<= 1) {
        const size_t h = len / 2, step = n / len;
        const size_t lo = half * (size_t)J->tid / (size_t)J->threads, hi = half * (size_t)(J->tid + 1) / (size_t)J->threads;
        for (size_t b = lo; b < hi; b++) {        /* butterfly b of this stage */
            const size_t blk = b / h, j = b % h, i0 = blk * len + j, i1 = i0 + h;
            NT_(fe) t, u = J->a[i0];
            NF_(mul)(&t, &J->a[i1], &J->tw[j * step]);
            NF_(add)(&J->a[i0], &u, &t);
            NF_(sub)(&J->a[i1], &u, &t);
        }
        if (J->threads > 1) pthread_barrier_wait(J->bar);
    }
    return NULL;
}

/* in-place transform with root omega (natural order in and out) */
static void NT_(transform)(NT_(fe) *a, unsigned log_n, const NT_(fe) *omega, int threads) {
    const size_t n = (size_t)1 << log_n;
    if (n == 1) return;
    for (size_t i = 0; i < n; i++) {               /* bit reversal */
        size_t r = 0;
        for (unsigned k = 0; k < log_n; k++) r |= ((i >> k) & 1) << (log_n - 1 - k);
        if (i < r) { NT_(fe) t = a[i]; a[i] = a[r]; a[r] = t; }
    }
    NT_(fe) *tw = (NT_(fe) *)malloc((n / 2) * sizeof(NT_(fe)));
    tw[0] = NF_(R1);
    for (size_t i = 1; i < n / 2; i++) NF_(mul)(&tw[i], &tw[i - 1], omega);
    if (threads < 1) threads = 1;
    if ((size_t)threads > n / 2) threads = (int)(n / 2);
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)threads);
    NT_(job) *jobs = (NT_(job) *)calloc((size_t)threads, sizeof(NT_(job)));
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; t++) {
        jobs[t].a = a; jobs[t].tw = tw; jobs[t].n = n; jobs[t].tid = t; jobs[t].threads = threads; jobs[t].bar = &bar;
        if (t) pthread_create(&th[t], NULL, NT_(stages), &jobs[t]);
    }
    NT_(stages)(&jobs[0]);
    for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
    pthread_barrier_destroy(&bar);
    free(jobs); free(th); free(tw);
}

static void NT_(distribute_powers)(NT_(fe) *a, size_t n, const NT_(fe) *g) {
    NT_(fe) w = NF_(R1);
    for (size_t i = 0; i < n; i++) {
        NF_(mul)(&a[i], &a[i], &w);
        NF_(mul)(&w, &w, g);
    }
}

/* fft / ifft / coset_fft / coset_ifft in place on n = 2^log_n Montgomery residues */
static int NT_(ntt)(NT_(fe) *a, unsigned log_n, int inverse, int coset, int threads) {
    NT_(init)();
    if (log_n > NT_S) return -1;
    const size_t n = (size_t)1 << log_n;
    NT_(fe) omega = NT_(ROOT);
    for (unsigned k = log_n; k < NT_S; k++) NF_(sqr)(&omega, &omega);
    if (!inverse) {
        if (coset) NT_(distribute_powers)(a, n, &NT_(GEN));
        NT_(transform)(a, log_n, &omega, threads);
    } else {
        NT_(fe) oinv, ninv;
        NF_(inv)(&oinv, &omega);
        NT_(transform)(a, log_n, &oinv, threads);
        NT_(from_small)(&ninv, (uint64_t)n);
        NF_(inv)(&ninv, &ninv);
        for (size_t i = 0; i < n; i++) NF_(mul)(&a[i], &a[i], &ninv);
        if (coset) NT_(distribute_powers)(a, n, &NT_(GEN_INV));
    }
    return 0;
}

/* h = (A B - C) / Z as n coefficients from the n evaluations of A, B, C (a, b, c are overwritten) */
static int NT_(witness_map)(NT_(fe) *a, NT_(fe) *b, NT_(fe) *c, unsigned log_n, NT_(fe) *h, int threads) {
    const size_t n = (size_t)1 << log_n;
    int rc;
    if ((rc = NT_(ntt)(a, log_n, 1, 0, threads)) || (rc = NT_(ntt)(b, log_n, 1, 0, threads))) return rc;
    NT_(ntt)(a, log_n, 0, 1, threads);
    NT_(ntt)(b, log_n, 0, 1, threads);
    NT_(ntt)(c, log_n, 1, 0, threads);
    NT_(ntt)(c, log_n, 0, 1, threads);
    NT_(fe) z = NT_(GEN), zinv;
    for (unsigned k = 0; k < log_n; k++) NF_(sqr)(&z, &z);   /* g^n */
    NF_(sub)(&z, &z, &NF_(R1));
    NF_(inv)(&zinv, &z);
    for (size_t i = 0; i < n; i++) {
        NT_(fe) t;
        NF_(mul)(&t, &a[i], &b[i]);
        NF_(sub)(&t, &t, &c[i]);
        NF_(mul)(&h[i], &t, &zinv);
    }
    return NT_(ntt)(h, log_n, 1, 1, threads);
}

#undef NT
#undef NTF
#undef NT_NL
#undef NT_GEN
#undef NT_GEN_NEG
#undef NT_S
