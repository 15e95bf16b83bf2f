// CPU check of the index / carry logic of csrc/fp_kara.cuh (Karatsuba operand product + separated Montgomery
// reduction).  The PTX carry-chain primitives are replaced by a faithful emulation (one CC flag, one instruction
// at a time), so everything except the instruction selection is exercised here, against plain 64-bit schoolbook
// arithmetic.  Test infrastructure: never linked into the product.
//   g++ -O1 -std=c++17 -I celo_bls_snark_rs_b200/csrc -I tools/experiments tools/experiments/kara_host_check.cpp -o /tmp/kara_host_check
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define __host__
#define __device__
#define B200_KARA_HOST_EMU 1
#define KDEV inline

namespace b200 {
namespace kara {
static uint32_t CC = 0;
KDEV void add_cc(uint32_t &r, uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; r = (uint32_t)s; CC = (uint32_t)(s >> 32); }
KDEV void addc_cc(uint32_t &r, uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + CC; r = (uint32_t)s; CC = (uint32_t)(s >> 32); }
KDEV void addc(uint32_t &r, uint32_t a, uint32_t b) { r = a + b + CC; }
KDEV void sub_cc(uint32_t &r, uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b; r = (uint32_t)s; CC = (uint32_t)(s >> 63); }
KDEV void subc_cc(uint32_t &r, uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b - CC; r = (uint32_t)s; CC = (uint32_t)(s >> 63); }
KDEV void subc(uint32_t &r, uint32_t a, uint32_t b) { r = a - b - CC; }
KDEV uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
KDEV void wmul(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b) { uint64_t p = (uint64_t)a * b; lo = (uint32_t)p; hi = (uint32_t)(p >> 32); }
// mad.lo.cc / madc.lo.cc d, a, b, c ; madc.hi.cc d, a, b, c -- one instruction at a time
static inline uint32_t i_mad_lo(uint32_t a, uint32_t b, uint32_t c, bool cin, bool cout) {
    uint64_t s = (uint64_t)(uint32_t)((uint64_t)a * b) + c + (cin ? CC : 0);
    if (cout) CC = (uint32_t)(s >> 32);
    return (uint32_t)s;
}
static inline uint32_t i_mad_hi(uint32_t a, uint32_t b, uint32_t c, bool cin, bool cout) {
    uint64_t s = (((uint64_t)a * b) >> 32) + c + (cin ? CC : 0);
    if (cout) CC = (uint32_t)(s >> 32);
    else if (s >> 32) { fprintf(stderr, "overflow in a top word\n"); exit(2); }
    return (uint32_t)s;
}
template <bool CIN>
KDEV void wmad(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b) {
    lo = i_mad_lo(a, b, lo, CIN, true);
    hi = i_mad_hi(a, b, hi, true, true);
}
template <bool CIN>
KDEV void wmad4(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b, uint32_t c_lo, uint32_t c_hi) {
    uint32_t l = i_mad_lo(a, b, c_lo, CIN, true);
    uint32_t h = i_mad_hi(a, b, c_hi, true, true);
    lo = l;
    hi = h;
}
template <bool CIN>
KDEV void wmad_top(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b, uint32_t c_lo) {
    uint32_t l = i_mad_lo(a, b, c_lo, CIN, true);
    uint32_t h = i_mad_hi(a, b, 0, true, false);
    lo = l;
    hi = h;
}
}  // namespace kara
}  // namespace b200

#include "fp_kara.cuh"
#include "params_gen.cuh"

using namespace b200;

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint32_t rnd() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 16);
}

template <int n>
static void ref_mul(uint32_t *t, const uint32_t *a, const uint32_t *b) {
    memset(t, 0, 8 * n);
    for (int i = 0; i < n; i++) {
        uint64_t c = 0;
        for (int j = 0; j < n; j++) {
            uint64_t s = (uint64_t)a[j] * b[i] + t[i + j] + c;
            t[i + j] = (uint32_t)s;
            c = s >> 32;
        }
        t[i + n] = (uint32_t)c;
    }
}
// t (2N limbs) / R mod p, result in [0, 2p) exactly as the device computes it (unique: t + m p with m = -t p^-1 mod R)
template <class P>
static void ref_redc(uint32_t *r, const uint32_t *t_in) {
    constexpr int N = P::N;
    uint32_t t[2 * N + 1];
    memcpy(t, t_in, 8 * N);
    t[2 * N] = 0;
    for (int i = 0; i < N; i++) {
        uint32_t m = t[i] * P::INV;
        uint64_t c = 0;
        for (int j = 0; j < N; j++) {
            uint64_t s = (uint64_t)m * P::mod(j) + t[i + j] + c;
            t[i + j] = (uint32_t)s;
            c = s >> 32;
        }
        for (int k = i + N; c && k <= 2 * N; k++) {
            uint64_t s = (uint64_t)t[k] + c;
            t[k] = (uint32_t)s;
            c = s >> 32;
        }
    }
    if (t[2 * N]) { fprintf(stderr, "reference overflow\n"); exit(2); }
    memcpy(r, t + N, 4 * N);
}

template <int n>
static void fill(uint32_t *x, int mode) {
    for (int i = 0; i < n; i++) x[i] = mode == 0 ? rnd() : mode == 1 ? 0xffffffffu : mode == 2 ? 0u : (rnd() & 1 ? 0xffffffffu : 0u);
}

template <int n, int BASE>
static int check_prod(int rounds) {
    int bad = 0;
    for (int it = 0; it < rounds; it++) {
        uint32_t a[n], b[n], t[2 * n], r[2 * n];
        fill<n>(a, it < 16 ? (it & 3) : 0);
        fill<n>(b, it < 16 ? (it >> 2) : 0);
        if (it == 16) memcpy(b, a, sizeof a);                       // equal halves -> zero differences
        if (it == 17) for (int i = 0; i < n / 2; i++) a[i] = a[n / 2 + i];
        kara::Prod<n, BASE>::run(t, a, b);
        ref_mul<n>(r, a, b);
        if (memcmp(t, r, sizeof t)) bad++;
    }
    printf("Prod<%d,%d>: %d mismatches of %d\n", n, BASE, bad, rounds);
    return bad;
}

template <class P, int BASE>
static int check_field(int rounds) {
    constexpr int N = P::N;
    int bad = 0;
    uint32_t p[N];
    for (int i = 0; i < N; i++) p[i] = P::mod(i);
    for (int it = 0; it < rounds; it++) {
        uint32_t a[N], b[N], c[N], d[N];
        // operands below p: random with the top limb cut, or p - 1 / 0 / 1 patterns
        auto gen = [&](uint32_t *x, int mode) {
            for (int i = 0; i < N; i++) x[i] = rnd();
            x[N - 1] %= p[N - 1];
            if (mode == 1) { memcpy(x, p, sizeof p); x[0] -= 1; }
            if (mode == 2) memset(x, 0, sizeof p);
            if (mode == 3) { memset(x, 0, sizeof p); x[0] = 1; }
        };
        gen(a, it < 16 ? (it & 3) : 0);
        gen(b, it < 16 ? (it >> 2) : 0);
        gen(c, it < 16 ? 1 : 0);
        gen(d, it < 16 ? ((it >> 1) & 3) : 0);
        uint32_t t[2 * N], u[2 * N], r[N], rr[N], tr[2 * N], ur[2 * N];
        kara::Prod<N, BASE>::run(t, a, b);
        kara::redc<P>(r, t, P::INV);
        ref_mul<N>(tr, a, b);
        ref_redc<P>(rr, tr);
        if (memcmp(r, rr, sizeof r)) bad++;
        // lazy sum of two products, one reduction
        kara::Prod<N, BASE>::run(u, c, d);
        uint64_t cy = 0;
        for (int i = 0; i < 2 * N; i++) {
            uint64_t s = (uint64_t)t[i] + u[i] + cy;
            t[i] = (uint32_t)s;
            cy = s >> 32;
        }
        kara::redc<P>(r, t, P::INV);
        ref_mul<N>(ur, c, d);
        cy = 0;
        for (int i = 0; i < 2 * N; i++) {
            uint64_t s = (uint64_t)tr[i] + ur[i] + cy;
            tr[i] = (uint32_t)s;
            cy = s >> 32;
        }
        ref_redc<P>(rr, tr);
        if (memcmp(r, rr, sizeof r)) bad++;
    }
    printf("field N=%d BASE=%d: %d mismatches of %d\n", N, BASE, bad, 2 * rounds);
    return bad;
}

int main() {
    int bad = 0;
    bad += check_prod<2, 2>(2000);
    bad += check_prod<4, 4>(2000);
    bad += check_prod<6, 6>(2000);
    bad += check_prod<8, 4>(2000);
    bad += check_prod<12, 6>(4000);
    bad += check_prod<12, 12>(2000);
    bad += check_prod<24, 12>(4000);
    bad += check_prod<24, 6>(4000);
    bad += check_field<Fq377Params, 6>(4000);
    bad += check_field<Fq377Params, 12>(1000);
    bad += check_field<Fq761Params, 12>(4000);
    bad += check_field<Fq761Params, 6>(4000);
    bad += check_field<Fr253Params, 4>(2000);
    printf(bad ? "FAIL\n" : "OK\n");
    return bad ? 1 : 0;
}
