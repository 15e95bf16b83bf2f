// Radix-2 number-theoretic transforms over the Groth16 scalar fields, for sm_100a.
//
// Device replacement for ark-poly 0.1.0 Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place
// and for ark-groth16 0.1.0 R1CStoQAP::witness_map's transform chain (3 iFFT, 3 coset FFT, pointwise
// (a b - c) / Z, 1 coset iFFT), reached from crates/epoch-snark/src/api/prover.rs:78 (BW6-761:
// Fr = BLS12-377 Fq, 377 bits, two-adicity 46) and :112 (BLS12-377 Fr, 253 bits, two-adicity 47).
// SURVEY.md section 8 row f1 / appendix A.4.  Elements are arkworks' Fp384 / Fp256 memory images
// (Montgomery residues), transformed in place.
//
// Schedule: log2(n) butterfly stages are cut into passes of <= 8 stages; one block moves a tile of
// 1024 elements (48 KB / 32 KB) into shared memory, runs the pass's stages there (512 threads, one
// butterfly per thread and stage) and writes the tile back, so a 2^24-point transform is 3 round
// trips through HBM.  Tiles of the strided passes keep >= 4 consecutive elements per row (192-byte
// runs).  Twiddles omega^i (i < n/2) come from a table built once per domain; inverse transforms
// read the same table mirrored (omega^-i = -omega^(n/2 - i)).
//   decimation in frequency: natural order in  -> bit-reversed order out
//   decimation in time:      bit-reversed in   -> natural order out
// so the witness map never permutes between an inverse and the following forward transform.
// One butterfly is one Montgomery product (2 N^2 multiply-adds) + add + sub: like the MSM, the
// transform is bound by the integer-multiply pipe, not by HBM.
#pragma once
#include "ec.cuh"

namespace b200 {

constexpr int NTT_TILE_LOG = 10;
constexpr int NTT_MAX_STAGES = 8;
constexpr int NTT_POW_LOG = 12;                      // two-level power tables: g^lo, g^(hi << 12)

// per-domain constants, built by k_ntt_setup (device memory)
template <class F>
struct NttConsts {
    typename F::Mem n_inv;                           // 1 / n
    typename F::Mem z_inv;                           // 1 / (g^n - 1): vanishing polynomial on the coset gH
    typename F::Mem w2[48];                          // omega_n^(2^k)
};

template <class F, class P>
__global__ void k_ntt_setup(int log_n, NttConsts<F> *__restrict__ c, typename F::Mem *__restrict__ pw /* [4][1 << NTT_POW_LOG] */) {
    if (blockIdx.x || threadIdx.x) return;
    F root, g, gi;
#pragma unroll
    for (int i = 0; i < F::N; i++) {
        root.l[i] = P::root(i);
        g.l[i] = P::gen(i);
        gi.l[i] = P::gen_inv(i);
    }
    for (int k = log_n; k < P::TWO_ADICITY; k++) root = root.sqr();          // omega_n
    F w = root;
    for (int k = 0; k < 48; k++) {
        c->w2[k] = w.store();
        w = w.sqr();
    }
    F two = F::one() + F::one(), nn = F::one();
    for (int k = 0; k < log_n; k++) nn = nn * two;
    c->n_inv = nn.inv().store();
    F gn = g;
    for (int k = 0; k < log_n; k++) gn = gn.sqr();
    c->z_inv = (gn - F::one()).inv().store();
    // seeds of the power tables: pw[0] = g^i, pw[1] = g^(i << 12), pw[2], pw[3] the same for g^-1
    F gh = g, gih = gi;
    for (int k = 0; k < NTT_POW_LOG; k++) {
        gh = gh.sqr();
        gih = gih.sqr();
    }
    const F seeds[4] = {g, gh, gi, gih};
    for (int t = 0; t < 4; t++) {
        pw[(size_t)t << NTT_POW_LOG] = F::one().store();
        pw[((size_t)t << NTT_POW_LOG) + 1] = seeds[t].store();
    }
}

// pw[t][i] = pw[t][1]^i by binary powering (thread per entry; 4 tables of 4096)
template <class F>
__global__ void __launch_bounds__(256) k_ntt_pow_tables(typename F::Mem *__restrict__ pw) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t t = i >> NTT_POW_LOG, e = i & ((1u << NTT_POW_LOG) - 1u);
    if (t >= 4 || e < 2) return;
    F base = F::load(pw[((size_t)t << NTT_POW_LOG) + 1]), acc = F::one();
    for (int b = 31 - __clz(e); b >= 0; b--) {
        acc = acc.sqr();
        if ((e >> b) & 1u) acc = acc * base;
    }
    pw[((size_t)t << NTT_POW_LOG) + e] = acc.store();
}

// tw[i] = omega_n^i, i < n / 2: product of the w2[k] over the set bits of i
template <class F>
__global__ void __launch_bounds__(256) k_ntt_twiddles(const NttConsts<F> *__restrict__ c, uint32_t half,
                                                      typename F::Mem *__restrict__ tw) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    F acc = F::one();
    for (uint32_t e = i, k = 0; e; e >>= 1, k++)
        if (e & 1u) acc = acc * F::load(c->w2[k]);
    tw[i] = acc.store();
}

struct NttPass {
    int log_n, tile_log;
    int s_lo, k;           // stages act on index bits [s_lo, s_lo + k)
    int cl, eu;            // tile = eu upper bits | k stage bits | cl low bits (cl + k + eu = tile_log)
    int inverse;           // omega^-1 twiddles
};

// element index of tile-local position l
B200_DEV uint32_t ntt_global_index(const NttPass &p, uint32_t tile, uint32_t l) {
    const int rest = p.s_lo - p.cl;                  // low bits enumerated by the tile id
    uint32_t low = l & ((1u << p.cl) - 1u);
    uint32_t r = (l >> p.cl) & ((1u << p.k) - 1u);
    uint32_t up = l >> (p.cl + p.k);
    uint32_t tile_low = tile & ((1u << rest) - 1u), tile_high = tile >> rest;
    return (((((tile_high << p.eu) | up) << p.k | r) << rest | tile_low) << p.cl) | low;
}

template <class F>
B200_DEV F ntt_twiddle(const typename F::Mem *__restrict__ tw, uint32_t idx, uint32_t half, int inverse) {
    if (!inverse) return F::load(ldg_mem(tw + idx));
    if (idx == 0) return F::one();
    return F::load(ldg_mem(tw + (half - idx))).neg();             // omega^-i = -omega^(n/2 - i)
}

// 128-bit vector copy of a field image that the same kernel also writes (plain loads, not the read-only path)
template <class M>
B200_DEV M ntt_ldg(const M *__restrict__ src) {
    M r;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(M) / 16); k++) d4[k] = s4[k];
    return r;
}

// one pass: DIT = false: Gentleman-Sande stages from bit s_lo + k - 1 down to s_lo;
//           DIT = true : Cooley-Tukey stages from bit s_lo up to s_lo + k - 1
template <class F, bool DIT>
__global__ void __launch_bounds__(1 << (NTT_TILE_LOG - 1), 2)          // 2 tiles (96 KB) per SM: <= 64 registers
k_ntt_pass(typename F::Mem *__restrict__ data, const typename F::Mem *__restrict__ tw, NttPass p) {
    extern __shared__ __align__(16) unsigned char ntt_smem[];
    using M = typename F::Mem;
    M *sm = reinterpret_cast<M *>(ntt_smem);
    const uint32_t tid = threadIdx.x, nthreads = 1u << (p.tile_log - 1), tile = blockIdx.x;
    const uint32_t half = 1u << (p.log_n - 1);
    if (tid >= nthreads) return;                     // (blockDim.x == nthreads; defensive)
    sm[tid] = ntt_ldg(data + ntt_global_index(p, tile, tid));
    sm[tid + nthreads] = ntt_ldg(data + ntt_global_index(p, tile, tid + nthreads));
    __syncthreads();
#pragma unroll 1
    for (int step = 0; step < p.k; step++) {
        const int j = DIT ? step : p.k - 1 - step;
        const int lb = p.cl + j, b = p.s_lo + j;     // local / global bit of this stage
        const uint32_t l0 = ((tid >> lb) << (lb + 1)) | (tid & ((1u << lb) - 1u)), l1 = l0 | (1u << lb);
        const uint32_t g0 = ntt_global_index(p, tile, l0);
        const uint32_t widx = (g0 & ((1u << b) - 1u)) << (p.log_n - 1 - b);
        F w = ntt_twiddle<F>(tw, widx, half, p.inverse);
        F u = F::load(sm[l0]), v = F::load(sm[l1]);
        if (DIT) {
            v = v * w;
            sm[l0] = (u + v).store();
            sm[l1] = (u - v).store();
        } else {
            sm[l0] = (u + v).store();
            sm[l1] = ((u - v) * w).store();
        }
        __syncthreads();
    }
    data[ntt_global_index(p, tile, tid)] = sm[tid];
    data[ntt_global_index(p, tile, tid + nthreads)] = sm[tid + nthreads];
}

B200_DEV uint32_t ntt_bitrev(uint32_t i, int log_n) { return log_n ? __brev(i) >> (32 - log_n) : 0u; }

// g^e (or g^-e) from the two-level tables
template <class F>
B200_DEV F ntt_gpow(const typename F::Mem *__restrict__ pw, uint32_t e, int inverse) {
    const typename F::Mem *lo = pw + ((size_t)(inverse ? 2 : 0) << NTT_POW_LOG), *hi = lo + ((size_t)1 << NTT_POW_LOG);
    F r = F::load(ldg_mem(lo + (e & ((1u << NTT_POW_LOG) - 1u))));
    uint32_t h = e >> NTT_POW_LOG;
    if (h) r = r * F::load(ldg_mem(hi + h));
    return r;
}

// mode bits: 1 = multiply by 1/n, 2 = multiply by g^e, 4 = multiply by g^-e, 8 = position i holds index bitrev(i)
enum : int { NTT_SCALE_NINV = 1, NTT_SCALE_G = 2, NTT_SCALE_GINV = 4, NTT_SCALE_BITREV = 8 };
template <class F>
__global__ void __launch_bounds__(256) k_ntt_scale(typename F::Mem *__restrict__ data, int log_n, int mode,
                                                   const NttConsts<F> *__restrict__ c, const typename F::Mem *__restrict__ pw) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> log_n) return;
    F x = F::load(ntt_ldg(data + i));
    if (mode & (NTT_SCALE_G | NTT_SCALE_GINV)) {
        uint32_t e = (mode & NTT_SCALE_BITREV) ? ntt_bitrev(i, log_n) : i;
        x = x * ntt_gpow<F>(pw, e, mode & NTT_SCALE_GINV);
    }
    if (mode & NTT_SCALE_NINV) x = x * F::load(c->n_inv);
    data[i] = x.store();
}

// in-place bit-reversal permutation
template <class F>
__global__ void __launch_bounds__(256) k_ntt_bitrev(typename F::Mem *__restrict__ data, int log_n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> log_n) return;
    uint32_t j = ntt_bitrev(i, log_n);
    if (i < j) {
        typename F::Mem a = ntt_ldg(data + i), b = ntt_ldg(data + j);
        data[i] = b;
        data[j] = a;
    }
}

// witness map, evaluation-domain step on the coset: a <- (a b - c) / Z(g)
template <class F>
__global__ void __launch_bounds__(256) k_ntt_quotient(typename F::Mem *__restrict__ a, const typename F::Mem *__restrict__ b,
                                                      const typename F::Mem *__restrict__ cc, uint32_t n,
                                                      const NttConsts<F> *__restrict__ c) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = F::load(ntt_ldg(a + i)) * F::load(ntt_ldg(b + i)) - F::load(ntt_ldg(cc + i));
    a[i] = (x * F::load(c->z_inv)).store();
}

// witness map, last step: out[bitrev(i)] = in[i] * g^-bitrev(i) / n   (coset iFFT tail: un-permute + un-shift)
template <class F>
__global__ void __launch_bounds__(256) k_ntt_unpermute_unshift(const typename F::Mem *__restrict__ in,
                                                               typename F::Mem *__restrict__ out, int log_n,
                                                               const NttConsts<F> *__restrict__ c,
                                                               const typename F::Mem *__restrict__ pw) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> log_n) return;
    uint32_t e = ntt_bitrev(i, log_n);
    F x = F::load(ntt_ldg(in + i)) * ntt_gpow<F>(pw, e, 1);
    out[e] = (x * F::load(c->n_inv)).store();
}

}  // namespace b200
