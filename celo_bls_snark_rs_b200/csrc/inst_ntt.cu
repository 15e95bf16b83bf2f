// Radix-2 NTT and the Groth16 witness-map transform chain: host-side launches (kernels in ntt.cuh).
#include "engine.cuh"
#include "ntt.cuh"

namespace b200 {

template <class F, class P>
static int ntt_prepare(Engine &E, NttDomain &D, int log_n, cudaStream_t st) {
    using M = typename F::Mem;
    if (D.log_n == log_n) return B200_OK;
    D.log_n = -1;
    int rc;
    const size_t half = log_n ? (size_t)1 << (log_n - 1) : 1;
    if ((rc = D.consts.reserve(sizeof(NttConsts<F>))) || (rc = D.pw.reserve(((size_t)4 << NTT_POW_LOG) * sizeof(M))) ||
        (rc = D.tw.reserve(half * sizeof(M))))
        return rc;
    NttConsts<F> *c = D.consts.as<NttConsts<F>>();
    k_ntt_setup<F, P><<<1, 1, 0, st>>>(log_n, c, D.pw.as<M>());
    LAUNCH_CHECK();
    k_ntt_pow_tables<F><<<(4 << NTT_POW_LOG) / 256, 256, 0, st>>>(D.pw.as<M>());
    LAUNCH_CHECK();
    k_ntt_twiddles<F><<<ceil_div(half, 256), 256, 0, st>>>(c, (uint32_t)half, D.tw.as<M>());
    LAUNCH_CHECK();
    D.log_n = log_n;
    (void)E;
    return B200_OK;
}

// all log_n stages over `data`; dit: bit-reversed in -> natural out, else natural in -> bit-reversed out
template <class F>
static int ntt_stages(const NttDomain &D, void *data, int log_n, bool dit, bool inverse, cudaStream_t st) {
    using M = typename F::Mem;
    if (log_n == 0) return B200_OK;
    const int tile_log = std::min(NTT_TILE_LOG, log_n);
    const int passes = (log_n + NTT_MAX_STAGES - 1) / NTT_MAX_STAGES;
    const int base = log_n / passes, rem = log_n % passes;
    const size_t smem = ((size_t)1 << tile_log) * sizeof(M);
    const unsigned blocks = 1u << (log_n - tile_log), threads = 1u << (tile_log - 1);
    int s = dit ? 0 : log_n;
    for (int i = 0; i < passes; i++) {
        NttPass p;
        p.log_n = log_n;
        p.tile_log = tile_log;
        p.k = base + (i < rem ? 1 : 0);
        p.s_lo = dit ? s : s - p.k;
        p.cl = std::min(p.s_lo, tile_log - p.k);
        p.eu = tile_log - p.k - p.cl;
        p.inverse = inverse ? 1 : 0;
        if (dit) k_ntt_pass<F, true><<<blocks, threads, smem, st>>>(reinterpret_cast<M *>(data), D.tw.as<M>(), p);
        else k_ntt_pass<F, false><<<blocks, threads, smem, st>>>(reinterpret_cast<M *>(data), D.tw.as<M>(), p);
        LAUNCH_CHECK();
        s = dit ? s + p.k : s - p.k;
    }
    return B200_OK;
}

template <class F>
static int ntt_scale(const NttDomain &D, void *data, int log_n, int mode, cudaStream_t st) {
    using M = typename F::Mem;
    k_ntt_scale<F><<<ceil_div((size_t)1 << log_n, 256), 256, 0, st>>>(reinterpret_cast<M *>(data), log_n, mode,
                                                                      D.consts.as<NttConsts<F>>(), D.pw.as<M>());
    LAUNCH_CHECK();
    return B200_OK;
}

// Radix2EvaluationDomain::{fft, ifft, coset_fft, coset_ifft}_in_place: natural order in and out
template <class F, class P>
static int ntt_transform_t(Engine &E, NttDomain &D, void *data, int log_n, int inverse, int coset, cudaStream_t st) {
    using M = typename F::Mem;
    int rc;
    if ((rc = ntt_prepare<F, P>(E, D, log_n, st))) return rc;
    if (coset && !inverse && (rc = ntt_scale<F>(D, data, log_n, NTT_SCALE_G, st))) return rc;      // distribute_powers(g)
    if ((rc = ntt_stages<F>(D, data, log_n, false, inverse != 0, st))) return rc;
    k_ntt_bitrev<F><<<ceil_div((size_t)1 << log_n, 256), 256, 0, st>>>(reinterpret_cast<M *>(data), log_n);
    LAUNCH_CHECK();
    if (inverse && (rc = ntt_scale<F>(D, data, log_n, NTT_SCALE_NINV | (coset ? NTT_SCALE_GINV : 0), st))) return rc;
    return B200_OK;
}

// R1CStoQAP::witness_map after the evaluation vectors a, b, c are built (appendix A.4 step 2)
template <class F, class P>
static int witness_map_t(Engine &E, NttDomain &D, void *a, void *b, void *c, int log_n, void *h, cudaStream_t st) {
    using M = typename F::Mem;
    int rc;
    if ((rc = ntt_prepare<F, P>(E, D, log_n, st))) return rc;
    const size_t n = (size_t)1 << log_n;
    for (void *x : {a, b, c}) {
        // ifft (decimation in frequency, output left bit-reversed), then coset shift and 1/n in one
        // element-wise pass, then the forward transform by decimation in time: natural order out
        if ((rc = ntt_stages<F>(D, x, log_n, false, true, st))) return rc;
        if ((rc = ntt_scale<F>(D, x, log_n, NTT_SCALE_NINV | NTT_SCALE_G | NTT_SCALE_BITREV, st))) return rc;
        if ((rc = ntt_stages<F>(D, x, log_n, true, false, st))) return rc;
    }
    k_ntt_quotient<F><<<ceil_div(n, 256), 256, 0, st>>>(reinterpret_cast<M *>(a), reinterpret_cast<const M *>(b),
                                                        reinterpret_cast<const M *>(c), (uint32_t)n,
                                                        D.consts.as<NttConsts<F>>());
    LAUNCH_CHECK();
    if ((rc = ntt_stages<F>(D, a, log_n, false, true, st))) return rc;            // coset ifft: DIF ...
    k_ntt_unpermute_unshift<F><<<ceil_div(n, 256), 256, 0, st>>>(reinterpret_cast<const M *>(a), reinterpret_cast<M *>(h),
                                                                 log_n, D.consts.as<NttConsts<F>>(), D.pw.as<M>());
    LAUNCH_CHECK();                                                                // ... un-permute, g^-i, 1/n
    return B200_OK;
}

int ntt_transform(Engine &E, int field, void *data, int log_n, int inverse, int coset, cudaStream_t st) {
    if (field == B200_FR_BLS12_377) return ntt_transform_t<Fr253, Fr253Params>(E, E.ntt[0], data, log_n, inverse, coset, st);
    return ntt_transform_t<Fq377, Fq377Params>(E, E.ntt[1], data, log_n, inverse, coset, st);
}
int witness_map(Engine &E, int field, void *a, void *b, void *c, int log_n, void *h, cudaStream_t st) {
    if (field == B200_FR_BLS12_377) return witness_map_t<Fr253, Fr253Params>(E, E.ntt[0], a, b, c, log_n, h, st);
    return witness_map_t<Fq377, Fq377Params>(E, E.ntt[1], a, b, c, log_n, h, st);
}

}  // namespace b200
