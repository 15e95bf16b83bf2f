// BLS12-377 multi-pairing: host-side launches (kernels in pairing_warp.cuh / pairing.cuh)
#include "curve_impl.cuh"
#include "pairing_warp.cuh"

namespace b200 {

// Cross-check kernels (one thread per pair, one pair per warp, one-warp final exponentiation) are compiled only with
// -DB200_WITH_CROSSCHECKS (python -m celo_bls_snark_rs_b200.build --crosschecks); the shipped library holds the product
// kernels alone.  With them compiled in: B200_PAIRING_THREAD=1, B200_PAIRING_SINGLE=1, B200_FINAL_EXP_WARP=1 select them.
#ifdef B200_WITH_CROSSCHECKS
static bool env_on(const char *name) { return getenv(name) && atoi(getenv(name)); }
static bool thread_path() { static const bool v = env_on("B200_PAIRING_THREAD"); return v; }
static bool warp_final_exp() { static const bool v = env_on("B200_FINAL_EXP_WARP"); return v; }
static bool single_pair_warps() { static const bool v = env_on("B200_PAIRING_SINGLE"); return v; }
#endif

// `warps` warps of the two-pairs-per-warp Miller kernel; B200_MILLER_WARPS = 1 | 2 | 4 warps per block (default 2)
// Up to 888 checks (one wave of six 128-thread blocks per SM) take the block kernel -- one block of four warps per two
// pairs: 2 pairs 1.48 -> 0.64 ms, 1024 pairs 1.81 -> 1.43 ms, 1776 pairs 2.25 -> 2.04 ms -- beyond that one warp per two
// pairs wins (2960 pairs: 3.15 against 3.24 ms, 4097: 4.22 against 4.46; profiles/r2_pairing_block.log).
// B200_MILLER_BLOCK = k overrides the limit.
static void launch_w2_miller(const AffineMem<PFq> *g1, const AffineMem<PFq2> *g2, uint32_t n, uint32_t warps, Fq12::Mem *vals,
                             cudaStream_t st) {
    static const int ww = getenv("B200_MILLER_WARPS") ? atoi(getenv("B200_MILLER_WARPS")) : 2;
    static const long block_max = getenv("B200_MILLER_BLOCK") ? atol(getenv("B200_MILLER_BLOCK")) : 888;
    if ((long)warps <= block_max) {
        k_b2_miller_loop<<<warps, B2_THREADS, 0, st>>>(g1, g2, n, vals);
        return;
    }
    if (ww == 4)
        k_w2_miller_loop<4><<<ceil_div(warps, 4), 128, 0, st>>>(g1, g2, n, vals);
    else if (ww == 1)
        k_w2_miller_loop<1><<<warps, 32, 0, st>>>(g1, g2, n, vals);
    else
        k_w2_miller_loop<2><<<ceil_div(warps, 2), 64, 0, st>>>(g1, g2, n, vals);
}

// Miller values of n pairs multiplied together -> d_out (one Fq12 image, before the final exponentiation)
int miller_product(Engine &E, const void *d_g1_packed, const void *d_g2_packed, size_t n, void *d_out, cudaStream_t st) {
    static_assert(sizeof(Fq12::Mem) == 576, "arkworks Fq12 image is 576 bytes");
    int rc = E.miller.reserve((n ? n : 1) * sizeof(Fq12::Mem));
    if (rc) return rc;
    Fq12::Mem *vals = E.miller.as<Fq12::Mem>();
    const AffineMem<PFq> *g1 = reinterpret_cast<const AffineMem<PFq> *>(d_g1_packed);
    const AffineMem<PFq2> *g2 = reinterpret_cast<const AffineMem<PFq2> *>(d_g2_packed);
    if (n == 0) {                                    // empty product: one pair of infinite points yields one
        if ((rc = E.native_bases.reserve(sizeof(AffineMem<PFq>) + sizeof(AffineMem<PFq2>)))) return rc;
        CUDA_TRY(cudaMemsetAsync(E.native_bases.p, 0, sizeof(AffineMem<PFq>) + sizeof(AffineMem<PFq2>), st));
        g1 = E.native_bases.as<AffineMem<PFq>>();
        g2 = reinterpret_cast<const AffineMem<PFq2> *>(g1 + 1);
        n = 1;
    }
#ifdef B200_WITH_CROSSCHECKS
    if (thread_path()) {
        constexpr int TH = 64;
        k_miller_loop<TH><<<ceil_div(n, TH), TH, 0, st>>>(g1, g2, (uint32_t)n, vals);
        LAUNCH_CHECK();
        k_fq12_product<128><<<1, 128, 0, st>>>(vals, (uint32_t)n);
        LAUNCH_CHECK();
    } else
#endif
    {
        uint32_t live = (uint32_t)n;                 // Miller values to fold
#ifdef B200_WITH_CROSSCHECKS
        if (single_pair_warps()) {
            k_w_miller_loop<<<ceil_div(n, W_WARPS), 32 * W_WARPS, 0, st>>>(g1, g2, (uint32_t)n, vals);
        } else
#endif
        {                                            // two pairs per warp share one Miller variable
            live = (uint32_t)((n + 1) / 2);
            launch_w2_miller(g1, g2, (uint32_t)n, live, vals, st);
        }
        LAUNCH_CHECK();
        // fold: live -> <= 4 * SMs -> <= 32 -> 1 partial products
        const uint32_t steps[3] = {(uint32_t)E.sm_count * W_WARPS, 32u, 1u};
        for (uint32_t stride : steps) {
            if (live <= stride) continue;
            k_w_fq12_strided_product<<<ceil_div(stride, W_WARPS), 32 * W_WARPS, 0, st>>>(vals, live, stride);
            LAUNCH_CHECK();
            live = stride;
        }
    }
    CUDA_TRY(cudaMemcpyAsync(d_out, vals, sizeof(Fq12::Mem), cudaMemcpyDeviceToDevice, st));
    return B200_OK;
}

// `count` independent two-pair checks (pairs 2b and 2b + 1 belong to check b): one warp runs both Miller loops of a check
// on a shared Miller variable, one block per check finishes it -> d_flags[b] = 1 iff the product is one
int pairing_checks_2(Engine &E, const void *d_g1_packed, const void *d_g2_packed, size_t count, int *d_flags, cudaStream_t st) {
    if (count == 0) return B200_OK;
    int rc = E.miller.reserve(count * sizeof(Fq12::Mem));
    if (rc) return rc;
    Fq12::Mem *vals = E.miller.as<Fq12::Mem>();
    launch_w2_miller(reinterpret_cast<const AffineMem<PFq> *>(d_g1_packed), reinterpret_cast<const AffineMem<PFq2> *>(d_g2_packed),
                     (uint32_t)(2 * count), (uint32_t)count, vals, st);
    LAUNCH_CHECK();
    k_b_final_exp<<<(unsigned)count, B_THREADS, 0, st>>>(vals, nullptr, d_flags);
    LAUNCH_CHECK();
    return B200_OK;
}

// product of `count` Fq12 images (Miller values, e.g. one per GPU) -> final exponentiation
int final_exp(Engine &E, const void *d_vals, size_t count, void *d_out, int *d_is_one, cudaStream_t st) {
    if (count == 0) return fail(B200_ERR_ARG, "final_exp needs at least one value");
    int rc = E.miller.reserve((count + 1) * sizeof(Fq12::Mem));
    if (rc) return rc;
    Fq12::Mem *vals = E.miller.as<Fq12::Mem>();
    if (d_vals != vals) CUDA_TRY(cudaMemcpyAsync(vals, d_vals, count * sizeof(Fq12::Mem), cudaMemcpyDeviceToDevice, st));
    if (count > 1) {
        k_fq12_product<128><<<1, 128, 0, st>>>(vals, (uint32_t)count);          // a handful of values (one per GPU)
        LAUNCH_CHECK();
    }
#ifdef B200_WITH_CROSSCHECKS
    if (thread_path()) {
        k_final_exp<<<1, 32, 0, st>>>(vals, reinterpret_cast<Fq12::Mem *>(d_out), d_is_one);   // d_out may be NULL
    } else if (warp_final_exp()) {
        k_w_final_exp<<<1, 32, 0, st>>>(vals, reinterpret_cast<Fq12::Mem *>(d_out), d_is_one, vals + count);
    } else
#endif
    {
        k_b_final_exp<<<1, B_THREADS, 0, st>>>(vals, reinterpret_cast<Fq12::Mem *>(d_out), d_is_one);
    }
    LAUNCH_CHECK();
    return B200_OK;
}

}  // namespace b200
