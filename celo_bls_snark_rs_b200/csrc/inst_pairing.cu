// BLS12-377 multi-pairing: host-side launches (see pairing.cuh)
#include "curve_impl.cuh"
#include "pairing.cuh"

namespace b200 {

// Miller values of n pairs multiplied together -> d_out (one Fq12 image, before the final exponentiation)
int miller_product(Engine &E, const void *d_g1_packed, const void *d_g2_packed, size_t n, void *d_out, cudaStream_t st) {
    static_assert(sizeof(Fq12::Mem) == 576, "arkworks Fq12 image is 576 bytes");
    size_t slots = n ? n : 1;
    int rc = E.miller.reserve(slots * sizeof(Fq12::Mem));
    if (rc) return rc;
    Fq12::Mem *vals = E.miller.as<Fq12::Mem>();
    constexpr int TH = 64;
    // n == 0: the kernel writes nothing; seed slot 0 with one via a 1-pair loop over infinite points
    if (n == 0) {
        CUDA_TRY(cudaMemsetAsync(E.miller.p, 0, sizeof(Fq12::Mem), st));
        int rc2 = E.native_bases.reserve(sizeof(AffineMem<PFq>) + sizeof(AffineMem<PFq2>));
        if (rc2) return rc2;
        CUDA_TRY(cudaMemsetAsync(E.native_bases.p, 0, sizeof(AffineMem<PFq>) + sizeof(AffineMem<PFq2>), st));
        const AffineMem<PFq> *z1 = E.native_bases.as<AffineMem<PFq>>();
        const AffineMem<PFq2> *z2 = reinterpret_cast<const AffineMem<PFq2> *>(z1 + 1);
        k_miller_loop<TH><<<1, TH, 0, st>>>(z1, z2, 1, vals);
        LAUNCH_CHECK();
        n = 1;
    } else {
        k_miller_loop<TH><<<ceil_div(n, TH), TH, 0, st>>>(reinterpret_cast<const AffineMem<PFq> *>(d_g1_packed),
                                                         reinterpret_cast<const AffineMem<PFq2> *>(d_g2_packed),
                                                         (uint32_t)n, vals);
        LAUNCH_CHECK();
    }
    constexpr int PT = 128;
    k_fq12_product<PT><<<1, PT, 0, st>>>(vals, (uint32_t)n);
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpyAsync(d_out, vals, sizeof(Fq12::Mem), cudaMemcpyDeviceToDevice, st));
    return B200_OK;
}

// product of `count` Fq12 images (Miller values, e.g. one per GPU) -> final exponentiation
int final_exp(Engine &E, const void *d_vals, size_t count, void *d_out, int *d_is_one, cudaStream_t st) {
    if (count == 0) return fail(B200_ERR_ARG, "final_exp needs at least one value");
    int rc = E.miller.reserve(count * sizeof(Fq12::Mem));
    if (rc) return rc;
    Fq12::Mem *vals = E.miller.as<Fq12::Mem>();
    if (d_vals != vals) CUDA_TRY(cudaMemcpyAsync(vals, d_vals, count * sizeof(Fq12::Mem), cudaMemcpyDeviceToDevice, st));
    if (count > 1) {
        k_fq12_product<128><<<1, 128, 0, st>>>(vals, (uint32_t)count);
        LAUNCH_CHECK();
    }
    k_final_exp<<<1, 32, 0, st>>>(vals, reinterpret_cast<Fq12::Mem *>(d_out), d_is_one);   // d_out may be NULL
    LAUNCH_CHECK();
    return B200_OK;
}

}  // namespace b200
