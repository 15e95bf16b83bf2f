// BW6-761 product of pairings and the Groth16 verification built on it: host-side launches
// (kernels in pairing_bw6.cuh).  Reference path: crates/epoch-snark/src/api/verifier.rs:35
// (ark-groth16 verify_proof), reached from the C-ABI `verify` (crates/bls-snark-sys/src/snark/mod.rs:23-45).
#include "curve_impl.cuh"
#include "pairing_bw6.cuh"
#include "pairing_bw6_coop.cuh"

namespace b200 {

static constexpr size_t FQ6_BYTES = 6 * sizeof(BImg);            // 576
static constexpr size_t AFF = sizeof(AffineMem<BFq>);            // 192: packed affine record of G1 and of G2

// Miller values of n pairs: 2 n Fq6 images (per pair: f_{u+1}, then f_{u^3-u^2-u}^q), not yet multiplied
int bw6_miller_values(Engine &E, const void *d_g1_packed, const void *d_g2_packed, size_t n, void *d_vals, cudaStream_t st) {
    (void)E;
    if (n == 0) return B200_OK;
    if (n > (1u << 20)) return fail(B200_ERR_ARG, "n = %zu pairs exceeds the 2^20 per-call limit", n);
    k_bw6_miller<<<(unsigned)(2 * n), BW6_THREADS, 0, st>>>(reinterpret_cast<const AffineMem<BFq> *>(d_g1_packed),
                                                            reinterpret_cast<const AffineMem<BFq> *>(d_g2_packed), (uint32_t)n,
                                                            reinterpret_cast<BImg *>(d_vals));
    LAUNCH_CHECK();
    return B200_OK;
}

// product of `count` Fq6 images -> final exponentiation; d_out / d_is_one may be NULL
int bw6_final_exp(Engine &E, const void *d_vals, size_t count, void *d_out, int *d_is_one, cudaStream_t st) {
    (void)E;
    if (count == 0) return fail(B200_ERR_ARG, "final exponentiation needs at least one value");
    // six warps, one output coefficient each, on the warp-cooperative field (pairing_bw6_coop.cuh);
    // B200_BW6_THREAD_KERNELS=1 selects the one-product-per-thread kernels (cross-check)
    static const bool thread_kernels = getenv("B200_BW6_THREAD_KERNELS") && atoi(getenv("B200_BW6_THREAD_KERNELS"));
    if (thread_kernels)
        k_bw6_final_exp<<<1, BW6_THREADS, 0, st>>>(reinterpret_cast<const BImg *>(d_vals), (uint32_t)count,
                                                   reinterpret_cast<BImg *>(d_out), d_is_one);
    else
        k_bw6_final_exp_coop<<<1, BW6C_THREADS, 0, st>>>(reinterpret_cast<const BImg *>(d_vals), (uint32_t)count,
                                                         reinterpret_cast<BImg *>(d_out), d_is_one);
    LAUNCH_CHECK();
    return B200_OK;
}

// y <- -y for the records selected by `mask` (bit i = record i); (0, 0) stays (0, 0)
__global__ void k_bw6_negate_y(AffineMem<BFq> *pts, uint32_t count, uint32_t mask) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count && ((mask >> i) & 1u)) b_st(pts[i].y, b_ld(pts[i].y).neg());
}

// one identity Fq6 image (the empty product)
__global__ void k_bw6_fq6_one(BImg *out) {
    if (threadIdx.x < 6) b_st(out[threadIdx.x], threadIdx.x == 0 ? BFq::one() : BFq::zero());
}

// host-pointer product of pairings: prod e(g1[i], g2[i]); out_fq6 = arkworks Fq6 image (576 B)
int bw6_multi_pairing_host(Engine &E, const void *g1, size_t stride1, const void *g2, size_t stride2, size_t n, void *out_fq6,
                           int *out_is_one) {
    cudaStream_t st = E.stream;
    int rc;
    if ((rc = E.result.reserve(FQ6_BYTES + 16)) || (rc = E.miller.reserve((2 * n + 1) * FQ6_BYTES))) return rc;
    char *vals = E.miller.as<char>();
    if (n) {
        if (stride1 % 4 || stride1 < AFF || stride2 % 4 || stride2 < AFF) return fail(B200_ERR_ARG, "bad stride");
        if ((rc = E.h2d_bases.reserve(n * stride1)) || (rc = E.h2d_g2.reserve(n * stride2)) ||
            (rc = E.native_bases.reserve(n * AFF)) || (rc = E.g2_packed.reserve(n * AFF)))
            return rc;
        CUDA_TRY(cudaMemcpyAsync(E.h2d_bases.p, g1, n * stride1, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(E.h2d_g2.p, g2, n * stride2, cudaMemcpyHostToDevice, st));
        if ((rc = pack_bases<G_761>(E.h2d_bases.p, stride1, n, E.native_bases.p, st))) return rc;
        if ((rc = pack_bases<G_761>(E.h2d_g2.p, stride2, n, E.g2_packed.p, st))) return rc;
        if ((rc = bw6_miller_values(E, E.native_bases.p, E.g2_packed.p, n, vals, st))) return rc;
    } else {
        k_bw6_fq6_one<<<1, 32, 0, st>>>(reinterpret_cast<BImg *>(vals));
        LAUNCH_CHECK();
    }
    char *res = E.result.as<char>();
    if ((rc = bw6_final_exp(E, vals, n ? 2 * n : 1, res, reinterpret_cast<int *>(res + FQ6_BYTES), st))) return rc;
    if (out_fq6) CUDA_TRY(cudaMemcpyAsync(out_fq6, res, FQ6_BYTES, cudaMemcpyDeviceToHost, st));
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, res + FQ6_BYTES, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (out_is_one) *out_is_one = flag;
    return B200_OK;
}

// The device part of verify_proof.  d_packed: packed affine records [A, (slot for g_ic), C, alpha] [B, gamma, delta, beta]
// [gamma_abc_0 .. gamma_abc_{nabc-1}]; d_scalars: nabc canonical scalars [1, input_0, ...].  alpha, gamma and delta are
// negated in place.
int bw6_groth16_verify_core(Engine &E, char *d_packed, size_t nabc, const void *d_scalars, int *out_verified) {
    cudaStream_t st = E.stream;
    int rc;
    if ((rc = E.result.reserve(FQ6_BYTES + 16 + 3 * sizeof(BImg))) || (rc = E.miller.reserve(8 * FQ6_BYTES))) return rc;
    char *res = E.result.as<char>();
    char *gic_jac = res + FQ6_BYTES + 16;                        // 288 B Jacobian image
    if ((rc = msm_native<G_761>(E, d_packed + 8 * AFF, d_scalars, nabc, gic_jac, st))) return rc;
    if ((rc = batch_to_affine<G_761>(gic_jac, 1, d_packed + 1 * AFF, st))) return rc;
    // negate alpha (record 3), gamma (5), delta (6)
    k_bw6_negate_y<<<1, 32, 0, st>>>(reinterpret_cast<AffineMem<BFq> *>(d_packed), 8, (1u << 3) | (1u << 5) | (1u << 6));
    LAUNCH_CHECK();
    char *vals = E.miller.as<char>();
    if ((rc = bw6_miller_values(E, d_packed, d_packed + 4 * AFF, 4, vals, st))) return rc;
    if ((rc = bw6_final_exp(E, vals, 8, nullptr, reinterpret_cast<int *>(res + FQ6_BYTES), st))) return rc;
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, res + FQ6_BYTES, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *out_verified = flag;
    return B200_OK;
}

// ark-groth16 verify_proof over BW6-761, all arithmetic on the device:
//   g_ic = gamma_abc[0] + sum_i input_i gamma_abc[i + 1]            (MSM with a leading unit scalar)
//   e(A, B) e(g_ic, -gamma) e(C, -delta) e(-alpha, beta) == 1
int bw6_groth16_verify(Engine &E, const b200_groth16_vk *vk, const void *proof_a, const void *proof_b, const void *proof_c,
                       const uint64_t *inputs, size_t num_inputs, int *out_verified) {
    cudaStream_t st = E.stream;
    const size_t stride = vk->stride, nabc = vk->num_gamma_abc;
    if (stride % 4 || stride < AFF) return fail(B200_ERR_ARG, "bad stride");
    int rc;
    // staging: 4 G1 records [A, g_ic, C, alpha], 4 G2 records [B, gamma, delta, beta], then gamma_abc
    const size_t raw_bytes = (8 + nabc) * stride, packed_bytes = (8 + nabc) * AFF;
    if ((rc = E.h2d_bases.reserve(raw_bytes)) || (rc = E.native_bases.reserve(packed_bytes)) ||
        (rc = E.scalars.reserve(nabc * 48)))
        return rc;
    std::vector<unsigned char> host(raw_bytes, 0);
    const void *g1s[4] = {proof_a, nullptr, proof_c, vk->alpha_g1};
    const void *g2s[4] = {proof_b, vk->gamma_g2, vk->delta_g2, vk->beta_g2};
    for (int i = 0; i < 4; i++) {
        if (g1s[i]) memcpy(host.data() + i * stride, g1s[i], stride);
        memcpy(host.data() + (4 + i) * stride, g2s[i], stride);
    }
    memcpy(host.data() + 8 * stride, vk->gamma_abc_g1, nabc * stride);
    std::vector<uint64_t> scal(nabc * 6, 0);
    scal[0] = 1;
    if (num_inputs) memcpy(scal.data() + 6, inputs, num_inputs * 48);
    CUDA_TRY(cudaMemcpyAsync(E.h2d_bases.p, host.data(), raw_bytes, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(E.scalars.p, scal.data(), nabc * 48, cudaMemcpyHostToDevice, st));
    char *packed = E.native_bases.as<char>();
    if ((rc = pack_bases<G_761>(E.h2d_bases.p, stride, 8 + nabc, packed, st))) return rc;
    // the host vectors must outlive the copies: pageable memory is staged synchronously by the runtime, but
    // do not rely on it
    CUDA_TRY(cudaStreamSynchronize(st));
    return bw6_groth16_verify_core(E, packed, nabc, E.scalars.p, out_verified);
}

}  // namespace b200
