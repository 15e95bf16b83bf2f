// Callers of the hot path, composed on the device: the two batch-verification flows of
// bls-crypto after message hashing (hash-to-curve itself is out of scope).
//   * Signature::batch_verify_hashes   crates/bls-crypto/src/bls/signature.rs:125-155
//   * Batch::verify                     crates/bls-crypto/src/bls/batch.rs:44-84
//     (PublicKey::batch public.rs:47-65, Signature::batch signature.rs:70-89,
//      PublicKey::verify_sig public.rs:94-120)
// Inputs are the Rust types' memory images: Signature = G1Projective (144 B),
// PublicKey = G2Projective (288 B), exponents after into_repr() (4 x u64, canonical).
#include <vector>

#include "curve_impl.cuh"
#include "pairing_params_gen.cuh"

namespace b200 {

static int upload(Buffer &b, const void *host, size_t bytes, cudaStream_t st) {
    int rc = b.reserve(bytes ? bytes : 16);
    if (rc) return rc;
    if (bytes) CUDA_TRY(cudaMemcpyAsync(b.p, host, bytes, cudaMemcpyHostToDevice, st));
    return B200_OK;
}

static int read_flag(Engine &E, int *d_flag, int *out, cudaStream_t st) {
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *out = flag;
    (void)E;
    return B200_OK;
}

// e(sigma, -g2) * prod e(H_i, pk_i) == 1
int batch_verify_hashes(Engine &E, const void *signature, const void *pubkeys, const void *hashes, size_t n,
                        int *out_verified) {
    cudaStream_t st = E.stream;
    int rc;
    const size_t G1J = 144, G2J = 288, G1A = 96, G2A = 192;
    // [sigma, H_0 .. H_{n-1}] Jacobian -> affine ; [pk_0 ..] Jacobian -> affine behind -g2
    if ((rc = E.v_g1jac.reserve((n + 1) * G1J)) || (rc = E.v_g2jac.reserve((n ? n : 1) * G2J)) ||
        (rc = E.v_g1aff.reserve((n + 1) * G1A)) || (rc = E.v_g2aff.reserve((n + 1) * G2A)) ||
        (rc = E.result.reserve(576 + 16)))
        return rc;
    char *g1j = E.v_g1jac.as<char>();
    CUDA_TRY(cudaMemcpyAsync(g1j, signature, G1J, cudaMemcpyHostToDevice, st));
    if (n) {
        CUDA_TRY(cudaMemcpyAsync(g1j + G1J, hashes, n * G1J, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(E.v_g2jac.p, pubkeys, n * G2J, cudaMemcpyHostToDevice, st));
    }
    char *g2a = E.v_g2aff.as<char>();
    CUDA_TRY(cudaMemcpyAsync(g2a, NEG_G2_GENERATOR_PACKED, G2A, cudaMemcpyHostToDevice, st));
    if ((rc = batch_to_affine<G1_377>(g1j, n + 1, E.v_g1aff.p, st))) return rc;      // into_affine() x (N + 1)
    if ((rc = batch_to_affine<G2_377>(E.v_g2jac.p, n, g2a + G2A, st))) return rc;
    char *res = E.result.as<char>();
    if ((rc = miller_product(E, E.v_g1aff.p, g2a, n + 1, res, st))) return rc;
    if ((rc = final_exp(E, res, 1, nullptr, reinterpret_cast<int *>(res + 576), st))) return rc;
    return read_flag(E, reinterpret_cast<int *>(res + 576), out_verified, st);
}

// Batch::verify with caller-supplied exponents: bpk = sum r_i pk_i, bsig = sum r_i sig_i,
// then e(bsig, -g2) * e(H, bpk) == 1
int batch_verify_strict_hash(Engine &E, const void *pubkeys, const void *signatures, const uint64_t *exponents,
                             size_t n, const void *message_hash, int *out_verified) {
    cudaStream_t st = E.stream;
    int rc;
    const size_t G1J = 144, G2J = 288, G1A = 96, G2A = 192;
    if ((rc = E.v_g1jac.reserve((n + 2) * G1J)) || (rc = E.v_g2jac.reserve((n + 1) * G2J)) ||
        (rc = E.v_g1aff.reserve((n + 2) * G1A)) || (rc = E.v_g2aff.reserve((n + 2) * G2A)) ||
        (rc = upload(E.scalars, exponents, n * 32, st)) || (rc = E.result.reserve(576 + 16)))
        return rc;
    char *g1j = E.v_g1jac.as<char>(), *g2j = E.v_g2jac.as<char>();
    char *g1a = E.v_g1aff.as<char>(), *g2a = E.v_g2aff.as<char>();
    if (n) {
        CUDA_TRY(cudaMemcpyAsync(g1j, signatures, n * G1J, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(g2j, pubkeys, n * G2J, cudaMemcpyHostToDevice, st));
    }
    // batch_normalization_into_affine (signature.rs:82, public.rs:58)
    if ((rc = batch_to_affine<G1_377>(g1j, n, g1a, st))) return rc;
    if ((rc = batch_to_affine<G2_377>(g2j, n, g2a, st))) return rc;
    // the two MSMs (signature.rs:85, public.rs:61); results land behind the inputs as Jacobian
    char *bsig_j = g1j + n * G1J, *bpk_j = g2j + n * G2J;
    if ((rc = msm_native<G1_377>(E, g1a, E.scalars.p, n, bsig_j, st))) return rc;
    if ((rc = msm_native<G2_377>(E, g2a, E.scalars.p, n, bpk_j, st))) return rc;
    // verify_sig (public.rs:94-120): pairs (bsig, -g2), (H, bpk)
    CUDA_TRY(cudaMemcpyAsync(bsig_j + G1J, message_hash, G1J, cudaMemcpyHostToDevice, st));
    char *p1 = g1a + n * G1A, *p2 = g2a + n * G2A;                // 2 G1 records, 2 G2 records
    if ((rc = batch_to_affine<G1_377>(bsig_j, 2, p1, st))) return rc;          // [bsig, H]
    CUDA_TRY(cudaMemcpyAsync(p2, NEG_G2_GENERATOR_PACKED, G2A, cudaMemcpyHostToDevice, st));
    if ((rc = batch_to_affine<G2_377>(bpk_j, 1, p2 + G2A, st))) return rc;     // [-g2, bpk]
    char *res = E.result.as<char>();
    if ((rc = miller_product(E, p1, p2, 2, res, st))) return rc;
    if ((rc = final_exp(E, res, 1, nullptr, reinterpret_cast<int *>(res + 576), st))) return rc;
    return read_flag(E, reinterpret_cast<int *>(res + 576), out_verified, st);
}

// Batch::verify for MANY batches in one pass (the shape batch_verify_strict is called with,
// crates/bls-snark-sys/src/signatures.rs:343-404, and the reference's own benchmark: 300 batches of 20,
// benches/batch_bls.rs:62-95): all keys and signatures normalised in two launches, every batch's G1 and G2 MSM as one
// block of k_small_msm each (the two curves side by side on two streams), all 2-pair checks in one Miller launch (one
// warp per batch) and one final-exponentiation launch (one block per batch).  Nothing but the flags returns to the host.
int batch_verify_strict_many(Engine &E, const b200_strict_batch *batches, size_t count, int *out_verified) {
    cudaStream_t st = E.stream, side = E.pipe_stream[1];
    const size_t G1J = 144, G2J = 288, G1A = 96, G2A = 192;
    std::vector<uint32_t> offsets(count + 1, 0);
    int bits = 1;
    for (size_t b = 0; b < count; b++) {
        offsets[b + 1] = offsets[b] + (uint32_t)batches[b].n;
        if (batches[b].n && (!batches[b].pubkeys || !batches[b].signatures || !batches[b].exponents)) return fail(B200_ERR_ARG, "null pointer in batch %zu", b);
        if (!batches[b].message_hash) return fail(B200_ERR_ARG, "null message hash in batch %zu", b);
    }
    const size_t total = offsets[count];
    // host gather: one contiguous block per kind (the handles of a batch are scattered heap objects on the caller's side)
    std::vector<uint8_t> sigs(total * G1J + 16), pks(total * G2J + 16), hashes(count * G1J + 16), neg(count * G2A + 16);
    std::vector<uint64_t> exps(4 * total + 4);
    for (size_t b = 0; b < count; b++) {
        const size_t lo = offsets[b], n = batches[b].n;
        memcpy(&sigs[lo * G1J], batches[b].signatures, n * G1J);
        memcpy(&pks[lo * G2J], batches[b].pubkeys, n * G2J);
        memcpy(&exps[4 * lo], batches[b].exponents, n * 32);
        memcpy(&hashes[b * G1J], batches[b].message_hash, G1J);
        memcpy(&neg[b * G2A], NEG_G2_GENERATOR_PACKED, G2A);
    }
    for (size_t i = 0; i < 4 * total; i++) {                      // highest set bit over all exponents
        if (!exps[i]) continue;
        const int top = 64 * (int)(i & 3) + 64 - __builtin_clzll(exps[i]);
        bits = std::max(bits, top);
    }
    int rc;
    if ((rc = E.v_g1jac.reserve((total + count) * G1J + 16)) || (rc = E.v_g2jac.reserve(total * G2J + 16)) ||
        (rc = E.v_g1aff.reserve((total + count) * G1A + 16)) || (rc = E.v_g2aff.reserve(total * G2A + 16)) ||
        (rc = E.v_pairs1.reserve(2 * count * G1A)) || (rc = E.v_pairs2.reserve(2 * count * G2A)) ||
        (rc = upload(E.scalars, exps.data(), total * 32, st)) || (rc = upload(E.v_offsets, offsets.data(), (count + 1) * 4, st)) ||
        (rc = E.v_flags.reserve(count * sizeof(int) + 16)))
        return rc;
    char *g1j = E.v_g1jac.as<char>(), *g2j = E.v_g2jac.as<char>(), *g1a = E.v_g1aff.as<char>(), *g2a = E.v_g2aff.as<char>();
    char *p1 = E.v_pairs1.as<char>(), *p2 = E.v_pairs2.as<char>();
    if (total) {
        CUDA_TRY(cudaMemcpyAsync(g1j, sigs.data(), total * G1J, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(g2j, pks.data(), total * G2J, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaMemcpyAsync(g1j + total * G1J, hashes.data(), count * G1J, cudaMemcpyHostToDevice, st));
    // pairs of batch b: (bsig_b, -g2), (H_b, bpk_b)   [verify_sig, public.rs:94-120]
    CUDA_TRY(cudaMemcpy2DAsync(p2, 2 * G2A, neg.data(), G2A, G2A, count, cudaMemcpyHostToDevice, st));
    // batch_normalization_into_affine of every signature and message hash (G1) and every key (G2)
    if ((rc = batch_to_affine<G1_377>(g1j, total + count, g1a, st))) return rc;
    if ((rc = batch_to_affine<G2_377>(g2j, total, g2a, st))) return rc;
    CUDA_TRY(cudaMemcpy2DAsync(p1 + G1A, 2 * G1A, g1a + total * G1A, G1A, G1A, count, cudaMemcpyDeviceToDevice, st));
    constexpr int TH = 128;                                       // 32 quads: one point each per pass
    CUDA_TRY(cudaEventRecord(E.ev_fork, st));
    CUDA_TRY(cudaStreamWaitEvent(side, E.ev_fork, 0));
    k_small_msm_quad<Fq377, 8, TH><<<(unsigned)count, TH, (TH / 4) * sizeof(XYZZMem<Fq377>), st>>>(
        reinterpret_cast<const AffineMem<Fq377> *>(g1a), E.scalars.as<uint32_t>(), E.v_offsets.as<uint32_t>(), bits,
        reinterpret_cast<AffineMem<Fq377> *>(p1), 2, 0);
    LAUNCH_CHECK();
    k_small_msm_quad<Fp2<Fq377>, 8, TH><<<(unsigned)count, TH, (TH / 4) * sizeof(XYZZMem<Fp2<Fq377>>), side>>>(
        reinterpret_cast<const AffineMem<Fp2<Fq377>> *>(g2a), E.scalars.as<uint32_t>(), E.v_offsets.as<uint32_t>(), bits,
        reinterpret_cast<AffineMem<Fp2<Fq377>> *>(p2), 2, 1);
    LAUNCH_CHECK();
    CUDA_TRY(cudaEventRecord(E.ev_join, side));
    CUDA_TRY(cudaStreamWaitEvent(st, E.ev_join, 0));
    if ((rc = pairing_checks_2(E, p1, p2, count, E.v_flags.as<int>(), st))) return rc;
    std::vector<int> flags(count, 0);
    CUDA_TRY(cudaMemcpyAsync(flags.data(), E.v_flags.p, count * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (size_t b = 0; b < count; b++) out_verified[b] = flags[b];
    return B200_OK;
}

}  // namespace b200
