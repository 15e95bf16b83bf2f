// Callers of the hot path, composed on the device: the two batch-verification flows of
// bls-crypto after message hashing (hash-to-curve itself is out of scope).
//   * Signature::batch_verify_hashes   crates/bls-crypto/src/bls/signature.rs:125-155
//   * Batch::verify                     crates/bls-crypto/src/bls/batch.rs:44-84
//     (PublicKey::batch public.rs:47-65, Signature::batch signature.rs:70-89,
//      PublicKey::verify_sig public.rs:94-120)
// Inputs are the Rust types' memory images: Signature = G1Projective (144 B),
// PublicKey = G2Projective (288 B), exponents after into_repr() (4 x u64, canonical).
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

}  // namespace b200
