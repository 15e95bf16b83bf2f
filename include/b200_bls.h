/* b200_bls.h -- C-ABI of the B200-native MSM / multi-pairing engine.
 *
 * This is the drop-in seam for the arkworks calls celo-bls-snark-rs makes on its
 * hot path ("B1 engine seam").  Every entry point names the
 * reference call it replaces.  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - return value: 0 = ok; non-zero = error (b200_last_error() has the text).
 *     Never aborts, never falls back to a CPU path: without a usable CUDA device
 *     every compute call returns B200_ERR_CUDA.
 *   - field elements are arkworks' in-memory form: Montgomery residues, 64-bit
 *     little-endian limbs (Fp384 = 6 limbs, Fp768 = 12 limbs; Fq2 = c0 | c1).
 *   - scalars are canonical (non-Montgomery) integers, little-endian 64-bit limbs,
 *     i.e. what PrimeField::into_repr() yields (signature.rs:83, public.rs:59):
 *     4 limbs for BLS12-377 Fr, 6 limbs for BW6-761 Fr.
 *   - affine bases are `n` records of `stride` bytes: x | y [| u8 infinity | pad].
 *     Pass stride = sizeof(GroupAffine<P>) from Rust (104 for BLS12-377 G1, 200 for
 *     the others); stride == 2 * coordinate bytes means "packed, no flag" and then
 *     (0, 0) denotes the point at infinity.
 *   - results are arkworks GroupProjective: Jacobian X | Y | Z (x = X/Z^2,
 *     y = Y/Z^3, infinity <=> Z == 0), 144 bytes (BLS12-377 G1) or 288 bytes.
 *     Compare results as group elements / canonical compressed bytes, never as raw
 *     (X, Y, Z): the representative depends on the order of additions.
 *   - the caller owns all buffers; nothing is retained after a call returns.
 *   - thread-safety: calls may come from any host thread; calls on one device
 *     serialise on an internal mutex (cgo callers use arbitrary threads).
 */
#ifndef B200_BLS_H
#define B200_BLS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    B200_BLS12_377_G1 = 0, /* G1Affine / G1Projective, Fq 6 limbs, Fr 4 limbs */
    B200_BLS12_377_G2 = 1, /* G2Affine / G2Projective, Fq2 coordinates       */
    B200_BW6_761_G1 = 2,   /* Fq 12 limbs, Fr 6 limbs                         */
    B200_BW6_761_G2 = 3
} b200_curve;

enum {
    B200_OK = 0,
    B200_ERR_ARG = 1,      /* null pointer, bad curve id, n too large, bad stride */
    B200_ERR_CUDA = 2,     /* no device / CUDA runtime failure                     */
    B200_ERR_STATE = 3     /* b200_init not called                                 */
};

/* Selects the CUDA device this thread's process-wide engine uses (-1: current device)
 * and creates its stream and workspace.  Idempotent per device.
 * Reference analogue: bls-snark-sys `init()` (crates/bls-snark-sys/src/lib.rs:29-34). */
int b200_init(int device);
/* b200_init(current device) unless an engine is already bound (the reference's `init()` is optional, lib.rs:29-34). */
int b200_ensure_init(void);
/* The device the engine is bound to, or -1.  Callers that stage their own device buffers next to the engine's calls
 * (csrc/sys_compat.cu) select it first: b200_init binds the CALLING thread only. */
int b200_bound_device(void);
/* One process, several GPUs (a Rust caller of Signature::batch or ark-groth16's MSMs is one process): one engine per listed
 * device; devices[0] is the primary engine every single-device entry point keeps using.  The *_sharded entry points below
 * cut one call into a contiguous slice per GPU, one host thread per GPU, partial results (144 / 288 /
 * 576 bytes) peer-copied to the primary GPU and combined there. */
int b200_init_devices(const int *devices, int count);
int b200_device_count(void);
void b200_shutdown(void);
const char *b200_last_error(void);

/* ---- MSM: replaces VariableBaseMSM::multi_scalar_mul ------------------------------
 * Host-pointer form (copies in, computes on the GPU, copies the 144/288-byte result out).
 *   crates/bls-crypto/src/bls/signature.rs:85   -> b200_msm_bls12_377_g1
 *   crates/bls-crypto/src/bls/public.rs:61      -> b200_msm_bls12_377_g2
 *   crates/epoch-snark/src/api/prover.rs:78     -> b200_msm_bw6_761_g1 / _g2 (via ark-groth16)
 *   crates/epoch-snark/src/api/prover.rs:112    -> b200_msm_bls12_377_g1 / _g2 (inner proof)
 * Semantics follow arkworks: min(len) is the caller's job (pass one n), zero scalars
 * and infinite bases contribute nothing, bases may repeat. */
int b200_msm(int curve, const void *bases, size_t stride, const uint64_t *scalars, size_t n, void *out_jacobian);
/* the same call spread over the GPUs of b200_init_devices (one GPU: identical to b200_msm) */
int b200_msm_sharded(int curve, const void *bases, size_t stride, const uint64_t *scalars, size_t n, void *out_jacobian);
int b200_msm_bls12_377_g1(const void *bases_104B, const uint64_t *scalars, size_t n, void *out_144B);
int b200_msm_bls12_377_g2(const void *bases_200B, const uint64_t *scalars, size_t n, void *out_288B);
int b200_msm_bw6_761_g1(const void *bases_200B, const uint64_t *scalars, size_t n, void *out_288B);
int b200_msm_bw6_761_g2(const void *bases_200B, const uint64_t *scalars, size_t n, void *out_288B);

/* Device-pointer form: everything already resident in HBM, asynchronous on `stream`
 * (a cudaStream_t; NULL = the engine's own stream).  d_bases_packed holds packed arkworks
 * residues (x | y, 16-byte aligned records of 2 * coordinate bytes, (0, 0) = infinity).
 * d_out_jacobian receives the arkworks GroupProjective image. */
int b200_msm_device(int curve, const void *d_bases_packed, const void *d_scalars, size_t n, void *d_out_jacobian,
                    void *stream);

/* Fixed-base use (a Groth16 proving key is multiplied by many witnesses,
 * crates/epoch-snark/src/api/prover.rs:78): convert the bases once with
 * b200_pack_bases_device into the engine's prepared form (same size as the packed
 * records; internal Montgomery radix), then call b200_msm_prepared_device per witness.
 * src: arkworks records at `stride` bytes in host (src_on_device = 0) or device memory. */
int b200_pack_bases_device(int curve, const void *src, size_t stride, size_t n, int src_on_device,
                           void *d_dst_prepared, void *stream);
int b200_msm_prepared_device(int curve, const void *d_bases_prepared, const void *d_scalars, size_t n,
                             void *d_out_jacobian, void *stream);

/* `count` independent MSMs over one curve (prepared / packed bases, device pointers), software-pipelined
 * inside the engine: the digit sort of job i+1 and the latency-bound tail (bucket reduce, window sums,
 * Horner) of job i-1 run beside the bucket accumulation of job i, on internal streams and two workspace
 * sets.  The reference issues such runs of MSMs itself: the three G1 MSMs of one Groth16 proof
 * (crates/epoch-snark/src/api/prover.rs:78), and one G1 + one G2 MSM per batch for every batch handed to
 * batch_verify_strict (crates/bls-snark-sys/src/signatures.rs:343-404).  All jobs are ordered after the
 * work already queued on `stream`, and `stream` waits for all of them before anything queued later. */
typedef struct {
    const void *d_bases_packed;
    const void *d_scalars;
    size_t n;
    void *d_out_jacobian;
} b200_msm_job;
int b200_msm_batch_device(int curve, const b200_msm_job *jobs, size_t count, void *stream);

/* out = sum of `count` Jacobian points (device pointers).  Used to combine per-GPU
 * partial MSM results after the all-gather. */
int b200_sum_jacobian_device(int curve, const void *d_points, size_t count, void *d_out_jacobian, void *stream);
/* d_out[b] = sum_{i < count} d_points[i * batch + b] for b < batch: the combine of a whole BATCH of sharded MSMs after one
 * all-gather of rank-major records of `batch` partial points each (one launch instead of `batch`). */
int b200_sum_jacobian_batch_device(int curve, const void *d_points, size_t count, size_t batch, void *d_out_jacobian,
                                   void *stream);
/* Host-pointer form: `count` contiguous GroupProjective images -> their sum; replaces the folds of
 * PublicKey::aggregate / Signature::aggregate (crates/bls-crypto/src/bls/public.rs:34, signature.rs:52). */
int b200_sum_jacobian(int curve, const void *points, size_t count, void *out_jacobian);

/* out = scalar * base for one point, host pointers, GroupProjective images in and out (the result is normalised:
 * (x, y, 1), or zero() = (1, 1, 0)).  Replaces the scalar multiplications of PrivateKey::sign_raw / to_public
 * (crates/bls-crypto/src/bls/secret.rs:65-72). */
int b200_scalar_mul(int curve, const void *base_jacobian, const uint64_t *scalar, void *out_jacobian);

/* d_out_packed[i] = scalars[i] * base, as packed affine records (device pointers;
 * base is one packed affine record).  Synthesises benchmark / test bases on the GPU. */
int b200_fixed_base_mul_device(int curve, const void *d_base_packed, const void *d_scalars, size_t n,
                               void *d_out_packed, void *stream);

/* d_out_packed[t * run_len + j] = (start_scalars[t] + j) * base: runs of consecutive multiples (one double-and-add per
 * run).  Synthetic base arrays of 2^22 .. 2^24 distinct points for the benchmarks at ~1 / run_len of the cost above. */
int b200_point_runs_device(int curve, const void *d_base_packed, const void *d_start_scalars, size_t runs, size_t run_len,
                           void *d_out_packed, void *stream);

/* Jacobian -> affine with Montgomery batch inversion; replaces
 * ProjectiveCurve::batch_normalization_into_affine (signature.rs:82, public.rs:58).
 * d_out_packed receives packed affine records, infinity as (0, 0). */
int b200_batch_to_affine_device(int curve, const void *d_jacobian, size_t n, void *d_out_packed, void *stream);

/* ---- multi-pairing: replaces Bls12_377::product_of_pairings ---------------------------------
 *   crates/bls-crypto/src/bls/signature.rs:149  (batch_verify_hashes: (sigma, -g2), (H_i, pk_i) ...)
 *   crates/bls-crypto/src/bls/public.rs:102     (verify_sig: 2 pairs)
 * g1: n G1Affine records at stride1 bytes (104 arkworks / 96 packed), g2: n G2Affine records at
 * stride2 (200 / 192).  Pairs with an infinite member are skipped, as arkworks does.
 * out_fq12 (may be NULL) receives the GT element as arkworks' Fq12 image: 12 Fq Montgomery
 * residues in the order c0.c0.c0, c0.c0.c1, c0.c1.c0, ... c1.c2.c1 (576 bytes).
 * out_is_one (may be NULL) receives 1 iff the product of pairings equals Fq12::one() -- the
 * comparison the reference makes at signature.rs:150 / public.rs:115. */
int b200_multi_pairing_bls12_377(const void *g1, size_t stride1, const void *g2, size_t stride2, size_t n,
                                 void *out_fq12, int *out_is_one);
/* the pairs split over the GPUs of b200_init_devices; same GT element (a product of Miller values is exact) */
int b200_multi_pairing_bls12_377_sharded(const void *g1, size_t stride1, const void *g2, size_t stride2, size_t n,
                                         void *out_fq12, int *out_is_one);
/* Device-pointer halves, for pipelines and for sharding the pairs across GPUs:
 * product of the Miller values of n pairs (packed records) -> one Fq12 image; then the product
 * of `count` such images (e.g. one per GPU after an all-gather) -> final exponentiation.
 * d_is_one is a device int (may be NULL). */
int b200_miller_product_bls12_377_device(const void *d_g1_packed, const void *d_g2_packed, size_t n, void *d_out_fq12,
                                         void *stream);
int b200_final_exp_bls12_377_device(const void *d_fq12_vals, size_t count, void *d_out_fq12, int *d_is_one,
                                    void *stream);

/* ---- the two batch-verification flows of bls-crypto, after message hashing -------------------
 * Inputs are the Rust types' memory images: Signature = G1Projective (144 B), PublicKey =
 * G2Projective (288 B), message hashes = G1Projective (what HashToCurve::hash returns).
 * *out_verified = 1 iff the reference would return Ok(()).
 *
 * b200_batch_verify_hashes: Signature::batch_verify_hashes (signature.rs:125-155) --
 *   into_affine() of all points, pairs (sigma, -g2), (H_i, pk_i), product == Fq12::one().
 *   (Length mismatch is the caller's UnevenNumKeysMessages check, signature.rs:130-132.)
 * b200_batch_verify_strict_hash: Batch::verify (batch.rs:44-84) with the exponents supplied by
 *   the caller (the reference draws exp_size random bytes per entry, batch.rs:51-65, and calls
 *   into_repr(): pass those canonical 4 x u64 values): batch_normalization_into_affine of the
 *   keys and signatures, G2 MSM (public.rs:61), G1 MSM (signature.rs:85), then verify_sig
 *   (public.rs:94-120) on the message hash. */
int b200_batch_verify_hashes(const void *signature, const void *pubkeys, const void *message_hashes, size_t n,
                             int *out_verified);
int b200_batch_verify_strict_hash(const void *pubkeys, const void *signatures, const uint64_t *exponents, size_t n,
                                  const void *message_hash, int *out_verified);
/* Batch::verify for `count` batches in one pass -- what batch_verify_strict is called with
 * (crates/bls-snark-sys/src/signatures.rs:343-404; the reference's benchmark: 300 batches of 20 signatures,
 * crates/bls-crypto/benches/batch_bls.rs:62-95).  Every batch's two MSMs, 2-pair Miller loops and final exponentiation
 * run side by side on the device; out_verified[b] = 1 iff Batch::verify of batch b would return Ok(()). */
typedef struct {
    const void *pubkeys;       /* n x 288 bytes, G2Projective images */
    const void *signatures;    /* n x 144 bytes, G1Projective images */
    const uint64_t *exponents; /* n x 4 limbs, canonical */
    size_t n;
    const void *message_hash;  /* 144 bytes, what HashToCurve::hash returned for the batch's message */
} b200_strict_batch;
int b200_batch_verify_strict_many(const b200_strict_batch *batches, size_t count, int *out_verified);

/* ---- batched hash-to-G1 ------------------------------------------------------------------------------
 * Replaces HashToCurve::hash for BLS12-377 G1 -- the step before the multi-pairing in
 * PublicKey::verify / Signature::batch_verify (crates/bls-crypto/src/bls/signature.rs:111-114):
 *   B200_HASHER_DIRECT     DIRECT_HASH_TO_G1    (hash_to_curve/try_and_increment.rs:36-38; hashers/direct.rs:20-80)
 *   B200_HASHER_COMPOSITE  COMPOSITE_HASH_TO_G1 (try_and_increment.rs:29-31; hashers/composite.rs:15-95: Bowe-Hopwood
 *                          CRH over ed-on-bw6-761 with the reference's ChaCha20-seeded generators, then the XOF)
 * flags: B200_HASH_COMPAT = the `compat` cargo feature's sign-bit rule (the reference's default,
 *   try_and_increment.rs:103-117); B200_HASH_CIP22 = TryAndIncrementCIP22 (try_and_increment_cip22.rs:60-134);
 *   B200_HASH_CRH_ONLY = Hasher::crh of each message alone: out receives n x 48 bytes (the composite CRH's 48-byte
 *   x coordinate, or the direct CRH's 32 bytes followed by zeros) instead of points.
 * domain: at most 8 bytes (BLSError::DomainTooLarge otherwise -> B200_ERR_ARG).  Each input is hashed as
 * counter | extra_data | message for counter = 0..254 until a candidate decodes to a point whose cofactor multiple
 * is not zero; out_jacobian receives n G1Projective memory images (144 B, what HashToCurve::hash returns),
 * out_attempts (may be NULL) the successful counters (hash_with_attempt).  All n inputs run in one launch,
 * one warp per input.  An input larger than the CRH's 156 240 bits, or one with no point in 255 attempts
 * (BLSError::HashToCurveError), fails the call with B200_ERR_ARG. */
enum { B200_HASHER_DIRECT = 0, B200_HASHER_COMPOSITE = 1 };
enum { B200_HASH_COMPAT = 1, B200_HASH_CIP22 = 2, B200_HASH_CRH_ONLY = 4 };
typedef struct {
    const uint8_t *message;
    size_t message_len;
    const uint8_t *extra_data;
    size_t extra_data_len;
} b200_hash_input;
int b200_hash_to_g1(int hasher, int flags, const uint8_t *domain, size_t domain_len, const b200_hash_input *inputs, size_t n,
                    void *out_jacobian, uint32_t *out_attempts);

/* ---- radix-2 NTT and the Groth16 witness map ------------------------------------------------------
 * Replaces ark-poly 0.1.0 Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place and the
 * transform chain of ark-groth16 0.1.0 R1CStoQAP::witness_map, which create_proof_no_zk runs right
 * before its MSMs (crates/epoch-snark/src/api/prover.rs:78 outer proof over BW6-761, :112 inner proof
 * over BLS12-377).  Elements are the scalar fields' arkworks memory images (Montgomery residues):
 * B200_FR_BLS12_377 = Fp256, 4 x u64 (32 B); B200_FR_BW6_761 = Fp384, 6 x u64 (48 B; this field is
 * BLS12-377's Fq).  The domain has n = 2^log_n points, group generator
 * TWO_ADIC_ROOT_OF_UNITY^(2^(s - log_n)) and coset offset GENERATOR (22 / -5), as in arkworks.
 * b200_ntt_device: in place, natural order in and out; inverse includes the 1/n scaling; coset
 *   variants multiply by GENERATOR^i before the forward / by GENERATOR^-i after the inverse transform.
 * b200_witness_map_device: d_a, d_b, d_c hold the n evaluations of the A, B, C combinations (c_i =
 *   a_i * b_i on constraint rows); they are overwritten.  d_h receives the n coefficients of
 *   (a b - c) / Z; the prover's h-query MSM consumes the first n - 1. */
enum { B200_FR_BLS12_377 = 0, B200_FR_BW6_761 = 1 };
int b200_ntt_device(int field, void *d_data, unsigned log_n, int inverse, int coset, void *stream);
int b200_witness_map_device(int field, void *d_a, void *d_b, void *d_c, unsigned log_n, void *d_h, void *stream);

/* ---- Groth16 prover arithmetic --------------------------------------------------------------------
 * Everything ark-groth16 0.1.0 create_proof_no_zk (= create_proof_with_reduction_no_zk, r = s = 0)
 * computes after constraint synthesis, for crates/epoch-snark/src/api/prover.rs:78 (family
 * B200_GROTH16_BW6_761: G1/G2 over BW6-761, scalars 6 x u64) and :112 (B200_GROTH16_BLS12_377).
 * Proving-key queries are device arrays of PACKED affine records (b200_pack_bases_device output),
 * resident across proofs:  a_query, b_g2_query: num_assign + 1 records (entry 0 is the constant-one
 * column), l_query: num_aux, h_query: 2^log_n - 1, alpha_g1 / beta_g2: one record each.
 * d_assignment: num_assign canonical scalars = public inputs (without the leading one) followed by the
 * num_aux auxiliary values (what arkworks passes to its MSMs after into_repr()).
 * d_a, d_b, d_c: the 2^log_n evaluation vectors of the witness map (Montgomery images; overwritten).
 * d_proof receives A | B | C as arkworks GroupProjective images (G1, G2, G1); the caller's
 * into_affine() of those is the Proof { a, b, c } the reference serialises. */
enum { B200_GROTH16_BLS12_377 = 0, B200_GROTH16_BW6_761 = 1 };
typedef struct {
    const void *a_query;
    const void *b_g2_query;
    const void *h_query;
    const void *l_query;
    const void *alpha_g1;
    const void *beta_g2;
} b200_groth16_pk;
int b200_groth16_prove_device(int family, const b200_groth16_pk *pk, const void *d_assignment, size_t num_assign,
                              size_t num_aux, void *d_a, void *d_b, void *d_c, unsigned log_n, void *d_proof,
                              void *stream);
/* The same proof split over `shards` GPUs (BASELINE config 5: "8 x B200"): every GPU runs the witness map and its
 * contiguous share of each of the four MSMs -> d_partials = a_acc | l_acc | h_acc | b_acc (3 G1 + 1 G2
 * GroupProjective images: 3 * 144 + 288 or 4 * 288 bytes); the single exchange is an all-gather of those records;
 * b200_groth16_assemble_device folds `shards` records (rank-major) into A | B | C.  shards = 1 is the call above. */
int b200_groth16_prove_partial_device(int family, const b200_groth16_pk *pk, const void *d_assignment, size_t num_assign,
                                      size_t num_aux, void *d_a, void *d_b, void *d_c, unsigned log_n, unsigned shard,
                                      unsigned shards, void *d_partials, void *stream);
int b200_groth16_assemble_device(int family, const b200_groth16_pk *pk, const void *d_partials, unsigned shards, void *d_proof,
                                 void *stream);

/* ---- BW6-761 product of pairings and Groth16 verification -------------------------------------------
 * Replaces BW6_761::product_of_pairings as reached from ark-groth16 0.1.0 verify_proof at
 * crates/epoch-snark/src/api/verifier.rs:35 (C-ABI caller: `verify`, crates/bls-snark-sys/src/snark/mod.rs:23-45).
 * G1 (y^2 = x^3 - 1) and G2 (y^2 = x^3 + 4) are both over Fq761: records are x | y [| flag] at `stride`
 * bytes (200 arkworks GroupAffine / 192 packed), Montgomery residues, 12 x u64 per coordinate.
 * The pairing is the optimal ate pairing e(P, Q) = (f_{u+1,Q}(P) f_{u^3-u^2-u,Q}(P)^q)^((q^6-1)/r).
 * out_fq6 (may be NULL) receives the value as arkworks' Fq6 image (c0.c0, c0.c1, c0.c2, c1.c0, c1.c1,
 * c1.c2; 576 bytes) for the exponent c (q^6 - 1) / r with a fixed 191-bit c prime to r (the hard part is
 * computed as f^R0(u) (f^q)^R1(u), csrc/pairing_bw6.cuh).  Upstream's addition chain raises to another fixed
 * multiple; the multiple changes the representative of GT but not any comparison with one, the only thing
 * verify_proof observes.  out_is_one (may be NULL): 1 iff the product of pairings is one. */
int b200_multi_pairing_bw6_761(const void *g1, size_t stride1, const void *g2, size_t stride2, size_t n,
                               void *out_fq6, int *out_is_one);
/* Device-pointer halves (packed 192-byte records): d_out_vals receives 2 n Fq6 images (power-basis
 * coefficient order, an engine-internal layout), whose product is the Miller value of the n pairs; the
 * final exponentiation multiplies `count` such images first (pairs may be split across streams or GPUs). */
int b200_miller_values_bw6_761_device(const void *d_g1_packed, const void *d_g2_packed, size_t n, void *d_out_vals,
                                      void *stream);
int b200_final_exp_bw6_761_device(const void *d_vals, size_t count, void *d_out_fq6, int *d_is_one, void *stream);

/* ark-groth16 verify_proof over BW6-761 (host pointers, arkworks memory images):
 *   g_ic = gamma_abc_g1[0] + sum_i public_inputs[i] * gamma_abc_g1[i + 1]
 *   e(A, B) * e(g_ic, -gamma_g2) * e(C, -delta_g2) == e(alpha_g1, beta_g2)
 * The struct mirrors VerifyingKey<BW6_761>: single points are one record each, gamma_abc_g1 is
 * num_gamma_abc records at `stride` bytes.  public_inputs: canonical scalars, 6 x u64 each (what
 * into_repr() yields); num_inputs + 1 != num_gamma_abc is upstream's MalformedVerifyingKey error and
 * returns B200_ERR_ARG.  *out_verified = 1 iff verify_proof would return Ok(true). */
typedef struct {
    const void *alpha_g1;
    const void *beta_g2;
    const void *gamma_g2;
    const void *delta_g2;
    const void *gamma_abc_g1;
    size_t num_gamma_abc;
    size_t stride;
} b200_groth16_vk;
int b200_groth16_verify_bw6_761(const b200_groth16_vk *vk, const void *proof_a, const void *proof_b, const void *proof_c,
                                const uint64_t *public_inputs, size_t num_inputs, int *out_verified);

/* ---- point decoding: replaces ark-serialize 0.1.0 GroupAffine::deserialize (compressed form) ------------
 *   crates/bls-snark-sys/src/snark/epoch_block.rs:154-196 (read_slice::<VerifyingKey / Proof>, read_pubkeys)
 * kind: 0 = BLS12-377 G2 (96 B: x.c0 | x.c1), 1 = BW6-761 G1, 2 = BW6-761 G2 (96 B: x); flags in the top two
 * bits of the last byte (bit 7: y is the larger root, bit 6: infinity).  Host pointers.  out_packed (may be
 * NULL): n packed affine records (192 B, Montgomery, (0, 0) for infinity or a rejected point); out_status: n
 * ints, 0 ok, 1 infinity, 2 coordinate >= modulus, 3 x not on the curve, 4 not in the prime-order subgroup
 * (only tested when check_subgroup != 0, as deserialize does; deserialize_unchecked does not). */
enum { B200_POINTS_BLS12_377_G2 = 0, B200_POINTS_BW6_761_G1 = 1, B200_POINTS_BW6_761_G2 = 2, B200_POINTS_BLS12_377_G1 = 3 };
int b200_deserialize_points(int kind, const void *bytes, size_t n, int check_subgroup, void *out_packed, int *out_status);
/* B200_POINTS_BLS12_377_G1 (Signature::deserialize): 48-byte encodings in, 96-byte packed affine records out.
 * b200_serialize_points: the inverse for the two BLS12-377 groups -- n GroupProjective memory images (kind 3: 144 B,
 * kind 0: 288 B; Signature / PublicKey) -> into_affine().serialize(): 48 / 96 bytes each, infinity as the flag byte.
 *   crates/bls-crypto/src/bls/public.rs:123-135, signature.rs (CanonicalSerialize for PublicKey / Signature) */
int b200_serialize_points(int kind, const void *jacobian_images, size_t n, void *out_bytes);

/* The reference's SNARK verifier entry point with a status code: same arguments as `verify`
 * (include/bls_snark_sys_compat.h, which this library also exports under that name), blocks passed by
 * pointer (const EpochBlockFFI *).  *out_ok = 1 iff the reference returns true; the return value is non-zero
 * only for engine failures (no CUDA device, runtime error) -- malformed inputs are *out_ok = 0. */
int b200_verify_epochs(const uint8_t *vk, size_t vk_len, const uint8_t *proof, size_t proof_len, const void *first_epoch,
                       const void *last_epoch, int *out_ok);

/* The public inputs epoch_snark::verify derives from the two blocks (verifier.rs:30-33): pack(hash(first) |
 * hash(last with aggregated key)), canonical scalars of 6 x u64; two of them for the Blake2s edge hashes.
 * Keys are decoded, checked and aggregated on the device; *out_ok = 0 when a block does not decode. */
int b200_epoch_public_inputs(const void *first_epoch, const void *last_epoch, uint64_t *out_inputs, size_t capacity,
                             size_t *out_count, int *out_ok);

/* Host helper of the verifier (no GPU): Blake2s, 32-byte digest, 8-byte personalisation
 * (crates/epoch-snark/src/epoch_block.rs:226-236 hashes the encoded blocks with "ULforout"). */
void b200_blake2s_personal(const uint8_t *data, size_t len, const uint8_t *personal8, uint8_t *out32);
/* The same with the full parameter block (digest length, fanout, depth, leaf length, node offset, inner length): the
 * DirectHasher's CRH / XOF parameters (crates/bls-crypto/src/hashers/direct.rs:23-79). */
void b200_blake2s_param(const uint8_t *data, size_t len, int digest_len, int fanout, int depth, uint32_t leaf_len, uint64_t node_offset,
                        int inner_len, const uint8_t *personal8, uint8_t *out);
/* Host helper (no GPU): EpochBlock::encode_to_bytes (cip22 = 0; round, entropies and maximum_validators unused) or
 * ::encode_inner_to_bytes_cip22 (cip22 = 1; entropies may be NULL, keys padded with the G2 generator up to
 * maximum_validators) from the keys' 96-byte compressed encodings.
 *   crates/epoch-snark/src/epoch_block.rs:106-114, 152-171, 191-211; crates/epoch-snark/src/encoding.rs:23-80
 * out_inner / out_extra: malloc'd, the caller frees; out_extra may be NULL when cip22 = 0. */
int b200_encode_epoch_block(int cip22, uint16_t index, uint8_t round, const uint8_t *epoch_entropy, const uint8_t *parent_entropy,
                            uint32_t maximum_non_signers, size_t maximum_validators, const uint8_t *keys96, size_t nkeys,
                            uint8_t **out_inner, size_t *out_inner_len, uint8_t **out_extra, size_t *out_extra_len);

/* Element-wise arithmetic in the coordinate field of `curve` (Fq, Fq2 or Fq761; Montgomery
 * form, device pointers): op 0 add, 1 sub, 2 mul, 3 square(a), 4 inverse(a), 5 neg(a),
 * 6 double(a).  Exists so the field layer can be checked against the oracle directly. */
int b200_field_op_device(int curve, int op, const void *d_a, const void *d_b, size_t n, void *d_out, void *stream);

/* Blocks until everything queued on the engine's stream (or `stream`) has finished. */
int b200_sync(void *stream);

/* Introspection for tests and bench: the window plan the engine picks for n. */
int b200_msm_plan(int curve, size_t n, int *window_bits, int *windows, uint32_t *buckets_per_window);
/* Timing of the dominant kernel (bucket accumulation) with CUDA events on the launching
 * stream: enable, run MSMs, then read the summed duration, launch count and pairs covered
 * (resets the counters; at most 256 launches are recorded between reads). */
int b200_profile_enable(int on);
int b200_profile_read(double *accumulate_ms, int *launches, uint64_t *pairs);
/* Number of kernel launches the engine has issued since b200_init. */
uint64_t b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200_BLS_H */
