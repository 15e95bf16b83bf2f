/* bls_snark_sys_compat.h -- the part of bls-snark-sys' own C-ABI that libb200bls.so re-exports under the
 * reference's names and conventions (seam "B2"), so that an existing cgo / C consumer
 * of bls-snark-sys can link the CUDA engine without source changes for these calls.
 *
 * Conventions kept from the reference (crates/bls-snark-sys/src/lib.rs:21-27): every function returns `bool`
 * success; any error (malformed bytes, a point off the curve or outside the prime-order subgroup, a failed
 * pairing check, no usable CUDA device) is logged to stderr and turned into `false`.
 */
#ifndef BLS_SNARK_SYS_COMPAT_H
#define BLS_SNARK_SYS_COMPAT_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* crates/bls-snark-sys/src/snark/epoch_block.rs:109-127 (#[repr(C)], 56 bytes, passed BY VALUE to verify).
 * epoch_entropy / parent_entropy: 16 bytes each or NULL (= None, epoch_block.rs:204-210);
 * pubkeys: pubkeys_num x 96 bytes, compressed BLS12-377 G2 (epoch_block.rs:168-196). */
typedef struct {
    uint16_t index;
    uint8_t round;
    const uint8_t *epoch_entropy;
    const uint8_t *parent_entropy;
    const uint8_t *pubkeys;
    size_t pubkeys_num;
    uint32_t maximum_non_signers;
    size_t maximum_validators;
} EpochBlockFFI;

/* crates/bls-snark-sys/src/snark/mod.rs:23-45.  vk: arkworks-compressed VerifyingKey<BW6_761>
 * (alpha_g1 | beta_g2 | gamma_g2 | delta_g2 | u64 len | gamma_abc_g1[len], 96 bytes per point), proof: A | B | C
 * (288 bytes).  Returns true iff epoch_snark::verify (crates/epoch-snark/src/api/verifier.rs:23-40) returns Ok:
 * both epoch blocks decode (every key on the curve and in the subgroup), the key and the proof decode, and
 * e(A, B) e(g_ic, -gamma) e(C, -delta) == e(alpha, beta) for the public inputs pack(hash(first) | hash(last)).
 * Binds the engine to the current CUDA device on first use (the reference's `init()` is optional too). */
bool verify(const uint8_t *vk, uint32_t vk_len, const uint8_t *proof, uint32_t proof_len, EpochBlockFFI first_epoch,
            EpochBlockFFI last_epoch);

/* ---- signature verification (crates/bls-snark-sys/src/signatures.rs, serialization.rs, utils.rs) ---------------------
 * Handles are heap objects holding the Rust types' memory images, as in the reference (Box::into_raw):
 * PublicKey = G2Projective (288 bytes), Signature = G1Projective (144 bytes), Montgomery limbs.  They are created by
 * deserialize_* / aggregate_* and released by destroy_*; byte buffers returned through out-pointers are released by
 * free_vec.  A handle made by the reference's library has the same layout and can be passed in as it is. */
typedef struct PublicKey PublicKey;
typedef struct Signature Signature;
typedef struct PrivateKey PrivateKey;          /* Fr, 32 bytes (Montgomery residue) */

/* utils.rs:75-82 */
typedef struct {
    const uint8_t *ptr;
    size_t len;
} Buffer;
/* utils.rs:20-32 (48 bytes) */
typedef struct {
    Buffer data;
    Buffer extra;
    const PublicKey *public_key;
    const Signature *sig;
} MessageFFI;
/* utils.rs:58-72 (64 bytes) */
typedef struct {
    Buffer data;
    Buffer extra;
    const PublicKey *const *public_keys;
    size_t public_keys_len;
    const Signature *const *signatures;
    size_t signatures_len;
} BatchMessageFFI;

/* serialization.rs:35-43, 81-88: ark-serialize compressed encodings (96 / 48 bytes) with every check of
 * G2Affine / G1Affine::deserialize (coordinate < modulus, on the curve, prime-order subgroup) -> a new handle. */
bool deserialize_public_key(const uint8_t *in_public_key_bytes, int in_public_key_bytes_len, PublicKey **out_public_key);
bool deserialize_signature(const uint8_t *in_signature_bytes, int in_signature_bytes_len, Signature **out_signature);
/* serialization.rs:63-70, 90-97: into_affine().serialize() -> a new 96 / 48-byte buffer (free_vec). */
bool serialize_public_key(const PublicKey *in_public_key, uint8_t **out_bytes, int *out_len);
bool serialize_signature(const Signature *in_signature, uint8_t **out_bytes, int *out_len);
/* serialization.rs:166-215: uncompressed x | y (96 / 192 bytes, canonical little-endian coordinates) -> the
 * compressed encoding (48 / 96 bytes, free_vec).  Integer comparisons only; runs on the host. */
bool compress_signature(const uint8_t *in_signature, int in_signature_len, uint8_t **out_signature, int *out_len);
bool compress_pubkey(const uint8_t *in_pubkey, int in_pubkey_len, uint8_t **out_pubkey, int *out_len);
/* ---- csrc/sys_compat_keys.cu: private keys, signing, hash helpers, uncompressed encodings ---------------------
 * lib.rs:29-34 */
bool init(void);
/* signatures.rs:19-42; serialization.rs:13-33, 224-234.  generate_private_key draws Fr::rand from the OS entropy source
 * (the reference: rand::thread_rng()); private_key_to_public_key is sk * g2 on the device. */
bool generate_private_key(PrivateKey **out_private_key);
bool deserialize_private_key(const uint8_t *in_private_key_bytes, int in_private_key_bytes_len, PrivateKey **out_private_key);
bool serialize_private_key(const PrivateKey *in_private_key, uint8_t **out_bytes, int *out_len);
bool destroy_private_key(PrivateKey *private_key);
bool private_key_to_public_key(const PrivateKey *in_private_key, PublicKey **out_public_key);
/* signatures.rs:44-91: hash-to-G1 on the device (same hasher selection as verify_signature), then hash * sk on the device */
bool sign_message(const PrivateKey *in_private_key, const uint8_t *in_message, int in_message_len, const uint8_t *in_extra_data,
                  int in_extra_data_len, bool should_use_composite, bool should_use_cip22, Signature **out_signature);
bool sign_pop(const PrivateKey *in_private_key, const uint8_t *in_message, int in_message_len, Signature **out_signature);
/* signatures.rs:93-242.  hash_direct*: G1Affine::write = canonical x | y | infinity byte (97 bytes).  hash_composite*:
 * G1Projective::write = canonical X | Y | Z (144 bytes) of the NORMALISED representative (x, y, 1) -- the reference writes
 * the representative its cofactor multiplication ended on; the point is the same.  hash_crh: the 48-byte composite CRH. */
bool hash_direct(const uint8_t *in_message, int in_message_len, uint8_t **out_hash, int *out_len, bool use_pop);
bool hash_direct_with_attempt(const uint8_t *in_message, int in_message_len, uint8_t **out_hash, int *out_len, int *out_attempt,
                              bool use_pop);
bool hash_composite(const uint8_t *in_message, int in_message_len, const uint8_t *in_extra_data, int in_extra_data_len, uint8_t **out_hash,
                    int *out_len);
bool hash_composite_cip22(const uint8_t *in_message, int in_message_len, const uint8_t *in_extra_data, int in_extra_data_len,
                          uint8_t **out_hash, int *out_len, uint8_t *attempt_counter);
bool hash_crh(const uint8_t *in_message, int in_message_len, int hash_bytes, uint8_t **out_hash, int *out_len);
/* signatures.rs:191-212: DirectHasher.hash(SIG_DOMAIN, message, hash_bytes) = XOF(CRH(message)), hash_bytes bytes */
bool hash_direct_first_step(const uint8_t *in_message, int in_message_len, int hash_bytes, uint8_t **out_hash, int *out_len);
/* snark/epoch_block.rs:16-106: EpochBlock::encode_inner_to_bytes_cip22 (inner bytes + extra-data bytes; entropies are 16 bytes
 * or NULL; keys padded with the G2 generator up to in_maximum_validators) and the pre-Donut EpochBlock::encode_to_bytes */
bool encode_epoch_block_to_bytes_cip22(unsigned short in_epoch_index, unsigned char in_round_number, const uint8_t *in_epoch_entropy,
                                       const uint8_t *in_parent_entropy, unsigned int in_maximum_non_signers, unsigned int in_maximum_validators,
                                       const PublicKey *const *in_added_public_keys, int in_added_public_keys_len, uint8_t **out_bytes,
                                       int *out_len, uint8_t **out_extra_data_bytes, int *out_extra_data_len);
bool encode_epoch_block_to_bytes(unsigned short in_epoch_index, unsigned int in_maximum_non_signers, const PublicKey *const *in_added_public_keys,
                                 int in_added_public_keys_len, uint8_t **out_bytes, int *out_len);
/* serialization.rs:44-61, 72-105: the cached decoder hands out the same key; serialize_uncompressed = canonical x | y
 * (192 / 96 bytes), infinity flag in bit 6 of the last byte */
bool deserialize_public_key_cached(const uint8_t *in_public_key_bytes, int in_public_key_bytes_len, PublicKey **out_public_key);
bool serialize_public_key_uncompressed(const PublicKey *in_public_key, uint8_t **out_bytes, int *out_len);
bool serialize_signature_uncompressed(const Signature *in_signature, uint8_t **out_bytes, int *out_len);
/* signatures.rs:454-483: aggregated - sum(keys) */
bool aggregate_public_keys_subtract(const PublicKey *in_aggregated_public_key, const PublicKey *const *in_public_keys, int in_public_keys_len,
                                    PublicKey **out_public_key);

/* serialization.rs:236-266 */
bool free_vec(uint8_t *bytes, int len);
bool destroy_public_key(PublicKey *public_key);
bool destroy_signature(Signature *signature);
/* signatures.rs:428-451, 485-505: group sums (PublicKey::aggregate / Signature::aggregate).  The reference's
 * aggregate_public_keys also consults a process-wide cache (cache.rs); the sum is the same. */
bool aggregate_public_keys(const PublicKey *const *in_public_keys, int in_public_keys_len, PublicKey **out_public_key);
bool aggregate_signatures(const Signature *const *in_signatures, int in_signatures_len, Signature **out_signature);
/* signatures.rs:244-276: PublicKey::verify over SIG_DOMAIN with the selected hash-to-G1:
 * (composite, cip22) = (true, true) COMPOSITE_HASH_TO_G1_CIP22, (true, false) COMPOSITE_HASH_TO_G1,
 * (false, false) DIRECT_HASH_TO_G1, (false, true) an error (returns false). */
bool verify_signature(const PublicKey *in_public_key, const uint8_t *in_message, int in_message_len, const uint8_t *in_extra_data,
                      int in_extra_data_len, const Signature *in_signature, bool should_use_composite, bool should_use_cip22,
                      bool *out_verified);
/* signatures.rs:407-425: PublicKey::verify_pop (POP_DOMAIN, DIRECT_HASH_TO_G1, no extra data). */
bool verify_pop(const PublicKey *in_public_key, const uint8_t *in_message, int in_message_len, const Signature *in_signature,
                bool *out_verified);
/* signatures.rs:290-333: aggregates the messages' signatures and runs Signature::batch_verify --
 * e(asig, -g2) * prod e(H(data_i, extra_i), pk_i) == 1; all messages are hashed in one launch. */
bool batch_verify_signature(const MessageFFI *messages_ptr, size_t messages_len, bool should_use_composite, bool should_use_cip22,
                            bool *verified);
/* signatures.rs:343-404: per batch Batch::verify (crates/bls-crypto/src/bls/batch.rs:44-84) with fresh random exponents
 * of byte_count_from_target_batch_size bytes; out_results[i] per batch; returns false unless every batch verified. */
bool batch_verify_strict(const BatchMessageFFI *in_batches_ptr, size_t in_batches_len, bool should_use_composite, bool should_use_cip22,
                         bool *out_results);

#ifdef __cplusplus
}
#endif
#endif /* BLS_SNARK_SYS_COMPAT_H */
