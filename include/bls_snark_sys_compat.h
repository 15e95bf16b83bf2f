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

#ifdef __cplusplus
}
#endif
#endif /* BLS_SNARK_SYS_COMPAT_H */
