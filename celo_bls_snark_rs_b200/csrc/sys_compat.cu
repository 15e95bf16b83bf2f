// The signature-verification part of bls-snark-sys' own C-ABI, re-exported under the reference's names and
// conventions over the CUDA engine (SURVEY.md section 8b seam "B2", row f2), so that a cgo / C consumer of
// libbls_snark_sys can link libb200bls for these calls without source changes:
//
//   crates/bls-snark-sys/src/signatures.rs:244-276   verify_signature
//   crates/bls-snark-sys/src/signatures.rs:290-333   batch_verify_signature   (Signature::batch_verify)
//   crates/bls-snark-sys/src/signatures.rs:343-404   batch_verify_strict      (Batch::verify, batch.rs:44-84)
//   crates/bls-snark-sys/src/signatures.rs:407-425   verify_pop
//   crates/bls-snark-sys/src/signatures.rs:428-451,485-505   aggregate_public_keys / aggregate_signatures
//   crates/bls-snark-sys/src/serialization.rs:35-105 (de)serialize_public_key / _signature
//   crates/bls-snark-sys/src/serialization.rs:224-266 destroy_public_key / destroy_signature / free_vec
//   crates/bls-snark-sys/src/utils.rs:20-82          MessageFFI, BatchMessageFFI, Buffer
//
// Conventions kept (lib.rs:21-27): every function returns `bool` success, an error is logged to stderr and becomes
// `false`, results go through out-pointers, `verified` is a separate out-bool.  Handles are what the reference's are
// -- heap objects holding the Rust types' memory images (PublicKey = G2Projective, 288 B; Signature = G1Projective,
// 144 B) -- allocated here with malloc and released by destroy_*; byte buffers by free_vec.
//
// This file is host glue only: every field / curve operation (point decoding and its checks, hashing to G1, sums,
// the MSMs of the strict batch, the pairings, the encodings) runs on the device through the b200_* entry points.
#include <cuda_runtime.h>
#include <sys/random.h>

#include <cerrno>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>
#include <vector>

#include "../../include/b200_bls.h"
#include "../../include/bls_snark_sys_compat.h"

namespace {

constexpr size_t PK_BYTES = 288, SIG_BYTES = 144;
const uint8_t SIG_DOMAIN[8] = {'U', 'L', 'f', 'o', 'r', 'x', 'o', 'f'};      // crates/bls-crypto/src/lib.rs:75
const uint8_t POP_DOMAIN[8] = {'U', 'L', 'f', 'o', 'r', 'p', 'o', 'p'};      // lib.rs:78
// Montgomery form of 1 in BLS12-377 Fq (the z coordinate of into_projective()), little-endian u64 limbs
const uint64_t FQ_ONE[6] = {0x02cdffffffffff68ull, 0x51409f837fffffb1ull, 0x9f7db3a98a7d3ff2ull,
                            0x7b4e97b76e7c6305ull, 0x4cf495bf803c84e8ull, 0x008d6661e2fdf49aull};

bool failed(const char *fn, const char *why) {
    fprintf(stderr, "[b200] %s -> false: %s\n", fn, why);
    return false;
}
bool engine_failed(const char *fn) { return failed(fn, b200_last_error()); }

// hasher selection of the reference: (composite, cip22); (false, true) is BLSError::HashToCurveError
bool hasher_of(bool composite, bool cip22, int *hasher, int *flags) {
    if (!composite && cip22) return false;
    *hasher = composite ? B200_HASHER_COMPOSITE : B200_HASHER_DIRECT;
    *flags = B200_HASH_COMPAT | (cip22 ? B200_HASH_CIP22 : 0);               // `compat` is the reference's default feature
    return true;
}

// sum of Jacobian images on the device (PublicKey::aggregate / Signature::aggregate): gathered into one host block,
// then b200_sum_jacobian stages, sums and reads back on the engine's own stream under the engine's lock
bool sum_images(int curve, const std::vector<const void *> &images, size_t bytes, void *out) {
    const size_t n = images.size();
    std::vector<uint8_t> host(n * bytes + 16);
    for (size_t i = 0; i < n; i++) memcpy(&host[i * bytes], images[i], bytes);
    return b200_sum_jacobian(curve, host.data(), n, out) == B200_OK;
}

// len bytes from the operating system's CSPRNG (getrandom(2)); std::random_device as the fallback
bool os_random(uint8_t *dst, size_t len) {
    size_t at = 0;
    while (at < len) {
        const ssize_t got = getrandom(dst + at, len - at, 0);
        if (got < 0) {
            if (errno == EINTR) continue;
            break;
        }
        at += (size_t)got;
    }
    if (at < len) {
        try {
            std::random_device rd;
            for (; at < len; at++) dst[at] = (uint8_t)rd();
        } catch (...) {
            return false;
        }
    }
    return true;
}

// ark_std::log2: ceil(log2(x)), 0 for x <= 1
size_t log2_ceil(size_t x) {
    size_t l = 0;
    while (((size_t)1 << l) < x) l++;
    return l;
}

template <class T>
bool deserialize_point(const char *fn, int kind, size_t enc, const uint8_t *bytes, int len, T **out) {
    if (!bytes || !out) return failed(fn, "null pointer");
    if (len < (int)enc) return failed(fn, "not enough bytes");               // SerializationError (io: unexpected end)
    if (b200_ensure_init() != B200_OK) return engine_failed(fn);
    uint8_t packed[192];
    int status = -1;
    if (b200_deserialize_points(kind, bytes, 1, 1, packed, &status) != B200_OK) return engine_failed(fn);
    if (status > 1) return failed(fn, status == 2 ? "coordinate not below the modulus" : status == 3 ? "x is not on the curve"
                                                                                                   : "point not in the prime-order subgroup");
    uint8_t *image = (uint8_t *)calloc(1, 3 * enc);
    if (!image) return failed(fn, "out of memory");
    if (status == 1) {                                                       // GroupProjective::zero() = (1, 1, 0)
        memcpy(image, FQ_ONE, 48);
        memcpy(image + enc, FQ_ONE, 48);
    } else {                                                                 // into_projective(): (x, y, 1)
        memcpy(image, packed, 2 * enc);
        memcpy(image + 2 * enc, FQ_ONE, 48);
    }
    *out = reinterpret_cast<T *>(image);
    return true;
}

template <class T>
bool serialize_point(const char *fn, int kind, size_t enc, const T *in, uint8_t **out_bytes, int *out_len) {
    if (!in || !out_bytes || !out_len) return failed(fn, "null pointer");
    if (b200_ensure_init() != B200_OK) return engine_failed(fn);
    uint8_t *buf = (uint8_t *)malloc(enc);
    if (!buf) return failed(fn, "out of memory");
    if (b200_serialize_points(kind, in, 1, buf) != B200_OK) {
        free(buf);
        return engine_failed(fn);
    }
    *out_bytes = buf;
    *out_len = (int)enc;
    return true;
}

// PublicKey::verify_sig (public.rs:94-120) after hashing on the device
bool verify_one(const char *fn, const void *pk, const uint8_t *domain, const uint8_t *msg, size_t msg_len, const uint8_t *extra,
                size_t extra_len, const void *sig, int hasher, int flags, bool *verified) {
    b200_hash_input in = {msg, msg_len, extra, extra_len};
    uint8_t h[SIG_BYTES];
    if (b200_hash_to_g1(hasher, flags, domain, 8, &in, 1, h, nullptr) != B200_OK) return engine_failed(fn);
    int ok = 0;
    if (b200_batch_verify_hashes(sig, pk, h, 1, &ok) != B200_OK) return engine_failed(fn);
    *verified = ok != 0;
    return true;
}

// ---- compress_signature / compress_pubkey: byte and integer work only (no field products), done on the host ----
// BLS12-377 Fq modulus, little-endian 32-bit words
const uint32_t FQ_MODULUS[12] = {0x00000001u, 0x8508c000u, 0x30000000u, 0x170b5d44u, 0xba094800u, 0x1ef3622fu,
                                 0x00f5138fu, 0x1a22d9f3u, 0x6ca1493bu, 0xc63b05c0u, 0x17c510eau, 0x01ae3a46u};
void load_words(const uint8_t *le48, uint32_t w[12]) {
    for (int i = 0; i < 12; i++) w[i] = (uint32_t)le48[4 * i] | ((uint32_t)le48[4 * i + 1] << 8) | ((uint32_t)le48[4 * i + 2] << 16) | ((uint32_t)le48[4 * i + 3] << 24);
}
bool words_less(const uint32_t a[12], const uint32_t b[12]) {
    for (int i = 11; i >= 0; i--)
        if (a[i] != b[i]) return a[i] < b[i];
    return false;
}
bool coord_valid(const uint8_t *le48) {                  // Fq::read: canonical integer below the modulus
    uint32_t w[12];
    load_words(le48, w);
    return words_less(w, FQ_MODULUS);
}
bool coord_is_zero(const uint8_t *le48) {
    for (int i = 0; i < 48; i++)
        if (le48[i]) return false;
    return true;
}
bool coord_over_half(const uint8_t *le48) {              // y > -y  <=>  y > (p - 1) / 2
    uint32_t w[12], half[12];
    load_words(le48, w);
    for (int i = 0; i < 12; i++) half[i] = (FQ_MODULUS[i] >> 1) | (i < 11 ? FQ_MODULUS[i + 1] << 31 : 0u);   // p odd: (p - 1) / 2 = p >> 1
    return words_less(half, w);
}
bool compress_bytes(const char *fn, const uint8_t *in, int len, size_t coords, uint8_t **out, int *out_len) {
    if (!in || !out || !out_len) return failed(fn, "null pointer");
    const size_t enc = 48 * coords;
    if (len < (int)(2 * enc)) return failed(fn, "not enough bytes");
    for (size_t k = 0; k < 2 * coords; k++)
        if (!coord_valid(in + 48 * k)) return failed(fn, "coordinate not below the modulus");
    uint8_t *buf = (uint8_t *)malloc(enc);
    if (!buf) return failed(fn, "out of memory");
    memcpy(buf, in, enc);                                 // x, little endian (Fq2: c0 | c1)
    const uint8_t *y = in + enc;
    // Fq2 ordering compares c1 first, then c0 (crates/epoch-snark/src/encoding.rs:31-33)
    const bool larger = coords == 2 && !coord_is_zero(y + 48) ? coord_over_half(y + 48) : coord_over_half(y);
    if (larger) buf[enc - 1] |= 0x80;
    *out = buf;
    *out_len = (int)enc;
    return true;
}

}  // namespace

extern "C" {

// serialization.rs:166-215: uncompressed x | y (canonical little-endian coordinates) -> the compressed encoding
bool compress_signature(const uint8_t *in_signature, int in_signature_len, uint8_t **out_signature, int *out_len) {
    return compress_bytes("compress_signature", in_signature, in_signature_len, 1, out_signature, out_len);
}
bool compress_pubkey(const uint8_t *in_pubkey, int in_pubkey_len, uint8_t **out_pubkey, int *out_len) {
    return compress_bytes("compress_pubkey", in_pubkey, in_pubkey_len, 2, out_pubkey, out_len);
}

bool deserialize_public_key(const uint8_t *in_public_key_bytes, int in_public_key_bytes_len, PublicKey **out_public_key) {
    return deserialize_point("deserialize_public_key", B200_POINTS_BLS12_377_G2, 96, in_public_key_bytes, in_public_key_bytes_len,
                             out_public_key);
}
bool deserialize_signature(const uint8_t *in_signature_bytes, int in_signature_bytes_len, Signature **out_signature) {
    return deserialize_point("deserialize_signature", B200_POINTS_BLS12_377_G1, 48, in_signature_bytes, in_signature_bytes_len,
                             out_signature);
}
bool serialize_public_key(const PublicKey *in_public_key, uint8_t **out_bytes, int *out_len) {
    return serialize_point("serialize_public_key", B200_POINTS_BLS12_377_G2, 96, in_public_key, out_bytes, out_len);
}
bool serialize_signature(const Signature *in_signature, uint8_t **out_bytes, int *out_len) {
    return serialize_point("serialize_signature", B200_POINTS_BLS12_377_G1, 48, in_signature, out_bytes, out_len);
}
bool destroy_public_key(PublicKey *public_key) {
    if (!public_key) return false;
    free(public_key);
    return true;
}
bool destroy_signature(Signature *signature) {
    if (!signature) return false;
    free(signature);
    return true;
}
bool free_vec(uint8_t *bytes, int len) {
    (void)len;
    if (!bytes) return false;
    free(bytes);
    return true;
}

bool aggregate_public_keys(const PublicKey *const *in_public_keys, int in_public_keys_len, PublicKey **out_public_key) {
    const char *fn = "aggregate_public_keys";
    if ((in_public_keys_len && !in_public_keys) || !out_public_key || in_public_keys_len < 0) return failed(fn, "null pointer");
    if (b200_ensure_init() != B200_OK) return engine_failed(fn);
    std::vector<const void *> images(in_public_keys, in_public_keys + in_public_keys_len);
    void *out = malloc(PK_BYTES);
    if (!out || !sum_images(B200_BLS12_377_G2, images, PK_BYTES, out)) {
        free(out);
        return engine_failed(fn);
    }
    *out_public_key = reinterpret_cast<PublicKey *>(out);
    return true;
}
bool aggregate_signatures(const Signature *const *in_signatures, int in_signatures_len, Signature **out_signature) {
    const char *fn = "aggregate_signatures";
    if ((in_signatures_len && !in_signatures) || !out_signature || in_signatures_len < 0) return failed(fn, "null pointer");
    if (b200_ensure_init() != B200_OK) return engine_failed(fn);
    std::vector<const void *> images(in_signatures, in_signatures + in_signatures_len);
    void *out = malloc(SIG_BYTES);
    if (!out || !sum_images(B200_BLS12_377_G1, images, SIG_BYTES, out)) {
        free(out);
        return engine_failed(fn);
    }
    *out_signature = reinterpret_cast<Signature *>(out);
    return true;
}

bool verify_signature(const PublicKey *in_public_key, const uint8_t *in_message, int in_message_len, const uint8_t *in_extra_data,
                      int in_extra_data_len, const Signature *in_signature, bool should_use_composite, bool should_use_cip22,
                      bool *out_verified) {
    const char *fn = "verify_signature";
    int hasher, flags;
    if (!in_public_key || !in_signature || !out_verified) return failed(fn, "null pointer");
    if (!hasher_of(should_use_composite, should_use_cip22, &hasher, &flags)) return failed(fn, "could not hash to curve");
    if (b200_ensure_init() != B200_OK) return engine_failed(fn);
    return verify_one(fn, in_public_key, SIG_DOMAIN, in_message, (size_t)in_message_len, in_extra_data, (size_t)in_extra_data_len,
                      in_signature, hasher, flags, out_verified);
}

bool verify_pop(const PublicKey *in_public_key, const uint8_t *in_message, int in_message_len, const Signature *in_signature,
                bool *out_verified) {
    const char *fn = "verify_pop";
    if (!in_public_key || !in_signature || !out_verified) return failed(fn, "null pointer");
    if (b200_ensure_init() != B200_OK) return engine_failed(fn);
    return verify_one(fn, in_public_key, POP_DOMAIN, in_message, (size_t)in_message_len, nullptr, 0, in_signature, B200_HASHER_DIRECT,
                      B200_HASH_COMPAT, out_verified);
}

bool batch_verify_signature(const MessageFFI *messages_ptr, size_t messages_len, bool should_use_composite, bool should_use_cip22,
                            bool *verified) {
    const char *fn = "batch_verify_signature";
    int hasher, flags;
    if ((messages_len && !messages_ptr) || !verified) return failed(fn, "null pointer");
    if (!hasher_of(should_use_composite, should_use_cip22, &hasher, &flags)) return failed(fn, "could not hash to curve");
    if (b200_ensure_init() != B200_OK) return engine_failed(fn);
    const size_t n = messages_len;
    std::vector<const void *> sigs(n);
    std::vector<b200_hash_input> inputs(n);
    std::vector<uint8_t> pks(n * PK_BYTES + 16), hashes(n * SIG_BYTES + 16);
    for (size_t i = 0; i < n; i++) {
        const MessageFFI &m = messages_ptr[i];
        if (!m.public_key || !m.sig) return failed(fn, "null handle");
        sigs[i] = m.sig;
        inputs[i] = {m.data.ptr, m.data.len, m.extra.ptr, m.extra.len};
        memcpy(&pks[i * PK_BYTES], m.public_key, PK_BYTES);
    }
    uint8_t asig[SIG_BYTES];
    if (!sum_images(B200_BLS12_377_G1, sigs, SIG_BYTES, asig)) return engine_failed(fn);            // Signature::aggregate
    // Signature::batch_verify (signature.rs:101-117): all messages hashed in one launch, then one product of n + 1 pairings
    if (b200_hash_to_g1(hasher, flags, SIG_DOMAIN, 8, inputs.data(), n, hashes.data(), nullptr) != B200_OK) return engine_failed(fn);
    int ok = 0;
    if (b200_batch_verify_hashes(asig, pks.data(), hashes.data(), n, &ok) != B200_OK) return engine_failed(fn);
    *verified = ok != 0;
    return true;
}

bool batch_verify_strict(const BatchMessageFFI *in_batches_ptr, size_t in_batches_len, bool should_use_composite, bool should_use_cip22,
                         bool *out_results) {
    const char *fn = "batch_verify_strict";
    if (in_batches_len && (!in_batches_ptr || !out_results)) return failed(fn, "null pointer");
    int hasher = 0, flags = 0;
    const bool hasher_ok = hasher_of(should_use_composite, should_use_cip22, &hasher, &flags);
    if (b200_ensure_init() != B200_OK) return engine_failed(fn);
    const size_t nb = in_batches_len;
    for (size_t b = 0; b < nb; b++) out_results[b] = false;
    if (nb == 0) return true;
    if (!hasher_ok) return failed(fn, "bad hash to curve configuration");      // every batch: Err -> false (signatures.rs:389-398)
    // one launch hashes the message of every batch (each Batch::verify hashes its own message once); should a message fail
    // to hash, the reference marks THAT batch false and goes on (signatures.rs:389-398): retry one by one to find it
    std::vector<b200_hash_input> inputs(nb);
    for (size_t b = 0; b < nb; b++) inputs[b] = {in_batches_ptr[b].data.ptr, in_batches_ptr[b].data.len, in_batches_ptr[b].extra.ptr, in_batches_ptr[b].extra.len};
    std::vector<uint8_t> hashes(nb * SIG_BYTES + 16);
    std::vector<char> hashed(nb, 1);
    if (b200_hash_to_g1(hasher, flags, SIG_DOMAIN, 8, inputs.data(), nb, hashes.data(), nullptr) != B200_OK) {
        for (size_t b = 0; b < nb; b++)
            hashed[b] = b200_hash_to_g1(hasher, flags, SIG_DOMAIN, 8, &inputs[b], 1, &hashes[b * SIG_BYTES], nullptr) == B200_OK;
    }
    // the reference draws every exponent from rand::thread_rng() (a CSPRNG seeded by the OS); here ONE getrandom() call
    // fills the exponents of all batches (std::random_device per 32-bit word was a system call each: 30 000 of them, 25 of
    // the 42 ms of a 300 x 20 call)
    size_t exp_total = 0;
    for (size_t b = 0; b < nb; b++) {
        const size_t n = in_batches_ptr[b].public_keys_len < in_batches_ptr[b].signatures_len ? in_batches_ptr[b].public_keys_len
                                                                                              : in_batches_ptr[b].signatures_len;
        exp_total += n * std::min<size_t>((128 + log2_ceil(n) + 7) / 8, 253 / 8);
    }
    std::vector<uint8_t> pool(exp_total + 1);
    if (!os_random(pool.data(), exp_total)) return failed(fn, "no entropy");
    size_t pool_at = 0;
    // every batch's keys, signatures and fresh exponents gathered, then ALL batches verified in one pass on the device
    std::vector<std::vector<uint8_t>> pks(nb), sigs(nb);
    std::vector<std::vector<uint64_t>> exps(nb);
    std::vector<b200_strict_batch> jobs;
    std::vector<size_t> job_batch;
    for (size_t b = 0; b < nb; b++) {
        if (!hashed[b]) continue;
        const BatchMessageFFI &batch = in_batches_ptr[b];
        const size_t n = batch.public_keys_len < batch.signatures_len ? batch.public_keys_len : batch.signatures_len;   // zip()
        pks[b].resize(n * PK_BYTES + 16);
        sigs[b].resize(n * SIG_BYTES + 16);
        exps[b].assign(4 * n + 4, 0);
        // byte_count_from_target_batch_size (batch.rs:23-28): min((128 + ceil(log2 n) + 7) / 8, 253 / 8) random bytes,
        // read little-endian (Fr::from_random_bytes)
        const size_t exp_bytes = std::min<size_t>((128 + log2_ceil(n) + 7) / 8, 253 / 8);
        for (size_t i = 0; i < n; i++) {
            if (!batch.public_keys[i] || !batch.signatures[i]) return failed(fn, "null handle");
            memcpy(&pks[b][i * PK_BYTES], batch.public_keys[i], PK_BYTES);
            memcpy(&sigs[b][i * SIG_BYTES], batch.signatures[i], SIG_BYTES);
            memcpy(&exps[b][4 * i], &pool[pool_at], exp_bytes);
            pool_at += exp_bytes;
        }
        jobs.push_back({pks[b].data(), sigs[b].data(), exps[b].data(), n, &hashes[b * SIG_BYTES]});
        job_batch.push_back(b);
    }
    std::vector<int> ok(jobs.size() + 1, 0);
    if (jobs.size() > 2) {
        if (b200_batch_verify_strict_many(jobs.data(), jobs.size(), ok.data()) != B200_OK) return engine_failed(fn);
    } else {
        // one or two batches: the bucket-method path has the shorter latency (the small-MSM kernel walks one
        // double-and-add chain per point, which only pays when hundreds of batches run side by side)
        for (size_t j = 0; j < jobs.size(); j++)
            if (b200_batch_verify_strict_hash(jobs[j].pubkeys, jobs[j].signatures, jobs[j].exponents, jobs[j].n, jobs[j].message_hash,
                                              &ok[j]) != B200_OK)
                return engine_failed(fn);
    }
    bool all_valid = true;
    for (size_t j = 0; j < jobs.size(); j++) out_results[job_batch[j]] = ok[j] != 0;
    for (size_t b = 0; b < nb; b++) all_valid = all_valid && out_results[b];
    if (!all_valid) return failed(fn, "signature verification failed");       // BLSError::VerificationFailed
    return true;
}

}  // extern "C"
