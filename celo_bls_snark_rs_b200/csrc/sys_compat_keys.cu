// The rest of bls-snark-sys' C-ABI, re-exported under the reference's names over the CUDA engine (SURVEY.md section 8b
// seam "B2"): private keys and signing, the hash_* byte helpers, the uncompressed encodings, key subtraction, init.
//
//   crates/bls-snark-sys/src/signatures.rs:19-91     generate_private_key, private_key_to_public_key, sign_message, sign_pop
//   crates/bls-snark-sys/src/signatures.rs:93-242    hash_direct, hash_direct_with_attempt, hash_composite, hash_crh, hash_direct_first_step, hash_composite_cip22
//   crates/bls-snark-sys/src/snark/epoch_block.rs:16-106 encode_epoch_block_to_bytes, encode_epoch_block_to_bytes_cip22
//   crates/bls-snark-sys/src/signatures.rs:454-483   aggregate_public_keys_subtract
//   crates/bls-snark-sys/src/serialization.rs:13-105 (de)serialize_private_key, deserialize_public_key_cached, serialize_*_uncompressed
//   crates/bls-snark-sys/src/serialization.rs:224-234 destroy_private_key          crates/bls-snark-sys/src/lib.rs:29-34 init
//
// Handles are the Rust types' memory images, as in sys_compat.cu: PrivateKey = Fr (32 bytes, the Montgomery residue).
// Host code here is byte and integer work only (Montgomery <-> canonical conversions of single field elements for the
// wire formats, negating a coordinate); hashing, scalar multiplications and sums run on the device.
// With these every one of the reference's 36 exports exists under its own name.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../include/b200_bls.h"
#include "../../include/bls_snark_sys_compat.h"
#include "pairing_params_gen.cuh"
#include "params_gen.cuh"

namespace {

using b200::Fq377Params;
using b200::Fr253Params;

constexpr size_t PK_BYTES = 288, SIG_BYTES = 144, SK_BYTES = 32;
const uint8_t SIG_DOMAIN[8] = {'U', 'L', 'f', 'o', 'r', 'x', 'o', 'f'};      // crates/bls-crypto/src/lib.rs:75
const uint8_t POP_DOMAIN[8] = {'U', 'L', 'f', 'o', 'r', 'p', 'o', 'p'};      // lib.rs:78

bool failed(const char *fn, const char *why) {
    fprintf(stderr, "[b200] %s -> false: %s\n", fn, why);
    return false;
}
bool engine_failed(const char *fn) { return failed(fn, b200_last_error()); }

// ---- one field element on the host: Montgomery product on 32-bit words (conversions for the wire formats only) ----
template <class P>
struct HostField {
    static constexpr int N = P::N;
    static void load(const uint8_t *le, uint32_t w[N]) { memcpy(w, le, 4 * N); }
    static void store(const uint32_t w[N], uint8_t *le) { memcpy(le, w, 4 * N); }
    static bool less_than_modulus(const uint32_t a[N]) {
        for (int i = N - 1; i >= 0; i--)
            if (a[i] != P::mod(i)) return a[i] < P::mod(i);
        return false;
    }
    // r = a * b / R mod p (CIOS)
    static void mul(const uint32_t a[N], const uint32_t b[N], uint32_t r[N]) {
        uint32_t t[N + 2] = {0};
        for (int i = 0; i < N; i++) {
            uint64_t c = 0;
            for (int j = 0; j < N; j++) {
                c += (uint64_t)a[j] * b[i] + t[j];
                t[j] = (uint32_t)c;
                c >>= 32;
            }
            c += t[N];
            t[N] = (uint32_t)c;
            t[N + 1] = (uint32_t)(c >> 32);
            const uint32_t m = t[0] * P::INV;
            c = (uint64_t)m * P::mod(0) + t[0];
            c >>= 32;
            for (int j = 1; j < N; j++) {
                c += (uint64_t)m * P::mod(j) + t[j];
                t[j - 1] = (uint32_t)c;
                c >>= 32;
            }
            c += t[N];
            t[N - 1] = (uint32_t)c;
            t[N] = t[N + 1] + (uint32_t)(c >> 32);
        }
        bool ge = t[N] != 0;
        if (!ge) {
            ge = true;
            for (int i = N - 1; i >= 0; i--)
                if (t[i] != P::mod(i)) {
                    ge = t[i] > P::mod(i);
                    break;
                }
        }
        if (ge) {
            uint64_t br = 0;
            for (int i = 0; i < N; i++) {
                uint64_t d = (uint64_t)t[i] - P::mod(i) - br;
                t[i] = (uint32_t)d;
                br = (d >> 32) & 1;
            }
        }
        memcpy(r, t, 4 * N);
    }
    static void from_mont(const uint32_t a[N], uint32_t r[N]) {
        uint32_t one[N] = {1};
        mul(a, one, r);
    }
    static void to_mont(const uint32_t a[N], uint32_t r[N]) {
        uint32_t r2[N];
        for (int i = 0; i < N; i++) r2[i] = P::r2(i);
        mul(a, r2, r);
    }
    static void neg(const uint32_t a[N], uint32_t r[N]) {                  // works on residues of either form
        bool zero = true;
        for (int i = 0; i < N; i++) zero = zero && a[i] == 0;
        if (zero) {
            memset(r, 0, 4 * N);
            return;
        }
        uint64_t br = 0;
        for (int i = 0; i < N; i++) {
            uint64_t d = (uint64_t)P::mod(i) - a[i] - br;
            r[i] = (uint32_t)d;
            br = (d >> 32) & 1;
        }
    }
};
using Fq = HostField<Fq377Params>;
using Fr = HostField<Fr253Params>;

void fq_one(uint8_t *le48) {
    uint32_t w[12];
    for (int i = 0; i < 12; i++) w[i] = Fq377Params::one(i);
    memcpy(le48, w, 48);
}
bool all_zero(const uint8_t *p, size_t n) {
    for (size_t i = 0; i < n; i++)
        if (p[i]) return false;
    return true;
}
// GroupProjective image (X | Y | Z, Montgomery) -> canonical affine coordinates.  0: finite, 1: the point at infinity,
// -1: the engine failed (b200_last_error has the text)
int image_to_affine_canonical(int curve, const void *image, size_t coords, uint8_t *xy /* 2 * coords * 48 */) {
    const size_t cb = 48 * coords;
    uint8_t aff[192], one_img[288];
    // normalise on the device: sum of one point is the point; batch-to-affine needs device buffers, so go through the
    // compressed encoder's sibling: scalar multiplication by one returns (x, y, 1)
    const uint64_t one_scalar[4] = {1, 0, 0, 0};
    if (b200_scalar_mul(curve, image, one_scalar, one_img) != B200_OK) return -1;
    if (all_zero(one_img + 2 * cb, cb)) return 1;                         // Z == 0
    memcpy(aff, one_img, 2 * cb);
    for (size_t k = 0; k < 2 * coords; k++) {
        uint32_t w[12], c[12];
        Fq::load(aff + 48 * k, w);
        Fq::from_mont(w, c);
        Fq::store(c, xy + 48 * k);
    }
    return 0;
}

uint8_t *leak(const std::vector<uint8_t> &v) {                             // what the reference hands out as a forgotten Vec
    uint8_t *p = (uint8_t *)malloc(v.size() ? v.size() : 1);
    if (p && !v.empty()) memcpy(p, v.data(), v.size());
    return p;
}
bool hand_out(const char *fn, const std::vector<uint8_t> &v, uint8_t **out, int *out_len) {
    uint8_t *p = leak(v);
    if (!p) return failed(fn, "out of memory");
    *out = p;
    *out_len = (int)v.size();
    return true;
}

bool hasher_of(bool composite, bool cip22, int *hasher, int *flags) {
    if (!composite && cip22) return false;
    *hasher = composite ? B200_HASHER_COMPOSITE : B200_HASHER_DIRECT;
    *flags = B200_HASH_COMPAT | (cip22 ? B200_HASH_CIP22 : 0);
    return true;
}

// sk (Montgomery image) -> canonical scalar limbs
void sk_scalar(const PrivateKey *sk, uint64_t out[4]) {
    uint32_t w[8], c[8];
    Fr::load(reinterpret_cast<const uint8_t *>(sk), w);
    Fr::from_mont(w, c);
    memcpy(out, c, 32);
}

bool sign_with(const char *fn, const PrivateKey *sk, const uint8_t *domain, const uint8_t *msg, int msg_len, const uint8_t *extra,
               int extra_len, int hasher, int flags, Signature **out) {
    if (!sk || !out || msg_len < 0 || extra_len < 0 || (msg_len && !msg) || (extra_len && !extra)) return failed(fn, "null pointer");
    if (b200_ensure_init() != B200_OK) return engine_failed(fn);
    b200_hash_input in = {msg, (size_t)msg_len, extra, (size_t)extra_len};
    uint8_t h[SIG_BYTES];
    if (b200_hash_to_g1(hasher, flags, domain, 8, &in, 1, h, nullptr) != B200_OK) return engine_failed(fn);
    uint64_t s[4];
    sk_scalar(sk, s);
    uint8_t *img = (uint8_t *)malloc(SIG_BYTES);
    if (!img) return failed(fn, "out of memory");
    if (b200_scalar_mul(B200_BLS12_377_G1, h, s, img) != B200_OK) {        // sign_raw: hash * sk (secret.rs:65-67)
        free(img);
        return engine_failed(fn);
    }
    *out = reinterpret_cast<Signature *>(img);
    return true;
}

bool hash_point(const char *fn, int hasher, int flags, const uint8_t *domain, const uint8_t *msg, int msg_len, const uint8_t *extra,
                int extra_len, uint8_t h[SIG_BYTES], uint32_t *attempt) {
    if (msg_len < 0 || extra_len < 0 || (msg_len && !msg) || (extra_len && !extra)) return failed(fn, "null pointer");
    if (b200_ensure_init() != B200_OK) return engine_failed(fn);
    b200_hash_input in = {msg, (size_t)msg_len, extra, (size_t)extra_len};
    if (b200_hash_to_g1(hasher, flags, domain, 8, &in, 1, h, attempt) != B200_OK) return engine_failed(fn);
    return true;
}

// G1Projective::write of the normalised representative (x, y, 1): canonical x | y | z, 144 bytes.  The reference writes
// whatever representative its cofactor multiplication ended on; the POINT is the same (INTEGRATION.md lists the deviation).
bool projective_bytes(const char *fn, const uint8_t h[SIG_BYTES], std::vector<uint8_t> *out) {
    out->assign(144, 0);
    uint8_t xy[96];
    const int st = image_to_affine_canonical(B200_BLS12_377_G1, h, 1, xy);
    if (st < 0) return engine_failed(fn);
    if (st == 1) {
        (*out)[0] = 1;                                                      // zero(): (1, 1, 0)
        (*out)[48] = 1;
        return true;
    }
    memcpy(out->data(), xy, 96);
    (*out)[96] = 1;
    return true;
}

}  // namespace

extern "C" {

bool init(void) { return b200_ensure_init() == B200_OK; }

// ---- private keys ----
bool generate_private_key(PrivateKey **out_private_key) {
    if (!out_private_key) return failed("generate_private_key", "null pointer");
    std::random_device rd;                                                  // the reference draws from rand::thread_rng()
    uint32_t w[8];
    do {                                                                    // Fr::rand: 253 random bits, rejected unless below r;
        for (uint32_t &x : w) x = rd();                                     // the accepted limbs ARE the memory image
        w[7] &= 0xffffffffu >> 3;
    } while (!Fr::less_than_modulus(w));
    uint8_t *img = (uint8_t *)malloc(SK_BYTES);
    if (!img) return failed("generate_private_key", "out of memory");
    memcpy(img, w, SK_BYTES);
    *out_private_key = reinterpret_cast<PrivateKey *>(img);
    return true;
}
bool deserialize_private_key(const uint8_t *in_bytes, int in_len, PrivateKey **out_private_key) {
    const char *fn = "deserialize_private_key";
    if (!in_bytes || !out_private_key) return failed(fn, "null pointer");
    if (in_len < (int)SK_BYTES) return failed(fn, "not enough bytes");
    uint32_t w[8], m[8];
    Fr::load(in_bytes, w);
    if (!Fr::less_than_modulus(w)) return failed(fn, "scalar not below the group order");
    Fr::to_mont(w, m);
    uint8_t *img = (uint8_t *)malloc(SK_BYTES);
    if (!img) return failed(fn, "out of memory");
    memcpy(img, m, SK_BYTES);
    *out_private_key = reinterpret_cast<PrivateKey *>(img);
    return true;
}
bool serialize_private_key(const PrivateKey *in_private_key, uint8_t **out_bytes, int *out_len) {
    const char *fn = "serialize_private_key";
    if (!in_private_key || !out_bytes || !out_len) return failed(fn, "null pointer");
    uint64_t s[4];
    sk_scalar(in_private_key, s);
    std::vector<uint8_t> v(SK_BYTES);
    memcpy(v.data(), s, SK_BYTES);
    return hand_out(fn, v, out_bytes, out_len);
}
bool destroy_private_key(PrivateKey *private_key) {
    if (!private_key) return false;
    memset(private_key, 0, SK_BYTES);
    free(private_key);
    return true;
}
bool private_key_to_public_key(const PrivateKey *in_private_key, PublicKey **out_public_key) {
    const char *fn = "private_key_to_public_key";
    if (!in_private_key || !out_public_key) return failed(fn, "null pointer");
    if (b200_ensure_init() != B200_OK) return engine_failed(fn);
    // g2 = -(-g2): the packed generator with y negated back, as a GroupProjective image (x, y, 1)
    uint8_t g2[PK_BYTES];
    memset(g2, 0, sizeof g2);
    memcpy(g2, b200::NEG_G2_GENERATOR_PACKED, 96);
    for (int k = 0; k < 2; k++) {
        uint32_t w[12], n[12];
        memcpy(w, reinterpret_cast<const uint8_t *>(b200::NEG_G2_GENERATOR_PACKED) + 96 + 48 * k, 48);
        Fq::neg(w, n);
        memcpy(g2 + 96 + 48 * k, n, 48);
    }
    fq_one(g2 + 192);
    uint64_t s[4];
    sk_scalar(in_private_key, s);
    uint8_t *img = (uint8_t *)malloc(PK_BYTES);
    if (!img) return failed(fn, "out of memory");
    if (b200_scalar_mul(B200_BLS12_377_G2, g2, s, img) != B200_OK) {
        free(img);
        return engine_failed(fn);
    }
    *out_public_key = reinterpret_cast<PublicKey *>(img);
    return true;
}

// ---- signing ----
bool sign_message(const PrivateKey *in_private_key, const uint8_t *in_message, int in_message_len, const uint8_t *in_extra_data,
                  int in_extra_data_len, bool should_use_composite, bool should_use_cip22, Signature **out_signature) {
    int hasher, flags;
    if (!hasher_of(should_use_composite, should_use_cip22, &hasher, &flags)) return failed("sign_message", "could not hash to curve");
    return sign_with("sign_message", in_private_key, SIG_DOMAIN, in_message, in_message_len, in_extra_data, in_extra_data_len, hasher,
                     flags, out_signature);
}
bool sign_pop(const PrivateKey *in_private_key, const uint8_t *in_message, int in_message_len, Signature **out_signature) {
    return sign_with("sign_pop", in_private_key, POP_DOMAIN, in_message, in_message_len, nullptr, 0, B200_HASHER_DIRECT, B200_HASH_COMPAT,
                     out_signature);
}

// ---- hash helpers ----
// G1Affine::write: canonical x | y | infinity byte (97 bytes)
static bool hash_direct_impl(const char *fn, const uint8_t *msg, int len, uint8_t **out_hash, int *out_len, int *out_attempt, bool use_pop) {
    if (!out_hash || !out_len) return failed(fn, "null pointer");
    uint8_t h[SIG_BYTES];
    uint32_t attempt = 0;
    if (!hash_point(fn, B200_HASHER_DIRECT, B200_HASH_COMPAT, use_pop ? POP_DOMAIN : SIG_DOMAIN, msg, len, nullptr, 0, h, &attempt)) return false;
    std::vector<uint8_t> v(97, 0);
    const int st = image_to_affine_canonical(B200_BLS12_377_G1, h, 1, v.data());
    if (st < 0) return engine_failed(fn);
    if (st == 1) {
        memset(v.data(), 0, 96);                                            // G1Affine::zero() = (0, 1, true)
        v[48] = 1;
        v[96] = 1;
    }
    if (out_attempt) *out_attempt = (int)attempt;
    return hand_out(fn, v, out_hash, out_len);
}
bool hash_direct(const uint8_t *in_message, int in_message_len, uint8_t **out_hash, int *out_len, bool use_pop) {
    return hash_direct_impl("hash_direct", in_message, in_message_len, out_hash, out_len, nullptr, use_pop);
}
bool hash_direct_with_attempt(const uint8_t *in_message, int in_message_len, uint8_t **out_hash, int *out_len, int *out_attempt,
                              bool use_pop) {
    if (!out_attempt) return failed("hash_direct_with_attempt", "null pointer");
    return hash_direct_impl("hash_direct_with_attempt", in_message, in_message_len, out_hash, out_len, out_attempt, use_pop);
}
bool hash_composite(const uint8_t *in_message, int in_message_len, const uint8_t *in_extra_data, int in_extra_data_len, uint8_t **out_hash,
                    int *out_len) {
    const char *fn = "hash_composite";
    if (!out_hash || !out_len) return failed(fn, "null pointer");
    uint8_t h[SIG_BYTES];
    if (!hash_point(fn, B200_HASHER_COMPOSITE, B200_HASH_COMPAT, SIG_DOMAIN, in_message, in_message_len, in_extra_data, in_extra_data_len, h, nullptr))
        return false;
    std::vector<uint8_t> v;
    return projective_bytes(fn, h, &v) && hand_out(fn, v, out_hash, out_len);
}
bool hash_composite_cip22(const uint8_t *in_message, int in_message_len, const uint8_t *in_extra_data, int in_extra_data_len,
                          uint8_t **out_hash, int *out_len, uint8_t *attempt_counter) {
    const char *fn = "hash_composite_cip22";
    if (!out_hash || !out_len || !attempt_counter) return failed(fn, "null pointer");
    uint8_t h[SIG_BYTES];
    uint32_t attempt = 0;
    if (!hash_point(fn, B200_HASHER_COMPOSITE, B200_HASH_COMPAT | B200_HASH_CIP22, SIG_DOMAIN, in_message, in_message_len, in_extra_data,
                    in_extra_data_len, h, &attempt))
        return false;
    *attempt_counter = (uint8_t)attempt;
    std::vector<uint8_t> v;
    return projective_bytes(fn, h, &v) && hand_out(fn, v, out_hash, out_len);
}
// COMPOSITE_HASHER.crh(SIG_DOMAIN, message, _): the 48-byte x coordinate of the Bowe-Hopwood point
bool hash_crh(const uint8_t *in_message, int in_message_len, int hash_bytes, uint8_t **out_hash, int *out_len) {
    const char *fn = "hash_crh";
    (void)hash_bytes;                                                       // the composite CRH ignores the requested size (composite.rs:60-75)
    if (!out_hash || !out_len || in_message_len < 0 || (in_message_len && !in_message)) return failed(fn, "null pointer");
    if (b200_ensure_init() != B200_OK) return engine_failed(fn);
    b200_hash_input in = {in_message, (size_t)in_message_len, nullptr, 0};
    std::vector<uint8_t> v(48, 0);
    if (b200_hash_to_g1(B200_HASHER_COMPOSITE, B200_HASH_CRH_ONLY, SIG_DOMAIN, 8, &in, 1, v.data(), nullptr) != B200_OK) return engine_failed(fn);
    return hand_out(fn, v, out_hash, out_len);
}

// DirectHasher.hash(SIG_DOMAIN, message, hash_bytes) = xof(crh(message)) (crates/bls-crypto/src/hashers/direct.rs:23-79,
// signatures.rs:191-212): pure byte work -- two or more Blake2s calls on the host, no field element involved
bool hash_direct_first_step(const uint8_t *in_message, int in_message_len, int hash_bytes, uint8_t **out_hash, int *out_len) {
    const char *fn = "hash_direct_first_step";
    if (!out_hash || !out_len || in_message_len < 0 || (in_message_len && !in_message) || hash_bytes < 0 || hash_bytes > 0xffff)
        return failed(fn, "bad argument");
    const uint64_t len_tag = (uint64_t)(uint16_t)hash_bytes << 32;             // xof_digest_length_to_node_offset
    uint8_t crh[32];
    b200_blake2s_param(in_message, (size_t)in_message_len, 32, 1, 1, 0, len_tag, 0, SIG_DOMAIN, crh);
    std::vector<uint8_t> v((size_t)hash_bytes, 0);
    const int blocks = (hash_bytes + 31) / 32;
    for (int i = 0; i < blocks; i++) {
        const int part = (i == blocks - 1 && hash_bytes % 32) ? hash_bytes % 32 : 32;
        b200_blake2s_param(crh, 32, part, 0, 0, 32, (uint64_t)i | len_tag, 32, SIG_DOMAIN, v.data() + 32 * i);
    }
    return hand_out(fn, v, out_hash, out_len);
}

// ---- epoch-block byte encodings (snark/epoch_block.rs:16-106): keys compressed on the device in one pass, bit shuffling on the host ----
static bool encode_block(const char *fn, int cip22, uint16_t index, uint8_t round, const uint8_t *epoch_entropy, const uint8_t *parent_entropy,
                         uint32_t maximum_non_signers, size_t maximum_validators, const PublicKey *const *keys, int nkeys, uint8_t **out_bytes,
                         int *out_len, uint8_t **out_extra, int *out_extra_len) {
    if (!out_bytes || !out_len || nkeys < 0 || (nkeys && !keys) || (cip22 && (!out_extra || !out_extra_len))) return failed(fn, "null pointer");
    std::vector<uint8_t> images((size_t)nkeys * PK_BYTES), keys96((size_t)nkeys * 96);
    for (int i = 0; i < nkeys; i++) {
        if (!keys[i]) return failed(fn, "null handle");
        memcpy(&images[(size_t)i * PK_BYTES], keys[i], PK_BYTES);
    }
    if (nkeys) {
        if (b200_ensure_init() != B200_OK) return engine_failed(fn);
        if (b200_serialize_points(B200_POINTS_BLS12_377_G2, images.data(), (size_t)nkeys, keys96.data()) != B200_OK) return engine_failed(fn);
    }
    uint8_t *inner = nullptr, *extra = nullptr;
    size_t inner_len = 0, extra_len = 0;
    if (b200_encode_epoch_block(cip22, index, round, epoch_entropy, parent_entropy, maximum_non_signers, maximum_validators, keys96.data(),
                                (size_t)nkeys, &inner, &inner_len, cip22 ? &extra : nullptr, cip22 ? &extra_len : nullptr) != B200_OK)
        return engine_failed(fn);
    *out_bytes = inner;
    *out_len = (int)inner_len;
    if (cip22) {
        *out_extra = extra;
        *out_extra_len = (int)extra_len;
    }
    return true;
}
bool encode_epoch_block_to_bytes_cip22(unsigned short in_epoch_index, unsigned char in_round_number, const uint8_t *in_epoch_entropy,
                                       const uint8_t *in_parent_entropy, unsigned int in_maximum_non_signers, unsigned int in_maximum_validators,
                                       const PublicKey *const *in_added_public_keys, int in_added_public_keys_len, uint8_t **out_bytes,
                                       int *out_len, uint8_t **out_extra_data_bytes, int *out_extra_data_len) {
    return encode_block("encode_epoch_block_to_bytes_cip22", 1, in_epoch_index, in_round_number, in_epoch_entropy, in_parent_entropy,
                        in_maximum_non_signers, in_maximum_validators, in_added_public_keys, in_added_public_keys_len, out_bytes, out_len,
                        out_extra_data_bytes, out_extra_data_len);
}
bool encode_epoch_block_to_bytes(unsigned short in_epoch_index, unsigned int in_maximum_non_signers, const PublicKey *const *in_added_public_keys,
                                 int in_added_public_keys_len, uint8_t **out_bytes, int *out_len) {
    return encode_block("encode_epoch_block_to_bytes", 0, in_epoch_index, 0, nullptr, nullptr, in_maximum_non_signers,
                        (size_t)(in_added_public_keys_len > 0 ? in_added_public_keys_len : 0), in_added_public_keys, in_added_public_keys_len,
                        out_bytes, out_len, nullptr, nullptr);
}

// ---- encodings ----
// serialize_uncompressed of the affine point: canonical x | y, the infinity flag in bit 6 of the last byte
static bool uncompressed(const char *fn, int curve, const void *image, size_t coords, uint8_t **out_bytes, int *out_len) {
    if (!image || !out_bytes || !out_len) return failed(fn, "null pointer");
    if (b200_ensure_init() != B200_OK) return engine_failed(fn);
    std::vector<uint8_t> v(96 * coords, 0);
    const int st = image_to_affine_canonical(curve, image, coords, v.data());
    if (st < 0) return engine_failed(fn);
    if (st == 1) {
        memset(v.data(), 0, v.size());
        v.back() |= 0x40;
    }
    return hand_out(fn, v, out_bytes, out_len);
}
bool serialize_public_key_uncompressed(const PublicKey *in_public_key, uint8_t **out_bytes, int *out_len) {
    return uncompressed("serialize_public_key_uncompressed", B200_BLS12_377_G2, in_public_key, 2, out_bytes, out_len);
}
bool serialize_signature_uncompressed(const Signature *in_signature, uint8_t **out_bytes, int *out_len) {
    return uncompressed("serialize_signature_uncompressed", B200_BLS12_377_G1, in_signature, 1, out_bytes, out_len);
}
// the reference's cache only memoises the decoding (cache.rs); the key handed out is the same
bool deserialize_public_key_cached(const uint8_t *in_public_key_bytes, int in_public_key_bytes_len, PublicKey **out_public_key) {
    return deserialize_public_key(in_public_key_bytes, in_public_key_bytes_len, out_public_key);
}

// aggregated - sum(keys): one device sum over [aggregated, -key_0, -key_1, ...]
bool aggregate_public_keys_subtract(const PublicKey *in_aggregated_public_key, const PublicKey *const *in_public_keys, int in_public_keys_len,
                                    PublicKey **out_public_key) {
    const char *fn = "aggregate_public_keys_subtract";
    if (!in_aggregated_public_key || (in_public_keys_len && !in_public_keys) || !out_public_key || in_public_keys_len < 0)
        return failed(fn, "null pointer");
    if (b200_ensure_init() != B200_OK) return engine_failed(fn);
    const size_t n = (size_t)in_public_keys_len;
    std::vector<uint8_t> host((n + 1) * PK_BYTES + 16);
    memcpy(host.data(), in_aggregated_public_key, PK_BYTES);
    for (size_t i = 0; i < n; i++) {
        if (!in_public_keys[i]) return failed(fn, "null handle");
        uint8_t *dst = &host[(i + 1) * PK_BYTES];
        memcpy(dst, in_public_keys[i], PK_BYTES);
        for (int k = 0; k < 2; k++) {                                       // -(X, Y, Z) = (X, -Y, Z)
            uint32_t w[12], neg[12];
            memcpy(w, dst + 96 + 48 * k, 48);
            Fq::neg(w, neg);
            memcpy(dst + 96 + 48 * k, neg, 48);
        }
    }
    void *out = malloc(PK_BYTES);
    if (!out || b200_sum_jacobian(B200_BLS12_377_G2, host.data(), n + 1, out) != B200_OK) {
        free(out);
        return engine_failed(fn);
    }
    *out_public_key = reinterpret_cast<PublicKey *>(out);
    return true;
}

}  // extern "C"
