// The SNARK verifier entry point of bls-snark-sys, rebuilt over the CUDA engine (SURVEY.md section 8b "B2").
//
//   bool verify(vk, vk_len, proof, proof_len, EpochBlockFFI first_epoch, EpochBlockFFI last_epoch)
//       crates/bls-snark-sys/src/snark/mod.rs:23-45        (the exported function this file replaces)
//       crates/bls-snark-sys/src/snark/epoch_block.rs:109-211 (EpochBlockFFI, read_pubkeys, read_epoch_entropy)
//       crates/epoch-snark/src/api/verifier.rs:23-40       (hash -> pack -> verify_proof)
//       crates/epoch-snark/src/epoch_block.rs:106-236      (CIP22 encodings, Blake2s "ULforout" edge hashes)
//       crates/epoch-snark/src/encoding.rs:23-80           (encode_public_key, encode_u16 / encode_u32)
//       crates/epoch-snark/src/gadgets/mod.rs:75-83        (pack: 376-bit big-endian chunks)
//
// Split of work: the host does what is byte shuffling (bit encodings, one Blake2s per epoch block, packing);
// every field / curve operation runs on the device: point decoding with the deserialisation checks (codec.cuh),
// aggregation of the last block's keys, g_ic, the four Miller loops and the final exponentiation
// (pairing_bw6.cuh).  There is no host big-integer code in this path.
#include "codec.cuh"
#include "curve_impl.cuh"

#include "../../include/bls_snark_sys_compat.h"

namespace b200 {

int bw6_groth16_verify_core(Engine &E, char *d_packed, size_t nabc, const void *d_scalars, int *out_verified);

// ---- Blake2s (RFC 7693), unkeyed, 32-byte digest, 8-byte personalisation -----------------------------
namespace {
const uint32_t BLAKE2S_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
const uint8_t BLAKE2S_SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};

inline uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

void blake2s_compress(uint32_t h[8], const uint8_t block[64], uint64_t t, bool last) {
    uint32_t m[16], v[16];
    for (int i = 0; i < 16; i++) m[i] = (uint32_t)block[4 * i] | ((uint32_t)block[4 * i + 1] << 8) | ((uint32_t)block[4 * i + 2] << 16) | ((uint32_t)block[4 * i + 3] << 24);
    for (int i = 0; i < 8; i++) {
        v[i] = h[i];
        v[8 + i] = BLAKE2S_IV[i];
    }
    v[12] ^= (uint32_t)t;
    v[13] ^= (uint32_t)(t >> 32);
    if (last) v[14] = ~v[14];
    auto G = [&](int a, int b, int c, int d, uint32_t x, uint32_t y) {
        v[a] = v[a] + v[b] + x;
        v[d] = rotr32(v[d] ^ v[a], 16);
        v[c] = v[c] + v[d];
        v[b] = rotr32(v[b] ^ v[c], 12);
        v[a] = v[a] + v[b] + y;
        v[d] = rotr32(v[d] ^ v[a], 8);
        v[c] = v[c] + v[d];
        v[b] = rotr32(v[b] ^ v[c], 7);
    };
    for (int r = 0; r < 10; r++) {
        const uint8_t *s = BLAKE2S_SIGMA[r];
        G(0, 4, 8, 12, m[s[0]], m[s[1]]);
        G(1, 5, 9, 13, m[s[2]], m[s[3]]);
        G(2, 6, 10, 14, m[s[4]], m[s[5]]);
        G(3, 7, 11, 15, m[s[6]], m[s[7]]);
        G(0, 5, 10, 15, m[s[8]], m[s[9]]);
        G(1, 6, 11, 12, m[s[10]], m[s[11]]);
        G(2, 7, 8, 13, m[s[12]], m[s[13]]);
        G(3, 4, 9, 14, m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[8 + i];
}
}  // namespace

// full parameter block (RFC 7693 section 2.5, sequential or tree parameters): what DirectHasher's XOF needs
// (crates/bls-crypto/src/hashers/direct.rs:41-79: fanout 0, depth 0, leaf 32, inner 32, node offset i | digest length << 32)
void blake2s_param(const uint8_t *data, size_t len, int digest_len, int fanout, int depth, uint32_t leaf_len, uint64_t node_offset,
                   int inner_len, const uint8_t personal[8], uint8_t *out) {
    uint32_t h[8];
    for (int i = 0; i < 8; i++) h[i] = BLAKE2S_IV[i];
    h[0] ^= (uint32_t)digest_len | ((uint32_t)fanout << 16) | ((uint32_t)depth << 24);
    h[1] ^= leaf_len;
    h[2] ^= (uint32_t)node_offset;
    h[3] ^= (uint32_t)((node_offset >> 32) & 0xffffu) | ((uint32_t)inner_len << 24);       // node depth 0
    h[6] ^= (uint32_t)personal[0] | ((uint32_t)personal[1] << 8) | ((uint32_t)personal[2] << 16) | ((uint32_t)personal[3] << 24);
    h[7] ^= (uint32_t)personal[4] | ((uint32_t)personal[5] << 8) | ((uint32_t)personal[6] << 16) | ((uint32_t)personal[7] << 24);
    uint8_t block[64];
    size_t off = 0;
    while (len - off > 64) {
        blake2s_compress(h, data + off, off + 64, false);
        off += 64;
    }
    memset(block, 0, 64);
    if (len - off) memcpy(block, data + off, len - off);
    blake2s_compress(h, block, len, true);
    for (int i = 0; i < digest_len; i++) out[i] = (uint8_t)(h[i >> 2] >> (8 * (i & 3)));
}
void blake2s_personal(const uint8_t *data, size_t len, const uint8_t personal[8], uint8_t out[32]) {
    blake2s_param(data, len, 32, 1, 1, 0, 0, 0, personal, out);                            // digest 32, no key, fanout 1, depth 1
}

// ---- CIP22 bit encodings (host) ------------------------------------------------------------------------
namespace {
using Bits = std::vector<uint8_t>;

void push_le_bytes(Bits &b, const uint8_t *bytes, size_t n) {      // every byte least-significant bit first
    for (size_t i = 0; i < n; i++)
        for (int k = 0; k < 8; k++) b.push_back((bytes[i] >> k) & 1u);
}
void push_le_int(Bits &b, uint64_t v, int nbytes) {
    uint8_t raw[8];
    for (int i = 0; i < nbytes; i++) raw[i] = (uint8_t)(v >> (8 * i));
    push_le_bytes(b, raw, nbytes);
}
// 377 bits of a little-endian 48-byte coordinate, most significant first
void push_coord_be(Bits &b, const uint8_t *le48) {
    for (int i = 376; i >= 0; i--) b.push_back((le48[i >> 3] >> (i & 7)) & 1u);
}
// encode_public_key from the canonical x (c0 | c1, flag bits of the last byte ignored) and the y bit
void push_public_key(Bits &b, const uint8_t *x96, bool y_over_half) {
    push_coord_be(b, x96);
    push_coord_be(b, x96 + 48);
    b.push_back(y_over_half ? 1 : 0);
}
void push_block_cip22(Bits &b, const EpochBlockFFI &blk, bool first) {
    push_le_int(b, blk.index, 2);
    const uint8_t *entropy = first ? blk.parent_entropy : blk.epoch_entropy;
    if (entropy) push_le_bytes(b, entropy, 16);
    else b.insert(b.end(), 128, 0);
    push_le_int(b, blk.maximum_non_signers, 4);
    for (size_t i = 0; i < blk.pubkeys_num; i++) {
        const uint8_t *key = blk.pubkeys + 96 * i;
        push_public_key(b, key, (key[95] >> 7) & 1u);               // the wire flag IS the y-over-half bit (same order on Fq2)
    }
    for (size_t i = blk.pubkeys_num; i < blk.maximum_validators; i++)
        push_public_key(b, reinterpret_cast<const uint8_t *>(G2_GENERATOR_X_CANONICAL), G2_GENERATOR_Y_OVER_HALF != 0);
}
// bits (big-endian string) -> little-endian bytes -> Blake2s("ULforout") -> 256 bits, bytes least-significant bit first
void hash_bits(const Bits &bits, Bits &out) {
    std::vector<uint8_t> bytes((bits.size() + 7) / 8, 0);
    const size_t n = bits.size();
    for (size_t i = 0; i < n; i++)
        if (bits[n - 1 - i]) bytes[i >> 3] |= (uint8_t)(1u << (i & 7));
    uint8_t digest[32];
    blake2s_personal(bytes.data(), bytes.size(), reinterpret_cast<const uint8_t *>("ULforout"), digest);
    push_le_bytes(out, digest, 32);
}
// pack: chunks of 376 bits, first bit most significant -> canonical 6 x u64 scalars
std::vector<uint64_t> pack_376(const Bits &bits) {
    const size_t chunks = (bits.size() + 375) / 376;
    std::vector<uint64_t> out(chunks * 6, 0);
    for (size_t c = 0; c < chunks; c++) {
        const size_t lo = c * 376, hi = std::min(bits.size(), lo + 376), width = hi - lo;
        for (size_t i = 0; i < width; i++)
            if (bits[lo + i]) {
                const size_t pos = width - 1 - i;
                out[c * 6 + (pos >> 6)] |= (uint64_t)1 << (pos & 63);
            }
    }
    return out;
}
}  // namespace

// bits (big-endian string) -> little-endian bytes (bls-gadgets utils.rs:2-21 bits_be_to_bytes_le)
static std::vector<uint8_t> bits_be_to_bytes_le(const Bits &bits) {
    std::vector<uint8_t> bytes((bits.size() + 7) / 8, 0);
    const size_t n = bits.size();
    for (size_t i = 0; i < n; i++)
        if (bits[n - 1 - i]) bytes[i >> 3] |= (uint8_t)(1u << (i & 7));
    return bytes;
}
// EpochBlock::encode_to_bytes (pre-Donut) and ::encode_inner_to_bytes_cip22 (crates/epoch-snark/src/epoch_block.rs:106-114,
// 152-171, 191-211) from the keys' compressed encodings (96 bytes each: the wire's sign flag is encode_public_key's y bit).
void epoch_block_encode(int cip22, uint16_t index, uint8_t round, const uint8_t *epoch_entropy, const uint8_t *parent_entropy,
                        uint32_t maximum_non_signers, size_t maximum_validators, const uint8_t *keys96, size_t nkeys,
                        std::vector<uint8_t> *inner, std::vector<uint8_t> *extra) {
    Bits bits, xbits;
    auto push_keys = [&]() {
        for (size_t i = 0; i < nkeys; i++) push_public_key(bits, keys96 + 96 * i, (keys96[96 * i + 95] >> 7) & 1u);
    };
    if (!cip22) {
        push_le_int(bits, index, 2);
        push_le_int(bits, maximum_non_signers, 4);
        push_keys();
        *inner = bits_be_to_bytes_le(bits);
        return;
    }
    push_le_int(xbits, index, 2);
    push_le_int(xbits, round, 1);
    push_le_int(xbits, maximum_non_signers, 4);
    for (const uint8_t *entropy : {epoch_entropy, parent_entropy}) {
        if (entropy) push_le_bytes(bits, entropy, 16);
        else bits.insert(bits.end(), 128, 0);
    }
    push_keys();
    for (size_t i = nkeys; i < maximum_validators; i++)
        push_public_key(bits, reinterpret_cast<const uint8_t *>(G2_GENERATOR_X_CANONICAL), G2_GENERATOR_Y_OVER_HALF != 0);
    *inner = bits_be_to_bytes_le(bits);
    *extra = bits_be_to_bytes_le(xbits);
}

// ---- subgroup membership, latency form ----------------------------------------------------------------------
// r * P == O on ONE BLOCK PER POINT with the warp-cooperative point operations of coop.cuh (one limb per lane, the
// independent products of a round on four warps): `verify` decodes ten BW6-761 points and eight BLS12-377 G2 keys, and
// with one thread per point (codec.cuh k_subgroup_check) every call paid one thread's chain of 376 doublings of
// 761-bit points -- 16 ms of the 37 ms entry point.  Same double-and-add, same exceptional cases (CoopPoint::add is
// exact), so the last addition lands on O exactly when the per-thread kernel's does.
template <class F, class RP>
__global__ void __launch_bounds__(COOP_THREADS) k_subgroup_check_coop(const AffineMem<F> *__restrict__ pts, uint32_t n,
                                                                      int *__restrict__ status) {
    using CP = CoopPoint<F>;
    using C = Coop<F>;
    constexpr int W = C::WORDS;
    __shared__ CoopSm<F> sm;
    const uint32_t i = blockIdx.x;
    if (i >= n || status[i] != DECODE_OK) return;                  // uniform over the block
    const typename C::Ctx c = C::Ctx::make();
    const uint32_t *src = reinterpret_cast<const uint32_t *>(pts + i);
    for (int k = threadIdx.x; k < 2 * W; k += blockDim.x) sm.in[k / W][k % W] = __ldg(src + k);      // x | y
    for (int k = threadIdx.x; k < 4 * W; k += blockDim.x) sm.pt[k / W][k % W] = 0u;                  // infinity
    if ((threadIdx.x >> 5) == 0) {
        C::one(c).store(c, sm.in[2]);
        C::one(c).store(c, sm.in[3]);
    }
    __syncthreads();
    CP::add(c, sm);
#pragma unroll 1
    for (int b = RP::BITS - 2; b >= 0; b--) {
        CP::dbl(c, sm);
        uint32_t word = 0;
#pragma unroll
        for (int k = 0; k < RP::N; k++) word = (k == (b >> 5)) ? RP::mod(k) : word;
        if ((word >> (b & 31)) & 1u) CP::add(c, sm);
    }
    __syncthreads();
    const C zz = C::load(c, sm.pt[2]);
    if (!C::is_zero(c, zz) && threadIdx.x == 0) status[i] = DECODE_NOT_IN_SUBGROUP;
}
// one block per point while all blocks are resident at once (latency: one chain); beyond that the per-thread kernel's
// throughput wins
constexpr size_t SUBGROUP_COOP_MAX = 1184;
template <class F, class RP>
static int subgroup_check(const void *d_pts, size_t n, int *d_status, cudaStream_t st) {
    static const bool force_thread = getenv("B200_SUBGROUP_THREAD") != nullptr;
    if (n <= SUBGROUP_COOP_MAX && !force_thread)
        k_subgroup_check_coop<F, RP><<<(unsigned)n, COOP_THREADS, 0, st>>>(reinterpret_cast<const AffineMem<F> *>(d_pts), (uint32_t)n, d_status);
    else
        k_subgroup_check<F, RP><<<(unsigned)ceil_div(n, 64), 64, 0, st>>>(reinterpret_cast<const AffineMem<F> *>(d_pts), (uint32_t)n, d_status);
    LAUNCH_CHECK();
    return B200_OK;
}

// ---- device-side decoding ------------------------------------------------------------------------------
// kind: 0 = BLS12-377 G2, 1 = BW6-761 G1, 2 = BW6-761 G2.  d_src: n x 96 bytes; d_out: n packed affine records;
// d_status: n ints (DECODE_*).  Subgroup membership is checked when `subgroup` is set.
int decode_points(int kind, const void *d_src, size_t n, int subgroup, void *d_out, int *d_status, cudaStream_t st) {
    if (n == 0) return B200_OK;
    const unsigned blocks = (unsigned)ceil_div(n, 64);
    if (kind == 0) {
        k_g2_377_decompress<<<blocks, 64, 0, st>>>(reinterpret_cast<const uint32_t *>(d_src), (uint32_t)n,
                                                   reinterpret_cast<AffineMem<CFq2> *>(d_out), d_status);
        LAUNCH_CHECK();
        int rc;
        if (subgroup && (rc = subgroup_check<CFq2, Fr253Params>(d_out, n, d_status, st))) return rc;
    } else if (kind == 1 || kind == 2) {
        k_bw6_decompress<<<blocks, 64, 0, st>>>(reinterpret_cast<const uint32_t *>(d_src), (uint32_t)n, 0u, kind == 2 ? (uint32_t)n : 0u,
                                                reinterpret_cast<AffineMem<Fq761> *>(d_out), d_status);
        LAUNCH_CHECK();
        int rc;
        if (subgroup && (rc = subgroup_check<Fq761, Fq377Params>(d_out, n, d_status, st))) return rc;
    } else if (kind == 3) {
        k_g1_377_decompress<<<blocks, 64, 0, st>>>(reinterpret_cast<const uint32_t *>(d_src), (uint32_t)n,
                                                   reinterpret_cast<AffineMem<CFq> *>(d_out), d_status);
        LAUNCH_CHECK();
        int rc;
        if (subgroup && (rc = subgroup_check<CFq, Fr253Params>(d_out, n, d_status, st))) return rc;
    } else {
        return fail(B200_ERR_ARG, "unknown point kind %d", kind);
    }
    return B200_OK;
}

int decode_points_host(Engine &E, int kind, const void *bytes, size_t n, int subgroup, void *out_packed, int *out_status) {
    cudaStream_t st = E.stream;
    int rc;
    if (n == 0) return B200_OK;
    const size_t enc = kind == 3 ? 48 : 96;                // bytes of one encoding; the packed record is twice that
    if ((rc = E.h2d_bases.reserve(n * enc)) || (rc = E.native_bases.reserve(n * 2 * enc)) || (rc = E.scalars.reserve(n * sizeof(int)))) return rc;
    CUDA_TRY(cudaMemcpyAsync(E.h2d_bases.p, bytes, n * enc, cudaMemcpyHostToDevice, st));
    if ((rc = decode_points(kind, E.h2d_bases.p, n, subgroup, E.native_bases.p, E.scalars.as<int>(), st))) return rc;
    if (out_packed) CUDA_TRY(cudaMemcpyAsync(out_packed, E.native_bases.p, n * 2 * enc, cudaMemcpyDeviceToHost, st));
    if (out_status) CUDA_TRY(cudaMemcpyAsync(out_status, E.scalars.p, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return B200_OK;
}

// GroupProjective images (host) -> compressed encodings (host): kind 0 = BLS12-377 G2 (288 B -> 96 B), 3 = G1 (144 B -> 48 B)
int encode_points_host(Engine &E, int kind, const void *images, size_t n, void *out_bytes) {
    if (kind != 0 && kind != 3) return fail(B200_ERR_ARG, "unknown point kind %d", kind);
    if (n == 0) return B200_OK;
    cudaStream_t st = E.stream;
    int rc;
    const size_t enc = kind == 3 ? 48 : 96;
    if ((rc = E.h2d_bases.reserve(n * 3 * enc)) || (rc = E.native_bases.reserve(n * enc))) return rc;
    CUDA_TRY(cudaMemcpyAsync(E.h2d_bases.p, images, n * 3 * enc, cudaMemcpyHostToDevice, st));
    const unsigned blocks = (unsigned)ceil_div(n, 64);
    if (kind == 3) k_g1_377_compress<<<blocks, 64, 0, st>>>(E.h2d_bases.as<JacobianMem<CFq>>(), (uint32_t)n, E.native_bases.as<uint32_t>());
    else k_g2_377_compress<<<blocks, 64, 0, st>>>(E.h2d_bases.as<JacobianMem<CFq2>>(), (uint32_t)n, E.native_bases.as<uint32_t>());
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpyAsync(out_bytes, E.native_bases.p, n * enc, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return B200_OK;
}

// ---- public inputs of the epoch SNARK ---------------------------------------------------------------------
// hash_first_last_epoch_block + pack (crates/epoch-snark/src/api/verifier.rs:30-33): decodes and checks both blocks' keys
// on the device, aggregates the last block's keys there, hashes and packs on the host.  *ok = 0 (with a reason) when
// the reference's EpochBlock::try_from / encoding would fail.
static const char *decode_reason(int st) {
    return st == DECODE_NOT_IN_SUBGROUP ? "a point is not in the prime-order subgroup"
           : st == DECODE_NOT_ON_CURVE  ? "an x coordinate is not on the curve"
           : st == DECODE_INFINITY      ? "a validator public key is the point at infinity"
                                        : "a coordinate is not below the field modulus";
}

// asynchronous half: uploads, decodes, checks and aggregates the keys on `st`; results stay in E.g2_packed
struct KeyWork {
    size_t n1 = 0, n2 = 0;
    int *d_status = nullptr;
    uint32_t *d_agg = nullptr;
};
static int epoch_keys_launch(Engine &E, const EpochBlockFFI &first, const EpochBlockFFI &last, KeyWork *kw, cudaStream_t st) {
    const size_t n1 = first.pubkeys_num, n2 = last.pubkeys_num, nkeys = n1 + n2;
    int rc;
    const size_t status_off = (nkeys + 1) * 192, agg_off = status_off + (nkeys + 1) * sizeof(int);
    if ((rc = E.h2d_g2.reserve((nkeys + 1) * 96)) || (rc = E.g2_packed.reserve(agg_off + 32 * sizeof(uint32_t)))) return rc;
    char *d_src = E.h2d_g2.as<char>(), *d_pts = E.g2_packed.as<char>();
    kw->n1 = n1;
    kw->n2 = n2;
    kw->d_status = reinterpret_cast<int *>(d_pts + status_off);
    kw->d_agg = reinterpret_cast<uint32_t *>(d_pts + agg_off);
    if (n1) CUDA_TRY(cudaMemcpyAsync(d_src, first.pubkeys, 96 * n1, cudaMemcpyHostToDevice, st));
    if (n2) CUDA_TRY(cudaMemcpyAsync(d_src + 96 * n1, last.pubkeys, 96 * n2, cudaMemcpyHostToDevice, st));
    if ((rc = decode_points(0, d_src, nkeys, 1, d_pts, kw->d_status, st))) return rc;      // G2Affine::deserialize per key
    k_g2_377_aggregate_emit<<<1, 32, 0, st>>>(reinterpret_cast<const AffineMem<CFq2> *>(d_pts + n1 * 192), kw->d_status + n1,
                                              (uint32_t)n2, kw->d_agg);
    LAUNCH_CHECK();
    return B200_OK;
}
// host half, after the stream has been synchronised: statuses -> encodings -> hashes -> packed inputs
static int epoch_keys_finish(const KeyWork &kw, const EpochBlockFFI &first, const EpochBlockFFI &last, std::vector<uint64_t> *inputs,
                             int *ok, std::string *why) {
    *ok = 0;
    auto reject = [&](const char *msg) {
        if (why) *why = msg;
        return B200_OK;
    };
    const size_t nkeys = kw.n1 + kw.n2;
    std::vector<int> status(nkeys);
    uint32_t agg[26];
    if (nkeys) CUDA_TRY(cudaMemcpy(status.data(), kw.d_status, nkeys * sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(agg, kw.d_agg, sizeof(agg), cudaMemcpyDeviceToHost));
    // an infinite validator key has no CIP22 encoding worth reproducing: rejected like an invalid one
    for (int s : status)
        if (s != DECODE_OK) return reject(decode_reason(s));
    if (agg[25]) return reject("the aggregated public key is the point at infinity");
    Bits enc_first, enc_last, hashes;
    push_block_cip22(enc_first, first, true);
    push_block_cip22(enc_last, last, false);
    push_public_key(enc_last, reinterpret_cast<const uint8_t *>(agg), agg[24] != 0);
    hash_bits(enc_first, hashes);
    hash_bits(enc_last, hashes);
    *inputs = pack_376(hashes);
    *ok = 1;
    return B200_OK;
}

int epoch_public_inputs(Engine &E, const EpochBlockFFI &first, const EpochBlockFFI &last, std::vector<uint64_t> *inputs, int *ok,
                        std::string *why) {
    *ok = 0;
    if ((first.pubkeys_num && !first.pubkeys) || (last.pubkeys_num && !last.pubkeys) || first.pubkeys_num + last.pubkeys_num > (1u << 20)) {
        if (why) *why = "null or oversized public-key array";
        return B200_OK;
    }
    KeyWork kw;
    int rc;
    if ((rc = epoch_keys_launch(E, first, last, &kw, E.stream))) return rc;
    CUDA_TRY(cudaStreamSynchronize(E.stream));
    return epoch_keys_finish(kw, first, last, inputs, ok, why);
}

// ---- verify ----------------------------------------------------------------------------------------------
// *out_ok = 1 iff the reference's `verify` would return true.  A non-zero return code is an engine failure
// (no device, CUDA error); malformed inputs are *out_ok = 0 with rc = 0, as the reference turns them into `false`.
int epoch_verify(Engine &E, const uint8_t *vk, size_t vk_len, const uint8_t *proof, size_t proof_len,
                 const EpochBlockFFI &first, const EpochBlockFFI &last, int *out_ok, std::string *why) {
    *out_ok = 0;
    auto reject = [&](const char *msg) {
        if (why) *why = msg;
        return B200_OK;
    };
    int rc, ok = 0;
    if ((first.pubkeys_num && !first.pubkeys) || (last.pubkeys_num && !last.pubkeys) || first.pubkeys_num + last.pubkeys_num > (1u << 20))
        return reject("null or oversized public-key array");
    // VerifyingKey<BW6_761>: alpha_g1 | beta_g2 | gamma_g2 | delta_g2 | u64 len | gamma_abc_g1[len]; Proof: A | B | C
    if (!vk || vk_len < 392) return reject("verifying key shorter than its fixed part");
    uint64_t nabc = 0;
    for (int i = 0; i < 8; i++) nabc |= (uint64_t)vk[384 + i] << (8 * i);
    if (nabc > (1u << 20) || vk_len < 392 + 96 * nabc) return reject("verifying key shorter than its gamma_abc list");
    if (!proof || proof_len < 288) return reject("proof shorter than 288 bytes");
    const size_t nbw = 8 + nabc;
    // compressed staging: [A, A (slot of g_ic), C, alpha] [B, gamma, delta, beta] [gamma_abc ...]
    std::vector<uint8_t> host(nbw * 96);
    const uint8_t *order[8] = {proof, proof, proof + 192, vk, proof + 96, vk + 192, vk + 288, vk + 96};
    for (int i = 0; i < 8; i++) memcpy(host.data() + 96 * i, order[i], 96);
    memcpy(host.data() + 96 * 8, vk + 392, 96 * nabc);
    cudaStream_t st = E.stream, side = E.pipe_stream[1];
    const size_t status_off = nbw * 192;
    if ((rc = E.h2d_bases.reserve(host.size())) || (rc = E.native_bases.reserve(status_off + nbw * sizeof(int)))) return rc;
    char *d_src = E.h2d_bases.as<char>(), *d_pts = E.native_bases.as<char>();
    int *d_status = reinterpret_cast<int *>(d_pts + status_off);
    // the two decodings are independent chains of single-thread field arithmetic: the validator keys (BLS12-377 G2) go to a
    // side stream, the key and the proof (BW6-761) stay on the main one
    CUDA_TRY(cudaEventRecord(E.ev_fork, st));
    CUDA_TRY(cudaStreamWaitEvent(side, E.ev_fork, 0));
    KeyWork kw;
    if ((rc = epoch_keys_launch(E, first, last, &kw, side))) return rc;
    CUDA_TRY(cudaEventRecord(E.ev_join, side));
    CUDA_TRY(cudaMemcpyAsync(d_src, host.data(), host.size(), cudaMemcpyHostToDevice, st));
    // G1Affine / G2Affine::deserialize over BW6-761: on the curve, in the subgroup (one launch: records 4..7 are G2)
    k_bw6_decompress<<<ceil_div(nbw, 64), 64, 0, st>>>(reinterpret_cast<const uint32_t *>(d_src), (uint32_t)nbw, 4u, 8u,
                                                       reinterpret_cast<AffineMem<Fq761> *>(d_pts), d_status);
    LAUNCH_CHECK();
    if ((rc = subgroup_check<Fq761, Fq377Params>(d_pts, nbw, d_status, st))) return rc;
    CUDA_TRY(cudaStreamWaitEvent(st, E.ev_join, 0));
    std::vector<int> status(nbw);
    CUDA_TRY(cudaMemcpyAsync(status.data(), d_status, nbw * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    std::vector<uint64_t> inputs;
    if ((rc = epoch_keys_finish(kw, first, last, &inputs, &ok, why))) return rc;
    if (!ok) return B200_OK;
    // an infinite element of the key / proof is a legal encoding: it decodes and simply fails the pairing check
    for (int s : status)
        if (s != DECODE_OK && s != DECODE_INFINITY) return reject(decode_reason(s));
    if (inputs.size() / 6 + 1 != nabc) return reject("malformed verifying key: public-input count does not match gamma_abc");
    std::vector<uint64_t> scal(nabc * 6, 0);
    scal[0] = 1;
    memcpy(scal.data() + 6, inputs.data(), inputs.size() * sizeof(uint64_t));
    if ((rc = E.scalars.reserve(nabc * 48))) return rc;
    CUDA_TRY(cudaMemcpyAsync(E.scalars.p, scal.data(), nabc * 48, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if ((rc = bw6_groth16_verify_core(E, d_pts, nabc, E.scalars.p, out_ok))) return rc;
    if (!*out_ok && why) *why = "pairing check failed";
    return B200_OK;
}

}  // namespace b200
