// Batched hash-to-G1 on the device: the step that precedes the multi-pairing in the
// reference's verification flows (crates/bls-crypto/src/bls/signature.rs:111-114).
//
//   crates/bls-crypto/src/hashers/direct.rs:20-80            DirectHasher: Blake2s CRH, Blake2Xs-style XOF
//   crates/bls-crypto/src/hashers/composite.rs:15-95         CompositeHasher: Bowe-Hopwood CRH (window 93 x 560 over
//                                                            ed-on-bw6-761, ChaCha20-seeded generators) + the same XOF
//   crates/bls-crypto/src/hash_to_curve/try_and_increment.rs:84-139        counter | extra | message, `compat` bit rule
//   crates/bls-crypto/src/hash_to_curve/try_and_increment_cip22.rs:60-134  CRH(message) once, XOF per counter
//   crates/bls-crypto/src/hash_to_curve/mod.rs:146-156       from_random_bytes (flags, masking, get_point_from_x)
//
// One warp per message runs the whole try-and-increment loop without host round trips:
//   * the Bowe-Hopwood sum is a table walk -- every 3-bit chunk selects (+-)(1..4) * 16^j * G_w, precomputed once per
//     process in affine form (30 MB in HBM); the 32 lanes add strided chunks (7 products per mixed addition) and a
//     shuffle tree folds them.  In the non-CIP22 flow the counter only touches chunks 0..2, so the rest of the sum is
//     computed once and each attempt adds three points;
//   * the 32 lanes then try 32 consecutive counters at once -- Blake2s (CRH of the direct hasher and both XOF blocks),
//     candidate decoding and the Tonelli-Shanks square root per lane; a ballot picks the lowest successful counter and
//     the warp multiplies that point by the cofactor with quad-cooperative XYZZ steps.
// The generators are derived at first use exactly as `setup_crh` does (composite.rs:53-72): ChaCha20 on the host
// (byte work; the stream positions do not depend on curve arithmetic), the candidate points tested on the device.
#include "codec.cuh"
#include "engine.cuh"

namespace b200 {

using HFq = Fq377;

enum : int { BH_WINDOW = 93, BH_WINDOWS = 560, BH_CHUNKS = BH_WINDOW * BH_WINDOWS, BH_SETUP_ATTEMPTS = 2048 };
enum : uint32_t { HASH_FAILED = 0xffffffffu };
enum : int { HASH_DEFAULT_OCC = 8 };

struct EdExt {
    HFq x, y, z, t;
};
struct alignas(16) EdExtMem {
    HFq::Mem x, y, z, t;
};
struct alignas(16) EdTabMem {                         // affine x, y and d x y
    HFq::Mem x, y, td;
};
struct HashMsg {                                      // bytes at data + off: extra_data | message
    uint32_t off, msg_len, extra_len, pad;
};

B200_DEV HFq ed_coeff_d() {                           // ed-on-bw6-761: -x^2 + y^2 = 1 + 79743 x^2 y^2
    uint32_t w[12] = {79743u, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    return fp_from_canonical<HFq>(w);
}
B200_DEV EdExt ed_identity() { return {HFq::zero(), HFq::one(), HFq::one(), HFq::zero()}; }

// unified (complete: a = -1 is a square, d is not) extended-coordinate addition, add-2008-hwcd
__device__ __noinline__ EdExt ed_add(EdExt p, EdExt q, HFq d) {
    HFq A = p.x * q.x, B = p.y * q.y, C = d * p.t * q.t, D = p.z * q.z;
    HFq E = (p.x + p.y) * (q.x + q.y) - A - B, F = D - C, G = D + C, H = B + A;
    return {E * F, G * H, F * G, E * H};
}
// the same with an affine table entry (z2 = 1, d t2 precomputed)
__device__ __noinline__ EdExt ed_madd(EdExt p, HFq x2, HFq y2, HFq td2) {
    HFq A = p.x * x2, B = p.y * y2, C = p.t * td2;
    HFq E = (p.x + p.y) * (x2 + y2) - A - B, F = p.z - C, G = p.z + C, H = B + A;
    return {E * F, G * H, F * G, E * H};
}

// ---- setup of the CRH generators ---------------------------------------------------------------------------
// `Standard` sampler of a twisted-Edwards point (ark-ec): x drawn as raw Montgomery limbs, y from x with the sign
// bit, times the cofactor 8.  cand: 13 words per attempt (12 limbs, greatest).
__global__ void __launch_bounds__(64) k_bh_candidates(const uint32_t *__restrict__ cand, uint32_t n, AffineMem<HFq> *__restrict__ out,
                                                      int *__restrict__ ok) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    HFq x;
#pragma unroll
    for (int k = 0; k < 12; k++) x.l[k] = cand[13 * i + k];
    const bool greatest = cand[13 * i + 12] != 0;
    const HFq one = HFq::one(), d = ed_coeff_d();
    HFq x2 = x.sqr(), den = d * x2 - one, y;
    ok[i] = 0;
    if (den.is_zero()) return;
    if (!fq377_sqrt((x2 + one).neg() * den.inv(), &y)) return;
    if (fp_canonical_over_half(fp_to_canonical(y)) != greatest) y = y.neg();
    EdExt p = {x, y, one, x * y};
#pragma unroll 1
    for (int k = 0; k < 3; k++) p = ed_add(p, p, d);
    HFq zi = p.z.inv();
    out[i] = {(p.x * zi).store(), (p.y * zi).store()};
    ok[i] = 1;
}
// window w: 16^j G_w for j < 93 (bowe_hopwood create_generators: four doublings between consecutive generators)
__global__ void __launch_bounds__(32) k_bh_powers(const AffineMem<HFq> *__restrict__ cand, const uint32_t *__restrict__ pick,
                                                  EdExtMem *__restrict__ powers) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= BH_WINDOWS) return;
    const HFq d = ed_coeff_d();
    HFq gx = HFq::load(cand[pick[w]].x), gy = HFq::load(cand[pick[w]].y);
    EdExt g = {gx, gy, HFq::one(), gx * gy};
#pragma unroll 1
    for (int j = 0; j < BH_WINDOW; j++) {
        powers[w * BH_WINDOW + j] = {g.x.store(), g.y.store(), g.z.store(), g.t.store()};
#pragma unroll 1
        for (int k = 0; k < 4; k++) g = ed_add(g, g, d);
    }
}
// chunk c: the four multiples a chunk can select, affine
__global__ void __launch_bounds__(64) k_bh_table(const EdExtMem *__restrict__ powers, EdTabMem *__restrict__ table) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= BH_CHUNKS) return;
    const HFq d = ed_coeff_d();
    EdExt g = {HFq::load(powers[c].x), HFq::load(powers[c].y), HFq::load(powers[c].z), HFq::load(powers[c].t)};
    EdExt m = g;
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        HFq zi = m.z.inv(), x = m.x * zi, y = m.y * zi;
        table[4 * c + k] = {x.store(), y.store(), (d * x * y).store()};
        m = ed_add(m, g, d);
    }
}

// ---- Blake2s on the device (RFC 7693, unkeyed) over a virtual concatenation counter | a | b -----------------
struct ByteSrc {
    uint32_t has_c;
    uint8_t c;
    const uint8_t *a;
    uint32_t la;
    const uint8_t *b;
    uint32_t lb;
    B200_DEV uint32_t len() const { return has_c + la + lb; }
    B200_DEV uint32_t get(uint32_t i) const {
        if (i < has_c) return c;
        i -= has_c;
        if (i < la) return a[i];
        i -= la;
        return i < lb ? b[i] : 0u;
    }
};
static __device__ const uint32_t D_BLAKE2S_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                                    0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static __device__ const uint8_t D_BLAKE2S_SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};

B200_DEV uint32_t rotr32d(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

__device__ __noinline__ void blake2s_compress_dev(uint32_t *h, const uint32_t *m, uint32_t t, bool last) {
    uint32_t v[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        v[i] = h[i];
        v[8 + i] = D_BLAKE2S_IV[i];
    }
    v[12] ^= t;
    if (last) v[14] = ~v[14];
#define B2S_G(a, b, c, d, x, y)                 \
    v[a] = v[a] + v[b] + (x);                   \
    v[d] = rotr32d(v[d] ^ v[a], 16);            \
    v[c] = v[c] + v[d];                         \
    v[b] = rotr32d(v[b] ^ v[c], 12);            \
    v[a] = v[a] + v[b] + (y);                   \
    v[d] = rotr32d(v[d] ^ v[a], 8);             \
    v[c] = v[c] + v[d];                         \
    v[b] = rotr32d(v[b] ^ v[c], 7);
#pragma unroll 1
    for (int r = 0; r < 10; r++) {
        const uint8_t *s = D_BLAKE2S_SIGMA[r];
        B2S_G(0, 4, 8, 12, m[s[0]], m[s[1]]);
        B2S_G(1, 5, 9, 13, m[s[2]], m[s[3]]);
        B2S_G(2, 6, 10, 14, m[s[4]], m[s[5]]);
        B2S_G(3, 7, 11, 15, m[s[6]], m[s[7]]);
        B2S_G(0, 5, 10, 15, m[s[8]], m[s[9]]);
        B2S_G(1, 6, 11, 12, m[s[10]], m[s[11]]);
        B2S_G(2, 7, 8, 13, m[s[12]], m[s[13]]);
        B2S_G(3, 4, 9, 14, m[s[14]], m[s[15]]);
    }
#undef B2S_G
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[8 + i];
}
// h <- digest words; p0..p3 are the first four words of the parameter block, pers the personalisation words
__device__ __noinline__ void blake2s_dev(uint32_t *h, const ByteSrc &src, uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3,
                                         uint32_t pers0, uint32_t pers1) {
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = D_BLAKE2S_IV[i];
    h[0] ^= p0;
    h[1] ^= p1;
    h[2] ^= p2;
    h[3] ^= p3;
    h[6] ^= pers0;
    h[7] ^= pers1;
    const uint32_t n = src.len();
    uint32_t off = 0, m[16];
#pragma unroll 1
    for (;;) {
        const bool last = n - off <= 64;
#pragma unroll 1
        for (int k = 0; k < 16; k++) {
            uint32_t w = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint32_t i = off + 4 * k + b;
                if (i < n) w |= src.get(i) << (8 * b);
            }
            m[k] = w;
        }
        blake2s_compress_dev(h, m, last ? n : off + 64, last);
        if (last) break;
        off += 64;
    }
}

// ---- Bowe-Hopwood sum ----------------------------------------------------------------------------------------
// bit i of the CRH input counter | body (the counter present iff has_c), zero beyond the end (padding to a chunk)
B200_DEV uint32_t bh_bit(uint32_t i, uint32_t has_c, uint32_t counter, const uint8_t *body, uint32_t body_len) {
    if (has_c) {
        if (i < 8) return (counter >> i) & 1u;
        i -= 8;
    }
    return (i >> 3) < body_len ? (body[i >> 3] >> (i & 7)) & 1u : 0u;
}
B200_DEV EdExt bh_add_chunk(EdExt acc, uint32_t c, uint32_t has_c, uint32_t counter, const uint8_t *body, uint32_t body_len,
                            const EdTabMem *__restrict__ table) {
    const uint32_t b0 = bh_bit(3 * c, has_c, counter, body, body_len), b1 = bh_bit(3 * c + 1, has_c, counter, body, body_len),
                   b2 = bh_bit(3 * c + 2, has_c, counter, body, body_len);
    const EdTabMem e = ldg_mem(table + 4 * (size_t)c + b0 + 2 * b1);
    return ed_madd(acc, HFq::load(e.x).cneg(b2), HFq::load(e.y), HFq::load(e.td).cneg(b2));
}
// chunks [first, count) of the input summed over the warp; the total lands on lane 0
B200_DEV EdExt bh_warp_sum(uint32_t first, uint32_t count, uint32_t has_c, const uint8_t *body, uint32_t body_len,
                           const EdTabMem *__restrict__ table, const HFq &d, int lane) {
    EdExt acc = ed_identity();
#pragma unroll 1
    for (uint32_t c = first + lane; c < count; c += 32) acc = bh_add_chunk(acc, c, has_c, 0, body, body_len, table);
#pragma unroll 1
    for (int off = 16; off >= 1; off >>= 1) {
        EdExt o = {acc.x.shfl(0xffffffffu, lane ^ off), acc.y.shfl(0xffffffffu, lane ^ off), acc.z.shfl(0xffffffffu, lane ^ off),
                   acc.t.shfl(0xffffffffu, lane ^ off)};
        if ((lane & off) == 0) acc = ed_add(acc, o, d);
    }
    return acc;
}
// x coordinate of the CRH point as the 48 canonical little-endian bytes `h.x.serialize` writes (composite.rs:78-84)
B200_DEV void ed_x_bytes(const EdExt &p, uint32_t *out12) {
    HFq x = fp_to_canonical(p.x * p.z.inv());
#pragma unroll
    for (int k = 0; k < 12; k++) out12[k] = x.l[k];
}

// ---- square root by a windowed discrete logarithm in the 2-Sylow subgroup --------------------------------------
// p - 1 = 2^46 t.  With x0 = a^((t+1)/2) and b = a^t = root^e (root of order 2^46; e is even iff a is a square),
// sqrt(a) = x0 root^(-e/2).  Tonelli-Shanks (codec.cuh) finds e one bit per round at ~46 squarings a round, and
// with 32 lanes on different branches the warp pays the longest chain (~1000 squarings).  Here e is read six bits
// at a time: digit i is identified by raising b (digits below i already cancelled) to 2^(46 - 6(i+1)) and looking
// the result up among the 64 powers of root^(2^40); x and b are then corrected from a table.  154 squarings, 16
// products and 8 look-ups per root, the same on every lane.
enum : int { TS_S = Fq377Params::TWO_ADICITY, TS_W = 6, TS_WINDOWS = (TS_S + TS_W - 1) / TS_W, TS_DIGITS = 1 << TS_W };
struct alignas(16) SqrtTables {
    HFq::Mem lookup[TS_DIGITS];                          // (root^(2^(46 - 6)))^d
    HFq::Mem half[TS_WINDOWS][TS_DIGITS];                // root^(-d 2^(6 i) / 2)   (window 0: even d only)
};
// one thread per (window, digit) plus one block row for the look-up powers
__global__ void __launch_bounds__(TS_DIGITS) k_sqrt_tables(SqrtTables *__restrict__ t) {
    const int d = threadIdx.x, i = blockIdx.x;
    HFq root;
#pragma unroll
    for (int k = 0; k < 12; k++) root.l[k] = Fq377Params::root(k);
    HFq base;
    int e;
    if (i == TS_WINDOWS) {                               // look-up row
        base = root;
        for (int k = 0; k < TS_S - TS_W; k++) base = base.sqr();
        e = d;
    } else {
        base = root.inv();
        for (int k = 0; k < TS_W * i - 1; k++) base = base.sqr();
        e = i == 0 ? d / 2 : d;
    }
    HFq acc = HFq::one();
#pragma unroll 1
    for (int k = 0; k < e; k++) acc = acc * base;
    if (i == TS_WINDOWS) t->lookup[d] = acc.store();
    else t->half[i][d] = acc.store();
}
__device__ __noinline__ bool fq377_sqrt_tab(HFq a, HFq *out, const SqrtTables *__restrict__ t) {
    if (a.is_zero()) {
        *out = a;
        return true;
    }
    HFq z = fp_pow_words<HFq>(a, FQ377_TS_EXP, FQ377_TS_EXP_BITS);              // a^((t - 1) / 2)
    HFq x = a * z;                                                              // a^((t + 1) / 2)
    HFq b = x * z;                                                              // a^t = root^e
#pragma unroll 1
    for (int i = 0; i < TS_WINDOWS; i++) {
        const int width = (i + 1) * TS_W <= TS_S ? TS_W : TS_S - i * TS_W;      // the last window holds 4 bits
        HFq c = b;
#pragma unroll 1
        for (int k = 0; k < TS_S - i * TS_W - width; k++) c = c.sqr();
        // c = lookup[d << (6 - width)]
        int d = -1;
#pragma unroll 1
        for (int j = 0; j < (1 << width) && d < 0; j++) {
            const HFq::Mem &m = t->lookup[j << (TS_W - width)];
            if (m.w[0] != c.l[0] || m.w[1] != c.l[1]) continue;
            if (HFq::load(m) == c) d = j;
        }
        if (d < 0 || (i == 0 && (d & 1))) return false;   // e odd: a is not a square (d < 0 cannot happen)
        if (d == 0) continue;
        const HFq h = HFq::load(t->half[i][d]);
        x = x * h;
        b = b * h.sqr();
    }
    *out = x;
    return x.sqr() == a;
}

// ---- the try-and-increment loop ------------------------------------------------------------------------------
enum : int { HASHER_DIRECT = 0, HASHER_COMPOSITE = 1 };
enum : int { HASH_FLAG_COMPAT = 1, HASH_FLAG_CIP22 = 2, HASH_FLAG_CRH_ONLY = 4 };

// cofactor of BLS12-377 G1, (x - 1)^2 / 3 = 0x170b5d44300000000000000000000000 (125 bits)
static __device__ const uint32_t G1_COFACTOR[4] = {0x00000000u, 0x00000000u, 0x30000000u, 0x170b5d44u};

// One counter of the try-and-increment loop on one lane: CRH (unless hashed once) -> XOF -> from_random_bytes ->
// y from x.  Returns false when the reference would move on to the next counter before the cofactor step.
__device__ __noinline__ bool try_counter(uint32_t c, bool cip22, bool composite, bool compat, const uint8_t *extra, uint32_t extra_len,
                                         uint32_t body_len, const uint32_t *inner, const EdExt &rest, const EdTabMem *__restrict__ table,
                                         const SqrtTables *__restrict__ sqrt_tables, uint32_t pers0, uint32_t pers1, HFq &x, HFq &y) {
    const uint32_t HB = 64;                               // hash_length(48), hash_to_curve/mod.rs:19-23
    const uint32_t inner_len = composite ? 48 : 32;
    uint32_t crh[12];
    ByteSrc xin;
    if (cip22) {
        xin = {1, (uint8_t)c, extra, extra_len, reinterpret_cast<const uint8_t *>(inner), inner_len};
    } else {
        if (composite) {
            EdExt s = rest;
            for (uint32_t k = 0; k < 3; k++) s = bh_add_chunk(s, k, 1, c, extra, body_len, table);
            ed_x_bytes(s, crh);
        } else {
            ByteSrc src = {1, (uint8_t)c, extra, body_len, nullptr, 0};
            blake2s_dev(crh, src, 0x01010020u, 0, 0, HB, pers0, pers1);
        }
        xin = {0, 0, reinterpret_cast<const uint8_t *>(crh), inner_len, nullptr, 0};
    }
    // XOF (direct.rs:41-79): two 32-byte blocks, fanout 0, depth 0, leaf 32, inner 32, node offset = block index
    uint32_t h0[8], h1[8], w[12];
    blake2s_dev(h0, xin, 32u, 32u, 0u, HB | (32u << 24), pers0, pers1);
    blake2s_dev(h1, xin, 32u, 32u, 1u, HB | (32u << 24), pers0, pers1);
    for (int k = 0; k < 8; k++) w[k] = h0[k];
    for (int k = 0; k < 4; k++) w[8 + k] = h1[k];
    // from_random_bytes: byte 47 carries the flags (bit 7 sign, bit 6 infinity); `compat` moves bit 1 into bit 7
    const bool positive = compat ? (w[11] >> 25) & 1u : (w[11] >> 31) & 1u;
    const bool infinity = (w[11] >> 30) & 1u;
    w[11] &= 0x01ffffffu;
    if (!fp_words_lt_modulus<HFq>(w)) return false;
    x = fp_from_canonical<HFq>(w);
    // (0, infinity flag) is the zero point, whose cofactor multiple is zero -> next counter
    if (x.is_zero() && infinity) return false;
    if (!fq377_sqrt_tab(x.sqr() * x + HFq::one(), &y, sqrt_tables)) return false;
    if (fp_canonical_over_half(fp_to_canonical(y)) != positive) y = y.neg();
    return true;
}

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(32, MIN_BLOCKS) k_hash_to_g1(const uint8_t *__restrict__ data, const HashMsg *__restrict__ msgs, uint32_t n,
                                                   int hasher, int flags, uint32_t pers0, uint32_t pers1,
                                                   const EdTabMem *__restrict__ table, const SqrtTables *__restrict__ sqrt_tables,
                                                   JacobianMem<HFq> *__restrict__ out,
                                                   uint32_t *__restrict__ attempts, uint32_t *__restrict__ crh_out) {
    const uint32_t i = blockIdx.x;
    const int lane = threadIdx.x;
    if (i >= n) return;
    const HashMsg hm = msgs[i];
    const uint8_t *extra = data + hm.off, *message = extra + hm.extra_len;
    const bool cip22 = flags & HASH_FLAG_CIP22, compat = flags & HASH_FLAG_COMPAT, composite = hasher == HASHER_COMPOSITE;
    const bool once = cip22 || (flags & HASH_FLAG_CRH_ONLY);   // the CRH covers the message alone
    const uint32_t HB = 64;                               // hash_length(48), hash_to_curve/mod.rs:19-23
    const HFq d = ed_coeff_d();
    __shared__ uint32_t inner[12];                        // CIP22: the CRH of the message, hashed once
    EdExt rest = ed_identity();                           // composite, counter-dependent flow: chunks 3.. of the sum
    const uint32_t body_len = hm.extra_len + hm.msg_len;
    if (composite) {
        if (once) {
            EdExt s = bh_warp_sum(0, (8 * hm.msg_len + 2) / 3, 0, message, hm.msg_len, table, d, lane);
            if (lane == 0) ed_x_bytes(s, inner);
        } else {
            rest = bh_warp_sum(3, (8 * (body_len + 1) + 2) / 3, 1, extra, body_len, table, d, lane);
        }
    } else if (once && lane == 0) {
        ByteSrc src = {0, 0, message, hm.msg_len, nullptr, 0};
        uint32_t h[8];
        blake2s_dev(h, src, 0x01010020u, 0, 0, HB, pers0, pers1);
        for (int k = 0; k < 8; k++) inner[k] = h[k];
    }
    __syncwarp();
    if (flags & HASH_FLAG_CRH_ONLY) {                     // Hasher::crh of the message alone (48 or 32 bytes)
        if (lane == 0)
            for (int k = 0; k < 12; k++) crh_out[12 * (size_t)i + k] = (composite || k < 8) ? inner[k] : 0u;
        return;
    }
    // The 32 lanes try 32 consecutive counters at once (a lone lane would cost the same issue slots); the lowest
    // successful counter wins, exactly as the sequential loop of the reference would find it.
    const unsigned full = 0xffffffffu;
    if (composite && !cip22)
        rest = {rest.x.shfl(full, 0), rest.y.shfl(full, 0), rest.z.shfl(full, 0), rest.t.shfl(full, 0)};
    const Quad Q;
    bool done = false;
#pragma unroll 1
    for (uint32_t first = 0; first < 255 && !done; first += 32) {          // NUM_TRIES, try_and_increment.rs:26
        const uint32_t c = first + lane;
        bool ok = false;
        HFq x = HFq::zero(), y = HFq::zero();
        if (c < 255)
            ok = try_counter(c, cip22, composite, compat, extra, hm.extra_len, body_len, inner, rest, table, sqrt_tables, pers0, pers1, x, y);
        unsigned winners = __ballot_sync(full, ok);
#pragma unroll 1
        while (winners && !done) {
            const int win = __ffs(winners) - 1;
            winners &= winners - 1;
            // scale_by_cofactor on every quad of the warp at once (quad-cooperative XYZZ steps, ec.cuh)
            const XYZZ<HFq> base = {x.shfl(full, win), y.shfl(full, win), HFq::one(), HFq::one()};
            XYZZ<HFq> acc = base;
#pragma unroll 1
            for (int b = 123; b >= 0; b--) {
                quad_dbl(Q, acc);
                if ((G1_COFACTOR[b >> 5] >> (b & 31)) & 1u) quad_add(Q, acc, base);
            }
            if (acc.is_inf()) continue;                   // scaled.is_zero(): the reference moves on to the next counter
            if (lane == 0) {
                out[i] = acc.to_jacobian().to_ark();
                attempts[i] = first + win;
            }
            done = true;
        }
    }
    if (!done && lane == 0) attempts[i] = HASH_FAILED;
}

// The same loop for large batches, in two launches.
//  * k_hash_prepare: one warp per message computes the counter-independent part of the CRH (the Bowe-Hopwood sum of
//    chunks 3.. or, for CIP22, the inner hash) into global memory -- a table walk that wants every warp it can get.
//  * k_hash_to_g1_packed<PACK>: PACK messages per warp.  A group of 32 / PACK lanes owns one message and tries that
//    many counters per round; once every message of the warp has a candidate point the groups run the cofactor
//    multiplication side by side (quad-cooperative steps, each quad on its group's point).  Per message this is
//    1 / PACK of a warp's square-root rounds and cofactor multiplications instead of one each, at a longer latency
//    for a single message (so small batches keep k_hash_to_g1).
struct alignas(16) HashPrep {
    uint32_t inner[12];                                   // CIP22: the CRH of the message
    EdExtMem rest;                                        // composite, counter-dependent flow: chunks 3.. of the sum
};
__global__ void __launch_bounds__(32) k_hash_prepare(const uint8_t *__restrict__ data, const HashMsg *__restrict__ msgs, uint32_t n,
                                                     int hasher, int flags, uint32_t pers0, uint32_t pers1,
                                                     const EdTabMem *__restrict__ table, HashPrep *__restrict__ prep) {
    const uint32_t i = blockIdx.x;
    const int lane = threadIdx.x;
    if (i >= n) return;
    const bool cip22 = flags & HASH_FLAG_CIP22, composite = hasher == HASHER_COMPOSITE;
    const HashMsg hm = msgs[i];
    const uint8_t *extra = data + hm.off, *message = extra + hm.extra_len;
    const uint32_t body_len = hm.extra_len + hm.msg_len;
    const HFq d = ed_coeff_d();
    if (composite) {
        if (cip22) {
            EdExt s = bh_warp_sum(0, (8 * hm.msg_len + 2) / 3, 0, message, hm.msg_len, table, d, lane);
            if (lane == 0) ed_x_bytes(s, prep[i].inner);
        } else {
            EdExt r = bh_warp_sum(3, (8 * (body_len + 1) + 2) / 3, 1, extra, body_len, table, d, lane);
            if (lane == 0) prep[i].rest = {r.x.store(), r.y.store(), r.z.store(), r.t.store()};
        }
    } else if (cip22 && lane == 0) {
        ByteSrc src = {0, 0, message, hm.msg_len, nullptr, 0};
        uint32_t h[8];
        blake2s_dev(h, src, 0x01010020u, 0, 0, 64u, pers0, pers1);
        for (int k = 0; k < 8; k++) prep[i].inner[k] = h[k];
    }
}

template <int PACK>
__global__ void __launch_bounds__(32) k_hash_to_g1_packed(const uint8_t *__restrict__ data, const HashMsg *__restrict__ msgs, uint32_t n,
                                                          int hasher, int flags, uint32_t pers0, uint32_t pers1,
                                                          const EdTabMem *__restrict__ table, const SqrtTables *__restrict__ sqrt_tables,
                                                          const HashPrep *__restrict__ prep, JacobianMem<HFq> *__restrict__ out,
                                                          uint32_t *__restrict__ attempts) {
    constexpr int LPM = 32 / PACK;                        // lanes (= counters per round) per message
    const int lane = threadIdx.x, slot = lane / LPM, q = lane % LPM, group = lane - q;
    const bool cip22 = flags & HASH_FLAG_CIP22, compat = flags & HASH_FLAG_COMPAT, composite = hasher == HASHER_COMPOSITE;
    const unsigned full = 0xffffffffu;
    // group `slot` owns message i; next_c / have / finished are identical on its lanes
    const uint32_t i = blockIdx.x * PACK + slot;
    const bool live = i < n;
    const HashMsg hm = live ? msgs[i] : HashMsg{0, 0, 0, 0};
    const uint8_t *extra = data + hm.off;
    const uint32_t body_len = hm.extra_len + hm.msg_len;
    const HashPrep *mine_prep = prep + (live ? i : 0);
    EdExt rest = ed_identity();
    if (live && composite && !cip22)
        rest = {HFq::load(mine_prep->rest.x), HFq::load(mine_prep->rest.y), HFq::load(mine_prep->rest.z), HFq::load(mine_prep->rest.t)};
    const Quad Q;
    const HFq one = HFq::one();
    uint32_t next_c = 0, cand_c = 0;
    bool have = false, finished = !live, failed = false;
    HFq cx = HFq::zero(), cy = HFq::zero();
#pragma unroll 1
    for (;;) {
#pragma unroll 1
        for (;;) {                                        // rounds of LPM counters per message until each has a candidate
            const bool need = !finished && !have;
            if (!__any_sync(full, need)) break;
            const uint32_t c = next_c + q;
            bool ok = false;
            HFq x = HFq::zero(), y = HFq::zero();
            if (need && c < 255)                          // NUM_TRIES, try_and_increment.rs:26
                ok = try_counter(c, cip22, composite, compat, extra, hm.extra_len, body_len, mine_prep->inner, rest, table, sqrt_tables, pers0,
                                 pers1, x, y);
            const unsigned mine = (__ballot_sync(full, ok) >> group) & ((1u << LPM) - 1u);
            const int w = mine ? __ffs(mine) - 1 : 0;     // the lowest successful counter of the group
            const HFq wx = x.shfl(full, group + w), wy = y.shfl(full, group + w);
            if (need) {
                if (mine) {
                    cx = wx;
                    cy = wy;
                    cand_c = next_c + w;
                    have = true;
                } else {
                    next_c += LPM;
                    if (next_c >= 255) finished = failed = true;
                }
            }
        }
        if (!__any_sync(full, have)) break;
        // scale_by_cofactor, every group on its own candidate (groups without one carry the identity and fall through)
        const XYZZ<HFq> base = have ? XYZZ<HFq>{cx, cy, one, one} : XYZZ<HFq>::inf();
        XYZZ<HFq> acc = base;
#pragma unroll 1
        for (int b = 123; b >= 0; b--) {
            quad_dbl(Q, acc);
            if ((G1_COFACTOR[b >> 5] >> (b & 31)) & 1u) quad_add(Q, acc, base);
        }
        if (have) {
            have = false;
            if (acc.is_inf()) {                           // scaled.is_zero(): the reference moves on to the next counter
                next_c = cand_c + 1;
                if (next_c >= 255) finished = failed = true;
            } else {
                if (q == 0) {
                    out[i] = acc.to_jacobian().to_ark();
                    attempts[i] = cand_c;
                }
                finished = true;
            }
        }
    }
    if (live && failed && q == 0) attempts[i] = HASH_FAILED;
}

// ---- host side -----------------------------------------------------------------------------------------------
namespace {
struct ChaCha20 {                                         // rand_chacha ChaChaRng: 64-bit counter, stream 0, words in order
    uint32_t key[8], buf[16];
    uint64_t counter = 0;
    int pos = 16;
    static uint32_t rotl(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
    void block() {
        uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3], key[4], key[5], key[6],
                          key[7], (uint32_t)counter, (uint32_t)(counter >> 32), 0, 0};
        uint32_t w[16];
        memcpy(w, s, sizeof(w));
        auto qr = [&](int a, int b, int c, int d) {
            w[a] += w[b]; w[d] = rotl(w[d] ^ w[a], 16);
            w[c] += w[d]; w[b] = rotl(w[b] ^ w[c], 12);
            w[a] += w[b]; w[d] = rotl(w[d] ^ w[a], 8);
            w[c] += w[d]; w[b] = rotl(w[b] ^ w[c], 7);
        };
        for (int r = 0; r < 10; r++) {
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
        }
        for (int k = 0; k < 16; k++) buf[k] = w[k] + s[k];
        counter++;
        pos = 0;
    }
    uint32_t next_u32() {
        if (pos == 16) block();
        return buf[pos++];
    }
};
bool words_below_modulus(const uint32_t *w) {
    for (int k = 11; k >= 0; k--) {
        if (w[k] != Fq377Params::mod(k)) return w[k] < Fq377Params::mod(k);
    }
    return false;
}
}  // namespace

// CompositeHasher::setup_crh (composite.rs:53-72) -> the affine multiples table of every chunk position
static int ensure_bh_table(Engine &E, cudaStream_t st) {
    if (E.bh_ready) return B200_OK;
    const uint8_t personal[8] = {'U', 'L', '_', 'p', 'r', 'n', 'g', 's'};
    uint8_t seed[32];
    blake2s_personal(reinterpret_cast<const uint8_t *>("ULTRALIGHT PRNG SEED"), 20, personal, seed);
    ChaCha20 rng;
    memcpy(rng.key, seed, 32);
    // each sampler attempt: Fq::rand (six u64 limbs taken as the Montgomery form, top 7 bits shaved, redrawn while
    // >= p), then one bool; whether the attempt yields a point does not move the stream, so all are drawn up front
    std::vector<uint32_t> cand(13 * (size_t)BH_SETUP_ATTEMPTS);
    for (int a = 0; a < BH_SETUP_ATTEMPTS; a++) {
        uint32_t *w = &cand[13 * (size_t)a];
        do {
            for (int k = 0; k < 12; k++) w[k] = rng.next_u32();
            w[11] &= 0x01ffffffu;
        } while (!words_below_modulus(w));
        w[12] = rng.next_u32() >> 31;
    }
    int rc;
    const size_t cand_bytes = cand.size() * 4, pts_bytes = (size_t)BH_SETUP_ATTEMPTS * sizeof(AffineMem<HFq>),
                 ok_bytes = (size_t)BH_SETUP_ATTEMPTS * 4, pick_bytes = (size_t)BH_WINDOWS * 4,
                 pow_bytes = (size_t)BH_CHUNKS * sizeof(EdExtMem);
    Buffer tmp;
    if ((rc = tmp.reserve(cand_bytes + pts_bytes + ok_bytes + pick_bytes + pow_bytes)) ||
        (rc = E.bh_table.reserve(4 * (size_t)BH_CHUNKS * sizeof(EdTabMem))))
        return rc;
    char *base = tmp.as<char>();
    uint32_t *d_cand = reinterpret_cast<uint32_t *>(base);
    AffineMem<HFq> *d_pts = reinterpret_cast<AffineMem<HFq> *>(base + cand_bytes);
    int *d_ok = reinterpret_cast<int *>(base + cand_bytes + pts_bytes);
    uint32_t *d_pick = reinterpret_cast<uint32_t *>(base + cand_bytes + pts_bytes + ok_bytes);
    EdExtMem *d_pow = reinterpret_cast<EdExtMem *>(base + cand_bytes + pts_bytes + ok_bytes + pick_bytes);
    std::vector<int> ok(BH_SETUP_ATTEMPTS);
    std::vector<uint32_t> pick;
    auto body = [&]() -> int {
        CUDA_TRY(cudaMemcpyAsync(d_cand, cand.data(), cand_bytes, cudaMemcpyHostToDevice, st));
        k_bh_candidates<<<ceil_div(BH_SETUP_ATTEMPTS, 64), 64, 0, st>>>(d_cand, BH_SETUP_ATTEMPTS, d_pts, d_ok);
        LAUNCH_CHECK();
        CUDA_TRY(cudaMemcpyAsync(ok.data(), d_ok, ok_bytes, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        for (int a = 0; a < BH_SETUP_ATTEMPTS && (int)pick.size() < BH_WINDOWS; a++)
            if (ok[a]) pick.push_back((uint32_t)a);
        if ((int)pick.size() < BH_WINDOWS) return fail(B200_ERR_STATE, "CRH setup: %zu generators from %d attempts", pick.size(), BH_SETUP_ATTEMPTS);
        CUDA_TRY(cudaMemcpyAsync(d_pick, pick.data(), pick_bytes, cudaMemcpyHostToDevice, st));
        k_bh_powers<<<ceil_div(BH_WINDOWS, 32), 32, 0, st>>>(d_pts, d_pick, d_pow);
        LAUNCH_CHECK();
        k_bh_table<<<ceil_div(BH_CHUNKS, 64), 64, 0, st>>>(d_pow, E.bh_table.as<EdTabMem>());
        LAUNCH_CHECK();
        CUDA_TRY(cudaStreamSynchronize(st));
        return B200_OK;
    };
    rc = body();
    tmp.release();
    if (rc == B200_OK) E.bh_ready = true;
    return rc;
}

int hash_to_g1(Engine &E, int hasher, int flags, const uint8_t *domain, size_t domain_len, const b200_hash_input *inputs, size_t n,
               void *out, uint32_t *out_attempts) {
    if (hasher != HASHER_DIRECT && hasher != HASHER_COMPOSITE) return fail(B200_ERR_ARG, "unknown hasher id %d", hasher);
    if (domain_len > 8) return fail(B200_ERR_ARG, "domain of %zu bytes (at most 8: BLSError::DomainTooLarge)", domain_len);
    if (n == 0) return B200_OK;
    const bool cip22 = flags & HASH_FLAG_CIP22, crh_only = flags & HASH_FLAG_CRH_ONLY, composite = hasher == HASHER_COMPOSITE;
    cudaStream_t st = E.stream;
    std::vector<HashMsg> metas(n);
    size_t total = 0;
    for (size_t i = 0; i < n; i++) {
        const b200_hash_input &in = inputs[i];
        if ((in.message_len && !in.message) || (in.extra_data_len && !in.extra_data)) return fail(B200_ERR_ARG, "null pointer in input %zu", i);
        const size_t crh_len = (cip22 || crh_only) ? in.message_len : 1 + in.extra_data_len + in.message_len;
        if (composite && crh_len * 8 > (size_t)BH_CHUNKS * 3)   // bowe_hopwood::CRH::evaluate's input-size check
            return fail(B200_ERR_ARG, "input %zu: %zu bytes exceed the CRH capacity of %d bits", i, crh_len, BH_CHUNKS * 3);
        if (total + in.extra_data_len + in.message_len > 0xfffffff0u) return fail(B200_ERR_ARG, "inputs exceed 4 GiB");
        metas[i] = {(uint32_t)total, (uint32_t)in.message_len, (uint32_t)in.extra_data_len, 0};
        total += in.extra_data_len + in.message_len;
    }
    std::vector<uint8_t> blob(total + 16);
    for (size_t i = 0; i < n; i++) {
        if (inputs[i].extra_data_len) memcpy(&blob[metas[i].off], inputs[i].extra_data, inputs[i].extra_data_len);
        if (inputs[i].message_len) memcpy(&blob[metas[i].off + inputs[i].extra_data_len], inputs[i].message, inputs[i].message_len);
    }
    int rc;
    if (composite && (rc = ensure_bh_table(E, st))) return rc;
    if (!E.sqrt_ready) {
        if ((rc = E.sqrt_tables.reserve(sizeof(SqrtTables)))) return rc;
        k_sqrt_tables<<<TS_WINDOWS + 1, TS_DIGITS, 0, st>>>(E.sqrt_tables.as<SqrtTables>());
        LAUNCH_CHECK();
        E.sqrt_ready = true;
    }
    const size_t meta_off = (blob.size() + 15) & ~(size_t)15, out_off = meta_off + n * sizeof(HashMsg), att_off = out_off + n * 144,
                 crh_off = att_off + ((n * 4 + 15) & ~(size_t)15), prep_off = crh_off + n * 48;
    if ((rc = E.hash_ws.reserve(prep_off + n * sizeof(HashPrep)))) return rc;
    char *base = E.hash_ws.as<char>();
    uint8_t personal[8] = {0};
    if (domain_len) memcpy(personal, domain, domain_len);
    uint32_t pers[2];
    memcpy(pers, personal, 8);
    CUDA_TRY(cudaMemcpyAsync(base, blob.data(), blob.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(base + meta_off, metas.data(), n * sizeof(HashMsg), cudaMemcpyHostToDevice, st));
    // resident warps per SM: the loop is a chain of dependent 377-bit products, so throughput comes from warps in
    // flight; the register cap that pays for them is chosen by measurement (profiles/r1_hash_to_g1.md)
    static const int occ = getenv("B200_HASH_OCC") ? atoi(getenv("B200_HASH_OCC")) : HASH_DEFAULT_OCC;
    auto launch = [&](auto kernel) {
        kernel<<<(unsigned)n, 32, 0, st>>>(reinterpret_cast<const uint8_t *>(base), reinterpret_cast<const HashMsg *>(base + meta_off),
                                           (uint32_t)n, hasher, flags, pers[0], pers[1], E.bh_table.as<EdTabMem>(),
                                           E.sqrt_tables.as<SqrtTables>(), reinterpret_cast<JacobianMem<HFq> *>(base + out_off),
                                           reinterpret_cast<uint32_t *>(base + att_off), reinterpret_cast<uint32_t *>(base + crh_off));
    };
    // large batches: 4 or 8 messages per warp (fewer field products per message); small ones: one message per warp,
    // 32 counters at once (lowest latency).  B200_HASH_PACKED = 0 / 4 / 8 forces the choice.
    static const int packed_env = getenv("B200_HASH_PACKED") ? atoi(getenv("B200_HASH_PACKED")) : -1;
    const size_t wave = (size_t)E.sm_count * 8;           // resident warps of these kernels (register-limited)
    const int pack = crh_only ? 0 : packed_env >= 0 ? packed_env : n > 4 * wave ? 8 : n > wave ? 4 : 0;
    if (pack) {
        HashPrep *d_prep = reinterpret_cast<HashPrep *>(base + prep_off);
        const HashMsg *d_msgs = reinterpret_cast<const HashMsg *>(base + meta_off);
        if (cip22 || composite) {
            k_hash_prepare<<<(unsigned)n, 32, 0, st>>>(reinterpret_cast<const uint8_t *>(base), d_msgs, (uint32_t)n, hasher, flags, pers[0],
                                                       pers[1], E.bh_table.as<EdTabMem>(), d_prep);
            LAUNCH_CHECK();
        }
        auto launch_packed = [&](auto kernel, int per_warp) {
            kernel<<<(unsigned)((n + per_warp - 1) / per_warp), 32, 0, st>>>(
                reinterpret_cast<const uint8_t *>(base), d_msgs, (uint32_t)n, hasher, flags, pers[0], pers[1], E.bh_table.as<EdTabMem>(),
                E.sqrt_tables.as<SqrtTables>(), d_prep, reinterpret_cast<JacobianMem<HFq> *>(base + out_off),
                reinterpret_cast<uint32_t *>(base + att_off));
        };
        if (pack >= 8) launch_packed(k_hash_to_g1_packed<8>, 8);
        else launch_packed(k_hash_to_g1_packed<4>, 4);
    } else if (occ >= 16) launch(k_hash_to_g1<16>);
    else if (occ >= 12) launch(k_hash_to_g1<12>);
    else launch(k_hash_to_g1<8>);
    LAUNCH_CHECK();
    if (crh_only) {
        CUDA_TRY(cudaMemcpyAsync(out, base + crh_off, n * 48, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        return B200_OK;
    }
    std::vector<uint32_t> att(n);
    CUDA_TRY(cudaMemcpyAsync(out, base + out_off, n * 144, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(att.data(), base + att_off, n * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (size_t i = 0; i < n; i++) {
        if (att[i] == HASH_FAILED) return fail(B200_ERR_ARG, "input %zu: no curve point in 255 attempts (BLSError::HashToCurveError)", i);
        if (out_attempts) out_attempts[i] = att[i];
    }
    return B200_OK;
}

}  // namespace b200
