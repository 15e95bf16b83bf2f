"""CPU oracle (TEST INFRASTRUCTURE, never on the product path) for the message-hashing step that
precedes the multi-pairing in the reference: hashers + try-and-increment hash-to-curve.

Restates, with Python integers:

* `DirectHasher` (crates/bls-crypto/src/hashers/direct.rs:8-80): Blake2s CRH and the Blake2Xs-style XOF
  (node offset = block index | digest length << 32).
* `CompositeHasher<bowe_hopwood::CRH<EdwardsParameters, Window{93, 560}>>`
  (crates/bls-crypto/src/hashers/composite.rs:15-95): Bowe-Hopwood Pedersen hash over ed-on-bw6-761 with
  generators drawn from ChaCha20 seeded by Blake2s("ULTRALIGHT PRNG SEED", personal "UL_prngs"); the CRH output
  is the x coordinate, 48 bytes little endian.  The CRH, its setup and the samplers live in un-vendored
  dependencies (ark-crypto-primitives @ fde39ab7, ark-ec / ark-ff @ 8d76d181, ark-ed-on-bw6-761 @ 6ed2450b,
  rand_chacha 0.3.1): their published algorithms are restated here and PINNED by the reference's own
  known-answer tests (composite.rs:104-189) and its 30 hash-to-curve vectors (hash_to_curve/mod.rs:413-513).
* `TryAndIncrement` (hash_to_curve/try_and_increment.rs:84-139, with and without the `compat` bit
  extraction) and `TryAndIncrementCIP22` (try_and_increment_cip22.rs:60-134) for BLS12-377 G1 and G2,
  `from_random_bytes` (hash_to_curve/mod.rs:146-156).
* `XorShiftRng` (rand_xorshift 0.2.0) and `generate_test_data` (hash_to_curve/mod.rs:215-234), which produce
  the inputs of the reference's vectors.
"""
from __future__ import annotations

import hashlib
import struct
from functools import lru_cache
from typing import List, Optional, Tuple

from . import oracle as O

def _modulus() -> int:
    return O.P


# ----------------------------------------------------------------------------- Blake2s hashers
def _node_offset(i: int, xof_digest_length: int) -> int:
    """direct.rs:8-18: the two little-endian bytes of the digest length sit at bits 32..47."""
    return i | ((xof_digest_length & 0xFFFF) << 32)


_B2S_IV = (0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19)
_B2S_SIGMA = (
    (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15), (14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3),
    (11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4), (7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8),
    (9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13), (2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9),
    (12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11), (13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10),
    (6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5), (10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0),
)


def blake2s(data: bytes, digest_size: int = 32, person: bytes = b"", fanout: int = 1, depth: int = 1,
            leaf_size: int = 0, node_offset: int = 0, node_depth: int = 0, inner_size: int = 0) -> bytes:
    """Unkeyed BLAKE2s (RFC 7693) with the full parameter block -- hashlib refuses fanout = depth = 0, which the
    reference's XOF uses (direct.rs:57-64)."""
    params = struct.pack("<BBBBIIHBB", digest_size, 0, fanout, depth, leaf_size, node_offset & 0xFFFFFFFF,
                         (node_offset >> 32) & 0xFFFF, node_depth, inner_size) + bytes(8) + person.ljust(8, b"\0")
    h = [iv ^ pw for iv, pw in zip(_B2S_IV, struct.unpack("<8I", params))]
    M = 0xFFFFFFFF

    def compress(block: bytes, t: int, last: bool):
        m = struct.unpack("<16I", block)
        v = h + list(_B2S_IV)
        v[12] ^= t & M
        v[13] ^= (t >> 32) & M
        if last:
            v[14] ^= M

        def g(a, b, c, d, x, y):
            v[a] = (v[a] + v[b] + x) & M; v[d] = _rotr(v[d] ^ v[a], 16)
            v[c] = (v[c] + v[d]) & M; v[b] = _rotr(v[b] ^ v[c], 12)
            v[a] = (v[a] + v[b] + y) & M; v[d] = _rotr(v[d] ^ v[a], 8)
            v[c] = (v[c] + v[d]) & M; v[b] = _rotr(v[b] ^ v[c], 7)

        for r in range(10):
            s = _B2S_SIGMA[r]
            g(0, 4, 8, 12, m[s[0]], m[s[1]]); g(1, 5, 9, 13, m[s[2]], m[s[3]])
            g(2, 6, 10, 14, m[s[4]], m[s[5]]); g(3, 7, 11, 15, m[s[6]], m[s[7]])
            g(0, 5, 10, 15, m[s[8]], m[s[9]]); g(1, 6, 11, 12, m[s[10]], m[s[11]])
            g(2, 7, 8, 13, m[s[12]], m[s[13]]); g(3, 4, 9, 14, m[s[14]], m[s[15]])
        for i in range(8):
            h[i] ^= v[i] ^ v[i + 8]

    n = len(data)
    off = 0
    while n - off > 64:
        compress(data[off:off + 64], off + 64, False)
        off += 64
    compress(data[off:].ljust(64, b"\0"), n, True)
    return struct.pack("<8I", *h)[:digest_size]


def _rotr(v, n):
    return ((v >> n) | (v << (32 - n))) & 0xFFFFFFFF


def direct_crh(domain: bytes, message: bytes, xof_digest_length: int) -> bytes:
    """direct.rs:23-39."""
    return blake2s(message, digest_size=32, person=domain, node_offset=_node_offset(0, xof_digest_length))


def xof(domain: bytes, hashed_message: bytes, xof_digest_length: int) -> bytes:
    """direct.rs:41-79."""
    if len(domain) > 8:
        raise ValueError("domain too large")
    num = (xof_digest_length + 31) // 32
    out = b""
    for i in range(num):
        hl = xof_digest_length % 32 if (i == num - 1 and xof_digest_length % 32) else 32
        out += blake2s(hashed_message, digest_size=hl, person=domain, fanout=0, depth=0, leaf_size=32,
                       inner_size=32, node_offset=_node_offset(i, xof_digest_length))
    return out


# ----------------------------------------------------------------------------- RNGs
class XorShiftRng:
    """rand_xorshift 0.2.0 (the generator behind every seeded input in the reference's tests)."""

    def __init__(self, seed: bytes):
        self.x, self.y, self.z, self.w = struct.unpack("<4I", seed)

    def next_u32(self) -> int:
        t = (self.x ^ (self.x << 11)) & 0xFFFFFFFF
        self.x, self.y, self.z = self.y, self.z, self.w
        self.w = (self.w ^ (self.w >> 19) ^ (t ^ (t >> 8))) & 0xFFFFFFFF
        return self.w

    def gen_u8(self) -> int:
        return self.next_u32() & 0xFF


REFERENCE_SEED = bytes([0x5d, 0xbe, 0x62, 0x59, 0x8d, 0x31, 0x3d, 0x76, 0x32, 0x37, 0xdb, 0x17, 0xe5, 0xbc, 0x06, 0x54])


def generate_test_data(rng: XorShiftRng) -> Tuple[bytes, bytes, bytes]:
    """hash_to_curve/mod.rs:215-234 -> (domain, msg, extra_data)."""
    msg = bytes(rng.gen_u8() for _ in range(rng.gen_u8()))
    domain = bytes(rng.gen_u8() for _ in range(8))
    extra = bytes(rng.gen_u8() for _ in range(rng.gen_u8()))
    return domain, msg, extra


def _rotl(v, n):
    return ((v << n) | (v >> (32 - n))) & 0xFFFFFFFF


class ChaCha20Rng:
    """rand_chacha ChaChaRng (20 rounds, 64-bit block counter in words 12-13, stream 0), read as a
    continuous stream of little-endian u32 words; next_u64 = low word then high word."""

    def __init__(self, seed: bytes):
        self.key = struct.unpack("<8I", seed)
        self.counter = 0
        self.buf: List[int] = []

    def _block(self):
        s = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574, *self.key,
             self.counter & 0xFFFFFFFF, (self.counter >> 32) & 0xFFFFFFFF, 0, 0]
        w = list(s)

        def qr(a, b, c, d):
            w[a] = (w[a] + w[b]) & 0xFFFFFFFF; w[d] = _rotl(w[d] ^ w[a], 16)
            w[c] = (w[c] + w[d]) & 0xFFFFFFFF; w[b] = _rotl(w[b] ^ w[c], 12)
            w[a] = (w[a] + w[b]) & 0xFFFFFFFF; w[d] = _rotl(w[d] ^ w[a], 8)
            w[c] = (w[c] + w[d]) & 0xFFFFFFFF; w[b] = _rotl(w[b] ^ w[c], 7)

        for _ in range(10):
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
        self.counter += 1
        self.buf = [(w[i] + s[i]) & 0xFFFFFFFF for i in range(16)]

    def next_u32(self) -> int:
        if not self.buf:
            self._block()
        return self.buf.pop(0)

    def next_u64(self) -> int:
        lo = self.next_u32()
        return lo | (self.next_u32() << 32)

    def gen_bool(self) -> bool:
        return (self.next_u32() >> 31) == 1


# ----------------------------------------------------------------------------- ed-on-bw6-761
ED_A = -1
ED_D = 79743
ED_COFACTOR = 8
BH_WINDOW_SIZE = 93          # composite.rs:23
BH_NUM_WINDOWS = 560         # composite.rs:24
BH_CHUNK = 3


def _ed_add(p1, p2):
    """Unified addition on a x^2 + y^2 = 1 + d x^2 y^2 in extended coordinates (X, Y, Z, T = XY/Z)."""
    m = _modulus()
    x1, y1, z1, t1 = p1
    x2, y2, z2, t2 = p2
    a = x1 * x2 % m
    b = y1 * y2 % m
    c = ED_D * t1 % m * t2 % m
    d = z1 * z2 % m
    e = ((x1 + y1) * (x2 + y2) - a - b) % m
    f = (d - c) % m
    g = (d + c) % m
    h = (b - ED_A * a) % m
    return (e * f % m, g * h % m, f * g % m, e * h % m)


def _ed_neg(p):
    m = _modulus()
    return ((-p[0]) % m, p[1], p[2], (-p[3]) % m)


ED_ZERO = (0, 1, 1, 0)


def _ed_affine(p) -> Tuple[int, int]:
    m = _modulus()
    zi = O.inv(p[2], m)
    return p[0] * zi % m, p[1] * zi % m


def _ed_point_from_x(x: int, greatest: bool):
    """ark-ec twisted_edwards_extended GroupAffine::get_point_from_x."""
    m = _modulus()
    x2 = x * x % m
    num = (ED_A * x2 - 1) % m
    den = (ED_D * x2 - 1) % m
    if den == 0:
        return None
    y = O.sqrt_mod(num * O.inv(den, m) % m, m)
    if y is None:
        return None
    negy = (-y) % m
    y = y if ((y < negy) ^ greatest) else negy
    return (x, y, 1, x * y % m)


def _fq_rand(rng: ChaCha20Rng) -> int:
    """ark-ff `Standard` sampler for Fp384: six random limbs taken AS the Montgomery representation, top 7 bits
    shaved, rejected unless below the modulus."""
    m = _modulus()
    rinv = O.inv(1 << 384, m)
    while True:
        limbs = [rng.next_u64() for _ in range(6)]
        limbs[5] &= (1 << 64) - 1 >> 7
        v = sum(l << (64 * i) for i, l in enumerate(limbs))
        if v < m:
            return v * rinv % m


def _ed_rand(rng: ChaCha20Rng):
    """ark-ec `Standard` sampler for a twisted-Edwards GroupProjective."""
    while True:
        x = _fq_rand(rng)
        greatest = rng.gen_bool()
        p = _ed_point_from_x(x, greatest)
        if p is not None:
            for _ in range(3):          # scale_by_cofactor, cofactor 8
                p = _ed_add(p, p)
            return p


@lru_cache(maxsize=None)
def bh_base_generators() -> Tuple[Tuple[int, int], ...]:
    """The 560 per-window base points of composite.rs:53-72 (setup_crh), affine."""
    seed = hashlib.blake2s(b"ULTRALIGHT PRNG SEED", digest_size=32, person=b"UL_prngs").digest()
    rng = ChaCha20Rng(seed)
    return tuple(_ed_affine(_ed_rand(rng)) for _ in range(BH_NUM_WINDOWS))


def bh_crh_point(message: bytes):
    """bowe_hopwood::CRH::evaluate: 3-bit chunks, (1 + b0 + 2 b1) * (-1)^b2 times the chunk's generator,
    generator j of a window = 16^j times the window base."""
    if len(message) * 8 > BH_WINDOW_SIZE * BH_NUM_WINDOWS * BH_CHUNK:
        raise ValueError("message too long for the CRH")
    bits = [(byte >> i) & 1 for byte in message for i in range(8)]
    while len(bits) % BH_CHUNK:
        bits.append(0)
    gens = bh_base_generators()
    acc = ED_ZERO
    per_window = BH_WINDOW_SIZE * BH_CHUNK
    for w in range(0, len(bits), per_window):
        gx, gy = gens[w // per_window]
        g = (gx, gy, 1, gx * gy % _modulus())
        seg = bits[w:w + per_window]
        for c in range(0, len(seg), BH_CHUNK):
            b0, b1, b2 = seg[c:c + 3]
            enc = g
            g2 = _ed_add(g, g)
            if b0:
                enc = _ed_add(enc, g)
            if b1:
                enc = _ed_add(enc, g2)
            if b2:
                enc = _ed_neg(enc)
            acc = _ed_add(acc, enc)
            g4 = _ed_add(g2, g2)
            g8 = _ed_add(g4, g4)
            g = _ed_add(g8, g8)
    return acc


def composite_crh(message: bytes) -> bytes:
    """composite.rs:78-84: x coordinate of the CRH point, 48 bytes LE."""
    x, _ = _ed_affine(bh_crh_point(message))
    return x.to_bytes(48, "little")


def composite_hash(domain: bytes, message: bytes, out_len: int) -> bytes:
    """hashers/mod.rs:33-41 with the composite hasher."""
    return xof(domain, composite_crh(message), out_len)


def direct_hash(domain: bytes, message: bytes, out_len: int) -> bytes:
    return xof(domain, direct_crh(domain, message, out_len), out_len)


# ----------------------------------------------------------------------------- try and increment
def hash_length(n: int) -> int:
    """hash_to_curve/mod.rs:19-23."""
    return -(-(n * 8) // 256) * 256 // 8


def from_random_bytes(curve: O.Curve, bs: bytes):
    """hash_to_curve/mod.rs:146-156 over ark-ff's from_random_bytes_with_flags::<YSignFlags>:
    returns None (reject), "zero" or an affine point."""
    nb = curve.coord_bytes * curve.ext_degree
    bs = bytearray(bs[:nb])
    flags = bs[nb - 1] & 0xC0
    m = curve.modulus
    coords = []
    for k in range(curve.ext_degree):
        part = bytearray(bs[k * curve.coord_bytes:(k + 1) * curve.coord_bytes])
        part[-1] &= 0x01          # each coordinate: every bit at or above MODULUS_BITS = 377 is masked off
        v = int.from_bytes(part, "little")
        if v >= m:
            return None
        coords.append(v)
    x = coords[0] if curve.ext_degree == 1 else tuple(coords)
    positive, infinity = bool(flags & 0x80), bool(flags & 0x40)
    if infinity and (x == 0 or x == (0, 0)):
        return "zero"
    rhs = curve.add(curve.mul(curve.mul(x, x), x), curve.b)
    y = O._sqrt_coord(curve, rhs)
    if y is None:
        return None
    if O._y_is_larger(curve, y) != positive:
        y = curve.neg(y)
    return (x, y)


def _cofactor(curve: O.Curve) -> int:
    return COFACTORS[curve.name]


def try_and_increment(curve: O.Curve, hasher, domain: bytes, message: bytes, extra: bytes, compat: bool = True,
                      cip22: bool = False):
    """try_and_increment.rs:84-139 / try_and_increment_cip22.rs:60-134 -> (affine point, attempt).
    `hasher` = (crh, xof) pair of callables crh(domain, msg, n), xof(domain, msg, n)."""
    crh, xof_ = hasher
    nb = curve.coord_bytes * curve.ext_degree
    hb = hash_length(nb)
    inner = crh(domain, message, hb) if cip22 else None
    for c in range(255):
        if cip22:
            cand = xof_(domain, bytes([c]) + extra + inner, hb)
        else:
            cand = xof_(domain, crh(domain, bytes([c]) + extra + message, hb), hb)
        cand = bytearray(cand[:nb])
        if compat:
            if cand[nb - 1] & 2:
                cand[nb - 1] |= 0x80
            else:
                cand[nb - 1] &= 0x7F
        p = from_random_bytes(curve, bytes(cand))
        if p is None or p == "zero":
            continue
        scaled = curve.pmul(p, _cofactor(curve))
        if scaled is None:
            continue
        return scaled, c
    raise ValueError("hash to curve failed")


COMPOSITE = (lambda d, m, n: composite_crh(m), xof)
DIRECT = (direct_crh, xof)

_X = 0x8508C00000000001
# G1: (x - 1)^2 / 3; G2: (x^8 - 4x^7 + 5x^6 - 4x^4 + 6x^3 - 4x^2 - 4x + 13) / 9  (order of the twist over Fq2 / r)
COFACTORS = {
    "bls12_377_g1": (_X - 1) ** 2 // 3,
    "bls12_377_g2": (_X ** 8 - 4 * _X ** 7 + 5 * _X ** 6 - 4 * _X ** 4 + 6 * _X ** 3 - 4 * _X ** 2 - 4 * _X + 13) // 9,
}
assert COFACTORS["bls12_377_g1"] == 0x170B5D44300000000000000000000000
