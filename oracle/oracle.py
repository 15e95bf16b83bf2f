"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the MSM / multi-pairing hot path.

This file is a plain-Python big-integer restatement of the arithmetic that
celo-bls-snark-rs reaches through arkworks at these call sites:

  * crates/bls-crypto/src/bls/signature.rs:70-89   Signature::batch  -> G1 MSM
  * crates/bls-crypto/src/bls/public.rs:47-65      PublicKey::batch  -> G2 MSM
  * crates/bls-crypto/src/bls/signature.rs:125-155 batch_verify_hashes -> (N+1)-pair product_of_pairings
  * crates/bls-crypto/src/bls/public.rs:94-120     verify_sig -> 2-pair product_of_pairings
  * crates/bls-crypto/src/bls/batch.rs:23-84       Batch::verify exponent sizing / order of ops
  * crates/epoch-snark/src/api/prover.rs:78,112    create_proof_no_zk -> 4 MSMs (BW6-761 / BLS12-377)

The algorithms themselves live in third-party crates that are NOT vendored in
/root/reference (Cargo.lock pins: ark-ec / ark-ff / ark-serialize 0.1.0 @
arkworks-rs/algebra#8d76d181, ark-bls12-377 / ark-bw6-761 0.1.0 @
arkworks-rs/curves#6ed2450b, ark-groth16 0.1.0 @ arkworks-rs/groth16#d8acb2b2).
What is restated here is their *published* algorithm (Pippenger bucket MSM,
short-Weierstrass group law, BLS12 optimal-ate pairing with the 2016/130 final
exponentiation), with textbook affine formulas wherever the result is a
canonical group element.

PARITY PINNING.  No reference test pins an MSM output or a BLS12-377 GT value
byte-for-byte, so for those the oracle is "parity
unpinned" at byte level and parity is argued through canonical forms: an MSM
result is a unique group element and the verify result a unique boolean.  The
building blocks ARE pinned against the reference's own golden vectors (see
tests/test_oracle_golden.py): compressed G1/G2 encodings from
crates/bls-crypto/src/hash_to_curve/mod.rs:415-426,438-449,474-485,497-508,
the G2 generator from crates/epoch-snark/src/epoch_block.rs:243-246 and the
BW6-761 verifying key / proof / BLS12-377 public keys from
crates/bls-snark-sys/src/snark/mod.rs:52-64.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product (celo_bls_snark_rs_b200)
never does.
"""
from __future__ import annotations

import hashlib
from typing import Iterable, List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------
# BLS12-377 parameters ("[verified]" constants)
# --------------------------------------------------------------------------
X = 0x8508C00000000001                      # curve seed, positive
R = X**4 - X**2 + 1                         # scalar field modulus (253 bit)
P = ((X - 1) ** 2 * R) // 3 + X             # base field modulus (377 bit)
assert R == 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001
assert P == 0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001
G1_COFACTOR = (X - 1) ** 2 // 3
FQ_BYTES = 48
FR_BYTES = 32
NONRESIDUE = -5 % P                         # Fq2 = Fq[u]/(u^2 + 5)

G1_GEN = (
    0x008848DEFE740A67C8FC6225BF87FF5485951E2CAA9D41BB188282C8BD37CB5CD5481512FFCD394EEAB9B16EB21BE9EF,
    0x01914A69C5102EFF1F674F5D30AFEEC4BD7FB348CA3E52D96D182AD44FB82305C2FE3D3634A9591AFD82DE55559C8EA6,
)

# --------------------------------------------------------------------------
# BW6-761 parameters
# --------------------------------------------------------------------------
Q761 = 0x0122E824FB83CE0AD187C94004FAFF3EB926186A81D14688528275EF8087BE41707BA638E584E91903CEBAFF25B423048689C8ED12F9FD9071DCD3DC73EBFF2E98A116C25667A8F8160CF8AEEAF0A437E6913E6870000082F49D00000000008B
R761 = P                                    # BW6-761 scalar field == BLS12-377 base field
FQ761_BYTES = 96


# --------------------------------------------------------------------------
# prime-field helpers
# --------------------------------------------------------------------------
def inv(a: int, m: int) -> int:
    return pow(a, -1, m)


def sqrt_mod(a: int, m: int) -> Optional[int]:
    """Square root mod prime m (Tonelli-Shanks; m = 3 mod 4 shortcut)."""
    a %= m
    if a == 0:
        return 0
    if pow(a, (m - 1) // 2, m) != 1:
        return None
    if m % 4 == 3:
        return pow(a, (m + 1) // 4, m)
    q, s = m - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 2
    while pow(z, (m - 1) // 2, m) != m - 1:
        z += 1
    mm, c, t, r = s, pow(z, q, m), pow(a, q, m), pow(a, (q + 1) // 2, m)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % m
            i += 1
        b = pow(c, 1 << (mm - i - 1), m)
        mm, c = i, b * b % m
        t, r = t * c % m, r * b % m
    return r


# --------------------------------------------------------------------------
# Fq2 over BLS12-377 Fq: (c0, c1) = c0 + c1*u, u^2 = -5
# --------------------------------------------------------------------------
Fq2 = Tuple[int, int]
FQ2_ZERO: Fq2 = (0, 0)
FQ2_ONE: Fq2 = (1, 0)


def f2_add(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_neg(a: Fq2) -> Fq2:
    return (-a[0] % P, -a[1] % P)


def f2_mul(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] * b[0] + NONRESIDUE * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_sqr(a: Fq2) -> Fq2:
    return f2_mul(a, a)


def f2_scale(a: Fq2, k: int) -> Fq2:
    return (a[0] * k % P, a[1] * k % P)


def f2_conj(a: Fq2) -> Fq2:
    return (a[0], -a[1] % P)


def f2_inv(a: Fq2) -> Fq2:
    n = inv((a[0] * a[0] - NONRESIDUE * a[1] * a[1]) % P, P)
    return (a[0] * n % P, -a[1] * n % P)


def f2_pow(a: Fq2, e: int) -> Fq2:
    r = FQ2_ONE
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_sqr(a)
        e >>= 1
    return r


def f2_sqrt(a: Fq2) -> Optional[Fq2]:
    """Square root in Fq2 via the norm trick (any root; caller picks the sign)."""
    if a == FQ2_ZERO:
        return FQ2_ZERO
    if a[1] == 0:
        s = sqrt_mod(a[0], P)
        if s is not None:
            return (s, 0)
        # a0 is a non-residue: sqrt = t*u with t^2 * (-5) = a0
        t = sqrt_mod(a[0] * inv(NONRESIDUE, P) % P, P)
        return None if t is None else (0, t)
    norm = (a[0] * a[0] - NONRESIDUE * a[1] * a[1]) % P
    alpha = sqrt_mod(norm, P)
    if alpha is None:
        return None
    two_inv = inv(2, P)
    delta = (a[0] + alpha) * two_inv % P
    x0 = sqrt_mod(delta, P)
    if x0 is None:
        delta = (a[0] - alpha) * two_inv % P
        x0 = sqrt_mod(delta, P)
        if x0 is None:
            return None
    x1 = a[1] * inv(2 * x0 % P, P) % P
    r = (x0, x1)
    return r if f2_sqr(r) == (a[0] % P, a[1] % P) else None


# --------------------------------------------------------------------------
# Fq6 = Fq2[v]/(v^3 - u), Fq12 = Fq6[w]/(w^2 - v)
# (tower as used by ark-bls12-377)
# --------------------------------------------------------------------------
XI: Fq2 = (0, 1)                            # the Fq6 non-residue u
Fq6 = Tuple[Fq2, Fq2, Fq2]
Fq12 = Tuple[Fq6, Fq6]
FQ6_ZERO: Fq6 = (FQ2_ZERO, FQ2_ZERO, FQ2_ZERO)
FQ6_ONE: Fq6 = (FQ2_ONE, FQ2_ZERO, FQ2_ZERO)
FQ12_ONE: Fq12 = (FQ6_ONE, FQ6_ZERO)


def f2_mul_xi(a: Fq2) -> Fq2:
    # (a0 + a1 u) * u = -5 a1 + a0 u
    return (NONRESIDUE * a[1] % P, a[0])


def f6_add(a: Fq6, b: Fq6) -> Fq6:
    return (f2_add(a[0], b[0]), f2_add(a[1], b[1]), f2_add(a[2], b[2]))


def f6_sub(a: Fq6, b: Fq6) -> Fq6:
    return (f2_sub(a[0], b[0]), f2_sub(a[1], b[1]), f2_sub(a[2], b[2]))


def f6_neg(a: Fq6) -> Fq6:
    return (f2_neg(a[0]), f2_neg(a[1]), f2_neg(a[2]))


def f6_mul(a: Fq6, b: Fq6) -> Fq6:
    a0, a1, a2 = a
    b0, b1, b2 = b
    c0 = f2_add(f2_mul(a0, b0), f2_mul_xi(f2_add(f2_mul(a1, b2), f2_mul(a2, b1))))
    c1 = f2_add(f2_add(f2_mul(a0, b1), f2_mul(a1, b0)), f2_mul_xi(f2_mul(a2, b2)))
    c2 = f2_add(f2_add(f2_mul(a0, b2), f2_mul(a1, b1)), f2_mul(a2, b0))
    return (c0, c1, c2)


def f6_mul_v(a: Fq6) -> Fq6:
    # (a0 + a1 v + a2 v^2) * v = xi*a2 + a0 v + a1 v^2
    return (f2_mul_xi(a[2]), a[0], a[1])


def f6_inv(a: Fq6) -> Fq6:
    a0, a1, a2 = a
    t0 = f2_sub(f2_sqr(a0), f2_mul_xi(f2_mul(a1, a2)))
    t1 = f2_sub(f2_mul_xi(f2_sqr(a2)), f2_mul(a0, a1))
    t2 = f2_sub(f2_sqr(a1), f2_mul(a0, a2))
    d = f2_add(f2_mul(a0, t0), f2_mul_xi(f2_add(f2_mul(a2, t1), f2_mul(a1, t2))))
    di = f2_inv(d)
    return (f2_mul(t0, di), f2_mul(t1, di), f2_mul(t2, di))


def f12_mul(a: Fq12, b: Fq12) -> Fq12:
    a0, a1 = a
    b0, b1 = b
    t0 = f6_mul(a0, b0)
    t1 = f6_mul(a1, b1)
    c0 = f6_add(t0, f6_mul_v(t1))
    c1 = f6_sub(f6_sub(f6_mul(f6_add(a0, a1), f6_add(b0, b1)), t0), t1)
    return (c0, c1)


def f12_sqr(a: Fq12) -> Fq12:
    return f12_mul(a, a)


def f12_conj(a: Fq12) -> Fq12:
    return (a[0], f6_neg(a[1]))


def f12_inv(a: Fq12) -> Fq12:
    a0, a1 = a
    d = f6_sub(f6_mul(a0, a0), f6_mul_v(f6_mul(a1, a1)))
    di = f6_inv(d)
    return (f6_mul(a0, di), f6_neg(f6_mul(a1, di)))


def f12_pow(a: Fq12, e: int) -> Fq12:
    r = FQ12_ONE
    while e:
        if e & 1:
            r = f12_mul(r, a)
        a = f12_sqr(a)
        e >>= 1
    return r


# Frobenius: computed, never hard-coded from memory.
# Basis of Fq12 over Fq2: w^k, k = 0..5, with w^2 = v, w^6 = xi.
# frob^j(c * w^k) = frob^j(c) * xi^(k (p^j - 1)/6) * w^k.
def _frob_coeffs(j: int) -> List[Fq2]:
    e = (P**j - 1) // 6
    g = f2_pow(XI, e)
    out = [FQ2_ONE]
    for _ in range(5):
        out.append(f2_mul(out[-1], g))
    return out


_FROB = {j: _frob_coeffs(j) for j in (1, 2, 3)}


def _f2_frob(a: Fq2, j: int) -> Fq2:
    return f2_conj(a) if j % 2 else a


def f12_frob(a: Fq12, j: int) -> Fq12:
    co = _FROB[j]
    # index in w-power basis: c0 = (w^0, w^2, w^4), c1 = (w^1, w^3, w^5)
    (a00, a01, a02), (a10, a11, a12) = a
    return (
        (f2_mul(_f2_frob(a00, j), co[0]), f2_mul(_f2_frob(a01, j), co[2]), f2_mul(_f2_frob(a02, j), co[4])),
        (f2_mul(_f2_frob(a10, j), co[1]), f2_mul(_f2_frob(a11, j), co[3]), f2_mul(_f2_frob(a12, j), co[5])),
    )


# --------------------------------------------------------------------------
# short-Weierstrass curves y^2 = x^3 + b (a = 0 for all four groups)
# A Curve bundles the coordinate-field ops so G1/G2/BW6 share one group law.
# Points are None (infinity) or (x, y).
# --------------------------------------------------------------------------
class Curve:
    def __init__(self, name, zero, add, sub, mul, inv_, neg, b, order_bits, scalar_mod, coord_bytes, ext_degree, modulus):
        self.name, self.zero = name, zero
        self.add, self.sub, self.mul, self.inv, self.neg = add, sub, mul, inv_, neg
        self.b = b
        self.scalar_bits = order_bits
        self.scalar_mod = scalar_mod
        self.coord_bytes = coord_bytes      # bytes of ONE base-prime-field element
        self.ext_degree = ext_degree
        self.modulus = modulus

    def on_curve(self, pt) -> bool:
        if pt is None:
            return True
        x, y = pt
        return self.mul(y, y) == self.add(self.mul(self.mul(x, x), x), self.b)

    def pneg(self, pt):
        return None if pt is None else (pt[0], self.neg(pt[1]))

    def padd(self, p1, p2):
        """Textbook affine addition (chord / tangent), the arbiter formula."""
        if p1 is None:
            return p2
        if p2 is None:
            return p1
        x1, y1 = p1
        x2, y2 = p2
        if x1 == x2:
            if y1 != y2 or y1 == self.zero:
                return None
            num = self.mul(self.mul(x1, x1), self._three)
            den = self.add(y1, y1)
        else:
            num = self.sub(y2, y1)
            den = self.sub(x2, x1)
        lam = self.mul(num, self.inv(den))
        x3 = self.sub(self.sub(self.mul(lam, lam), x1), x2)
        y3 = self.sub(self.mul(lam, self.sub(x1, x3)), y1)
        return (x3, y3)

    def pmul(self, pt, k: int):
        if k < 0:
            return self.pmul(self.pneg(pt), -k)
        acc = None
        while k:
            if k & 1:
                acc = self.padd(acc, pt)
            pt = self.padd(pt, pt)
            k >>= 1
        return acc

    def msm_naive(self, bases: Sequence, scalars: Sequence[int]):
        """sum_i scalars[i] * bases[i], one double-and-add per pair (the arbiter)."""
        acc = None
        for b, s in zip(bases, scalars):
            acc = self.padd(acc, self.pmul(b, s))
        return acc


def _fp_curve(name, m, b, scalar_mod, coord_bytes):
    c = Curve(
        name, 0,
        lambda a, b_: (a + b_) % m, lambda a, b_: (a - b_) % m, lambda a, b_: a * b_ % m,
        lambda a: inv(a, m), lambda a: -a % m,
        b % m, scalar_mod.bit_length(), scalar_mod, coord_bytes, 1, m,
    )
    c._three = 3
    return c


G1 = _fp_curve("bls12_377_g1", P, 1, R, FQ_BYTES)
TWIST_B: Fq2 = (0, (-inv(5, P)) % P)        # b / u  (D-twist), = (0, 1551986...874906)
assert TWIST_B[1] == 155198655607781456406391640216936120121836107652948796323930557600032281009004493664981332883744016074664192874906
G2 = Curve("bls12_377_g2", FQ2_ZERO, f2_add, f2_sub, f2_mul, f2_inv, f2_neg, TWIST_B, R.bit_length(), R, FQ_BYTES, 2, P)
G2._three = (3, 0)
BW6_G1 = _fp_curve("bw6_761_g1", Q761, -1, R761, FQ761_BYTES)
BW6_G2 = _fp_curve("bw6_761_g2", Q761, 4, R761, FQ761_BYTES)
CURVES = {c.name: c for c in (G1, G2, BW6_G1, BW6_G2)}
assert G1.on_curve(G1_GEN)


# --------------------------------------------------------------------------
# arkworks canonical serialization (ark-serialize 0.1.0; flags mirrored by the
# repo's own YSignFlags, crates/bls-crypto/src/hash_to_curve/mod.rs:118-144)
#   x little-endian (Fq2: c0 || c1); last byte: bit7 = y is the larger root,
#   bit6 = infinity.
# --------------------------------------------------------------------------
def _y_is_larger(curve: Curve, y) -> bool:
    m = curve.modulus
    if curve.ext_degree == 1:
        return y > (-y % m)
    ny = f2_neg(y)
    # Fq2 ordering: compare c1 first, then c0 (crates/epoch-snark/src/encoding.rs:31-33)
    return (y[1], y[0]) > (ny[1], ny[0])


def _coord_to_bytes(curve: Curve, c) -> bytes:
    if curve.ext_degree == 1:
        return int(c).to_bytes(curve.coord_bytes, "little")
    return b"".join(int(ci).to_bytes(curve.coord_bytes, "little") for ci in c)


def _coord_from_bytes(curve: Curve, bs: bytes):
    n = curve.coord_bytes
    if curve.ext_degree == 1:
        return int.from_bytes(bs[:n], "little")
    return tuple(int.from_bytes(bs[i * n:(i + 1) * n], "little") for i in range(curve.ext_degree))


def serialize_compressed(curve: Curve, pt) -> bytes:
    size = curve.coord_bytes * curve.ext_degree
    if pt is None:
        out = bytearray(size)
        out[-1] |= 1 << 6
        return bytes(out)
    out = bytearray(_coord_to_bytes(curve, pt[0]))
    if _y_is_larger(curve, pt[1]):
        out[-1] |= 1 << 7
    return bytes(out)


def serialize_uncompressed(curve: Curve, pt) -> bytes:
    """x || y little-endian (crates/bls-snark-sys/src/serialization.rs:167-189);
    infinity flag (bit6) lives in the last byte of y."""
    size = curve.coord_bytes * curve.ext_degree
    if pt is None:
        out = bytearray(2 * size)
        out[-1] |= 1 << 6
        return bytes(out)
    return _coord_to_bytes(curve, pt[0]) + _coord_to_bytes(curve, pt[1])


def _sqrt_coord(curve: Curve, a):
    if curve.ext_degree == 1:
        return sqrt_mod(a, curve.modulus)
    return f2_sqrt(a)


def deserialize_compressed(curve: Curve, bs: bytes):
    size = curve.coord_bytes * curve.ext_degree
    assert len(bs) == size, (len(bs), size)
    raw = bytearray(bs)
    flags = raw[-1]
    raw[-1] &= 0x3F
    if flags & 0x40:
        return None
    x = _coord_from_bytes(curve, bytes(raw))
    rhs = curve.add(curve.mul(curve.mul(x, x), x), curve.b)
    y = _sqrt_coord(curve, rhs)
    if y is None:
        raise ValueError("x is not on the curve")
    if _y_is_larger(curve, y) != bool(flags & 0x80):
        y = curve.neg(y)
    return (x, y)


# BLS12-377 G2 generator: decoded from the reference's own golden encoding
# (crates/epoch-snark/src/epoch_block.rs:243-246).
_G2_GEN_X: Fq2 = (
    0x018480BE71C785FEC89630A2A3841D01C565F071203E50317EA501F557DB6B9B71889F52BB53540274E3E48F7C005196,
    0x00EA6040E700403170DC5A51B1B140D5532777EE6651CECBE7223ECE0799C9DE5CF89984BFF76FE6B26BFEFA6EA16AFE,
)


def _g2_gen():
    y = f2_sqrt(f2_add(f2_mul(f2_sqr(_G2_GEN_X), _G2_GEN_X), TWIST_B))
    assert y is not None
    if not _y_is_larger(G2, y):             # golden vector has the "over half" bit = 1
        y = f2_neg(y)
    return (_G2_GEN_X, y)


G2_GEN = _g2_gen()


# --------------------------------------------------------------------------
# VariableBaseMSM::multi_scalar_mul, restated:
# same window rule, zero/unit-scalar handling, running-sum reduction and
# high->low combine as ark-ec 0.1.0.  Group ops are the affine arbiter
# formulas, so this is slow and only used on small n.
# --------------------------------------------------------------------------
def ln_without_floats(a: int) -> int:
    # ark-std log2(a) = ceil(log2 a); ln ~ log2 * 69 / 100
    return ((a - 1).bit_length() if a > 1 else 0) * 69 // 100


def msm_window_bits(n: int) -> int:
    return 3 if n < 32 else ln_without_floats(n) + 2


def msm_pippenger(curve: Curve, bases: Sequence, scalars: Sequence[int]):
    n = min(len(bases), len(scalars))
    c = msm_window_bits(n)
    pairs = [(s, b) for s, b in zip(scalars[:n], bases[:n]) if s != 0]
    window_sums = []
    for w_start in range(0, curve.scalar_bits, c):
        res = None
        buckets = [None] * ((1 << c) - 1)
        for s, b in pairs:
            if s == 1:
                if w_start == 0:
                    res = curve.padd(res, b)
            else:
                d = (s >> w_start) & ((1 << c) - 1)
                if d:
                    buckets[d - 1] = curve.padd(buckets[d - 1], b)
        running = None
        for bk in reversed(buckets):
            running = curve.padd(running, bk)
            res = curve.padd(res, running)
        window_sums.append(res)
    total = None
    for ws in reversed(window_sums[1:]):
        total = curve.padd(total, ws)
        for _ in range(c):
            total = curve.padd(total, total)
    return curve.padd(window_sums[0], total)


# --------------------------------------------------------------------------
# BLS12-377 optimal-ate pairing, restated from ark-ec 0.1.0 models::bls12
# (SURVEY.md appendix A.3).  G2Prepared line coefficients in homogeneous
# projective coordinates, D-twist, mul_by_034, 2016/130 final exponentiation.
# --------------------------------------------------------------------------
_TWO_INV = inv(2, P)
_X_BITS_AFTER_MSB = [int(b) for b in bin(X)[3:]]     # 63 bits


def _g2_doubling_step(r):
    rx, ry, rz = r
    a = f2_scale(f2_mul(rx, ry), _TWO_INV)
    b = f2_sqr(ry)
    c = f2_sqr(rz)
    e = f2_mul(TWIST_B, f2_add(f2_add(c, c), c))
    f = f2_add(f2_add(e, e), e)
    g = f2_scale(f2_add(b, f), _TWO_INV)
    h = f2_sub(f2_sqr(f2_add(ry, rz)), f2_add(b, c))
    i = f2_sub(e, b)
    j = f2_sqr(rx)
    e2 = f2_sqr(e)
    nrx = f2_mul(a, f2_sub(b, f))
    nry = f2_sub(f2_sqr(g), f2_add(f2_add(e2, e2), e2))
    nrz = f2_mul(b, h)
    return (nrx, nry, nrz), (f2_neg(h), f2_add(f2_add(j, j), j), i)


def _g2_addition_step(r, q):
    rx, ry, rz = r
    qx, qy = q
    theta = f2_sub(ry, f2_mul(qy, rz))
    lam = f2_sub(rx, f2_mul(qx, rz))
    c = f2_sqr(theta)
    d = f2_sqr(lam)
    e = f2_mul(lam, d)
    f = f2_mul(rz, c)
    g = f2_mul(rx, d)
    h = f2_sub(f2_add(e, f), f2_add(g, g))
    nrx = f2_mul(lam, h)
    nry = f2_sub(f2_mul(theta, f2_sub(g, h)), f2_mul(e, ry))
    nrz = f2_mul(rz, e)
    j = f2_sub(f2_mul(theta, qx), f2_mul(lam, qy))
    return (nrx, nry, nrz), (lam, f2_neg(theta), j)


def g2_prepare(q) -> List[Tuple[Fq2, Fq2, Fq2]]:
    """G2Prepared::from(G2Affine): 63 doubling + 6 addition line triples."""
    if q is None:
        return []
    r = (q[0], q[1], FQ2_ONE)
    coeffs = []
    for bit in _X_BITS_AFTER_MSB:
        r, l = _g2_doubling_step(r)
        coeffs.append(l)
        if bit:
            r, l = _g2_addition_step(r, q)
            coeffs.append(l)
    return coeffs


def f12_mul_by_034(f: Fq12, c0: Fq2, d0: Fq2, d1: Fq2) -> Fq12:
    """f * (c0 + d0*w*... ) for the D-twist sparse element with non-zero
    Fq2 coefficients at positions 0 (c0.c0), 3 (c1.c0) and 4 (c1.c1)."""
    sparse: Fq12 = ((c0, FQ2_ZERO, FQ2_ZERO), (d0, d1, FQ2_ZERO))
    return f12_mul(f, sparse)


def _ell(f: Fq12, coeffs, p) -> Fq12:
    c0, c1, c2 = coeffs
    px, py = p
    # D-twist: c0 *= P.y ; c1 *= P.x ; f.mul_by_034(c0, c1, c2)
    return f12_mul_by_034(f, f2_scale(c0, py), f2_scale(c1, px), c2)


def miller_loop(pairs: Sequence[Tuple[Optional[Tuple[int, int]], object]]) -> Fq12:
    """Bls12::miller_loop over (G1Affine, G2Affine) pairs; pairs with an
    infinite member are skipped, as arkworks does."""
    prepared = [(p, g2_prepare(q)) for p, q in pairs if p is not None and q is not None]
    idx = [0] * len(prepared)
    f = FQ12_ONE
    for bit in _X_BITS_AFTER_MSB:
        f = f12_sqr(f)
        for k, (p, co) in enumerate(prepared):
            f = _ell(f, co[idx[k]], p)
            idx[k] += 1
        if bit:
            for k, (p, co) in enumerate(prepared):
                f = _ell(f, co[idx[k]], p)
                idx[k] += 1
    return f                                 # x > 0: no conjugation


def _exp_by_x(f: Fq12) -> Fq12:
    return f12_pow(f, X)


def final_exponentiation(f: Fq12) -> Fq12:
    """Bls12::final_exponentiation, eprint 2016/130 table 1 chain."""
    f1 = f12_conj(f)
    f2 = f12_inv(f)
    r = f12_mul(f1, f2)
    r = f12_mul(f12_frob(r, 2), r)           # easy part: (p^6 - 1)(p^2 + 1)
    y0 = f12_conj(f12_sqr(r))                # cyclotomic: conj == inverse
    y5 = _exp_by_x(r)
    y1 = f12_sqr(y5)
    y3 = f12_mul(y0, y5)
    y0 = _exp_by_x(y3)
    y2 = _exp_by_x(y0)
    y4 = f12_mul(_exp_by_x(y2), y1)
    y1 = _exp_by_x(y4)
    y3 = f12_conj(y3)
    y1 = f12_mul(f12_mul(y1, y3), r)
    y3 = f12_conj(r)
    y0 = f12_frob(f12_mul(y0, r), 3)
    y4 = f12_frob(f12_mul(y4, y3), 1)
    y5 = f12_frob(f12_mul(y5, y2), 2)
    return f12_mul(f12_mul(f12_mul(y5, y0), y4), y1)


def product_of_pairings(pairs) -> Fq12:
    """PairingEngine::product_of_pairings (signature.rs:149, public.rs:102)."""
    return final_exponentiation(miller_loop(pairs))


def pairing(p, q) -> Fq12:
    return product_of_pairings([(p, q)])


# --------------------------------------------------------------------------
# callers of the hot path, restated (the behavioural contract)
# --------------------------------------------------------------------------
def batch_verify_hashes(sig, pubkeys: Sequence, message_hashes: Sequence) -> bool:
    """Signature::batch_verify_hashes (signature.rs:125-155): first pair is
    (sigma, -g2), then (H_i, pk_i); true iff the product is one."""
    if len(pubkeys) != len(message_hashes):
        raise ValueError("UnevenNumKeysMessages")
    els = [(sig, G2.pneg(G2_GEN))]
    els += [(h, pk) for h, pk in zip(message_hashes, pubkeys)]
    return product_of_pairings(els) == FQ12_ONE


def verify_hash(pk, message_hash, sig) -> bool:
    """PublicKey::verify_sig after hashing (public.rs:94-120)."""
    return product_of_pairings([(sig, G2.pneg(G2_GEN)), (message_hash, pk)]) == FQ12_ONE


def batch_exponent_bytes(n: int, target_security: int = 128) -> int:
    """byte_count_from_target_batch_size (batch.rs:23-28): ark_std::log2 is ceil."""
    log2n = (n - 1).bit_length() if n > 1 else 0
    return min((target_security + log2n + 7) // 8, R.bit_length() // 8)


def strict_batch_verify_hash(entries: Sequence, message_hash, exponents: Sequence[int]) -> bool:
    """Batch::verify (batch.rs:44-84) with the exponents supplied by the caller
    (the reference draws them from thread_rng)."""
    pks = [pk for pk, _ in entries]
    sigs = [s for _, s in entries]
    bpk = msm_pippenger(G2, pks, exponents)
    bsig = msm_pippenger(G1, sigs, exponents)
    return verify_hash(bpk, message_hash, bsig)


# --------------------------------------------------------------------------
# deterministic synthetic inputs: SplitMix64 seeded with
# the first 8 bytes of the reference's test seed (hash_to_curve/mod.rs:290-293)
# --------------------------------------------------------------------------
SEED = 0x5DBE62598D313D76
_M64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int = SEED):
        self.s = seed & _M64

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    def below(self, m: int) -> int:
        """Uniform-ish integer < m: draw ceil(bits/64) words, mask to the bit
        length of m, reject if >= m (same rule as the C and CUDA generators)."""
        bits = m.bit_length()
        words = (bits + 63) // 64
        while True:
            v = 0
            for i in range(words):
                v |= self.next() << (64 * i)
            v &= (1 << bits) - 1
            if v < m:
                return v
