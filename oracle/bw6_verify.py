"""CPU restatement of the epoch-SNARK verifier path.  TEST
INFRASTRUCTURE ONLY (see oracle/oracle.py header).

Follows, on the reference side:
  * crates/bls-snark-sys/src/snark/mod.rs:23-45        C-ABI `verify`
  * crates/bls-snark-sys/src/snark/epoch_block.rs:129-196   EpochBlockFFI -> EpochBlock (96-byte compressed G2 keys)
  * crates/epoch-snark/src/api/verifier.rs:23-40       hash_first_last_epoch_block -> pack -> verify_proof
  * crates/epoch-snark/src/epoch_block.rs:106-236      CIP22 bit encodings, Blake2s("ULforout") edge hashes
  * crates/epoch-snark/src/encoding.rs:23-80           encode_public_key / encode_u16 / encode_u32
  * crates/epoch-snark/src/gadgets/mod.rs:75-83        pack (chunks of CAPACITY = 376 bits, big-endian)
and, upstream (ark-groth16 0.1.0 verify_proof):
  g_ic = gamma_abc[0] + sum_i input_i * gamma_abc[i + 1];
  e(A, B) * e(g_ic, -gamma) * e(C, -delta) == e(alpha, beta).

The pairing used here is NOT arkworks' optimal-ate BW6 loop: the check above is an identity between
pairing values and holds for every non-degenerate bilinear pairing on (G1, G2), so the oracle uses the
textbook reduced Tate pairing t(P, Q) = f_{r,P}(psi(Q))^((q^6 - 1) / r) over F_q^6 = F_q[w] / (w^6 + 4),
with psi the M-twist isomorphism E': y^2 = x^3 + 4 -> E: y^2 = x^3 - 1, (x, y) -> (x / w^2, y / w^3)
(w^6 = -4).  PINNED: the reference's own known-answer test
(crates/bls-snark-sys/src/snark/mod.rs:52-119: real VK, proof and epoch blocks, expected `true`) passes
through this code in tests/test_oracle_bw6_verify.py, and any corrupted byte makes it fail.
"""
from __future__ import annotations

import hashlib
from typing import List, Optional, Sequence

from . import oracle as O

Q = O.Q761
R = O.R761                      # group order = BLS12-377 base-field modulus (377 bits)

# ---- F_q^6 = F_q[w] / (w^6 + 4), elements are lists of 6 ints -------------------------------------------


def f6_one():
    return [1, 0, 0, 0, 0, 0]


def f6_mul(a, b):
    t = [0] * 11
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                t[i + j] += x * y
    return [(t[k] - 4 * (t[k + 6] if k < 5 else 0)) % Q for k in range(6)]


def f6_pow(a, e):
    r, base = f6_one(), a
    while e:
        if e & 1:
            r = f6_mul(r, base)
        base = f6_mul(base, base)
        e >>= 1
    return r


_NEG_QUARTER = (-pow(4, -1, Q)) % Q            # 1 / w^6


def _psi(qpt):
    """E'(F_q) -> E(F_q^6): (x / w^2, y / w^3) = (x w^4 / w^6, y w^3 / w^6)."""
    x, y = qpt
    return (x * _NEG_QUARTER % Q, y * _NEG_QUARTER % Q)        # coefficients of w^4 and of w^3


def _line(t, p2, lam, xq4, yq3):
    """l(Q) = (y_Q - y_T) - lam (x_Q - x_T) with x_Q = xq4 w^4, y_Q = yq3 w^3."""
    xt, yt = t
    return [(-yt + lam * xt) % Q, 0, 0, yq3, (-lam * xq4) % Q, 0]


def tate_pairing_unreduced(p, qpt):
    """Miller function f_{r,P} evaluated at psi(Q) (vertical lines dropped: x_Q lies in F_q^3)."""
    if p is None or qpt is None:
        return f6_one()
    xq4, yq3 = _psi(qpt)
    f, t = f6_one(), p
    for bit in bin(R)[3:]:
        lam = 3 * t[0] * t[0] * pow(2 * t[1], -1, Q) % Q
        f = f6_mul(f6_mul(f, f), _line(t, None, lam, xq4, yq3))
        t = O.BW6_G1.padd(t, t)
        if bit == "1":
            if t[0] == p[0]:                   # T = -P at the very end: the line is vertical, f unchanged
                t = None
                continue
            lam = (t[1] - p[1]) * pow(t[0] - p[0], -1, Q) % Q
            f = f6_mul(f, _line(t, p, lam, xq4, yq3))
            t = O.BW6_G1.padd(t, p)
    assert t is None
    return f


_FINAL_EXP = (Q ** 6 - 1) // R
assert (Q ** 6 - 1) % R == 0


def pairing_product(pairs) -> List[int]:
    """prod_i t(P_i, Q_i) in F_q^6 (one final exponentiation)."""
    f = f6_one()
    for p, qpt in pairs:
        f = f6_mul(f, tate_pairing_unreduced(p, qpt))
    return f6_pow(f, _FINAL_EXP)


# ---- Groth16 verification over BW6-761 --------------------------------------------------------------------


def parse_vk(raw: bytes):
    n = int.from_bytes(raw[384:392], "little")
    assert len(raw) == 392 + 96 * n
    g1 = lambda b: O.deserialize_compressed(O.BW6_G1, b)
    g2 = lambda b: O.deserialize_compressed(O.BW6_G2, b)
    return {"alpha": g1(raw[0:96]), "beta": g2(raw[96:192]), "gamma": g2(raw[192:288]), "delta": g2(raw[288:384]),
            "gamma_abc": [g1(raw[392 + 96 * i:392 + 96 * (i + 1)]) for i in range(n)]}


def parse_proof(raw: bytes):
    assert len(raw) == 288
    return (O.deserialize_compressed(O.BW6_G1, raw[0:96]), O.deserialize_compressed(O.BW6_G2, raw[96:192]),
            O.deserialize_compressed(O.BW6_G1, raw[192:288]))


def verify_proof(vk, proof, inputs: Sequence[int]) -> bool:
    a, b, c = proof
    if len(inputs) + 1 != len(vk["gamma_abc"]):
        return False
    g_ic = vk["gamma_abc"][0]
    for x, base in zip(inputs, vk["gamma_abc"][1:]):
        g_ic = O.BW6_G1.padd(g_ic, O.BW6_G1.pmul(base, x))
    lhs = pairing_product([(a, b), (g_ic, O.BW6_G2.pneg(vk["gamma"])), (c, O.BW6_G2.pneg(vk["delta"]))])
    rhs = pairing_product([(vk["alpha"], vk["beta"])])
    return lhs == rhs


# ---- public inputs: epoch-block edge hashes ----------------------------------------------------------------


def _le_bits(value: int, nbytes: int):
    return [(b >> i) & 1 for b in value.to_bytes(nbytes, "little") for i in range(8)]


def bytes_le_to_bits_be(bs: bytes, take: int):
    return [(b >> i) & 1 for b in bs for i in range(8)][:take][::-1]


def bytes_le_to_bits_le(bs: bytes, take: int):
    return bytes_le_to_bits_be(bs, take)[::-1]


def bits_be_to_bytes_le(bits):
    rev = bits[::-1]
    return bytes(sum(c << i for i, c in enumerate(rev[k:k + 8])) for k in range(0, len(rev), 8))


def encode_public_key(pk) -> List[int]:
    (x0, x1), (y0, y1) = pk
    half = (O.P - 1) // 2
    over_half = y1 > half or (y1 == 0 and y0 > half)
    return (bytes_le_to_bits_be(x0.to_bytes(48, "little"), 377) + bytes_le_to_bits_be(x1.to_bytes(48, "little"), 377)
            + [int(over_half)])


class EpochBlock:
    ENTROPY_BYTES = 16

    def __init__(self, index, round_, epoch_entropy: Optional[bytes], parent_entropy: Optional[bytes], maximum_non_signers,
                 maximum_validators, pubkeys):
        self.index, self.round, self.epoch_entropy, self.parent_entropy = index, round_, epoch_entropy, parent_entropy
        self.maximum_non_signers, self.maximum_validators, self.pubkeys = maximum_non_signers, maximum_validators, pubkeys

    @classmethod
    def from_ffi(cls, index, round_, epoch_entropy, parent_entropy, maximum_non_signers, maximum_validators, pubkey_bytes, num):
        keys = [O.deserialize_compressed(O.G2, pubkey_bytes[96 * i:96 * (i + 1)]) for i in range(num)]
        return cls(index, round_, epoch_entropy, parent_entropy, maximum_non_signers, maximum_validators, keys)

    def _entropy_bits(self, entropy):
        data = entropy if entropy is not None else bytes(self.ENTROPY_BYTES * 8)
        return bytes_le_to_bits_le(data, self.ENTROPY_BYTES * 8)

    def encode_to_bits_cip22(self, first: bool):
        bits = _le_bits(self.index, 2)
        bits += self._entropy_bits(self.parent_entropy if first else self.epoch_entropy)
        bits += _le_bits(self.maximum_non_signers, 4)
        for pk in self.pubkeys:
            bits += encode_public_key(pk)
        for _ in range(max(0, self.maximum_validators - len(self.pubkeys))):
            bits += encode_public_key(O.G2_GEN)
        return bits

    def encode_to_bits(self):
        """epoch_block.rs:106-114 (the pre-Donut encoding: index, maximum_non_signers, keys)."""
        bits = _le_bits(self.index, 2) + _le_bits(self.maximum_non_signers, 4)
        for pk in self.pubkeys:
            bits += encode_public_key(pk)
        return bits

    def encode_to_bytes(self):
        """epoch_block.rs:191-193."""
        return bits_be_to_bytes_le(self.encode_to_bits())

    def encode_first_epoch_to_bytes_cip22(self):
        """epoch_block.rs:184-188."""
        return bits_be_to_bytes_le(self.encode_to_bits_cip22(True))

    def encode_inner_to_bytes_cip22(self):
        """epoch_block.rs:152-171, 205-211: (inner bytes, extra-data bytes)."""
        extra = _le_bits(self.index, 2) + _le_bits(self.round, 1) + _le_bits(self.maximum_non_signers, 4)
        bits = self._entropy_bits(self.epoch_entropy) + self._entropy_bits(self.parent_entropy)
        for pk in self.pubkeys:
            bits += encode_public_key(pk)
        for _ in range(max(0, self.maximum_validators - len(self.pubkeys))):
            bits += encode_public_key(O.G2_GEN)
        return bits_be_to_bytes_le(bits), bits_be_to_bytes_le(extra)

    def blake2_first_epoch_cip22(self):
        return hash_to_bits(bits_be_to_bytes_le(self.encode_to_bits_cip22(True)))

    def blake2_last_epoch_with_aggregated_pk_cip22(self):
        agg = None
        for pk in self.pubkeys:
            agg = O.G2.padd(agg, pk)
        return hash_to_bits(bits_be_to_bytes_le(self.encode_to_bits_cip22(False) + encode_public_key(agg)))


def hash_to_bits(data: bytes):
    digest = hashlib.blake2s(data, digest_size=32, person=b"ULforout").digest()
    return bytes_le_to_bits_le(digest, 256)


def pack(bits, capacity=376):
    out = []
    for k in range(0, len(bits), capacity):
        v = 0
        for b in bits[k:k + capacity]:
            v = (v << 1) | b
        out.append(v)
    return out


def verify(vk_bytes: bytes, first: EpochBlock, last: EpochBlock, proof_bytes: bytes) -> bool:
    """epoch_snark::verify (api/verifier.rs:23-40)."""
    bits = first.blake2_first_epoch_cip22() + last.blake2_last_epoch_with_aggregated_pk_cip22()
    return verify_proof(parse_vk(vk_bytes), parse_proof(proof_bytes), pack(bits))
