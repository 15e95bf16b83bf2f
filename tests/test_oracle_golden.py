"""Pins the Python oracle's building blocks against the reference's own golden
vectors.  CPU only."""
import json
import os

import pytest

from oracle import oracle as O

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))


def _hexes(key):
    return [bytes.fromhex(h) for h in GOLD[key]["hex"]]


@pytest.mark.parametrize("key", ["hash_to_g1_compat_pre_donut", "hash_to_g1_compat_cip22", "hash_to_g1_non_compat"])
def test_g1_golden_points_decode_on_curve_in_subgroup_and_reencode(key):
    # crates/bls-crypto/src/hash_to_curve/mod.rs:415-426,438-449,474-485
    for raw in _hexes(key):
        pt = O.deserialize_compressed(O.G1, raw)
        assert pt is not None and O.G1.on_curve(pt)
        assert O.G1.pmul(pt, O.R) is None            # hash output is cofactor-cleared
        assert O.serialize_compressed(O.G1, pt) == raw


def test_g1_compat_and_non_compat_vectors_share_x():
    # the two feature flavours differ only in the y-sign flag (mod.rs:415-426 vs :474-485)
    for a, b in zip(_hexes("hash_to_g1_compat_pre_donut"), _hexes("hash_to_g1_non_compat")):
        assert a[:-1] == b[:-1] and (a[-1] & 0x3F) == (b[-1] & 0x3F)


def test_g2_golden_points_decode_on_twist_in_subgroup_and_reencode():
    # crates/bls-crypto/src/hash_to_curve/mod.rs:497-508
    for raw in _hexes("hash_to_g2_non_compat"):
        pt = O.deserialize_compressed(O.G2, raw)
        assert pt is not None and O.G2.on_curve(pt)
        assert O.G2.pmul(pt, O.R) is None
        assert O.serialize_compressed(O.G2, pt) == raw


def _bytes_le_to_bits_be(bs, take):
    bits = [(b >> i) & 1 for b in bs for i in range(8)][:take]
    return bits[::-1]


def _bits_be_to_bytes_le(bits):
    rev = bits[::-1]
    return bytes(sum(c << i for i, c in enumerate(rev[k:k + 8])) for k in range(0, len(rev), 8))


def _le_bits(value, nbytes):
    return [(b >> i) & 1 for b in value.to_bytes(nbytes, "little") for i in range(8)]


def test_epoch_block_encoding_reproduced_from_oracle_g2_generator():
    """encode_first_epoch_to_bytes_cip22 (crates/epoch-snark/src/epoch_block.rs:106-131,
    encoding.rs:23-47) of 10 x G2 generator == golden hex at epoch_block.rs:243."""
    (x0, x1), y = O.G2_GEN
    assert O.G2.on_curve(O.G2_GEN) and O.G2.pmul(O.G2_GEN, O.R) is None
    half = (O.P - 1) // 2
    over_half = y[1] > half or (y[1] == 0 and y[0] > half)
    pk_bits = (_bytes_le_to_bits_be(x0.to_bytes(48, "little"), 377)
               + _bytes_le_to_bits_be(x1.to_bytes(48, "little"), 377) + [int(over_half)])
    bits = _le_bits(120, 2)
    bits += _bytes_le_to_bits_be(bytes([254] * 16), 128)[::-1]      # parent entropy (EpochType::First)
    bits += _le_bits(3, 4)
    bits += pk_bits * 10
    assert _bits_be_to_bytes_le(bits).hex() == GOLD["epoch_block_encoding_with_entropy"]["hex"]


def test_epoch_block_encodings_of_the_oracle_class():
    """The four encoding KATs of crates/epoch-snark/src/epoch_block.rs:243-320 through oracle/bw6_verify.EpochBlock
    (the class the GPU tests of encode_epoch_block_to_bytes[_cip22] compare with)."""
    from oracle import bw6_verify as V
    keys = [O.G2_GEN] * 10
    e255, e254 = bytes([255] * 16), bytes([254] * 16)
    assert V.EpochBlock(120, 5, e255, e254, 3, 10, keys).encode_first_epoch_to_bytes_cip22().hex() == \
        GOLD["epoch_block_encoding_with_entropy"]["hex"]
    assert V.EpochBlock(120, 5, None, None, 3, 10, keys).encode_first_epoch_to_bytes_cip22().hex() == \
        GOLD["epoch_block_encoding_without_entropy"]["hex"]
    assert V.EpochBlock(120, 5, e255, e254, 3, 11, keys).encode_first_epoch_to_bytes_cip22().hex() == \
        GOLD["epoch_block_encoding_with_entropy_padded"]["hex"]
    assert V.EpochBlock(120, 10, None, None, 3, 10, keys).encode_to_bytes().hex() == GOLD["epoch_block_encoding_before_donut"]["hex"]
    inner, extra = V.EpochBlock(120, 5, e255, e254, 3, 11, keys).encode_inner_to_bytes_cip22()
    assert len(inner) == (256 + 11 * 755 + 7) // 8 and len(extra) == 7


def test_ffi_pubkeys_decode_as_bls12_377_g2():
    # crates/bls-snark-sys/src/snark/mod.rs:56-58 : 4 + 4 compressed G2 keys of 96 bytes
    for key in ("bls12_377_first_pubkeys", "bls12_377_last_pubkeys"):
        raw = bytes.fromhex(GOLD[key]["hex"])
        assert len(raw) == 4 * 96
        for i in range(4):
            chunk = raw[96 * i:96 * (i + 1)]
            pt = O.deserialize_compressed(O.G2, chunk)
            assert O.G2.on_curve(pt) and O.G2.pmul(pt, O.R) is None
            assert O.serialize_compressed(O.G2, pt) == chunk


def test_bw6_groth16_vk_and_proof_decode():
    """VK = alpha_g1 | beta_g2 | gamma_g2 | delta_g2 | u64 len | gamma_abc_g1[len];
    proof = A(G1) | B(G2) | C(G1)   (crates/bls-snark-sys/src/snark/mod.rs:52-54)."""
    vk = bytes.fromhex(GOLD["bw6_groth16_vk"]["hex"])
    proof = bytes.fromhex(GOLD["bw6_groth16_proof"]["hex"])
    assert len(vk) == 680 and len(proof) == 288
    n = int.from_bytes(vk[384:392], "little")
    assert n == 3 and 392 + 96 * n == len(vk)
    layout = [(O.BW6_G1, vk[0:96]), (O.BW6_G2, vk[96:192]), (O.BW6_G2, vk[192:288]), (O.BW6_G2, vk[288:384])]
    layout += [(O.BW6_G1, vk[392 + 96 * i:392 + 96 * (i + 1)]) for i in range(n)]
    layout += [(O.BW6_G1, proof[0:96]), (O.BW6_G2, proof[96:192]), (O.BW6_G1, proof[192:288])]
    for curve, raw in layout:
        pt = O.deserialize_compressed(curve, raw)
        assert pt is not None and curve.on_curve(pt)
        assert curve.pmul(pt, O.R761) is None
        assert O.serialize_compressed(curve, pt) == raw


def test_field_constants():
    assert O.P.bit_length() == 377 and O.R.bit_length() == 253 and O.Q761.bit_length() == 761
    assert pow(O.NONRESIDUE, (O.P - 1) // 2, O.P) == O.P - 1          # -5 is a non-residue
    assert (-pow(O.P, -1, 1 << 64)) % (1 << 64) == 0x8508BFFFFFFFFFFF
    assert (-pow(O.R, -1, 1 << 64)) % (1 << 64) == 0x0A117FFFFFFFFFFF
    assert O.G1.pmul(O.G1_GEN, O.R) is None
