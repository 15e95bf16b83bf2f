"""The reference's SNARK verifier entry point on the GPU, driven through the symbol bls-snark-sys itself exports
(`verify`, crates/bls-snark-sys/src/snark/mod.rs:23-45), with the reference's own known-answer test restated
(`simple_verifier_groth16_with_entropy`, mod.rs:68-119: raw VK / proof / public-key bytes, expected true), and the
device-side point decoding (ark-serialize deserialize semantics) against the oracle on the same golden bytes."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bw6_kat import GOLD, kat_blocks, kat_inputs  # noqa: E402

from oracle import bw6_verify as V  # noqa: E402
from oracle import cref as C  # noqa: E402
from oracle import oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu

VK = bytes.fromhex(GOLD["bw6_groth16_vk"]["hex"])
PROOF = bytes.fromhex(GOLD["bw6_groth16_proof"]["hex"])
FIRST_KEYS = bytes.fromhex(GOLD["bls12_377_first_pubkeys"]["hex"])
LAST_KEYS = bytes.fromhex(GOLD["bls12_377_last_pubkeys"]["hex"])


@pytest.fixture(scope="module")
def eng():
    from celo_bls_snark_rs_b200 import engine as E
    E.init(0)
    return E


def blocks(eng, first_index=0, last_index=2, first_keys=FIRST_KEYS, last_keys=LAST_KEYS, max_validators=4, epoch_entropy=bytes([3] * 16),
           parent_entropy=bytes([2] * 16), max_non_signers=1):
    first = eng.EpochBlock(first_index, 0, bytes([1] * 16), parent_entropy, max_non_signers, max_validators, first_keys)   # mod.rs:80-89
    last = eng.EpochBlock(last_index, 0, epoch_entropy, bytes([2] * 16), max_non_signers, max_validators, last_keys)       # mod.rs:94-103
    return first, last


# ---- point decoding ------------------------------------------------------------------------------------------

def _check_decode(eng, kind, layout, curve, data):
    recs, status = eng.deserialize_points(kind, data)
    assert list(status) == [0] * (len(data) // 96)
    want = [O.deserialize_compressed(curve, data[96 * i:96 * (i + 1)]) for i in range(len(data) // 96)]
    assert layout.affine_from_records(recs) == want


def test_decode_reference_public_keys(eng):
    L = C.LAYOUTS["bls12_377_g2"]
    _check_decode(eng, eng.POINTS_BLS12_377_G2, L, O.G2, FIRST_KEYS + LAST_KEYS)
    # both y signs: re-encode the negated keys
    neg = b"".join(O.serialize_compressed(O.G2, O.G2.pneg(O.deserialize_compressed(O.G2, FIRST_KEYS[96 * i:96 * (i + 1)])))
                   for i in range(4))
    _check_decode(eng, eng.POINTS_BLS12_377_G2, L, O.G2, neg)


def test_decode_reference_key_and_proof(eng):
    g1 = VK[0:96] + VK[392:] + PROOF[0:96] + PROOF[192:288]
    g2 = VK[96:384] + PROOF[96:192]
    _check_decode(eng, eng.POINTS_BW6_761_G1, C.LAYOUTS["bw6_761_g1"], O.BW6_G1, g1)
    _check_decode(eng, eng.POINTS_BW6_761_G2, C.LAYOUTS["bw6_761_g2"], O.BW6_G2, g2)


def test_decode_rejections(eng):
    # x with no y on the curve
    x = 5
    while O.sqrt_mod((x ** 3 - 1) % O.Q761, O.Q761) is not None:
        x += 1
    off_curve = x.to_bytes(96, "little")
    # x >= modulus
    too_big = (O.Q761 + 3).to_bytes(96, "little")
    # on the curve, outside the prime-order subgroup (the G1 cofactor is huge: a random curve point is outside)
    x = 7
    while O.sqrt_mod((x ** 3 - 1) % O.Q761, O.Q761) is None:
        x += 1
    outside = x.to_bytes(96, "little")
    infinity = bytes(95) + bytes([0x40])
    recs, status = eng.deserialize_points(eng.POINTS_BW6_761_G1, off_curve + too_big + outside + infinity + VK[0:96])
    assert list(status) == [eng.DECODE_NOT_ON_CURVE, eng.DECODE_BAD_COORD, eng.DECODE_NOT_IN_SUBGROUP, eng.DECODE_INFINITY, eng.DECODE_OK]
    # deserialize_unchecked semantics: no subgroup test
    _, status = eng.deserialize_points(eng.POINTS_BW6_761_G1, outside, check_subgroup=False)
    assert list(status) == [eng.DECODE_OK]
    # BLS12-377 G2: a key whose x is off the twist, and one on the twist but outside G2
    p = O.P
    bad = None
    for c0 in range(1, 50):
        xx = (c0, 1)
        rhs = O.f2_add(O.f2_mul(O.f2_sqr(xx), xx), O.TWIST_B)
        y = O.f2_sqrt(rhs)
        if y is None and bad is None:
            bad = c0.to_bytes(48, "little") + (1).to_bytes(48, "little")
        if y is not None:
            outside2 = c0.to_bytes(48, "little") + (1).to_bytes(48, "little")
    _, status = eng.deserialize_points(eng.POINTS_BLS12_377_G2, bad + outside2 + (p + 1).to_bytes(48, "little") + bytes(48))
    assert list(status) == [eng.DECODE_NOT_ON_CURVE, eng.DECODE_NOT_IN_SUBGROUP, eng.DECODE_BAD_COORD]


# ---- public inputs --------------------------------------------------------------------------------------------

def _oracle_inputs(first_index, last_index, first_keys, last_keys, max_validators, epoch_entropy, parent_entropy, mns=1):
    first = V.EpochBlock.from_ffi(first_index, 0, bytes([1] * 16), parent_entropy, mns, max_validators, first_keys, len(first_keys) // 96)
    last = V.EpochBlock.from_ffi(last_index, 0, epoch_entropy, bytes([2] * 16), mns, max_validators, last_keys, len(last_keys) // 96)
    return V.pack(first.blake2_first_epoch_cip22() + last.blake2_last_epoch_with_aggregated_pk_cip22())


def test_public_inputs_match_oracle(eng):
    assert eng.epoch_public_inputs(*blocks(eng)) == kat_inputs()
    # generator padding (fewer keys than maximum_validators), None entropy, other indices, uneven key counts
    cases = [dict(max_validators=7), dict(epoch_entropy=None, parent_entropy=None), dict(first_index=41, last_index=1234, max_non_signers=3),
             dict(first_keys=FIRST_KEYS[:96], last_keys=LAST_KEYS + FIRST_KEYS[96:288], max_validators=9)]
    for kw in cases:
        o = dict(first_index=0, last_index=2, first_keys=FIRST_KEYS, last_keys=LAST_KEYS, max_validators=4,
                 epoch_entropy=bytes([3] * 16), parent_entropy=bytes([2] * 16), max_non_signers=1)
        o.update(kw)
        want = _oracle_inputs(o["first_index"], o["last_index"], o["first_keys"], o["last_keys"], o["max_validators"],
                              o["epoch_entropy"], o["parent_entropy"], o["max_non_signers"])
        assert eng.epoch_public_inputs(*blocks(eng, **kw)) == want, kw


def test_public_inputs_reject_bad_keys(eng):
    bad = bytearray(LAST_KEYS)
    bad[5] ^= 1                                                     # a different x: off the twist or outside G2
    assert eng.epoch_public_inputs(*blocks(eng, last_keys=bytes(bad))) is None


# ---- verify -----------------------------------------------------------------------------------------------------

def test_reference_kat_through_the_exported_verify(eng):
    """simple_verifier_groth16_with_entropy (snark/mod.rs:68-119): assert!(verify(...))."""
    first, last = blocks(eng)
    assert eng.verify_epochs(VK, PROOF, first, last) is True
    ok, _ = eng.verify_epochs_status(VK, PROOF, first, last)
    assert ok is True
    # the oracle agrees on the same bytes
    ofirst, olast = kat_blocks()
    assert V.verify(VK, ofirst, olast, PROOF) is True


def test_verify_rejects_what_the_reference_rejects(eng):
    f, l = blocks(eng, last_index=3)                                # a different public input
    assert eng.verify_epochs(VK, PROOF, f, l) is False
    f, l = blocks(eng, epoch_entropy=None)
    assert eng.verify_epochs(VK, PROOF, f, l) is False
    f, l = blocks(eng, max_validators=5)
    assert eng.verify_epochs(VK, PROOF, f, l) is False
    f, l = blocks(eng)
    assert eng.verify_epochs(VK, PROOF[:96] + PROOF[192:288] + PROOF[96:192], f, l) is False       # C where B belongs: not on the twist / wrong proof
    swapped = PROOF[192:288] + PROOF[96:192] + PROOF[0:96]         # valid points, wrong proof
    ok, why = eng.verify_epochs_status(VK, swapped, f, l)
    assert ok is False and "pairing check failed" in why
    assert eng.verify_epochs(VK[:400], PROOF, f, l) is False        # truncated key
    assert eng.verify_epochs(VK, PROOF[:200], f, l) is False        # truncated proof
    bad = bytearray(FIRST_KEYS)
    bad[100] ^= 4
    f, l = blocks(eng, first_keys=bytes(bad))                       # first block no longer decodes
    assert eng.verify_epochs(VK, PROOF, f, l) is False
