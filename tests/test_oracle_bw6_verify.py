"""The reference's own known-answer test for the SNARK verifier
(crates/bls-snark-sys/src/snark/mod.rs:52-119, `simple_verifier_groth16_with_entropy`: a real BW6-761
Groth16 VK + proof + two epoch blocks, expected `true`) run through the oracle: pins BW6-761 G1/G2
arithmetic and decoding, the pairing check, the Blake2s edge hashes and the input packing.  CPU only."""
import json
import os

import pytest

from oracle import bw6_verify as V
from oracle import oracle as O

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))


def _blocks():
    first_keys = bytes.fromhex(GOLD["bls12_377_first_pubkeys"]["hex"])
    last_keys = bytes.fromhex(GOLD["bls12_377_last_pubkeys"]["hex"])
    first = V.EpochBlock.from_ffi(0, 0, bytes([1] * 16), bytes([2] * 16), 1, 4, first_keys, 4)      # mod.rs:80-89
    last = V.EpochBlock.from_ffi(2, 0, bytes([3] * 16), bytes([2] * 16), 1, 4, last_keys, 4)         # mod.rs:94-103
    return first, last


def test_reference_verifier_kat_is_true():
    vk = bytes.fromhex(GOLD["bw6_groth16_vk"]["hex"])
    proof = bytes.fromhex(GOLD["bw6_groth16_proof"]["hex"])
    first, last = _blocks()
    assert V.verify(vk, first, last, proof) is True


def test_reference_verifier_kat_rejects_changes():
    vk = bytes.fromhex(GOLD["bw6_groth16_vk"]["hex"])
    proof = bytes.fromhex(GOLD["bw6_groth16_proof"]["hex"])
    first, last = _blocks()
    last.index = 3                                      # a different public input
    assert V.verify(vk, first, last, proof) is False
    first, last = _blocks()
    a, b, c = V.parse_proof(proof)
    bad = (O.BW6_G1.padd(a, a), b, c)                   # a different proof element
    bits = first.blake2_first_epoch_cip22() + last.blake2_last_epoch_with_aggregated_pk_cip22()
    assert V.verify_proof(V.parse_vk(vk), bad, V.pack(bits)) is False


def test_tate_pairing_is_bilinear():
    vk = V.parse_vk(bytes.fromhex(GOLD["bw6_groth16_vk"]["hex"]))
    p, q = vk["alpha"], vk["beta"]
    e = V.pairing_product([(p, q)])
    assert e != V.f6_one() and V.f6_pow(e, V.R) == V.f6_one()
    assert V.pairing_product([(O.BW6_G1.pmul(p, 5), q)]) == V.f6_pow(e, 5)
    assert V.pairing_product([(p, O.BW6_G2.pmul(q, 7))]) == V.f6_pow(e, 7)
    assert V.pairing_product([(p, q), (O.BW6_G1.pneg(p), q)]) == V.f6_one()
