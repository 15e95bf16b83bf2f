"""The reference's verifier known-answer test (crates/bls-snark-sys/src/snark/mod.rs:52-119) as oracle objects:
verifying key, proof and the packed public inputs derived from the two epoch blocks."""
import json
import os

from oracle import bw6_verify as V

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.json")))
VK = V.parse_vk(bytes.fromhex(GOLD["bw6_groth16_vk"]["hex"]))
PROOF = V.parse_proof(bytes.fromhex(GOLD["bw6_groth16_proof"]["hex"]))


def kat_blocks():
    first_keys = bytes.fromhex(GOLD["bls12_377_first_pubkeys"]["hex"])
    last_keys = bytes.fromhex(GOLD["bls12_377_last_pubkeys"]["hex"])
    first = V.EpochBlock.from_ffi(0, 0, bytes([1] * 16), bytes([2] * 16), 1, 4, first_keys, 4)      # mod.rs:80-89
    last = V.EpochBlock.from_ffi(2, 0, bytes([3] * 16), bytes([2] * 16), 1, 4, last_keys, 4)         # mod.rs:94-103
    return first, last


def kat_inputs():
    first, last = kat_blocks()
    return V.pack(first.blake2_first_epoch_cip22() + last.blake2_last_epoch_with_aggregated_pk_cip22())
