"""Extracts the golden vectors the reference's own tests hold for the building
blocks of the hot path into tests/golden/reference_vectors.json.

Run once in the build container (where /root/reference exists):
    python tests/golden/extract_reference_vectors.py
The GPU box has no /root/reference; tests read only the committed JSON.
"""
import json
import os
import re

REF = "/root/reference/crates"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")


def hex_strings_in(path, lo, hi, min_len=64):
    with open(path) as f:
        lines = f.readlines()[lo - 1:hi]
    return [m for ln in lines for m in re.findall(r'"([0-9a-f]{%d,})"' % min_len, ln)]


def const_str(path, name):
    src = open(path).read()
    m = re.search(name + r'\s*:\s*&str\s*=\s*"([0-9a-f]+)"', src)
    return m.group(1)


h2c = f"{REF}/bls-crypto/src/hash_to_curve/mod.rs"
snark = f"{REF}/bls-snark-sys/src/snark/mod.rs"
epoch = f"{REF}/epoch-snark/src/epoch_block.rs"

vec = {
    "_source": "celo-org/celo-bls-snark-rs @ 1c59d25, extracted by tests/golden/extract_reference_vectors.py",
    "hash_to_g1_compat_pre_donut": {"cite": "crates/bls-crypto/src/hash_to_curve/mod.rs:415-426",
                                    "hex": hex_strings_in(h2c, 413, 427)},
    "hash_to_g1_compat_cip22": {"cite": "crates/bls-crypto/src/hash_to_curve/mod.rs:438-449",
                                "hex": hex_strings_in(h2c, 436, 450)},
    "hash_to_g1_non_compat": {"cite": "crates/bls-crypto/src/hash_to_curve/mod.rs:474-485",
                              "hex": hex_strings_in(h2c, 472, 486)},
    "hash_to_g2_non_compat": {"cite": "crates/bls-crypto/src/hash_to_curve/mod.rs:497-508",
                              "hex": hex_strings_in(h2c, 495, 509)},
    "bw6_groth16_vk": {"cite": "crates/bls-snark-sys/src/snark/mod.rs:54", "hex": const_str(snark, "ENTROPY_VK")},
    "bw6_groth16_proof": {"cite": "crates/bls-snark-sys/src/snark/mod.rs:52", "hex": const_str(snark, "ENTROPY_PROOF")},
    "bls12_377_first_pubkeys": {"cite": "crates/bls-snark-sys/src/snark/mod.rs:56",
                                "hex": const_str(snark, "ENTROPY_FIRST_PUBKEYS")},
    "bls12_377_last_pubkeys": {"cite": "crates/bls-snark-sys/src/snark/mod.rs:58",
                               "hex": const_str(snark, "ENTROPY_LAST_PUBKEYS")},
    "epoch_block_encoding_with_entropy": {"cite": "crates/epoch-snark/src/epoch_block.rs:243",
                                          "hex": const_str(epoch, "EXPECTED_ENCODING_WITH_ENTROPY")},
    "epoch_block_encoding_with_entropy_padded": {"cite": "crates/epoch-snark/src/epoch_block.rs:244",
                                                 "hex": const_str(epoch, "EXPECTED_ENCODING_WITH_ENTROPY_PADDED")},
    "epoch_block_encoding_without_entropy": {"cite": "crates/epoch-snark/src/epoch_block.rs:245",
                                             "hex": const_str(epoch, "EXPECTED_ENCODING_WITHOUT_ENTROPY")},
    "epoch_block_encoding_before_donut": {"cite": "crates/epoch-snark/src/epoch_block.rs:246",
                                          "hex": const_str(epoch, "EXPECTED_ENCODING_BEFORE_DONUT")},
}


def test_fn_hex(path, fn_name):
    """All hex literals (>= 4 digits) inside one #[test] function body."""
    src = open(path).read()
    body = src[src.index("fn " + fn_name + "("):]
    nxt = body.find("#[test]")
    body = body if nxt < 0 else body[:nxt]
    return re.findall(r'"([0-9a-f]{4,})"', body)


direct = f"{REF}/bls-crypto/src/hashers/direct.rs"
composite = f"{REF}/bls-crypto/src/hashers/composite.rs"
for label, path, cite in (("direct", direct, "crates/bls-crypto/src/hashers/direct.rs:87-171"),
                          ("composite", composite, "crates/bls-crypto/src/hashers/composite.rs:104-189")):
    fns = ["test_crh_empty", "test_crh_random", "test_xof_random_96", "test_hash_random"]
    if label == "composite":
        fns += ["test_xof_random_768", "test_xof_random_769"]
    else:
        fns += ["test_blake2s_test_vectors"]
    vec["hasher_kats_" + label] = {"cite": cite, "hex": {fn: test_fn_hex(path, fn) for fn in fns}}

for k, v in vec.items():
    if isinstance(v, dict):
        n = len(v["hex"]) if isinstance(v["hex"], list) else len(v["hex"]) // 2
        print(k, n)
json.dump(vec, open(OUT, "w"), indent=1)
