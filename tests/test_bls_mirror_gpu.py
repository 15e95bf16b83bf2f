"""The reference's behavioural tests for the callers of the hot path, restated against the
mirror API (crates/bls-crypto/src/bls/signature.rs:181-426; message hashing replaced by hash
points h_i * g1, which is what batch_verify_hashes takes in the reference)."""
import numpy as np
import pytest

from oracle import cref as C
from oracle import oracle as O

pytestmark = pytest.mark.gpu
L1, L2 = C.LAYOUTS["bls12_377_g1"], C.LAYOUTS["bls12_377_g2"]


@pytest.fixture(scope="module")
def bls():
    import torch
    from celo_bls_snark_rs_b200 import bls, engine
    engine.init(0)
    torch.cuda.set_device(0)
    return bls


def g1_image(k):      # k * g1 as a G1Projective image with a non-trivial Z (double-and-add output)
    return C.scalar_mul(L1, O.G1_GEN, k % O.R)


def g2_image(k):
    return C.scalar_mul(L2, O.G2_GEN, k % O.R)


def keygen_sign(rng, n, h):
    """n keys signing the hash point h * g1: returns (sks, pubkeys images, signature images)."""
    sks = [rng.below(O.R - 1) + 1 for _ in range(n)]
    return sks, [g2_image(s) for s in sks], [g1_image(s * h) for s in sks]


def test_aggregated_sig(bls):
    # signature.rs:181-229 test_aggregated_sig: two keys, same message; aggregate verifies under apk
    rng = O.SplitMix64(1)
    h = rng.below(O.R)
    sks, pks, sigs = keygen_sign(rng, 2, h)
    apk = bls.PublicKey.aggregate([bls.PublicKey(p) for p in pks])
    asig = bls.Signature.aggregate([bls.Signature(s) for s in sigs])
    apk.verify_hash(g1_image(h), asig)
    with pytest.raises(bls.VerificationFailed):
        apk.verify_hash(g1_image(h + 1), asig)
    with pytest.raises(bls.VerificationFailed):
        bls.PublicKey(pks[0]).verify_hash(g1_image(h), asig)
    # aggregation is the group sum
    assert L2.jacobian_to_affine(apk.image) == O.G2.pmul(O.G2_GEN, sum(sks) % O.R)


@pytest.mark.parametrize("num_batches,batch_size", [(1, 1), (3, 4), (7, 5)])
def test_batch_verify_hashes(bls, num_batches, batch_size):
    # signature.rs:329-361: per batch an aggregate key / aggregate signature over its own message
    rng = O.SplitMix64(10 * num_batches + batch_size)
    apks, hashes, asigs = [], [], []
    for _ in range(num_batches):
        h = rng.below(O.R)
        sks, pks, sigs = keygen_sign(rng, batch_size, h)
        apks.append(bls.PublicKey.aggregate([bls.PublicKey(p) for p in pks]))
        asigs.append(bls.Signature.aggregate([bls.Signature(s) for s in sigs]))
        hashes.append(g1_image(h))
    asig = bls.Signature.aggregate(asigs)
    asig.batch_verify_hashes(apks, hashes)
    with pytest.raises(bls.UnevenNumKeysMessages):
        asig.batch_verify_hashes(apks, hashes[:-1])
    if num_batches > 1:
        with pytest.raises(bls.VerificationFailed):
            asig.batch_verify_hashes(apks, hashes[1:] + hashes[:1])


def test_batch_verify_strict(bls):
    # signature.rs:390-426: 10 good signatures verify; adding one over a different message fails
    rng = O.SplitMix64(42)
    h = rng.below(O.R)
    sks, pks, sigs = keygen_sign(rng, 10, h)
    batch = bls.Batch.new()
    for p, s in zip(pks, sigs):
        batch.add(bls.PublicKey(p), bls.Signature(s))
    batch.verify_hash(g1_image(h))                       # random exponents, as the reference draws them
    batch.verify_hash(g1_image(h), exponents=[rng.below(1 << 136) for _ in range(10)])
    bad_sk = rng.below(O.R)
    batch.add(bls.PublicKey(g2_image(bad_sk)), bls.Signature(g1_image(bad_sk * (h + 5))))
    with pytest.raises(bls.VerificationFailed):
        batch.verify_hash(g1_image(h))


def test_batch_returns_none_on_length_mismatch_and_matches_oracle(bls):
    # signature.rs:76-80 / public.rs:53-56
    rng = O.SplitMix64(7)
    ks = [rng.below(O.R) for _ in range(5)]
    es = [rng.below(1 << 144) for _ in range(5)]
    sigs = [bls.Signature(g1_image(k)) for k in ks]
    pks = [bls.PublicKey(g2_image(k)) for k in ks]
    assert bls.Signature.batch(es[:4], sigs) is None
    assert bls.PublicKey.batch(es, pks[:3]) is None
    want = sum(k * e for k, e in zip(ks, es)) % O.R
    assert L1.jacobian_to_affine(bls.Signature.batch(es, sigs).image) == O.G1.pmul(O.G1_GEN, want)
    assert L2.jacobian_to_affine(bls.PublicKey.batch(es, pks).image) == O.G2.pmul(O.G2_GEN, want)


def test_exponent_byte_count(bls):
    assert bls.byte_count_from_target_batch_size(4096) == 18 and bls.byte_count_from_target_batch_size(20) == 17
