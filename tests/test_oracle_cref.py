"""C oracle (oracle/cpu_ref.c) vs the Python big-int arbiter (oracle/oracle.py). CPU only."""
import numpy as np
import pytest

from oracle import cref as C
from oracle import oracle as O

GENS = {
    "bls12_377_g1": O.G1_GEN,
    "bls12_377_g2": O.G2_GEN,
}


def _bw6_point(curve, rng):
    """random curve point (not necessarily in the r-torsion; irrelevant to MSM parity)."""
    while True:
        x = rng.below(O.Q761)
        y = O.sqrt_mod((x * x * x + curve.b) % O.Q761, O.Q761)
        if y is not None:
            return (x, y)


def gen_for(name, rng):
    if name in GENS:
        return GENS[name]
    return _bw6_point(O.CURVES[name], rng)


@pytest.mark.parametrize("name", list(O.CURVES))
def test_scalar_mul_and_codecs(name):
    L = C.LAYOUTS[name]
    rng = O.SplitMix64(1)
    g = gen_for(name, rng)
    for k in (0, 1, 2, 3, rng.below(L.curve.scalar_mod), L.curve.scalar_mod - 1):
        got = L.jacobian_to_affine(C.scalar_mul(L, g, k))
        assert got == L.curve.pmul(g, k), (name, k)


@pytest.mark.parametrize("name", list(O.CURVES))
@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 100])
def test_msm_matches_python_pippenger_and_naive(name, n):
    L = C.LAYOUTS[name]
    rng = O.SplitMix64(1000 + n)
    g = gen_for(name, rng)
    ks = [rng.below(L.curve.scalar_mod) for _ in range(n)]
    packed = C.fixed_base_batch(L, g, ks)
    pts = L.affine_from_records(packed)
    scalars = [rng.below(L.curve.scalar_mod) for _ in range(n)]
    if n >= 31:
        scalars[0] = 0                       # skipped
        scalars[1] = 1                       # unit-scalar fast path
        scalars[2] = L.curve.scalar_mod - 1
        pts[4] = pts[3]                      # duplicate bases (P + P inside a bucket)
        scalars[4] = scalars[3]
        pts[6] = L.curve.pneg(pts[5])        # P + (-P) inside a bucket
        scalars[6] = scalars[5]
        pts[7] = None                        # infinity base
    bases = L.affine_records(pts)
    want = L.curve.msm_naive(pts, scalars)
    got = L.jacobian_to_affine(C.msm(L, bases, L.scalars_array(scalars), threads=4))
    assert got == want
    if n <= 33:
        assert O.msm_pippenger(L.curve, pts, scalars) == want
    # packed layout (no flag byte) must agree as well
    got2 = L.jacobian_to_affine(C.msm(L, L.affine_records(pts, L.packed_stride), L.scalars_array(scalars), threads=1))
    assert got2 == want


def test_window_rule_matches_arkworks_table():
    # n=2^16->13, 2^20->15, 2^22->17, 2^24->18, 4096->10, 20->3
    assert [O.msm_window_bits(n) for n in (1 << 16, 1 << 20, 1 << 22, 1 << 24, 4096, 20)] == [13, 15, 17, 18, 10, 3]


def test_batch_exponent_bytes():
    # batch.rs:23-28: n = 4096 -> 18 bytes (144-bit exponents); capped at 31
    assert O.batch_exponent_bytes(4096) == 18
    assert O.batch_exponent_bytes(20) == 17
    assert O.batch_exponent_bytes(1 << 200) == 31


def test_c_direct_hash_to_g1_matches_python_oracle():
    """The C port of DIRECT_HASH_TO_G1 (the CPU figure of tools/bench_hash.py) against oracle/hash_to_curve.py,
    which is pinned on the reference's hasher KATs and hash-to-curve vectors."""
    from oracle import hash_to_curve as HC
    L = C.LAYOUTS["bls12_377_g1"]
    rng = O.SplitMix64(404)
    seen = set()
    for i in range(24):
        msg = bytes(rng.below(256) for _ in range(rng.below(200) if i else 0))
        extra = bytes(rng.below(256) for _ in range(rng.below(30) if i > 1 else 0))
        for compat in (True, False):
            image, att = C.hash_to_g1_direct(b"ULforxof", msg, extra, compat)
            pt, want_att = HC.try_and_increment(O.G1, HC.DIRECT, b"ULforxof", msg, extra, compat=compat)
            assert L.jacobian_compressed(image) == O.serialize_compressed(O.G1, pt) and att == want_att
            seen.add(att)
    assert len(seen) >= 3
    image, _ = C.hash_to_g1_direct(b"abc", b"short domain", b"")
    pt, _ = HC.try_and_increment(O.G1, HC.DIRECT, b"abc", b"short domain", b"")
    assert L.jacobian_compressed(image) == O.serialize_compressed(O.G1, pt)


def test_c_composite_hasher_reproduces_the_reference_vectors():
    """The C port of the composite hasher (Bowe-Hopwood CRH with the ChaCha20-derived generators) against the
    reference's own CRH KATs (hashers/composite.rs:104-131) and its 30 G1 hash-to-curve vectors
    (hash_to_curve/mod.rs:413-487) -- a second, independent pin next to oracle/hash_to_curve.py."""
    import json
    import os
    from oracle import hash_to_curve as HC
    V = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
    k = V["hasher_kats_composite"]["hex"]
    rng = HC.XorShiftRng(HC.REFERENCE_SEED)
    assert C.bh_crh(b"").hex() == k["test_crh_empty"][0]
    assert C.bh_crh(bytes(rng.gen_u8() for _ in range(32))).hex() == k["test_crh_random"][0]
    with pytest.raises(ValueError):
        C.bh_crh(bytes(19531))
    L = C.LAYOUTS["bls12_377_g1"]
    for key, compat, cip22 in (("hash_to_g1_compat_pre_donut", True, False), ("hash_to_g1_compat_cip22", True, True),
                               ("hash_to_g1_non_compat", False, False)):
        rng = HC.XorShiftRng(HC.REFERENCE_SEED)
        for want in V[key]["hex"]:
            d, m, e = HC.generate_test_data(rng)
            image, _ = C.hash_to_g1_composite(d, m, e, compat, cip22)
            assert L.jacobian_compressed(image).hex() == want
    # and against the Python oracle on a multi-window message, attempts included
    msg, extra = bytes(range(256)) * 3, b"extra"
    for cip22 in (False, True):
        image, att = C.hash_to_g1_composite(b"ULforxof", msg, extra, True, cip22)
        pt, want_att = HC.try_and_increment(O.G1, HC.COMPOSITE, b"ULforxof", msg, extra, compat=True, cip22=cip22)
        assert L.jacobian_compressed(image) == O.serialize_compressed(O.G1, pt) and att == want_att


# ---- C product of pairings (oracle/pairing_tmpl.h) vs the Python restatement -------------------------------
def _pairing_inputs(n, seed):
    rng = O.SplitMix64(seed)
    L1, L2 = C.LAYOUTS["bls12_377_g1"], C.LAYOUTS["bls12_377_g2"]
    p1 = L1.affine_from_records(C.fixed_base_batch(L1, O.G1_GEN, [rng.below(O.R) for _ in range(n)]))
    p2 = L2.affine_from_records(C.fixed_base_batch(L2, O.G2_GEN, [rng.below(O.R) for _ in range(n)]))
    return L1, L2, p1, p2


@pytest.mark.parametrize("n", [0, 1, 2, 5])
def test_c_multi_pairing_bytes_match_python(n):
    L1, L2, p1, p2 = _pairing_inputs(n, 40 + n)
    if n >= 5:
        p1[3] = None                         # pairs with an infinite member are skipped
        p2[4] = None
    want = O.product_of_pairings(list(zip(p1, p2)))
    for threads, stride in ((1, None), (3, "packed")):
        g1 = L1.affine_records(p1, L1.packed_stride if stride else None)
        g2 = L2.affine_records(p2, L2.packed_stride if stride else None)
        is_one, gt = C.multi_pairing(g1, g2, n, threads=threads)
        assert gt == C.fq12_to_ark_bytes(want), (n, threads)
        assert is_one == (want == O.FQ12_ONE)
    # the two halves separately: Miller value, then its final exponentiation
    _, mv = C.multi_pairing(L1.affine_records(p1), L2.affine_records(p2), n, phase=1)
    assert mv == C.fq12_to_ark_bytes(O.miller_loop(list(zip(p1, p2))))
    _, gt2 = C.multi_pairing(L1.affine_records(p1[:0]), L2.affine_records(p2[:0]), 0, phase=2, value=mv)
    assert gt2 == C.fq12_to_ark_bytes(want)


def test_c_multi_pairing_bilinear_check():
    """e(a G1, b G2) e(-ab G1, G2) == 1: the shape of the 2-pair check of public.rs:102"""
    rng = O.SplitMix64(77)
    a, b = rng.below(O.R), rng.below(O.R)
    L1, L2 = C.LAYOUTS["bls12_377_g1"], C.LAYOUTS["bls12_377_g2"]
    p1 = [O.G1.pmul(O.G1_GEN, a), O.G1.pneg(O.G1.pmul(O.G1_GEN, a * b % O.R))]
    p2 = [O.G2.pmul(O.G2_GEN, b), O.G2_GEN]
    ok, _ = C.multi_pairing(L1.affine_records(p1), L2.affine_records(p2))
    assert ok
    p1[1] = O.G1.pmul(O.G1_GEN, 5)
    ok, _ = C.multi_pairing(L1.affine_records(p1), L2.affine_records(p2))
    assert not ok


# ---- C radix-2 transforms (oracle/ntt_tmpl.h) vs oracle/ntt.py ------------------------------------------------
@pytest.mark.parametrize("fname", ["fr_bls12_377", "fr_bw6_761"])
@pytest.mark.parametrize("log_n", [0, 1, 3, 6])
def test_c_ntt_matches_python(fname, log_n):
    from oracle import ntt as N
    f = N.FIELDS[fname]
    rng = O.SplitMix64(9 + log_n)
    n = 1 << log_n
    vals = [rng.below(f.p) for _ in range(n)]
    arr = f.to_mont_array(vals)
    for inverse, coset, ref in ((0, 0, N.fft), (1, 0, N.ifft), (0, 1, N.coset_fft), (1, 1, N.coset_ifft)):
        got = f.from_mont_array(C.ntt(f.id, arr, log_n, bool(inverse), bool(coset), threads=3))
        assert got == ref(f, vals), (fname, log_n, inverse, coset)
    if log_n >= 1:
        b, c = [rng.below(f.p) for _ in range(n)], [rng.below(f.p) for _ in range(n)]
        got = f.from_mont_array(C.witness_map(f.id, arr, f.to_mont_array(b), f.to_mont_array(c), log_n, threads=2))
        assert got == N.witness_map(f, vals, b, c)
