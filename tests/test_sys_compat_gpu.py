"""bls-snark-sys' own signature-verification entry points, called through the symbols the library re-exports
(include/bls_snark_sys_compat.h; crates/bls-snark-sys/src/signatures.rs:244-505, serialization.rs:35-266), with
bls-crypto's behavioural tests restated on top (crates/bls-crypto/src/bls/signature.rs:181-426: aggregated signatures,
batch_verify over epochs, strict batches with one bad entry).  Keys, hashes and signatures come from the oracle;
handles are made from compressed bytes by deserialize_*, as a cgo consumer makes them."""
import ctypes

import pytest

from oracle import cref as C
from oracle import hash_to_curve as H
from oracle import oracle as O

pytestmark = pytest.mark.gpu
L1, L2 = C.LAYOUTS["bls12_377_g1"], C.LAYOUTS["bls12_377_g2"]
SIG_DOMAIN, POP_DOMAIN = b"ULforxof", b"ULforpop"
HASHERS = {(True, True): (H.COMPOSITE, True), (True, False): (H.COMPOSITE, False), (False, False): (H.DIRECT, False)}


@pytest.fixture(scope="module")
def lib():
    from celo_bls_snark_rs_b200 import engine as E
    return E.load()                                   # no b200_init: the entry points bind the engine themselves


def _handle(lib, fn, data: bytes):
    out = ctypes.c_void_p()
    assert getattr(lib, fn)(data, len(data), ctypes.byref(out)), fn
    return out


def _bytes_of(lib, fn, handle) -> bytes:
    ptr, n = ctypes.c_void_p(), ctypes.c_int()
    assert getattr(lib, fn)(handle, ctypes.byref(ptr), ctypes.byref(n))
    data = ctypes.string_at(ptr, n.value)
    assert lib.free_vec(ptr, n.value)
    return data


def pk_bytes(sk):
    return O.serialize_compressed(O.G2, O.G2.pmul(O.G2_GEN, sk))


def sign_bytes(sk, msg, extra, cfg, domain=SIG_DOMAIN):
    hasher, cip22 = HASHERS[cfg]
    h, _ = H.try_and_increment(O.G1, hasher, domain, msg, extra, compat=True, cip22=cip22)
    return O.serialize_compressed(O.G1, O.G1.pmul(h, sk))


def test_struct_sizes():
    from celo_bls_snark_rs_b200 import engine as E
    assert ctypes.sizeof(E.MessageFFI) == 48 and ctypes.sizeof(E.BatchMessageFFI) == 64      # utils.rs:20-72


def test_serialization_round_trip_and_checks(lib):
    rng = O.SplitMix64(3)
    for _ in range(3):
        sk = rng.below(O.R - 1) + 1
        pkb, sgb = pk_bytes(sk), O.serialize_compressed(O.G1, O.G1.pmul(O.G1_GEN, sk))
        pk, sg = _handle(lib, "deserialize_public_key", pkb), _handle(lib, "deserialize_signature", sgb)
        # the handle is the Rust type's memory image: (x, y, 1) in Montgomery form
        assert L2.jacobian_to_affine(ctypes.string_at(pk, 288)) == O.G2.pmul(O.G2_GEN, sk)
        assert L1.jacobian_to_affine(ctypes.string_at(sg, 144)) == O.G1.pmul(O.G1_GEN, sk)
        assert _bytes_of(lib, "serialize_public_key", pk) == pkb
        assert _bytes_of(lib, "serialize_signature", sg) == sgb
        assert lib.destroy_public_key(pk) and lib.destroy_signature(sg)
    # infinity
    inf1, inf2 = O.serialize_compressed(O.G1, None), O.serialize_compressed(O.G2, None)
    sg, pk = _handle(lib, "deserialize_signature", inf1), _handle(lib, "deserialize_public_key", inf2)
    assert _bytes_of(lib, "serialize_signature", sg) == inf1 and _bytes_of(lib, "serialize_public_key", pk) == inf2
    # G1Affine::deserialize rejects: x off the curve, a curve point outside the prime-order subgroup, x >= p, short input
    out = ctypes.c_void_p()
    x = 1
    while O.sqrt_mod((x ** 3 + 1) % O.P, O.P) is not None:
        x += 1
    assert not lib.deserialize_signature(x.to_bytes(48, "little"), 48, ctypes.byref(out))
    x = 2
    while True:
        y = O.sqrt_mod((x ** 3 + 1) % O.P, O.P)
        if y is not None and O.G1.pmul((x, y), O.R) is not None:
            break
        x += 1
    assert not lib.deserialize_signature(O.serialize_compressed(O.G1, (x, y)), 48, ctypes.byref(out))
    assert not lib.deserialize_signature(((O.P + 5) | 0).to_bytes(48, "little"), 48, ctypes.byref(out))
    assert not lib.deserialize_signature(b"\x01" * 47, 47, ctypes.byref(out))
    assert not lib.deserialize_public_key(b"\x01" * 96, 96, ctypes.byref(out)) or lib.destroy_public_key(out)
    assert not lib.destroy_signature(None) and not lib.free_vec(None, 0)


@pytest.mark.parametrize("cfg", [(True, True), (True, False), (False, False)])
def test_verify_signature(lib, cfg):
    sk, msg, extra = 0x1234567890ABCDEF1234567, b"message to sign", b"extra"
    pk = _handle(lib, "deserialize_public_key", pk_bytes(sk))
    sg = _handle(lib, "deserialize_signature", sign_bytes(sk, msg, extra, cfg))
    ok = ctypes.c_bool(False)
    assert lib.verify_signature(pk, msg, len(msg), extra, len(extra), sg, cfg[0], cfg[1], ctypes.byref(ok)) and ok.value
    assert lib.verify_signature(pk, msg + b"!", len(msg) + 1, extra, len(extra), sg, cfg[0], cfg[1], ctypes.byref(ok)) and not ok.value
    assert lib.verify_signature(pk, msg, len(msg), extra, len(extra) - 1, sg, cfg[0], cfg[1], ctypes.byref(ok)) and not ok.value
    other = tuple(k for k in HASHERS if k != cfg)[0]
    assert lib.verify_signature(pk, msg, len(msg), extra, len(extra), sg, other[0], other[1], ctypes.byref(ok)) and not ok.value
    # (false, true) is BLSError::HashToCurveError in the reference: the call itself fails
    assert not lib.verify_signature(pk, msg, len(msg), extra, len(extra), sg, False, True, ctypes.byref(ok))


def test_verify_pop_and_aggregates(lib):
    rng = O.SplitMix64(8)
    sks = [rng.below(O.R - 1) + 1 for _ in range(4)]
    msg = b"proof of possession"
    pks = [_handle(lib, "deserialize_public_key", pk_bytes(s)) for s in sks]
    pops = [_handle(lib, "deserialize_signature", sign_bytes(s, msg, b"", (False, False), POP_DOMAIN)) for s in sks]
    ok = ctypes.c_bool(False)
    assert lib.verify_pop(pks[0], msg, len(msg), pops[0], ctypes.byref(ok)) and ok.value
    assert lib.verify_pop(pks[1], msg, len(msg), pops[0], ctypes.byref(ok)) and not ok.value
    # signature.rs:181-229 test_aggregated_sig: the aggregate verifies under the aggregate key
    apk, asig = ctypes.c_void_p(), ctypes.c_void_p()
    arr_pk = (ctypes.c_void_p * 4)(*[p.value for p in pks])
    arr_sg = (ctypes.c_void_p * 4)(*[s.value for s in pops])
    assert lib.aggregate_public_keys(arr_pk, 4, ctypes.byref(apk)) and lib.aggregate_signatures(arr_sg, 4, ctypes.byref(asig))
    assert _bytes_of(lib, "serialize_public_key", apk) == pk_bytes(sum(sks) % O.R)
    assert lib.verify_pop(apk, msg, len(msg), asig, ctypes.byref(ok)) and ok.value
    assert lib.verify_pop(apk, msg, len(msg), pops[0], ctypes.byref(ok)) and not ok.value


def _epochs(lib, rng, num_epochs, num_validators, cfg):
    """signature.rs:232-326 test_batch_verify: per epoch one message, its validators' aggregate key and signature."""
    from celo_bls_snark_rs_b200 import engine as E
    msgs, handles = [], []
    for e in range(num_epochs):
        data, extra = bytes(rng.below(256) for _ in range(32)), bytes(rng.below(256) for _ in range(e % 3))
        sks = [rng.below(O.R - 1) + 1 for _ in range(num_validators)]
        agg = sum(sks) % O.R                                              # aggregate key / signature of the epoch
        pk = _handle(lib, "deserialize_public_key", pk_bytes(agg))
        sg = _handle(lib, "deserialize_signature", sign_bytes(agg, data, extra, cfg))
        handles.append((pk, sg, data, extra))
        msgs.append(E.MessageFFI(E.FFIBuffer(data, len(data)), E.FFIBuffer(extra, len(extra)), pk.value, sg.value))
    return (E.MessageFFI * num_epochs)(*msgs), handles


@pytest.mark.parametrize("cfg", [(True, True), (False, False)])
def test_batch_verify_signature(lib, cfg):
    arr, keep = _epochs(lib, O.SplitMix64(31), 6, 3, cfg)
    ok = ctypes.c_bool(False)
    assert lib.batch_verify_signature(arr, 6, cfg[0], cfg[1], ctypes.byref(ok)) and ok.value
    arr[2].sig, arr[3].sig = arr[3].sig, arr[2].sig          # the aggregate is unchanged, the pairing product too
    assert lib.batch_verify_signature(arr, 6, cfg[0], cfg[1], ctypes.byref(ok)) and ok.value
    arr[2].public_key, arr[3].public_key = arr[3].public_key, arr[2].public_key
    assert lib.batch_verify_signature(arr, 6, cfg[0], cfg[1], ctypes.byref(ok)) and not ok.value
    assert not lib.batch_verify_signature(arr, 6, False, True, ctypes.byref(ok))
    assert keep


def test_batch_verify_strict(lib):
    """signature.rs:390-426 test_batch_verify_strict: ten good entries verify; one entry signed over another message
    makes its batch fail, the other batches keep their own results, and the call returns false."""
    from celo_bls_snark_rs_b200 import engine as E
    rng = O.SplitMix64(57)
    cfg = (True, True)
    batches, keep = [], []
    for b in range(3):
        data, extra = bytes(rng.below(256) for _ in range(24 + b)), b"\x07" * b
        sks = [rng.below(O.R - 1) + 1 for _ in range(10)]
        pks = [_handle(lib, "deserialize_public_key", pk_bytes(s)) for s in sks]
        sgs = [_handle(lib, "deserialize_signature", sign_bytes(s, data, extra, cfg)) for s in sks]
        if b == 1:                                          # the 11th signature is over a different message
            bad = rng.below(O.R - 1) + 1
            pks.append(_handle(lib, "deserialize_public_key", pk_bytes(bad)))
            sgs.append(_handle(lib, "deserialize_signature", sign_bytes(bad, data + b"x", extra, cfg)))
        apk = (ctypes.c_void_p * len(pks))(*[p.value for p in pks])
        asg = (ctypes.c_void_p * len(sgs))(*[s.value for s in sgs])
        keep.append((data, extra, pks, sgs, apk, asg))
        batches.append(E.BatchMessageFFI(E.FFIBuffer(data, len(data)), E.FFIBuffer(extra, len(extra)), apk, len(pks), asg, len(sgs)))
    arr = (E.BatchMessageFFI * 3)(*batches)
    results = (ctypes.c_bool * 3)()
    assert not lib.batch_verify_strict(arr, 3, True, True, results)
    assert list(results) == [True, False, True]
    good = (E.BatchMessageFFI * 2)(batches[0], batches[2])
    results2 = (ctypes.c_bool * 2)()
    assert lib.batch_verify_strict(good, 2, True, True, results2) and list(results2) == [True, True]
    # wrong hasher for these signatures: every batch fails; (false, true) marks every batch false as well
    assert not lib.batch_verify_strict(good, 2, False, False, results2) and list(results2) == [False, False]
    assert not lib.batch_verify_strict(good, 2, False, True, results2) and list(results2) == [False, False]


# ---- private keys, signing, hash helpers, uncompressed encodings (csrc/sys_compat_keys.cu) ---------------------------
def _out_bytes(lib, call):
    ptr, n = ctypes.c_void_p(), ctypes.c_int()
    assert call(ctypes.byref(ptr), ctypes.byref(n))
    data = ctypes.string_at(ptr, n.value)
    assert lib.free_vec(ptr, n.value)
    return data


def test_private_key_sign_and_verify_round_trip(lib):
    """test_simple_sig / test_pop of crates/bls-crypto/src/bls/signature.rs:158-180 and secret.rs through the exported symbols:
    keys and signatures made by the library verify, against the oracle's own values too."""
    assert lib.init()
    rng = O.SplitMix64(21)
    sk_int = rng.below(O.R - 1) + 1
    sk = _handle(lib, "deserialize_private_key", sk_int.to_bytes(32, "little"))
    assert _out_bytes(lib, lambda p, n: lib.serialize_private_key(sk, p, n)) == sk_int.to_bytes(32, "little")
    pk = ctypes.c_void_p()
    assert lib.private_key_to_public_key(sk, ctypes.byref(pk))
    assert _bytes_of(lib, "serialize_public_key", pk) == pk_bytes(sk_int)
    msg, extra = b"hello celo", b"xtra"
    for cfg in HASHERS:
        sig = ctypes.c_void_p()
        assert lib.sign_message(sk, msg, len(msg), extra, len(extra), cfg[0], cfg[1], ctypes.byref(sig))
        assert _bytes_of(lib, "serialize_signature", sig) == sign_bytes(sk_int, msg, extra, cfg)
        ok = ctypes.c_bool(False)
        assert lib.verify_signature(pk, msg, len(msg), extra, len(extra), sig, cfg[0], cfg[1], ctypes.byref(ok)) and ok.value
        assert lib.verify_signature(pk, b"other", 5, extra, len(extra), sig, cfg[0], cfg[1], ctypes.byref(ok)) and not ok.value
        assert lib.destroy_signature(sig)
    out = ctypes.c_void_p()
    assert not lib.sign_message(sk, msg, len(msg), extra, len(extra), False, True, ctypes.byref(out))      # (false, true): HashToCurveError
    pop = ctypes.c_void_p()
    pkb = pk_bytes(sk_int)
    assert lib.sign_pop(sk, pkb, len(pkb), ctypes.byref(pop))
    ok = ctypes.c_bool(False)
    assert lib.verify_pop(pk, pkb, len(pkb), pop, ctypes.byref(ok)) and ok.value
    # a generated key is a valid scalar and signs verifiably
    gen = ctypes.c_void_p()
    assert lib.generate_private_key(ctypes.byref(gen))
    g_int = int.from_bytes(_out_bytes(lib, lambda p, n: lib.serialize_private_key(gen, p, n)), "little")
    assert 0 < g_int < O.R
    gpk, gsig = ctypes.c_void_p(), ctypes.c_void_p()
    assert lib.private_key_to_public_key(gen, ctypes.byref(gpk)) and lib.sign_message(gen, msg, len(msg), None, 0, True, True, ctypes.byref(gsig))
    assert lib.verify_signature(gpk, msg, len(msg), None, 0, gsig, True, True, ctypes.byref(ok)) and ok.value
    assert not lib.deserialize_private_key(O.R.to_bytes(32, "little"), 32, ctypes.byref(out))                  # not below the order
    for h in (sk, gen):
        assert lib.destroy_private_key(h)


def test_hash_helpers_and_uncompressed_encodings(lib):
    msg, extra = b"some message", b"\x01\x02"
    fe = lambda v: v.to_bytes(48, "little")
    # hash_direct: G1Affine::write = canonical x | y | infinity byte
    for pop in (False, True):
        pt, att = H.try_and_increment(O.G1, H.DIRECT, POP_DOMAIN if pop else SIG_DOMAIN, msg, b"", compat=True, cip22=False)
        want = fe(pt[0]) + fe(pt[1]) + b"\x00"
        assert _out_bytes(lib, lambda p, n: lib.hash_direct(msg, len(msg), p, n, pop)) == want
        a = ctypes.c_int(-1)
        assert _out_bytes(lib, lambda p, n: lib.hash_direct_with_attempt(msg, len(msg), p, n, ctypes.byref(a), pop)) == want and a.value == att
    # hash_composite / _cip22: the point, as canonical (x, y, 1)
    pt, _ = H.try_and_increment(O.G1, H.COMPOSITE, SIG_DOMAIN, msg, extra, compat=True, cip22=False)
    assert _out_bytes(lib, lambda p, n: lib.hash_composite(msg, len(msg), extra, len(extra), p, n)) == fe(pt[0]) + fe(pt[1]) + fe(1)
    pt, att = H.try_and_increment(O.G1, H.COMPOSITE, SIG_DOMAIN, msg, extra, compat=True, cip22=True)
    c = ctypes.c_uint8(255)
    got = _out_bytes(lib, lambda p, n: lib.hash_composite_cip22(msg, len(msg), extra, len(extra), p, n, ctypes.byref(c)))
    assert got == fe(pt[0]) + fe(pt[1]) + fe(1) and c.value == att
    assert _out_bytes(lib, lambda p, n: lib.hash_crh(msg, len(msg), 48, p, n)) == H.COMPOSITE[0](SIG_DOMAIN, msg, 48)
    # uncompressed encodings and the key subtraction
    rng = O.SplitMix64(33)
    sks = [rng.below(O.R - 1) + 1 for _ in range(4)]
    pks = [_handle(lib, "deserialize_public_key_cached", pk_bytes(s)) for s in sks]
    sig = _handle(lib, "deserialize_signature", O.serialize_compressed(O.G1, O.G1.pmul(O.G1_GEN, sks[0])))
    assert _out_bytes(lib, lambda p, n: lib.serialize_public_key_uncompressed(pks[0], p, n)) == \
        O.serialize_uncompressed(O.G2, O.G2.pmul(O.G2_GEN, sks[0]))
    assert _out_bytes(lib, lambda p, n: lib.serialize_signature_uncompressed(sig, p, n)) == \
        O.serialize_uncompressed(O.G1, O.G1.pmul(O.G1_GEN, sks[0]))
    arr = (ctypes.c_void_p * 4)(*[p.value for p in pks])
    agg = ctypes.c_void_p()
    assert lib.aggregate_public_keys(arr, 4, ctypes.byref(agg))
    sub = (ctypes.c_void_p * 2)(pks[1].value, pks[3].value)
    rest = ctypes.c_void_p()
    assert lib.aggregate_public_keys_subtract(agg, sub, 2, ctypes.byref(rest))
    assert _bytes_of(lib, "serialize_public_key", rest) == pk_bytes((sks[0] + sks[2]) % O.R)
    inf = _handle(lib, "deserialize_signature", O.serialize_compressed(O.G1, None))
    assert _out_bytes(lib, lambda p, n: lib.serialize_signature_uncompressed(inf, p, n)) == O.serialize_uncompressed(O.G1, None)


def test_epoch_block_byte_exports(lib):
    """encode_epoch_block_to_bytes / _cip22 (crates/bls-snark-sys/src/snark/epoch_block.rs:16-106) on key HANDLES: the keys are
    normalised and compressed on the device, the bit strings assembled on the host.  The pre-Donut KAT of
    crates/epoch-snark/src/epoch_block.rs:285-295 through the exported symbol, then random keys against the oracle class."""
    import json
    import os
    from oracle import bw6_verify as V
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
    gen = [_handle(lib, "deserialize_public_key", O.serialize_compressed(O.G2, O.G2_GEN)) for _ in range(10)]
    arr = (ctypes.c_void_p * 10)(*[h.value for h in gen])
    got = _out_bytes(lib, lambda p, n: lib.encode_epoch_block_to_bytes(120, 3, arr, 10, p, n))
    assert got.hex() == gold["epoch_block_encoding_before_donut"]["hex"]
    rng = O.SplitMix64(77)
    sks = [rng.below(O.R - 1) + 1 for _ in range(5)]
    pts = [O.G2.pmul(O.G2_GEN, s) for s in sks]
    hs = [_handle(lib, "deserialize_public_key", O.serialize_compressed(O.G2, p)) for p in pts]
    # an aggregate is a non-normalised Jacobian image: the device has to normalise it
    agg = ctypes.c_void_p()
    arr5 = (ctypes.c_void_p * 5)(*[h.value for h in hs])
    assert lib.aggregate_public_keys(arr5, 5, ctypes.byref(agg))
    agg_pt = None
    for p in pts:
        agg_pt = O.G2.padd(agg_pt, p)
    keys, key_pts = (ctypes.c_void_p * 6)(*([h.value for h in hs] + [agg.value])), pts + [agg_pt]
    assert _out_bytes(lib, lambda p, n: lib.encode_epoch_block_to_bytes(7, 2, keys, 6, p, n)) == \
        V.EpochBlock(7, 0, None, None, 2, 6, key_pts).encode_to_bytes()
    for ee, pe, maxv in ((bytes(range(16)), bytes(range(16, 32)), 6), (None, bytes([9] * 16), 8), (bytes([1] * 16), None, 3)):
        inner, n_in, extra, n_ex = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_void_p(), ctypes.c_int()
        assert lib.encode_epoch_block_to_bytes_cip22(513, 9, ee, pe, 4, maxv, keys, 6, ctypes.byref(inner), ctypes.byref(n_in),
                                                     ctypes.byref(extra), ctypes.byref(n_ex))
        want_inner, want_extra = V.EpochBlock(513, 9, ee, pe, 4, maxv, key_pts).encode_inner_to_bytes_cip22()
        assert ctypes.string_at(inner, n_in.value) == want_inner and ctypes.string_at(extra, n_ex.value) == want_extra
        assert lib.free_vec(inner, n_in.value) and lib.free_vec(extra, n_ex.value)
    # no keys at all: only the header (and the padding)
    assert _out_bytes(lib, lambda p, n: lib.encode_epoch_block_to_bytes(1, 1, None, 0, p, n)) == V.EpochBlock(1, 0, None, None, 1, 0, []).encode_to_bytes()
    msg = b"first step"
    assert _out_bytes(lib, lambda p, n: lib.hash_direct_first_step(msg, len(msg), 96, p, n)) == H.direct_hash(SIG_DOMAIN, msg, 96)
