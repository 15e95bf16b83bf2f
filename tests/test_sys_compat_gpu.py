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
