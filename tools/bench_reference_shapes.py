"""The reference's own benchmark shapes (crates/bls-crypto/benches/batch_bls.rs:12-95: 300 epochs x 20 validators, messages
and extra data of 32 bytes, COMPOSITE_HASH_TO_G1_CIP22), through the symbols the CUDA library re-exports for bls-snark-sys
(verify_signature / batch_verify_signature / batch_verify_strict), with the C port timed beside each shape on the host:

    per-epoch aggregate screening       300 x PublicKey::verify                     -> 300 x verify_signature
    all epoch aggregate screening       one Signature::batch_verify over 300 msgs   -> batch_verify_signature
    per-epoch batch verification        300 x Batch::verify (n = 20)                -> batch_verify_strict, 300 batches in one call
                                                                                       (and as 300 calls of one batch)
    per-epoch individual verification   6000 x PublicKey::verify                    -> a sample of verify_signature calls

Small-n, many-call shapes are where a GPU path can lose to the CPU: this tool is the evidence either way.
    PYTHONPATH=. python tools/bench_reference_shapes.py [--epochs 300] [--validators 20]
Prints one JSON line.  Keys, hashes and signatures: hashes from the CUDA hash-to-G1 (pinned against the reference's vectors),
scalar multiplications from the C port -- valid by construction, checked by every shape verifying true and a corrupted one false."""
import argparse
import ctypes
import json
import time

import numpy as np

from celo_bls_snark_rs_b200 import engine as E
from oracle import cref as C
from oracle import oracle as O

L1, L2 = C.LAYOUTS["bls12_377_g1"], C.LAYOUTS["bls12_377_g2"]
SIG_DOMAIN = b"ULforxof"


def timed(fn, reps=1):
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) * 1e3 / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=300)
    ap.add_argument("--validators", type=int, default=20)
    ap.add_argument("--individual-sample", type=int, default=100)
    args = ap.parse_args()
    ne, nv = args.epochs, args.validators
    E.init(0)
    lib = E.load()
    rng = np.random.default_rng(2024)
    prng = O.SplitMix64(7)
    msgs = [(rng.integers(0, 256, 32, dtype=np.uint8).tobytes(), rng.integers(0, 256, 32, dtype=np.uint8).tobytes()) for _ in range(ne)]
    hashes, _ = E.hash_to_g1(E.HASHER_COMPOSITE, SIG_DOMAIN, msgs, compat=True, cip22=True)        # 144-byte G1Projective images
    h_aff = [L1.jacobian_to_affine(h) for h in hashes]
    sks = [[prng.below(O.R - 1) + 1 for _ in range(nv)] for _ in range(ne)]
    flat = [s for row in sks for s in row]
    pk_aff = C.fixed_base_batch(L2, O.G2_GEN, flat)                      # packed affine records
    one = L1.fe_to_mont_bytes(1)
    pk_img = [pk_aff[i].tobytes() + one + bytes(48) for i in range(len(flat))]                       # (x, y, 1) images, 288 B
    sig_img = [C.scalar_mul(L1, h_aff[e], sks[e][v]) for e in range(ne) for v in range(nv)]          # Jacobian images, 144 B
    keep = {"pk": [ctypes.create_string_buffer(b, 288) for b in pk_img], "sig": [ctypes.create_string_buffer(b, 144) for b in sig_img]}
    addr = ctypes.addressof
    vp = ctypes.c_void_p

    def aggregate(fn, bufs, size):
        arr = (vp * len(bufs))(*[addr(b) for b in bufs])
        out = vp()
        assert getattr(lib, fn)(arr, len(bufs), ctypes.byref(out)), fn
        img = ctypes.string_at(out, size)
        (lib.destroy_public_key if size == 288 else lib.destroy_signature)(out)
        return ctypes.create_string_buffer(img, size)

    apk = [aggregate("aggregate_public_keys", keep["pk"][e * nv:(e + 1) * nv], 288) for e in range(ne)]
    asig = [aggregate("aggregate_signatures", keep["sig"][e * nv:(e + 1) * nv], 144) for e in range(ne)]
    res = {"tool": "bench_reference_shapes", "epochs": ne, "validators": nv, "hasher": "COMPOSITE_HASH_TO_G1_CIP22", "shapes": {}}
    ok = ctypes.c_bool(False)

    # ---- per-epoch aggregate screening: 300 x PublicKey::verify ----
    def screening():
        good = True
        for e in range(ne):
            assert lib.verify_signature(addr(apk[e]), msgs[e][0], 32, msgs[e][1], 32, addr(asig[e]), True, True, ctypes.byref(ok))
            good = good and ok.value
        return good
    screening()
    gpu_ms, good = timed(screening)
    assert good

    def cpu_verify(pk_buf, sig_buf, m):
        h, _ = C.hash_to_g1_composite(SIG_DOMAIN, m[0], m[1], compat=True, cip22=True)
        p1 = L1.affine_records([L1.jacobian_to_affine(bytes(sig_buf)), L1.jacobian_to_affine(h)])
        p2 = L2.affine_records([O.G2.pneg(O.G2_GEN), L2.jacobian_to_affine(bytes(pk_buf))])
        return C.multi_pairing(p1, p2, 2)[0]

    k = min(ne, 40)
    cpu_ms, _ = timed(lambda: all(cpu_verify(apk[e], asig[e], msgs[e]) for e in range(k)))
    res["shapes"]["per_epoch_aggregate_screening"] = {"calls": ne, "gpu_ms": gpu_ms, "gpu_ms_per_call": gpu_ms / ne,
                                                      "cpu_port_ms": cpu_ms * ne / k, "cpu_port_ms_per_call": cpu_ms / k,
                                                      "cpu_sample_calls": k, "export": "verify_signature"}

    # ---- all epoch aggregate screening: one batch_verify over 300 messages ----
    marr = (E.MessageFFI * ne)(*[E.MessageFFI(E.FFIBuffer(m, 32), E.FFIBuffer(x, 32), addr(apk[e]), addr(asig[e]))
                                 for e, (m, x) in enumerate(msgs)])
    assert lib.batch_verify_signature(marr, ne, True, True, ctypes.byref(ok)) and ok.value
    gpu_ms, _ = timed(lambda: lib.batch_verify_signature(marr, ne, True, True, ctypes.byref(ok)), reps=5)
    assert ok.value

    def cpu_batch_verify():
        hs = [C.hash_to_g1_composite(SIG_DOMAIN, m, x, compat=True, cip22=True)[0] for m, x in msgs]
        total = None
        for b in asig:
            total = O.G1.padd(total, L1.jacobian_to_affine(bytes(b)))
        p1 = L1.affine_records([total] + [L1.jacobian_to_affine(h) for h in hs])
        p2 = L2.affine_records([O.G2.pneg(O.G2_GEN)] + [L2.jacobian_to_affine(bytes(b)) for b in apk])
        return C.multi_pairing(p1, p2, ne + 1)[0]
    cpu_ms, good = timed(cpu_batch_verify)
    assert good
    res["shapes"]["all_epoch_aggregate_screening"] = {"messages": ne, "gpu_ms": gpu_ms, "cpu_port_ms": cpu_ms, "export": "batch_verify_signature",
                                                      "note": "the CPU figure includes Python glue around the C hash / pairing calls (a few ms)"}

    # ---- per-epoch batch verification: 300 x Batch::verify(n = 20) ----
    pk_ptrs = [(vp * nv)(*[addr(b) for b in keep["pk"][e * nv:(e + 1) * nv]]) for e in range(ne)]
    sig_ptrs = [(vp * nv)(*[addr(b) for b in keep["sig"][e * nv:(e + 1) * nv]]) for e in range(ne)]
    barr = (E.BatchMessageFFI * ne)(*[E.BatchMessageFFI(E.FFIBuffer(m, 32), E.FFIBuffer(x, 32), pk_ptrs[e], nv, sig_ptrs[e], nv)
                                      for e, (m, x) in enumerate(msgs)])
    results = (ctypes.c_bool * ne)()
    assert lib.batch_verify_strict(barr, ne, True, True, results) and all(results)
    gpu_all_ms, _ = timed(lambda: lib.batch_verify_strict(barr, ne, True, True, results), reps=5)
    assert all(results)

    def strict_one_by_one():
        good = True
        for e in range(ne):
            good = lib.batch_verify_strict(ctypes.byref(barr[e]), 1, True, True, results) and good
        return good
    gpu_each_ms, good = timed(strict_one_by_one)
    assert good
    # one bad signature in one batch: that batch false, the others true, the call false
    saved = sig_ptrs[7][3]
    sig_ptrs[7][3] = addr(keep["sig"][0])
    assert not lib.batch_verify_strict(barr, ne, True, True, results)
    assert [bool(r) for r in results] == [e != 7 for e in range(ne)]
    sig_ptrs[7][3] = saved

    exp_bytes = min((128 + int(np.ceil(np.log2(nv))) + 7) // 8, 31)

    def cpu_batch(e):
        h, _ = C.hash_to_g1_composite(SIG_DOMAIN, msgs[e][0], msgs[e][1], compat=True, cip22=True)
        ex = np.zeros((nv, 4), dtype=np.uint64)
        ex.view(np.uint8).reshape(nv, 32)[:, :exp_bytes] = rng.integers(0, 256, (nv, exp_bytes), dtype=np.uint8)
        sg = np.frombuffer(b"".join(sig_img[e * nv:(e + 1) * nv]), dtype=np.uint8).reshape(nv, 144)
        g1 = L1.affine_records([L1.jacobian_to_affine(s.tobytes()) for s in sg])     # batch_normalization_into_affine
        bsig = C.msm(L1, g1, ex, threads=1)
        bpk = C.msm(L2, C.with_flags(L2, pk_aff[e * nv:(e + 1) * nv]), ex, threads=1)
        p1 = L1.affine_records([L1.jacobian_to_affine(bsig), L1.jacobian_to_affine(h)])
        p2 = L2.affine_records([O.G2.pneg(O.G2_GEN), L2.jacobian_to_affine(bpk)])
        return C.multi_pairing(p1, p2, 2)[0]
    k = min(ne, 40)
    cpu_ms, good = timed(lambda: all(cpu_batch(e) for e in range(k)))
    assert good
    res["shapes"]["per_epoch_batch_verification"] = {"batches": ne, "batch_size": nv, "gpu_ms_one_call": gpu_all_ms,
                                                     "gpu_ms_call_per_batch": gpu_each_ms, "cpu_port_ms": cpu_ms * ne / k,
                                                     "cpu_sample_batches": k, "export": "batch_verify_strict",
                                                     "note": "the CPU figure includes Python glue (point conversions) around the C calls"}

    # ---- per-epoch individual verification: 6000 x PublicKey::verify (sampled) ----
    k = min(args.individual_sample, ne * nv)

    def individual():
        good = True
        for i in range(k):
            e = i // nv
            assert lib.verify_signature(addr(keep["pk"][i]), msgs[e][0], 32, msgs[e][1], 32, addr(keep["sig"][i]), True, True, ctypes.byref(ok))
            good = good and ok.value
        return good
    gpu_ms, good = timed(individual)
    assert good
    kc = min(k, 30)
    cpu_ms, good = timed(lambda: all(cpu_verify(keep["pk"][i], keep["sig"][i], msgs[i // nv]) for i in range(kc)))
    assert good
    res["shapes"]["per_epoch_individual_verification"] = {"calls_in_shape": ne * nv, "gpu_ms_per_call": gpu_ms / k, "gpu_ms_projected": gpu_ms / k * ne * nv,
                                                          "cpu_port_ms_per_call": cpu_ms / kc, "cpu_port_ms_projected": cpu_ms / kc * ne * nv,
                                                          "sampled_calls": k, "export": "verify_signature"}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
