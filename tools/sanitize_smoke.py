"""Small end-to-end pass over every kernel family, meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
Checks results against the oracle as it goes (so a sanitizer-clean run is also a correct run)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from celo_bls_snark_rs_b200 import engine as E   # noqa: E402
from oracle import cref as C                      # noqa: E402
from oracle import inputs as H                    # noqa: E402
from oracle import ntt as N                       # noqa: E402
from oracle import oracle as O                    # noqa: E402


def main():
    E.init(0)
    dev = torch.device("cuda:0")
    # MSM: all curves, edge cases, one over-populated bucket, batch pipeline
    for name, n in (("bls12_377_g1", 700), ("bls12_377_g2", 200), ("bw6_761_g1", 200)):
        L = C.LAYOUTS[name]
        pts, scalars = H.edge_case_inputs(name, n, 7)
        bases, sc = L.affine_records(pts), L.scalars_array(scalars)
        sc[50:50 + n // 3] = 0
        sc[50:50 + n // 3, 0] = 0x2B
        want = L.jacobian_compressed(C.msm(L, bases, sc))
        assert L.jacobian_compressed(E.msm(L.id, bases, sc)) == want, name
        packed = torch.from_numpy(L.affine_records(pts, L.packed_stride).copy()).to(dev)
        d_sc = torch.from_numpy(sc.view(np.int64).copy()).to(dev)
        out = torch.zeros((3, L.jac_bytes), dtype=torch.uint8, device=dev)
        E.msm_batch_device(L.id, [(packed.data_ptr(), d_sc.data_ptr(), n, out[j].data_ptr()) for j in range(3)])
        E.sync()
        assert all(L.jacobian_compressed(out[j].cpu().numpy().tobytes()) == want for j in range(3)), name
    # pairing
    L1, L2 = C.LAYOUTS["bls12_377_g1"], C.LAYOUTS["bls12_377_g2"]
    g1, g2 = H.signature_batch(4, 3)
    ok, _ = E.multi_pairing(L1.affine_records(g1), L2.affine_records(g2))
    assert ok is True
    # NTT + witness map
    for f in N.FIELDS.values():
        rng = O.SplitMix64(5)
        a = [rng.below(f.p) for _ in range(1 << 11)]
        d = torch.from_numpy(f.to_mont_array(a).view(np.int64).copy()).to(dev)
        E.ntt_device(f.id, d.data_ptr(), 11, False, True)
        E.sync()
        assert np.array_equal(d.cpu().numpy().view(np.uint64).reshape(-1, f.limbs), f.to_mont_array(N.coset_fft(f, a)))
    # hash-to-G1 (both hashers, CIP22, ragged inputs) and the bls-snark-sys signature entry points on top of it
    import ctypes
    from oracle import hash_to_curve as HC
    G1L = C.LAYOUTS["bls12_377_g1"]
    inputs = [(b"", b""), (b"m" * 40, b"x"), (bytes(range(200)), b"extra data")]
    for hasher, oh, cip22 in ((E.HASHER_DIRECT, HC.DIRECT, False), (E.HASHER_COMPOSITE, HC.COMPOSITE, False),
                              (E.HASHER_COMPOSITE, HC.COMPOSITE, True)):
        images, att = E.hash_to_g1(hasher, b"ULforxof", inputs, cip22=cip22)
        for (m, e), img, a in zip(inputs, images, att):
            pt, c = HC.try_and_increment(O.G1, oh, b"ULforxof", m, e, compat=True, cip22=cip22)
            assert G1L.jacobian_compressed(img) == O.serialize_compressed(O.G1, pt) and a == c
    lib = E.load()
    sk, msg = 0xC0FFEE1234567, b"sanitizer"
    pk, sg, ok = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_bool(False)
    hpt, _ = HC.try_and_increment(O.G1, HC.COMPOSITE, b"ULforxof", msg, b"", compat=True, cip22=True)
    pkb, sgb = O.serialize_compressed(O.G2, O.G2.pmul(O.G2_GEN, sk)), O.serialize_compressed(O.G1, O.G1.pmul(hpt, sk))
    assert lib.deserialize_public_key(pkb, 96, ctypes.byref(pk)) and lib.deserialize_signature(sgb, 48, ctypes.byref(sg))
    assert lib.verify_signature(pk, msg, len(msg), b"", 0, sg, True, True, ctypes.byref(ok)) and ok.value
    ptr, n = ctypes.c_void_p(), ctypes.c_int()
    assert lib.serialize_signature(sg, ctypes.byref(ptr), ctypes.byref(n)) and ctypes.string_at(ptr, n.value) == sgb
    assert lib.free_vec(ptr, n.value) and lib.destroy_public_key(pk) and lib.destroy_signature(sg)
    # the SNARK verifier entry point on the reference's known-answer instance: point decoding, the warp-cooperative subgroup
    # checks, g_ic, the BW6-761 Miller loops and the warp-cooperative final exponentiation
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    from bw6_kat import GOLD
    vk_b, proof_b = bytes.fromhex(GOLD["bw6_groth16_vk"]["hex"]), bytes.fromhex(GOLD["bw6_groth16_proof"]["hex"])
    first = E.EpochBlock(0, 0, bytes([1] * 16), bytes([2] * 16), 1, 4, bytes.fromhex(GOLD["bls12_377_first_pubkeys"]["hex"]))
    last = E.EpochBlock(2, 0, bytes([3] * 16), bytes([2] * 16), 1, 4, bytes.fromhex(GOLD["bls12_377_last_pubkeys"]["hex"]))
    assert E.verify_epochs(vk_b, proof_b, first, last)
    print("sanitize smoke ok; launches =", E.launch_count())
    E.shutdown()


if __name__ == "__main__":
    main()
