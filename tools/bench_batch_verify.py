"""BASELINE config 2 end to end: a batch of n = 4096 (public key, message, signature) triples through bls-snark-sys' own
entry point `batch_verify_signature` as exported by the CUDA library (crates/bls-snark-sys/src/signatures.rs:290-333):
signature aggregation, hash-to-G1 of all n raw messages, the (n + 1)-pair product of pairings and the final
exponentiation -- host buffers in, one bool out.  Not the headline bench; prints one JSON line.
    PYTHONPATH=. python tools/bench_batch_verify.py [--n 4096] [--steps 5]
The triples are valid by construction without n host signatures: every key is pk = sk * g2 for one sk, message i is
random, signature 0 is sk * sum_i H(m_i) (hashes and sum from the device, one oracle-free MSM of size 1) and the other
signatures are the identity -- the aggregate is the correct one and no step's cost depends on the values."""
import argparse
import ctypes
import json
import time

import numpy as np
import torch

from celo_bls_snark_rs_b200 import bls, engine as E
from tools.bench_sweep import generator_bytes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    n = args.n
    E.init(0)
    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    lib = E.load()
    rng = np.random.default_rng(7)
    msgs = [(rng.integers(0, 256, 32, dtype=np.uint8).tobytes(), b"\x01\x02") for _ in range(n)]
    sk = np.array([[0x1F2E3D4C5B6A7988, 0x1122334455667788, 0x99AABBCCDDEEFF00, 0x0123456789ABCDEF >> 8]], dtype=np.uint64)
    out = {"tool": "bench_batch_verify", "n": n, "steps": args.steps, "cases": []}
    # pk = sk * g2 (one record, reused by every message)
    g2 = torch.from_numpy(np.frombuffer(generator_bytes(E.BLS12_377_G2), dtype=np.uint8).copy()).to(dev)
    d_sk = torch.from_numpy(sk.view(np.int64)).to(dev)
    pk_aff = torch.empty(192, dtype=torch.uint8, device=dev)
    E.fixed_base_mul_device(E.BLS12_377_G2, g2.data_ptr(), d_sk.data_ptr(), 1, pk_aff.data_ptr())
    E.sync()
    one = bytes.fromhex("68ffffffffffcd02b1ffff7f839f4051f23f7d8aa9b37d9f05637c6eb7974e7be8843c80bf95f44c9af4fde261668d00")
    pk_img = pk_aff.cpu().numpy().tobytes() + one + bytes(48)
    zero_sig = one + one + bytes(48)                                   # G1Projective::zero() = (1, 1, 0)
    for name, hasher, composite, cip22 in (("direct", bls.DIRECT_HASH_TO_G1, False, False),
                                           ("composite_cip22", bls.COMPOSITE_HASH_TO_G1_CIP22, True, True)):
        hashes = hasher.hash_many(bls.SIG_DOMAIN, msgs)
        hsum = bls.Signature.aggregate([bls.Signature(h) for h in hashes])
        sig0 = bls.Signature.batch([int(sum(int(sk[0, j]) << (64 * j) for j in range(4)))], [hsum]).image
        pk_buf = ctypes.create_string_buffer(pk_img, 288)
        sig_bufs = [ctypes.create_string_buffer(sig0, 144), ctypes.create_string_buffer(zero_sig, 144)]
        arr = (E.MessageFFI * n)(*[E.MessageFFI(E.FFIBuffer(m, len(m)), E.FFIBuffer(e, len(e)), ctypes.addressof(pk_buf),
                                                ctypes.addressof(sig_bufs[0 if i == 0 else 1])) for i, (m, e) in enumerate(msgs)])
        ok = ctypes.c_bool(False)
        assert lib.batch_verify_signature(arr, n, composite, cip22, ctypes.byref(ok)) and ok.value, name
        times = []
        for _ in range(args.steps):                                  # per-call wall clock; the median is reported (host
            t0 = time.perf_counter()                                 # allocations and pageable copies make single calls noisy)
            lib.batch_verify_signature(arr, n, composite, cip22, ctypes.byref(ok))
            times.append((time.perf_counter() - t0) * 1e3)
        ms, best = float(np.median(times)), min(times)
        assert ok.value
        arr[n // 2].data = E.FFIBuffer(b"tampered", 8)                # one wrong message -> false
        assert lib.batch_verify_signature(arr, n, composite, cip22, ctypes.byref(ok)) and not ok.value
        out["cases"].append({"hasher": name, "e2e_ms": round(ms, 3), "best_ms": round(best, 3), "max_ms": round(max(times), 3),
                             "signatures_per_s": round(n / ms * 1e3)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
