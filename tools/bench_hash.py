"""Batched hash-to-G1 (b200_hash_to_g1): the messages of a 4096-signature batch
(BASELINE config 2) hashed in one launch, timed through the host-pointer C-ABI (host messages in, 144-byte
G1Projective images out), next to the Python oracle on a small sample.  Not the headline bench; one JSON line.
    PYTHONPATH=. python tools/bench_hash.py [--n 4096] [--steps 5]"""
import argparse
import json
import time

from celo_bls_snark_rs_b200 import engine as E
from oracle import cref as C
from oracle import hash_to_curve as H
from oracle import oracle as O


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--no-oracle", action="store_true", help="skip timing the Python oracle sample")
    args = ap.parse_args()
    E.init(0)
    rng = O.SplitMix64(5)
    res = {"tool": "bench_hash", "n": args.n, "steps": args.steps, "cases": []}
    t0 = time.perf_counter()
    E.hash_crh(E.HASHER_COMPOSITE, b"", [b"x"])             # builds the CRH table (first use)
    res["crh_table_setup_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
    for msg_len in (32, 1024):
        inputs = [(bytes(rng.below(256) for _ in range(msg_len)), b"\x01\x02") for _ in range(args.n)]
        for name, hasher, oh, cip22 in (("direct", E.HASHER_DIRECT, H.DIRECT, False),
                                        ("composite", E.HASHER_COMPOSITE, H.COMPOSITE, False),
                                        ("composite_cip22", E.HASHER_COMPOSITE, H.COMPOSITE, True)):
            batch = E.HashBatch(inputs)                 # ctypes marshalling of the Python byte strings: not timed
            flags = E.HASH_COMPAT | (E.HASH_CIP22 if cip22 else 0)
            batch.run(hasher, b"ULforxof", flags)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                batch.run(hasher, b"ULforxof", flags)    # the C-ABI call: host messages in, points + counters out
            ms = (time.perf_counter() - t0) * 1e3 / args.steps
            att = batch.results()[1]
            sample = [] if args.no_oracle else inputs[:4]
            t0 = time.perf_counter()
            for m, e in sample:
                H.try_and_increment(O.G1, oh, b"ULforxof", m, e, compat=True, cip22=cip22)
            cpu_ms = (time.perf_counter() - t0) * 1e3 / max(len(sample), 1)
            # C port of the same hasher (oracle/cpu_ref.c), one host thread, on a bounded sample
            sample_c = inputs[:256] if name == "direct" else inputs[:32 if msg_len <= 64 else 8]
            t0 = time.perf_counter()
            for m, e in sample_c:
                if name == "direct":
                    C.hash_to_g1_direct(b"ULforxof", m, e, True)
                else:
                    C.hash_to_g1_composite(b"ULforxof", m, e, True, cip22)
            c_ms = round((time.perf_counter() - t0) * 1e3 / len(sample_c), 4)
            res["cases"].append({"hasher": name, "message_bytes": msg_len, "e2e_ms": round(ms, 3), "cpu_c_port_ms_per_hash_1_thread": c_ms,
                                 "hashes_per_s": round(args.n / ms * 1e3), "max_attempt": max(att),
                                 "python_oracle_ms_per_hash": None if args.no_oracle else round(cpu_ms, 2)})
    import os
    res["host_cores"] = os.cpu_count()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
