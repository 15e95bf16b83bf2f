"""BASELINE config 2: batch aggregate BLS verify, 4096 signatures -> a 4097-pair BLS12-377
multi-pairing (Signature::batch_verify_hashes, crates/bls-crypto/src/bls/signature.rs:125-155).
Not the headline bench (bench.py); prints one JSON line.
    PYTHONPATH=. python tools/bench_pairing.py [--pairs 4097] [--steps 10]
Pairs are (a_i * G1, b_i * G2) generated on the device (the Miller loop's cost does not depend on
the values); `device_ms` = Miller loop + product tree + final exponentiation with packed inputs
resident (CUDA events), `e2e_ms` = b200_multi_pairing_bls12_377 on pinned arkworks-layout host
records (H2D + pack + compute + D2H of the flag)."""
import argparse
import ctypes
import json
import time

import numpy as np
import torch

from celo_bls_snark_rs_b200 import engine as E
from tools.bench_sweep import generator_bytes, scalars


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4097)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    n = args.pairs
    E.init(0)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    pts = {}
    for cid in (E.BLS12_377_G1, E.BLS12_377_G2):
        gen = torch.from_numpy(np.frombuffer(generator_bytes(cid), dtype=np.uint8).copy()).to(dev)
        ks = torch.from_numpy(scalars(n, 4, 60, 11 + cid).view(np.int64)).to(dev)
        out = torch.empty((n, E.PACKED_STRIDE[cid]), dtype=torch.uint8, device=dev)
        E.fixed_base_mul_device(cid, gen.data_ptr(), ks.data_ptr(), n, out.data_ptr(), sp)
        pts[cid] = out
    torch.cuda.synchronize()
    g1, g2 = pts[E.BLS12_377_G1], pts[E.BLS12_377_G2]
    mill = torch.zeros(576, dtype=torch.uint8, device=dev)
    gt = torch.zeros(576, dtype=torch.uint8, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)

    def step():
        E.miller_product_device(g1.data_ptr(), g2.data_ptr(), n, mill.data_ptr(), sp)
        E.final_exp_device(mill.data_ptr(), 1, gt.data_ptr(), flag.data_ptr(), sp)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(stream)
    for _ in range(args.steps):
        E.miller_product_device(g1.data_ptr(), g2.data_ptr(), n, mill.data_ptr(), sp)
    e1.record(stream)
    for _ in range(args.steps):
        E.final_exp_device(mill.data_ptr(), 1, gt.data_ptr(), flag.data_ptr(), sp)
    e2.record(stream)
    torch.cuda.synchronize()
    miller_ms, fe_ms = e0.elapsed_time(e1) / args.steps, e1.elapsed_time(e2) / args.steps

    # end to end: arkworks-layout host records (104 B / 200 B), pinned
    h1 = torch.zeros((n, 104), dtype=torch.uint8).pin_memory()
    h2 = torch.zeros((n, 200), dtype=torch.uint8).pin_memory()
    h1[:, :96].copy_(g1.cpu())
    h2[:, :192].copy_(g2.cpu())
    lib = E.load()
    f = ctypes.c_int(0)
    for _ in range(2):
        E._check(lib.b200_multi_pairing_bls12_377(h1.data_ptr(), 104, h2.data_ptr(), 200, n, None, ctypes.byref(f)))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        E._check(lib.b200_multi_pairing_bls12_377(h1.data_ptr(), 104, h2.data_ptr(), 200, n, None, ctypes.byref(f)))
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    dev_ms = miller_ms + fe_ms
    print(json.dumps({"workload": f"BLS12-377 multi-pairing, {n} pairs (batch_verify_hashes of {n - 1} signatures)",
                      "miller_ms": round(miller_ms, 3), "final_exp_ms": round(fe_ms, 3), "device_ms": round(dev_ms, 3),
                      "pairings_per_s": round(n / dev_ms * 1e3), "e2e_ms": round(e2e_ms, 3),
                      "e2e_pairings_per_s": round(n / e2e_ms * 1e3), "h2d_bytes": n * 304,
                      "algorithmic_bytes": n * 288 + 576}))


if __name__ == "__main__":
    main()
