"""One process, all GPUs of the box, behind the C-ABI: b200_msm_sharded / b200_multi_pairing_bls12_377_sharded against their
single-GPU twins on the same pinned host arrays (the path a Rust caller of Signature::batch takes, signature.rs:70-89).
    PYTHONPATH=. python tools/bench_multi_device.py [--log2n 22] [--curve bls12_377_g1] [--pairs 4097]
Prints one JSON line: ms per call and Mpairs/s at 1 and at N GPUs, results checked equal (canonical compressed bytes)."""
import argparse
import json
import time

import numpy as np
import torch

from celo_bls_snark_rs_b200 import engine as E
from oracle import cref as C
from tools.bench_sweep import generator_bytes, scalars


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) * 1e3 / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=22)
    ap.add_argument("--curve", default="bls12_377_g1")
    ap.add_argument("--pairs", type=int, default=4097)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    ngpu = torch.cuda.device_count()
    L = C.LAYOUTS[args.curve]
    cid, n = L.id, 1 << args.log2n
    E.init(0)
    dev = torch.device("cuda:0")
    limbs = E.SCALAR_BYTES[cid] // 8
    top = 60 if limbs == 4 else 56
    gen = torch.from_numpy(np.frombuffer(generator_bytes(cid), dtype=np.uint8).copy()).to(dev)
    st = np.zeros((n // 32, limbs), dtype=np.uint64)
    st[:, 0] = np.random.default_rng(1).integers(1 << 40, 1 << 62, size=n // 32, dtype=np.uint64)
    d_st = torch.from_numpy(st.view(np.int64)).to(dev)
    d_bases = torch.empty((n, E.PACKED_STRIDE[cid]), dtype=torch.uint8, device=dev)
    E.point_runs_device(cid, gen.data_ptr(), d_st.data_ptr(), n // 32, 32, d_bases.data_ptr())
    E.sync()
    h_bases = torch.zeros((n, E.ARK_STRIDE[cid]), dtype=torch.uint8).pin_memory()
    h_bases[:, :E.PACKED_STRIDE[cid]].copy_(d_bases.cpu())
    h_sc = torch.from_numpy(scalars(n, limbs, top, 2).view(np.int64)).pin_memory()
    del d_bases
    out1, outn = np.zeros(E.JAC_BYTES[cid], dtype=np.uint8), np.zeros(E.JAC_BYTES[cid], dtype=np.uint8)
    lib = E.load()
    ms1, _ = timed(lambda: E._check(lib.b200_msm(cid, h_bases.data_ptr(), E.ARK_STRIDE[cid], h_sc.data_ptr(), n, out1.ctypes.data)), args.reps)
    res = {"tool": "bench_multi_device", "gpus": ngpu, "curve": args.curve, "log2n": args.log2n,
           "msm_1gpu_ms": ms1, "msm_1gpu_Mpairs_s": n / ms1 / 1e3}
    if ngpu > 1:
        E.init_devices(list(range(ngpu)))
        msn, _ = timed(lambda: E._check(lib.b200_msm_sharded(cid, h_bases.data_ptr(), E.ARK_STRIDE[cid], h_sc.data_ptr(), n, outn.ctypes.data)), args.reps)
        res.update({"msm_sharded_ms": msn, "msm_sharded_Mpairs_s": n / msn / 1e3, "speedup": ms1 / msn,
                    "parity": L.jacobian_compressed(out1.tobytes()) == L.jacobian_compressed(outn.tobytes())})
        # pairs: config 2's shape
        from oracle import inputs as H
        L1, L2 = C.LAYOUTS["bls12_377_g1"], C.LAYOUTS["bls12_377_g2"]
        g1, g2 = H.signature_batch(args.pairs - 1, 5)
        r1, r2 = L1.affine_records(g1), L2.affine_records(g2)
        p1, (ok1, gt1) = timed(lambda: E.multi_pairing(r1, r2), args.reps)
        pn, (okn, gtn) = timed(lambda: E.multi_pairing_sharded(r1, r2), args.reps)
        res.update({"pairing_pairs": args.pairs, "pairing_1gpu_ms": p1, "pairing_sharded_ms": pn,
                    "pairing_parity": bool(ok1 and okn and gt1 == gtn)})
    print(json.dumps(res))


if __name__ == "__main__":
    main()
