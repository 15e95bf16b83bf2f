"""K MSMs issued one by one (b200_msm_device) against the same K through b200_msm_batch_device
(software-pipelined in the engine).  PYTHONPATH=. python tools/bench_batch.py [--log2n 20] [--k 16]"""
import argparse
import json

import numpy as np
import torch

from celo_bls_snark_rs_b200 import engine as E
from tools.bench_sweep import generator_bytes, scalars


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=20)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--curve", default="bls12_377_g1")
    args = ap.parse_args()
    cid = E.CURVE_IDS[args.curve]
    limbs = E.SCALAR_BYTES[cid] // 8
    top = 60 if limbs == 4 else 56
    n = 1 << args.log2n
    E.init(0)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    gen = torch.from_numpy(np.frombuffer(generator_bytes(cid), dtype=np.uint8).copy()).to(dev)
    sets = []
    for s in range(2):
        ks = torch.from_numpy(scalars(n, limbs, top, 10 + s).view(np.int64)).to(dev)
        bases = torch.empty((n, E.PACKED_STRIDE[cid]), dtype=torch.uint8, device=dev)
        E.fixed_base_mul_device(cid, gen.data_ptr(), ks.data_ptr(), n, bases.data_ptr(), sp)
        sc = torch.from_numpy(scalars(n, limbs, top, 20 + s).view(np.int64)).to(dev)
        sets.append((bases, sc))
    out = torch.zeros((args.k, E.JAC_BYTES[cid]), dtype=torch.uint8, device=dev)
    jobs = [(sets[i & 1][0].data_ptr(), sets[i & 1][1].data_ptr(), n, out[i].data_ptr()) for i in range(args.k)]
    torch.cuda.synchronize()

    def seq():
        for b, s, nn, o in jobs:
            E.msm_device(cid, b, s, nn, o, sp)

    def bat():
        E.msm_batch_device(cid, jobs, sp)

    res = {}
    for name, fn in (("sequential", seq), ("batch", bat)):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        res[name + "_ms_per_msm"] = round(e0.elapsed_time(e1) / args.k, 3)
    res["Mpairs_s_batch"] = round(n / res["batch_ms_per_msm"] / 1e3, 1)
    print(json.dumps({"curve": args.curve, "log2n": args.log2n, "k": args.k, **res}))


if __name__ == "__main__":
    main()
