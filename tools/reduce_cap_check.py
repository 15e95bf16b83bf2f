"""Bucket reduce with one thread per segment: the register-capped form against the uncapped one, inside one process
(B200_REDUCE_THREAD_CAP is read at every call).  Per (curve, n): ms per MSM for both forms (CUDA events, 3 warm-ups, 5 timed)
and whether the two results are the same point (affine bytes after b200_batch_to_affine_device).
    PYTHONPATH=. python tools/reduce_cap_check.py [bw6_761_g1:22 bls12_377_g1:24 ...]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from celo_bls_snark_rs_b200 import engine as E  # noqa: E402
from bench_sweep import generator_bytes, scalars  # noqa: E402

NAMES = {"bls12_377_g1": E.BLS12_377_G1, "bls12_377_g2": E.BLS12_377_G2, "bw6_761_g1": E.BW6_761_G1}


def main():
    cases = [a for a in sys.argv[1:] if ":" in a] or ["bw6_761_g1:20", "bw6_761_g1:22", "bls12_377_g1:22", "bls12_377_g1:24"]
    E.init(0)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    for case in cases:
        name, lg = case.split(":")
        cid, n = NAMES[name], 1 << int(lg)
        limbs = E.SCALAR_BYTES[cid] // 8
        top = 60 if limbs == 4 else 56
        gen = torch.from_numpy(np.frombuffer(generator_bytes(cid), dtype=np.uint8).copy()).to(dev)
        ks = torch.from_numpy(scalars(n, limbs, top, 1).view(np.int64)).to(dev)
        sc = torch.from_numpy(scalars(n, limbs, top, 2).view(np.int64)).to(dev)
        bases = torch.empty((n, E.PACKED_STRIDE[cid]), dtype=torch.uint8, device=dev)
        E.fixed_base_mul_device(cid, gen.data_ptr(), ks.data_ptr(), n, bases.data_ptr(), sp)
        res = {"curve": name, "log2n": int(lg)}
        points = []
        for cap in ("0", "1"):
            os.environ["B200_REDUCE_THREAD_CAP"] = cap
            out = torch.zeros(E.JAC_BYTES[cid], dtype=torch.uint8, device=dev)
            for _ in range(3):
                E.msm_device(cid, bases.data_ptr(), sc.data_ptr(), n, out.data_ptr(), sp)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5 if int(lg) <= 22 else 3
            e0.record(stream)
            for _ in range(reps):
                E.msm_device(cid, bases.data_ptr(), sc.data_ptr(), n, out.data_ptr(), sp)
            e1.record(stream)
            torch.cuda.synchronize()
            res["ms_cap" + cap] = round(e0.elapsed_time(e1) / reps, 3)
            aff = torch.zeros(E.PACKED_STRIDE[cid], dtype=torch.uint8, device=dev)
            E.batch_to_affine_device(cid, out.data_ptr(), 1, aff.data_ptr(), sp)
            torch.cuda.synchronize()
            points.append(aff.cpu().numpy().tobytes())
        res["same_point"] = points[0] == points[1] and any(points[0])
        print(json.dumps(res), flush=True)
        del bases, ks, sc
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
