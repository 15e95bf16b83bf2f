"""MSM part of the epoch-snark Groth16 prover (BASELINE config 5):
the four VariableBaseMSM calls inside ark-groth16's create_proof_no_zk for the outer BW6-761
circuit (crates/epoch-snark/src/api/prover.rs:78), on a synthetic witness of the estimated shape
(150 validators x 1 epoch: ~10^7 constraints, domain 2^24):
    a_query  (G1)  n = 10^7   witness-density scalars (~49% 0, ~49% 1, 2% dense)
    b_g2     (G2)  n = 10^7   same scalars
    l_query  (G1)  n = 10^7   same scalars
    h_query  (G1)  n = 2^24-1 dense 377-bit scalars
Constraint synthesis (Rust) and the 7 NTTs are NOT included; this is labelled "prover MSM
arithmetic".  Proving-key bases are resident and prepared once (b200_pack_bases_device), as a
prover would hold them.  --scale k divides all sizes by 2^k for quick runs.
    PYTHONPATH=. python tools/bench_groth16_msms.py [--scale 2]"""
import json
import sys

import numpy as np
import torch

from celo_bls_snark_rs_b200 import engine as E
from tools.bench_sweep import generator_bytes, scalars


def witness_scalars(n, seed):
    rng = np.random.default_rng(seed)
    sc = np.zeros((n, 6), dtype=np.uint64)
    kind = rng.integers(0, 100, size=n)
    sc[(kind >= 49) & (kind < 98), 0] = 1
    dense = kind >= 98
    sc[dense] = scalars(int(dense.sum()), 6, 56, seed + 1)
    return sc


def main():
    scale = int(sys.argv[sys.argv.index("--scale") + 1]) if "--scale" in sys.argv else 0
    E.init(0)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    n_w = 10_000_000 >> scale
    n_h = ((1 << 24) - 1) >> scale
    cid = E.BW6_761_G1
    n_max = max(n_w, n_h)
    gen = torch.from_numpy(np.frombuffer(generator_bytes(cid), dtype=np.uint8).copy()).to(dev)
    ks = torch.from_numpy(scalars(n_max, 6, 56, 11).view(np.int64)).to(dev)
    bases = torch.empty((n_max, 192), dtype=torch.uint8, device=dev)        # one base set reused for all queries
    E.fixed_base_mul_device(cid, gen.data_ptr(), ks.data_ptr(), n_max, bases.data_ptr(), sp)
    torch.cuda.synchronize()
    del ks
    d_w = torch.from_numpy(witness_scalars(n_w, 5).view(np.int64)).to(dev)
    d_h = torch.from_numpy(scalars(n_h, 6, 56, 7).view(np.int64)).to(dev)
    out = torch.empty(4 * 288, dtype=torch.uint8, device=dev)
    jobs = [("a_query", E.BW6_761_G1, d_w, n_w), ("b_g2_query", E.BW6_761_G2, d_w, n_w),
            ("l_query", E.BW6_761_G1, d_w, n_w), ("h_query", E.BW6_761_G1, d_h, n_h)]
    res = {}
    for rep in range(2):                                                   # first pass warms the workspace
        total = 0.0
        for i, (name, c, sc, n) in enumerate(jobs):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            E.msm_device(c, bases.data_ptr(), sc.data_ptr(), n, out.data_ptr() + 288 * i, sp)
            e1.record(stream)
            torch.cuda.synchronize()
            res[name] = round(e0.elapsed_time(e1), 2)
            total += e0.elapsed_time(e1)
    print(json.dumps({"workload": "epoch-snark outer Groth16 prover, MSM arithmetic only (no synthesis, no NTT)",
                      "curve": "bw6_761", "witness_n": n_w, "h_n": n_h, "ms": res, "total_ms": round(total, 2),
                      "n_gpus": 1}))


if __name__ == "__main__":
    main()
