"""BASELINE config 5 -- epoch-snark Groth16 prove, ARITHMETIC PART (witness map + 4 MSMs + assembly)
through b200_groth16_prove_device, on a synthetic witness of the estimated shape of the outer
BW6-761 circuit for 150 validators x 1 epoch (~10^7 constraints, domain
2^24; a/b/l assignments ~49 % zeros, ~49 % ones, 2 % dense; h dense).  Constraint synthesis (serial
Rust, crates/epoch-snark/src/gadgets) is NOT included and cannot run here; this is labelled
"prover arithmetic".  The proving key is synthetic (k_i * G, one base array shared by the queries:
cost does not depend on the values) and resident, as a prover holds it.
    PYTHONPATH=. python tools/bench_groth16_prove.py [--log-n 24] [--family bw6_761|bls12_377] [--reps 2]"""
import argparse
import json

import numpy as np
import torch

from celo_bls_snark_rs_b200 import engine as E
from tools.bench_sweep import generator_bytes, scalars


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=22)
    ap.add_argument("--family", default="bw6_761")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    outer = args.family == "bw6_761"
    fam = E.GROTH16_BW6_761 if outer else E.GROTH16_BLS12_377
    g1, g2 = (E.BW6_761_G1, E.BW6_761_G2) if outer else (E.BLS12_377_G1, E.BLS12_377_G2)
    limbs, top = (6, 56) if outer else (4, 60)
    n = 1 << args.log_n
    num_assign = int(n * 0.6)                          # ~10^7 variables on a 2^24 domain
    num_aux = num_assign - 2
    E.init(0)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream

    def synth_points(cid, count, seed):
        gen = torch.from_numpy(np.frombuffer(generator_bytes(cid), dtype=np.uint8).copy()).to(dev)
        ks = torch.from_numpy(scalars(count, limbs, top, seed).view(np.int64)).to(dev)
        out = torch.empty((count, E.PACKED_STRIDE[cid]), dtype=torch.uint8, device=dev)
        E.fixed_base_mul_device(cid, gen.data_ptr(), ks.data_ptr(), count, out.data_ptr(), sp)
        torch.cuda.synchronize()
        return out

    b1 = synth_points(g1, n, 3)                        # a_query / l_query / h_query share one base array
    b2 = synth_points(g2, num_assign + 1, 4)
    pk = E.Groth16Pk(b1.data_ptr(), b2.data_ptr(), b1.data_ptr(), b1.data_ptr(), b1.data_ptr(), b2.data_ptr())
    rng = np.random.default_rng(5)
    assign = np.zeros((num_assign, limbs), dtype=np.uint64)
    kind = rng.integers(0, 100, size=num_assign)
    assign[(kind >= 49) & (kind < 98), 0] = 1
    dense = kind >= 98
    assign[dense] = scalars(int(dense.sum()), limbs, top, 6)
    d_assign = torch.from_numpy(assign.view(np.int64)).to(dev)
    ev = scalars(n, limbs, top - 1, 7)                 # arbitrary residues < p as evaluation vectors
    src = [torch.from_numpy(ev.view(np.int64).copy()).to(dev) for _ in range(3)]
    work = [torch.empty_like(s) for s in src]
    proof = torch.zeros(3 * 288, dtype=torch.uint8, device=dev)
    times = []
    for rep in range(args.reps + 1):
        for w, s in zip(work, src):
            w.copy_(s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        E.groth16_prove_device(fam, pk, d_assign.data_ptr(), num_assign, num_aux, work[0].data_ptr(), work[1].data_ptr(),
                               work[2].data_ptr(), args.log_n, proof.data_ptr(), sp)
        e1.record(stream)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    print(json.dumps({"workload": f"Groth16 prover arithmetic ({args.family}): witness map 2^{args.log_n} + MSMs a/l/h (G1), b (G2) + assembly; "
                                  "synthetic witness, no constraint synthesis", "log2_domain": args.log_n,
                      "num_assign": num_assign, "first_ms": round(times[0], 2), "ms": round(min(times[1:]), 2), "n_gpus": 1}))


if __name__ == "__main__":
    main()
