"""Latency of the BW6-761 Groth16 verification on one B200: the reference's own
known-answer instance (crates/bls-snark-sys/src/snark/mod.rs:52-119) through b200_groth16_verify_bw6_761 (host
pointers, wall clock around the synchronous call) and its two kernels timed with CUDA events on the launching stream.

    python tools/bench_bw6_verify.py [reps]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from bw6_kat import PROOF, VK, kat_inputs  # noqa: E402  (test fixture: the reference's KAT instance)
from celo_bls_snark_rs_b200 import engine as E  # noqa: E402
from oracle import cref as C  # noqa: E402  (input preparation only)


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    E.init(0)
    L1, L2 = C.LAYOUTS["bw6_761_g1"], C.LAYOUTS["bw6_761_g2"]
    inputs = L1.scalars_array(kat_inputs())
    r1, r2 = (lambda p: L1.affine_records(p)), (lambda p: L2.affine_records(p))
    args = (r1([VK["alpha"]]), r2([VK["beta"]]), r2([VK["gamma"]]), r2([VK["delta"]]), r1(VK["gamma_abc"]),
            r1([PROOF[0]]), r2([PROOF[1]]), r1([PROOF[2]]), inputs)
    assert E.groth16_verify_bw6(*args) is True
    wall = []
    for _ in range(reps):
        t0 = time.perf_counter()
        ok = E.groth16_verify_bw6(*args)
        wall.append((time.perf_counter() - t0) * 1e3)
        assert ok
    # kernels alone: 4 pairs (the shape of one verification), device-resident, on torch's current stream
    g1 = torch.from_numpy(L1.affine_records([PROOF[0], VK["gamma_abc"][0], PROOF[2], VK["alpha"]], 192)).cuda()
    g2 = torch.from_numpy(L2.affine_records([PROOF[1], VK["gamma"], VK["delta"], VK["beta"]], 192)).cuda()
    vals = torch.zeros(8 * 576, dtype=torch.uint8, device="cuda")
    out = torch.zeros(576, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()                      # a real stream: handle 0 would select the engine's own
    st = stream.cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    abc = torch.from_numpy(L1.affine_records(VK["gamma_abc"], 192)).cuda()
    scal = torch.from_numpy(L1.scalars_array([1] + kat_inputs())).cuda()
    jac = torch.zeros(288, dtype=torch.uint8, device="cuda")
    aff = torch.zeros(192, dtype=torch.uint8, device="cuda")
    mil, fin, gic, toa = [], [], [], []
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        for _ in range(reps + 2):
            ev[0].record()
            E.msm_device(E.BW6_761_G1, abc.data_ptr(), scal.data_ptr(), 3, jac.data_ptr(), st)
            ev[1].record()
            E.batch_to_affine_device(E.BW6_761_G1, jac.data_ptr(), 1, aff.data_ptr(), st)
            ev[2].record()
            E.miller_values_bw6_device(g1.data_ptr(), g2.data_ptr(), 4, vals.data_ptr(), st)
            ev[3].record()
            E.final_exp_bw6_device(vals.data_ptr(), 8, out.data_ptr(), 0, st)
            ev[4].record()
            stream.synchronize()
            gic.append(ev[0].elapsed_time(ev[1]))
            toa.append(ev[1].elapsed_time(ev[2]))
            mil.append(ev[2].elapsed_time(ev[3]))
            fin.append(ev[3].elapsed_time(ev[4]))
    mil, fin, gic, toa = sorted(mil[2:]), sorted(fin[2:]), sorted(gic[2:]), sorted(toa[2:])
    # the reference's entry point on raw bytes (decoding + subgroup checks + hashing + the above)
    from bw6_kat import GOLD
    vk_b, proof_b = bytes.fromhex(GOLD["bw6_groth16_vk"]["hex"]), bytes.fromhex(GOLD["bw6_groth16_proof"]["hex"])
    first = E.EpochBlock(0, 0, bytes([1] * 16), bytes([2] * 16), 1, 4, bytes.fromhex(GOLD["bls12_377_first_pubkeys"]["hex"]))
    last = E.EpochBlock(2, 0, bytes([3] * 16), bytes([2] * 16), 1, 4, bytes.fromhex(GOLD["bls12_377_last_pubkeys"]["hex"]))
    assert E.verify_epochs(vk_b, proof_b, first, last)
    full = []
    for _ in range(reps):
        t0 = time.perf_counter()
        assert E.verify_epochs(vk_b, proof_b, first, last)
        full.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps({"workload": "BW6-761 Groth16 verify_proof, 2 public inputs (reference KAT instance)",
                      "verify_call_ms_median": sorted(wall)[len(wall) // 2], "verify_call_ms_min": min(wall),
                      "g_ic_msm_ms": gic[len(gic) // 2], "g_ic_to_affine_ms": toa[len(toa) // 2],
                      "miller_4_pairs_ms": mil[len(mil) // 2], "final_exp_ms": fin[len(fin) // 2],
                      "verify_entry_point_ms_median": sorted(full)[len(full) // 2], "reps": reps}))


if __name__ == "__main__":
    main()
