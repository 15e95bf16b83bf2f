"""Timing of the radix-2 NTT and of the Groth16 witness-map transform chain (not the headline bench).
    PYTHONPATH=. python tools/bench_ntt.py [--logs=16,20,24]
One JSON line per (field, log2 n): forward transform ms, witness map ms (7 transforms + element-wise
passes), elements/s, and achieved GB/s counting the ALGORITHMIC traffic of one transform as one read +
one write of the vector (2 * n * element bytes)."""
import json
import sys

import numpy as np
import torch

from celo_bls_snark_rs_b200 import engine as E


def main():
    logs = [16, 20, 22, 24]
    for a in sys.argv:
        if a.startswith("--logs="):
            logs = [int(v) for v in a.split("=", 1)[1].split(",")]
    E.init(0)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    for fid, name, limbs, bits in ((E.FR_BW6_761, "fr_bw6_761", 6, 376), (E.FR_BLS12_377, "fr_bls12_377", 4, 252)):
        for lg in logs:
            n = 1 << lg
            rng = np.random.default_rng(lg)
            raw = rng.integers(0, 1 << 62, size=(n, limbs), dtype=np.uint64)
            raw[:, -1] &= np.uint64((1 << (bits - 64 * (limbs - 1))) - 1)
            bufs = [torch.from_numpy(raw.view(np.int64).copy()).to(dev) for _ in range(4)]
            for _ in range(2):
                E.ntt_device(fid, bufs[0].data_ptr(), lg, False, False, sp)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            reps = 5
            ev[0].record(stream)
            for _ in range(reps):
                E.ntt_device(fid, bufs[0].data_ptr(), lg, False, False, sp)
            ev[1].record(stream)
            for _ in range(reps):
                E.witness_map_device(fid, bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), lg, bufs[3].data_ptr(), sp)
            ev[2].record(stream)
            torch.cuda.synchronize()
            fft_ms, wm_ms = ev[0].elapsed_time(ev[1]) / reps, ev[1].elapsed_time(ev[2]) / reps
            print(json.dumps({"field": name, "log2n": lg, "fft_ms": round(fft_ms, 3), "witness_map_ms": round(wm_ms, 3),
                              "fft_Melem_s": round(n / fft_ms / 1e3, 1),
                              "fft_algorithmic_GBs": round(2 * n * limbs * 8 / fft_ms / 1e6, 1)}), flush=True)
            del bufs


if __name__ == "__main__":
    main()
