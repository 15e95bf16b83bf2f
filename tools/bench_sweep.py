"""Device-resident MSM timing sweep over curves and sizes (not the headline bench: bench.py).
Bases are k_i * G generated on the device; BW6-761 uses a seeded curve point as G.
    PYTHONPATH=. python tools/bench_sweep.py [--quick] [--curve=bw6_761_g1] [--logs=12,20]
Prints one JSON line per (curve, n): ms per MSM (CUDA events, 3 warm-ups, 5 timed), Mpairs/s."""
import json
import sys

import numpy as np
import torch

from celo_bls_snark_rs_b200 import engine as E

P377 = 0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001
Q761 = 0x0122E824FB83CE0AD187C94004FAFF3EB926186A81D14688528275EF8087BE41707BA638E584E91903CEBAFF25B423048689C8ED12F9FD9071DCD3DC73EBFF2E98A116C25667A8F8160CF8AEEAF0A437E6913E6870000082F49D00000000008B
G1 = (0x008848DEFE740A67C8FC6225BF87FF5485951E2CAA9D41BB188282C8BD37CB5CD5481512FFCD394EEAB9B16EB21BE9EF,
      0x01914A69C5102EFF1F674F5D30AFEEC4BD7FB348CA3E52D96D182AD44FB82305C2FE3D3634A9591AFD82DE55559C8EA6)
G2X = (0x018480BE71C785FEC89630A2A3841D01C565F071203E50317EA501F557DB6B9B71889F52BB53540274E3E48F7C005196,
       0x00EA6040E700403170DC5A51B1B140D5532777EE6651CECBE7223ECE0799C9DE5CF89984BFF76FE6B26BFEFA6EA16AFE)


def mont(v, p, nbytes):
    return ((v << (8 * nbytes)) % p).to_bytes(nbytes, "little")


def generator_bytes(curve):
    if curve == E.BLS12_377_G1:
        return mont(G1[0], P377, 48) + mont(G1[1], P377, 48)
    if curve == E.BLS12_377_G2:
        from oracle import oracle as O            # tools/ may use the oracle for input construction only
        (x0, x1), (y0, y1) = O.G2_GEN
        return b"".join(mont(v, P377, 48) for v in (x0, x1, y0, y1))
    b = -1 if curve == E.BW6_761_G1 else 4
    x = 5
    while True:                                   # first x >= 5 on the curve (q = 3 mod 4)
        rhs = (x * x * x + b) % Q761
        y = pow(rhs, (Q761 + 1) // 4, Q761)
        if y * y % Q761 == rhs:
            return mont(x, Q761, 96) + mont(y, Q761, 96)
        x += 1


def scalars(n, limbs, top_bits, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 63, size=(n, limbs), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, limbs), dtype=np.uint64)
    a[:, -1] &= np.uint64((1 << top_bits) - 1)
    return a


def main():
    quick = "--quick" in sys.argv
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--curve=")]
    logs_override = [[int(v) for v in a.split("=", 1)[1].split(",")] for a in sys.argv if a.startswith("--logs=")]
    E.init(0)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    plan = [(E.BLS12_377_G1, "bls12_377_g1", [12, 16, 20, 22] if quick else [10, 12, 14, 16, 18, 20, 22, 24]),
            (E.BLS12_377_G2, "bls12_377_g2", [12, 16, 20] if quick else [10, 12, 14, 16, 18, 20, 22]),
            (E.BW6_761_G1, "bw6_761_g1", [12, 16, 20] if quick else [12, 14, 16, 18, 20, 22])]
    for cid, name, logs in plan:
        if only and name not in only:
            continue
        if logs_override:
            logs = logs_override[0]
        limbs = E.SCALAR_BYTES[cid] // 8
        top = 60 if limbs == 4 else 56
        gen = torch.from_numpy(np.frombuffer(generator_bytes(cid), dtype=np.uint8).copy()).to(dev)
        for lg in logs:
            n = 1 << lg
            ks = torch.from_numpy(scalars(n, limbs, top, 1).view(np.int64)).to(dev)
            sc = torch.from_numpy(scalars(n, limbs, top, 2).view(np.int64)).to(dev)
            bases = torch.empty((n, E.PACKED_STRIDE[cid]), dtype=torch.uint8, device=dev)
            E.fixed_base_mul_device(cid, gen.data_ptr(), ks.data_ptr(), n, bases.data_ptr(), sp)
            out = torch.empty(E.JAC_BYTES[cid], dtype=torch.uint8, device=dev)
            for _ in range(3):
                E.msm_device(cid, bases.data_ptr(), sc.data_ptr(), n, out.data_ptr(), sp)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5 if lg <= 22 else 2
            e0.record(stream)
            for _ in range(reps):
                E.msm_device(cid, bases.data_ptr(), sc.data_ptr(), n, out.data_ptr(), sp)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            c, w, nb = E.msm_plan(cid, n)
            print(json.dumps({"curve": name, "log2n": lg, "ms": round(ms, 3), "Mpairs_s": round(n / ms / 1e3, 2),
                              "window_bits": c, "windows": w}), flush=True)
            del bases, ks, sc


if __name__ == "__main__":
    main()
