"""BASELINE config 4: one MSM of 2^log2n pairs, input-chunk sharded across the ranks of one box
(strong scaling): local bucket MSM on n / world pairs -> all-gather of the partial Jacobian points ->
local sum on every rank (celo_bls_snark_rs_b200/sharded.py).  Launch with torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sharded.py \
        [--curve bw6_761_g1] [--log2n 22] [--steps 5]
Rank 0 prints one JSON line (CUDA-event time, max over ranks)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from celo_bls_snark_rs_b200 import engine as E                      # noqa: E402
from celo_bls_snark_rs_b200.sharded import ShardedMsm, shard_bounds   # noqa: E402
from tools.bench_sweep import generator_bytes, scalars                # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--curve", default="bw6_761_g1")
    ap.add_argument("--log2n", type=int, default=22)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    E.init(local)
    cid = E.CURVE_IDS[args.curve]
    limbs = E.SCALAR_BYTES[cid] // 8
    top = 60 if limbs == 4 else 56
    lo, hi = shard_bounds(1 << args.log2n, world, rank)
    n = hi - lo
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    gen = torch.from_numpy(np.frombuffer(generator_bytes(cid), dtype=np.uint8).copy()).to(dev)
    ks = torch.from_numpy(scalars(n, limbs, top, 100 + rank).view(np.int64)).to(dev)
    bases = torch.empty((n, E.PACKED_STRIDE[cid]), dtype=torch.uint8, device=dev)
    E.fixed_base_mul_device(cid, gen.data_ptr(), ks.data_ptr(), n, bases.data_ptr(), sp)
    sc = torch.from_numpy(scalars(n, limbs, top, 200 + rank).view(np.int64)).to(dev)
    job = ShardedMsm(cid, dev)
    for _ in range(2):
        job.run(bases, sc, n, sp)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        job.run(bases, sc, n, sp)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    if rank == 0:
        print(json.dumps({"workload": f"{args.curve} MSM, n=2^{args.log2n} total, chunk-sharded x{world} (strong scaling)",
                          "n_gpus": world, "pairs_per_gpu": n, "ms_per_msm": round(ms, 3),
                          "Mpairs_s": round((1 << args.log2n) / ms / 1e3, 2)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
