"""N > 1 host logic on CPU: world_size-2 gloo run of the chunk partition + all-gather of partial
results.  The per-rank MSM and the final point sum are done by the CPU oracle here (there is no
GPU); what is under test is the partitioning and the exchange layout used by ShardedMsm."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from celo_bls_snark_rs_b200 import sharded
from oracle import cref as C
from oracle import inputs as H
from oracle import oracle as O


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 8, 1000, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            spans = [sharded.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharded.shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        name = "bls12_377_g1"
        L = C.LAYOUTS[name]
        pts, scalars = H.edge_case_inputs(name, n, 5150)          # same seeded inputs on every rank
        lo, hi = sharded.shard_bounds(n, world, rank)
        part = C.msm(L, L.affine_records(pts[lo:hi]), L.scalars_array(scalars[lo:hi]), threads=1)
        t = torch.from_numpy(np.frombuffer(part, dtype=np.uint8).copy())
        gathered = sharded.gather_partials(t)
        assert gathered.numel() == world * L.jac_bytes
        raw = gathered.numpy().tobytes()
        assert raw[rank * L.jac_bytes:(rank + 1) * L.jac_bytes] == part    # rank-major layout
        total = None
        for r in range(world):
            total = L.curve.padd(total, L.jacobian_to_affine(raw[r * L.jac_bytes:(r + 1) * L.jac_bytes]))
        q.put((rank, O.serialize_compressed(L.curve, total)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_msm_matches_single_msm():
    n, world = 101, 2
    name = "bls12_377_g1"
    L = C.LAYOUTS[name]
    pts, scalars = H.edge_case_inputs(name, n, 5150)
    want = O.serialize_compressed(L.curve, L.curve.msm_naive(pts, scalars))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == want and got[1] == want


def _pairing_worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g1, g2 = H.signature_batch(n - 1, 31)                    # (sigma, -g2), (H_i, pk_i): product == 1
        lo, hi = sharded.shard_bounds(n, world, rank)
        part = O.miller_loop(list(zip(g1[lo:hi], g2[lo:hi])))    # this rank's Miller value (oracle stands in for the GPU)
        t = torch.from_numpy(np.frombuffer(C.fq12_to_ark_bytes(part), dtype=np.uint8).copy())
        gathered = sharded.gather_partials(t)
        raw = gathered.numpy().tobytes()
        assert len(raw) == world * 576 and raw[rank * 576:(rank + 1) * 576] == t.numpy().tobytes()
        prod = O.FQ12_ONE
        for r in range(world):
            prod = O.f12_mul(prod, C.fq12_from_ark_bytes(raw[r * 576:(r + 1) * 576]))
        q.put((rank, O.final_exponentiation(prod) == O.FQ12_ONE))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_pairing_product_is_exact():
    """ShardedPairing's exchange on CPU: Miller values of the two chunks, all-gathered and multiplied,
    give the same verdict as the single product (a valid 5-signature batch -> one)."""
    n, world = 6, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pairing_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == {0: True, 1: True}


def _oracle_hash_fn(hasher, domain, inputs, compat=True, cip22=False):
    """Stands in for engine.hash_to_g1 on CPU: the oracle's hash points as G1Projective images (z = 1)."""
    from oracle import hash_to_curve as HC
    L = C.LAYOUTS["bls12_377_g1"]
    images, attempts = [], []
    for m, e in inputs:
        pt, c = HC.try_and_increment(O.G1, HC.DIRECT, domain, m, e, compat=compat, cip22=cip22)
        images.append(L.fe_to_mont_bytes(pt[0]) + L.fe_to_mont_bytes(pt[1]) + L.fe_to_mont_bytes(1))
        attempts.append(c)
    return images, attempts


def _hash_worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        inputs = [(bytes([i, 7 * i % 251]) * (1 + i % 5), bytes([i]) * (i % 3)) for i in range(n)]
        job = sharded.ShardedHashToG1(0, hash_fn=_oracle_hash_fn)
        q.put((rank, job.run(b"ULforxof", inputs)))
    finally:
        dist.destroy_process_group()


def test_gather_ragged_layout_single_process():
    # world = 1 degenerates to a copy; the ragged arithmetic itself is covered by the two-rank run below
    for n, world in ((0, 2), (1, 2), (5, 2), (7, 3), (8, 8)):
        spans = [sharded.shard_bounds(n, world, r) for r in range(world)]
        assert sum(hi - lo for lo, hi in spans) == n and max(hi - lo for lo, hi in spans) == (-(-n // world) if n else 0)


def test_two_rank_gloo_sharded_hash_to_g1_is_in_message_order():
    """ShardedHashToG1 on CPU: 7 messages over 2 ranks (slices of 4 and 3, so the gather is ragged); every rank ends
    with all 7 hash points in message order, equal to hashing the batch in one piece."""
    n, world = 7, 2
    inputs = [(bytes([i, 7 * i % 251]) * (1 + i % 5), bytes([i]) * (i % 3)) for i in range(n)]
    want, _ = _oracle_hash_fn(0, b"ULforxof", inputs)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_hash_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == want and got[1] == want
