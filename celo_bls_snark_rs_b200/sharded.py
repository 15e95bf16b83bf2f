"""Chunk-sharded MSM across the GPUs of one box.

An MSM is a sum, so the (base, scalar) arrays are split into `world` contiguous chunks, every
rank reduces its chunk to one partial Jacobian point with the local bucket MSM, and the single
exchange step is an all-gather of those partials (144 or 288 bytes each) followed by a local
sum on every rank.  NCCL cannot reduce elliptic-curve points, hence all-gather + add rather
than all-reduce.  torch.distributed provides the plumbing; the arithmetic stays in the CUDA
library (engine.sum_jacobian_device).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist

from . import engine as E


def _current_stream(device) -> int:
    """torch's current stream on `device`: NCCL collectives are issued there, so the engine calls around them must be
    too (stream 0 would select the engine's private non-blocking stream, unordered with NCCL)."""
    if torch.device(device).type != "cuda":
        return 0
    return torch.cuda.current_stream(device).cuda_stream or 1      # 1 = cudaStreamLegacy: torch's default stream, by handle


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous chunk [lo, hi) of rank `rank`; the first n % world ranks get one extra pair."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_partials(partial: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather of one partial result per rank: uint8 [jac_bytes] -> uint8 [world * jac_bytes],
    rank-major.  Works on any backend (NCCL on the GPUs, gloo in the CPU tests)."""
    world = dist.get_world_size(group)
    out = torch.empty(world * partial.numel(), dtype=partial.dtype, device=partial.device)
    dist.all_gather_into_tensor(out, partial.contiguous(), group=group)
    return out


class ShardedMsm:
    """Per-rank state of a sharded MSM: the local chunk stays resident in HBM."""

    def __init__(self, curve: int, device: torch.device, group=None):
        self.curve, self.device, self.group = curve, device, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        jb = E.JAC_BYTES[curve]
        self.partial = torch.zeros(jb, dtype=torch.uint8, device=device)
        self.gathered = torch.zeros(self.world * jb, dtype=torch.uint8, device=device)
        self.result = torch.zeros(jb, dtype=torch.uint8, device=device)

    def run(self, d_bases: torch.Tensor, d_scalars: torch.Tensor, n_local: int, stream: int = 0) -> torch.Tensor:
        """Local MSM over this rank's chunk, exchange, combine.  Asynchronous on `stream`
        (which must be torch's current stream so NCCL orders after the MSM)."""
        stream = stream or _current_stream(self.device)
        E.msm_device(self.curve, d_bases.data_ptr(), d_scalars.data_ptr(), n_local, self.partial.data_ptr(), stream)
        if self.world == 1:
            return self.partial
        dist.all_gather_into_tensor(self.gathered, self.partial, group=self.group)
        E.sum_jacobian_device(self.curve, self.gathered.data_ptr(), self.world, self.result.data_ptr(), stream)
        return self.result

    def run_batch(self, jobs, stream: int = 0) -> torch.Tensor:
        """`jobs` = [(d_bases, d_scalars, n_local), ...]: independent sharded MSMs.  The local MSMs go
        through the engine's pipelined batch entry (sort / accumulate / tail of consecutive MSMs
        overlap), then every MSM gets its own all-gather + sum, as in run().  Returns uint8
        [len(jobs), jac_bytes] results (the partials themselves when world == 1)."""
        stream = stream or _current_stream(self.device)
        jb = E.JAC_BYTES[self.curve]
        k = len(jobs)
        if getattr(self, "_batch_k", 0) < k:
            self._partials = torch.zeros((k, jb), dtype=torch.uint8, device=self.device)
            self._gathered = torch.zeros((k, self.world * jb), dtype=torch.uint8, device=self.device)
            self._results = torch.zeros((k, jb), dtype=torch.uint8, device=self.device)
            self._batch_k = k
        E.msm_batch_device(self.curve, [(b.data_ptr(), s.data_ptr(), n, self._partials[i].data_ptr())
                                        for i, (b, s, n) in enumerate(jobs)], stream)
        if self.world == 1:
            return self._partials[:k]
        # ONE all-gather for the whole batch (rank-major records of k partial points each) and one combine launch
        # (block b adds the `world` partials of MSM b); round 1 issued k gathers and k sums after the batch
        flat = self._gathered.view(-1)[:self.world * k * jb]
        dist.all_gather_into_tensor(flat, self._partials[:k].reshape(-1), group=self.group)
        E.sum_jacobian_batch_device(self.curve, flat.data_ptr(), self.world, k, self._results.data_ptr(), stream)
        return self._results[:k]


class ShardedPairing:
    """Chunk-sharded product of pairings: every rank runs the Miller loops of
    its chunk of pairs and multiplies them into ONE Fq12 Miller value (576 bytes); the single exchange
    is an all-gather of those values; every rank then multiplies the `world` values and runs the one
    final exponentiation.  A product of Miller values is exact, so the GT bytes equal the single-GPU
    result (crates/bls-crypto/src/bls/signature.rs:149 semantics, pairs in any order)."""

    def __init__(self, device: torch.device, group=None):
        self.device, self.group = device, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.partial = torch.zeros(E.FQ12_BYTES, dtype=torch.uint8, device=device)
        self.gathered = torch.zeros(self.world * E.FQ12_BYTES, dtype=torch.uint8, device=device)
        self.gt = torch.zeros(E.FQ12_BYTES, dtype=torch.uint8, device=device)
        self.is_one = torch.zeros(1, dtype=torch.int32, device=device)

    def run(self, d_g1: torch.Tensor, d_g2: torch.Tensor, n_local: int, stream: int = 0):
        """d_g1 / d_g2: this rank's packed affine records (96 / 192 bytes each).  Returns (gt, is_one)
        device tensors, asynchronous on `stream` (torch's current stream, so NCCL orders after it)."""
        stream = stream or _current_stream(self.device)
        E.miller_product_device(d_g1.data_ptr() if n_local else 0, d_g2.data_ptr() if n_local else 0, n_local,
                                self.partial.data_ptr(), stream)
        vals = self.partial
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.partial, group=self.group)
            vals = self.gathered
        E.final_exp_device(vals.data_ptr(), self.world, self.gt.data_ptr(), self.is_one.data_ptr(), stream)
        return self.gt, self.is_one


def gather_ragged(local: torch.Tensor, n: int, record_bytes: int, group=None) -> torch.Tensor:
    """All-gather of per-rank slices of `n` fixed-size records split by shard_bounds (slice lengths differ by at
    most one record): every rank pads its slice to the longest one, the padded blocks are all-gathered rank-major and
    the padding is dropped -> uint8 [n * record_bytes] in record order on every rank."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    longest = -(-n // world) if n else 0
    lo, hi = shard_bounds(n, world, rank)
    assert local.numel() == (hi - lo) * record_bytes
    block = torch.zeros(max(longest, 1) * record_bytes, dtype=torch.uint8, device=local.device)
    block[:local.numel()] = local
    out = torch.empty(world * block.numel(), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out, block, group=group)
    parts = []
    for r in range(world):
        rlo, rhi = shard_bounds(n, world, r)
        parts.append(out[r * block.numel(): r * block.numel() + (rhi - rlo) * record_bytes])
    return torch.cat(parts) if parts else out[:0]


class ShardedHashToG1:
    """Message hashing of a batch split over the ranks (independent units, no data-path
    collective inside the hot loop): rank r hashes messages shard_bounds(n, world, r) with b200_hash_to_g1 and the
    144-byte results are all-gathered so that every rank holds all n hash points, in message order, for its share
    of the pairing product (ShardedPairing).  `hash_fn(hasher, domain, inputs, compat, cip22) -> (images, attempts)`
    defaults to the CUDA engine; the gloo CPU test passes the oracle in its place."""

    def __init__(self, hasher: int, cip22: bool = False, compat: bool = True, device=None, group=None, hash_fn=None):
        self.hasher, self.cip22, self.compat, self.group = hasher, cip22, compat, group
        self.device = device if device is not None else torch.device("cpu")
        self.hash_fn = hash_fn if hash_fn is not None else E.hash_to_g1

    def run(self, domain: bytes, inputs):
        n = len(inputs)
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        lo, hi = shard_bounds(n, world, rank)
        images, _ = self.hash_fn(self.hasher, domain, list(inputs[lo:hi]), compat=self.compat, cip22=self.cip22)
        local = torch.frombuffer(bytearray(b"".join(images)), dtype=torch.uint8).to(self.device) if images else \
            torch.zeros(0, dtype=torch.uint8, device=self.device)
        flat = gather_ragged(local, n, 144, self.group).cpu().numpy().tobytes()
        return [flat[144 * i:144 * (i + 1)] for i in range(n)]


class ShardedGroth16:
    """Groth16 prover arithmetic split over the ranks (BASELINE config 5: "8 x B200"; crates/epoch-snark/src/api/prover.rs:78,112).
    Every rank holds the whole proving key and witness (as a prover node does), runs the witness map and its contiguous
    share of the four MSMs (b200_groth16_prove_partial_device), the ONE exchange is an all-gather of the 4-point partial
    records, and every rank assembles A | B | C from them (b200_groth16_assemble_device)."""

    def __init__(self, family: int, device: torch.device, group=None):
        self.family, self.device, self.group = family, device, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        pb = E.groth16_partial_bytes(family)
        self.partial = torch.zeros(pb, dtype=torch.uint8, device=device)
        self.gathered = torch.zeros(self.world * pb, dtype=torch.uint8, device=device)
        self.proof = torch.zeros(2 * (144 if family == E.GROTH16_BLS12_377 else 288) + 288, dtype=torch.uint8, device=device)

    def run(self, pk, d_assign: torch.Tensor, num_assign: int, num_aux: int, d_a: torch.Tensor, d_b: torch.Tensor,
            d_c: torch.Tensor, log_n: int, stream: int = 0) -> torch.Tensor:
        """Asynchronous on `stream` (torch's current stream, so NCCL orders after the local share)."""
        stream = stream or torch.cuda.current_stream().cuda_stream
        E.groth16_prove_partial_device(self.family, pk, d_assign.data_ptr(), num_assign, num_aux, d_a.data_ptr(), d_b.data_ptr(),
                                       d_c.data_ptr(), log_n, self.rank, self.world, self.partial.data_ptr(), stream)
        vals = self.partial
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.partial, group=self.group)
            vals = self.gathered
        E.groth16_assemble_device(self.family, pk, vals.data_ptr(), self.world, self.proof.data_ptr(), stream)
        return self.proof
