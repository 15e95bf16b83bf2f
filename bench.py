#!/usr/bin/env python
"""Headline benchmark: BLS12-377 G1 MSM, n = 2^20 pairs per GPU (BASELINE.json metric
"G1 MSM Mscalar-muls/s (BLS12-377, n=2^20) at 1/2/4/8 GPU"), plus the other BASELINE configs as
sub-results of the same JSON line.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: arkworks-algorithm C port

One "step" = one MSM over one batch of synthetic (base, scalar) pairs.  At N GPUs every rank owns a
contiguous chunk of 2^20 pairs of one N*2^20-pair MSM (weak scaling): local bucket MSM -> NCCL
all-gather of the 144-byte partial Jacobian points -> local sum kernel (NCCL cannot reduce
elliptic-curve points).  The K timed steps are K complete, independent MSMs
queued through b200_msm_batch_device, which software-pipelines consecutive MSMs inside the engine;
`sequential_ms_per_step` is the same MSM issued one call at a time (b200_msm_device).

Printed JSON (rank 0, one line):
  value     pairs/s (in Mpairs/s) with inputs already resident in HBM, CUDA-event timed,
            max over ranks, barrier + synchronize on both sides
  e2e       the same metric through the host-pointer C-ABI call b200_msm(): pinned host
            buffers in arkworks layout -> H2D -> pack -> MSM -> D2H of the result
            (`e2e.pageable`: the same call on ordinary malloc'd buffers, what a Rust Vec is)
  parity    the result of the LAST timed MSM against the C port of arkworks' algorithm on the
            SAME inputs (canonical compressed bytes); at N > 1 every rank checks its partial and
            rank 0 checks the gathered-and-summed point against the sum of the CPU partials.
            The process exits non-zero when any check fails.
  roofline  dominant kernel (k_bucket_accumulate) vs the measured HBM peak, and `int_pipe_frac`:
            algorithmic 32x32->64 multiply-adds / (32 per clock per SM): the bound that applies
  cpu_baseline  the C oracle (arkworks-algorithm port) on the box's host cores, same inputs
  configs   BASELINE configs 2, 4, 5 and the strong-scaling reading of config 3, each with ms,
            value, parity, roofline, cpu_baseline (DESIGN.md section 5 defines every figure)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LOG2_N = 20
CURVE = "bls12_377_g1"
BYTES_PER_PAIR = 128            # 32 B scalar + 2 x 48 B affine coordinates
SEED = 0x5DBE62598D313D76       # first 8 bytes of the reference's test seed (hash_to_curve/mod.rs:290-293)
METRIC = "G1 MSM Mscalar-muls/s (BLS12-377, n=2^20)"

P377 = 0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001
# wide (32x32->64) multiply-adds of one XYZZ mixed addition (8 M + 2 S; a 12-limb product is 288, its squaring 222;
# 24 limbs: 1152 / 876; an Fq2 product is three Fq products, an Fq2 squaring two) -- the integer-pipe roof counts these
MADS_PER_MADD = {"bls12_377_g1": 8 * 288 + 2 * 222, "bls12_377_g2": (8 * 3 + 2 * 2) * 288, "bw6_761_g1": 8 * 1152 + 2 * 876}
PIPE_MADS_PER_CLK_SM = 32       # measured (profiles/r1_field_layer_notes.md): one IMAD.WIDE warp instruction per 4 cycles per SMSP
SM_COUNT = 148


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def rand_scalars(n: int, limbs: int, top_bits: int, seed: int) -> np.ndarray:
    """n uniform scalars of 64 * (limbs - 1) + top_bits bits as uint64 [n, limbs]; numpy PCG64."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 63, size=(n, limbs), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, limbs), dtype=np.uint64)
    a[:, -1] &= np.uint64((1 << top_bits) - 1)
    return a


def uniform_scalars(n: int, seed: int) -> np.ndarray:
    """n uniform 252-bit scalars (< r of BLS12-377) as uint64 [n, 4]."""
    return rand_scalars(n, 4, 60, seed)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self, wait_s: float = 8.0):
        """Start of the timed region: waits for nvidia-smi's first row (its start-up can outlast a short timed
        region on a cold box) and remembers where the region's samples begin."""
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < wait_s:
            time.sleep(0.02)
        self.first = len(self.rows)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = self.rows[getattr(self, "first", 0):] or self.rows      # samples taken during the timed region
        num = lambda s: s.replace(".", "", 1).isdigit()
        sm = [float(r[0]) for r in rows if r and num(r[0])]
        mx = [float(r[1]) for r in rows if len(r) > 1 and num(r[1])]
        pw = [float(r[6]) for r in rows if len(r) > 6 and num(r[6])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(pw) if pw else None}


# ----------------------------------------------------------------------------------------
# CPU arm: the arkworks-algorithm C port (oracle/cpu_ref.c), window-parallel threads
# ----------------------------------------------------------------------------------------
def cpu_inputs(n: int):
    """Bases for the CPU arm: 1024 distinct oracle points tiled to n (an MSM's cost does not
    depend on the base values); scalars uniform."""
    from oracle import cref as C
    from oracle import oracle as O
    L = C.LAYOUTS[CURVE]
    rng = O.SplitMix64(SEED)
    packed = C.fixed_base_batch(L, O.G1_GEN, [rng.below(O.R) for _ in range(1024)])
    bases = C.with_flags(L, np.tile(packed, (max(1, n // 1024), 1))[:n])
    return L, bases, uniform_scalars(n, SEED & 0xFFFFFFFF)


def cpu_msm_rate(n: int, reps: int, threads: int):
    from oracle import cref as C
    L, bases, sc = cpu_inputs(n)
    C.msm(L, bases[:1024], sc[:1024], threads=threads)           # warm the library
    t0 = time.perf_counter()
    for _ in range(reps):
        C.msm(L, bases, sc, threads=threads)
    dt = (time.perf_counter() - t0) / reps
    return n / dt, dt, C.msm_window_tasks(L, n)


def pick_cpu_sample(threads: int, budget_s: float = 6.0):
    """Largest n = 2^k <= 2^20 whose single MSM fits the time budget on this host."""
    rate, _, _ = cpu_msm_rate(1 << 13, 1, threads)
    k = LOG2_N
    while k > 13 and (1 << k) / (rate * 1.6) > budget_s:          # larger n runs ~1.6x faster per pair
        k -= 1
    return k


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    k = pick_cpu_sample(cores)
    n = 1 << k
    from oracle import cref as C
    L, bases, sc = cpu_inputs(n)
    tasks = C.msm_window_tasks(L, n)
    threads = min(cores, tasks)
    for _ in range(args.warmup):
        C.msm(L, bases, sc, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        C.msm(L, bases, sc, threads=threads)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt / 1e6
    sample = f"n=2^{k} pairs per step (full workload is 2^{LOG2_N}), {args.steps} steps"
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mpairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (377-bit Montgomery)", "data": "synthetic",
        "config": {"workload": f"BLS12-377 G1 Pippenger MSM on host cores, arkworks window rule, n=2^{k}",
                   "note": "arkworks-algorithm C port (oracle/cpu_ref.c), not the arkworks binary: no Rust toolchain"},
        "cpu_baseline": {"value": val, "unit": "Mpairs/s", "cores": threads, "kind": "port", "sample": sample,
                         "host_cores": cores, "window_tasks": tasks},
        "e2e": {"value": val, "unit": "Mpairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ----------------------------------------------------------------------------------------
# CUDA arm
# ----------------------------------------------------------------------------------------
class Ctx:
    """Per-rank benchmark context shared by the headline and the sub-results."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from celo_bls_snark_rs_b200 import engine as E
        self.torch, self.dist, self.E, self.args = torch, dist, E, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        E.init(self.local)
        self.stream = torch.cuda.Stream(device=self.dev)   # an explicit stream: kernels, NCCL and the timing events share it
        torch.cuda.set_stream(self.stream)
        self.sp = self.stream.cuda_stream
        assert self.sp != 0
        self.cores = os.cpu_count() or 1
        self.cpu_threads = max(1, self.cores // self.world)  # every rank runs its own CPU check
        self.clock_hz = 1.965e9

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return list(vals)
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def all_true(self, flag: bool) -> bool:
        if self.world == 1:
            return bool(flag)
        t = self.torch.tensor([1 if flag else 0], dtype=self.torch.int32, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    def gather_objects(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def timed(self, fn, steps, warmup=1):
        """ms per step of `fn` on the bench stream (CUDA events, barrier + synchronize both sides, max over ranks)."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            fn()
        e1.record(self.stream)
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1) / steps)[0]

    def up(self, arr):
        return self.torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy()).to(self.dev)

    def pipe_frac(self, mads: float, ms: float, gpus: int = 1) -> float:
        return mads / (PIPE_MADS_PER_CLK_SM * SM_COUNT * self.clock_hz * gpus * ms / 1e3) if ms else 0.0


def generator_record(cid) -> bytes:
    from tools.bench_sweep import generator_bytes
    return generator_bytes(cid)


def gen_points(ctx: Ctx, cid: int, lo: int, hi: int, total: int, seed: int, run: int = 32):
    """Packed affine records [lo, hi) of a synthetic array of `total` distinct points: runs of `run` consecutive multiples
    (s_t + j) * G with seeded 2^60+-bit starts s_t -- every rank derives the same global array and keeps its slice."""
    E, torch = ctx.E, ctx.torch
    limbs = E.SCALAR_BYTES[cid] // 8
    assert lo % run == 0 and (hi % run == 0 or hi == total) and total % run == 0
    starts = rand_scalars(total // run, limbs, 40, seed)
    starts[:, 1:] = 0                                   # 64-bit starts: a short double-and-add per run
    starts[:, 0] |= np.uint64(1 << 62)
    sl = starts[lo // run:(hi + run - 1) // run]
    d_gen = ctx.up(np.frombuffer(generator_record(cid), dtype=np.uint8))
    d_st = ctx.up(sl)
    out = torch.empty((len(sl) * run, E.PACKED_STRIDE[cid]), dtype=torch.uint8, device=ctx.dev)
    E.point_runs_device(cid, d_gen.data_ptr(), d_st.data_ptr(), len(sl), run, out.data_ptr(), ctx.sp)
    torch.cuda.synchronize()
    return out[:hi - lo]


def cpu_point_sum(L, jacobians):
    """sum of GroupProjective images with the Python oracle's affine group law -> canonical compressed bytes"""
    from oracle import oracle as O
    acc = None
    for raw in jacobians:
        acc = L.curve.padd(acc, L.jacobian_to_affine(raw))
    return O.serialize_compressed(L.curve, acc)


def check_sharded_msm(ctx: Ctx, name: str, d_bases, sc_np, n_local: int, gpu_partial: bytes, gpu_result: bytes, sample: int = 0):
    """CUDA against the C port on the same inputs.  sample = 0: this rank's whole chunk (its GPU partial against the CPU
    MSM of the chunk); sample > 0: a separate GPU MSM over the chunk's first `sample` pairs against the CPU one (CPU too
    slow for the chunk).  Always: rank 0 recomputes the combine -- the sum of every rank's partial -- on the CPU and
    compares it with the point the GPUs agreed on.  Returns (ok, detail)."""
    from oracle import cref as C
    E, torch = ctx.E, ctx.torch
    L = C.LAYOUTS[name]
    cid = L.id
    detail = {}
    if sample and sample < n_local:
        m = sample
        out = torch.zeros(E.JAC_BYTES[cid], dtype=torch.uint8, device=ctx.dev)
        d_s = ctx.up(sc_np[:m])
        E.msm_device(cid, d_bases.data_ptr(), d_s.data_ptr(), m, out.data_ptr(), ctx.sp)
        torch.cuda.synchronize()
        del d_s
        got = out.cpu().numpy().tobytes()
        detail["checked_pairs_per_rank"] = m
    else:
        m, got = n_local, gpu_partial
        detail["checked_pairs_per_rank"] = n_local
    bases_h = d_bases[:m].cpu().numpy()
    t0 = time.perf_counter()
    want = C.msm(L, bases_h, sc_np[:m], threads=ctx.cpu_threads)
    detail["cpu_check_s"] = round(time.perf_counter() - t0, 2)
    ok_local = L.jacobian_compressed(got) == L.jacobian_compressed(want)
    ok = ctx.all_true(ok_local)
    parts = ctx.gather_objects(gpu_partial)
    if ctx.rank == 0:
        combine_ok = cpu_point_sum(L, parts) == L.jacobian_compressed(gpu_result)
        detail["combine_checked"] = True
        ok = ok and combine_ok
    return ctx.all_true(ok), detail


def msm_roofline(ctx: Ctx, name: str, n_per_launch: float, kernel_ms: float, bytes_per_pair: int, windows: int, traffic=None):
    peak, peak_src = measured_peak_gbs()
    achieved = n_per_launch * bytes_per_pair / (kernel_ms / 1e3) / 1e9 if kernel_ms else 0.0
    return {"bound": "hbm", "kernel": "k_bucket_accumulate", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src, "kernel_ms": kernel_ms,
            "algorithmic_bytes_per_launch": n_per_launch * bytes_per_pair,
            "int_pipe_frac": ctx.pipe_frac(windows * n_per_launch * MADS_PER_MADD[name], kernel_ms),
            "int_pipe_note": f"{windows} windows x n mixed additions x {MADS_PER_MADD[name]} wide multiply-adds against "
                             f"{PIPE_MADS_PER_CLK_SM}/clk/SM x {SM_COUNT} SMs x {ctx.clock_hz / 1e9:.3f} GHz: the roof that binds"}


def cpu_msm_baseline(ctx: Ctx, name: str, bases_h, sc_np, budget_s: float = 5.0):
    """C port on a bounded prefix of the same inputs; threads = min(cores, windows) (arkworks parallelises over windows only)."""
    from oracle import cref as C
    L = C.LAYOUTS[name]
    k = 12
    C.msm(L, bases_h[:256], sc_np[:256], threads=1)
    t0 = time.perf_counter()
    C.msm(L, bases_h[:1 << k], sc_np[:1 << k], threads=min(ctx.cores, C.msm_window_tasks(L, 1 << k)))
    rate = (1 << k) / (time.perf_counter() - t0)
    while (1 << (k + 1)) <= len(bases_h) and (1 << (k + 1)) / (rate * 1.3) < budget_s:
        k += 1
    n = 1 << k
    tasks = C.msm_window_tasks(L, n)
    threads = min(ctx.cores, tasks)
    t0 = time.perf_counter()
    C.msm(L, bases_h[:n], sc_np[:n], threads=threads)
    dt = time.perf_counter() - t0
    return {"value": n / dt / 1e6, "unit": "Mpairs/s", "cores": threads, "kind": "port", "host_cores": ctx.cores,
            "sample": f"one MSM over the first 2^{k} pairs of the same inputs ({dt:.2f} s), arkworks window rule, "
                      f"{threads} threads over {tasks} windows"}


# ---- BASELINE config 3: the headline ---------------------------------------------------------
def headline(ctx: Ctx):
    torch, dist, E, args = ctx.torch, ctx.dist, ctx.E, ctx.args
    from celo_bls_snark_rs_b200.sharded import ShardedMsm
    world, rank, dev, stream, sp = ctx.world, ctx.rank, ctx.dev, ctx.stream, ctx.sp
    cid = E.BLS12_377_G1
    n = 1 << args.log2n
    # ---- synthetic inputs, resident in HBM: bases = k_i * G on the GPU, two rotating input sets ----
    d_gen = ctx.up(np.frombuffer(generator_record(cid), dtype=np.uint8))
    sets = []
    for s in range(2):
        ks = uniform_scalars(n, (SEED + 1000 * rank + s) & 0xFFFFFFFF)
        sc = uniform_scalars(n, (SEED + 77 + 1000 * rank + s) & 0xFFFFFFFF)
        d_ks = torch.from_numpy(ks.view(np.int64)).to(dev)
        d_bases = torch.empty((n, 96), dtype=torch.uint8, device=dev)
        E.fixed_base_mul_device(cid, d_gen.data_ptr(), d_ks.data_ptr(), n, d_bases.data_ptr(), sp)
        d_sc = torch.from_numpy(sc.view(np.int64)).to(dev)
        sets.append((d_bases, d_sc, sc))
    torch.cuda.synchronize()

    job = ShardedMsm(cid, dev)
    d_part, d_all, d_res = job.partial, job.gathered, job.result
    last = {}

    def steps(k):
        # k complete MSMs (rotating input sets), pipelined inside the engine; for world > 1 each is
        # followed by its all-gather of partials and the local sum
        last["results"] = job.run_batch([(sets[i & 1][0], sets[i & 1][1], n) for i in range(k)], sp)
        last["k"] = k

    sampler = ClockSampler(ctx.local)                   # every rank samples its own GPU (per_rank in the JSON line)
    sampler.start()
    steps(args.warmup)
    ctx.barrier()
    # the same MSM one call at a time (no overlap between consecutive MSMs), for reference
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    seq_steps = min(args.steps, 5)
    s0.record(stream)
    for i in range(seq_steps):
        job.run(sets[i & 1][0], sets[i & 1][1], n, sp)
    s1.record(stream)
    ctx.barrier()
    seq_ms = s0.elapsed_time(s1) / seq_steps
    E.profile_enable(True)
    sampler.mark()
    launches0 = E.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    e0.record(stream)
    steps(args.steps)
    e1.record(stream)
    ctx.barrier()
    ms = e0.elapsed_time(e1)
    launches = E.launch_count() - launches0 + (args.steps if world > 1 else 0)     # + NCCL all-gather kernels
    acc_ms, acc_launches, acc_pairs = E.profile_read()
    E.profile_enable(False)
    clocks = sampler.stop()
    if clocks and clocks.get("sm_mhz"):
        ctx.clock_hz = clocks["sm_mhz"] * 1e6
    # per-rank view of the timed region: where a weak-scaling loss comes from (kernel time per GPU, clocks, power)
    per_rank = ctx.gather_objects({"rank": rank, "ms_per_step": ms / args.steps, "kernel_ms": acc_ms / max(acc_launches, 1),
                                   "sm_mhz": clocks.get("sm_mhz"), "power_w_max": clocks.get("power_w_max"),
                                   "reasons": clocks.get("reasons")})

    # ---- parity of the LAST timed MSM: this rank's partial and the combined point against the C port ----
    k_last = args.steps - 1
    gpu_result = last["results"][k_last].cpu().numpy().tobytes()
    gpu_partial = job._partials[k_last].cpu().numpy().tobytes()
    parity, parity_detail = check_sharded_msm(ctx, CURVE, sets[k_last & 1][0], sets[k_last & 1][2], n, gpu_partial, gpu_result)

    # ---- end to end through the host-pointer C-ABI: pinned and pageable arkworks-layout buffers ----
    h_bases_np = np.zeros((n, 104), dtype=np.uint8)
    h_bases_np[:, :96] = sets[0][0].cpu().numpy()
    h_sc_np = np.ascontiguousarray(sets[0][2])
    h_bases = torch.from_numpy(h_bases_np.copy()).pin_memory()
    h_sc = torch.from_numpy(h_sc_np.view(np.int64).copy()).pin_memory()
    out = np.zeros(144, dtype=np.uint8)

    def e2e_step(bases_ptr, sc_ptr):
        E.msm_host_ptrs(cid, bases_ptr, 104, sc_ptr, n, out)
        if world > 1:
            d_part.copy_(torch.from_numpy(out), non_blocking=False)
            dist.all_gather_into_tensor(d_all, d_part)
            E.sum_jacobian_device(cid, d_all.data_ptr(), world, d_res.data_ptr(), sp)
            return d_res.cpu()
        return out

    def e2e_time(bases_ptr, sc_ptr, steps_):
        for _ in range(2):
            e2e_step(bases_ptr, sc_ptr)
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps_):
            e2e_step(bases_ptr, sc_ptr)
        ctx.barrier()
        return (time.perf_counter() - t0) * 1e3 / steps_

    e2e_ms = e2e_time(h_bases.data_ptr(), h_sc.data_ptr(), args.steps)
    e2e_out = out.copy().tobytes()
    pageable_ms = e2e_time(h_bases_np.ctypes.data, h_sc_np.ctypes.data, min(args.steps, 10))
    from oracle import cref as C
    L = C.LAYOUTS[CURVE]
    # the host-pointer path must land on the same point as the device path did for set 0
    dev0 = torch.zeros(144, dtype=torch.uint8, device=dev)
    E.msm_device(cid, sets[0][0].data_ptr(), sets[0][1].data_ptr(), n, dev0.data_ptr(), sp)
    torch.cuda.synchronize()
    e2e_parity = ctx.all_true(L.jacobian_compressed(e2e_out) == L.jacobian_compressed(dev0.cpu().numpy().tobytes()) and
                              L.jacobian_compressed(out.tobytes()) == L.jacobian_compressed(e2e_out))
    ms, e2e_ms, pageable_ms = ctx.max_over_ranks(ms, e2e_ms, pageable_ms)

    line = None
    if rank == 0:
        total_pairs = n * world * args.steps
        value = total_pairs / (ms / 1e3) / 1e6
        kernel_ms = acc_ms / max(acc_launches, 1)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("k_bucket_accumulate_bytes_per_launch")
            except Exception:
                traffic = None
        c, w, nb = E.msm_plan(cid, n)
        roof = msm_roofline(ctx, CURVE, acc_pairs / max(acc_launches, 1), kernel_ms, BYTES_PER_PAIR, w, traffic)
        roof["kernel_share_of_step"] = kernel_ms / (ms / args.steps) if ms else None
        roof["note"] = ("integer-ALU-bound path (about 160 377-bit Montgomery products per pair): the HBM fraction is small by "
                        "construction, int_pipe_frac is the fraction of the roof that applies; see DESIGN.md.  kernel_ms is "
                        "event-timed on the accumulate stream while the neighbouring MSMs' sort and tail kernels share the GPU")
        line = {
            "metric": METRIC, "value": value, "unit": "Mpairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 limbs (377-bit Montgomery, integer pipe)", "data": "synthetic",
            "config": {"workload": f"BLS12-377 G1 Pippenger MSM, n=2^{args.log2n} pairs per GPU, uniform 252-bit scalars, "
                                   "bases k_i*G generated on device", "pairs_per_gpu": n, "window_bits": c, "windows": w,
                       "buckets_per_window": nb, "parallelism": f"chunk-sharded x{world} + all-gather of 144 B partials",
                       "l2": "two rotating input sets; inputs (128 MiB) + workspace (>160 MiB) exceed the 126 MB L2",
                       "pipeline": "K independent MSMs through b200_msm_batch_device: sort / accumulate / tail of consecutive "
                                   "MSMs overlap on three streams, two workspace sets",
                       "sequential_ms_per_step": seq_ms},
            "parity": bool(parity and e2e_parity),
            "parity_detail": dict(parity_detail, checked="last timed MSM of every rank against oracle/cpu_ref.c on the same inputs "
                                                         "(canonical compressed bytes); e2e result against the device-path result",
                                  e2e_matches_device=bool(e2e_parity)),
            "e2e": {"value": total_pairs / args.steps / (e2e_ms / 1e3) / 1e6, "unit": "Mpairs/s", "h2d_bytes_per_step": n * (104 + 32),
                    "d2h_bytes_per_step": 144, "ms_per_step": e2e_ms,
                    "api": "b200_msm (host pointers, pinned arkworks-layout records)",
                    "pageable": {"value": n * world / (pageable_ms / 1e3) / 1e6, "ms_per_step": pageable_ms,
                                 "note": "the same call on ordinary (pageable) host memory, what a Rust Vec<G1Affine> is"}},
            "gpu_launches": int(launches),
            "roofline": roof,
            "clocks": clocks,
            "per_rank": per_rank,
        }
        if not args.no_cpu:
            line["cpu_baseline"] = cpu_msm_baseline(ctx, CURVE, sets[0][0].cpu().numpy(), sets[0][2], budget_s=6.0)
    del sets
    torch.cuda.empty_cache()
    return line, bool(parity and e2e_parity)


# ---- BASELINE config 3, strong-scaling reading: ONE MSM of 2^20 pairs split over the ranks -----------------
def sub_config3_strong(ctx: Ctx):
    torch, E = ctx.torch, ctx.E
    from celo_bls_snark_rs_b200.sharded import ShardedMsm, shard_bounds
    cid, n = E.BLS12_377_G1, 1 << ctx.args.log2n
    lo, hi = shard_bounds(n, ctx.world, ctx.rank)
    ks = uniform_scalars(n, (SEED + 31) & 0xFFFFFFFF)[lo:hi]
    sc = np.ascontiguousarray(uniform_scalars(n, (SEED + 32) & 0xFFFFFFFF)[lo:hi])
    d_gen = ctx.up(np.frombuffer(generator_record(cid), dtype=np.uint8))
    d_ks, d_sc = ctx.up(ks), ctx.up(sc)
    d_bases = torch.empty((hi - lo, 96), dtype=torch.uint8, device=ctx.dev)
    E.fixed_base_mul_device(cid, d_gen.data_ptr(), d_ks.data_ptr(), hi - lo, d_bases.data_ptr(), ctx.sp)
    job = ShardedMsm(cid, ctx.dev)
    ms = ctx.timed(lambda: job.run(d_bases, d_sc, hi - lo, ctx.sp), steps=10, warmup=2)
    res = job.run(d_bases, d_sc, hi - lo, ctx.sp)
    torch.cuda.synchronize()
    ok, detail = check_sharded_msm(ctx, CURVE, d_bases, sc, hi - lo, job.partial.cpu().numpy().tobytes(), res.cpu().numpy().tobytes())
    return {"workload": f"ONE BLS12-377 G1 MSM of 2^{ctx.args.log2n} pairs split over {ctx.world} GPU(s): b200_msm_device per rank "
                        "(single call, no batch pipelining), all-gather of 144 B partials, local sum",
            "scaling": "strong", "n_gpus": ctx.world, "ms": ms, "value": n / ms / 1e3, "unit": "Mpairs/s", "parity": ok,
            "parity_detail": detail}, ok


# ---- BASELINE config 4: BW6-761 G1 MSM, n = 2^22, input-chunk sharded -------------------------------------
def sub_config4(ctx: Ctx):
    torch, E = ctx.torch, ctx.E
    from celo_bls_snark_rs_b200.sharded import ShardedMsm, shard_bounds
    name, cid = "bw6_761_g1", E.BW6_761_G1
    n = 1 << ctx.args.bw6_log2n
    lo, hi = shard_bounds(n // 32, ctx.world, ctx.rank)
    lo, hi = lo * 32, hi * 32                              # whole runs of the generator per rank
    d_bases = gen_points(ctx, cid, lo, hi, n, seed=41)
    sc = np.ascontiguousarray(rand_scalars(n, 6, 56, 42)[lo:hi])   # < 2^376 < r: canonical
    d_sc = ctx.up(sc)
    job = ShardedMsm(cid, ctx.dev)
    run = lambda: job.run(d_bases, d_sc, hi - lo, ctx.sp)
    run()
    E.profile_enable(True)
    ms = ctx.timed(run, steps=3, warmup=1)
    acc_ms, acc_launches, acc_pairs = E.profile_read()
    E.profile_enable(False)
    res = run()
    torch.cuda.synchronize()
    ok, detail = check_sharded_msm(ctx, name, d_bases, sc, hi - lo, job.partial.cpu().numpy().tobytes(), res.cpu().numpy().tobytes(),
                                   sample=1 << 14)
    out = None
    if ctx.rank == 0:
        c, w, nb = E.msm_plan(cid, hi - lo)
        kernel_ms = acc_ms / max(acc_launches, 1)
        roof = msm_roofline(ctx, name, acc_pairs / max(acc_launches, 1), kernel_ms, 240, w)
        roof["kernel_share_of_step"] = kernel_ms / ms if ms else None
        out = {"workload": f"BW6-761 G1 MSM, n=2^{ctx.args.bw6_log2n} pairs (uniform 376-bit scalars, distinct bases), contiguous chunks "
                           f"over {ctx.world} GPU(s), all-gather of 288 B partials + local sum", "scaling": "strong",
               "n_gpus": ctx.world, "ms": ms, "value": n / ms / 1e3, "unit": "Mpairs/s", "window_bits": c, "windows": w,
               "parity": ok, "parity_detail": dict(detail, full_size="tests/test_msm_gpu.py compares the 2^22 MSM with the C port"),
               "roofline": roof}
        if not ctx.args.no_cpu:
            out["cpu_baseline"] = cpu_msm_baseline(ctx, name, d_bases[:1 << 15].cpu().numpy(), sc[:1 << 15], budget_s=5.0)
    return out, ok


# ---- BASELINE config 2: 4096 signatures -> a 4097-pair product of pairings ----------------------------------
def sub_config2(ctx: Ctx):
    torch, E = ctx.torch, ctx.E
    from celo_bls_snark_rs_b200.sharded import ShardedPairing, shard_bounds
    from oracle import cref as C
    n_sig = ctx.args.signatures
    g1c, g2c = E.BLS12_377_G1, E.BLS12_377_G2
    gen1 = np.frombuffer(generator_record(g1c), dtype=np.uint8)
    gen2 = np.frombuffer(generator_record(g2c), dtype=np.uint8)
    d_gen1, d_gen2 = ctx.up(gen1), ctx.up(gen2)
    sk = uniform_scalars(n_sig, 51)
    hs = uniform_scalars(n_sig, 52)
    d_sk, d_hs = ctx.up(sk), ctx.up(hs)
    d_pk = torch.empty((n_sig, 192), dtype=torch.uint8, device=ctx.dev)      # pk_i = sk_i * g2
    d_h = torch.empty((n_sig, 96), dtype=torch.uint8, device=ctx.dev)        # H_i = h_i * g1 (stands for the message hash)
    E.fixed_base_mul_device(g2c, d_gen2.data_ptr(), d_sk.data_ptr(), n_sig, d_pk.data_ptr(), ctx.sp)
    E.fixed_base_mul_device(g1c, d_gen1.data_ptr(), d_hs.data_ptr(), n_sig, d_h.data_ptr(), ctx.sp)
    d_sig_j = torch.zeros(144, dtype=torch.uint8, device=ctx.dev)            # sigma = sum sk_i H_i: the valid aggregate
    E.msm_device(g1c, d_h.data_ptr(), d_sk.data_ptr(), n_sig, d_sig_j.data_ptr(), ctx.sp)
    d_sig = torch.zeros(96, dtype=torch.uint8, device=ctx.dev)
    E.batch_to_affine_device(g1c, d_sig_j.data_ptr(), 1, d_sig.data_ptr(), ctx.sp)
    torch.cuda.synchronize()
    # pairs as signature.rs:135-150 builds them: (sigma, -g2), then (H_i, pk_i)
    neg = lambda b: ((P377 - int.from_bytes(b, "little")) % P377).to_bytes(48, "little")
    g2b = gen2.tobytes()
    neg_g2 = g2b[:96] + neg(g2b[96:144]) + neg(g2b[144:192])
    n = n_sig + 1
    g1_all = torch.cat([d_sig.reshape(1, 96), d_h]).contiguous()
    g2_all = torch.cat([ctx.up(np.frombuffer(neg_g2, dtype=np.uint8)).reshape(1, 192), d_pk]).contiguous()
    lo, hi = shard_bounds(n, ctx.world, ctx.rank)
    l1, l2 = g1_all[lo:hi].contiguous(), g2_all[lo:hi].contiguous()
    job = ShardedPairing(ctx.dev)
    ms = ctx.timed(lambda: job.run(l1, l2, hi - lo, ctx.sp), steps=10, warmup=2)
    mill = torch.zeros(576, dtype=torch.uint8, device=ctx.dev)
    miller_ms = ctx.timed(lambda: E.miller_product_device(l1.data_ptr(), l2.data_ptr(), hi - lo, mill.data_ptr(), ctx.sp), steps=10)
    gt, flag = job.run(l1, l2, hi - lo, ctx.sp)
    torch.cuda.synchronize()
    gt_gpu, is_one = gt.cpu().numpy().tobytes(), bool(flag.item())
    out, ok = None, True
    if ctx.rank == 0:
        h1 = np.zeros((n, 104), dtype=np.uint8)
        h2 = np.zeros((n, 200), dtype=np.uint8)
        h1[:, :96] = g1_all.cpu().numpy()
        h2[:, :192] = g2_all.cpu().numpy()
        # CPU: arkworks' schedule is ONE serial Miller loop over all pairs (threads = 1); timed, and its GT bytes are the check
        t0 = time.perf_counter()
        cpu_one, gt_cpu = C.multi_pairing(h1, h2, n, threads=1)
        cpu_s = time.perf_counter() - t0
        ok = bool(is_one and cpu_one and gt_gpu == gt_cpu)
        # a corrupted batch must fail on both sides
        bad1 = h1.copy()
        bad1[n // 2, :96] = h1[n // 2 - 1, :96]
        bad_ok, _ = E.multi_pairing(bad1, h2, n, want_gt=False)
        ok = ok and not bad_ok
        p1, p2 = torch.from_numpy(h1).pin_memory(), torch.from_numpy(h2).pin_memory()
        e2e = []
        for i in range(7):
            t0 = time.perf_counter()
            good, _ = E.multi_pairing(p1.numpy(), p2.numpy(), n, want_gt=False)
            e2e.append((time.perf_counter() - t0) * 1e3)
            ok = ok and good
        e2e_ms = statistics.median(e2e[2:])
        mads = n * 4800 * 288
        out = {"workload": f"batch aggregate BLS verify: {n_sig} signatures -> {n}-pair BLS12-377 product of pairings "
                           f"(Signature::batch_verify_hashes, signature.rs:125-155), pairs split over {ctx.world} GPU(s)",
               "n_gpus": ctx.world, "scaling": "strong", "ms": ms, "miller_ms": miller_ms, "value": n_sig / ms * 1e3,
               "unit": "signatures/s", "parity": ok,
               "parity_detail": {"checked": "GT bytes of all pairs against the C port of product_of_pairings (oracle/pairing_tmpl.h), "
                                            "== 1 on the valid batch, != 1 on a corrupted one", "pairs": n},
               "e2e": {"ms": e2e_ms, "value": n_sig / e2e_ms * 1e3, "unit": "signatures/s", "h2d_bytes_per_step": n * 304,
                       "d2h_bytes_per_step": 4, "api": "b200_multi_pairing_bls12_377 (host pointers, arkworks records, one GPU)"},
               "roofline": {"bound": "hbm", "kernel": "k_w2_miller_loop", "achieved": n * 288 / (miller_ms / 1e3) / 1e9,
                            "peak": measured_peak_gbs()[0], "unit": "GB/s", "frac": n * 288 / (miller_ms / 1e3) / 1e9 / measured_peak_gbs()[0],
                            "traffic": None, "kernel_ms": miller_ms, "algorithmic_bytes_per_launch": n * 288,
                            "int_pipe_frac": ctx.pipe_frac(mads, miller_ms, ctx.world),
                            "int_pipe_note": "arkworks' operation count (about 4800 Fq products per pair: G2 prepare + 69 sparse "
                                             "line products) x 288 wide multiply-adds, against the integer-pipe roof"},
               "cpu_baseline": {"value": n_sig / cpu_s, "unit": "signatures/s", "cores": 1, "kind": "port", "host_cores": ctx.cores,
                                "sample": f"all {n} pairs once ({cpu_s:.2f} s): one serial Miller loop + final exponentiation, "
                                          "as ark-ec 0.1.0 schedules it"}}
    return out, ctx.all_true(ok)


# ---- BASELINE config 5: epoch-snark Groth16 prove, arithmetic part -------------------------------------------
def _prover_inputs(ctx: Ctx, family: str, log_n: int):
    E, torch = ctx.E, ctx.torch
    outer = family == "bw6_761"
    g1, g2 = (E.BW6_761_G1, E.BW6_761_G2) if outer else (E.BLS12_377_G1, E.BLS12_377_G2)
    limbs, top = (6, 56) if outer else (4, 60)
    n = 1 << log_n
    num_assign = int(n * 0.6) // 32 * 32 - 1              # ~10^7 variables on a 2^24 domain
    num_aux = num_assign - 2
    b1 = gen_points(ctx, g1, 0, n, n, seed=61)            # a_query / l_query / h_query share one base array (cost is value-free)
    b2 = gen_points(ctx, g2, 0, num_assign + 1, num_assign + 1, seed=62)
    pk = E.Groth16Pk(b1.data_ptr(), b2.data_ptr(), b1.data_ptr(), b1.data_ptr(), b1.data_ptr(), b2.data_ptr())
    rng = np.random.default_rng(5)
    assign = np.zeros((num_assign, limbs), dtype=np.uint64)
    kind = rng.integers(0, 100, size=num_assign)
    assign[(kind >= 49) & (kind < 98), 0] = 1             # a witness is mostly bits: ~49 % zeros, ~49 % ones, 2 % dense
    dense = kind >= 98
    assign[dense] = rand_scalars(int(dense.sum()), limbs, top, 6)
    ev = rand_scalars(n, limbs, top - 1, 7)               # arbitrary residues < p as evaluation vectors
    d_assign = ctx.up(assign)
    src = [ctx.up(ev) for _ in range(3)]
    return dict(fam=E.GROTH16_BW6_761 if outer else E.GROTH16_BLS12_377, pk=pk, keep=(b1, b2), d_assign=d_assign, num_assign=num_assign,
                num_aux=num_aux, src=src, work=[torch.empty_like(s) for s in src], assign=assign, ev=ev, limbs=limbs)


def _prover_time(ctx: Ctx, P, log_n: int, steps: int):
    from celo_bls_snark_rs_b200.sharded import ShardedGroth16
    torch = ctx.torch
    job = ShardedGroth16(P["fam"], ctx.dev)
    times = []
    for rep in range(steps + 1):
        for w, s in zip(P["work"], P["src"]):
            w.copy_(s)
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ctx.stream)
        proof = job.run(P["pk"], P["d_assign"], P["num_assign"], P["num_aux"], P["work"][0], P["work"][1], P["work"][2], log_n, ctx.sp)
        e1.record(ctx.stream)
        ctx.barrier()
        times.append(ctx.max_over_ranks(e0.elapsed_time(e1))[0])
    return min(times[1:]), proof.cpu().numpy().tobytes()


def _prover_cpu(ctx: Ctx, family: str, P, log_n: int):
    """The same composition on the host with the C ports (BW6-761 family): witness map, four MSMs, the group sums
    -> (seconds, A | B | C compressed)."""
    from oracle import cref as C
    from oracle import oracle as O
    assert family == "bw6_761"
    L1, L2, field = C.LAYOUTS["bw6_761_g1"], C.LAYOUTS["bw6_761_g2"], 1
    n = 1 << log_n
    b1, b2 = P["keep"][0].cpu().numpy(), P["keep"][1].cpu().numpy()
    na, nx = P["num_assign"], P["num_aux"]
    th = ctx.cores
    import ctypes
    t0 = time.perf_counter()
    h = C.witness_map(field, P["ev"], P["ev"], P["ev"], log_n, threads=th)
    # into_repr: Montgomery residues -> canonical integers, through the C field's own product with 1
    lib, hc, vp = C.lib(), np.zeros_like(h), ctypes.c_void_p
    for i in range(n - 1):
        lib.cpu_ref_from_mont(6, h[i:i + 1].ctypes.data_as(vp), hc[i:i + 1].ctypes.data_as(vp))
    tw = lambda L, m: min(th, C.msm_window_tasks(L, max(m, 1)))
    a_acc = C.msm(L1, b1[1:na + 1], P["assign"], threads=tw(L1, na))
    l_acc = C.msm(L1, b1[:nx], P["assign"][na - nx:], threads=tw(L1, nx))
    h_acc = C.msm(L1, b1[:n - 1], hc[:n - 1], threads=tw(L1, n))
    b_acc = C.msm(L2, b2[1:na + 1], P["assign"], threads=tw(L2, na))
    dt = time.perf_counter() - t0
    aff = lambda L, rec: L.affine_from_records(rec.reshape(1, -1))[0]
    A = L1.curve.padd(L1.curve.padd(aff(L1, b1[0]), L1.jacobian_to_affine(a_acc)), aff(L1, b1[0]))
    B = L2.curve.padd(L2.curve.padd(aff(L2, b2[0]), L2.jacobian_to_affine(b_acc)), aff(L2, b2[0]))
    Cc = L1.curve.padd(L1.jacobian_to_affine(l_acc), L1.jacobian_to_affine(h_acc))
    return dt, O.serialize_compressed(L1.curve, A) + O.serialize_compressed(L2.curve, B) + O.serialize_compressed(L1.curve, Cc)


def _prover_roofline(ctx: Ctx, outer, n):
    """Dominant kernel of the outer proof: k_bucket_accumulate<Fq761> of the h MSM (dense 377-bit scalars; the a / l / b MSMs
    run over witness-density scalars: half zeros, half ones -- one mixed addition per unit scalar, no windows)."""
    E = ctx.E
    share = max(1, (n - 1) // ctx.world)
    c, w, nb = E.msm_plan(E.BW6_761_G1, share)
    na = outer["num_assign"] // ctx.world
    madds = w * share + 3 * (0.49 * na + 0.02 * na * w)             # h, then a / l / b: ones once, dense ones per window
    peak, src = measured_peak_gbs()
    kms = outer["accumulate_ms_per_proof"]
    achieved = outer["msm_pairs_per_proof_per_gpu"] * 240 / (kms / 1e3) / 1e9 if kms else 0.0
    return {"bound": "hbm", "kernel": "k_bucket_accumulate<Fq761> (4 launches per proof)", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": None, "peak_source": src, "kernel_ms": kms,
            "kernel_share_of_step": kms / outer["ms"] if outer["ms"] else None,
            "algorithmic_bytes_per_launch": outer["msm_pairs_per_proof_per_gpu"] * 240 / 4,
            "int_pipe_frac": ctx.pipe_frac(madds * MADS_PER_MADD["bw6_761_g1"], kms)}


def sub_config5(ctx: Ctx):
    from oracle import cref as C
    from oracle import oracle as O
    args = ctx.args
    res, ok_all = {}, True
    for family, log_n, label in (("bw6_761", args.prove_log2n, "outer"), ("bls12_377", max(args.prove_log2n - 2, 8), "inner")):
        P = _prover_inputs(ctx, family, log_n)
        ctx.E.profile_enable(True)
        ms, _ = _prover_time(ctx, P, log_n, steps=2)
        acc_ms, acc_launches, acc_pairs = ctx.E.profile_read()
        ctx.E.profile_enable(False)
        res[label] = {"family": family, "log2_domain": log_n, "num_assign": P["num_assign"], "ms": ms,
                      "accumulate_ms_per_proof": acc_ms / 3, "msm_pairs_per_proof_per_gpu": acc_pairs / 3}
        del P
        ctx.torch.cuda.empty_cache()
    # parity and the CPU figure on a bounded domain: the same code path (shards and all) against the C composition
    s_log = args.prove_sample_log2n
    P = _prover_inputs(ctx, "bw6_761", s_log)
    gpu_ms, proof = _prover_time(ctx, P, s_log, steps=1)
    out = None
    if ctx.rank == 0:
        L1, L2 = C.LAYOUTS["bw6_761_g1"], C.LAYOUTS["bw6_761_g2"]
        cpu_s, want = _prover_cpu(ctx, "bw6_761", P, s_log)
        got = L1.jacobian_compressed(proof[:288]) + L2.jacobian_compressed(proof[288:576]) + L1.jacobian_compressed(proof[576:])
        ok_all = got == want
        total = res["outer"]["ms"] + res["inner"]["ms"]
        n = 1 << args.prove_log2n
        out = {"workload": "epoch-snark Groth16 prove, ARITHMETIC PART after constraint synthesis (witness map: 7 transforms; MSMs a / l / h "
                           "in G1, b in G2; assembly) for the outer BW6-761 proof and the inner BLS12-377 proof on synthetic witnesses of "
                           "the estimated shape (SURVEY.md section 8d cfg5); every GPU runs the witness map and its share of each MSM",
               "n_gpus": ctx.world, "scaling": "strong", "ms": total, "outer": res["outer"], "inner": res["inner"],
               "value": n / (res["outer"]["ms"] / 1e3) / 1e6, "unit": "M domain points/s (outer proof)", "parity": ok_all,
               "parity_detail": {"checked": f"A | B | C of the BW6-761 composite at domain 2^{s_log} (same sharded code path) against the C "
                                            "composition: C witness map + C MSMs + oracle group sums, canonical compressed bytes",
                                 "gpu_ms_at_sample": gpu_ms},
               "cpu_baseline": {"value": (1 << s_log) / cpu_s / 1e6, "unit": "M domain points/s", "cores": ctx.cores, "kind": "port",
                                "sample": f"BW6-761 prover arithmetic at domain 2^{s_log} ({cpu_s:.2f} s): C witness map (threads over "
                                          "butterflies) + four C MSMs (threads over windows)"},
               "roofline": _prover_roofline(ctx, res["outer"], n),
               "excluded": "constraint synthesis (serial Rust over the R1CS gadgets) cannot run here and is not part of the figure"}
    return out, ctx.all_true(ok_all)


def run_b200(args):
    ctx = Ctx(args)
    line, ok = headline(ctx)
    subs = {}
    if not args.headline_only:
        for key, fn in (("config3_msm_2p20_strong", sub_config3_strong), ("config2_batch_verify_4096", sub_config2),
                        ("config4_bw6_761_g1_msm", sub_config4), ("config5_groth16_prove", sub_config5)):
            t0 = time.perf_counter()
            try:
                sub, sub_ok = fn(ctx)
            except Exception as exc:                                  # a broken sub-result must not hide the headline
                sub, sub_ok = {"error": f"{type(exc).__name__}: {exc}", "trace": traceback.format_exc()[-1500:]}, False
                if ctx.world > 1:
                    raise
            ok = ok and sub_ok
            if sub is not None:
                sub["bench_wall_s"] = round(time.perf_counter() - t0, 1)
            subs[key] = sub
            ctx.torch.cuda.empty_cache()
    if ctx.rank == 0:
        line["configs"] = subs
        line["parity"] = bool(ok)
        emit(line)
    if ctx.world > 1:
        ctx.dist.destroy_process_group()
    if not ok:
        raise SystemExit("bench.py: PARITY FAILED (see the `parity` fields of the JSON line)")


_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: keep a private handle on it and point fd 1 at stderr, so that whatever a native
    library prints there (NCCL's version banner under NCCL_DEBUG=VERSION, which ignores NCCL_DEBUG_FILE) cannot precede it."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=LOG2_N)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs (parity checks still run)")
    ap.add_argument("--headline-only", action="store_true", help="config 3 only (no sub-results)")
    ap.add_argument("--bw6-log2n", type=int, default=22, help="config 4 size")
    ap.add_argument("--signatures", type=int, default=4096, help="config 2 size")
    ap.add_argument("--prove-log2n", type=int, default=24, help="config 5: domain of the outer proof (inner: 4x smaller)")
    ap.add_argument("--prove-sample-log2n", type=int, default=12, help="config 5: domain of the parity / CPU sample")
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
