"""Parity of the CUDA MSM (through the C-ABI) against the CPU oracles.  Needs a B200."""
import numpy as np
import pytest

from oracle import cref as C
from oracle import oracle as O
from oracle import inputs as H

pytestmark = pytest.mark.gpu

CURVES = ["bls12_377_g1", "bls12_377_g2", "bw6_761_g1", "bw6_761_g2"]


@pytest.fixture(scope="module")
def eng():
    from celo_bls_snark_rs_b200 import engine as E
    E.init(0)
    return E


@pytest.mark.parametrize("name", CURVES)
@pytest.mark.parametrize("n", [0, 1, 2, 7, 31, 32, 33, 100, 257])
def test_msm_small_with_edge_cases_matches_naive_oracle(eng, name, n):
    L = C.LAYOUTS[name]
    pts, scalars = H.edge_case_inputs(name, n, 4242 + n)
    want = O.serialize_compressed(L.curve, L.curve.msm_naive(pts, scalars)) if n <= 33 else \
        L.jacobian_compressed(C.msm(L, L.affine_records(pts), L.scalars_array(scalars)))
    sc = L.scalars_array(scalars)
    got_ark = eng.msm(L.id, L.affine_records(pts), sc)                       # arkworks stride + flags
    got_packed = eng.msm(L.id, L.affine_records(pts, L.packed_stride), sc)   # packed, (0,0) = infinity
    assert L.jacobian_compressed(got_ark) == want
    assert L.jacobian_compressed(got_packed) == want


@pytest.mark.parametrize("name", CURVES)
def test_msm_all_same_base_and_scalar(eng, name):
    """Every bucket sees only duplicates: exercises the P + P path exclusively."""
    L = C.LAYOUTS[name]
    p = H.random_points(name, 1, 9)[0]
    n, k = 64, 0x1234567
    got = eng.msm(L.id, L.affine_records([p] * n), L.scalars_array([k] * n))
    assert L.jacobian_to_affine(got) == L.curve.pmul(p, n * k)


@pytest.mark.parametrize("name,n", [("bls12_377_g1", 1 << 14), ("bls12_377_g2", 1 << 12), ("bw6_761_g1", 1 << 12)])
def test_msm_medium_matches_c_oracle(eng, name, n):
    L = C.LAYOUTS[name]
    bases = L.affine_records(H.random_points(name, n, 77, distinct=256))
    sc = H.random_scalars_array(L, n, 5)
    want = L.jacobian_compressed(C.msm(L, bases, sc))
    assert L.jacobian_compressed(eng.msm(L.id, bases, sc)) == want


def test_named_entry_points(eng):
    import ctypes
    lib = eng.load()
    for name, fn in (("bls12_377_g1", lib.b200_msm_bls12_377_g1), ("bls12_377_g2", lib.b200_msm_bls12_377_g2),
                     ("bw6_761_g1", lib.b200_msm_bw6_761_g1), ("bw6_761_g2", lib.b200_msm_bw6_761_g2)):
        L = C.LAYOUTS[name]
        pts, scalars = H.edge_case_inputs(name, 40, 11)
        bases, sc = L.affine_records(pts), L.scalars_array(scalars)
        assert bases.strides[0] == eng.ARK_STRIDE[L.id]
        out = np.zeros(L.jac_bytes, dtype=np.uint8)
        rc = fn(bases.ctypes.data_as(ctypes.c_void_p), sc.ctypes.data_as(ctypes.c_void_p), 40,
                out.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0
        assert L.jacobian_to_affine(out.tobytes()) == L.curve.msm_naive(pts, scalars)


def test_error_paths(eng):
    lib = eng.load()
    assert lib.b200_msm(99, None, 104, None, 0, None) == 1           # bad curve id
    assert lib.b200_msm(0, None, 104, None, 5, None) == 1            # null pointers
    assert b"null" in lib.b200_last_error() or b"unknown" in lib.b200_last_error()


def test_bls12_377_g1_full_size_linearity_and_c_oracle(eng):
    """BASELINE config 3 (n = 2^20): size-independent properties + one full CPU comparison.
    MSM(b, s) + MSM(b, t) == MSM(b, s + t mod r), and MSM(b, s) == C oracle."""
    import torch
    name, n = "bls12_377_g1", 1 << 20
    L = C.LAYOUTS[name]
    dev = torch.device("cuda:0")
    # bases: k_i * G computed on the GPU from seeded scalars, spot-checked against the oracle
    ks = H.random_scalars_array(L, n, 123)
    d_ks = torch.from_numpy(ks.view(np.int64)).to(dev)
    d_gen = torch.from_numpy(L.affine_records([O.G1_GEN], L.packed_stride).copy()).to(dev)
    d_bases = torch.empty((n, L.packed_stride), dtype=torch.uint8, device=dev)
    eng.fixed_base_mul_device(L.id, d_gen.data_ptr(), d_ks.data_ptr(), n, d_bases.data_ptr())
    eng.sync()
    bases = d_bases.cpu().numpy()
    for i in (0, 1, 12345, n - 1):
        k = L.scalars_from_array(ks[i:i + 1])[0]
        assert L.affine_from_records(bases[i:i + 1])[0] == L.jacobian_to_affine(C.scalar_mul(L, O.G1_GEN, k))
    s = H.random_scalars_array(L, n, 1)
    t = H.random_scalars_array(L, n, 2)
    s_int, t_int = L.scalars_from_array(s[:64]), L.scalars_from_array(t[:64])
    # s + t mod r with Python ints would be slow for 2^20 rows: top bits were cleared, so s + t < 2r;
    # do the conditional subtraction with numpy object arrays once.
    st_int = [(a + b) % O.R for a, b in zip(L.scalars_from_array(s), L.scalars_from_array(t))]
    st = L.scalars_array(st_int)
    assert st_int[:64] == [(a + b) % O.R for a, b in zip(s_int, t_int)]
    d_out = torch.empty(3 * L.jac_bytes, dtype=torch.uint8, device=dev)
    outs = []
    for j, arr in enumerate((s, t, st)):
        d_sc = torch.from_numpy(arr.view(np.int64)).to(dev)
        eng.msm_device(L.id, d_bases.data_ptr(), d_sc.data_ptr(), n, d_out.data_ptr() + j * L.jac_bytes)
        eng.sync()
        outs.append(L.jacobian_to_affine(d_out[j * L.jac_bytes:(j + 1) * L.jac_bytes].cpu().numpy().tobytes()))
    assert L.curve.padd(outs[0], outs[1]) == outs[2]
    want = L.jacobian_to_affine(C.msm(L, bases, s))
    assert outs[0] == want
    # the host-pointer API must agree with the device API
    assert L.jacobian_to_affine(eng.msm(L.id, bases, s)) == want


@pytest.mark.parametrize("name", ["bls12_377_g1", "bls12_377_g2", "bw6_761_g1"])
def test_batch_to_affine_and_sum_jacobian(eng, name):
    import torch
    L = C.LAYOUTS[name]
    dev = torch.device("cuda:0")
    pts = H.random_points(name, 21, 3, distinct=21)
    # Jacobian inputs with non-trivial Z: k * P from the C oracle's double-and-add, plus infinity
    jac = [C.scalar_mul(L, p, 5 + i) for i, p in enumerate(pts)]
    jac[4] = C.scalar_mul(L, pts[4], 0)
    want = [L.curve.pmul(p, 5 + i) for i, p in enumerate(pts)]
    want[4] = None
    d_jac = torch.from_numpy(np.frombuffer(b"".join(jac), dtype=np.uint8).copy()).to(dev)
    d_aff = torch.empty((len(pts), L.packed_stride), dtype=torch.uint8, device=dev)
    eng.batch_to_affine_device(L.id, d_jac.data_ptr(), len(pts), d_aff.data_ptr())
    d_sum = torch.empty(L.jac_bytes, dtype=torch.uint8, device=dev)
    eng.sum_jacobian_device(L.id, d_jac.data_ptr(), len(pts), d_sum.data_ptr())
    eng.sync()
    assert L.affine_from_records(d_aff.cpu().numpy()) == want
    acc = None
    for w in want:
        acc = L.curve.padd(acc, w)
    assert L.jacobian_to_affine(d_sum.cpu().numpy().tobytes()) == acc


@pytest.mark.parametrize("name", ["bls12_377_g1", "bls12_377_g2", "bw6_761_g1"])
def test_fixed_base_mul_matches_oracle(eng, name):
    import torch
    L = C.LAYOUTS[name]
    dev = torch.device("cuda:0")
    rng = O.SplitMix64(31)
    g = H.generator(name, rng)
    ks = [0, 1, 2, L.curve.scalar_mod - 1] + [rng.below(L.curve.scalar_mod) for _ in range(29)]
    d_g = torch.from_numpy(L.affine_records([g], L.packed_stride).copy()).to(dev)
    d_k = torch.from_numpy(L.scalars_array(ks).view(np.int64)).to(dev)
    d_o = torch.empty((len(ks), L.packed_stride), dtype=torch.uint8, device=dev)
    eng.fixed_base_mul_device(L.id, d_g.data_ptr(), d_k.data_ptr(), len(ks), d_o.data_ptr())
    eng.sync()
    got = L.affine_from_records(d_o.cpu().numpy())
    want = [L.jacobian_to_affine(C.scalar_mul(L, g, k)) for k in ks]
    assert got == want


@pytest.mark.parametrize("name,n", [("bls12_377_g1", 1 << 15), ("bw6_761_g1", 1 << 13), ("bls12_377_g2", 1 << 12)])
def test_msm_groth16_like_density(eng, name, n):
    """Witness-shaped scalars (crates/epoch-snark/src/api/prover.rs:78): ~49 % zeros, ~47 % ones,
    a run of one repeated small scalar (an over-populated bucket) and a few dense ones."""
    L = C.LAYOUTS[name]
    bases = L.affine_records(H.random_points(name, n, 909, distinct=128))
    rng = np.random.default_rng(17)
    sc = np.zeros((n, L.scalar_limbs), dtype=np.uint64)
    kind = rng.integers(0, 100, size=n)
    sc[(kind >= 49) & (kind < 96), 0] = 1                     # unit scalars
    sc[kind >= 96] = H.random_scalars_array(L, int((kind >= 96).sum()), 3)
    rep = slice(n // 4, n // 4 + min(3000, n // 3))           # > BIG_BUCKET entries in one bucket per window
    sc[rep] = 0
    sc[rep, 0] = 0x2B
    want = L.jacobian_compressed(C.msm(L, bases, sc))
    assert L.jacobian_compressed(eng.msm(L.id, bases, sc)) == want


def test_msm_all_unit_scalars(eng):
    L = C.LAYOUTS["bls12_377_g1"]
    n = 5000
    pts = H.random_points("bls12_377_g1", n, 4, distinct=50)
    acc = None
    for p in pts:
        acc = L.curve.padd(acc, p)
    got = eng.msm(L.id, L.affine_records(pts), L.scalars_array([1] * n))
    assert L.jacobian_to_affine(got) == acc


@pytest.mark.parametrize("name,n,run", [("bls12_377_g1", 1 << 15, 12000), ("bw6_761_g1", 1 << 14, 9000)])
def test_msm_huge_bucket_is_sliced(eng, name, n, run):
    """More than HUGE_BUCKET (8192) points with one and the same scalar: the bucket is summed by
    HUGE_SLICES blocks (k_huge_buckets / k_huge_finish) and must still equal the oracle."""
    L = C.LAYOUTS[name]
    bases = L.affine_records(H.random_points(name, n, 31337, distinct=64))
    sc = H.random_scalars_array(L, n, 8)
    sc[100:100 + run] = 0
    sc[100:100 + run, 0] = 0x1D3
    want = L.jacobian_compressed(C.msm(L, bases, sc))
    assert L.jacobian_compressed(eng.msm(L.id, bases, sc)) == want


@pytest.mark.parametrize("name", ["bls12_377_g1", "bw6_761_g1"])
def test_msm_batch_pipeline_matches_single_calls(eng, name):
    """b200_msm_batch_device: independent MSMs of different sizes (one of them empty), pipelined over
    the engine's internal streams, must give the same group elements as one call each and as the oracle."""
    import torch
    L = C.LAYOUTS[name]
    dev = torch.device("cuda:0")
    sizes = [3000, 0, 1 << 12, 257, 5000, 1]
    jobs, keep, want = [], [], []
    d_out = torch.zeros((len(sizes), L.jac_bytes), dtype=torch.uint8, device=dev)
    for j, n in enumerate(sizes):
        bases = L.affine_records(H.random_points(name, n, 400 + j, distinct=64), L.packed_stride)
        sc = H.random_scalars_array(L, n, 500 + j)
        d_b = torch.from_numpy(bases.copy()).to(dev) if n else None
        d_s = torch.from_numpy(sc.view(np.int64).copy()).to(dev) if n else None
        keep += [d_b, d_s]
        jobs.append((d_b.data_ptr() if n else 0, d_s.data_ptr() if n else 0, n, d_out[j].data_ptr()))
        want.append(L.jacobian_compressed(C.msm(L, bases, sc)))
    for _ in range(2):                                   # twice: workspace sets are reused across calls
        d_out.zero_()
        eng.msm_batch_device(L.id, jobs)
        eng.sync()
        got = [L.jacobian_compressed(d_out[j].cpu().numpy().tobytes()) for j in range(len(sizes))]
        assert got == want


def test_sharded_run_batch_single_rank(eng):
    """ShardedMsm.run_batch (what bench.py times) on one rank: K pipelined MSMs == K single calls."""
    import torch
    from celo_bls_snark_rs_b200.sharded import ShardedMsm
    L = C.LAYOUTS["bls12_377_g1"]
    dev = torch.device("cuda:0")
    n = 2000
    sets = []
    for j in range(2):
        bases = L.affine_records(H.random_points("bls12_377_g1", n, 70 + j, distinct=32), L.packed_stride)
        sc = H.random_scalars_array(L, n, 80 + j)
        sets.append((torch.from_numpy(bases.copy()).to(dev), torch.from_numpy(sc.view(np.int64).copy()).to(dev),
                     L.jacobian_compressed(C.msm(L, bases, sc))))
    job = ShardedMsm(L.id, dev)
    out = job.run_batch([(sets[i & 1][0], sets[i & 1][1], n) for i in range(5)])
    eng.sync()
    for i in range(5):
        assert L.jacobian_compressed(out[i].cpu().numpy().tobytes()) == sets[i & 1][2]


def test_config4_bw6_761_large_linearity_and_c_oracle(eng):
    """BASELINE config 4's curve at sizes the Python oracle cannot reach: BW6-761 G1, n = 2^18 against
    the C oracle, and n = 2^20 through linearity MSM(b, s) + MSM(b, t) == MSM(b, s + t) (no reduction of
    s + t: the seeded base point need not lie in the prime-order subgroup; s, t < 2^376 so the sum fits)."""
    import torch
    name = "bw6_761_g1"
    L = C.LAYOUTS[name]
    dev = torch.device("cuda:0")
    n = 1 << 20
    rng = O.SplitMix64(4)
    g = H.generator(name, rng)
    ks = H.random_scalars_array(L, n, 321)
    d_gen = torch.from_numpy(L.affine_records([g], L.packed_stride).copy()).to(dev)
    d_ks = torch.from_numpy(ks.view(np.int64)).to(dev)
    d_bases = torch.empty((n, L.packed_stride), dtype=torch.uint8, device=dev)
    eng.fixed_base_mul_device(L.id, d_gen.data_ptr(), d_ks.data_ptr(), n, d_bases.data_ptr())
    eng.sync()
    m = 1 << 18
    bases_m = d_bases[:m].cpu().numpy()
    s = H.random_scalars_array(L, n, 11)
    t = H.random_scalars_array(L, n, 12)
    st = L.scalars_array([a + b for a, b in zip(L.scalars_from_array(s), L.scalars_from_array(t))])
    d_out = torch.empty(4 * L.jac_bytes, dtype=torch.uint8, device=dev)
    keep = []
    jobs = []
    for j, (arr, cnt) in enumerate(((s, n), (t, n), (st, n), (s, m))):
        d_sc = torch.from_numpy(arr.view(np.int64)).to(dev)
        keep.append(d_sc)
        jobs.append((d_bases.data_ptr(), d_sc.data_ptr(), cnt, d_out.data_ptr() + j * L.jac_bytes))
    eng.msm_batch_device(L.id, jobs)                    # the four MSMs as one pipelined batch
    eng.sync()
    raw = d_out.cpu().numpy().tobytes()
    pts = [L.jacobian_to_affine(raw[j * L.jac_bytes:(j + 1) * L.jac_bytes]) for j in range(4)]
    assert L.curve.padd(pts[0], pts[1]) == pts[2]
    assert pts[3] == L.jacobian_to_affine(C.msm(L, bases_m, s[:m]))


@pytest.mark.parametrize("name", ["bls12_377_g1", "bls12_377_g2", "bw6_761_g1"])
def test_batch_normalisation_4096_points_with_infinities(eng, name):
    """Row a3 at a real size: batch_normalization_into_affine (signature.rs:82, public.rs:58) over 4096 Jacobian points
    with non-trivial Z, infinities sprinkled in (also a whole batch-of-8 group of them), against the oracle."""
    import torch
    L = C.LAYOUTS[name]
    dev = torch.device("cuda:0")
    n = 4096
    base = H.random_points(name, 64, 31, distinct=64)
    jac, want = [], []
    for i in range(n):
        k = 0 if (i % 97 == 5 or 800 <= i < 808) else 3 + (i % 11)
        j = C.scalar_mul(L, base[i % 64], k)
        jac.append(j)
        want.append(L.jacobian_to_affine(j))
    d_jac = torch.from_numpy(np.frombuffer(b"".join(jac), dtype=np.uint8).copy()).to(dev)
    d_aff = torch.empty((n, L.packed_stride), dtype=torch.uint8, device=dev)
    eng.batch_to_affine_device(L.id, d_jac.data_ptr(), n, d_aff.data_ptr())
    eng.sync()
    got = L.affine_from_records(d_aff.cpu().numpy())
    assert got == want and want[5] is None and want[803] is None


def test_bw6_761_g2_msm_4096(eng):
    """BW6-761 G2 (the b_g2_query MSM of the outer proof, prover.rs:78) at n = 2^12 with the edge cases, against the C port"""
    name, n = "bw6_761_g2", 1 << 12
    L = C.LAYOUTS[name]
    pts, scalars = H.edge_case_inputs(name, n, 77)
    bases, sc = L.affine_records(pts), L.scalars_array(scalars)
    assert L.jacobian_compressed(eng.msm(L.id, bases, sc)) == L.jacobian_compressed(C.msm(L, bases, sc))


def test_config4_bw6_761_g1_full_size_against_c_port(eng):
    """BASELINE config 4 at its full size, n = 2^22: the device MSM and the host-pointer MSM against the C port of arkworks'
    algorithm on the same inputs (the C port takes about half a minute on 16 cores)."""
    import torch
    name, n = "bw6_761_g1", 1 << 22
    L = C.LAYOUTS[name]
    dev = torch.device("cuda:0")
    rng = O.SplitMix64(5)
    g = H.generator(name, rng)
    starts = np.zeros((n // 32, 6), dtype=np.uint64)
    starts[:, 0] = np.random.default_rng(9).integers(1 << 40, 1 << 62, size=n // 32, dtype=np.uint64)
    d_gen = torch.from_numpy(L.affine_records([g], L.packed_stride).copy()).to(dev)
    d_st = torch.from_numpy(starts.view(np.int64)).to(dev)
    d_bases = torch.empty((n, L.packed_stride), dtype=torch.uint8, device=dev)
    eng.point_runs_device(L.id, d_gen.data_ptr(), d_st.data_ptr(), n // 32, 32, d_bases.data_ptr())
    eng.sync()
    bases = d_bases.cpu().numpy()
    for i in (0, 1, 33, n - 1):                                    # the generator kernel itself, spot-checked
        k = int(starts[i // 32, 0]) + i % 32
        assert L.affine_from_records(bases[i:i + 1])[0] == L.jacobian_to_affine(C.scalar_mul(L, g, k))
    sc = H.random_scalars_array(L, n, 10)
    d_sc = torch.from_numpy(sc.view(np.int64)).to(dev)
    d_out = torch.empty(L.jac_bytes, dtype=torch.uint8, device=dev)
    eng.msm_device(L.id, d_bases.data_ptr(), d_sc.data_ptr(), n, d_out.data_ptr())
    eng.sync()
    want = L.jacobian_compressed(C.msm(L, bases, sc))
    assert L.jacobian_compressed(d_out.cpu().numpy().tobytes()) == want
    assert L.jacobian_compressed(eng.msm(L.id, bases, sc)) == want     # chunk-fed host path, pageable memory


@pytest.mark.parametrize("name,log_n", [("bls12_377_g1", 19), ("bls12_377_g2", 18), ("bw6_761_g1", 18)])
def test_host_pointer_msm_pinned_and_pageable_chunks(eng, name, log_n):
    """The chunk-fed host path (one bucket set resumed chunk by chunk) from page-locked and from ordinary memory, with
    unit / zero scalars and duplicates spread over the chunks, at a ragged size; against the device path and the C port."""
    import torch
    L = C.LAYOUTS[name]
    n = (1 << log_n) + 12345
    pts = H.random_points(name, 512, 3, distinct=512)
    recs = L.affine_records(pts)
    bases = np.tile(recs, (n // 512 + 1, 1))[:n].copy()
    sc = H.random_scalars_array(L, n, 4)
    sc[::7] = 0
    sc[3::11] = 0
    sc[3::11, 0] = 1                                               # unit scalars in every chunk
    bases[5::1001, :] = 0
    bases[5::1001, 2 * L.coord_bytes] = 1                          # infinity flags
    want = L.jacobian_compressed(C.msm(L, bases, sc))
    assert L.jacobian_compressed(eng.msm(L.id, bases, sc)) == want                  # pageable: staged through pinned slots
    pb, ps = torch.from_numpy(bases).pin_memory(), torch.from_numpy(sc.view(np.int64)).pin_memory()
    out = np.zeros(L.jac_bytes, dtype=np.uint8)
    eng.msm_host_ptrs(L.id, pb.data_ptr(), bases.strides[0], ps.data_ptr(), n, out)
    assert L.jacobian_compressed(out.tobytes()) == want                             # pinned: copied from directly
