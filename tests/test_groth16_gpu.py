"""Groth16 prover arithmetic on the GPU (b200_groth16_prove_device) against the composition of the
oracles: witness map (oracle/ntt.py) + arkworks-algorithm MSMs (oracle/cpu_ref.c) + the group sums of
ark-groth16's create_proof_no_zk (crates/epoch-snark/src/api/prover.rs:78,112).
The proving key is synthetic (random curve points): the arithmetic does not care, and the reference
ships no proving key.  Compared as canonical compressed points."""
import numpy as np
import pytest

from oracle import cref as C
from oracle import inputs as H
from oracle import ntt as N
from oracle import oracle as O

pytestmark = pytest.mark.gpu

FAMILIES = {
    "bls12_377": (0, "bls12_377_g1", "bls12_377_g2", N.FR_BLS12_377),
    "bw6_761": (1, "bw6_761_g1", "bw6_761_g2", N.FR_BW6_761),
}


@pytest.fixture(scope="module")
def eng():
    from celo_bls_snark_rs_b200 import engine as E
    E.init(0)
    return E


@pytest.mark.parametrize("family,log_n,num_inputs", [("bls12_377", 6, 3), ("bw6_761", 6, 2), ("bls12_377", 9, 5), ("bw6_761", 8, 1),
                                                     ("bw6_761", 16, 3), ("bls12_377", 16, 2)])
def test_prover_arithmetic_matches_oracle_composition(eng, family, log_n, num_inputs):
    import torch
    fam_id, g1n, g2n, f = FAMILIES[family]
    L1, L2 = C.LAYOUTS[g1n], C.LAYOUTS[g2n]
    dev = torch.device("cuda:0")
    n = 1 << log_n
    num_aux = n - 2 * num_inputs - 5                      # constraints + inputs fit the domain
    num_assign = num_inputs + num_aux
    rng = O.SplitMix64(99 + log_n)
    # witness-shaped assignment: zeros, ones and dense values
    assign = [(0 if k % 3 == 0 else 1 if k % 3 == 1 else rng.below(f.p)) for k in range(num_assign)]
    a = [rng.below(f.p) for _ in range(n)]
    b = [rng.below(f.p) for _ in range(n)]
    c = [x * y % f.p for x, y in zip(a, b)]
    for i in range(n - num_inputs - 1, n):
        b[i] = c[i] = 0
    pts1 = H.random_points(g1n, 2 * num_assign + n + 8, 5, distinct=48)
    pts2 = H.random_points(g2n, num_assign + 4, 6, distinct=24)
    a_query, l_query = pts1[:num_assign + 1], pts1[num_assign + 1:num_assign + 1 + num_aux]
    h_query = pts1[2 * num_assign + 1:2 * num_assign + n]
    alpha, beta = pts1[-1], pts2[-1]
    b_query = pts2[:num_assign + 1]
    assert len(h_query) == n - 1 and len(l_query) == num_aux

    # ---- oracle composition ----
    if log_n <= 10:
        h = N.witness_map(f, a, b, c)
    else:                                                 # the C port of the same chain (pinned against N.witness_map in the CPU suite)
        h = f.from_mont_array(C.witness_map(f.id, f.to_mont_array(a), f.to_mont_array(b), f.to_mont_array(c), log_n, threads=8))
    sc = lambda vals: L1.scalars_array(vals)
    msm = lambda L, pts, vals: L.jacobian_to_affine(C.msm(L, L.affine_records(pts), sc(vals))) if vals else None
    a_acc = msm(L1, a_query[1:], assign)
    l_acc = msm(L1, l_query, assign[num_inputs:])
    h_acc = msm(L1, h_query, h[:n - 1])
    b_acc = msm(L2, b_query[1:], assign)
    want_a = L1.curve.padd(L1.curve.padd(a_query[0], a_acc), alpha)
    want_b = L2.curve.padd(L2.curve.padd(b_query[0], b_acc), beta)
    want_c = L1.curve.padd(l_acc, h_acc)

    # ---- device ----
    up = lambda arr: torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).copy()).to(dev)
    d = {k: up(L.affine_records(v, L.packed_stride)) for k, L, v in
         (("a", L1, a_query), ("b", L2, b_query), ("h", L1, h_query), ("l", L1, l_query), ("alpha", L1, [alpha]), ("beta", L2, [beta]))}
    pk = eng.Groth16Pk(d["a"].data_ptr(), d["b"].data_ptr(), d["h"].data_ptr(), d["l"].data_ptr(), d["alpha"].data_ptr(),
                       d["beta"].data_ptr())
    d_assign = up(sc(assign))
    da, db, dc = (up(f.to_mont_array(v)) for v in (a, b, c))
    d_proof = torch.zeros(2 * L1.jac_bytes + L2.jac_bytes, dtype=torch.uint8, device=dev)
    eng.groth16_prove_device(fam_id, pk, d_assign.data_ptr(), num_assign, num_aux, da.data_ptr(), db.data_ptr(), dc.data_ptr(),
                             log_n, d_proof.data_ptr())
    eng.sync()
    raw = d_proof.cpu().numpy().tobytes()
    j1, j2 = L1.jac_bytes, L2.jac_bytes
    assert L1.jacobian_to_affine(raw[:j1]) == want_a
    assert L2.jacobian_to_affine(raw[j1:j1 + j2]) == want_b
    assert L1.jacobian_to_affine(raw[j1 + j2:]) == want_c
