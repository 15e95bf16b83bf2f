"""One process, several GPUs behind the C-ABI (b200_init_devices, b200_msm_sharded, b200_multi_pairing_bls12_377_sharded):
a Rust caller of Signature::batch (crates/bls-crypto/src/bls/signature.rs:70-89) or of ark-groth16's MSMs is ONE process.
Runs on every GPU the box has; on a one-GPU box the same device is listed twice (B200_ALLOW_DUP_DEVICES=1), which still
drives the whole sharded path: one engine and host thread per slice, peer copies of the partial results, the combine."""
import os

import numpy as np
import pytest

from oracle import cref as C
from oracle import inputs as H
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import torch
    from celo_bls_snark_rs_b200 import engine as E
    count = torch.cuda.device_count()
    os.environ["B200_ALLOW_DUP_DEVICES"] = "1"
    E.init(0)
    E.init_devices(list(range(count)) if count > 1 else [0, 0])
    assert E.device_count() >= 2
    yield E
    E.shutdown()                                                   # back to one engine for the other modules
    E.init(0)


@pytest.mark.parametrize("name,n", [("bls12_377_g1", (1 << 19) + 77), ("bls12_377_g2", 30000), ("bw6_761_g1", 70001), ("bls12_377_g1", 900)])
def test_msm_sharded_matches_c_port(eng, name, n):
    L = C.LAYOUTS[name]
    pts = H.random_points(name, 256, 8, distinct=256)
    bases = np.tile(L.affine_records(pts), (n // 256 + 1, 1))[:n].copy()
    sc = H.random_scalars_array(L, n, 9)
    sc[::5] = 0
    sc[1::9] = 0
    sc[1::9, 0] = 1
    want = L.jacobian_compressed(C.msm(L, bases, sc))
    assert L.jacobian_compressed(eng.msm_sharded(L.id, bases, sc)) == want
    assert L.jacobian_compressed(eng.msm(L.id, bases, sc)) == want       # the primary engine still serves single-device calls


def test_multi_pairing_sharded_gt_bytes(eng):
    L1, L2 = C.LAYOUTS["bls12_377_g1"], C.LAYOUTS["bls12_377_g2"]
    n = 600
    g1, g2 = H.signature_batch(n, 11)
    r1, r2 = L1.affine_records(g1), L2.affine_records(g2)
    ok, gt = eng.multi_pairing_sharded(r1, r2)
    assert ok is True and gt == C.fq12_to_ark_bytes(O.FQ12_ONE)
    g1b, g2b = H.signature_batch(n, 11, corrupt=100)
    rb1, rb2 = L1.affine_records(g1b), L2.affine_records(g2b)
    bad, gt_bad = eng.multi_pairing_sharded(rb1, rb2)
    cpu_ok, gt_cpu = C.multi_pairing(rb1, rb2, n + 1, threads=8)
    assert bad is False and cpu_ok is False and gt_bad == gt_cpu
    assert eng.multi_pairing(rb1, rb2)[1] == gt_cpu
