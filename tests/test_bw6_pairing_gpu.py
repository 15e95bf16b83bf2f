"""BW6-761 product of pairings and Groth16 verification on the GPU (through the C-ABI) against the oracle's
restatement (oracle/bw6_ate.py), and the reference's own verifier known-answer test
(crates/bls-snark-sys/src/snark/mod.rs:52-119: real VK, proof and epoch blocks, expected `true`) through
b200_groth16_verify_bw6_761.  Bit-exact: the Fq6 value is compared byte for byte."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bw6_kat import PROOF, VK, kat_inputs  # noqa: E402

from oracle import bw6_ate as A  # noqa: E402
from oracle import cref as C  # noqa: E402
from oracle import oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
L1, L2 = C.LAYOUTS["bw6_761_g1"], C.LAYOUTS["bw6_761_g2"]


@pytest.fixture(scope="module")
def eng():
    from celo_bls_snark_rs_b200 import engine as E
    E.init(0)
    return E


def fq6_ark_bytes(v):
    """power basis -> arkworks Fq6 image: (c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2), a_k = c_{k & 1}[k >> 1]."""
    order = [0, 2, 4, 1, 3, 5]
    return b"".join(L1.fe_to_mont_bytes(v[k]) for k in order)


def _pairs():
    # the verifying key and the proof of the reference's known-answer test are real r-torsion points
    g1 = [VK["alpha"], PROOF[0], PROOF[2], VK["gamma_abc"][0], VK["gamma_abc"][1]]
    g2 = [VK["beta"], PROOF[1], VK["gamma"], VK["delta"], O.BW6_G2.pmul(VK["beta"], 0xC0FFEE)]
    return g1, g2


@pytest.mark.parametrize("n", [0, 1, 2, 5])
def test_fq6_bytes_match_oracle(eng, n):
    g1, g2 = (x[:n] for x in _pairs())
    if n >= 5:
        g1[1] = None                                   # a pair with an infinite member contributes one
        g2[3] = None
    want = A.product_of_pairings(list(zip(g1, g2)))
    is_one, gt = eng.multi_pairing_bw6(L1.affine_records(g1), L2.affine_records(g2), n)
    assert gt == fq6_ark_bytes(want)
    assert is_one == (want == A.f6_one())
    is_one2, gt2 = eng.multi_pairing_bw6(L1.affine_records(g1, L1.packed_stride), L2.affine_records(g2, L2.packed_stride), n)
    assert gt2 == gt and is_one2 == is_one


def test_bilinearity_and_cancellation_on_device(eng):
    p, q = VK["alpha"], VK["beta"]
    a, b = 0x1234567, 0x7654321
    e_ab = eng.multi_pairing_bw6(L1.affine_records([O.BW6_G1.pmul(p, a)]), L2.affine_records([O.BW6_G2.pmul(q, b)]))[1]
    e_1 = eng.multi_pairing_bw6(L1.affine_records([O.BW6_G1.pmul(p, a * b % A.R)]), L2.affine_records([q]))[1]
    assert e_ab == e_1
    one, _ = eng.multi_pairing_bw6(L1.affine_records([p, O.BW6_G1.pneg(p)]), L2.affine_records([q, q]), want_gt=False)
    assert one is True


def _verify(eng, vk, proof, inputs, stride=None):
    rec1 = lambda pts: L1.affine_records(pts, stride)
    rec2 = lambda pts: L2.affine_records(pts, stride)
    return eng.groth16_verify_bw6(rec1([vk["alpha"]]), rec2([vk["beta"]]), rec2([vk["gamma"]]), rec2([vk["delta"]]),
                                  rec1(vk["gamma_abc"]), rec1([proof[0]]), rec2([proof[1]]), rec1([proof[2]]),
                                  L1.scalars_array(inputs))


def test_reference_verifier_kat_on_device(eng):
    """simple_verifier_groth16_with_entropy (snark/mod.rs:68-119): verify(...) == true."""
    inputs = kat_inputs()
    assert _verify(eng, VK, PROOF, inputs) is True
    assert _verify(eng, VK, PROOF, inputs, L1.packed_stride) is True


def test_reference_verifier_kat_rejects_changes_on_device(eng):
    inputs = kat_inputs()
    assert _verify(eng, VK, PROOF, [inputs[0], inputs[1] ^ 1]) is False            # a different public input
    a, b, c = PROOF
    assert _verify(eng, VK, (O.BW6_G1.padd(a, a), b, c), inputs) is False          # a different proof element
    assert _verify(eng, VK, (a, b, O.BW6_G1.pneg(c)), inputs) is False
    vk2 = dict(VK, delta=O.BW6_G2.padd(VK["delta"], VK["delta"]))
    assert _verify(eng, vk2, PROOF, inputs) is False
    with pytest.raises(eng.B200Error):                                               # MalformedVerifyingKey
        _verify(eng, VK, PROOF, inputs[:1])


def test_sharded_miller_values_equal_single_shot(eng):
    """The device-pointer halves: Miller values of two chunks of pairs, one final exponentiation."""
    import torch
    g1, g2 = _pairs()
    want = fq6_ark_bytes(A.product_of_pairings(list(zip(g1, g2))))
    d1 = torch.from_numpy(L1.affine_records(g1, L1.packed_stride)).cuda()
    d2 = torch.from_numpy(L2.affine_records(g2, L2.packed_stride)).cuda()
    vals = torch.zeros(2 * 5 * 576, dtype=torch.uint8, device="cuda")
    out = torch.zeros(576, dtype=torch.uint8, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    eng.miller_values_bw6_device(d1.data_ptr(), d2.data_ptr(), 2, vals.data_ptr())
    eng.miller_values_bw6_device(d1.data_ptr() + 2 * 192, d2.data_ptr() + 2 * 192, 3, vals.data_ptr() + 4 * 576)
    eng.final_exp_bw6_device(vals.data_ptr(), 10, out.data_ptr(), flag.data_ptr())
    eng.sync()
    assert out.cpu().numpy().tobytes() == want and int(flag.item()) == 0
