"""The optimal-ate restatement the CUDA kernels follow (oracle/bw6_ate.py) pinned on the CPU: bilinearity and
non-degeneracy on the reference's own verifying key, agreement with the independent reduced Tate pairing of
oracle/bw6_verify.py, and the reference's verifier known-answer test
(crates/bls-snark-sys/src/snark/mod.rs:52-119, expected `true`).  CPU only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bw6_kat import PROOF, VK, kat_inputs  # noqa: E402

from oracle import bw6_ate as A  # noqa: E402
from oracle import bw6_verify as V  # noqa: E402
from oracle import oracle as O  # noqa: E402


def test_loop_counts_and_constants():
    assert (A.LOOP_1 + A.Q * A.LOOP_2) % A.R == 0
    assert sum(d << i for i, d in enumerate(A.LOOP_2_DIGITS)) == A.LOOP_2
    assert all(not (a and b) for a, b in zip(A.LOOP_2_DIGITS, A.LOOP_2_DIGITS[1:]))      # non-adjacent
    x = [3, 1, 4, 1, 5, 9]
    assert A.f6_frob(x, 1) == A.f6_pow(x, A.Q) and A.f6_conj(x) == A.f6_pow(x, A.Q ** 3)
    assert A.f6_mul(x, A.f6_inv(x)) == A.f6_one()


def test_optimal_ate_is_bilinear_and_non_degenerate():
    p, q = VK["alpha"], VK["beta"]
    e = A.product_of_pairings([(p, q)])
    assert e != A.f6_one() and A.f6_pow(e, A.R) == A.f6_one()
    assert A.product_of_pairings([(O.BW6_G1.pmul(p, 5), q)]) == A.f6_pow(e, 5)
    assert A.product_of_pairings([(p, O.BW6_G2.pmul(q, 7))]) == A.f6_pow(e, 7)
    assert A.product_of_pairings([(p, q), (O.BW6_G1.pneg(p), q)]) == A.f6_one()
    assert A.product_of_pairings([(None, q), (p, None)]) == A.f6_one()


def test_reference_verifier_kat_through_the_ate_pairing():
    inputs = kat_inputs()
    assert A.verify_proof(VK, PROOF, inputs) is True
    assert V.verify_proof(VK, PROOF, inputs) is True                       # the Tate oracle agrees
    bad_inputs = [inputs[0], inputs[1] ^ 1]
    assert A.verify_proof(VK, PROOF, bad_inputs) is False and V.verify_proof(VK, PROOF, bad_inputs) is False
    a, b, c = PROOF
    assert A.verify_proof(VK, (O.BW6_G1.padd(a, a), b, c), inputs) is False
