"""CPU restatement of the radix-2 transforms on the Groth16 path.  TEST INFRASTRUCTURE ONLY:
nothing under celo_bls_snark_rs_b200/ imports this (see oracle/oracle.py header).

What it restates (the sources are third-party and absent from /root/reference: ark-poly 0.1.0 @
arkworks-rs/algebra#8d76d181, ark-groth16 0.1.0 @ arkworks-rs/groth16#d8acb2b2, Cargo.lock:175-215):
  * Radix2EvaluationDomain::{fft, ifft, coset_fft, coset_ifft}_in_place -- natural order in and out,
    group generator = TWO_ADIC_ROOT_OF_UNITY^(2^(s - log n)), ifft scales by 1/n, the coset offset is
    F::multiplicative_generator() (distribute_powers before the forward / after the inverse transform);
  * R1CStoQAP::witness_map's transform chain, entered from create_proof_no_zk at
    crates/epoch-snark/src/api/prover.rs:78 (BW6-761, Fr = BLS12-377 Fq) and :112 (BLS12-377 Fr):
    ifft(a), ifft(b), coset_fft(a), coset_fft(b), ab = a*b, ifft(c), coset_fft(c), ab -= c,
    ab /= Z(g) = g^n - 1, coset_ifft(ab).
PARITY UNPINNED at byte level: the reference holds no golden vector for an FFT or for h; the field
constants are pinned indirectly (tools/gen_params.py: 22^((r-1)/2^47) equals the published
BLS12-377 Fr root of unity; the Montgomery limbs of -5 and (-5)^((q-1)/2^46) equal the published Fq
constants).  Correctness of the transform itself is checked against the O(n^2) definition below and
through polynomial identities (tests/test_oracle_ntt.py).
"""
from __future__ import annotations

import numpy as np

from . import oracle as O


class ScalarField:
    def __init__(self, name, fid, p, gen, two_adicity, limbs):
        self.name, self.id, self.p, self.gen, self.s, self.limbs = name, fid, p, gen % p, two_adicity, limbs
        self.two_adic_root = pow(self.gen, (p - 1) >> two_adicity, p)
        self.mont_r = (1 << (64 * limbs)) % p
        self.mont_rinv = pow(self.mont_r, -1, p)

    def root_of_unity(self, log_n: int) -> int:
        assert log_n <= self.s
        return pow(self.two_adic_root, 1 << (self.s - log_n), self.p)

    # arkworks memory images: Montgomery residues, 64-bit little-endian limbs
    def to_mont_array(self, values) -> np.ndarray:
        out = np.zeros((len(values), self.limbs), dtype=np.uint64)
        for i, v in enumerate(values):
            m = v * self.mont_r % self.p
            for j in range(self.limbs):
                out[i, j] = (m >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
        return out

    def from_mont_array(self, arr: np.ndarray):
        return [sum(int(v) << (64 * j) for j, v in enumerate(row)) * self.mont_rinv % self.p for row in arr]


FR_BLS12_377 = ScalarField("fr_bls12_377", 0, O.R, 22, 47, 4)
FR_BW6_761 = ScalarField("fr_bw6_761", 1, O.P, -5, 46, 6)
FIELDS = {f.name: f for f in (FR_BLS12_377, FR_BW6_761)}


def dft_naive(f: ScalarField, a, omega):
    """out[k] = sum_j a[j] omega^(jk): the definition (tiny n only)."""
    n, p = len(a), f.p
    return [sum(a[j] * pow(omega, j * k, p) for j in range(n)) % p for k in range(n)]


def _transform(f: ScalarField, a, omega):
    """Recursive radix-2 Cooley-Tukey, natural order in and out (len(a) a power of two)."""
    n, p = len(a), f.p
    if n == 1:
        return list(a)
    w2 = omega * omega % p
    even, odd = _transform(f, a[0::2], w2), _transform(f, a[1::2], w2)
    out = [0] * n
    w = 1
    for k in range(n // 2):
        t = w * odd[k] % p
        out[k] = (even[k] + t) % p
        out[k + n // 2] = (even[k] - t) % p
        w = w * omega % p
    return out


def _log2(n):
    assert n and n & (n - 1) == 0
    return n.bit_length() - 1


def fft(f, a):
    return _transform(f, a, f.root_of_unity(_log2(len(a))))


def ifft(f, a):
    n = len(a)
    ninv = pow(n, -1, f.p)
    return [v * ninv % f.p for v in _transform(f, a, pow(f.root_of_unity(_log2(n)), -1, f.p))]


def _distribute_powers(f, a, g):
    out, w = [], 1
    for v in a:
        out.append(v * w % f.p)
        w = w * g % f.p
    return out


def coset_fft(f, a):
    return fft(f, _distribute_powers(f, a, f.gen))


def coset_ifft(f, a):
    return _distribute_powers(f, ifft(f, a), pow(f.gen, -1, f.p))


def witness_map(f, a, b, c):
    """h = (A B - C) / Z as n coefficients, from the n evaluations of A, B, C (appendix A.4 step 2)."""
    n, p = len(a), f.p
    ca, cb, cc = coset_fft(f, ifft(f, a)), coset_fft(f, ifft(f, b)), coset_fft(f, ifft(f, c))
    zinv = pow(pow(f.gen, n, p) - 1, -1, p)
    ab = [(x * y - z) * zinv % p for x, y, z in zip(ca, cb, cc)]
    return coset_ifft(f, ab)


def poly_eval(f, coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % f.p
    return acc
