"""The NTT oracle (oracle/ntt.py) against the O(n^2) definition and polynomial identities (CPU)."""
import pytest

from oracle import ntt as N
from oracle import oracle as O


@pytest.mark.parametrize("fname", list(N.FIELDS))
def test_domain_constants(fname):
    f = N.FIELDS[fname]
    assert (f.p - 1) % (1 << f.s) == 0 and ((f.p - 1) >> f.s) % 2 == 1          # two-adicity
    assert pow(f.gen, (f.p - 1) // 2, f.p) == f.p - 1                           # the generator is a non-residue
    assert pow(f.two_adic_root, 1 << (f.s - 1), f.p) == f.p - 1                 # exact order 2^s
    for log_n in (0, 1, 5, 24):
        w = f.root_of_unity(log_n)
        assert pow(w, 1 << log_n, f.p) == 1 and (log_n == 0 or pow(w, 1 << (log_n - 1), f.p) == f.p - 1)


def test_published_bls12_377_fr_root_of_unity():
    # ark-bls12-377 Fr TWO_ADIC_ROOT_OF_UNITY as published (decimal), reproduced from GENERATOR = 22
    assert N.FR_BLS12_377.two_adic_root == 8065159656716812877374967518403273466521432693661810619979959746626482506078


@pytest.mark.parametrize("fname", list(N.FIELDS))
@pytest.mark.parametrize("n", [1, 2, 8, 32])
def test_fft_is_the_dft(fname, n):
    f = N.FIELDS[fname]
    rng = O.SplitMix64(1000 + n)
    a = [rng.below(f.p) for _ in range(n)]
    w = f.root_of_unity(n.bit_length() - 1)
    assert N.fft(f, a) == N.dft_naive(f, a, w)
    assert N.ifft(f, N.fft(f, a)) == a
    assert N.coset_ifft(f, N.coset_fft(f, a)) == a
    # coset_fft evaluates the polynomial at g * w^k
    ev = N.coset_fft(f, a)
    for k in (0, n - 1):
        assert ev[k] == N.poly_eval(f, a, f.gen * pow(w, k, f.p) % f.p)


@pytest.mark.parametrize("fname", list(N.FIELDS))
def test_witness_map_identity(fname):
    """h Z = A B - C as polynomials (checked at a point off the domain), deg h <= n - 2."""
    f = N.FIELDS[fname]
    n = 64
    rng = O.SplitMix64(7)
    a = [rng.below(f.p) for _ in range(n)]
    b = [rng.below(f.p) for _ in range(n)]
    c = [x * y % f.p for x, y in zip(a, b)]                   # a satisfied system: c_i = a_i b_i on the domain
    h = N.witness_map(f, a, b, c)
    assert h[n - 1] == 0
    pa, pb, pc = N.ifft(f, a), N.ifft(f, b), N.ifft(f, c)
    tau = rng.below(f.p)
    z = (pow(tau, n, f.p) - 1) % f.p
    lhs = N.poly_eval(f, h, tau) * z % f.p
    rhs = (N.poly_eval(f, pa, tau) * N.poly_eval(f, pb, tau) - N.poly_eval(f, pc, tau)) % f.p
    assert lhs == rhs
