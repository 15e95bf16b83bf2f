"""CUDA radix-2 NTT / Groth16 witness map (through the C-ABI) against the oracle restatement of
ark-poly's Radix2EvaluationDomain and ark-groth16's witness_map (oracle/ntt.py).  Bit-exact: field
elements are unique, so the arkworks Montgomery images must be identical."""
import numpy as np
import pytest

from oracle import ntt as N
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from celo_bls_snark_rs_b200 import engine as E
    E.init(0)
    return E


def _dev(arr):
    import torch
    return torch.from_numpy(arr.view(np.int64).copy()).to("cuda:0")


def _host(t, limbs):
    return t.cpu().numpy().view(np.uint64).reshape(-1, limbs)


def _rand(f, n, seed):
    rng = O.SplitMix64(seed)
    return [rng.below(f.p) for _ in range(n)]


@pytest.mark.parametrize("fname", list(N.FIELDS))
@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 8, 9, 10, 11, 13])
def test_four_transforms_match_oracle(eng, fname, log_n):
    f = N.FIELDS[fname]
    a = _rand(f, 1 << log_n, 50 + log_n)
    a[0] = 0
    a[-1] = f.p - 1
    img = f.to_mont_array(a)
    for inverse, coset, ref in ((0, 0, N.fft), (1, 0, N.ifft), (0, 1, N.coset_fft), (1, 1, N.coset_ifft)):
        d = _dev(img)
        eng.ntt_device(f.id, d.data_ptr(), log_n, bool(inverse), bool(coset))
        eng.sync()
        assert np.array_equal(_host(d, f.limbs), f.to_mont_array(ref(f, a))), (fname, log_n, inverse, coset)


@pytest.mark.parametrize("fname", list(N.FIELDS))
@pytest.mark.parametrize("log_n", [1, 4, 10, 12])
def test_witness_map_matches_oracle(eng, fname, log_n):
    import torch
    f = N.FIELDS[fname]
    n = 1 << log_n
    a, b = _rand(f, n, 1), _rand(f, n, 2)
    c = [x * y % f.p for x, y in zip(a, b)]
    for i in range(n - max(1, n // 8), n):          # instance rows and padding: b = c = 0 there
        b[i] = c[i] = 0
    want = f.to_mont_array(N.witness_map(f, a, b, c))
    da, db, dc = (_dev(f.to_mont_array(v)) for v in (a, b, c))
    dh = torch.zeros_like(da)
    eng.witness_map_device(f.id, da.data_ptr(), db.data_ptr(), dc.data_ptr(), log_n, dh.data_ptr())
    eng.sync()
    assert np.array_equal(_host(dh, f.limbs), want)


@pytest.mark.parametrize("fname,log_n", [("fr_bls12_377", 20), ("fr_bw6_761", 20), ("fr_bw6_761", 22)])
def test_large_transform_properties(eng, fname, log_n):
    """Sizes the Python oracle cannot reach: round trip, and single outputs of the forward transform
    against a direct evaluation of the polynomial (Horner on the CPU)."""
    import torch
    f = N.FIELDS[fname]
    n = 1 << log_n
    rng = np.random.default_rng(log_n)
    raw = rng.integers(0, 1 << 62, size=(n, f.limbs), dtype=np.uint64)      # arbitrary residues < p (top limb < 2^62 > p's? no:
    raw[:, -1] &= np.uint64((1 << (f.p.bit_length() - 64 * (f.limbs - 1) - 1)) - 1)   # clear the top bits: value < p
    d = _dev(raw)
    orig = d.clone()
    eng.ntt_device(f.id, d.data_ptr(), log_n, False, False)
    eng.sync()
    fwd = _host(d, f.limbs).copy()
    coeffs = f.from_mont_array(raw)
    w = f.root_of_unity(log_n)
    for k in (0, 1, n // 2 + 3, n - 1):
        assert f.from_mont_array(fwd[k:k + 1])[0] == N.poly_eval(f, coeffs, pow(w, k, f.p))
    eng.ntt_device(f.id, d.data_ptr(), log_n, True, False)
    eng.sync()
    assert torch.equal(d, orig)
    eng.ntt_device(f.id, d.data_ptr(), log_n, False, True)
    eng.ntt_device(f.id, d.data_ptr(), log_n, True, True)
    eng.sync()
    assert torch.equal(d, orig)


def test_ntt_argument_errors(eng):
    lib = eng.load()
    assert lib.b200_ntt_device(7, None, 4, 0, 0, None) == 1
    assert lib.b200_ntt_device(0, None, 4, 0, 0, None) == 1
    assert lib.b200_ntt_device(0, 1, 40, 0, 0, None) == 1
