"""Field layer (Fq377, Fq2 over it, Fq761) on the GPU vs Python big-int arithmetic."""
import numpy as np
import pytest

from oracle import cref as C
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _ref(L, op, a, b):
    c = L.curve
    if op == "add": return c.add(a, b)
    if op == "sub": return c.sub(a, b)
    if op == "mul": return c.mul(a, b)
    if op == "sqr": return c.mul(a, a)
    if op == "neg": return c.neg(a)
    if op == "dbl": return c.add(a, a)
    if op == "inv": return c.inv(a) if a != c.zero else c.zero
    raise ValueError(op)


@pytest.mark.parametrize("name", ["bls12_377_g1", "bls12_377_g2", "bw6_761_g1"])
def test_field_ops_match_python(name):
    import torch
    from celo_bls_snark_rs_b200 import engine as E
    E.init(0)
    L = C.LAYOUTS[name]
    rng = O.SplitMix64(99)
    m = L.modulus
    deg = L.curve.ext_degree
    n = 64

    def rand_el(i):
        special = [0, 1, m - 1, m - 2, 2, (m - 1) // 2]
        def one():
            return special[i % len(special)] if i < 12 else rng.below(m)
        return one() if deg == 1 else (one(), rng.below(m) if i % 3 else 0)

    a = [rand_el(i) for i in range(n)]
    b = [rand_el(n - 1 - i) for i in range(n)]
    dev = torch.device("cuda:0")
    enc = lambda xs: torch.from_numpy(np.frombuffer(b"".join(L.fe_to_mont_bytes(x) for x in xs), dtype=np.uint8).copy()).to(dev)
    d_a, d_b = enc(a), enc(b)
    d_o = torch.empty_like(d_a)
    cb = L.coord_bytes
    # coop_*: the same operations through the warp-cooperative routines (one limb per lane, csrc/coop.cuh)
    for op in ("add", "sub", "mul", "sqr", "neg", "dbl", "inv", "coop_add", "coop_sub", "coop_mul", "coop_sqr", "coop_neg",
               "coop_dbl"):
        d_o.zero_()
        E.field_op_device(L.id, op, d_a.data_ptr(), d_b.data_ptr(), n, d_o.data_ptr())
        E.sync()
        raw = d_o.cpu().numpy().tobytes()
        got = [L.fe_from_mont_bytes(raw[i * cb:(i + 1) * cb]) for i in range(n)]
        want = [_ref(L, op.replace("coop_", ""), x, y) for x, y in zip(a, b)]
        assert got == want, (name, op, [i for i in range(n) if got[i] != want[i]][:5])
        # outputs must be canonical residues (fully reduced), i.e. the raw Montgomery words < p
        words = [int.from_bytes(raw[i * L.curve.coord_bytes:(i + 1) * L.curve.coord_bytes], "little")
                 for i in range(n * deg)]
        assert all(w < m for w in words), (name, op)
