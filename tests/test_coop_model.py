"""The lane-level model of the warp-cooperative field arithmetic (tools/coop_model.py, mirrored by csrc/coop.cuh)
against plain integers, for the three moduli the kernels use."""
import random

import pytest

from tools import coop_model as M

P377 = 0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001
Q761 = 0x0122E824FB83CE0AD187C94004FAFF3EB926186A81D14688528275EF8087BE41707BA638E584E91903CEBAFF25B423048689C8ED12F9FD9071DCD3DC73EBFF2E98A116C25667A8F8160CF8AEEAF0A437E6913E6870000082F49D00000000008B


@pytest.mark.parametrize("p,n", [(P377, 12), (Q761, 24)])
def test_coop_field_ops_match_integers(p, n):
    R = 1 << (32 * n)
    q = (-pow(p, -1, R)) % R
    pl, ql = M.limbs_of(p, n), M.limbs_of(q, n)
    rng = random.Random(5 + n)
    edge = [0, 1, p - 1, p - 2, (1 << 32) - 1, ((1 << 64) - 1) << 32, R % p, (p - 1) // 2, ((1 << (p.bit_length() - 1)) - 1)]
    vals = edge + [rng.randrange(p) for _ in range(40)]
    rinv = pow(R, -1, p)
    for a in vals:
        for b in rng.sample(vals, 6) + [a, 0, p - 1]:
            al, bl = M.limbs_of(a, n), M.limbs_of(b, n)
            assert M.value_of(M.coop_mul(al, bl, pl, ql, n)) == a * b * rinv % p
            assert M.value_of(M.coop_add(al, bl, pl, n)) == (a + b) % p
            assert M.value_of(M.coop_sub(al, bl, pl, n)) == (a - b) % p


@pytest.mark.parametrize("p,n", [(P377, 12), (Q761, 24)])
def test_coop_dot_products_share_one_reduction(p, n):
    """CoopOps::accumulate x K + CoopOps::reduce (the six-term coefficient sums of pairing_bw6_coop.cuh): worst-case
    operands (p - 1 everywhere) stay inside the 96-bit columns and one conditional subtraction lands in [0, p)."""
    R = 1 << (32 * n)
    q = (-pow(p, -1, R)) % R
    pl, ql = M.limbs_of(p, n), M.limbs_of(q, n)
    rinv = pow(R, -1, p)
    rng = random.Random(11 + n)
    for k in (1, 2, 3, 6, 12):
        for trial in range(6):
            vals = [(p - 1, p - 1)] * k if trial == 0 else [(rng.randrange(p), rng.randrange(p)) for _ in range(k)]
            got = M.coop_dot([(M.limbs_of(a, n), M.limbs_of(b, n)) for a, b in vals], pl, ql, n)
            assert M.value_of(got) == sum(a * b for a, b in vals) * rinv % p
