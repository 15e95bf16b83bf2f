"""CPU restatement of the BW6-761 pairing as the device computes it.
TEST INFRASTRUCTURE ONLY (see oracle/oracle.py header).

Reference call site: crates/epoch-snark/src/api/verifier.rs:35 (`verify_proof`, ark-groth16 0.1.0) ->
`BW6_761::product_of_pairings`.  Upstream (ark-ec 0.1.0 models/bw6, eprint 2020/351 algorithm 5) computes

    e(P, Q) = ( f_{u+1,Q}(P) * f_{u^3-u^2-u,Q}(P)^q ) ^ ((q^6 - 1) / r)

(the optimal ate pairing: (u + 1) + q (u^3 - u^2 - u) = 0 mod r, the second loop in non-adjacent form).
This file restates that pairing with the data layout of the CUDA kernels (csrc/pairing_bw6.cuh):

  * F_q^6 in the power basis F_q[w] / (w^6 + 4); arkworks' tower element (c0 + c1 v), c_i = a + b u + c u^2,
    u^3 = -4, v^2 = u, is sum_k a_k w^k with a_{2j} = c0[j], a_{2j+1} = c1[j];
  * G2 on the M-twist E': y^2 = x^3 + 4 over F_q, untwisted by (x, y) -> (x / w^2, y / w^3);
  * the running point T in Jacobian coordinates, lines scaled by elements of F_q and by w^3 (both vanish in
    the final exponentiation): a line is (c0, c2, c3), the coefficients of 1, w^2, w^3;
  * final exponentiation = easy part (q^3 - 1)(q + 1), then the hard part raised to HARD_MULTIPLE = c (q^2 - q + 1) / r
    (see below; here by plain square-and-multiply, on the device as f^R0 (f^q)^R1).

PINNING.  The value is pinned three ways in tests/test_oracle_bw6_ate.py: bilinearity and non-degeneracy on
the reference's own verifying key, agreement of the Groth16 verification boolean with the independent reduced
Tate pairing of oracle/bw6_verify.py, and the reference's known-answer test
(crates/bls-snark-sys/src/snark/mod.rs:52-119, expected `true`).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

from . import oracle as O

Q = O.Q761
R = O.R761
U = O.X                                   # BLS12-377 seed; BW6-761 is parameterised by the same u
LOOP_1 = U + 1
LOOP_2 = U ** 3 - U ** 2 - U
assert (LOOP_1 + Q * LOOP_2) % R == 0

HARD_EXP = (Q * Q - Q + 1) // R
assert (Q * Q - Q + 1) % R == 0
# The device raises to a multiple of the hard part that splits into two short polynomials of the seed:
# R0(u) + q R1(u) = c (q^2 - q + 1) / r, c a 191-bit integer prime to r -- still a non-degenerate bilinear pairing
# (e^c, e the pairing with the plain exponent), and f^R0 (f^q)^R1 is one joint square-and-multiply of 575 steps.
HARD_R0 = -103 * U**7 + 70 * U**6 + 269 * U**5 - 197 * U**4 - 314 * U**3 - 73 * U**2 - 263 * U - 220
HARD_R1 = 103 * U**9 - 276 * U**8 + 77 * U**7 + 492 * U**6 - 445 * U**5 - 65 * U**4 + 452 * U**3 - 181 * U**2 + 34 * U + 229
HARD_MULTIPLE = HARD_R0 + Q * HARD_R1
assert HARD_MULTIPLE % HARD_EXP == 0 and HARD_MULTIPLE > 0
HARD_COFACTOR = HARD_MULTIPLE // HARD_EXP
assert HARD_COFACTOR % R != 0 and HARD_COFACTOR.bit_length() == 191

GAMMA = pow(-4 % Q, (Q - 1) // 6, Q)      # w^q = GAMMA * w
assert pow(GAMMA, 3, Q) == Q - 1           # -4 is a non-residue: w^(q^3) = -w


def naf(k: int) -> List[int]:
    """Non-adjacent form, least significant digit first."""
    out = []
    while k:
        if k & 1:
            d = 2 - (k & 3)
            k -= d
        else:
            d = 0
        out.append(d)
        k >>= 1
    return out


LOOP_1_DIGITS = [int(b) for b in bin(LOOP_1)[2:]][::-1]          # plain bits (upstream ATE_LOOP_COUNT_1)
LOOP_2_DIGITS = naf(LOOP_2)                                         # signed digits (upstream ATE_LOOP_COUNT_2)


# ---- F_q^6, power basis ------------------------------------------------------------------------------------

def f6_one():
    return [1, 0, 0, 0, 0, 0]


def f6_mul(a, b):
    t = [0] * 11
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                t[i + j] += x * y
    return [(t[k] - 4 * (t[k + 6] if k < 5 else 0)) % Q for k in range(6)]


def f6_conj(a):
    """a^(q^3): w -> -w."""
    return [(-x) % Q if k & 1 else x for k, x in enumerate(a)]


def f6_frob(a, j=1):
    g = pow(GAMMA, j, Q)
    return [x * pow(g, k, Q) % Q for k, x in enumerate(a)]


def f6_inv(a):
    """Norm to F_q^3 (even powers), norm to F_q, one base-field inversion."""
    c = f6_conj(a)
    n = f6_mul(a, c)                       # in F_q^3 = F_q[w^2]
    assert n[1] == n[3] == n[5] == 0
    m = f6_mul(f6_frob(n, 1), f6_frob(n, 2))
    big = f6_mul(n, m)                     # in F_q
    assert big[1:] == [0, 0, 0, 0, 0]
    s = pow(big[0], -1, Q)
    return [x * s % Q for x in f6_mul(c, m)]


def f6_pow(a, e):
    r = f6_one()
    for bit in bin(e)[2:]:
        r = f6_mul(r, r)
        if bit == "1":
            r = f6_mul(r, a)
    return r


# ---- Miller loops ------------------------------------------------------------------------------------------

def _dbl_step(t, xp, yp):
    """T <- 2T on E' (a = 0, dbl-2009-l); tangent line at T evaluated at P, scaled by 2 Y Z^3 w^3."""
    x, y, z = t
    a = x * x % Q
    b = y * y % Q
    zz = z * z % Q
    yz = y * z % Q
    c = b * b % Q
    s = (x + b) * (x + b) % Q
    e = 3 * a % Q
    f = e * e % Q
    d = 2 * (s - a - c) % Q
    z3 = 2 * yz % Q
    x3 = (f - 2 * d) % Q
    y3 = (e * (d - x3) - 8 * c) % Q
    ezz = e * zz % Q
    ex = e * x % Q
    z3zz = z3 * zz % Q
    line = ((ex - 2 * b) % Q, (-ezz * xp) % Q, z3zz * yp % Q)
    return (x3, y3, z3), line


def _add_step(t, q2, xp, yp):
    """T <- T + Q2 (madd-2007-bl without the doubling factors); chord through T and Q2 at P, scaled by Z3 w^3."""
    x, y, z = t
    x2, y2 = q2
    zz = z * z % Q
    u2 = x2 * zz % Q
    s2 = y2 * z % Q * zz % Q
    h = (u2 - x) % Q
    r = (s2 - y) % Q
    hh = h * h % Q
    hhh = h * hh % Q
    v = x * hh % Q
    x3 = (r * r - hhh - 2 * v) % Q
    y3 = (r * (v - x3) - y * hhh) % Q
    z3 = z * h % Q
    line = ((r * x2 - y2 * z3) % Q, (-r * xp) % Q, z3 * yp % Q)
    return (x3, y3, z3), line


def _mul_line(f, line):
    c0, c2, c3 = line
    return f6_mul(f, [c0, 0, c2, c3, 0, 0])


def miller_sub_loop(p, q2, digits: Sequence[int]):
    """f_{k,Q}(P) for k given by its (signed) digits, least significant first; top digit must be 1."""
    xp, yp = p
    t = (q2[0], q2[1], 1)
    neg = (q2[0], (-q2[1]) % Q)
    f = f6_one()
    for d in reversed(digits[:-1]):
        f = f6_mul(f, f)
        t, line = _dbl_step(t, xp, yp)
        f = _mul_line(f, line)
        if d:
            t, line = _add_step(t, q2 if d > 0 else neg, xp, yp)
            f = _mul_line(f, line)
    return f


def miller_loop(p: Optional[Tuple[int, int]], q2: Optional[Tuple[int, int]]):
    if p is None or q2 is None:
        return f6_one()
    f1 = miller_sub_loop(p, q2, LOOP_1_DIGITS)
    f2 = miller_sub_loop(p, q2, LOOP_2_DIGITS)
    return f6_mul(f1, f6_frob(f2, 1))


def final_exponentiation(f):
    r = f6_mul(f6_conj(f), f6_inv(f))          # f^(q^3 - 1)
    r = f6_mul(f6_frob(r, 1), r)               # ^(q + 1)
    return f6_pow(r, HARD_MULTIPLE)


def product_of_pairings(pairs):
    f = f6_one()
    for p, q2 in pairs:
        f = f6_mul(f, miller_loop(p, q2))
    return final_exponentiation(f)


def verify_proof(vk, proof, inputs: Sequence[int]) -> bool:
    """ark-groth16 verify_proof with the check folded into one product:
    e(A, B) e(g_ic, -gamma) e(C, -delta) e(-alpha, beta) == 1."""
    a, b, c = proof
    if len(inputs) + 1 != len(vk["gamma_abc"]):
        return False
    g_ic = vk["gamma_abc"][0]
    for x, base in zip(inputs, vk["gamma_abc"][1:]):
        g_ic = O.BW6_G1.padd(g_ic, O.BW6_G1.pmul(base, x))
    out = product_of_pairings([(a, b), (g_ic, O.BW6_G2.pneg(vk["gamma"])), (c, O.BW6_G2.pneg(vk["delta"])),
                               (O.BW6_G1.pneg(vk["alpha"]), vk["beta"])])
    return out == f6_one()
