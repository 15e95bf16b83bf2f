"""Shared input builders for the parity tests (CPU oracle side; TEST INFRASTRUCTURE ONLY)."""
import numpy as np

from . import cref as C
from . import oracle as O

GENS = {"bls12_377_g1": O.G1_GEN, "bls12_377_g2": O.G2_GEN}


def generator(name, rng):
    if name in GENS:
        return GENS[name]
    curve = O.CURVES[name]
    while True:                                   # BW6-761: random curve point
        x = rng.below(O.Q761)
        y = O.sqrt_mod((x * x * x + curve.b) % O.Q761, O.Q761)
        if y is not None:
            return (x, y)


def random_points(name, n, seed, distinct=64):
    """n oracle points: `distinct` random multiples of a generator, tiled."""
    L = C.LAYOUTS[name]
    rng = O.SplitMix64(seed)
    g = generator(name, rng)
    k = min(n, distinct)
    packed = C.fixed_base_batch(L, g, [rng.below(L.curve.scalar_mod) for _ in range(k)])
    pts = L.affine_from_records(packed)
    return [pts[i % k] for i in range(n)] if k else []


def random_scalars_array(L, n, seed):
    """uniform scalars < modulus as a uint64 [n, limbs] array (numpy, fast)."""
    rng = np.random.default_rng(seed)
    bits = L.curve.scalar_bits
    arr = rng.integers(0, 1 << 63, size=(n, L.scalar_limbs), dtype=np.uint64) * np.uint64(2) + \
        rng.integers(0, 2, size=(n, L.scalar_limbs), dtype=np.uint64)
    top_bits = bits - 64 * (L.scalar_limbs - 1) - 1           # stay below the modulus: clear the top bit
    arr[:, -1] &= np.uint64((1 << top_bits) - 1)
    return arr


def edge_case_inputs(name, n, seed):
    """points/scalars with the cases SURVEY.md section 7 step 2 lists."""
    L = C.LAYOUTS[name]
    rng = O.SplitMix64(seed)
    pts = random_points(name, n, seed)
    scalars = [rng.below(L.curve.scalar_mod) for _ in range(n)]
    if n >= 8:
        scalars[0] = 0
        scalars[1] = 1
        scalars[2] = L.curve.scalar_mod - 1
        pts[4], scalars[4] = pts[3], scalars[3]                  # P + P in every bucket it touches
        pts[6], scalars[6] = L.curve.pneg(pts[5]), scalars[5]    # P + (-P)
        pts[7] = None                                            # infinity base
    return pts, scalars


def signature_batch(n, seed, corrupt=None):
    """n (pk_i, H_i) pairs plus the aggregate signature, as in Signature::batch_verify_hashes
    (crates/bls-crypto/src/bls/signature.rs:125-155): pk_i = sk_i * g2, H_i = h_i * g1,
    sigma = sum sk_i * H_i.  Returns the (G1 list, G2 list) of the N + 1 pairs, first pair
    (sigma, -g2).  corrupt = index of a message hash to replace (verification must then fail)."""
    L1, L2 = C.LAYOUTS["bls12_377_g1"], C.LAYOUTS["bls12_377_g2"]
    rng = O.SplitMix64(seed)
    sks = [rng.below(O.R - 1) + 1 for _ in range(n)]
    hs = [rng.below(O.R - 1) + 1 for _ in range(n)]
    pks = L2.affine_from_records(C.fixed_base_batch(L2, O.G2_GEN, sks))
    if corrupt is not None:
        hs_used = list(hs)
        hs_used[corrupt] = (hs[corrupt] + 1) % O.R or 1
    else:
        hs_used = hs
    hashes = L1.affine_from_records(C.fixed_base_batch(L1, O.G1_GEN, hs_used))
    sig_scalar = sum(s * h for s, h in zip(sks, hs)) % O.R
    sigma = L1.jacobian_to_affine(C.scalar_mul(L1, O.G1_GEN, sig_scalar))
    g1s = [sigma] + hashes
    g2s = [O.G2.pneg(O.G2_GEN)] + pks
    return g1s, g2s
