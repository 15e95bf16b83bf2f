"""Lane-level model of the warp-cooperative Montgomery product of csrc/coop.cuh (limb i of every operand lives
in lane i of a group of N lanes; shuffles are list indexing, ballots are bit masks).  The CUDA code follows this
file statement by statement; tests/test_coop_model.py checks the model against plain integer arithmetic, so an
algorithmic slip (column assignment, carry look-ahead, the low-half product of the reduction) shows up on the
CPU, before a GPU round trip.

    product  T = a * b            lane i owns columns i and N + i: exactly N partial products per lane
    quotient m = T_low * q mod R  (q = -p^-1 mod R, R = 2^(32 N)): low product, lane i owns column i
    result   (T + m * p) / R      same column split as the first product, then one conditional subtraction
Column sums are 96-bit (three words); `normalize` turns them into limbs: neighbour words by shuffle, then the
remaining one-bit carries by a generate / propagate look-ahead on ballot masks.
"""
M32 = (1 << 32) - 1


def limbs_of(v, n):
    return [(v >> (32 * i)) & M32 for i in range(n)]


def value_of(l):
    return sum(x << (32 * i) for i, x in enumerate(l))


def lookahead(gen_mask, prop_mask):
    """carry-in mask: bit c = carry entering position c (cout[c] = g[c] | p[c] & cin[c])"""
    x, y = gen_mask | prop_mask, gen_mask
    return (x + y) ^ x ^ y


def normalize(lo_cols, hi_cols, n, low_only=False):
    """lo_cols[i] / hi_cols[i]: 96-bit column sums of columns i and n + i, held by lane i -> limbs (low, high)."""
    w = lambda v, k: (v >> (32 * k)) & M32
    lo, hi = [0] * n, [0] * n
    klo, khi = [0] * n, [0] * n
    for i in range(n):                                  # every lane, in parallel
        s1, s2 = (i - 1) % n, (i - 2) % n               # shuffle sources
        # low column i: words of columns i - 1, i - 2 (nothing below column 0)
        y = w(lo_cols[i], 0) + (w(lo_cols[s1], 1) if i >= 1 else 0) + (w(lo_cols[s2], 2) if i >= 2 else 0)
        lo[i], klo[i] = y & M32, y >> 32
        # high column n + i: neighbours are high columns, or the top low columns for i < 2
        y = w(hi_cols[i], 0) + (w(hi_cols[s1], 1) if i >= 1 else w(lo_cols[s1], 1)) + \
            (w(hi_cols[s2], 2) if i >= 2 else w(lo_cols[s2], 2))
        hi[i], khi[i] = y & M32, y >> 32
    glo = ghi = plo = phi = 0
    xlo, xhi = [0] * n, [0] * n
    for i in range(n):
        s1 = (i - 1) % n
        y = lo[i] + (klo[s1] if i >= 1 else 0)
        xlo[i] = y & M32
        glo |= (y >> 32) << i
        plo |= (xlo[i] == M32) << i
        y = hi[i] + (khi[s1] if i >= 1 else klo[s1])
        xhi[i] = y & M32
        ghi |= (y >> 32) << i
        phi |= (xhi[i] == M32) << i
    if low_only:
        cin = lookahead(glo, plo)
        return [(xlo[i] + ((cin >> i) & 1)) & M32 for i in range(n)], None
    cin = lookahead(glo | (ghi << n), plo | (phi << n))
    return ([(xlo[i] + ((cin >> i) & 1)) & M32 for i in range(n)],
            [(xhi[i] + ((cin >> (n + i)) & 1)) & M32 for i in range(n)])


def cond_sub(r, p, n):
    """r in [0, 2p) -> [0, p): lane-parallel r - p with a borrow look-ahead, kept when no borrow leaves the top"""
    gen = prop = 0
    d = [0] * n
    for i in range(n):
        d[i] = (r[i] - p[i]) & M32
        gen |= (r[i] < p[i]) << i
        prop |= (r[i] == p[i]) << i
    x, y = gen | prop, gen
    s = x + y
    bin_mask = s ^ x ^ y
    borrow_out = (s >> n) & 1
    return list(r) if borrow_out else [(d[i] - ((bin_mask >> i) & 1)) & M32 for i in range(n)]


def coop_mul(a, b, p, q, n):
    """a, b, p, q: lists of n limbs (limb i in lane i).  Returns the Montgomery product a * b / R mod p."""
    lo, hi = [0] * n, [0] * n
    for j in range(n):                                  # n steps; every lane does one product per step
        for i in range(n):
            aj, bk = a[j], b[(i - j) % n]               # two shuffles: broadcast of a_j, rotation of b
            if j <= i:
                lo[i] += aj * bk
            else:
                hi[i] += aj * bk
    t_lo, t_hi = normalize(lo, hi, n)
    mc = [0] * n
    for j in range(n):
        for i in range(n):
            if j <= i:
                mc[i] += t_lo[j] * q[(i - j) % n]
    m, _ = normalize(mc, [0] * n, n, low_only=True)
    lo, hi = list(t_lo), list(t_hi)                     # T enters the column sums of m * p
    for j in range(n):
        for i in range(n):
            mj, pk = m[j], p[(i - j) % n]
            if j <= i:
                lo[i] += mj * pk
            else:
                hi[i] += mj * pk
    u_lo, u_hi = normalize(lo, hi, n)
    assert all(x == 0 for x in u_lo), "low half of T + m p must vanish"
    return cond_sub(u_hi, p, n)


def coop_dot(pairs, p, q, n):
    """sum of products sharing ONE reduction (csrc/coop.cuh CoopOps::accumulate x K, then CoopOps::reduce): the column
    sums of every a * b go into the same accumulators.  Needs 96-bit columns (K n products of 64 bits per lane) and
    K p < R; returns sum(a * b) / R mod p."""
    lo, hi = [0] * n, [0] * n
    for a, b in pairs:
        for j in range(n):
            for i in range(n):
                if j <= i:
                    lo[i] += a[j] * b[(i - j) % n]
                else:
                    hi[i] += a[j] * b[(i - j) % n]
    assert all(x < (1 << 96) for x in lo + hi), "column accumulators are 96 bits wide"
    t_lo, t_hi = normalize(lo, hi, n)
    mc = [0] * n
    for j in range(n):
        for i in range(n):
            if j <= i:
                mc[i] += t_lo[j] * q[(i - j) % n]
    m, _ = normalize(mc, [0] * n, n, low_only=True)
    lo, hi = list(t_lo), list(t_hi)
    for j in range(n):
        for i in range(n):
            if j <= i:
                lo[i] += m[j] * p[(i - j) % n]
            else:
                hi[i] += m[j] * p[(i - j) % n]
    u_lo, u_hi = normalize(lo, hi, n)
    assert all(x == 0 for x in u_lo), "low half of T + m p must vanish"
    return cond_sub(u_hi, p, n)


def coop_add(a, b, p, n):
    gen = prop = 0
    s = [0] * n
    for i in range(n):
        y = a[i] + b[i]
        s[i] = y & M32
        gen |= (y >> 32) << i
        prop |= (s[i] == M32) << i
    cin = lookahead(gen, prop)
    return cond_sub([(s[i] + ((cin >> i) & 1)) & M32 for i in range(n)], p, n)


def coop_sub(a, b, p, n):
    gen = prop = 0
    d = [0] * n
    for i in range(n):
        d[i] = (a[i] - b[i]) & M32
        gen |= (a[i] < b[i]) << i
        prop |= (a[i] == b[i]) << i
    x, y = gen | prop, gen
    s = x + y
    bin_mask, borrow_out = s ^ x ^ y, (s >> n) & 1
    d = [(d[i] - ((bin_mask >> i) & 1)) & M32 for i in range(n)]
    if not borrow_out:
        return d
    gen = prop = 0                                       # a < b: add p back (the carry out of the top is dropped)
    s2 = [0] * n
    for i in range(n):
        y = d[i] + p[i]
        s2[i] = y & M32
        gen |= (y >> 32) << i
        prop |= (s2[i] == M32) << i
    cin = lookahead(gen, prop)
    return [(s2[i] + ((cin >> i) & 1)) & M32 for i in range(n)]
