// Short-Weierstrass group law (a = 0: BLS12-377 G1/G2, BW6-761 G1/G2) for sm_100a.
//
// Device replacement for ark-ec 0.1.0 short_weierstrass_jacobian as used by
// VariableBaseMSM (crates/bls-crypto/src/bls/signature.rs:85, public.rs:61).
// Buckets are kept in extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ,
// ZZ^3 = ZZZ^2): the mixed addition is 8M + 2S against 7M + 4S for Jacobian and has
// no field doubling chains.  The tail (window combine) and everything that crosses
// the C-ABI is plain Jacobian (X/Z^2, Y/Z^3), arkworks' GroupProjective.
// All exceptional cases (infinity operands, P + P, P + (-P)) are handled exactly:
// duplicate bases are the common case for this caller, not a corner case
// (crates/epoch-snark/src/api/prover.rs:154-157 pads with the generator).
#pragma once
#include "fp.cuh"

namespace b200 {

// everything except the mixed add of the accumulate kernel is off the hot path: out of line
#define B200_COLD __device__ __noinline__

// ---------------- Fq2 = Fq[u] / (u^2 + 5) over BLS12-377 Fq -------------------------
template <class B>
struct Fp2 {
    B c0, c1;
    struct alignas(16) Mem {
        typename B::Mem c0, c1;
    };
    B200_DEV static Fp2 from_ark(const Mem &m) { return {B::from_ark(m.c0), B::from_ark(m.c1)}; }
    B200_DEV Mem to_ark() const { return {c0.to_ark(), c1.to_ark()}; }
    B200_DEV static Fp2 load(const Mem &m) { return {B::load(m.c0), B::load(m.c1)}; }
    B200_DEV Mem store() const { return {c0.store(), c1.store()}; }
    B200_DEV static Fp2 zero() { return {B::zero(), B::zero()}; }
    B200_DEV static Fp2 one() { return {B::one(), B::zero()}; }
    B200_DEV bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    B200_DEV bool operator==(const Fp2 &o) const { return c0 == o.c0 && c1 == o.c1; }
    B200_DEV friend Fp2 operator+(const Fp2 &a, const Fp2 &b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
    B200_DEV friend Fp2 operator-(const Fp2 &a, const Fp2 &b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
    B200_DEV Fp2 dbl() const { return {c0.dbl(), c1.dbl()}; }
    B200_DEV Fp2 neg() const { return {c0.neg(), c1.neg()}; }
    B200_DEV Fp2 cneg(bool f) const { return {c0.cneg(f), c1.cneg(f)}; }
    B200_DEV Fp2 shfl(unsigned mask, int src_lane) const { return {c0.shfl(mask, src_lane), c1.shfl(mask, src_lane)}; }
    B200_DEV static Fp2 sel4(int q, const Fp2 &a0, const Fp2 &a1, const Fp2 &a2, const Fp2 &a3) {
        return {B::sel4(q, a0.c0, a1.c0, a2.c0, a3.c0), B::sel4(q, a0.c1, a1.c1, a2.c1, a3.c1)};
    }
    B200_DEV static B mul5(const B &a) {
        B t = a.dbl().dbl();
        return t + a;
    }
    // Karatsuba: 3 base multiplications; u^2 = -5.  Out of line (one shared body of ~1.2k
    // instructions per product, like the 24-limb field) to bound code size.
    __device__ __noinline__ static Fp2 mul_outline(Fp2 a, Fp2 b) {
        B v0 = a.c0 * b.c0;
        B v1 = a.c1 * b.c1;
        B t = (a.c0 + a.c1) * (b.c0 + b.c1);
        return {v0 - mul5(v1), t - v0 - v1};
    }
    // (a0 + a1 u)^2 = a0^2 - 5 a1^2 + 2 a0 a1 u
    __device__ __noinline__ static Fp2 sqr_outline(Fp2 a) {
        B v0 = a.c0.sqr();
        B v1 = a.c1.sqr();
        B t = a.c0 * a.c1;
        return {v0 - mul5(v1), t.dbl()};
    }
    // (arguments by value: see the note at Fp::mul_outline)
    B200_DEV friend Fp2 operator*(const Fp2 &a, const Fp2 &b) { return mul_outline(a, b); }
    B200_DEV Fp2 sqr() const { return sqr_outline(*this); }
};

// ---------------- points --------------------------------------------------------------
// *Mem types are the memory images (packed 32-bit words); the plain types live in registers.
// "ark" conversions change the Montgomery radix at the C-ABI boundary, load/store keep the
// engine's native radix (buckets, partial sums).
template <class F>
struct alignas(16) AffineMem {
    typename F::Mem x, y;
};
template <class F>
struct alignas(16) JacobianMem {
    typename F::Mem x, y, z;
};
template <class F>
struct alignas(16) XYZZMem {
    typename F::Mem x, y, zz, zzz;
};

// 128-bit vector copy of a memory image (global -> registers)
template <class M>
B200_DEV M ldg_mem(const M *__restrict__ src) {
    static_assert(sizeof(M) % 16 == 0, "memory images are 16-byte multiples");
    M r;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(M) / 16); k++) d4[k] = __ldg(s4 + k);
    return r;
}

template <class F>
struct Affine {                                     // finite point; (0, 0) encodes infinity
    F x, y;
    B200_DEV bool is_inf() const { return x.is_zero() && y.is_zero(); }
    B200_DEV static Affine load(const AffineMem<F> &m) { return {F::load(m.x), F::load(m.y)}; }
    B200_DEV static Affine from_ark(const AffineMem<F> &m) { return {F::from_ark(m.x), F::from_ark(m.y)}; }
    B200_DEV AffineMem<F> store() const { return {x.store(), y.store()}; }
    B200_DEV AffineMem<F> to_ark() const { return {x.to_ark(), y.to_ark()}; }
};

template <class F>
struct Jacobian {
    F x, y, z;
    B200_DEV static Jacobian inf() { return {F::one(), F::one(), F::zero()}; }
    B200_DEV bool is_inf() const { return z.is_zero(); }
    B200_DEV static Jacobian from_ark(const JacobianMem<F> &m) {
        return {F::from_ark(m.x), F::from_ark(m.y), F::from_ark(m.z)};
    }
    B200_DEV JacobianMem<F> to_ark() const { return {x.to_ark(), y.to_ark(), z.to_ark()}; }

    // Out-of-line bodies take and return points BY VALUE (see the note at Fp::mul_outline).
    // dbl-2009-l (2M + 5S)
    B200_COLD static Jacobian dbl_outline(Jacobian p) {
        if (p.is_inf()) return p;
        F a = p.x.sqr();
        F b = p.y.sqr();
        F c = b.sqr();
        F d = ((p.x + b).sqr() - a - c).dbl();
        F e = a.dbl() + a;
        F f = e.sqr();
        Jacobian r;
        r.z = (p.z * p.y).dbl();
        r.x = f - d.dbl();
        r.y = e * (d - r.x) - c.dbl().dbl().dbl();
        return r;
    }
    B200_DEV void dbl() { *this = dbl_outline(*this); }
    // add-2007-bl (11M + 5S)
    B200_COLD static Jacobian add_outline(Jacobian p, Jacobian q) {
        if (q.is_inf()) return p;
        if (p.is_inf()) return q;
        F z1z1 = p.z.sqr();
        F z2z2 = q.z.sqr();
        F u1 = p.x * z2z2;
        F u2 = q.x * z1z1;
        F s1 = p.y * q.z * z2z2;
        F s2 = q.y * p.z * z1z1;
        if (u1 == u2) {
            if (s1 == s2) return dbl_outline(p);
            return inf();
        }
        F h = u2 - u1;
        F i = h.dbl().sqr();
        F j = h * i;
        F r = (s2 - s1).dbl();
        F v = u1 * i;
        Jacobian o;
        o.x = r.sqr() - j - v.dbl();
        o.y = r * (v - o.x) - (s1 * j).dbl();
        o.z = ((p.z + q.z).sqr() - z1z1 - z2z2) * h;
        return o;
    }
    B200_DEV void add(const Jacobian &q) { *this = add_outline(*this, q); }
};

template <class F>
struct XYZZ {
    F x, y, zz, zzz;
    B200_DEV static XYZZ inf() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
    B200_DEV bool is_inf() const { return zz.is_zero(); }
    B200_DEV static XYZZ load(const XYZZMem<F> &m) {
        return {F::load(m.x), F::load(m.y), F::load(m.zz), F::load(m.zzz)};
    }
    B200_DEV XYZZMem<F> store() const { return {x.store(), y.store(), zz.store(), zzz.store()}; }

    // mdbl-2008-s-1: 2 * (affine point)
    B200_COLD static XYZZ dbl_affine(F px, F py) {
        F u = py.dbl();
        F v = u.sqr();
        F w = u * v;
        F s = px * v;
        F xx = px.sqr();
        F m = xx.dbl() + xx;
        XYZZ r;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * py;
        r.zz = v;
        r.zzz = w;
        return r;
    }
    // dbl-2008-s-1
    B200_COLD static XYZZ dbl_outline(XYZZ p) {
        if (p.is_inf()) return p;
        F u = p.y.dbl();
        F v = u.sqr();
        F w = u * v;
        F s = p.x * v;
        F xx = p.x.sqr();
        F m = xx.dbl() + xx;
        XYZZ r;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * p.y;
        r.zz = v * p.zz;
        r.zzz = w * p.zzz;
        return r;
    }
    B200_DEV void dbl() { *this = dbl_outline(*this); }
    // madd-2008-s (8M + 2S): this += (px, py), a finite affine point
    B200_DEV void madd(const F &px, const F &py) {
        if (is_inf()) {
            x = px;
            y = py;
            zz = F::one();
            zzz = F::one();
            return;
        }
        F p = px * zz - x;
        F r = py * zzz - y;
        if (p.is_zero()) {
            if (r.is_zero()) *this = dbl_affine(px, py);   // P + P
            else *this = inf();                            // P + (-P)
            return;
        }
        F pp = p.sqr();
        F ppp = p * pp;
        F q = x * pp;
        F x3 = r.sqr() - ppp - q.dbl();
        y = r * (q - x3) - y * ppp;
        x = x3;
        zz = zz * pp;
        zzz = zzz * ppp;
    }
    // add-2008-s (12M + 2S)
    B200_COLD static XYZZ add_outline(XYZZ a, XYZZ o) {
        if (o.is_inf()) return a;
        if (a.is_inf()) return o;
        F u1 = a.x * o.zz;
        F u2 = o.x * a.zz;
        F s1 = a.y * o.zzz;
        F s2 = o.y * a.zzz;
        F p = u2 - u1;
        F r = s2 - s1;
        if (p.is_zero()) {
            if (r.is_zero()) return dbl_outline(a);
            return inf();
        }
        F pp = p.sqr();
        F ppp = p * pp;
        F q = u1 * pp;
        XYZZ t;
        t.x = r.sqr() - ppp - q.dbl();
        t.y = r * (q - t.x) - s1 * ppp;
        t.zz = a.zz * o.zz * pp;
        t.zzz = a.zzz * o.zzz * ppp;
        return t;
    }
    B200_DEV void add(const XYZZ &o) { *this = add_outline(*this, o); }
    // (X, Y, ZZ, ZZZ) -> Jacobian (X*ZZ, Y*ZZZ, ZZ)
    B200_COLD static Jacobian<F> to_jacobian_outline(XYZZ p) {
        if (p.is_inf()) return Jacobian<F>::inf();
        return {p.x * p.zz, p.y * p.zzz, p.zz};
    }
    B200_DEV Jacobian<F> to_jacobian() const { return to_jacobian_outline(*this); }
};

// ---------------- quad-cooperative point operations ------------------------------------------
// The tail of the MSM (bucket reduction, window sums, Horner combine) is a few thousand
// point operations deep but only a few thousand threads wide: it is bound by the LATENCY of
// dependent field products (about 0.94 us each in one warp, profiles/r1_field_layer_notes.md),
// and most of the GPU idles.  Four consecutive lanes (a "quad") therefore share one point
// operation: all four hold identical copies of the operands, each computes ONE of the up to
// four independent products of a round, and the products are exchanged with shuffles.  An
// XYZZ addition drops from 14 sequential products to 4 rounds, a doubling from 9 to 3.
struct Quad {
    int q, base;
    unsigned mask;
    B200_DEV Quad() {
        int lane = threadIdx.x & 31;
        q = lane & 3;
        base = lane & ~3;
        mask = 0xFu << base;
    }
    // p_k = a_k * b_k for k = 0..3, every lane receives all four
    template <class F>
    B200_DEV void mul4(const F &a0, const F &b0, const F &a1, const F &b1, const F &a2, const F &b2, const F &a3,
                       const F &b3, F &p0, F &p1, F &p2, F &p3) const {
        F p = F::sel4(q, a0, a1, a2, a3) * F::sel4(q, b0, b1, b2, b3);
        p0 = p.shfl(mask, base);
        p1 = p.shfl(mask, base + 1);
        p2 = p.shfl(mask, base + 2);
        p3 = p.shfl(mask, base + 3);
    }
};

// Quad whose products go through the field's ONE out-of-line body (Fp::mul_outline; Fp2 products already do): for
// kernels with many warps in different phases of long loops, where seven inlined ~600-instruction products per point
// operation overflow the instruction cache (the many-small-MSMs kernel: 26 us per scalar bit either way until then).
template <class F> struct OutlineMul {
    B200_DEV static F mul(const F &a, const F &b) { return a * b; }
};
template <class P> struct OutlineMul<Fp<P>> {
    B200_DEV static Fp<P> mul(const Fp<P> &a, const Fp<P> &b) { return Fp<P>::mul_outline(a, b); }
};
struct QuadShared : Quad {
    template <class F>
    B200_DEV void mul4(const F &a0, const F &b0, const F &a1, const F &b1, const F &a2, const F &b2, const F &a3,
                       const F &b3, F &p0, F &p1, F &p2, F &p3) const {
        F p = OutlineMul<F>::mul(F::sel4(q, a0, a1, a2, a3), F::sel4(q, b0, b1, b2, b3));
        p0 = p.shfl(mask, base);
        p1 = p.shfl(mask, base + 1);
        p2 = p.shfl(mask, base + 2);
        p3 = p.shfl(mask, base + 3);
    }
};

// dbl-2008-s-1 in 3 rounds.  All lanes of the quad must call with identical arguments.
template <class F, class QT>
B200_DEV void quad_dbl(const QT &Q, XYZZ<F> &a) {
    if (a.is_inf()) return;
    F u = a.y.dbl();
    F v, xx, d0, d1;
    Q.mul4(u, u, a.x, a.x, u, u, a.x, a.x, v, xx, d0, d1);
    F m = xx.dbl() + xx;
    F w, s, mm, zz3;
    Q.mul4(u, v, a.x, v, m, m, v, a.zz, w, s, mm, zz3);
    F x3 = mm - s.dbl();
    F t1, t2, zzz3;
    Q.mul4(m, s - x3, w, a.y, w, a.zzz, w, a.zzz, t1, t2, zzz3, d0);
    a.x = x3;
    a.y = t1 - t2;
    a.zz = zz3;
    a.zzz = zzz3;
}

// add-2008-s in 4 rounds (exceptional cases exact, as in XYZZ::add_outline).
template <class F, class QT>
B200_DEV void quad_add(const QT &Q, XYZZ<F> &a, const XYZZ<F> &o) {
    if (o.is_inf()) return;
    if (a.is_inf()) {
        a = o;
        return;
    }
    F u1, u2, s1, s2;
    Q.mul4(a.x, o.zz, o.x, a.zz, a.y, o.zzz, o.y, a.zzz, u1, u2, s1, s2);
    F p = u2 - u1;
    F r = s2 - s1;
    if (p.is_zero()) {
        if (r.is_zero()) quad_dbl(Q, a);
        else a = XYZZ<F>::inf();
        return;
    }
    F pp, rr, zza, zzza;
    Q.mul4(p, p, r, r, a.zz, o.zz, a.zzz, o.zzz, pp, rr, zza, zzza);
    F ppp, qq, zz3, d0;
    Q.mul4(p, pp, u1, pp, zza, pp, zza, pp, ppp, qq, zz3, d0);
    F x3 = rr - ppp - qq.dbl();
    F t1, t2, zzz3;
    Q.mul4(r, qq - x3, s1, ppp, zzza, ppp, zzza, ppp, t1, t2, zzz3, d0);
    a.x = x3;
    a.y = t1 - t2;
    a.zz = zz3;
    a.zzz = zzz3;
}

}  // namespace b200
