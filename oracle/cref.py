"""ctypes front-end of the C oracle (oracle/cpu_ref.c) plus the byte-level codecs
shared by the tests: arkworks in-memory layouts (Montgomery, 64-bit LE limbs)
<-> Python integers.  TEST INFRASTRUCTURE ONLY (see oracle/oracle.py header).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

from . import oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcpu_ref.so")

CURVE_ID = {"bls12_377_g1": 0, "bls12_377_g2": 1, "bw6_761_g1": 2, "bw6_761_g2": 3}


class CurveLayout:
    """Byte layout of one group in arkworks memory."""

    def __init__(self, curve: O.Curve):
        self.curve = curve
        self.id = CURVE_ID[curve.name]
        self.fe_limbs = curve.coord_bytes // 8                  # limbs of one prime-field element
        self.coord_bytes = curve.coord_bytes * curve.ext_degree  # bytes of one coordinate
        self.packed_stride = 2 * self.coord_bytes
        self.ark_stride = 2 * self.coord_bytes + 8              # x | y | bool + padding
        self.jac_bytes = 3 * self.coord_bytes
        self.scalar_limbs = (curve.scalar_bits + 63) // 64
        self.modulus = curve.modulus
        self.mont_r = (1 << (64 * self.fe_limbs)) % self.modulus
        self.mont_rinv = pow(self.mont_r, -1, self.modulus)

    # --- field element codecs -------------------------------------------------
    def fe_to_mont_bytes(self, v) -> bytes:
        if self.curve.ext_degree == 1:
            return (v * self.mont_r % self.modulus).to_bytes(self.curve.coord_bytes, "little")
        return b"".join((c * self.mont_r % self.modulus).to_bytes(self.curve.coord_bytes, "little") for c in v)

    def fe_from_mont_bytes(self, bs: bytes):
        n = self.curve.coord_bytes
        vals = [int.from_bytes(bs[i * n:(i + 1) * n], "little") * self.mont_rinv % self.modulus
                for i in range(self.curve.ext_degree)]
        return vals[0] if self.curve.ext_degree == 1 else tuple(vals)

    # --- point codecs ----------------------------------------------------------
    def affine_records(self, points: Sequence, stride: Optional[int] = None) -> np.ndarray:
        """points -> uint8 array [n, stride] in arkworks GroupAffine layout
        (x | y | infinity flag); stride == packed_stride drops the flag and
        encodes infinity as (0, 0)."""
        stride = stride or self.ark_stride
        out = np.zeros((len(points), stride), dtype=np.uint8)
        cb = self.coord_bytes
        one = self.fe_to_mont_bytes(1 if self.curve.ext_degree == 1 else (1, 0))
        for i, pt in enumerate(points):
            if pt is None:
                if stride > 2 * cb:
                    out[i, cb:2 * cb] = np.frombuffer(one, dtype=np.uint8)   # arkworks zero() = (0, 1, inf)
                    out[i, 2 * cb] = 1
                continue
            out[i, :cb] = np.frombuffer(self.fe_to_mont_bytes(pt[0]), dtype=np.uint8)
            out[i, cb:2 * cb] = np.frombuffer(self.fe_to_mont_bytes(pt[1]), dtype=np.uint8)
        return out

    def affine_from_records(self, arr: np.ndarray):
        if len(arr) == 0:
            return []
        arr = np.ascontiguousarray(arr).view(np.uint8).reshape(len(arr), -1)
        cb = self.coord_bytes
        pts = []
        for rec in arr:
            raw = rec.tobytes()
            if len(raw) > 2 * cb and raw[2 * cb]:
                pts.append(None)
                continue
            x = self.fe_from_mont_bytes(raw[:cb])
            y = self.fe_from_mont_bytes(raw[cb:2 * cb])
            zero = 0 if self.curve.ext_degree == 1 else (0, 0)
            pts.append(None if (len(raw) == 2 * cb and x == zero and y == zero) else (x, y))
        return pts

    def jacobian_to_affine(self, raw: bytes):
        """arkworks GroupProjective bytes (X|Y|Z Montgomery) -> oracle affine point."""
        cb = self.coord_bytes
        c = self.curve
        x, y, z = (self.fe_from_mont_bytes(raw[i * cb:(i + 1) * cb]) for i in range(3))
        if z == c.zero:
            return None
        zi = c.inv(z)
        zi2 = c.mul(zi, zi)
        return (c.mul(x, zi2), c.mul(y, c.mul(zi2, zi)))

    def jacobian_compressed(self, raw: bytes) -> bytes:
        """The parity comparand: canonical arkworks compressed bytes."""
        return O.serialize_compressed(self.curve, self.jacobian_to_affine(raw))

    def scalars_array(self, scalars: Sequence[int]) -> np.ndarray:
        out = np.zeros((len(scalars), self.scalar_limbs), dtype=np.uint64)
        for i, s in enumerate(scalars):
            for j in range(self.scalar_limbs):
                out[i, j] = (s >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
        return out

    def scalars_from_array(self, arr: np.ndarray):
        return [sum(int(v) << (64 * j) for j, v in enumerate(row)) for row in arr]


LAYOUTS = {name: CurveLayout(c) for name, c in O.CURVES.items()}


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
            for f in ("cpu_ref.c", "fp_tmpl.h", "ec_tmpl.h", "pairing_tmpl.h", "ntt_tmpl.h")):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.cpu_ref_init()
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def msm(layout: CurveLayout, bases: np.ndarray, scalars: np.ndarray, threads: int = 0) -> bytes:
    """arkworks-algorithm Pippenger on the host; returns Jacobian bytes."""
    bases = np.ascontiguousarray(bases)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    n = min(len(bases), len(scalars))
    stride = bases.strides[0] if bases.ndim == 2 else layout.ark_stride
    out = np.zeros(layout.jac_bytes, dtype=np.uint8)
    threads = threads or os.cpu_count() or 1
    lib().cpu_ref_msm(layout.id, _ptr(bases), ctypes.c_size_t(stride), _ptr(scalars), ctypes.c_size_t(n),
                      _ptr(out), threads)
    return out.tobytes()


def msm_window_tasks(layout: CurveLayout, n: int) -> int:
    c = O.msm_window_bits(n)
    return (layout.curve.scalar_bits + c - 1) // c


def fixed_base_batch(layout: CurveLayout, base, scalars: Sequence[int]) -> np.ndarray:
    """[k * base for k in scalars] as packed affine records (x|y)."""
    rec = layout.affine_records([base])
    sc = layout.scalars_array(scalars)
    out = np.zeros((len(scalars), layout.packed_stride), dtype=np.uint8)
    lib().cpu_ref_fixed_base_batch(layout.id, _ptr(rec), ctypes.c_size_t(rec.strides[0]), _ptr(sc),
                                   ctypes.c_size_t(len(scalars)), _ptr(out))
    return out


def scalar_mul(layout: CurveLayout, base, k: int) -> bytes:
    rec = layout.affine_records([base])
    sc = layout.scalars_array([k])
    out = np.zeros(layout.jac_bytes, dtype=np.uint8)
    lib().cpu_ref_scalar_mul(layout.id, _ptr(rec), ctypes.c_size_t(rec.strides[0]), _ptr(sc), _ptr(out))
    return out.tobytes()


def with_flags(layout: CurveLayout, packed: np.ndarray) -> np.ndarray:
    """packed (x|y) records -> arkworks-stride records with a clear infinity flag."""
    out = np.zeros((len(packed), layout.ark_stride), dtype=np.uint8)
    out[:, :layout.packed_stride] = packed
    return out


# ---- BLS12-377 GT codec (arkworks Fq12 image: 12 Montgomery Fq residues, tower order) ----------
def fq12_to_ark_bytes(f) -> bytes:
    L = LAYOUTS["bls12_377_g1"]
    out = b""
    for c6 in f:                 # c0, c1 (Fq6)
        for c2 in c6:            # c0, c1, c2 (Fq2)
            for c in c2:         # c0, c1 (Fq)
                out += L.fe_to_mont_bytes(c)
    return out


def fq12_from_ark_bytes(raw: bytes):
    L = LAYOUTS["bls12_377_g1"]
    vals = [L.fe_from_mont_bytes(raw[48 * i:48 * (i + 1)]) for i in range(12)]
    f2 = [(vals[2 * i], vals[2 * i + 1]) for i in range(6)]
    return ((f2[0], f2[1], f2[2]), (f2[3], f2[4], f2[5]))


def hash_to_g1_direct(domain: bytes, message: bytes, extra: bytes, compat: bool = True):
    """DIRECT_HASH_TO_G1 by the C port (cpu_ref_hash_to_g1_direct) -> (144-byte G1Projective image, attempt)."""
    out = np.zeros(144, dtype=np.uint8)
    att = ctypes.c_uint32(0)
    rc = lib().cpu_ref_hash_to_g1_direct(domain.ljust(8, b"\0"), message, ctypes.c_size_t(len(message)), extra,
                                         ctypes.c_size_t(len(extra)), int(compat), _ptr(out), ctypes.byref(att))
    if rc != 0:
        raise ValueError("hash to curve failed")
    return out.tobytes(), att.value


def bh_crh(message: bytes) -> bytes:
    """CompositeHasher::crh by the C port -> 48 bytes (x coordinate of the Bowe-Hopwood point)."""
    out = np.zeros(48, dtype=np.uint8)
    if lib().cpu_ref_bh_crh(message, ctypes.c_size_t(len(message)), _ptr(out)) != 0:
        raise ValueError("message too long for the CRH")
    return out.tobytes()


def hash_to_g1_composite(domain: bytes, message: bytes, extra: bytes, compat: bool = True, cip22: bool = False):
    """COMPOSITE_HASH_TO_G1 / ..._CIP22 by the C port -> (144-byte G1Projective image, attempt)."""
    out = np.zeros(144, dtype=np.uint8)
    att = ctypes.c_uint32(0)
    rc = lib().cpu_ref_hash_to_g1_composite(domain.ljust(8, b"\0"), message, ctypes.c_size_t(len(message)), extra,
                                            ctypes.c_size_t(len(extra)), int(compat), int(cip22), _ptr(out), ctypes.byref(att))
    if rc != 0:
        raise ValueError("hash to curve failed" if rc == 1 else "message too long for the CRH")
    return out.tobytes(), att.value


# ---- BLS12-377 product of pairings and the prover's radix-2 transforms (pairing_tmpl.h, ntt_tmpl.h) ----------
def multi_pairing(g1: np.ndarray, g2: np.ndarray, n: Optional[int] = None, threads: int = 1, phase: int = 0,
                  value: Optional[bytes] = None):
    """cpu_ref_multi_pairing over arkworks-layout records (uint8 [n, 104 | 96] and [n, 200 | 192]).
    phase 0: product of pairings; 1: Miller value only; 2: final exponentiation of `value`.
    threads = 1 is arkworks' own serial Miller loop.  Returns (is_one, 576-byte Fq12 image)."""
    g1 = np.ascontiguousarray(g1)
    g2 = np.ascontiguousarray(g2)
    if n is None:
        n = min(len(g1), len(g2))
    out = np.zeros(576, dtype=np.uint8)
    if value is not None:
        out[:] = np.frombuffer(value, dtype=np.uint8)
    flag = ctypes.c_int(0)
    s1 = g1.strides[0] if n else 104
    s2 = g2.strides[0] if n else 200
    rc = lib().cpu_ref_multi_pairing(_ptr(g1) if n else None, ctypes.c_size_t(s1), _ptr(g2) if n else None, ctypes.c_size_t(s2),
                                     ctypes.c_size_t(n), _ptr(out), ctypes.byref(flag), int(threads), int(phase))
    assert rc == 0
    return bool(flag.value), out.tobytes()


def ntt(field_id: int, data: np.ndarray, log_n: int, inverse: bool, coset: bool, threads: int = 1) -> np.ndarray:
    """In-place fft / ifft / coset_fft / coset_ifft on a uint64 [n, limbs] array of Montgomery residues (a copy is returned)."""
    a = np.ascontiguousarray(data, dtype=np.uint64).copy()
    rc = lib().cpu_ref_ntt(int(field_id), _ptr(a), ctypes.c_uint(log_n), int(inverse), int(coset), int(threads))
    assert rc == 0
    return a


def witness_map(field_id: int, a: np.ndarray, b: np.ndarray, c: np.ndarray, log_n: int, threads: int = 1) -> np.ndarray:
    a, b, c = (np.ascontiguousarray(x, dtype=np.uint64).copy() for x in (a, b, c))
    h = np.zeros_like(a)
    rc = lib().cpu_ref_witness_map(int(field_id), _ptr(a), _ptr(b), _ptr(c), ctypes.c_uint(log_n), _ptr(h), int(threads))
    assert rc == 0
    return h
