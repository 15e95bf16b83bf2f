"""CPU-only checks of the drop-in boundary: the library loads and exports every symbol
include/b200_bls.h declares; without a GPU compute calls fail loudly (no fallback)."""
import os
import re

import pytest

from celo_bls_snark_rs_b200 import engine as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200_bls.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = E.load()
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(E.EXPORTS) == syms
    # the bls-snark-sys names re-exported for existing C / cgo consumers (include/bls_snark_sys_compat.h)
    compat = open(os.path.join(ROOT, "include", "bls_snark_sys_compat.h")).read()
    compat = re.sub(r"/\*.*?\*/", "", compat, flags=re.S)
    names = sorted(set(re.findall(r"\bbool\s+([a-z_0-9]+)\s*\(", compat)))
    assert names == sorted(E.COMPAT_EXPORTS)
    for s in names:
        assert hasattr(lib, s), s


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(E.B200Error):
        E.init(0)
    lib = E.load()
    assert lib.b200_msm(0, None, 104, None, 0, None) != 0


def test_product_never_imports_the_oracle():
    """The product path may not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "celo_bls_snark_rs_b200")
    banned = re.compile(r"(import\s+oracle|from\s+oracle|from\s+\.\.?oracle|oracle/|cpu_ref|libcpu_ref|cref)")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not banned.search(text), f
    # and the shared library has no dependency on the C oracle
    import subprocess
    out = subprocess.run(["ldd", E.LIB_PATH], capture_output=True, text=True).stdout
    assert "cpu_ref" not in out


def test_compress_entry_points_on_host():
    """compress_signature / compress_pubkey (crates/bls-snark-sys/src/serialization.rs:166-215) are integer work done
    on the host: uncompressed oracle encodings in, the oracle's compressed encodings out."""
    import ctypes
    from oracle import oracle as O
    lib = E.load()
    rng = O.SplitMix64(12)
    for curve, gen, fn, size in ((O.G1, O.G1_GEN, "compress_signature", 48), (O.G2, O.G2_GEN, "compress_pubkey", 96)):
        for _ in range(6):
            pt = curve.pmul(gen, rng.below(O.R - 1) + 1)
            for p in (pt, curve.pneg(pt)):
                raw = O.serialize_uncompressed(curve, p)
                ptr, n = ctypes.c_void_p(), ctypes.c_int()
                assert getattr(lib, fn)(raw, len(raw), ctypes.byref(ptr), ctypes.byref(n)) and n.value == size
                assert ctypes.string_at(ptr, n.value) == O.serialize_compressed(curve, p)
                assert lib.free_vec(ptr, n.value)
        ptr, n = ctypes.c_void_p(), ctypes.c_int()
        bad = (O.P).to_bytes(48, "little") * (2 * size // 48)          # coordinate == modulus: Fq::read fails
        assert not getattr(lib, fn)(bad, len(bad), ctypes.byref(ptr), ctypes.byref(n))
        assert not getattr(lib, fn)(b"\0" * (2 * size - 1), 2 * size - 1, ctypes.byref(ptr), ctypes.byref(n))


def test_plan_is_sane():
    c, w, nb = E.msm_plan(E.BLS12_377_G1, 1 << 20)
    assert nb == 1 << (c - 1) and w * c >= 254
    c, w, nb = E.msm_plan(E.BW6_761_G1, 1 << 22)
    assert w * c >= 378


def test_byte_helper_exports_on_host():
    """hash_direct_first_step and b200_encode_epoch_block are host byte work (no GPU): checked here against the oracle's
    DirectHasher and against the reference's own encoding KATs (crates/epoch-snark/src/epoch_block.rs:243-320)."""
    import ctypes
    import json
    from oracle import hash_to_curve as H
    from oracle import oracle as O
    lib = E.load()
    for msg, n in ((b"", 32), (b"abc", 64), (bytes(range(200)), 96), (b"x" * 65, 33), (b"y", 1), (b"z" * 64, 0)):
        ptr, ln = ctypes.c_void_p(), ctypes.c_int()
        assert lib.hash_direct_first_step(msg, len(msg), n, ctypes.byref(ptr), ctypes.byref(ln))
        assert ctypes.string_at(ptr, ln.value) == H.direct_hash(b"ULforxof", msg, n)
        assert lib.free_vec(ptr, ln.value)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))
    key = O.serialize_compressed(O.G2, O.G2_GEN)
    lib.b200_encode_epoch_block.argtypes = [ctypes.c_int, ctypes.c_uint16, ctypes.c_uint8, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_uint32,
                                            ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p),
                                            ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]
    inner, n_in, extra, n_ex = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_void_p(), ctypes.c_size_t()
    assert lib.b200_encode_epoch_block(0, 120, 0, None, None, 3, 10, key * 10, 10, ctypes.byref(inner), ctypes.byref(n_in), None, None) == 0
    assert ctypes.string_at(inner, n_in.value).hex() == gold["epoch_block_encoding_before_donut"]["hex"]
    lib.free_vec(inner, n_in.value)
    # the inner CIP22 encoding is both entropies + keys + padding: the first-epoch KAT minus its 16 + 32 leading bits, checked
    # bit for bit against the oracle class that reproduces all four KATs (tests/test_oracle_golden.py)
    from oracle import bw6_verify as V
    for ee, pe, maxv in ((bytes([255] * 16), bytes([254] * 16), 10), (None, bytes([7] * 16), 12), (None, None, 10)):
        assert lib.b200_encode_epoch_block(1, 120, 5, ee, pe, 3, maxv, key * 10, 10, ctypes.byref(inner), ctypes.byref(n_in),
                                           ctypes.byref(extra), ctypes.byref(n_ex)) == 0
        want_inner, want_extra = V.EpochBlock(120, 5, ee, pe, 3, maxv, [O.G2_GEN] * 10).encode_inner_to_bytes_cip22()
        assert ctypes.string_at(inner, n_in.value) == want_inner and ctypes.string_at(extra, n_ex.value) == want_extra
        lib.free_vec(inner, n_in.value)
        lib.free_vec(extra, n_ex.value)
