"""ctypes binding of libb200bls.so (include/b200_bls.h) -- the only compute path.

There is no CPU fallback: if the shared library is missing or no CUDA device is
usable, every entry point raises.  torch is used by callers for device memory and
streams only; this module passes raw device pointers across the C-ABI.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200bls.so")

BLS12_377_G1, BLS12_377_G2, BW6_761_G1, BW6_761_G2 = 0, 1, 2, 3
CURVE_IDS = {"bls12_377_g1": 0, "bls12_377_g2": 1, "bw6_761_g1": 2, "bw6_761_g2": 3}

# bytes of one coordinate / one scalar / one Jacobian point, and the arkworks GroupAffine stride
COORD_BYTES = {0: 48, 1: 96, 2: 96, 3: 96}
SCALAR_BYTES = {0: 32, 1: 32, 2: 48, 3: 48}
JAC_BYTES = {0: 144, 1: 288, 2: 288, 3: 288}
ARK_STRIDE = {0: 104, 1: 200, 2: 200, 3: 200}
PACKED_STRIDE = {k: 2 * v for k, v in COORD_BYTES.items()}

EXPORTS = [
    "b200_init", "b200_init_devices", "b200_device_count", "b200_msm_sharded", "b200_multi_pairing_bls12_377_sharded", "b200_shutdown", "b200_last_error", "b200_msm", "b200_msm_bls12_377_g1", "b200_msm_bls12_377_g2",
    "b200_msm_bw6_761_g1", "b200_msm_bw6_761_g2", "b200_msm_device", "b200_pack_bases_device", "b200_msm_prepared_device",
    "b200_sum_jacobian_device", "b200_sum_jacobian_batch_device", "b200_sum_jacobian", "b200_scalar_mul", "b200_fixed_base_mul_device", "b200_point_runs_device", "b200_batch_to_affine_device", "b200_sync",
    "b200_msm_plan", "b200_launch_count", "b200_profile_enable", "b200_profile_read",
    "b200_field_op_device", "b200_multi_pairing_bls12_377", "b200_miller_product_bls12_377_device",
    "b200_final_exp_bls12_377_device", "b200_batch_verify_hashes", "b200_batch_verify_strict_hash", "b200_batch_verify_strict_many",
    "b200_ntt_device", "b200_witness_map_device", "b200_groth16_prove_device", "b200_groth16_prove_partial_device", "b200_groth16_assemble_device", "b200_msm_batch_device",
    "b200_multi_pairing_bw6_761", "b200_miller_values_bw6_761_device", "b200_final_exp_bw6_761_device",
    "b200_groth16_verify_bw6_761", "b200_deserialize_points", "b200_verify_epochs", "b200_epoch_public_inputs",
    "b200_blake2s_personal", "b200_blake2s_param", "b200_encode_epoch_block", "b200_hash_to_g1", "b200_serialize_points", "b200_ensure_init", "b200_bound_device",
]
# the reference's own symbols re-exported by the library (include/bls_snark_sys_compat.h)
COMPAT_EXPORTS = ["verify", "deserialize_public_key", "deserialize_signature", "serialize_public_key", "serialize_signature",
                  "free_vec", "destroy_public_key", "destroy_signature", "aggregate_public_keys", "aggregate_signatures",
                  "verify_signature", "verify_pop", "batch_verify_signature", "batch_verify_strict", "compress_signature",
                  "compress_pubkey", "init", "generate_private_key", "deserialize_private_key", "serialize_private_key",
                  "destroy_private_key", "private_key_to_public_key", "sign_message", "sign_pop", "hash_direct",
                  "hash_direct_with_attempt", "hash_composite", "hash_composite_cip22", "hash_crh", "deserialize_public_key_cached",
                  "serialize_public_key_uncompressed", "serialize_signature_uncompressed", "aggregate_public_keys_subtract",
                  "hash_direct_first_step", "encode_epoch_block_to_bytes", "encode_epoch_block_to_bytes_cip22"]


class Groth16Pk(ctypes.Structure):
    """b200_groth16_pk: device pointers to the packed proving-key queries."""
    _fields_ = [(k, ctypes.c_void_p) for k in ("a_query", "b_g2_query", "h_query", "l_query", "alpha_g1", "beta_g2")]


class Groth16Vk(ctypes.Structure):
    """b200_groth16_vk: host pointers to the arkworks VerifyingKey<BW6_761> members."""
    _fields_ = [(k, ctypes.c_void_p) for k in ("alpha_g1", "beta_g2", "gamma_g2", "delta_g2", "gamma_abc_g1")] + [
        ("num_gamma_abc", ctypes.c_size_t), ("stride", ctypes.c_size_t)]


class EpochBlockFFI(ctypes.Structure):
    """bls-snark-sys EpochBlockFFI (crates/bls-snark-sys/src/snark/epoch_block.rs:109-127), #[repr(C)], 56 bytes."""
    _fields_ = [("index", ctypes.c_uint16), ("round", ctypes.c_uint8), ("epoch_entropy", ctypes.c_void_p),
                ("parent_entropy", ctypes.c_void_p), ("pubkeys", ctypes.c_void_p), ("pubkeys_num", ctypes.c_size_t),
                ("maximum_non_signers", ctypes.c_uint32), ("maximum_validators", ctypes.c_size_t)]


class HashInput(ctypes.Structure):
    """b200_hash_input"""
    _fields_ = [("message", ctypes.c_char_p), ("message_len", ctypes.c_size_t), ("extra_data", ctypes.c_char_p),
                ("extra_data_len", ctypes.c_size_t)]


class FFIBuffer(ctypes.Structure):
    """bls-snark-sys Buffer (utils.rs:75-82)"""
    _fields_ = [("ptr", ctypes.c_char_p), ("len", ctypes.c_size_t)]


class MessageFFI(ctypes.Structure):
    """bls-snark-sys MessageFFI (utils.rs:20-32), 48 bytes"""
    _fields_ = [("data", FFIBuffer), ("extra", FFIBuffer), ("public_key", ctypes.c_void_p), ("sig", ctypes.c_void_p)]


class BatchMessageFFI(ctypes.Structure):
    """bls-snark-sys BatchMessageFFI (utils.rs:58-72), 64 bytes"""
    _fields_ = [("data", FFIBuffer), ("extra", FFIBuffer), ("public_keys", ctypes.POINTER(ctypes.c_void_p)),
                ("public_keys_len", ctypes.c_size_t), ("signatures", ctypes.POINTER(ctypes.c_void_p)),
                ("signatures_len", ctypes.c_size_t)]


class MsmJob(ctypes.Structure):
    """b200_msm_job"""
    _fields_ = [("d_bases_packed", ctypes.c_void_p), ("d_scalars", ctypes.c_void_p), ("n", ctypes.c_size_t),
                ("d_out_jacobian", ctypes.c_void_p)]


class B200Error(RuntimeError):
    pass


_lib = None


def load() -> ctypes.CDLL:
    """Loads the CUDA library.  Raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(f"{LIB_PATH} is missing: run `python -m celo_bls_snark_rs_b200.build` (no CPU fallback exists)")
    lib = ctypes.CDLL(LIB_PATH)
    vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    lib.b200_init.argtypes = [i32]
    lib.b200_last_error.restype = ctypes.c_char_p
    lib.b200_msm.argtypes = [i32, vp, sz, vp, sz, vp]
    lib.b200_msm_sharded.argtypes = [i32, vp, sz, vp, sz, vp]
    lib.b200_init_devices.argtypes = [ctypes.POINTER(i32), i32]
    lib.b200_multi_pairing_bls12_377_sharded.argtypes = [vp, sz, vp, sz, sz, vp, ctypes.POINTER(i32)]
    for name in ("b200_msm_bls12_377_g1", "b200_msm_bls12_377_g2", "b200_msm_bw6_761_g1", "b200_msm_bw6_761_g2"):
        getattr(lib, name).argtypes = [vp, vp, sz, vp]
    lib.b200_msm_device.argtypes = [i32, vp, vp, sz, vp, vp]
    lib.b200_pack_bases_device.argtypes = [i32, vp, sz, sz, i32, vp, vp]
    lib.b200_msm_prepared_device.argtypes = [i32, vp, vp, sz, vp, vp]
    lib.b200_sum_jacobian_device.argtypes = [i32, vp, sz, vp, vp]
    lib.b200_sum_jacobian.argtypes = [i32, vp, sz, vp]
    lib.b200_sum_jacobian_batch_device.argtypes = [i32, vp, sz, sz, vp, vp]
    lib.b200_fixed_base_mul_device.argtypes = [i32, vp, vp, sz, vp, vp]
    lib.b200_point_runs_device.argtypes = [i32, vp, vp, sz, sz, vp, vp]
    lib.b200_batch_to_affine_device.argtypes = [i32, vp, sz, vp, vp]
    lib.b200_field_op_device.argtypes = [i32, i32, vp, vp, sz, vp, vp]
    lib.b200_multi_pairing_bls12_377.argtypes = [vp, sz, vp, sz, sz, vp, ctypes.POINTER(i32)]
    lib.b200_miller_product_bls12_377_device.argtypes = [vp, vp, sz, vp, vp]
    lib.b200_final_exp_bls12_377_device.argtypes = [vp, sz, vp, vp, vp]
    lib.b200_multi_pairing_bw6_761.argtypes = [vp, sz, vp, sz, sz, vp, ctypes.POINTER(i32)]
    lib.b200_miller_values_bw6_761_device.argtypes = [vp, vp, sz, vp, vp]
    lib.b200_final_exp_bw6_761_device.argtypes = [vp, sz, vp, vp, vp]
    lib.b200_groth16_verify_bw6_761.argtypes = [ctypes.POINTER(Groth16Vk), vp, vp, vp, vp, sz, ctypes.POINTER(i32)]
    lib.b200_deserialize_points.argtypes = [i32, vp, sz, i32, vp, vp]
    lib.b200_verify_epochs.argtypes = [vp, sz, vp, sz, ctypes.POINTER(EpochBlockFFI), ctypes.POINTER(EpochBlockFFI),
                                       ctypes.POINTER(i32)]
    lib.b200_epoch_public_inputs.argtypes = [ctypes.POINTER(EpochBlockFFI), ctypes.POINTER(EpochBlockFFI), vp, sz,
                                             ctypes.POINTER(sz), ctypes.POINTER(i32)]
    lib.b200_blake2s_personal.argtypes = [vp, sz, vp, vp]
    lib.b200_blake2s_personal.restype = None
    lib.verify.argtypes = [vp, ctypes.c_uint32, vp, ctypes.c_uint32, EpochBlockFFI, EpochBlockFFI]
    lib.verify.restype = ctypes.c_bool
    lib.b200_batch_verify_hashes.argtypes = [vp, vp, vp, sz, ctypes.POINTER(i32)]
    lib.b200_batch_verify_strict_hash.argtypes = [vp, vp, vp, sz, vp, ctypes.POINTER(i32)]
    lib.b200_ntt_device.argtypes = [i32, vp, ctypes.c_uint, i32, i32, vp]
    lib.b200_witness_map_device.argtypes = [i32, vp, vp, vp, ctypes.c_uint, vp, vp]
    lib.b200_groth16_prove_device.argtypes = [i32, ctypes.POINTER(Groth16Pk), vp, sz, sz, vp, vp, vp, ctypes.c_uint, vp, vp]
    lib.b200_groth16_prove_partial_device.argtypes = [i32, ctypes.POINTER(Groth16Pk), vp, sz, sz, vp, vp, vp, ctypes.c_uint,
                                                      ctypes.c_uint, ctypes.c_uint, vp, vp]
    lib.b200_groth16_assemble_device.argtypes = [i32, ctypes.POINTER(Groth16Pk), vp, ctypes.c_uint, vp, vp]
    lib.b200_msm_batch_device.argtypes = [i32, ctypes.POINTER(MsmJob), sz, vp]
    lib.b200_hash_to_g1.argtypes = [i32, i32, ctypes.c_char_p, sz, ctypes.POINTER(HashInput), sz, vp, vp]
    lib.b200_serialize_points.argtypes = [i32, vp, sz, vp]
    cb, pp, pb, ci = ctypes.c_bool, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_bool), ctypes.c_int
    pci, pu8 = ctypes.POINTER(ci), ctypes.POINTER(ctypes.c_uint8)
    for name, args in (("generate_private_key", [pp]), ("deserialize_private_key", [vp, ci, pp]), ("serialize_private_key", [vp, pp, pci]),
                       ("destroy_private_key", [vp]), ("private_key_to_public_key", [vp, pp]),
                       ("sign_message", [vp, vp, ci, vp, ci, cb, cb, pp]), ("sign_pop", [vp, vp, ci, pp]),
                       ("hash_direct", [vp, ci, pp, pci, cb]), ("hash_direct_with_attempt", [vp, ci, pp, pci, pci, cb]),
                       ("hash_composite", [vp, ci, vp, ci, pp, pci]), ("hash_composite_cip22", [vp, ci, vp, ci, pp, pci, pu8]),
                       ("hash_crh", [vp, ci, ci, pp, pci]), ("deserialize_public_key_cached", [vp, ci, pp]),
                       ("serialize_public_key_uncompressed", [vp, pp, pci]), ("serialize_signature_uncompressed", [vp, pp, pci]),
                       ("aggregate_public_keys_subtract", [vp, pp, ci, pp]), ("init", []),
                       ("hash_direct_first_step", [vp, ci, ci, pp, pci]),
                       ("encode_epoch_block_to_bytes", [ctypes.c_ushort, ctypes.c_uint, pp, ci, pp, pci]),
                       ("encode_epoch_block_to_bytes_cip22", [ctypes.c_ushort, ctypes.c_ubyte, vp, vp, ctypes.c_uint, ctypes.c_uint, pp, ci,
                                                              pp, pci, pp, pci])):
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = args, cb
    for name, args in (("deserialize_public_key", [vp, ci, pp]), ("deserialize_signature", [vp, ci, pp]),
                       ("serialize_public_key", [vp, pp, ctypes.POINTER(ci)]), ("serialize_signature", [vp, pp, ctypes.POINTER(ci)]),
                       ("compress_signature", [vp, ci, pp, ctypes.POINTER(ci)]), ("compress_pubkey", [vp, ci, pp, ctypes.POINTER(ci)]),
                       ("free_vec", [vp, ci]), ("destroy_public_key", [vp]), ("destroy_signature", [vp]),
                       ("aggregate_public_keys", [pp, ci, pp]), ("aggregate_signatures", [pp, ci, pp]),
                       ("verify_signature", [vp, vp, ci, vp, ci, vp, cb, cb, pb]), ("verify_pop", [vp, vp, ci, vp, pb]),
                       ("batch_verify_signature", [ctypes.POINTER(MessageFFI), sz, cb, cb, pb]),
                       ("batch_verify_strict", [ctypes.POINTER(BatchMessageFFI), sz, cb, cb, pb])):
        getattr(lib, name).argtypes = args
        getattr(lib, name).restype = cb
    lib.b200_sync.argtypes = [vp]
    lib.b200_msm_plan.argtypes = [i32, sz, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(ctypes.c_uint32)]
    lib.b200_launch_count.restype = ctypes.c_uint64
    lib.b200_profile_enable.argtypes = [i32]
    lib.b200_profile_read.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i32),
                                      ctypes.POINTER(ctypes.c_uint64)]
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        raise B200Error(f"b200 error {rc}: {load().b200_last_error().decode()}")


def init(device: int = -1):
    _check(load().b200_init(device))


def init_devices(devices):
    """One engine per listed GPU in THIS process (b200_init_devices); devices[0] is the primary engine."""
    arr = (ctypes.c_int * len(devices))(*devices)
    _check(load().b200_init_devices(arr, len(devices)))


def device_count() -> int:
    return load().b200_device_count()


def shutdown():
    load().b200_shutdown()


def launch_count() -> int:
    return int(load().b200_launch_count())


def msm_plan(curve: int, n: int):
    c, w, nb = ctypes.c_int(), ctypes.c_int(), ctypes.c_uint32()
    _check(load().b200_msm_plan(curve, n, ctypes.byref(c), ctypes.byref(w), ctypes.byref(nb)))
    return c.value, w.value, nb.value


def _hptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def msm(curve: int, bases: np.ndarray, scalars: np.ndarray, n: Optional[int] = None) -> bytes:
    """Host-pointer MSM (b200_msm).  bases: uint8 [n, stride] (arkworks or packed
    records); scalars: uint64 [n, limbs].  Returns the Jacobian result bytes."""
    bases = np.ascontiguousarray(bases)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    if n is None:
        n = min(len(bases), len(scalars))
    stride = bases.strides[0] if (bases.ndim == 2 and len(bases)) else ARK_STRIDE[curve]
    out = np.zeros(JAC_BYTES[curve], dtype=np.uint8)
    _check(load().b200_msm(curve, _hptr(bases), stride, _hptr(scalars), n, _hptr(out)))
    return out.tobytes()


def msm_sharded(curve: int, bases: np.ndarray, scalars: np.ndarray, n: Optional[int] = None) -> bytes:
    """b200_msm_sharded: one host-pointer MSM spread over the GPUs of init_devices."""
    bases = np.ascontiguousarray(bases)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    if n is None:
        n = min(len(bases), len(scalars))
    stride = bases.strides[0] if n else ARK_STRIDE[curve]
    out = np.zeros(JAC_BYTES[curve], dtype=np.uint8)
    _check(load().b200_msm_sharded(curve, _hptr(bases) if n else None, stride, _hptr(scalars) if n else None, n, _hptr(out)))
    return out.tobytes()


def multi_pairing_sharded(g1: np.ndarray, g2: np.ndarray, n: Optional[int] = None, want_gt: bool = True):
    """b200_multi_pairing_bls12_377_sharded: pairs split over the GPUs of init_devices.  Returns (is_one, gt | None)."""
    g1, g2 = np.ascontiguousarray(g1), np.ascontiguousarray(g2)
    if n is None:
        n = min(len(g1), len(g2))
    out = np.zeros(576, dtype=np.uint8)
    flag = ctypes.c_int(0)
    _check(load().b200_multi_pairing_bls12_377_sharded(_hptr(g1), g1.strides[0] if n else 104, _hptr(g2), g2.strides[0] if n else 200, n,
                                                       _hptr(out) if want_gt else None, ctypes.byref(flag)))
    return bool(flag.value), (out.tobytes() if want_gt else None)


def msm_host_ptrs(curve: int, bases_ptr: int, stride: int, scalars_ptr: int, n: int, out: np.ndarray):
    """b200_msm on raw host addresses (e.g. pinned torch tensors' data_ptr())."""
    _check(load().b200_msm(curve, bases_ptr, stride, scalars_ptr, n, _hptr(out)))


def msm_device(curve: int, d_bases: int, d_scalars: int, n: int, d_out: int, stream: int = 0):
    _check(load().b200_msm_device(curve, d_bases, d_scalars, n, d_out, stream or None))


def msm_batch_device(curve: int, jobs, stream: int = 0):
    """jobs: sequence of (d_bases_packed, d_scalars, n, d_out) -- independent MSMs, pipelined in the engine."""
    arr = (MsmJob * len(jobs))(*[MsmJob(b or None, s or None, n, o) for b, s, n, o in jobs])
    _check(load().b200_msm_batch_device(curve, arr, len(jobs), stream or None))


def msm_prepared_device(curve: int, d_bases_prepared: int, d_scalars: int, n: int, d_out: int, stream: int = 0):
    _check(load().b200_msm_prepared_device(curve, d_bases_prepared, d_scalars, n, d_out, stream or None))


def pack_bases_device(curve: int, src: int, stride: int, n: int, src_on_device: bool, d_dst: int, stream: int = 0):
    _check(load().b200_pack_bases_device(curve, src, stride, n, int(src_on_device), d_dst, stream or None))


def sum_jacobian_device(curve: int, d_points: int, count: int, d_out: int, stream: int = 0):
    _check(load().b200_sum_jacobian_device(curve, d_points, count, d_out, stream or None))


def sum_jacobian_batch_device(curve: int, d_points: int, count: int, batch: int, d_out: int, stream: int = 0):
    """d_out[b] = sum_i d_points[i * batch + b]: combines a batch of sharded MSMs after ONE all-gather."""
    _check(load().b200_sum_jacobian_batch_device(curve, d_points, count, batch, d_out, stream or None))


def fixed_base_mul_device(curve: int, d_base: int, d_scalars: int, n: int, d_out: int, stream: int = 0):
    _check(load().b200_fixed_base_mul_device(curve, d_base, d_scalars, n, d_out, stream or None))


def point_runs_device(curve: int, d_base: int, d_start_scalars: int, runs: int, run_len: int, d_out: int, stream: int = 0):
    """d_out[t * run_len + j] = (start[t] + j) * base as packed affine records (b200_point_runs_device)."""
    _check(load().b200_point_runs_device(curve, d_base, d_start_scalars, runs, run_len, d_out, stream or None))


def batch_to_affine_device(curve: int, d_jac: int, n: int, d_out: int, stream: int = 0):
    _check(load().b200_batch_to_affine_device(curve, d_jac, n, d_out, stream or None))


FIELD_OPS = {"add": 0, "sub": 1, "mul": 2, "sqr": 3, "inv": 4, "neg": 5, "dbl": 6,
             # the warp-cooperative routines of csrc/coop.cuh (one warp per element)
             "coop_add": 8, "coop_sub": 9, "coop_mul": 10, "coop_sqr": 11, "coop_neg": 13, "coop_dbl": 14}


def field_op_device(curve: int, op: str, d_a: int, d_b: int, n: int, d_out: int, stream: int = 0):
    _check(load().b200_field_op_device(curve, FIELD_OPS[op], d_a, d_b, n, d_out, stream or None))


FQ12_BYTES = 576


def multi_pairing(g1: np.ndarray, g2: np.ndarray, n: Optional[int] = None, want_gt: bool = True):
    """Host-pointer product of pairings over BLS12-377 (b200_multi_pairing_bls12_377).
    g1: uint8 [n, 104 | 96], g2: uint8 [n, 200 | 192].  Returns (is_one, gt_bytes | None)."""
    g1 = np.ascontiguousarray(g1)
    g2 = np.ascontiguousarray(g2)
    if n is None:
        n = min(len(g1), len(g2))
    s1 = g1.strides[0] if n else 104
    s2 = g2.strides[0] if n else 200
    out = np.zeros(FQ12_BYTES, dtype=np.uint8)
    flag = ctypes.c_int(0)
    _check(load().b200_multi_pairing_bls12_377(_hptr(g1), s1, _hptr(g2), s2, n, _hptr(out) if want_gt else None,
                                               ctypes.byref(flag)))
    return bool(flag.value), (out.tobytes() if want_gt else None)


def miller_product_device(d_g1: int, d_g2: int, n: int, d_out: int, stream: int = 0):
    _check(load().b200_miller_product_bls12_377_device(d_g1 or None, d_g2 or None, n, d_out, stream or None))


def final_exp_device(d_vals: int, count: int, d_out: int, d_is_one: int = 0, stream: int = 0):
    _check(load().b200_final_exp_bls12_377_device(d_vals, count, d_out or None, d_is_one or None, stream or None))


FQ6_BW6_BYTES = 576


def multi_pairing_bw6(g1: np.ndarray, g2: np.ndarray, n: Optional[int] = None, want_gt: bool = True):
    """Host-pointer product of pairings over BW6-761 (b200_multi_pairing_bw6_761).
    g1, g2: uint8 [n, 200 | 192].  Returns (is_one, fq6_bytes | None)."""
    g1 = np.ascontiguousarray(g1)
    g2 = np.ascontiguousarray(g2)
    if n is None:
        n = min(len(g1), len(g2))
    s1 = g1.strides[0] if n else 200
    s2 = g2.strides[0] if n else 200
    out = np.zeros(FQ6_BW6_BYTES, dtype=np.uint8)
    flag = ctypes.c_int(0)
    _check(load().b200_multi_pairing_bw6_761(_hptr(g1), s1, _hptr(g2), s2, n, _hptr(out) if want_gt else None,
                                             ctypes.byref(flag)))
    return bool(flag.value), (out.tobytes() if want_gt else None)


def miller_values_bw6_device(d_g1: int, d_g2: int, n: int, d_out: int, stream: int = 0):
    _check(load().b200_miller_values_bw6_761_device(d_g1 or None, d_g2 or None, n, d_out or None, stream or None))


def final_exp_bw6_device(d_vals: int, count: int, d_out: int, d_is_one: int = 0, stream: int = 0):
    _check(load().b200_final_exp_bw6_761_device(d_vals, count, d_out or None, d_is_one or None, stream or None))


def groth16_verify_bw6(alpha_g1: np.ndarray, beta_g2: np.ndarray, gamma_g2: np.ndarray, delta_g2: np.ndarray,
                       gamma_abc_g1: np.ndarray, proof_a: np.ndarray, proof_b: np.ndarray, proof_c: np.ndarray,
                       public_inputs: np.ndarray) -> bool:
    """ark-groth16 verify_proof over BW6-761 (b200_groth16_verify_bw6_761).  Points are uint8 records of one
    common stride (200 arkworks / 192 packed), gamma_abc_g1 is [k, stride]; public_inputs uint64 [k - 1, 6]."""
    arrs = [np.ascontiguousarray(a) for a in (alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1, proof_a, proof_b, proof_c)]
    abc = arrs[4].reshape(-1, arrs[4].shape[-1])
    inputs = np.ascontiguousarray(public_inputs, dtype=np.uint64).reshape(-1, 6)
    vk = Groth16Vk(_hptr(arrs[0]), _hptr(arrs[1]), _hptr(arrs[2]), _hptr(arrs[3]), _hptr(abc), len(abc), abc.shape[-1])
    flag = ctypes.c_int(0)
    _check(load().b200_groth16_verify_bw6_761(ctypes.byref(vk), _hptr(arrs[5]), _hptr(arrs[6]), _hptr(arrs[7]),
                                              _hptr(inputs) if len(inputs) else None, len(inputs), ctypes.byref(flag)))
    return bool(flag.value)


POINTS_BLS12_377_G2, POINTS_BW6_761_G1, POINTS_BW6_761_G2 = 0, 1, 2
DECODE_OK, DECODE_INFINITY, DECODE_BAD_COORD, DECODE_NOT_ON_CURVE, DECODE_NOT_IN_SUBGROUP = 0, 1, 2, 3, 4


def deserialize_points(kind: int, data: bytes, check_subgroup: bool = True):
    """Compressed arkworks points (96 bytes each) -> (packed affine records uint8 [n, 192], status int32 [n])
    (b200_deserialize_points; ark-serialize GroupAffine::deserialize on the device)."""
    raw = np.frombuffer(bytes(data), dtype=np.uint8)
    assert len(raw) % 96 == 0
    n = len(raw) // 96
    out = np.zeros((n, 192), dtype=np.uint8)
    status = np.zeros(n, dtype=np.int32)
    _check(load().b200_deserialize_points(kind, _hptr(raw) if n else None, n, int(check_subgroup), _hptr(out) if n else None,
                                          _hptr(status) if n else None))
    return out, status


class EpochBlock:
    """Owner of the buffers an EpochBlockFFI points at (what a cgo caller keeps alive across the call)."""

    def __init__(self, index: int, round_: int, epoch_entropy: Optional[bytes], parent_entropy: Optional[bytes],
                 maximum_non_signers: int, maximum_validators: int, pubkeys: bytes):
        assert len(pubkeys) % 96 == 0
        self._keep = [np.frombuffer(bytes(b), dtype=np.uint8).copy() if b is not None else None
                      for b in (epoch_entropy, parent_entropy, pubkeys)]
        ptr = lambda a: _hptr(a) if a is not None and len(a) else None
        self.ffi = EpochBlockFFI(index, round_, ptr(self._keep[0]), ptr(self._keep[1]), ptr(self._keep[2]),
                                 len(pubkeys) // 96, maximum_non_signers, maximum_validators)


def verify_epochs(vk: bytes, proof: bytes, first: EpochBlock, last: EpochBlock) -> bool:
    """bls-snark-sys `verify` (crates/bls-snark-sys/src/snark/mod.rs:23-45) through the library's export of that
    very symbol: structs by value, bool result."""
    vk_a, proof_a = np.frombuffer(bytes(vk), dtype=np.uint8), np.frombuffer(bytes(proof), dtype=np.uint8)
    return bool(load().verify(_hptr(vk_a), len(vk_a), _hptr(proof_a), len(proof_a), first.ffi, last.ffi))


def verify_epochs_status(vk: bytes, proof: bytes, first: EpochBlock, last: EpochBlock):
    """b200_verify_epochs: (ok, reason) -- engine failures raise."""
    vk_a, proof_a = np.frombuffer(bytes(vk), dtype=np.uint8), np.frombuffer(bytes(proof), dtype=np.uint8)
    ok = ctypes.c_int(0)
    _check(load().b200_verify_epochs(_hptr(vk_a), len(vk_a), _hptr(proof_a), len(proof_a), ctypes.byref(first.ffi),
                                     ctypes.byref(last.ffi), ctypes.byref(ok)))
    return bool(ok.value), load().b200_last_error().decode()


def epoch_public_inputs(first: EpochBlock, last: EpochBlock):
    """b200_epoch_public_inputs: list of canonical scalars (ints), or None when a block does not decode."""
    out = np.zeros((8, 6), dtype=np.uint64)
    count, ok = ctypes.c_size_t(0), ctypes.c_int(0)
    _check(load().b200_epoch_public_inputs(ctypes.byref(first.ffi), ctypes.byref(last.ffi), _hptr(out), 8, ctypes.byref(count),
                                           ctypes.byref(ok)))
    if not ok.value:
        return None
    return [sum(int(v) << (64 * j) for j, v in enumerate(row)) for row in out[:count.value]]


def blake2s_personal(data: bytes, personal: bytes) -> bytes:
    """Host helper of the verifier (no GPU needed): Blake2s-256 with an 8-byte personalisation."""
    assert len(personal) == 8
    d = np.frombuffer(bytes(data), dtype=np.uint8)
    p = np.frombuffer(bytes(personal), dtype=np.uint8)
    out = np.zeros(32, dtype=np.uint8)
    load().b200_blake2s_personal(_hptr(d) if len(d) else None, len(d), _hptr(p), _hptr(out))
    return out.tobytes()


FR_BLS12_377, FR_BW6_761 = 0, 1
FR_BYTES = {0: 32, 1: 48}


HASHER_DIRECT, HASHER_COMPOSITE = 0, 1
HASH_COMPAT, HASH_CIP22, HASH_CRH_ONLY = 1, 2, 4


class HashBatch:
    """The b200_hash_input array of a batch plus its output buffers, built once and reusable across calls (the
    marshalling of thousands of small Python byte strings is host work outside the C-ABI)."""

    def __init__(self, inputs):
        self.inputs = list(inputs)                       # keeps the byte strings alive
        self.n = len(self.inputs)
        self.arr = (HashInput * max(self.n, 1))(*[HashInput(m, len(m), e, len(e)) for m, e in self.inputs])
        self.out = np.zeros(max(self.n, 1) * 144, dtype=np.uint8)
        self.att = np.zeros(max(self.n, 1), dtype=np.uint32)

    def run(self, hasher: int, domain: bytes, flags: int):
        _check(load().b200_hash_to_g1(hasher, flags, domain, len(domain), self.arr, self.n, _hptr(self.out), _hptr(self.att)))

    def results(self):
        return [self.out[144 * i:144 * (i + 1)].tobytes() for i in range(self.n)], self.att[:self.n].tolist()


def hash_to_g1(hasher: int, domain: bytes, inputs, compat: bool = True, cip22: bool = False):
    """HashToCurve::hash for every (message, extra_data) of `inputs` in one launch ->
    (n x 144-byte G1Projective images, attempts)."""
    batch = inputs if isinstance(inputs, HashBatch) else HashBatch(inputs)
    batch.run(hasher, domain, (HASH_COMPAT if compat else 0) | (HASH_CIP22 if cip22 else 0))
    return batch.results()


def hash_crh(hasher: int, domain: bytes, messages):
    """Hasher::crh of each message (composite: 48 bytes; direct: 32 bytes)."""
    n = len(messages)
    arr = (HashInput * max(n, 1))(*[HashInput(m, len(m), b"", 0) for m in messages])
    out = np.zeros(max(n, 1) * 48, dtype=np.uint8)
    _check(load().b200_hash_to_g1(hasher, HASH_CRH_ONLY, domain, len(domain), arr, n, _hptr(out), None))
    size = 48 if hasher == HASHER_COMPOSITE else 32
    return [out[48 * i:48 * i + size].tobytes() for i in range(n)]


def ntt_device(field: int, d_data: int, log_n: int, inverse: bool = False, coset: bool = False, stream: int = 0):
    """In-place radix-2 transform of 2^log_n scalar-field elements (arkworks Montgomery images)."""
    _check(load().b200_ntt_device(field, d_data, log_n, int(inverse), int(coset), stream or None))


def witness_map_device(field: int, d_a: int, d_b: int, d_c: int, log_n: int, d_h: int, stream: int = 0):
    """Groth16 witness-map transform chain: h = (a b - c) / Z as coefficients; a, b, c are clobbered."""
    _check(load().b200_witness_map_device(field, d_a, d_b, d_c, log_n, d_h, stream or None))


GROTH16_BLS12_377, GROTH16_BW6_761 = 0, 1


def groth16_prove_device(family: int, pk: "Groth16Pk", d_assignment: int, num_assign: int, num_aux: int, d_a: int,
                         d_b: int, d_c: int, log_n: int, d_proof: int, stream: int = 0):
    """Groth16 prover arithmetic after synthesis (witness map + 4 MSMs + assembly); d_proof = A | B | C Jacobian."""
    _check(load().b200_groth16_prove_device(family, ctypes.byref(pk), d_assignment or None, num_assign, num_aux, d_a, d_b, d_c,
                                            log_n, d_proof, stream or None))


def groth16_partial_bytes(family: int) -> int:
    """size of one shard's record a_acc | l_acc | h_acc | b_acc"""
    return 3 * 144 + 288 if family == GROTH16_BLS12_377 else 4 * 288


def groth16_prove_partial_device(family: int, pk: "Groth16Pk", d_assignment: int, num_assign: int, num_aux: int, d_a: int,
                                 d_b: int, d_c: int, log_n: int, shard: int, shards: int, d_partials: int, stream: int = 0):
    """One GPU's share of a proof: witness map + its slice of each MSM -> d_partials (groth16_partial_bytes)."""
    _check(load().b200_groth16_prove_partial_device(family, ctypes.byref(pk), d_assignment or None, num_assign, num_aux, d_a,
                                                    d_b, d_c, log_n, shard, shards, d_partials, stream or None))


def groth16_assemble_device(family: int, pk: "Groth16Pk", d_partials: int, shards: int, d_proof: int, stream: int = 0):
    """`shards` gathered partial records (rank-major) -> A | B | C."""
    _check(load().b200_groth16_assemble_device(family, ctypes.byref(pk), d_partials, shards, d_proof, stream or None))


def profile_enable(on: bool = True):
    _check(load().b200_profile_enable(int(on)))


def profile_read():
    """(summed k_bucket_accumulate ms, launches, scalar-point pairs) since the last read."""
    ms, k, pairs = ctypes.c_double(), ctypes.c_int(), ctypes.c_uint64()
    _check(load().b200_profile_read(ctypes.byref(ms), ctypes.byref(k), ctypes.byref(pairs)))
    return ms.value, k.value, pairs.value


def sync(stream: int = 0):
    _check(load().b200_sync(stream or None))
