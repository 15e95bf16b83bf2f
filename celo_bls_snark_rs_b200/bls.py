"""Host-side mirror of the bls-crypto types that call the hot path, over the CUDA library.

Same names, argument meaning and error behaviour as the reference for this path, so the parity
tests read like the reference's own tests (crates/bls-crypto/src/bls/signature.rs:181-426):

  Signature.aggregate / Signature.batch / Signature.batch_verify_hashes   bls/signature.rs:61-155
  PublicKey.aggregate / PublicKey.batch / PublicKey.verify_hash            bls/public.rs:38-120
  Batch.new / add / verify_hash                                            bls/batch.rs:14-84

Message hashing (HashToCurve) is out of scope: every method takes the
message hash as a G1 point, as `batch_verify_hashes` does in the reference.  Objects hold the
Rust types' memory images (G1Projective 144 B, G2Projective 288 B); all arithmetic runs on the
GPU through the C-ABI -- there is no host fallback.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import engine as E

FR_MODULUS = 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001
SECURITY_BOUND = 128


class BLSError(Exception):
    """crates/bls-crypto/src/lib.rs:85-113 (the variants this path can raise)."""


class VerificationFailed(BLSError):
    pass


class UnevenNumKeysMessages(BLSError):
    pass


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


def _sum_points(curve: int, images: Sequence[bytes]) -> bytes:
    jb = E.JAC_BYTES[curve]
    buf = np.frombuffer(b"".join(images), dtype=np.uint8).copy() if images else np.zeros(0, dtype=np.uint8)
    d_in = torch.from_numpy(buf).to(_dev()) if len(images) else torch.zeros(jb, dtype=torch.uint8, device=_dev())
    d_out = torch.empty(jb, dtype=torch.uint8, device=_dev())
    E.sum_jacobian_device(curve, d_in.data_ptr(), len(images), d_out.data_ptr())
    E.sync()
    return d_out.cpu().numpy().tobytes()


def _batch(curve: int, exponents: Sequence[int], images: Sequence[bytes]) -> Optional[bytes]:
    """batch_normalization_into_affine + VariableBaseMSM::multi_scalar_mul (signature.rs:70-89)."""
    if len(images) != len(exponents):
        return None                                  # "takes the min length of the two": refuse
    n = len(images)
    jb, ab = E.JAC_BYTES[curve], E.PACKED_STRIDE[curve]
    sc = np.zeros((n, 4), dtype=np.uint64)
    for i, e in enumerate(exponents):
        e = int(e) % FR_MODULUS
        for j in range(4):
            sc[i, j] = (e >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    dev = _dev()
    d_jac = torch.from_numpy(np.frombuffer(b"".join(images), dtype=np.uint8).copy()).to(dev) if n else \
        torch.zeros(jb, dtype=torch.uint8, device=dev)
    d_aff = torch.empty(max(n, 1) * ab, dtype=torch.uint8, device=dev)
    d_sc = torch.from_numpy(sc.view(np.int64)).to(dev) if n else torch.zeros(4, dtype=torch.int64, device=dev)
    d_out = torch.empty(jb, dtype=torch.uint8, device=dev)
    E.batch_to_affine_device(curve, d_jac.data_ptr(), n, d_aff.data_ptr())
    E.msm_device(curve, d_aff.data_ptr(), d_sc.data_ptr(), n, d_out.data_ptr())
    E.sync()
    return d_out.cpu().numpy().tobytes()


SIG_DOMAIN, POP_DOMAIN = b"ULforxof", b"ULforpop"      # crates/bls-crypto/src/lib.rs:75-78


class HashToG1:
    """HashToCurve<Output = G1Projective> (hash_to_curve/mod.rs:8-17) over the CUDA engine; `hash_many` is the
    batched form the verification flows use (one launch for all messages)."""

    def __init__(self, hasher: int, cip22: bool = False, compat: bool = True):
        self.hasher, self.cip22, self.compat = hasher, cip22, compat

    def hash_many(self, domain: bytes, inputs):
        return E.hash_to_g1(self.hasher, domain, list(inputs), compat=self.compat, cip22=self.cip22)[0]

    def hash(self, domain: bytes, message: bytes, extra_data: bytes = b"") -> bytes:
        return self.hash_many(domain, [(message, extra_data)])[0]

    def hash_with_attempt(self, domain: bytes, message: bytes, extra_data: bytes = b""):
        images, attempts = E.hash_to_g1(self.hasher, domain, [(message, extra_data)], compat=self.compat, cip22=self.cip22)
        return images[0], attempts[0]


# try_and_increment.rs:29-38, try_and_increment_cip22.rs:24-33
COMPOSITE_HASH_TO_G1 = HashToG1(E.HASHER_COMPOSITE)
DIRECT_HASH_TO_G1 = HashToG1(E.HASHER_DIRECT)
COMPOSITE_HASH_TO_G1_CIP22 = HashToG1(E.HASHER_COMPOSITE, cip22=True)


class Signature:
    """A BLS signature on G1 (signature.rs:17): wraps a G1Projective image."""

    def __init__(self, image: bytes):
        assert len(image) == 144
        self.image = bytes(image)

    @staticmethod
    def aggregate(signatures: Sequence["Signature"]) -> "Signature":
        return Signature(_sum_points(E.BLS12_377_G1, [s.image for s in signatures]))

    @staticmethod
    def batch(exponents: Sequence[int], signatures: Sequence["Signature"]) -> Optional["Signature"]:
        out = _batch(E.BLS12_377_G1, exponents, [s.image for s in signatures])
        return None if out is None else Signature(out)

    def batch_verify(self, pubkeys: Sequence["PublicKey"], domain: bytes, messages, hash_to_g1: HashToG1) -> None:
        """signature.rs:101-117: messages = [(message, extra_data)], hashed on the device in one launch."""
        if len(pubkeys) != len(messages):
            raise UnevenNumKeysMessages()
        self.batch_verify_hashes(pubkeys, hash_to_g1.hash_many(domain, messages))

    def batch_verify_hashes(self, pubkeys: Sequence["PublicKey"], message_hashes: Sequence[bytes]) -> None:
        """Raises VerificationFailed / UnevenNumKeysMessages; returns None on success (Ok(()))."""
        if len(pubkeys) != len(message_hashes):
            raise UnevenNumKeysMessages()
        n = len(pubkeys)
        pk = np.frombuffer(b"".join(p.image for p in pubkeys), dtype=np.uint8) if n else np.zeros(1, dtype=np.uint8)
        hs = np.frombuffer(b"".join(message_hashes), dtype=np.uint8) if n else np.zeros(1, dtype=np.uint8)
        sig = np.frombuffer(self.image, dtype=np.uint8)
        ok = ctypes.c_int(0)
        vp = ctypes.c_void_p
        E._check(E.load().b200_batch_verify_hashes(sig.ctypes.data_as(vp), pk.ctypes.data_as(vp),
                                                   hs.ctypes.data_as(vp), n, ctypes.byref(ok)))
        if not ok.value:
            raise VerificationFailed()


class PublicKey:
    """A BLS public key on G2 (public.rs:16): wraps a G2Projective image."""

    def __init__(self, image: bytes):
        assert len(image) == 288
        self.image = bytes(image)

    @staticmethod
    def aggregate(public_keys: Sequence["PublicKey"]) -> "PublicKey":
        return PublicKey(_sum_points(E.BLS12_377_G2, [p.image for p in public_keys]))

    @staticmethod
    def batch(exponents: Sequence[int], public_keys: Sequence["PublicKey"]) -> Optional["PublicKey"]:
        out = _batch(E.BLS12_377_G2, exponents, [p.image for p in public_keys])
        return None if out is None else PublicKey(out)

    def verify(self, message: bytes, extra_data: bytes, signature: Signature, hash_to_g1: HashToG1) -> None:
        """public.rs:70-79 (SIG_DOMAIN)."""
        self.verify_hash(hash_to_g1.hash(SIG_DOMAIN, message, extra_data), signature)

    def verify_pop(self, message: bytes, signature: Signature, hash_to_g1: HashToG1) -> None:
        """public.rs:85-92 (POP_DOMAIN, no extra data)."""
        self.verify_hash(hash_to_g1.hash(POP_DOMAIN, message, b""), signature)

    def verify_hash(self, message_hash: bytes, signature: Signature) -> None:
        """verify_sig after hashing (public.rs:94-120): e(sig, -g2) * e(H, pk) == 1."""
        signature.batch_verify_hashes([self], [message_hash])


def byte_count_from_target_batch_size(size: int, target_security: int = SECURITY_BOUND) -> int:
    """batch.rs:23-28 (ark_std::log2 is the ceiling log)."""
    log2 = (size - 1).bit_length() if size > 1 else 0
    return min((target_security + log2 + 7) // 8, FR_MODULUS.bit_length() // 8)


class Batch:
    """Strict batch verifier context for one message (batch.rs:14-84)."""

    def __init__(self):
        self.entries: List = []

    @staticmethod
    def new() -> "Batch":
        return Batch()

    def add(self, public_key: PublicKey, signature: Signature) -> None:
        self.entries.append((public_key, signature))

    def verify_hash(self, message_hash: bytes, exponents: Optional[Sequence[int]] = None) -> None:
        n = len(self.entries)
        if exponents is None:                        # batch.rs:51-65: exp_size random bytes per entry
            k = byte_count_from_target_batch_size(n)
            exponents = [int.from_bytes(os.urandom(k), "little") for _ in range(n)]
        if len(exponents) != n:
            raise ValueError("Uneven number of exponents and public keys")      # the reference panics here
        sc = np.zeros((max(n, 1), 4), dtype=np.uint64)
        for i, e in enumerate(exponents):
            e = int(e) % FR_MODULUS
            for j in range(4):
                sc[i, j] = (e >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
        pk = np.frombuffer(b"".join(p.image for p, _ in self.entries), dtype=np.uint8) if n else np.zeros(1, np.uint8)
        sg = np.frombuffer(b"".join(s.image for _, s in self.entries), dtype=np.uint8) if n else np.zeros(1, np.uint8)
        mh = np.frombuffer(message_hash, dtype=np.uint8)
        ok = ctypes.c_int(0)
        vp = ctypes.c_void_p
        E._check(E.load().b200_batch_verify_strict_hash(pk.ctypes.data_as(vp), sg.ctypes.data_as(vp),
                                                        sc.ctypes.data_as(vp), n, mh.ctypes.data_as(vp),
                                                        ctypes.byref(ok)))
        if not ok.value:
            raise VerificationFailed()
