"""Host logic of the SNARK verifier entry point, checked on the CPU (no compute calls): the Blake2s the C++ side
uses for the epoch-block edge hashes (crates/epoch-snark/src/epoch_block.rs:226-236, personalisation "ULforout")
against hashlib, and the C layout of EpochBlockFFI (crates/bls-snark-sys/src/snark/epoch_block.rs:109-127)."""
import ctypes
import hashlib

import pytest

from celo_bls_snark_rs_b200 import engine as E


@pytest.mark.parametrize("n", [0, 1, 31, 63, 64, 65, 127, 128, 129, 1000, 14183])
def test_blake2s_matches_hashlib(n):
    data = bytes((i * 131 + 7) & 0xFF for i in range(n))
    for person in (b"ULforout", b"\x00" * 8, b"12345678"):
        assert E.blake2s_personal(data, person) == hashlib.blake2s(data, digest_size=32, person=person).digest()


def test_epoch_block_ffi_layout():
    # #[repr(C)] on x86-64: u16, u8, (pad) 3 pointers, usize, u32, (pad) usize = 56 bytes
    assert ctypes.sizeof(E.EpochBlockFFI) == 56
    offs = {name: getattr(E.EpochBlockFFI, name).offset for name, _ in E.EpochBlockFFI._fields_}
    assert offs == {"index": 0, "round": 2, "epoch_entropy": 8, "parent_entropy": 16, "pubkeys": 24, "pubkeys_num": 32,
                    "maximum_non_signers": 40, "maximum_validators": 48}
