"""In-tree build of libb200bls.so (sm_100a only).  nvcc cross-compiles without a GPU.

    python -m celo_bls_snark_rs_b200.build [--force] [--verbose]

One translation unit per curve so the heavy kernels compile in parallel; objects are
cached by source mtime under csrc/_obj/.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libb200bls.so")
UNITS = ["engine.cu", "inst_g1_377.cu", "inst_g2_377.cu", "inst_761.cu", "inst_pairing.cu", "inst_bw6_pairing.cu", "inst_epoch_verify.cu", "inst_verify.cu", "inst_ntt.cu", "inst_groth16.cu", "inst_hash.cu", "sys_compat.cu", "sys_compat_keys.cu"]
HEADERS = ["fp.cuh", "ec.cuh", "msm.cuh", "engine.cuh", "curve_impl.cuh", "params_gen.cuh", "pairing.cuh", "pairing_warp.cuh",
           "pairing_params_gen.cuh", "ntt.cuh", "pairing_bw6.cuh", "codec.cuh", "coop.cuh", "pairing_bw6_params_gen.cuh", "msm_afftree.cuh", "pairing_bw6_coop.cuh",
           os.path.join("..", "..", "include", "b200_bls.h"), os.path.join("..", "..", "include", "bls_snark_sys_compat.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _compile(unit: str, force: bool, verbose: bool) -> str:
    src = os.path.join(CSRC, unit)
    obj = os.path.join(OBJ, unit.replace(".cu", ".o"))
    deps = [src] + [os.path.join(CSRC, h) for h in HEADERS]
    if force or _stale(obj, deps):
        cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = res.stdout + res.stderr
        with open(obj + ".log", "w") as f:
            f.write(log)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {unit}:\n{log[-4000:]}")
        if verbose:
            print(log)
    return obj


def build(force: bool = False, verbose: bool = False, crosschecks: bool = False) -> str:
    """crosschecks: also compile the cross-check kernels (one-thread-per-pair pairing, batched-affine accumulation, the
    one-quad Horner ...) that the B200_* environment switches select; the shipped library leaves them out."""
    os.makedirs(OBJ, exist_ok=True)
    marker = os.path.join(OBJ, "crosschecks.flag")
    had = os.path.exists(marker)
    if crosschecks != had:
        force = True
        (open(marker, "w").close() if crosschecks else os.remove(marker))
    if crosschecks and "-DB200_WITH_CROSSCHECKS" not in FLAGS:
        FLAGS.append("-DB200_WITH_CROSSCHECKS")
    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        objs = list(ex.map(lambda u: _compile(u, force, verbose), UNITS))
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, crosschecks="--crosschecks" in sys.argv))
