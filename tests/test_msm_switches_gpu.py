"""The kernel forms the MSM picks by size stay bit-exact when forced at small sizes.  Needs a B200.

The engine reads its switches once per process, so each setting runs in a child process:
the child only computes through the C-ABI and prints the compressed result; the oracle stays in this process."""
import json
import os
import subprocess
import sys

import pytest

from oracle import cref as C
from oracle import inputs as H

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import json, sys
import numpy as np
from celo_bls_snark_rs_b200 import engine as E
E.init(0)
job = json.load(open(sys.argv[1]))
out = {}
for key, (cid, bases_path, scalars_path, stride, limbs) in job.items():
    bases = np.fromfile(bases_path, dtype=np.uint8).reshape(-1, stride)
    sc = np.fromfile(scalars_path, dtype=np.uint64).reshape(-1, limbs)
    out[key] = E.msm(cid, bases, sc).hex()
print("RESULT " + json.dumps(out))
"""

CASES = [("bls12_377_g1", "edge", 257), ("bls12_377_g1", "random", 1 << 13), ("bls12_377_g2", "edge", 257),
         ("bls12_377_g2", "random", 1 << 11), ("bw6_761_g1", "edge", 257), ("bw6_761_g1", "random", 1 << 11)]


@pytest.fixture(scope="module")
def workload(tmp_path_factory):
    d = tmp_path_factory.mktemp("switches")
    job, want = {}, {}
    for name, kind, n in CASES:
        L = C.LAYOUTS[name]
        if kind == "edge":
            pts, scalars = H.edge_case_inputs(name, n, 991 + n)
            bases, sc = L.affine_records(pts), L.scalars_array(scalars)
        else:
            bases = L.affine_records(H.random_points(name, n, 313, distinct=128))
            sc = H.random_scalars_array(L, n, 17)
        key = f"{name}:{kind}:{n}"
        bp, sp = d / f"{name}_{kind}.bases", d / f"{name}_{kind}.scalars"
        bases.tofile(bp)
        sc.tofile(sp)
        job[key] = (L.id, str(bp), str(sp), int(bases.strides[0]), int(sc.shape[1]))
        want[key] = L.jacobian_compressed(C.msm(L, bases, sc))
    jp = d / "job.json"
    jp.write_text(json.dumps(job))
    return str(jp), want


def run_child(job_path, env_extra):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), **env_extra)
    r = subprocess.run([sys.executable, "-c", CHILD, job_path], env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


@pytest.mark.parametrize("env_extra", [
    {"B200_REDUCE_THREAD_MIN": "0"},                              # one thread per bucket-reduce segment everywhere
    {"B200_REDUCE_THREAD_MIN": "0", "B200_MSM_SEG": "16"},
    {"B200_REDUCE_THREAD_MIN": "0", "B200_REDUCE_THREAD_CAP": "1"},   # ... with the register cap (more resident blocks)
    {"B200_MSM_ACC_SM": "0"},                                     # register-only accumulate kernel
    {"B200_MSM_ACC_SM": "2"},                                     # shared-memory slots at the higher occupancy
    {"B200_COMBINE": "quad"}, {"B200_COMBINE": "coop"},           # both Horner kernels on every curve
    {"B200_MSM_C": "9"},                                          # a narrow window: many windows, few buckets
], ids=lambda e: ",".join(f"{k[5:]}={v}" for k, v in e.items()))
def test_forced_kernel_forms_match_c_oracle(workload, env_extra):
    job_path, want = workload
    got = run_child(job_path, env_extra)
    for key, w in want.items():
        L = C.LAYOUTS[key.split(":")[0]]
        assert L.jacobian_compressed(bytes.fromhex(got[key])) == w, (key, env_extra)
