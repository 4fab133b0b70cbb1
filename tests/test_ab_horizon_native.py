"""The bit-parallel value of a horizon-1 alpha-beta node (squava_b200/csrc/ab_horizon.cuh) against a per-child walk
over the line tables, on random boards: the SAME __host__ __device__ code the kernel runs, compiled for the host
(tests/native/ab_horizon_check.cu).  No GPU needed."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc (host-only compile)")
def test_horizon_node_value_equals_per_child_walk(tmp_path):
    exe = str(tmp_path / "ab_horizon_check")
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                           os.path.join(ROOT, "tests", "native", "ab_horizon_check.cu")])
    for seed in (1, 2):
        out = subprocess.run([exe, "150000", str(seed)], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout
        tag, cases, nonzero, terminal = out.stdout.split()
        assert tag == "ok" and int(cases) == 150000 and int(nonzero) > 50000 and int(terminal) > 10000
