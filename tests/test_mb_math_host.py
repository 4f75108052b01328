"""The lean float64 exp / sqrt sequences of the covariance epilogue (mellon_b200/csrc/mb_math.cuh) are written
so that they also compile for the host: this test builds them with g++ and checks them against libm over the
argument ranges the kernels can see (the device build differs only in the reciprocal-square-root seed, which the
GPU tests cover through the kernel parity tests and tools/microbench_fp64 measures directly)."""

import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_lean_exp_and_sqrt_against_libm(tmp_path):
    exe = tmp_path / "mb_math_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(HERE, "host", "mb_math_check.cpp")], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    max_rel_exp, max_rel_sqrt, exp0, exp709, clamp_neg, clamp_zero, clamp_pos = map(float, out)
    assert max_rel_exp < 6e-16          # table + degree-5 polynomial: ~2 ulp
    assert max_rel_sqrt < 2.3e-16       # seed + two Newton steps: correctly rounded on every sample
    assert exp0 == 1.0 and exp709 == 0.0
    assert clamp_neg == 1e-300 and clamp_zero == 1e-300 and clamp_pos == 2e-300
