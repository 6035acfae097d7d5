"""Solver-tail algebra of the CUDA path on the host (rgbid-slam_b200/csrc/se3.cuh is __host__ __device__): the Cholesky-based
covariance (inverse of the packed SPD normal matrix) against the general Gauss-Jordan inverse, and the packed Cholesky
solve.  No GPU, no oracle."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_spd_inverse_and_packed_solve_match_gauss_jordan(tmp_path):
    src = os.path.join(ROOT, "tests", "cpp", "test_se3_host.cpp")
    exe = str(tmp_path / "test_se3_host")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([gxx, "-O2", "-std=c++17", "-o", exe, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
