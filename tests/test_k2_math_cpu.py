"""The register-level luma interpolation of K2, checked on the CPU.

broadway_b200/csrc/k2_math.cuh (dp4a horizontal taps, biased 16-bit lane pairs for the vertical taps, one blend
formula for the 16 fractional positions) compiles as plain C++ too; tests/native/k2_math_check.cpp runs it against a
scalar statement of 8.4.2.2.1 (the reference's h264bsd_reconstruct.c:491-1791) over every fractional position, every
superset of the warp-level operand votes, and random / saturating / flat 9x9 windows (123 M sample comparisons)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_luma_interpolation_matches_scalar(tmp_path):
    exe = str(tmp_path / "k2chk")
    subprocess.run(["g++", "-O2", "-I" + os.path.join(ROOT, "broadway_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "native", "k2_math_check.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 mismatches" in r.stdout
