"""The packed (two lines per 32-bit register) edge filter of K4, checked on the CPU.

broadway_b200/csrc/k4_simd.cuh compiles as plain C++ too; tests/native/k4_simd_check.cpp runs it against a
scalar statement of 8.7.2.3 / 8.7.2.4 (the reference's h264bsd_deblocking.c:649-1121) over every alpha/tc0
table row, every bS, luma and chroma, random / flat / extreme samples (25 M register comparisons)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_packed_edge_filter_matches_scalar(tmp_path):
    exe = str(tmp_path / "k4chk")
    subprocess.run(["g++", "-O2", "-I" + os.path.join(ROOT, "broadway_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "native", "k4_simd_check.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 mismatches" in r.stdout
