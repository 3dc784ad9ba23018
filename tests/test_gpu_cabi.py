"""Runs tests/cabi/test_cabi.c -- a plain C program that drives the C ABI (include/chdb_gpu.h) the way a Rust or C
host does: no Python, no pyarrow between the test and the library."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cabi", "test_cabi")


def build_cabi_test() -> str:
    lib_dir = os.path.join(ROOT, "chapterhouseqe_b200")
    src = os.path.join(ROOT, "tests", "cabi", "test_cabi.c")
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(lib_dir, "libchdb_gpu.so"))):
        subprocess.check_call(["gcc", "-O2", "-Wall", "-Wextra", "-std=c11", "-I", os.path.join(ROOT, "include"), src, "-o", BIN,
                               "-L", lib_dir, "-lchdb_gpu", "-Wl,-rpath," + lib_dir, "-lm"])
    return BIN


def test_cabi_program_builds_and_links():
    """CPU: the C program compiles against the header and links every symbol it uses."""
    assert os.path.exists(build_cabi_test())


@pytest.mark.gpu
def test_cabi_program_passes_on_gpu():
    r = subprocess.run([build_cabi_test()], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout
