"""Builds and runs the reference's own unit tests restated against the C++ host mirror
(include/bitnuc.hpp -> libbitnuc_cuda.so).  The binary exits non-zero on the first failed check."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _build(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = tmp_path / "test_reference_suite"
    lib_dir = ROOT / "bitnuc_b200"
    subprocess.run([gxx, "-std=c++17", "-O1", "-Wall", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cpp" / "test_reference_suite.cpp"),
                    "-L", str(lib_dir), "-lbitnuc_cuda", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)], check=True)
    return exe


def test_cpp_mirror_compiles_and_refuses_to_run_without_a_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if not torch.cuda.is_available():
        r = subprocess.run([str(exe)], capture_output=True, text=True)
        assert r.returncode != 0 and "bitnuc-cuda" in (r.stdout + r.stderr)  # loud failure, no CPU fallback


@pytest.mark.gpu
def test_reference_suite_through_cpp_mirror(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "reference suite passed" in r.stdout
