"""include/cora_b200.hpp (the C++ mirror of CORA::Problem / solveCORA over the C-ABI): compiled with g++
against the built library and run.  CPU: error path only (no CPU fallback); GPU: operators + staircase."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(lib):
    from cora_b200 import capi
    src = os.path.join(ROOT, "tests", "cpp", "test_shim.cpp")
    out = os.path.join(ROOT, "tests", "cpp", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "test_shim")
    libdir = os.path.dirname(capi.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", src, "-o", exe, "-L" + libdir, "-lcora_b200",
           "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64"]
    subprocess.run(cmd, check=True)
    return exe


def test_cpp_shim_compiles_and_fails_loudly_without_gpu(lib):
    from cora_b200 import capi
    exe = _build(lib)
    if capi.device_count() > 0:
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK (cpu)" in r.stdout and "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_cpp_shim_on_gpu(lib):
    exe = _build(lib)
    r = subprocess.run([exe], capture_output=True, text=True)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK (gpu)" in r.stdout
