"""tests/cpp/integration_tests.cpp: the reference's tests/integration_tests.rs restated in C++ against
include/veils_stft.hpp (the compiled-language host class over the C ABI).  Host-only tests run
everywhere; the reconstruction / determinism tests need the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "integration_tests.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "integration_tests")


@pytest.fixture(scope="module")
def exe():
    from veils_b200 import build
    build.build()
    deps = [SRC, os.path.join(ROOT, "include", "veils_stft.hpp"), os.path.join(ROOT, "include", "veils_cuda.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-o", EXE, SRC, "-L" + os.path.join(ROOT, "veils_b200"),
                        "-l:libveils_cuda.so", "-Wl,-rpath,$ORIGIN/../../veils_b200"], check=True)
    return EXE


def test_host_only_tests(exe):
    r = subprocess.run([exe, "--host-only"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "5 passed; 0 failed" in r.stdout


@pytest.mark.gpu
def test_all_reference_integration_tests(exe):
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "11 passed; 0 failed" in r.stdout
    for name in ("test_perfect_reconstruction_impulse", "test_different_fft_modes", "test_deterministic_behavior"):
        assert f"test {name} ... ok" in r.stdout
