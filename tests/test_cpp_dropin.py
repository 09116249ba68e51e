"""The header-only C++ drop-in layer (include/RandBLAS.hh) compiled with the host compiler and linked against
librandblas_b200.so. CPU part: it builds and its host-side checks (state arithmetic, distributions, error
behaviour) pass without a GPU. GPU part: the full program, written like the reference's gtest suites, with plain
host buffers going through the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_dropin")


def _build():
    from randblas_b200 import build as rb_build
    rb_build.build()
    subprocess.check_call(["bash", os.path.join(ROOT, "tests", "cpp", "build.sh")])


def test_cpp_layer_builds_and_host_checks_pass():
    _build()
    r = subprocess.run([EXE, "--host"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


@pytest.mark.gpu
def test_cpp_layer_full_program_on_gpu():
    if not os.path.exists(EXE):
        _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout
