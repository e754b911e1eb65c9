"""The C++ adaptor (reference's PCR::*Register interface over the C ABI) compiles, links against libpcr_cuda.so and —
on a GPU — registers a synthetic room with all three back ends."""
import os
import subprocess
import pytest
from simpleslam_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_adaptor")


def _build():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    libdir = os.path.dirname(capi.LIB_PATH)
    cmd = [cxx, "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "simpleslam_b200", "cpp"), "-I" + os.path.join(ROOT, "simpleslam_b200", "cpp", "standin"),
           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_adaptor.cpp"), "-o", BIN, "-L" + libdir, "-lpcr_cuda",
           "-Wl,-rpath," + libdir, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"]
    subprocess.check_call(cmd)


def test_adaptor_compiles_links_and_fails_loudly_without_gpu():
    import torch
    _build()
    r = subprocess.run([BIN], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 3 and "no CPU fallback" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_adaptor_registers_on_gpu():
    _build()
    r = subprocess.run([BIN], capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("converged=1") >= 2
