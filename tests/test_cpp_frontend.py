"""The C++ headless frontend (simpleslam_b200/cpp/frontend/HeadlessOdometry.hpp over the C ABI) builds everywhere, fails
loudly without a GPU, and on a GPU reproduces the ctypes mirror frame for frame."""
import os
import struct
import subprocess
import numpy as np
import pytest
from simpleslam_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_frontend")


def _build():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    libdir = os.path.dirname(capi.LIB_PATH)
    cmd = [cxx, "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "simpleslam_b200", "cpp"), "-I" + os.path.join(ROOT, "simpleslam_b200", "cpp", "standin"),
           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_frontend.cpp"), "-o", BIN, "-L" + libdir, "-lpcr_cuda",
           "-Wl,-rpath," + libdir, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"]
    subprocess.check_call(cmd)


def _write_frames(path, frames):
    with open(path, "wb") as f:
        f.write(struct.pack("<q", len(frames)))
        for fr in frames:
            f.write(struct.pack("<d", fr["stamp"]))
            f.write(np.ascontiguousarray(fr["local_odom"].T, dtype=np.float64).tobytes())
            sc = np.ascontiguousarray(fr["scan"], dtype=np.float32)
            f.write(struct.pack("<q", len(sc)))
            f.write(sc.tobytes())


def test_cpp_frontend_builds_and_fails_loudly_without_gpu(tmp_path):
    import torch
    _build()
    if torch.cuda.is_available():
        pytest.skip("covered by the gpu test")
    r = subprocess.run([BIN, "loam", str(tmp_path / "none.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_frontend_matches_python_mirror(tmp_path):
    from simpleslam_b200 import frontend, workloads
    _build()
    seq = workloads.c5_sequence(40)
    fin, fout = str(tmp_path / "frames.bin"), str(tmp_path / "poses.bin")
    _write_frames(fin, seq["frames"])
    r = subprocess.run([BIN, "loam", fin, fout], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(fout, "rb").read()
    n = len(seq["frames"])
    rec = np.frombuffer(raw[: n * 17 * 8], dtype=np.float64).reshape(n, 17)
    nk, sm = struct.unpack("<qq", raw[n * 17 * 8:])
    lo = frontend.LidarOdometry("loam")
    for k, f in enumerate(seq["frames"]):
        P = lo.generateOdom(f["scan"], f["stamp"], f["local_odom"])
        Pc = rec[k, :16].reshape(4, 4).T
        assert np.allclose(P, Pc, rtol=0, atol=1e-9), (k, np.abs(P - Pc).max())
        assert bool(np.frombuffer(rec[k, 16:].tobytes(), dtype=np.int64)[0]) == lo.converged[k]
    assert nk == len(lo.map.keyframes) and sm == lo.map.submap_size
    lo.close()
