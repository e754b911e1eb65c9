"""pcr_multi_*: several contexts / GPUs in one process (loc.cpp mode). On a 1-GPU box every context
sits on device 0 (same code path: blob export, peer copy, import, sharded concurrent batches); with >= 2 GPUs the real
device list is used as well. Results must equal one context registering the whole batch."""
import os
import subprocess
import numpy as np
import pytest
import torch
import data
from simpleslam_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_loc_multi")


def _build():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.check_call([cxx, "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_loc_multi.cpp"), "-o", BIN,
                           "-L" + libdir, "-lpcr_cuda", "-Wl,-rpath," + libdir, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"])


def test_loc_multi_host_builds_and_fails_loudly_without_gpu():
    _build()
    r = subprocess.run([BIN, "loam", "2", "same"], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 3 and "no CPU fallback" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["loam", "ndt"])
def test_cpp_loc_host_uses_several_contexts(method):
    _build()
    runs = [[method, "3", "same"]]
    if torch.cuda.device_count() >= 2:
        runs.append([method, str(min(torch.cuda.device_count(), 8))])
    for a in runs:
        r = subprocess.run([BIN] + a, capture_output=True, text=True)
        print(r.stdout)
        assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("method,case_fn", [(capi.PCR_LOAM, data.loam_case), (capi.PCR_NDT, data.ndt_case)])
def test_multi_equals_single_context(method, case_fn):
    case = case_fn()
    rng = np.random.RandomState(4)
    srcs, Ts = [], []
    for k in range(9):
        keep = rng.rand(len(case["src"])) < rng.uniform(0.5, 1.0)
        srcs.append(np.ascontiguousarray(case["src"][keep]))
        pert = np.concatenate([rng.uniform(-0.3, 0.3, 3) * [1, 1, 0.2], np.deg2rad(rng.uniform(-2, 2, 3)) * [0.3, 0.3, 1]])
        Ts.append(case["T_true"] @ synth.se3_exp(pert))
    offs = np.concatenate([[0], np.cumsum([len(s) for s in srcs])])
    cat = np.concatenate(srcs)
    c = capi.Context(method)
    c.set_target(case["dst"])
    sT, sconv = c.batch_align(cat, offs, Ts)
    c.close()
    devs = [0, 0, 0, 0] if torch.cuda.device_count() < 2 else list(range(min(torch.cuda.device_count(), 4)))
    m = capi.MultiContext(method, devs)
    m.set_target(case["dst"])
    info = m.broadcast_info()
    assert info["blob_bytes"] > 0
    mT, mconv = m.batch_align(cat, offs, Ts)
    for a, b, ca, cb in zip(sT, mT, sconv, mconv):
        # a shard of 2-3 scans partitions the wave differently from the 9-scan batch: equal to rounding
        assert ca == cb and np.allclose(a, b, rtol=0, atol=1e-6 if method == capi.PCR_NDT else 1e-9)
    m.close()


@pytest.mark.gpu
def test_api_calls_leave_the_callers_current_device_alone():
    """A host with CUDA work of its own on another GPU: every entry point switches to the context's device and back."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    case = data.ndt_case()
    torch.cuda.set_device(0)
    ctx = capi.Context(capi.PCR_NDT, device=1)
    assert torch.cuda.current_device() == 0
    ctx.set_target(case["dst"])
    T, conv = ctx.align(case["src"], case["T_guess"])
    assert torch.cuda.current_device() == 0 and conv
    ref = capi.Context(capi.PCR_NDT, device=0)
    ref.set_target(case["dst"])
    T0, _ = ref.align(case["src"], case["T_guess"])
    assert np.array_equal(T, T0)
    del ctx
    assert torch.cuda.current_device() == 0
