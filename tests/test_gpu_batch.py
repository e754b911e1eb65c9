"""Batched independent registrations against a static map (loc.cpp pattern) and the target blob
used for the multi-GPU broadcast."""
import numpy as np
import pytest
import torch
import data
from simpleslam_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _scans(case, n, seed, spread=0.3):
    rng = np.random.RandomState(seed)
    srcs, Ts = [], []
    for k in range(n):
        keep = rng.rand(len(case["src"])) < rng.uniform(0.5, 1.0)
        srcs.append(np.ascontiguousarray(case["src"][keep]))
        pert = np.concatenate([rng.uniform(-spread, spread, 3) * [1, 1, 0.2], np.deg2rad(rng.uniform(-2, 2, 3)) * [0.3, 0.3, 1]])
        Ts.append(case["T_true"] @ synth.se3_exp(pert))
    offs = np.concatenate([[0], np.cumsum([len(s) for s in srcs])])
    return srcs, Ts, offs


@pytest.mark.parametrize("method,case_fn", [(capi.PCR_LOAM, data.loam_case), (capi.PCR_NDT, data.ndt_case)])
def test_batch_equals_singles(method, case_fn):
    case = case_fn()
    c = capi.Context(method)
    c.set_target(case["dst"])
    srcs, Ts, offs = _scans(case, 7, 3)
    srcs.insert(3, np.zeros((0, 8), np.float32))  # an empty scan in the middle of the batch
    Ts.insert(3, case["T_guess"])
    offs = np.concatenate([[0], np.cumsum([len(s) for s in srcs])])
    singles = [c.align(s, T) for s, T in zip(srcs, Ts)]
    bT, bconv = c.batch_align(np.concatenate(srcs), offs, Ts)
    for (sT, sconv), T, conv in zip(singles, bT, bconv):
        assert sconv == conv
        # a batch uses fewer blocks per scan than a single registration, so the fixed-order FP64 reductions associate
        # differently: equal to rounding, not bit for bit (each call on its own is deterministic, see below)
        assert np.allclose(sT, T, rtol=0, atol=1e-9), "batched result differs from the single-scan result"
    bT2, _ = c.batch_align(np.concatenate(srcs), offs, Ts)
    assert all(np.array_equal(a, b) for a, b in zip(bT, bT2)), "the same call must be bit-reproducible"
    c.close()


@pytest.mark.parametrize("method,case_fn", [(capi.PCR_LOAM, data.loam_case), (capi.PCR_NDT, data.ndt_case), (capi.PCR_VGICP, data.vgicp_case)])
def test_target_blob_roundtrip(method, case_fn):
    """export the built index into a device blob, import it into a second context: identical registrations"""
    case = case_fn()
    a = capi.Context(method)
    a.set_target(case["dst"])
    nbytes = a.target_blob_size()
    blob = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    a.target_export(blob.data_ptr(), nbytes)
    torch.cuda.synchronize()
    b = capi.Context(method)
    b.target_import(blob.data_ptr(), nbytes)
    Ta, ca = a.align(case["src"], case["T_guess"])
    Tb, cb = b.align(case["src"], case["T_guess"])
    assert ca == cb and np.array_equal(Ta, Tb)
    if method == capi.PCR_VGICP:  # the imported target also carries the kNN grid behind getFitnessScore
        assert a.fitness() == b.fitness()
    a.close(); b.close()


def test_device_resident_inputs():
    case = data.loam_case()
    c = capi.Context(capi.PCR_LOAM)
    dst = torch.from_numpy(case["dst"]).cuda()
    src = torch.from_numpy(case["src"]).cuda()
    torch.cuda.synchronize()
    c.set_target_device(dst.data_ptr(), dst.shape[0], 32)
    T1, c1 = c.align_device(src.data_ptr(), src.shape[0], 32, case["T_guess"])
    c.set_target(case["dst"])
    T2, c2 = c.align(case["src"], case["T_guess"])
    assert c1 == c2 and np.array_equal(T1, T2)
    # device downsample
    raw = torch.from_numpy(case["raw"]).cuda()
    out = torch.empty((raw.shape[0], 8), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    m = c.voxel_downsample_device(raw.data_ptr(), raw.shape[0], 32, 0.5, out.data_ptr(), raw.shape[0])
    assert np.array_equal(out[:m].cpu().numpy(), c.voxel_downsample(case["raw"], 0.5))
    c.close()


@pytest.mark.parametrize("method,case_fn", [(capi.PCR_LOAM, data.loam_case), (capi.PCR_NDT, data.ndt_case), (capi.PCR_VGICP, data.vgicp_case)])
def test_chunked_host_batch_equals_one_piece(method, case_fn):
    """pcr_batch_align cuts a large HOST batch into chunks of whole scans whose uploads overlap the registration of earlier
    chunks (uploader thread + copy stream). Forced here with a 1 MB chunk size: same poses / flags as the one-piece call,
    from pinned and from pageable memory, and getFitnessScore still finds the last scan."""
    import os
    case = case_fn()
    c = capi.Context(method)
    c.set_target(case["dst"])
    srcs, Ts, offs = _scans(case, 12 if method != capi.PCR_VGICP else 8, 5)
    cat = np.ascontiguousarray(np.concatenate(srcs))
    assert cat.nbytes > 2 * (1 << 20)   # at least two 1 MB chunks (the call only chunks batches of >= 2 chunk sizes)
    ref_T, ref_conv = c.batch_align(cat, offs, Ts)
    ref_fit = c.fitness() if method == capi.PCR_VGICP else None
    os.environ["PCR_BATCH_CHUNK_MB"] = "1"
    try:
        pinned = torch.from_numpy(cat).pin_memory().numpy()
        for host in (cat, pinned):
            T, conv = c.batch_align(host, offs, Ts)
            st = c.stats()
            assert st["n_source"] == len(cat)
            for a, b, ca, cb in zip(ref_T, T, ref_conv, conv):
                assert ca == cb and np.allclose(a, b, rtol=0, atol=1e-6 if method == capi.PCR_NDT else 1e-9)
            if method == capi.PCR_VGICP:
                assert c.fitness() == ref_fit
    finally:
        os.environ.pop("PCR_BATCH_CHUNK_MB", None)
    c.close()
