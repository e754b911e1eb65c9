"""Inputs the reference's callers can produce and the product must survive the way the reference pipeline does:
NaN / Inf records (the reference strips them in LidarDataProxy.cpp:47 and PCL's voxel grids skip them), the keyframe
cache of pcr_submap_build (keyed by keyframe id, bounded), a failed fitness computation (fails closed, like
pcl::Registration::getFitnessScore's max()), and index blobs / files that do not match the context."""
import os
import numpy as np
import pytest
import torch
import data
from oracle import pyfrontend as opf
from oracle import pyoracle as orc
from simpleslam_b200 import capi, workloads

pytestmark = pytest.mark.gpu


def _with_bad_rows(cloud, seed, n_bad=37):
    """cloud with NaN / +-Inf records spliced in; returns (dirty, mask of the clean rows inside dirty)"""
    rng = np.random.RandomState(seed)
    pos = np.sort(rng.choice(len(cloud) + n_bad, n_bad, replace=False))
    dirty = np.zeros((len(cloud) + n_bad, cloud.shape[1]), np.float32)
    clean = np.ones(len(dirty), bool)
    clean[pos] = False
    dirty[clean] = cloud
    vals = [np.nan, np.inf, -np.inf]
    for k, p in enumerate(pos):
        dirty[p] = cloud[k % len(cloud)]
        dirty[p, k % 3] = vals[(k // 3) % 3]
    return dirty, clean


def test_voxel_downsample_drops_nonfinite_like_removeNaN_plus_VoxelGrid():
    case = data.loam_case()
    dirty, clean = _with_bad_rows(case["raw"], 0)
    c = capi.Context(capi.PCR_LOAM)
    got = c.voxel_downsample(dirty, 0.5)
    ref = orc.voxel_downsample(dirty[clean], 0.5)["points"]   # removeNaNFromPointCloud, then pcl::VoxelGrid
    assert np.isfinite(got).all() and np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    allbad = np.full((5, 8), np.nan, np.float32)
    assert len(c.voxel_downsample(allbad, 0.5)) == 0
    c.close()


@pytest.mark.parametrize("method,case_fn", [(capi.PCR_LOAM, data.loam_case), (capi.PCR_NDT, data.ndt_case), (capi.PCR_VGICP, data.vgicp_case)])
def test_nonfinite_records_in_target_and_source(method, case_fn):
    case = case_fn()
    c = capi.Context(method)
    c.set_target(case["dst"])
    T0, c0 = c.align(case["src"], case["T_guess"])
    f0 = c.fitness() if method == capi.PCR_VGICP else None
    ddst, _ = _with_bad_rows(case["dst"], 1)
    c.set_target(ddst)                                  # dirty map: same index as the clean one
    T1, c1 = c.align(case["src"], case["T_guess"])
    assert c1 == c0 and np.array_equal(T1, T0)
    dsrc, _ = _with_bad_rows(case["src"], 2)
    T2, c2 = c.align(dsrc, case["T_guess"])             # dirty scan: the bad records contribute nothing
    assert c2 == c0 and np.isfinite(T2).all()
    dt, dr = data.pose_err(T2, T0)
    # LOAM / NDT keep the records and skip them (block partition shifts -> rounding), VGICP compacts its staging copy
    assert dt < 1e-6 and dr < 1e-6, (dt, dr)
    if method == capi.PCR_VGICP:
        assert np.array_equal(T2, T0) and c.fitness() == f0
    c.close()


def test_submap_cache_is_keyed_by_id_and_bounded():
    seq = workloads.c5_sequence(30)
    fr = seq["frames"]
    ids = [0, 4, 9, 13, 17, 21]
    clouds = [np.ascontiguousarray(fr[i]["scan"]) for i in ids]
    poses = [fr[i]["truth"] for i in ids]

    def ref(sel):
        return orc.voxel_downsample(np.concatenate([opf.transform_cloud_f32(clouds[k], poses[k]) for k in sel]), 0.5)["points"]

    c = capi.Context(capi.PCR_LOAM)
    pts, _ = c.submap_build(clouds[:4], poses[:4], 0.5, ids=ids[:4])
    assert np.array_equal(pts.view(np.uint32), ref(range(4)).view(np.uint32))
    assert c.submap_cache_info()["entries"] == 4
    # temporaries at fresh addresses under the same ids: served from the cache; a DIFFERENT cloud at a recycled address
    # under a new id must not be confused with anything cached (the old cache keyed by host pointer could be)
    pts, _ = c.submap_build([cl.copy() for cl in clouds[:4]], poses[:4], 0.5, ids=ids[:4])
    assert np.array_equal(pts.view(np.uint32), ref(range(4)).view(np.uint32))
    L = min(len(clouds[0]), len(clouds[4]))
    buf = np.array(clouds[0][:L], copy=True)
    c.submap_build([buf], [poses[0]], 0.5, ids=[100])
    buf[...] = clouds[4][:L]                      # same host address, same count, other content, NEW id
    pts_b, _ = c.submap_build([buf], [poses[4]], 0.5, ids=[101])
    ref_b = orc.voxel_downsample(opf.transform_cloud_f32(buf, poses[4]), 0.5)["points"]
    assert np.array_equal(pts_b.view(np.uint32), ref_b.view(np.uint32))
    # the count under an id changes -> re-uploaded
    half = np.ascontiguousarray(clouds[1][: len(clouds[1]) // 2])
    pts, _ = c.submap_build([clouds[0], half], poses[:2], 0.5, ids=ids[:2])
    exp = orc.voxel_downsample(np.concatenate([opf.transform_cloud_f32(clouds[0], poses[0]), opf.transform_cloud_f32(half, poses[1])]), 0.5)["points"]
    assert np.array_equal(pts.view(np.uint32), exp.view(np.uint32))
    # without ids nothing is cached
    n0 = c.submap_cache_info()["entries"]
    pts, _ = c.submap_build(clouds[4:6], poses[4:6], 0.5)
    assert np.array_equal(pts.view(np.uint32), ref([4, 5]).view(np.uint32)) and c.submap_cache_info()["entries"] == n0
    # a small budget: only the keyframes of the submap being built survive
    c.submap_cache_budget(1)
    pts, _ = c.submap_build(clouds[2:5], poses[2:5], 0.5, ids=ids[2:5])
    assert np.array_equal(pts.view(np.uint32), ref([2, 3, 4]).view(np.uint32))
    assert c.submap_cache_info()["entries"] == 3
    c.submap_cache_clear()
    assert c.submap_cache_info() == dict(bytes=0, entries=0)
    c.close()


def test_fitness_fails_closed():
    case = data.vgicp_case()
    c = capi.Context(capi.PCR_VGICP)
    c.set_target(case["dst"])
    with pytest.raises(capi.PcrError):   # nothing aligned yet
        c.fitness()
    import ctypes
    s = ctypes.c_double(0.25)
    rc = capi.lib().pcr_fitness(c._h, ctypes.byref(s))
    assert rc != 0 and s.value > 1e300, "a failed fitness call must not read as a perfect match (pcl returns max())"
    c.align(case["src"], case["T_guess"])
    assert 0 < c.fitness() < 1.0
    c.close()


@pytest.mark.parametrize("method,case_fn", [(capi.PCR_LOAM, data.loam_case), (capi.PCR_NDT, data.ndt_case), (capi.PCR_VGICP, data.vgicp_case)])
def test_corrupt_or_foreign_index_blobs_are_rejected(method, case_fn, tmp_path):
    case = case_fn()
    a = capi.Context(method)
    a.set_target(case["dst"])
    n = a.target_blob_size()
    blob = torch.empty(n, dtype=torch.uint8, device="cuda")
    a.target_export(blob.data_ptr(), n)
    torch.cuda.synchronize()
    b = capi.Context(method)
    with pytest.raises(capi.PcrError):            # truncated
        b.target_import(blob.data_ptr(), n // 2)
    bad = blob.clone()
    bad[64:72] = 255                              # a section size far beyond the blob
    with pytest.raises(capi.PcrError):
        b.target_import(bad.data_ptr(), n)
    other = capi.Context(capi.PCR_NDT if method != capi.PCR_NDT else capi.PCR_LOAM)
    with pytest.raises(capi.PcrError):            # another back end's blob
        other.target_import(blob.data_ptr(), n)
    # an index built with other parameters is refused instead of silently accepted
    kw = {capi.PCR_LOAM: dict(loam_max_knn_d2=2.0), capi.PCR_NDT: dict(ndt_min_points=9), capi.PCR_VGICP: dict(vgicp_k=12)}[method]
    d = capi.Context(method, **kw)
    with pytest.raises(capi.PcrError):
        d.target_import(blob.data_ptr(), n)
    # files: truncated and garbage
    path = str(tmp_path / "t.idx")
    a.target_save(path)
    raw = open(path, "rb").read()
    open(path, "wb").write(raw[: len(raw) // 3])
    with pytest.raises(capi.PcrError):
        b.target_load(path)
    open(path, "wb").write(b"PCRIDX01" + b"\xff" * 64)
    with pytest.raises(capi.PcrError):
        b.target_load(path)
    b.target_import(blob.data_ptr(), n)           # and the genuine blob still imports
    Ta, _ = a.align(case["src"], case["T_guess"])
    Tb, _ = b.align(case["src"], case["T_guess"])
    assert np.array_equal(Ta, Tb)
    for x in (a, b, other, d):
        x.close()


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["ndt", "loam"])
def test_host_packed_upload_equals_plain_upload(method, monkeypatch):
    """Clouds that come from pageable memory are packed to float4 by the context's host threads (hostpack.hpp) instead of
    being copied raw and packed on the device: same bits on the device, same poses; all record layouts."""
    case = data.ndt_case() if method == "ndt" else data.loam_case()
    mid = capi.PCR_NDT if method == "ndt" else capi.PCR_LOAM
    out = {}
    for force in ("0", "1"):
        monkeypatch.setenv("PCR_HOST_PACK", force)
        ctx = capi.Context(mid)
        ctx.set_target(case["dst"])
        out[force] = ctx.align(case["src"], case["T_guess"])
        # record layouts other than PointXYZI: xyz + padding (16 B), xyz + intensity at float 4 (20 B), bare xyz (12 B)
        xyz = np.ascontiguousarray(case["src"][:, :3])
        for width in (4, 5, 3):
            rec = np.zeros((len(xyz), width), np.float32)
            rec[:, :3] = xyz
            T, conv = ctx.align(rec, case["T_guess"])
            assert np.array_equal(T, out[force][0]) and conv == out[force][1], (force, width)
        ds = ctx.voxel_downsample(case["dst"], 0.5)
        out[force] += (ds,)
    assert np.array_equal(out["0"][0], out["1"][0]) and out["0"][1] == out["1"][1]
    assert np.array_equal(out["0"][2], out["1"][2])


@pytest.mark.gpu
def test_trim_device_cache_hands_buffers_back_and_the_library_keeps_working():
    """pcr_trim_device_cache: buffers parked by destroyed contexts go back to the driver — whole slabs of small buffers only
    once every buffer of the slab is parked — and later calls allocate again with the same results."""
    case = data.loam_case()
    res = []
    for rep in range(2):
        ctx = capi.Context(capi.PCR_LOAM)
        keep = capi.Context(capi.PCR_LOAM)          # holds buffers of the same slabs while the other context goes away
        keep.set_target(case["dst"][::7])
        ctx.set_target(case["dst"])
        res.append(ctx.align(case["src"], case["T_guess"])[0])
        ctx.voxel_downsample(case["src"], 0.5)
        ctx.close()
        freed = capi.trim_device_cache()
        assert freed > 0
        assert np.array_equal(keep.align(case["src"], case["T_guess"])[0], keep.align(case["src"], case["T_guess"])[0])
        keep.close()
        capi.trim_device_cache()
    assert np.array_equal(res[0], res[1])


@pytest.mark.gpu
def test_registered_host_memory_gives_the_same_results():
    """pcr_host_register / pcr_host_unregister: a page-locked map takes the plain-DMA upload path; same index, same pose.
    Registering twice and unregistering memory that is not registered are not errors."""
    case = data.ndt_case()
    dst = np.array(case["dst"], copy=True)
    ctx = capi.Context(capi.PCR_NDT)
    ctx.set_target(dst)
    T0, c0 = ctx.align(case["src"], case["T_guess"])
    capi.host_register(dst)
    capi.host_register(dst)
    try:
        ctx.set_target(dst)
        T1, c1 = ctx.align(case["src"], case["T_guess"])
    finally:
        capi.host_unregister(dst)
        capi.host_unregister(dst)
    assert c0 == c1 and np.array_equal(T0, T1)
