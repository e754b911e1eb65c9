"""A1 voxel-grid downsample: CUDA (through the C-ABI) vs the oracle — bit-exact keys, membership, centroids."""
import os
import numpy as np
import pytest
import data
from oracle import pyoracle as orc
from simpleslam_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(capi.PCR_LOAM)
    yield c
    c.close()


def _check(ctx, pts, leaf):
    o = orc.voxel_downsample(pts, leaf)
    g = ctx.voxel_downsample(pts, leaf)
    d = ctx.debug_voxel()
    assert len(g) == len(o["points"])
    if not o["overflow"]:
        assert np.array_equal(d["keys"], o["keys"])
        assert np.array_equal(d["out_keys"], o["out_keys"])
        assert np.array_equal(d["counts"], o["counts"])
        assert np.array_equal(d["grid"], o["grid"])
    assert np.array_equal(g.view(np.uint32), o["points"].view(np.uint32)), "centroids are not bit-exact"
    return g


@pytest.mark.parametrize("leaf", [0.5, 0.1, 1.0, 0.3])
def test_downsample_scan(ctx, leaf):
    raw = data.loam_case()["raw"].copy()
    raw[:, 4] = np.arange(len(raw), dtype=np.float32) % 97  # exercise the intensity accumulator
    _check(ctx, raw, leaf)


def test_downsample_map_and_golden(ctx):
    _check(ctx, data.loam_case()["raw_map"], 0.5)
    g = np.load(os.path.join(data.GOLDEN, "loam_small.npz"))
    raw = data.xyzi(g["raw"][:, :3], g["raw"][:, 4])
    out = ctx.voxel_downsample(raw, float(g["vd_leaf"]))
    d = ctx.debug_voxel()
    assert np.array_equal(d["keys"], g["vd_keys"]) and np.array_equal(d["counts"], g["vd_counts"])
    assert np.array_equal(out[:, :5].view(np.uint32), g["vd_points"].view(np.uint32))


def test_downsample_strides(ctx):
    raw = data.loam_case()["raw"]
    a = ctx.voxel_downsample(raw, 0.5)                       # 32-byte PointXYZI
    b = ctx.voxel_downsample(np.ascontiguousarray(raw[:, :4]), 0.5)  # 16-byte records (no intensity)
    c = ctx.voxel_downsample(np.ascontiguousarray(raw[:, :3]), 0.5)  # 12-byte records
    assert np.array_equal(a[:, :4], b[:, :4]) and np.array_equal(a[:, :4], c[:, :4])


def test_downsample_edges(ctx):
    assert len(ctx.voxel_downsample(np.zeros((0, 8), np.float32), 0.5)) == 0
    one = data.xyzi(np.array([[1.5, -2.25, 0.125]], np.float32), [7.0])
    assert np.array_equal(_check(ctx, one, 0.5)[0, :5], np.array([1.5, -2.25, 0.125, 1.0, 7.0], np.float32))
    # grid overflow (PCL returns the input unchanged)
    big = data.xyzi(np.array([[0, 0, 0], [3000, 3000, 3000], [1, 2, 3]], np.float32))
    out = _check(ctx, big, 0.5)
    assert np.array_equal(out[:, :3], big[:, :3])
    # ragged: all points in one voxel; points exactly on voxel faces; negative coordinates
    same = data.xyzi(np.tile(np.array([[0.1, 0.2, 0.3]], np.float32), (1000, 1)))
    assert len(_check(ctx, same, 0.5)) == 1
    faces = data.xyzi(np.array([[0, 0, 0], [0.5, 0.5, 0.5], [-0.5, -0.5, -0.5], [1.0, -1.0, 0.0], [0.49999997, 0, 0]], np.float32))
    _check(ctx, faces, 0.5)


def test_downsample_idempotent_membership(ctx):
    """size-independent property at full scan size: every output voxel key is unique and counts sum to n."""
    raw = data.ndt_case()["dst"]
    ctx.voxel_downsample(raw, 0.4)
    d = ctx.debug_voxel()
    assert np.all(np.diff(d["out_keys"].astype(np.int64)) > 0)
    assert d["counts"].sum() == len(raw)
