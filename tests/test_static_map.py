"""PCD reader (pcp::loadPCDFile), static-map load (MapManager(pcd_file), MapManager.cpp:52-84) and the
on-disk index cache. The reader needs no GPU; the load / cache round trips are GPU tests."""
import os
import numpy as np
import pytest
import data
from oracle import pyoracle as orc
from simpleslam_b200 import capi


def _cloud(n, seed=0):
    rng = np.random.RandomState(seed)
    pts = np.zeros((n, 8), np.float32)
    pts[:, :3] = rng.uniform(-40, 40, (n, 3)) * [1, 1, 0.1]
    pts[:, 3] = 1.0
    pts[:, 4] = rng.rand(n)
    return pts


def test_pcd_reader_binary_ascii_and_extra_fields(tmp_path):
    pts = _cloud(777)
    for binary in (True, False):
        f = tmp_path / ("b.pcd" if binary else "a.pcd")
        capi.write_pcd(f, pts, binary=binary)
        assert np.array_equal(capi.read_pcd(f), pts)
    # a file with extra / reordered fields (rgb, ring) and CRLF line ends: only x y z intensity are taken
    n = 50
    rec = np.zeros(n, dtype=[("ring", "<u2"), ("x", "<f4"), ("rgb", "<f4"), ("z", "<f4"), ("y", "<f4"), ("intensity", "<f4")])
    rec["x"], rec["y"], rec["z"], rec["intensity"] = pts[:n, 0], pts[:n, 1], pts[:n, 2], pts[:n, 4]
    hdr = ("# .PCD v0.7\r\nVERSION 0.7\r\nFIELDS ring x rgb z y intensity\r\nSIZE 2 4 4 4 4 4\r\nTYPE U F F F F F\r\nCOUNT 1 1 1 1 1 1\r\n"
           "WIDTH %d\r\nHEIGHT 1\r\nVIEWPOINT 0 0 0 1 0 0 0\r\nPOINTS %d\r\nDATA binary\r\n" % (n, n))
    f = tmp_path / "x.pcd"
    f.write_bytes(hdr.encode() + rec.tobytes())
    assert np.array_equal(capi.read_pcd(f), pts[:n])
    # xyz-only cloud: intensity 0
    f2 = tmp_path / "xyz.pcd"
    f2.write_bytes(("FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 3\nHEIGHT 1\nPOINTS 3\nDATA ascii\n1 2 3\n4 5 6\n7 8 9\n").encode())
    r = capi.read_pcd(f2)
    assert np.array_equal(r[:, :3], [[1, 2, 3], [4, 5, 6], [7, 8, 9]]) and (r[:, 4] == 0).all() and (r[:, 3] == 1).all()


def test_pcd_reader_errors(tmp_path):
    with pytest.raises(capi.PcrError):
        capi.read_pcd(tmp_path / "missing.pcd")
    f = tmp_path / "trunc.pcd"
    pts = _cloud(100)
    capi.write_pcd(f, pts)
    raw = f.read_bytes()
    f.write_bytes(raw[:-100])
    with pytest.raises(capi.PcrError):
        capi.read_pcd(f)
    g = tmp_path / "double.pcd"   # float64 coordinates are not PointXYZI
    g.write_bytes(b"FIELDS x y z\nSIZE 8 8 8\nTYPE F F F\nCOUNT 1 1 1\nPOINTS 0\nDATA binary\n")
    with pytest.raises(capi.PcrError):
        capi.read_pcd(g)


@pytest.mark.gpu
@pytest.mark.parametrize("method,case_fn", [(capi.PCR_LOAM, data.loam_case), (capi.PCR_NDT, data.ndt_case)])
def test_static_map_load_and_index_cache(tmp_path, method, case_fn):
    case = case_fn()
    raw_map = case.get("raw_map", case["dst"])
    pcd = tmp_path / "map.pcd"
    capi.write_pcd(pcd, raw_map)
    leaf = 0.5
    ref_map = orc.voxel_downsample(raw_map, leaf)["points"]
    a = capi.Context(method)
    m = a.static_map_load(pcd, leaf)                      # MapManager(pcd_file): load + downsample + register
    assert m == len(ref_map)
    Ta, ca = a.align(case["src"], case["T_guess"])
    b = capi.Context(method)
    Tb, cb = b.scan2map(case["src"], ref_map, case["T_guess"])
    assert ca == cb and np.array_equal(Ta, Tb), "static map loaded from the PCD == the oracle-downsampled map"
    idx = tmp_path / "map.idx"
    a.target_save(idx)                                    # on-disk index: a fresh context starts without rebuilding
    c = capi.Context(method)
    c.target_load(idx)
    Tc, cc = c.align(case["src"], case["T_guess"])
    assert cc == ca and np.array_equal(Tc, Ta)
    assert os.path.getsize(idx) > 16
    with pytest.raises(capi.PcrError):                    # an index of the other method is refused
        other = capi.Context(capi.PCR_NDT if method == capi.PCR_LOAM else capi.PCR_LOAM)
        other.target_load(idx)
    with pytest.raises(capi.PcrError):
        c.static_map_load(tmp_path / "nope.pcd", leaf)
    a.close(); b.close(); c.close()
