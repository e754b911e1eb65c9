"""ScanContext descriptor, keys, descriptor distance and query — GPU (pcr_scancontext_*) against the
numpy restatement (oracle/pyscancontext.py); oracle known-answer checks run on the CPU."""
import numpy as np
import pytest
import data
from oracle import pyscancontext as osc
from oracle import pyoracle as orc


def _ring_cloud():
    # points on circles of known radius / height: bins are known in closed form
    pts = []
    for r, z in [(10.0, 1.0), (39.0, 3.0), (79.0, -0.5), (90.0, 9.0)]:   # 90 m is beyond PC_MAX_RADIUS
        for a in np.deg2rad(np.arange(0.5, 360.0, 1.0)):
            pts.append([r * np.cos(a), r * np.sin(a), z])
    return data.xyzi(np.array(pts, np.float32))


def test_oracle_descriptor_known_answers():
    d = osc.make_scancontext(_ring_cloud(), lidar_height=2.0)
    assert d.shape == (20, 60)
    assert np.allclose(d[2], 3.0) and np.allclose(d[9], 5.0) and np.allclose(d[19], 1.5)   # rings ceil(r / 4): 10 m -> 3, 39 -> 10, 79 -> 20
    assert np.count_nonzero(d) == 3 * 60                                                    # the 90 m circle is dropped
    assert np.allclose(osc.ring_key(d)[[2, 9, 19]], [3.0, 5.0, 1.5]) and np.allclose(osc.sector_key(d), (3.0 + 5.0 + 1.5) / 20)
    # a rotated copy: identical descriptor up to a column shift, found only by the all-shift alignment
    d2 = osc.circshift(d, 17)
    d2[0, 5] = 4.0                                                                           # make the columns distinguishable
    d1 = osc.circshift(d2, -17)
    assert osc.distance(d1, d1)[0] < 1e-12
    assert osc.distance(d2, d1, sector_key_align=True)[1] == 17
    assert osc.distance(d2, d1, sector_key_align=False)[1] in (0, 1, 2, 3, 57, 58, 59)      # the reference never leaves shift 0 +- 3


@pytest.mark.gpu
def test_gpu_descriptor_and_distance_parity():
    from simpleslam_b200 import capi, workloads
    seq = workloads.c5_sequence(8, speed=30.0)
    clouds = [orc.voxel_downsample(f["scan"], 0.5)["points"] for f in seq["frames"]] + [_ring_cloud(), np.zeros((0, 8), np.float32)]
    c = capi.Context(capi.PCR_LOAM)
    desc, rk, sk = c.scancontext_make(clouds, 2.0)
    for k, cl in enumerate(clouds):
        o = osc.make_scancontext(cl, 2.0)
        assert np.array_equal(desc[k], o), "descriptor bins must match bit for bit (k=%d)" % k
        assert np.array_equal(rk[k], osc.ring_key(o)) or np.allclose(rk[k], osc.ring_key(o), rtol=0, atol=1e-15)
        assert np.allclose(sk[k], osc.sector_key(o), rtol=0, atol=1e-15)
    sym = len(clouds) - 2   # the circles: rotation symmetric, its sector key is constant -> every alignment shift ties exactly
    for align in (False, True):
        pairs = [(i, j) for i in range(len(clouds)) for j in range(len(clouds)) if not (align and sym in (i, j))]
        dist, shift = c.scancontext_distance(desc, pairs, 0.1, align)
        for (i, j), dg, sg in zip(pairs, dist, shift):
            do, so = osc.distance(desc[i], desc[j], 0.1, align)
            assert sg == so, (i, j, align)
            assert dg == do or abs(dg - do) <= 1e-12, (i, j, align, dg, do)
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("align", [False, True])
def test_gpu_query_parity_and_revisits(align):
    from simpleslam_b200 import capi, frontend, workloads
    seq = workloads.c5_sequence(80, speed=30.0)           # 3 m between keyframes: the figure-8 crosses itself around keyframe 65
    clouds = [orc.voxel_downsample(f["scan"], 0.5)["points"] for f in seq["frames"]]
    truth = [f["truth"] for f in seq["frames"]]
    clouds.append(clouds[7].copy())                       # an exact revisit of keyframe 7 with the same heading
    truth.append(truth[7])
    c = capi.Context(capi.PCR_LOAM)
    g = frontend.ScanContext(c, sector_key_align=align)
    o = osc.OracleScanContext(sector_key_align=align)
    g.addContext(*clouds)
    for cl in clouds:
        o.add(cl)
    hits = []
    for i in range(len(clouds)):
        qg, qo = g.query(i), o.query(i)
        assert qg[0] == qo[0] and abs(qg[1] - qo[1]) < 1e-12, (i, qg, qo)
        if qg[0] >= 0:
            hits.append((i, qg[0]))
            gap = np.linalg.norm(truth[i][:2, 3] - truth[qg[0]][:2, 3])
            assert gap < 15.0, "a ScanContext match should be a nearby place"
    assert (len(clouds) - 1, 7) in hits                   # the same-heading revisit is found either way (distance 0)
    c.close()
