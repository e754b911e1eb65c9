"""Edge cases of the exact k-NN on the nested Morton / hash grid (knn.cuh) against the oracle's brute-force search:
exact ties (duplicates, lattices), fewer points than k, extreme density skew (a tight cluster plus far outliers: the
top-level ring expansion), and the fitness score of clouds that do not overlap (queries outside the target's box)."""
import numpy as np
import pytest
import data
from oracle import pyoracle as orc
from simpleslam_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(capi.PCR_VGICP)
    yield c
    c.close()


def _check(ctx, pts, k):
    cloud = data.xyzi(pts.astype(np.float32))
    _, idx = ctx.gicp_covariances(cloud, k, want_idx=True)
    oi, _ = orc.knn(cloud, cloud[:, :3].astype(np.float64), k, metric_float=True, brute=True, threads=8)
    assert np.array_equal(idx, oi.astype(np.int32))


def test_lattice_ties_and_duplicates(ctx):
    g = np.arange(12, dtype=np.float32) * 0.25
    lattice = np.stack(np.meshgrid(g, g, g[:6], indexing="ij"), -1).reshape(-1, 3)  # every neighbour distance ties many times
    _check(ctx, lattice, 20)
    dup = np.concatenate([lattice[:200], lattice[:200], lattice[:50]])              # exact duplicates: (d2 = 0, index) order
    _check(ctx, dup, 20)
    _check(ctx, dup, 1)
    _check(ctx, lattice + 1000.0, 7)                                                # large coordinates: coarser float grid


def test_fewer_points_than_k(ctx):
    rng = np.random.RandomState(0)
    pts = rng.uniform(-3, 3, (11, 3))
    cloud = data.xyzi(pts.astype(np.float32))
    _, idx = ctx.gicp_covariances(cloud, 20, want_idx=True)
    oi, _ = orc.knn(cloud, cloud[:, :3].astype(np.float64), 20, metric_float=True, brute=True)
    assert np.array_equal(idx, oi.astype(np.int32)) and (idx[:, 11:] == -1).all()
    one = data.xyzi(np.zeros((1, 3), np.float32))
    _, idx1 = ctx.gicp_covariances(one, 20, want_idx=True)
    assert idx1[0, 0] == 0 and (idx1[0, 1:] == -1).all()


def test_density_skew_and_far_outliers(ctx):
    rng = np.random.RandomState(1)
    cluster = rng.normal(0, 0.004, (4000, 3))                      # 4000 points inside a few finest-level cells
    shell = rng.normal(0, 1.0, (3000, 3)) * [20, 20, 2]
    far = rng.uniform(-1, 1, (25, 3)) * [400, 400, 30]             # isolated points: 20-NN radius of tens of metres
    _check(ctx, np.concatenate([cluster, shell, far]), 20)


def test_fitness_of_non_overlapping_clouds():
    rng = np.random.RandomState(2)
    tgt = data.xyzi((rng.uniform(0, 1, (5000, 3)) * [30, 20, 3]).astype(np.float32))
    src = data.xyzi((rng.uniform(0, 1, (2000, 3)) * [30, 20, 3] + [250.0, -80.0, 10.0]).astype(np.float32))  # far outside the target's box
    c = capi.Context(capi.PCR_VGICP, vgicp_max_iters=1)
    c.set_target(tgt)
    T, _ = c.align(src, np.eye(4))
    f = c.fitness()
    Tf = T.astype(np.float32)
    q = np.empty((len(src), 3), np.float32)
    for a in range(3):
        q[:, a] = ((Tf[a, 0] * src[:, 0] + Tf[a, 1] * src[:, 1]) + Tf[a, 2] * src[:, 2]) + Tf[a, 3]
    _, d2 = orc.knn(tgt, q.astype(np.float64), 1, metric_float=True, brute=True, threads=8)
    assert f > 1000 and abs(d2.mean() - f) <= 1e-9 * f
    c.close()
