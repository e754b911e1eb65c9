"""Oracle kNN (grid, (d2, idx) tie-break) vs brute force and vs the REFERENCE'S OWN nanoflann (oracle/_ref) — CPU only."""
import numpy as np
import pytest
from oracle import pyoracle as orc


def _cloud(n, seed, quantise=None):
    rng = np.random.RandomState(seed)
    p = rng.uniform(-20, 20, size=(n, 3)).astype(np.float32)
    p[:, 2] *= 0.1
    if quantise:
        p = (np.round(p / quantise) * quantise).astype(np.float32)  # creates exact distance ties
    out = np.zeros((n, 8), np.float32)
    out[:, :3] = p
    out[:, 3] = 1
    return out


@pytest.mark.parametrize("metric_float", [False, True])
@pytest.mark.parametrize("k", [1, 5, 20])
def test_grid_equals_brute(metric_float, k):
    m = _cloud(4000, 0)
    q = _cloud(300, 1)[:, :3].astype(np.float64)
    gi, gd = orc.knn(m, q, k, metric_float=metric_float, cell=1.0)
    bi, bd = orc.knn(m, q, k, metric_float=metric_float, brute=True)
    assert np.array_equal(gi, bi) and np.array_equal(gd, bd)


def test_ties_resolved_by_index():
    m = _cloud(3000, 2, quantise=0.5)
    q = _cloud(200, 3, quantise=0.25)[:, :3].astype(np.float64)
    gi, gd = orc.knn(m, q, 5, cell=1.0)
    bi, bd = orc.knn(m, q, 5, brute=True)
    assert np.array_equal(gi, bi) and np.array_equal(gd, bd)
    # with ties present, equal distances must come in ascending index order
    for r in range(len(q)):
        for j in range(4):
            if gd[r, j] == gd[r, j + 1]:
                assert gi[r, j] < gi[r, j + 1]


def test_fewer_points_than_k_and_far_queries():
    m = _cloud(3, 4)
    q = np.array([[0.0, 0.0, 0.0], [1e4, -1e4, 50.0]])
    gi, gd = orc.knn(m, q, 5)
    assert (gi[:, 3:] == -1).all() and (gi[:, :3] >= 0).all()
    bi, bd = orc.knn(m, q, 5, brute=True)
    assert np.array_equal(gi, bi)


@pytest.mark.parametrize("metric_float", [False, True])
def test_against_reference_nanoflann(metric_float):
    """Distances must be identical; indices may differ only inside groups of equal distance (nanoflann breaks ties by
    tree visit order, the oracle by index — SURVEY §7 hard part 1)."""
    if orc.ref_lib() is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    for seed, quant in ((5, None), (6, 0.5)):
        m = _cloud(5000, seed, quantise=quant)
        q = _cloud(400, seed + 10)[:, :3].astype(np.float64)
        gi, gd = orc.knn(m, q, 5, metric_float=metric_float)
        ri, rd = orc.ref_knn(m, q, 5, metric_float=metric_float)
        assert np.array_equal(gd, rd)
        diff_rows = np.nonzero((gi != ri).any(1))[0]
        for r in diff_rows:  # differing rows: same multiset of distances, all differences inside tie groups
            for j in range(5):
                if gi[r, j] != ri[r, j]:
                    assert (gd[r] == gd[r, j]).sum() >= 2 or gd[r, j] == gd[r, 4]
        if quant is None:
            assert len(diff_rows) == 0
