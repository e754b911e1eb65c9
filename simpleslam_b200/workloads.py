"""Synthetic benchmark workloads of BASELINE.json, shared by bench.py and the tests.
Bench/test utility: it only produces host arrays (scans, maps, poses); `downsample` is injected by the caller
(the CUDA path in our arm, the CPU oracle in the reference arm — both are bit-identical, tests/test_gpu_voxel.py)."""
import numpy as np
from . import synth

SEED = 1234


def _loop_pose(sc, i, n, cx, cy, radius, z=2.0):
    a = 2 * np.pi * i / n
    return sc.free_pose_near(cx + radius * np.cos(a), cy + radius * np.sin(a), z, a + np.pi / 2)


def _perturb(rng, max_t, max_deg):
    t = rng.uniform(-max_t, max_t, 3) * [1.0, 1.0, 0.15]
    r = np.deg2rad(rng.uniform(-max_deg, max_deg, 3)) * [0.2, 0.2, 1.0]
    return synth.se3_exp(np.concatenate([t, r]))


def c2_ndt(downsample, n_scans, n_map_scans=192, seed_offset=0):
    """C2: 64-beam scan (~130k pts, no source downsample) vs ~2M-point static map (0.2 m downsample of 192 scans on two
    concentric loops), NDT resolution 1.0 / DIRECT7; initial guess = truth o exp(<=0.5 m, <=3 deg)."""
    sc = synth.Scene(seed=SEED, tiles=(2, 2))
    clouds = []
    for i in range(n_map_scans):
        T = _loop_pose(sc, i, n_map_scans, 200.0, 200.0, 100.0 if i % 3 else 45.0)  # two concentric loops
        clouds.append(synth.transform_cloud(T, sc.scan(T, "hdl64", seed=300 + i)))
    raw = np.concatenate(clouds)
    del clouds
    dst = downsample(raw, 0.2)
    rng = np.random.RandomState(17 + seed_offset)
    scans, truths, guesses = [], [], []
    for k in range(n_scans):
        i = (7 + 11 * (k + seed_offset * 1009)) % n_map_scans
        T = _loop_pose(sc, i + 0.37, n_map_scans, 200.0, 200.0, 100.0 + rng.uniform(-1.5, 1.5))
        scans.append(sc.scan(T, "hdl64", seed=5000 + k + 100000 * seed_offset))
        truths.append(T)
        guesses.append(T @ _perturb(rng, 0.5, 3.0))
    return dict(name="C2 NDT scan-to-map: 64-beam scan (~130k pts) vs ~2M-pt map, res 1.0, DIRECT7", method="ndt", dst=dst, scans=scans,
                truths=truths, guesses=guesses, raw_map_points=len(raw))


def c1_loam(downsample, n_scans, seed_offset=0, n_map_scans=200, radius=50.0):
    """C1: VLP-16 scan (28.8k rays, 0.5 m downsample) vs ~200k-point local submap (0.5 m downsample of 200 scans along a
    50 m-radius loop), LOAM; guess = truth o exp([0.3,-0.2,0.05 m; 0.5,-0.5,2 deg]-scale perturbations)."""
    sc = synth.Scene(seed=SEED, tiles=(2, 2))
    clouds, poses = [], []
    for i in range(n_map_scans):
        T = _loop_pose(sc, i, n_map_scans, 200.0, 200.0, radius)
        poses.append(T)
        clouds.append(synth.transform_cloud(T, sc.scan(T, "vlp16", seed=100 + i)))
    dst = downsample(np.concatenate(clouds), 0.5)
    rng = np.random.RandomState(23 + seed_offset)
    scans, truths, guesses = [], [], []
    for k in range(n_scans):
        T = poses[(5 + 7 * (k + 1009 * seed_offset)) % len(poses)] @ synth.se3_exp([0.4, 0.1, 0, 0, 0, 0.01])
        raw = sc.scan(T, "vlp16", seed=7000 + k + 100000 * seed_offset)
        scans.append(downsample(raw, 0.5))
        truths.append(T)
        guesses.append(T @ _perturb(rng, 0.3, 2.0))
    return dict(name="C1 LOAM scan2map: VLP-16 scan (0.5 m downsample) vs ~200k-pt submap", method="loam", dst=dst, scans=scans, truths=truths,
                guesses=guesses)


def c3_vgicp(n_pairs, seed_offset=0):
    """C3: 128-beam scan pairs (~260k pts each) 1.5 m / 3 deg apart, VGICP resolution 1.0."""
    sc = synth.Scene(seed=SEED, tiles=(1, 1))
    rng = np.random.RandomState(31 + seed_offset)
    pairs = []
    for k in range(n_pairs):
        Ta = sc.free_pose_near(80 + 9 * k, 100 + 3 * k, 2.0, 0.2 * k)
        Tb = Ta @ synth.se3_exp([1.5, 0.2, 0.0, 0.0, 0.0, np.deg2rad(3.0)])
        dst = sc.scan(Ta, "os128", seed=9000 + 2 * k)
        src = sc.scan(Tb, "os128", seed=9001 + 2 * k)
        T_true = np.linalg.inv(Ta) @ Tb
        pairs.append(dict(src=src, dst=dst, T_true=T_true, T_guess=T_true @ _perturb(rng, 0.2, 1.0)))
    return dict(name="C3 VGICP scan-to-scan: 128-beam scans (~260k pts each), res 1.0", method="vgicp", pairs=pairs)


def _big_map(sc, extent, spacing, seed=7, strips=16):
    """direct surface sampling of the whole world, strip-parallel (ctypes releases the GIL)"""
    from concurrent.futures import ThreadPoolExecutor
    edges = np.linspace(0.0, extent, strips + 1)
    with ThreadPoolExecutor(max_workers=min(strips, 16)) as ex:
        parts = list(ex.map(lambda k: sc.sample_map(0.0, edges[k], extent, edges[k + 1], spacing, seed=seed + k), range(strips)))
    return np.concatenate(parts)


C4_TILES, C4_SPACING = 4, 0.24


def c4_map(downsample, tiles=C4_TILES, spacing=C4_SPACING, keep_raw=False):
    """the static map of C4: surfaces of tiles x tiles 200 m tiles sampled at `spacing`, voxel-downsampled at 0.2 m (~20M points)"""
    sc = synth.Scene(seed=SEED, tiles=(tiles, tiles))
    raw = _big_map(sc, 200.0 * tiles, spacing)
    dst = downsample(raw, 0.2)
    return (dst, raw) if keep_raw else (dst, len(raw))


def c4_scans(method, downsample, lo, hi, n_total, seed_offset=0, tiles=C4_TILES, workers=8):
    """scans [lo, hi) of the n_total-scan localisation job (every rank generates only its shard; scan k is the same whatever
    the sharding): random poses all over the map, LOAM: VLP-16 scans downsampled at 0.5 m, NDT: raw 64-beam scans,
    guess = truth o small perturbation."""
    from concurrent.futures import ThreadPoolExecutor
    sc = synth.Scene(seed=SEED, tiles=(tiles, tiles))
    extent = 200.0 * tiles
    rng = np.random.RandomState(41 + seed_offset)
    truths, guesses = [], []
    for k in range(n_total):  # the pose stream is drawn for the whole job so that shards agree on it
        T = sc.free_pose_near(rng.uniform(60.0, extent - 60.0), rng.uniform(60.0, extent - 60.0), 2.0, rng.uniform(-np.pi, np.pi))
        G = T @ (_perturb(rng, 0.3, 2.0) if method == "loam" else _perturb(rng, 0.5, 3.0))
        truths.append(T)
        guesses.append(G)
    sensor, base = ("vlp16", 11000) if method == "loam" else ("hdl64", 12000)

    def one(k):
        return np.ascontiguousarray(sc.scan(truths[k], sensor, seed=base + k + 100000 * seed_offset))
    with ThreadPoolExecutor(max_workers=workers) as ex:  # the raycaster releases the GIL
        raw = list(ex.map(one, range(lo, hi)))
    scans = [downsample(r, 0.5) for r in raw] if method == "loam" else raw
    return scans, truths[lo:hi], guesses[lo:hi]


def c4_batched(method, downsample, n_scans, seed_offset=0, tiles=C4_TILES, spacing=C4_SPACING):
    """C4 batched localisation (loc.cpp mode): independent scans at random poses all over a static map of ~20M points
    (4x4 tiles of 200 m, surfaces sampled at 0.24 m then voxel-downsampled at 0.2 m). LOAM: VLP-16 scans downsampled at
    0.5 m; NDT: raw 64-beam scans. Guess = truth o small perturbation."""
    dst, n_raw = c4_map(downsample, tiles, spacing)
    scans, truths, guesses = c4_scans(method, downsample, 0, n_scans, n_scans, seed_offset, tiles)
    name = "C4 batched localisation (%s): independent %s scans vs a %.1fM-pt static map (%dx%d tiles)" % (
        method.upper(), "VLP-16 (0.5 m downsample)" if method == "loam" else "64-beam", len(dst) / 1e6, tiles, tiles)
    return dict(name=name, method=method, dst=dst, scans=scans, truths=truths, guesses=guesses, raw_map_points=n_raw)


def c5_sequence(n_frames, sensor="vlp16", speed=5.0, rate=10.0, seed_offset=0):
    """C5 offline LIO sequence: a figure-8 (lemniscate, ~75 m lobes) driven at `speed` m/s and sampled at `rate` Hz through
    the 2x2-tile world; per frame a raw scan in the sensor frame, a stamp, the ground-truth pose in the map frame (= first
    sensor frame) and a local odometry pose (truth + accumulated wheel/IMU-like noise: 5 mm, 0.02 deg per step)."""
    sc = synth.Scene(seed=SEED, tiles=(2, 2))
    step = speed / rate
    a = 75.0
    # arc-length parametrisation by dense sampling of x = a cos t / (1 + sin^2 t), y = a sin t cos t / (1 + sin^2 t)
    tt = np.linspace(0.0, 2 * np.pi, 200001)
    px = 200.0 + a * np.cos(tt) / (1 + np.sin(tt) ** 2)
    py = 200.0 + a * np.sin(tt) * np.cos(tt) / (1 + np.sin(tt) ** 2)
    seg = np.concatenate([[0.0], np.cumsum(np.hypot(np.diff(px), np.diff(py)))])
    total = seg[-1]
    frames, T0inv, odom = [], None, np.eye(4)
    rng = np.random.RandomState(77 + seed_offset)
    prev_truth = None
    for k in range(n_frames):
        sdist = (k * step) % total
        i = int(np.searchsorted(seg, sdist))
        i = min(max(i, 1), len(tt) - 1)
        yaw = np.arctan2(py[i] - py[i - 1], px[i] - px[i - 1])
        T = sc.free_pose_near(px[i], py[i], 2.0, yaw, step=0.5)
        if T0inv is None:
            T0inv = np.linalg.inv(T)
        truth = T0inv @ T
        if prev_truth is not None:
            noise = synth.se3_exp(np.concatenate([rng.normal(0, 0.005, 3) * [1, 1, 0], np.deg2rad(rng.normal(0, 0.02, 3)) * [0, 0, 1]]))
            odom = odom @ (np.linalg.inv(prev_truth) @ truth) @ noise
        prev_truth = truth
        scan = np.ascontiguousarray(sc.scan(T, sensor, seed=20000 + k + 100000 * seed_offset))
        frames.append(dict(scan=scan, stamp=k / rate, truth=truth, local_odom=odom.copy()))
    return dict(name="C5 offline LIO mapping: %d-frame %s figure-8 at %.0f m/s, %.0f Hz" % (n_frames, sensor, speed, rate), frames=frames)


def shard(n_items, rank, world):
    """contiguous block partition of n_items over `world` ranks (scans i -> contiguous blocks)"""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
