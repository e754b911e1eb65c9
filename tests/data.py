"""Shared synthetic cases for the tests (deterministic; SURVEY.md §8(d) scaled down so the oracle finishes in seconds)."""
import functools
import os
import numpy as np
from simpleslam_b200 import synth
from oracle import pyoracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

PERTURB = [0.3, -0.2, 0.05, np.deg2rad(0.5), np.deg2rad(-0.5), np.deg2rad(2.0)]


@functools.lru_cache(maxsize=None)
def scene():
    return synth.Scene(seed=1234)


def xyzi(points_xyz, intensity=None):
    n = len(points_xyz)
    out = np.zeros((n, 8), np.float32)
    out[:, :3] = points_xyz
    out[:, 3] = 1.0
    if intensity is not None:
        out[:, 4] = intensity
    return out


@functools.lru_cache(maxsize=None)
def loam_case(n_map_scans=12, sensor="vlp16"):
    """C1-like: one scan (0.5 m downsample) against a submap of merged scans along a 10 m arc (0.5 m downsample)."""
    sc = scene()
    T0 = sc.free_pose_near(100, 100, 2.0, 0.3)
    raw = sc.scan(T0, sensor, seed=1)
    clouds = []
    for i in range(n_map_scans):
        Ti = sc.free_pose_near(100 + i * 0.9 - 5, 100 + 0.3 * i, 2.0, 0.3 + 0.02 * i)
        clouds.append(synth.transform_cloud(Ti, sc.scan(Ti, sensor, seed=100 + i)))
    raw_map = np.concatenate(clouds)
    src = orc.voxel_downsample(raw, 0.5)["points"]
    dst = orc.voxel_downsample(raw_map, 0.5)["points"]
    Tg = T0 @ synth.se3_exp(PERTURB)
    return dict(raw=raw, raw_map=raw_map, src=src, dst=dst, T_true=T0, T_guess=Tg)


@functools.lru_cache(maxsize=None)
def ndt_case(n_map_scans=10):
    """C2-like, scaled: raw VLP-16 scan (no source downsample) vs a 0.2 m-downsampled multi-scan map."""
    sc = scene()
    T0 = sc.free_pose_near(100, 100, 2.0, 0.3)
    src = sc.scan(T0, "vlp16", seed=1)
    clouds = []
    for i in range(n_map_scans):
        Ti = sc.free_pose_near(100 + i * 1.5 - 7, 100 + 0.3 * i, 2.0, 0.3 + 0.02 * i)
        clouds.append(synth.transform_cloud(Ti, sc.scan(Ti, "hdl64", seed=200 + i)))
    dst = orc.voxel_downsample(np.concatenate(clouds), 0.2)["points"]
    Tg = T0 @ synth.se3_exp([0.25, -0.15, 0.03, np.deg2rad(0.4), np.deg2rad(-0.3), np.deg2rad(1.5)])
    return dict(src=src, dst=dst, T_true=T0, T_guess=Tg)


@functools.lru_cache(maxsize=None)
def vgicp_case():
    """C3-like, scaled: two raw VLP-16 scans 1.5 m / 3 deg apart."""
    sc = scene()
    Ta = sc.free_pose_near(100, 100, 2.0, 0.3)
    Tb = Ta @ synth.se3_exp([1.5, 0.2, 0.0, 0.0, 0.0, np.deg2rad(3.0)])
    dst = sc.scan(Ta, "vlp16", seed=11)
    src = sc.scan(Tb, "vlp16", seed=12)
    T_true = np.linalg.inv(Ta) @ Tb  # source -> target frame
    Tg = T_true @ synth.se3_exp([0.2, -0.1, 0.02, 0.0, 0.0, np.deg2rad(1.0)])
    return dict(src=src, dst=dst, T_true=T_true, T_guess=Tg)


def pose_err(Ta, Tb):
    """translation (m) and rotation (rad) distance between two 4x4 poses"""
    dt = float(np.linalg.norm(Ta[:3, 3] - Tb[:3, 3]))
    # chordal form: well conditioned near zero (arccos of the trace loses half the digits there)
    f = np.linalg.norm(Ta[:3, :3] - Tb[:3, :3])
    return dt, float(2.0 * np.arcsin(min(1.0, f / (2.0 * np.sqrt(2.0)))))


def rel_err(a, b):
    """Frobenius / 2-norm relative error, magnitude-normalised (SURVEY §8(d) parity thresholds)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(np.linalg.norm(b), 1e-300)
    return float(np.linalg.norm(a - b) / den)
