"""Synthetic lidar world (bench / test utility). Thin ctypes wrapper over synth.cpp."""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SENSORS = {
    # name: (beams, el_min_deg, el_max_deg, az_steps, max_range)
    "vlp16": (16, -15.0, 15.0, 1800, 100.0),
    "hdl64": (64, -25.0, 3.0, 2048, 120.0),
    "os128": (128, -22.5, 22.5, 2048, 120.0),
}


def build(force=False):
    so = os.path.join(_HERE, "libsynth.so")
    src = os.path.join(_HERE, "synth.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O3", "-std=c++17", "-fPIC", "-fopenmp", "-shared", "-o", so, src])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libsynth.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        L.synth_scene_create.restype = ctypes.c_void_p
        L.synth_scene_create.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int]
        L.synth_scene_destroy.argtypes = [ctypes.c_void_p]
        L.synth_point_free.restype = ctypes.c_int
        L.synth_point_free.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double]
        L.synth_scan.restype = ctypes.c_size_t
        L.synth_scan.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                 ctypes.c_double, ctypes.c_double, ctypes.c_uint64, ctypes.c_void_p]
        L.synth_sample_map.restype = ctypes.c_size_t
        L.synth_sample_map.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_double, ctypes.c_double, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t]
        _LIB = L
    return _LIB


def se3_exp(x):
    """manifolds::exp ordering [rho; omega] (numpy double, used only to build test poses)."""
    x = np.asarray(x, dtype=np.float64)
    rho, w = x[:3], x[3:]
    t = np.linalg.norm(w)
    T = np.eye(4)
    if t < 1e-12:
        T[:3, 3] = rho
        return T
    a = w / t
    ah = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    aa = np.outer(a, a)
    R = np.cos(t) * np.eye(3) + (1 - np.cos(t)) * aa + np.sin(t) * ah
    V = np.sin(t) / t * np.eye(3) + (1 - np.sin(t) / t) * aa + (1 - np.cos(t)) / t * ah
    T[:3, :3] = R
    T[:3, 3] = V @ rho
    return T


def pose_xyz_yaw(x, y, z, yaw):
    T = np.eye(4)
    c, s = np.cos(yaw), np.sin(yaw)
    T[:3, :3] = [[c, -s, 0], [s, c, 0], [0, 0, 1]]
    T[:3, 3] = [x, y, z]
    return T


def transform_cloud(T, pts):
    """float32 transform of PointXYZI records (n,8) by a 4x4 (host utility for map assembly)."""
    out = pts.copy()
    Tf = np.asarray(T, dtype=np.float64)
    xyz = pts[:, :3].astype(np.float64) @ Tf[:3, :3].T + Tf[:3, 3]
    out[:, :3] = xyz.astype(np.float32)
    return out


class Scene:
    def __init__(self, seed=1234, tiles=(1, 1), tile_size=200.0, boxes_per_tile=40, cyls_per_tile=200):
        self.tiles = tiles
        self.tile_size = tile_size
        self._h = _lib().synth_scene_create(seed, tiles[0], tiles[1], tile_size, boxes_per_tile, cyls_per_tile)

    def __del__(self):
        try:
            if self._h:
                _lib().synth_scene_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def is_free(self, x, y, z=2.0, margin=1.5):
        return bool(_lib().synth_point_free(self._h, x, y, z, margin))

    def scan(self, T, sensor="vlp16", noise=0.02, seed=0):
        beams, el0, el1, az, rng = SENSORS[sensor]
        out = np.empty((beams * az, 8), dtype=np.float32)
        Tc = np.ascontiguousarray(np.asarray(T, dtype=np.float64).T)  # column-major
        n = _lib().synth_scan(self._h, Tc.ctypes.data, beams, el0, el1, az, rng, noise, seed, out.ctypes.data)
        return out[:n].copy()

    def sample_map(self, x0, y0, x1, y1, spacing, noise=0.02, seed=7):
        n = _lib().synth_sample_map(self._h, x0, y0, x1, y1, spacing, noise, seed, None, 0)
        out = np.empty((n, 8), dtype=np.float32)
        m = _lib().synth_sample_map(self._h, x0, y0, x1, y1, spacing, noise, seed, out.ctypes.data, n)
        assert m == n
        return out

    def free_pose_near(self, x, y, z=2.0, yaw=0.0, step=1.0):
        """nearest free-space sensor pose on a small spiral around (x, y)."""
        k = 0
        while not self.is_free(x, y, z):
            k += 1
            x += step * np.cos(k * 2.4) * np.sqrt(k)
            y += step * np.sin(k * 2.4) * np.sqrt(k)
        return pose_xyz_yaw(x, y, z, yaw)
