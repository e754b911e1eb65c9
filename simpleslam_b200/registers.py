"""Python mirror of the reference's PCR plugin interface (same names, argument meaning, error behaviour), used by
the tests and bench.py. The C++ adaptor with the reference's exact virtual interface is simpleslam_b200/cpp/PCR/.

Reference: PCR/include/PCR/PointCloudRegister.hpp:12-38 (abstract `scan2Map(src, dst, res) -> bool`,
`getFitnessScore()`), LoamRegister / NdtRegister / VgicpRegister, and the config-string selection at
frontend/src/LidarOdometry.cpp:32,44-53 (unknown string -> runtime_error).
"""
import numpy as np
from . import capi


class PointCloudRegister:
    """Abstract register. `scan2Map(src, dst, res)`: src/dst are (n, 8) float32 PointXYZI clouds (or (n, 4)/(n, 3)),
    `res` is a 4x4 float64 pose T_map<-scan, refined IN PLACE; returns the reference's `isConverge`."""
    METHOD = None

    def __init__(self, cores=4, device=0, **params):
        self.cores = cores  # cfg["cores"] (PointCloudRegister.hpp:30-31); the GPU path ignores it
        self.isConverge = False
        self._ctx = capi.Context(self.METHOD, device=device, cores=cores, **params)
        self._cache_key = None

    def getFitnessScore(self):
        return 0.0  # PointCloudRegister.hpp:34

    def scan2Map(self, src, dst, res):
        res = np.asarray(res)
        T, conv = self._ctx.scan2map(src, dst, res)
        res[...] = T
        self.isConverge = conv
        return conv

    # localisation mode (test/loc.cpp): static map registered once, many scans aligned against it
    def setStaticMap(self, dst):
        self._ctx.set_target(dst)

    def alignToStaticMap(self, src, res):
        res = np.asarray(res)
        T, conv = self._ctx.align(src, res)
        res[...] = T
        self.isConverge = conv
        return conv

    @property
    def ctx(self):
        return self._ctx


class LoamRegister(PointCloudRegister):
    METHOD = capi.PCR_LOAM


class NdtRegister(PointCloudRegister):
    METHOD = capi.PCR_NDT


class VgicpRegister(PointCloudRegister):
    METHOD = capi.PCR_VGICP

    def initForLC(self):
        self._ctx.init_for_lc()

    def getFitnessScore(self):
        return self._ctx.fitness()


def make_register(pcr_type, **kw):
    """cfg["frontend"]["pcr"] -> register (frontend/src/LidarOdometry.cpp:44-53)."""
    if pcr_type == "loam":
        return LoamRegister(**kw)
    if pcr_type == "ndt":
        return NdtRegister(**kw)
    if pcr_type == "vgicp":
        return VgicpRegister(**kw)
    raise RuntimeError("such pcr type(%s) is not exist, please implemented your self!" % pcr_type)


def voxelDownSample(cloud, grid_size, ctx=None):
    """pcp::voxelDownSample (common/pcp/pcp.hpp:15-28)."""
    own = ctx is None
    if own:
        ctx = capi.Context(capi.PCR_LOAM)
    try:
        return ctx.voxel_downsample(cloud, grid_size)
    finally:
        if own:
            ctx.close()
