"""ctypes binding of include/pcr_cuda.h (the C-ABI shared library csrc/libpcr_cuda.so).

The product path has NO CPU fallback: if the CUDA library is missing or no sm_100 device is present, loading /
construction fails loudly (north_star; SURVEY.md §8b "Plugin selection").
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = (os.environ.get("PCR_LIB") or "").strip() or os.path.join(_HERE, "csrc", "libpcr_cuda.so")  # PCR_LIB: a tuning variant built by build.py
_LIB = None

PCR_LOAM, PCR_NDT, PCR_VGICP = 0, 1, 2
PCR_NDT_KDTREE, PCR_NDT_DIRECT26, PCR_NDT_DIRECT7, PCR_NDT_DIRECT1 = 0, 1, 2, 3
PCR_LSQ_LM, PCR_LSQ_GN = 0, 1

# every symbol include/pcr_cuda.h declares (tests check the library exports all of them)
SYMBOLS = [
    "pcr_default_params", "pcr_create", "pcr_destroy", "pcr_last_error", "pcr_vgicp_init_for_lc", "pcr_set_profiling",
    "pcr_set_target", "pcr_set_target_device", "pcr_align", "pcr_align_device", "pcr_scan2map", "pcr_batch_align",
    "pcr_batch_align_device", "pcr_fitness", "pcr_get_stats", "pcr_voxel_downsample", "pcr_voxel_downsample_device", "pcr_downsample_align", "pcr_host_register", "pcr_host_unregister",
    "pcr_target_blob_size", "pcr_target_export", "pcr_target_import", "pcr_debug_voxel", "pcr_loam_linearize",
    "pcr_loam_get_logs", "pcr_ndt_num_leaves", "pcr_ndt_get_leaves", "pcr_ndt_derivatives", "pcr_ndt_hessian",
    "pcr_gicp_covariances", "pcr_vgicp_num_voxels", "pcr_vgicp_get_voxels", "pcr_vgicp_evaluate",
    "pcr_submap_build", "pcr_submap_cache_clear", "pcr_submap_cache_budget", "pcr_submap_cache_info", "pcr_target_save", "pcr_target_load", "pcr_read_pcd", "pcr_static_map_load",
    "pcr_scancontext_make", "pcr_scancontext_distance", "pcr_loam_last_shape",
    "pcr_multi_create", "pcr_multi_destroy", "pcr_multi_last_error", "pcr_multi_set_target", "pcr_multi_batch_align", "pcr_multi_get_broadcast", "pcr_trim_device_cache", "pcr_set_logger",
]


class Params(ctypes.Structure):
    _fields_ = [
        ("method", ctypes.c_int32), ("device", ctypes.c_int32), ("cores", ctypes.c_int32),
        ("loam_max_iters", ctypes.c_int32), ("loam_max_knn_d2", ctypes.c_float), ("loam_plane_thresh", ctypes.c_float),
        ("loam_point_thresh", ctypes.c_float), ("loam_pos_converge", ctypes.c_float), ("loam_rot_converge", ctypes.c_float),
        ("ndt_resolution", ctypes.c_float), ("ndt_search", ctypes.c_int32), ("ndt_max_iters", ctypes.c_int32),
        ("ndt_step_size", ctypes.c_double), ("ndt_outlier_ratio", ctypes.c_double), ("ndt_trans_eps", ctypes.c_double),
        ("ndt_min_points", ctypes.c_int32), ("ndt_eig_mult", ctypes.c_double),
        ("vgicp_resolution", ctypes.c_double), ("vgicp_k", ctypes.c_int32), ("vgicp_max_iters", ctypes.c_int32),
        ("vgicp_optimizer", ctypes.c_int32), ("vgicp_lm_max_iters", ctypes.c_int32), ("vgicp_rot_eps", ctypes.c_double),
        ("vgicp_trans_eps", ctypes.c_double), ("vgicp_lm_init_lambda", ctypes.c_double),
    ]


class Stats(ctypes.Structure):
    _fields_ = [
        ("iterations", ctypes.c_int32), ("evaluations", ctypes.c_int32), ("hessian_evals", ctypes.c_int32), ("converged", ctypes.c_int32),
        ("n_source", ctypes.c_int64), ("n_target", ctypes.c_int64), ("n_residuals", ctypes.c_int64), ("kernel_launches", ctypes.c_int64), ("n_pairs", ctypes.c_int64), ("n_point_evals", ctypes.c_int64),
        ("n_index_reads", ctypes.c_int64),
        ("score", ctypes.c_double), ("ms_total", ctypes.c_float), ("ms_hot_kernel", ctypes.c_float),
        ("hot_kernel_launches", ctypes.c_int32), ("pad", ctypes.c_int32),
        ("ms_aux_kernel", ctypes.c_float), ("aux_kernel_launches", ctypes.c_int32), ("n_aux_items", ctypes.c_int64),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_ if f != "pad"}


class LoamIterLog(ctypes.Structure):
    _fields_ = [("T_before", ctypes.c_double * 16), ("JtJ", ctypes.c_double * 36), ("JtE", ctypes.c_double * 6),
                ("x", ctypes.c_double * 6), ("n", ctypes.c_int64), ("converged", ctypes.c_int32), ("pad", ctypes.c_int32)]


def read_pcd(path):
    """pcr_read_pcd: (n, 8) float32 PointXYZI records of a PCD file (host only, needs no GPU)"""
    n = ctypes.c_size_t(0)
    rc = lib().pcr_read_pcd(str(path).encode(), None, ctypes.c_size_t(0), ctypes.byref(n))
    if rc != 0:
        raise PcrError(rc, "cannot read PCD file %s" % path)
    out = np.empty((max(n.value, 1), 8), np.float32)
    rc = lib().pcr_read_pcd(str(path).encode(), _vp(out), ctypes.c_size_t(n.value), ctypes.byref(n))
    if rc != 0:
        raise PcrError(rc, "truncated PCD file %s" % path)
    return out[:n.value]


def write_pcd(path, pts, binary=True):
    """test / bench utility: write (n, >=5) float32 PointXYZI records as a PCD v0.7 file (x y z intensity)"""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    n = len(pts)
    xyzi = np.ascontiguousarray(np.concatenate([pts[:, :3], pts[:, 4:5]], axis=1))
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
           "WIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA %s\n" % (n, n, "binary" if binary else "ascii"))
    with open(path, "wb") as f:
        f.write(hdr.encode())
        if binary:
            f.write(xyzi.tobytes())
        else:
            for r in xyzi:
                f.write(("%.9g %.9g %.9g %.9g\n" % tuple(float(v) for v in r)).encode())


class PcrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pcr error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Load csrc/libpcr_cuda.so. Raises (never falls back) when the CUDA extension is missing."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("CUDA extension %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.pcr_last_error.restype = ctypes.c_char_p
        L.pcr_last_error.argtypes = [ctypes.c_void_p]
        L.pcr_destroy.argtypes = [ctypes.c_void_p]
        _LIB = L
    return _LIB


def _vp(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return ctypes.c_void_p(a.ctypes.data)
    return ctypes.c_void_p(int(a))


def _cloud(pts):
    a = np.ascontiguousarray(pts, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] < 3:
        raise ValueError("cloud must be (n, >=3) float32")
    return a, a.shape[0], a.shape[1] * 4


def _T_in(T):
    return np.ascontiguousarray(np.asarray(T, dtype=np.float64).T).reshape(16).copy()


def _T_out(buf):
    return np.asarray(buf, dtype=np.float64).reshape(4, 4).T.copy()


def default_params(method, device=0):
    p = Params()
    lib().pcr_default_params(int(method), ctypes.byref(p))
    p.device = device
    return p


LOG_FN = ctypes.CFUNCTYPE(None, ctypes.c_int32, ctypes.c_char_p, ctypes.c_void_p)
_log_keepalive = None


def set_logger(fn):
    """fn(level, message) receives the library's log lines (3 = error, 2 = warning); None restores the stderr default"""
    global _log_keepalive
    if fn is None:
        _log_keepalive = None
        lib().pcr_set_logger(None, None)
        return
    cb = LOG_FN(lambda level, msg, user: fn(int(level), msg.decode(errors="replace")))
    _log_keepalive = cb   # the C side keeps the pointer
    lib().pcr_set_logger(cb, None)


def host_register(a):
    """page-lock a numpy array the caller keeps handing to the library (pcr_host_register)"""
    rc = lib().pcr_host_register(ctypes.c_void_p(a.ctypes.data), ctypes.c_size_t(a.nbytes))
    if rc != 0:
        raise PcrError(rc, "pcr_host_register failed")


def host_unregister(a):
    lib().pcr_host_unregister(ctypes.c_void_p(a.ctypes.data))


def trim_device_cache():
    """hand the process-wide cache of released device buffers back to the driver; returns the bytes freed"""
    b = ctypes.c_size_t(0)
    lib().pcr_trim_device_cache(ctypes.byref(b))
    return b.value


class MultiContext:
    """pcr_multi: one context per device in one process (loc.cpp mode with several GPUs, no Python-side collective)"""

    def __init__(self, method, devices, **overrides):
        p = default_params(method, devices[0])
        for k, v in overrides.items():
            setattr(p, k, v)
        devs = (ctypes.c_int32 * len(devices))(*devices)
        self._h = ctypes.c_void_p()
        L = lib()
        L.pcr_multi_last_error.restype = ctypes.c_char_p
        L.pcr_multi_last_error.argtypes = [ctypes.c_void_p]
        L.pcr_multi_destroy.argtypes = [ctypes.c_void_p]
        rc = L.pcr_multi_create(ctypes.byref(p), devs, ctypes.c_size_t(len(devices)), ctypes.byref(self._h))
        if rc != 0:
            self._h = None
            raise PcrError(rc, L.pcr_last_error(None).decode())

    def close(self):
        if getattr(self, "_h", None):
            lib().pcr_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise PcrError(rc, lib().pcr_multi_last_error(self._h).decode())

    def set_target(self, dst):
        a, n, st = _cloud(dst)
        self._check(lib().pcr_multi_set_target(self._h, _vp(a), ctypes.c_size_t(n), ctypes.c_size_t(st)))

    def batch_align(self, src_concat, offsets, Ts):
        offs = np.ascontiguousarray(offsets, dtype=np.uint64)
        ns = len(offs) - 1
        Tb = np.concatenate([_T_in(T) for T in Ts]) if ns else np.zeros(0)
        conv = np.zeros(max(ns, 1), np.int32)
        a, n, st = _cloud(src_concat)
        self._check(lib().pcr_multi_batch_align(self._h, _vp(a), _vp(offs), ctypes.c_size_t(ns), ctypes.c_size_t(st), _vp(Tb), _vp(conv)))
        return [_T_out(Tb[i * 16:(i + 1) * 16]) for i in range(ns)], conv[:ns].astype(bool)

    def broadcast_info(self):
        b, ms = ctypes.c_size_t(0), ctypes.c_double(0)
        self._check(lib().pcr_multi_get_broadcast(self._h, ctypes.byref(b), ctypes.byref(ms)))
        return dict(blob_bytes=b.value, copy_ms=ms.value)


class Context:
    """One registration context (= one reference register instance): a CUDA stream + device buffers."""

    def __init__(self, method, device=0, params=None, **overrides):
        p = params if params is not None else default_params(method, device)
        p.method = int(method)
        p.device = device
        for k, v in overrides.items():
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
        self.params = p
        self._h = ctypes.c_void_p()
        rc = lib().pcr_create(ctypes.byref(p), ctypes.byref(self._h))
        if rc != 0:
            msg = lib().pcr_last_error(None).decode()
            self._h = None
            raise PcrError(rc, msg)

    def close(self):
        if getattr(self, "_h", None):
            lib().pcr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise PcrError(rc, lib().pcr_last_error(self._h).decode())

    # -- configuration
    def init_for_lc(self):
        self._check(lib().pcr_vgicp_init_for_lc(self._h))

    def set_profiling(self, on=True):
        self._check(lib().pcr_set_profiling(self._h, int(on)))

    def stats(self):
        s = Stats()
        self._check(lib().pcr_get_stats(self._h, ctypes.byref(s)))
        return s.as_dict()

    # -- target / align
    def set_target(self, dst):
        a, n, st = _cloud(dst)
        self._check(lib().pcr_set_target(self._h, _vp(a), ctypes.c_size_t(n), ctypes.c_size_t(st)))

    def set_target_device(self, dev_ptr, n, stride):
        self._check(lib().pcr_set_target_device(self._h, _vp(dev_ptr), ctypes.c_size_t(n), ctypes.c_size_t(stride)))

    def align(self, src, T):
        a, n, st = _cloud(src)
        Tb = _T_in(T)
        conv = ctypes.c_int32(0)
        self._check(lib().pcr_align(self._h, _vp(a), ctypes.c_size_t(n), ctypes.c_size_t(st), _vp(Tb), ctypes.byref(conv)))
        return _T_out(Tb), bool(conv.value)

    def align_device(self, dev_ptr, n, stride, T):
        Tb = _T_in(T)
        conv = ctypes.c_int32(0)
        self._check(lib().pcr_align_device(self._h, _vp(dev_ptr), ctypes.c_size_t(n), ctypes.c_size_t(stride), _vp(Tb), ctypes.byref(conv)))
        return _T_out(Tb), bool(conv.value)

    def scan2map(self, src, dst, T):
        a, n, st = _cloud(src)
        d, m, dt = _cloud(dst)
        Tb = _T_in(T)
        conv = ctypes.c_int32(0)
        self._check(lib().pcr_scan2map(self._h, _vp(a), ctypes.c_size_t(n), ctypes.c_size_t(st), _vp(d), ctypes.c_size_t(m), ctypes.c_size_t(dt),
                                       _vp(Tb), ctypes.byref(conv)))
        return _T_out(Tb), bool(conv.value)

    def batch_align(self, src_concat, offsets, Ts, device_ptr=None, stride=None):
        offs = np.ascontiguousarray(offsets, dtype=np.uint64)
        ns = len(offs) - 1
        Tb = np.concatenate([_T_in(T) for T in Ts]) if ns else np.zeros(0)
        conv = np.zeros(max(ns, 1), np.int32)
        if device_ptr is None:
            a, n, st = _cloud(src_concat)
            self._check(lib().pcr_batch_align(self._h, _vp(a), _vp(offs), ctypes.c_size_t(ns), ctypes.c_size_t(st), _vp(Tb), _vp(conv)))
        else:
            self._check(lib().pcr_batch_align_device(self._h, _vp(device_ptr), _vp(offs), ctypes.c_size_t(ns), ctypes.c_size_t(stride), _vp(Tb),
                                                     _vp(conv)))
        return [_T_out(Tb[i * 16:(i + 1) * 16]) for i in range(ns)], conv[:ns].astype(bool)

    def fitness(self):
        s = ctypes.c_double(0)
        self._check(lib().pcr_fitness(self._h, ctypes.byref(s)))
        return s.value

    # -- voxel downsample
    def voxel_downsample(self, pts, leaf):
        a, n, st = _cloud(pts)
        out = np.empty((max(n, 1), 8), np.float32)
        m = ctypes.c_size_t(0)
        self._check(lib().pcr_voxel_downsample(self._h, _vp(a), ctypes.c_size_t(n), ctypes.c_size_t(st), ctypes.c_float(leaf), _vp(out),
                                               ctypes.c_size_t(n), ctypes.byref(m)))
        self._ds_n = n
        self._ds_m = m.value
        return out[:m.value].copy()

    def downsample_align(self, scan, leaf, T):
        """one frame of LidarOdometry::generateOdom: voxel filter + scan2Map against the resident target, the downsampled
        scan never leaves the device. Returns (pose, converged, points left)."""
        a, n, st = _cloud(scan)
        Tb = _T_in(T)
        conv = ctypes.c_int32(0)
        m = ctypes.c_size_t(0)
        self._check(lib().pcr_downsample_align(self._h, _vp(a), ctypes.c_size_t(n), ctypes.c_size_t(st), ctypes.c_float(leaf), _vp(Tb),
                                               ctypes.byref(conv), ctypes.byref(m)))
        return _T_out(Tb), bool(conv.value), m.value

    def voxel_downsample_device(self, dev_in, n, stride, leaf, dev_out, cap):
        m = ctypes.c_size_t(0)
        self._check(lib().pcr_voxel_downsample_device(self._h, _vp(dev_in), ctypes.c_size_t(n), ctypes.c_size_t(stride), ctypes.c_float(leaf),
                                                      _vp(dev_out), ctypes.c_size_t(cap), ctypes.byref(m)))
        return m.value

    def debug_voxel(self):
        n, m = self._ds_n, self._ds_m
        keys = np.empty(max(n, 1), np.int32); ok = np.empty(max(m, 1), np.int32); oc = np.empty(max(m, 1), np.int32)
        grid = np.zeros(9, np.int32)
        self._check(lib().pcr_debug_voxel(self._h, _vp(keys), _vp(ok), _vp(oc), _vp(grid)))
        return dict(keys=keys[:n], out_keys=ok[:m], counts=oc[:m], grid=grid)

    # -- submap assembly (MapManager::updateMap / loopFindNearKeyframes): transform + concat + downsample on the device;
    # the result becomes the current target. ids (optional): keyframe ids — device copies of the clouds are cached by id
    # (bounded, LRU), never by host address; without ids every cloud is uploaded again
    def submap_build(self, clouds, poses, leaf, want_points=True, ids=None):
        k = len(clouds)
        arrs = [_cloud(cl) for cl in clouds]
        ptrs = (ctypes.c_void_p * max(k, 1))(*[a.ctypes.data for a, _, _ in arrs])
        counts = (ctypes.c_size_t * max(k, 1))(*[n for _, n, _ in arrs])
        idarr = (ctypes.c_int64 * max(k, 1))(*[int(i) for i in ids]) if ids is not None else None
        stride = arrs[0][2] if k else 32
        assert all(st == stride for _, _, st in arrs), "all keyframe clouds must share one record stride"
        P = np.ascontiguousarray(np.concatenate([_T_in(T) for T in poses]) if k else np.zeros(16))
        total = int(sum(n for _, n, _ in arrs))
        out = np.empty((max(total, 1), 8), np.float32) if want_points else None
        m = ctypes.c_size_t(0)
        self._check(lib().pcr_submap_build(self._h, ptrs, counts, idarr, ctypes.c_size_t(k), ctypes.c_size_t(stride), _vp(P), ctypes.c_float(leaf),
                                           _vp(out), ctypes.c_size_t(total), ctypes.byref(m)))
        return (out[:m.value].copy() if want_points else None), m.value

    def submap_cache_budget(self, nbytes):
        self._check(lib().pcr_submap_cache_budget(self._h, ctypes.c_size_t(int(nbytes))))

    def submap_cache_info(self):
        b, e = ctypes.c_size_t(0), ctypes.c_size_t(0)
        self._check(lib().pcr_submap_cache_info(self._h, ctypes.byref(b), ctypes.byref(e)))
        return dict(bytes=b.value, entries=e.value)

    def submap_cache_clear(self):
        self._check(lib().pcr_submap_cache_clear(self._h))

    # -- target blob (multi-GPU)
    def target_blob_size(self):
        b = ctypes.c_size_t(0)
        self._check(lib().pcr_target_blob_size(self._h, ctypes.byref(b)))
        return b.value

    def target_export(self, dev_ptr, cap):
        self._check(lib().pcr_target_export(self._h, _vp(dev_ptr), ctypes.c_size_t(cap)))

    def target_import(self, dev_ptr, nbytes):
        self._check(lib().pcr_target_import(self._h, _vp(dev_ptr), ctypes.c_size_t(nbytes)))

    # -- ScanContext
    def scancontext_make(self, clouds, lidar_height=2.0):
        arrs = [_cloud(cl) for cl in clouds]
        k = len(arrs)
        if k == 0:
            return np.zeros((0, 20, 60)), np.zeros((0, 20)), np.zeros((0, 60))
        stride = arrs[0][2]
        cat = np.ascontiguousarray(np.concatenate([a for a, _, _ in arrs]))
        offs = np.concatenate([[0], np.cumsum([n for _, n, _ in arrs])]).astype(np.uint64)
        desc = np.empty((k, 20, 60)); rk = np.empty((k, 20)); sk = np.empty((k, 60))
        self._check(lib().pcr_scancontext_make(self._h, _vp(cat), _vp(offs), ctypes.c_size_t(k), ctypes.c_size_t(stride), ctypes.c_float(lidar_height),
                                               _vp(desc), _vp(rk), _vp(sk)))
        return desc, rk, sk

    def scancontext_distance(self, descs, pairs, search_ratio=0.1, sector_key_align=False):
        descs = np.ascontiguousarray(descs, dtype=np.float64).reshape(-1, 1200)
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        n = len(pairs)
        dist = np.empty(max(n, 1)); shift = np.empty(max(n, 1), np.int32)
        self._check(lib().pcr_scancontext_distance(self._h, _vp(descs), ctypes.c_size_t(len(descs)), _vp(pairs), ctypes.c_size_t(n),
                                                   ctypes.c_float(search_ratio), ctypes.c_int32(int(sector_key_align)), _vp(dist), _vp(shift)))
        return dist[:n], shift[:n]

    # -- on-disk index cache / static map
    def target_save(self, path):
        self._check(lib().pcr_target_save(self._h, str(path).encode()))

    def target_load(self, path):
        self._check(lib().pcr_target_load(self._h, str(path).encode()))

    def static_map_load(self, pcd_path, leaf):
        m = ctypes.c_size_t(0)
        self._check(lib().pcr_static_map_load(self._h, str(pcd_path).encode(), ctypes.c_float(leaf), ctypes.byref(m)))
        return m.value

    # -- parity / introspection
    def loam_linearize(self, src, T):
        a, n, st = _cloud(src)
        Tb = _T_in(T)
        idx = np.empty((max(n, 1), 5), np.int32); status = np.empty(max(n, 1), np.int32)
        JtJ = np.empty(36); JtE = np.empty(6); nacc = ctypes.c_int64(0)
        self._check(lib().pcr_loam_linearize(self._h, _vp(a), ctypes.c_size_t(n), ctypes.c_size_t(st), _vp(Tb), _vp(idx), _vp(status), _vp(JtJ),
                                             _vp(JtE), ctypes.byref(nacc)))
        return dict(knn_idx=idx[:n], status=status[:n], JtJ=JtJ.reshape(6, 6), JtE=JtE, n=nacc.value)

    def loam_last_shape(self):
        sh = np.zeros(3, np.int32)
        self._check(lib().pcr_loam_last_shape(self._h, _vp(sh)))
        return dict(lpq=int(sh[0]), tile=int(sh[1]), split=bool(sh[2]))

    def loam_logs(self, cap=64):
        logs = (LoamIterLog * cap)()
        n = ctypes.c_int32(0)
        self._check(lib().pcr_loam_get_logs(self._h, logs, cap, ctypes.byref(n)))
        out = []
        for i in range(n.value):
            lg = logs[i]
            out.append(dict(T_before=_T_out(lg.T_before), JtJ=np.array(lg.JtJ).reshape(6, 6), JtE=np.array(lg.JtE), x=np.array(lg.x), n=lg.n,
                            converged=bool(lg.converged)))
        return out

    def ndt_leaves(self):
        n = ctypes.c_size_t(0)
        grid = np.zeros(9, np.int32)
        self._check(lib().pcr_ndt_num_leaves(self._h, ctypes.byref(n), _vp(grid)))
        L = n.value
        keys = np.empty(max(L, 1), np.int32); npts = np.empty(max(L, 1), np.int32)
        mean = np.empty((max(L, 1), 3)); cov = np.empty((max(L, 1), 3, 3)); icov = np.empty((max(L, 1), 3, 3))
        self._check(lib().pcr_ndt_get_leaves(self._h, _vp(keys), _vp(npts), _vp(mean), _vp(cov), _vp(icov)))
        return dict(keys=keys[:L], npts=npts[:L], mean=mean[:L], cov=cov[:L], icov=icov[:L], min_b=grid[:3].copy(), max_b=grid[3:6].copy(),
                    div_b=grid[6:9].copy())

    def ndt_derivatives(self, src, p, compute_hessian=True, Tf=None):
        a, n, st = _cloud(src)
        p = np.ascontiguousarray(p, dtype=np.float64)
        Tfc = np.ascontiguousarray(np.asarray(Tf, dtype=np.float32).T).reshape(16).copy() if Tf is not None else None
        sc = ctypes.c_double(0); g = np.empty(6); H = np.empty(36)
        self._check(lib().pcr_ndt_derivatives(self._h, _vp(a), ctypes.c_size_t(n), ctypes.c_size_t(st), _vp(p), _vp(Tfc), int(compute_hessian),
                                              ctypes.byref(sc), _vp(g), _vp(H)))
        return dict(score=sc.value, g=g, H=H.reshape(6, 6))

    def ndt_hessian(self, src, p):
        a, n, st = _cloud(src)
        p = np.ascontiguousarray(p, dtype=np.float64)
        H = np.empty(36)
        self._check(lib().pcr_ndt_hessian(self._h, _vp(a), ctypes.c_size_t(n), ctypes.c_size_t(st), _vp(p), _vp(H)))
        return H.reshape(6, 6)

    def gicp_covariances(self, pts, k=20, want_idx=False):
        a, n, st = _cloud(pts)
        covs = np.empty((max(n, 1), 3, 3))
        idx = np.empty((max(n, 1), k), np.int32) if want_idx else None
        self._check(lib().pcr_gicp_covariances(self._h, _vp(a), ctypes.c_size_t(n), ctypes.c_size_t(st), k, _vp(covs), _vp(idx)))
        return (covs[:n], idx[:n]) if want_idx else covs[:n]

    def vgicp_voxels(self):
        n = ctypes.c_size_t(0)
        self._check(lib().pcr_vgicp_num_voxels(self._h, ctypes.byref(n)))
        V = n.value
        coords = np.empty((max(V, 1), 3), np.int32); npts = np.empty(max(V, 1), np.int32)
        mean = np.empty((max(V, 1), 3)); cov = np.empty((max(V, 1), 3, 3))
        self._check(lib().pcr_vgicp_get_voxels(self._h, _vp(coords), _vp(npts), _vp(mean), _vp(cov)))
        return dict(coords=coords[:V], npts=npts[:V], mean=mean[:V], cov=cov[:V])

    def vgicp_evaluate(self, src, T0, Ti=None, want_hb=True):
        a, n, st = _cloud(src)
        T0b = _T_in(T0)
        Tib = _T_in(Ti if Ti is not None else T0)
        cost = ctypes.c_double(0); H = np.empty(36) if want_hb else None; b = np.empty(6) if want_hb else None
        nc = ctypes.c_int64(0)
        self._check(lib().pcr_vgicp_evaluate(self._h, _vp(a), ctypes.c_size_t(n), ctypes.c_size_t(st), _vp(T0b), _vp(Tib), ctypes.byref(cost),
                                             _vp(H), _vp(b), ctypes.byref(nc)))
        return dict(cost=cost.value, H=H.reshape(6, 6) if want_hb else None, b=b, n=nc.value)
