"""ctypes front-end of the CPU oracle (oracle/pcr_oracle.h). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
PARITY UNPINNED by reference fixtures (none exist); kNN is pinned against the reference's own nanoflann (oracle/_ref).
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

c_f = ctypes.c_void_p


def build(force=False):
    so = os.path.join(_HERE, "libpcr_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("orc_loam.cpp", "orc_ndt.cpp", "orc_vgicp.cpp", "orc_common.hpp", "orc_linalg.hpp", "pcr_oracle.h")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "libpcr_oracle.so"], stdout=subprocess.DEVNULL)
    ref_so = os.path.join(_HERE, "_ref", "libref_nanoflann.so")
    if os.path.isdir("/root/reference/third_parties/nanoflann") and (force or not os.path.exists(ref_so)):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return so


class LoamIterLog(ctypes.Structure):
    _fields_ = [("T_before", ctypes.c_double * 16), ("JtJ", ctypes.c_double * 36), ("JtE", ctypes.c_double * 6),
                ("x", ctypes.c_double * 6), ("n", ctypes.c_int64), ("converged", ctypes.c_int32), ("pad", ctypes.c_int32)]


class NdtResult(ctypes.Structure):
    _fields_ = [("T", ctypes.c_double * 16), ("trans_probability", ctypes.c_double), ("converged", ctypes.c_int32),
                ("nr_iterations", ctypes.c_int32), ("n_derivative_evals", ctypes.c_int32), ("n_hessian_evals", ctypes.c_int32),
                ("p_final", ctypes.c_double * 6)]


class VgicpResult(ctypes.Structure):
    _fields_ = [("T", ctypes.c_double * 16), ("converged", ctypes.c_int32), ("nr_iterations", ctypes.c_int32),
                ("n_linearize", ctypes.c_int32), ("n_error_evals", ctypes.c_int32)]


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libpcr_oracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        L.orc_ndt_create.restype = ctypes.c_void_p
        L.orc_vgicp_create.restype = ctypes.c_void_p
        L.orc_ndt_num_leaves.restype = ctypes.c_size_t
        L.orc_vgicp_num_voxels.restype = ctypes.c_size_t
        for f in ("orc_ndt_derivatives", "orc_ndt_derivatives_T", "orc_vgicp_linearize", "orc_vgicp_error", "orc_fitness"):
            getattr(L, f).restype = ctypes.c_double
        _LIB = L
    return _LIB


def ref_lib():
    """The reference's own nanoflann compiled into oracle/_ref (None when not built)."""
    global _REF
    if _REF is None:
        so = os.path.join(_HERE, "_ref", "libref_nanoflann.so")
        if not os.path.exists(so):
            return None
        L = ctypes.CDLL(so)
        L.refnf_create_f64.restype = ctypes.c_void_p
        L.refnf_knn_f64.restype = ctypes.c_size_t
        _REF = L
    return _REF


def _pts(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2
    return a, a.shape[0], a.shape[1]


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def _T(T):
    """4x4 row-major numpy -> column-major double[16] buffer"""
    return np.ascontiguousarray(np.asarray(T, dtype=np.float64).T).reshape(16)


def _Tback(buf):
    return np.array(buf, dtype=np.float64).reshape(4, 4).T.copy()


def voxel_downsample(pts, leaf):
    a, n, st = _pts(pts)
    keys = np.empty(n, np.int32)
    out = np.empty((max(n, 1), 8), np.float32)
    okeys = np.empty(max(n, 1), np.int32)
    ocnt = np.empty(max(n, 1), np.int32)
    m = ctypes.c_size_t(0)
    grid = np.zeros(9, np.int32)
    rc = lib().orc_voxel_downsample(_p(a), ctypes.c_size_t(n), ctypes.c_size_t(st), ctypes.c_float(leaf), _p(keys), _p(out), _p(okeys),
                                    _p(ocnt), ctypes.byref(m), _p(grid))
    m = m.value
    return dict(points=out[:m].copy(), keys=keys, out_keys=okeys[:m].copy(), counts=ocnt[:m].copy(), overflow=bool(rc), grid=grid)


def knn(map_pts, queries, k, metric_float=False, brute=False, cell=1.0, threads=8):
    a, n, st = _pts(map_pts)
    q = np.ascontiguousarray(queries, dtype=np.float64)
    nq = q.shape[0]
    idx = np.empty((nq, k), np.int64)
    d2 = np.empty((nq, k), np.float64)
    lib().orc_knn(_p(a), ctypes.c_size_t(n), ctypes.c_size_t(st), _p(q), ctypes.c_size_t(nq), k, int(metric_float), int(brute),
                  ctypes.c_float(cell), threads, _p(idx), _p(d2))
    return idx, d2


def neighbourhood27(map_pts, queries, cell=1.0, threads=8):
    """C-bar of SURVEY.md §8(d): mean number of map points in the 27 cells around a query"""
    a, n, st = _pts(map_pts)
    q = np.ascontiguousarray(queries, dtype=np.float64)
    L = lib()
    L.orc_neighbourhood27.restype = ctypes.c_double
    return L.orc_neighbourhood27(_p(a), ctypes.c_size_t(n), ctypes.c_size_t(st), _p(q), ctypes.c_size_t(q.shape[0]), ctypes.c_float(cell), threads)


def ref_knn(map_pts, queries, k, metric_float=False):
    L = ref_lib()
    if L is None:
        return None
    a, n, st = _pts(map_pts)
    q = np.ascontiguousarray(queries, dtype=np.float64)
    nq = q.shape[0]
    idx = np.empty((nq, k), np.int64)
    d2 = np.empty((nq, k), np.float64)
    f = L.refnf_knn_batch_f32 if metric_float else L.refnf_knn_batch_f64
    f(_p(a), ctypes.c_size_t(n), ctypes.c_size_t(st), _p(q), ctypes.c_size_t(nq), k, _p(idx), _p(d2))
    return idx, d2


def loam_linearize(src, dst, T, threads=8):
    s, ns, ss = _pts(src)
    d, nm, ds = _pts(dst)
    Tc = _T(T)
    idx = np.empty((ns, 5), np.int64)
    d2 = np.empty((ns, 5), np.float64)
    status = np.empty(ns, np.int32)
    resid = np.empty(ns, np.float64)
    J = np.empty((ns, 6), np.float64)
    JtJ = np.empty(36, np.float64)
    JtE = np.empty(6, np.float64)
    n = ctypes.c_int64(0)
    lib().orc_loam_linearize(_p(s), ctypes.c_size_t(ns), ctypes.c_size_t(ss), _p(d), ctypes.c_size_t(nm), ctypes.c_size_t(ds), _p(Tc), threads,
                             _p(idx), _p(d2), _p(status), _p(resid), _p(J), _p(JtJ), _p(JtE), ctypes.byref(n))
    return dict(knn_idx=idx, knn_d2=d2, status=status, resid=resid, J=J, JtJ=JtJ.reshape(6, 6), JtE=JtE, n=n.value)


def loam_align(src, dst, T, threads=8, max_iters=8):
    s, ns, ss = _pts(src)
    d, nm, ds = _pts(dst)
    Tc = _T(T)
    logs = (LoamIterLog * max_iters)()
    nit = ctypes.c_int32(0)
    conv = ctypes.c_int32(0)
    lib().orc_loam_align(_p(s), ctypes.c_size_t(ns), ctypes.c_size_t(ss), _p(d), ctypes.c_size_t(nm), ctypes.c_size_t(ds), _p(Tc), threads,
                         max_iters, logs, ctypes.byref(nit), ctypes.byref(conv))
    its = []
    for i in range(nit.value):
        lg = logs[i]
        its.append(dict(T_before=_Tback(lg.T_before), JtJ=np.array(lg.JtJ).reshape(6, 6), JtE=np.array(lg.JtE), x=np.array(lg.x), n=lg.n,
                        converged=bool(lg.converged)))
    return dict(T=_Tback(Tc), iters=its, converged=bool(conv.value))


def se3_exp(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    T = np.empty(16, np.float64)
    lib().orc_se3_exp(_p(x), _p(T))
    return _Tback(T)


def t2se3(T):
    Tc = _T(T)
    lib().orc_t2se3(_p(Tc))
    return _Tback(Tc)


class Ndt:
    SEARCH = {"KDTREE": 0, "DIRECT26": 1, "DIRECT7": 2, "DIRECT1": 3}

    def __init__(self, dst, resolution=1.0):
        self.d, nm, ds = _pts(dst)
        self.h = ctypes.c_void_p(lib().orc_ndt_create(_p(self.d), ctypes.c_size_t(nm), ctypes.c_size_t(ds), ctypes.c_float(resolution)))

    def __del__(self):
        try:
            lib().orc_ndt_destroy(self.h)
        except Exception:
            pass

    def leaves(self):
        grid = np.zeros(9, np.int32)
        n = lib().orc_ndt_num_leaves(self.h, _p(grid))
        keys = np.empty(n, np.int32); npts = np.empty(n, np.int32)
        mean = np.empty((n, 3)); cov = np.empty((n, 3, 3)); icov = np.empty((n, 3, 3))
        lib().orc_ndt_get_leaves(self.h, _p(keys), _p(npts), _p(mean), _p(cov), _p(icov))
        return dict(keys=keys, npts=npts, mean=mean, cov=cov, icov=icov, min_b=grid[:3].copy(), max_b=grid[3:6].copy(), div_b=grid[6:9].copy())

    def derivatives(self, src, p, search="DIRECT7", compute_hessian=True, threads=8, Tf=None):
        s, ns, ss = _pts(src)
        p = np.ascontiguousarray(p, dtype=np.float64)
        g = np.empty(6); H = np.empty(36); nb = np.empty(ns, np.int32)
        if Tf is None:
            sc = lib().orc_ndt_derivatives(self.h, _p(s), ctypes.c_size_t(ns), ctypes.c_size_t(ss), _p(p), self.SEARCH[search],
                                           int(compute_hessian), threads, _p(g), _p(H), _p(nb))
        else:
            Tfc = np.ascontiguousarray(np.asarray(Tf, dtype=np.float32).T).reshape(16)
            sc = lib().orc_ndt_derivatives_T(self.h, _p(s), ctypes.c_size_t(ns), ctypes.c_size_t(ss), _p(Tfc), _p(p), self.SEARCH[search],
                                             int(compute_hessian), threads, _p(g), _p(H), _p(nb))
        return dict(score=sc, g=g, H=H.reshape(6, 6), nb=nb)

    def hessian(self, src, p, search="DIRECT7"):
        s, ns, ss = _pts(src)
        p = np.ascontiguousarray(p, dtype=np.float64)
        H = np.empty(36)
        lib().orc_ndt_hessian(self.h, _p(s), ctypes.c_size_t(ns), ctypes.c_size_t(ss), _p(p), self.SEARCH[search], _p(H))
        return H.reshape(6, 6)

    def align(self, src, T, search="DIRECT7", threads=8, max_iterations=35, trans_eps=0.1, step_size=0.1):
        s, ns, ss = _pts(src)
        Tc = _T(T)
        res = NdtResult()
        lib().orc_ndt_align(self.h, _p(s), ctypes.c_size_t(ns), ctypes.c_size_t(ss), _p(Tc), self.SEARCH[search], threads, max_iterations,
                            ctypes.c_double(trans_eps), ctypes.c_double(step_size), ctypes.byref(res))
        return dict(T=_Tback(res.T), converged=bool(res.converged), nr_iterations=res.nr_iterations, n_derivative_evals=res.n_derivative_evals,
                    n_hessian_evals=res.n_hessian_evals, trans_probability=res.trans_probability, p_final=np.array(res.p_final))


def euler_xyz_f32(R):
    R = np.ascontiguousarray(R, dtype=np.float32).reshape(9)
    out = np.empty(3, np.float32)
    lib().orc_euler_xyz_f32(_p(R), _p(out))
    return out


def gicp_covariances(pts, k=20, threads=8, want_idx=False):
    a, n, st = _pts(pts)
    covs = np.empty((n, 3, 3))
    idx = np.empty((n, k), np.int64) if want_idx else None
    lib().orc_gicp_covariances(_p(a), ctypes.c_size_t(n), ctypes.c_size_t(st), k, threads, _p(covs), _p(idx))
    return (covs, idx) if want_idx else covs


class Vgicp:
    def __init__(self, dst, resolution=1.0, k=20, threads=8, target_covs=None):
        self.d, nm, ds = _pts(dst)
        tc = np.ascontiguousarray(target_covs, dtype=np.float64) if target_covs is not None else None
        self.h = ctypes.c_void_p(lib().orc_vgicp_create(_p(self.d), ctypes.c_size_t(nm), ctypes.c_size_t(ds), ctypes.c_double(resolution), k,
                                                        threads, _p(tc)))

    def __del__(self):
        try:
            lib().orc_vgicp_destroy(self.h)
        except Exception:
            pass

    def voxels(self):
        n = lib().orc_vgicp_num_voxels(self.h)
        coords = np.empty((n, 3), np.int32); npts = np.empty(n, np.int32); mean = np.empty((n, 3)); cov = np.empty((n, 3, 3))
        lib().orc_vgicp_get_voxels(self.h, _p(coords), _p(npts), _p(mean), _p(cov))
        return dict(coords=coords, npts=npts, mean=mean, cov=cov)

    def linearize(self, src, src_covs, T, threads=8):
        s, ns, ss = _pts(src)
        sc = np.ascontiguousarray(src_covs, dtype=np.float64)
        Tc = _T(T)
        H = np.empty(36); b = np.empty(6); n = ctypes.c_int64(0)
        cost = lib().orc_vgicp_linearize(self.h, _p(s), ctypes.c_size_t(ns), ctypes.c_size_t(ss), _p(sc), _p(Tc), threads, _p(H), _p(b), ctypes.byref(n))
        return dict(cost=cost, H=H.reshape(6, 6), b=b, n=n.value)

    def error(self, src, src_covs, T0, Ti, threads=8):
        s, ns, ss = _pts(src)
        sc = np.ascontiguousarray(src_covs, dtype=np.float64)
        t0, ti = _T(T0), _T(Ti)  # keep the buffers alive across the call
        return lib().orc_vgicp_error(self.h, _p(s), ctypes.c_size_t(ns), ctypes.c_size_t(ss), _p(sc), _p(t0), _p(ti), threads)

    def align(self, src, T, src_covs=None, threads=8, optimizer="LM", max_iterations=64, rot_eps=2e-3, trans_eps=5e-4):
        s, ns, ss = _pts(src)
        sc = np.ascontiguousarray(src_covs, dtype=np.float64) if src_covs is not None else None
        res = VgicpResult()
        tg = _T(T)
        lib().orc_vgicp_align(self.h, _p(s), ctypes.c_size_t(ns), ctypes.c_size_t(ss), _p(sc), _p(tg), threads, 0 if optimizer == "LM" else 1,
                              max_iterations, ctypes.c_double(rot_eps), ctypes.c_double(trans_eps), ctypes.byref(res))
        return dict(T=_Tback(res.T), converged=bool(res.converged), nr_iterations=res.nr_iterations, n_linearize=res.n_linearize,
                    n_error_evals=res.n_error_evals)


def fitness(src, dst, T, max_range=float("inf"), threads=8):
    s, ns, ss = _pts(src)
    d, nm, ds = _pts(dst)
    mr = 1.7976931348623157e308 if max_range == float("inf") else max_range
    tt = _T(T)
    return lib().orc_fitness(_p(s), ctypes.c_size_t(ns), ctypes.c_size_t(ss), _p(d), ctypes.c_size_t(nm), ctypes.c_size_t(ds), _p(tt),
                             ctypes.c_double(mr), threads)
