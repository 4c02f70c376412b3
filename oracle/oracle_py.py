"""ctypes bindings for the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference``
legs may import this module.  Nothing under ``erasor_b200/`` does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_float, c_int, c_int32, c_long, c_size_t, c_uint8, c_uint32, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class OracleParamsC(ctypes.Structure):
    """oracle::Params (oracle/erasor_oracle.hpp)."""
    _fields_ = [
        ("max_range", c_double), ("min_h", c_double), ("max_h", c_double), ("th_bin_max_h", c_double),
        ("scan_ratio_threshold", c_double), ("rejection_ratio", c_double), ("gf_dist_thr", c_double),
        ("gf_th_seeds_height", c_double), ("map_voxel_size", c_double),
        ("num_rings", c_int), ("num_sectors", c_int), ("num_lowest_pts", c_int), ("minimum_num_pts", c_int),
        ("gf_iter", c_int), ("gf_num_lpr", c_int), ("version", c_int),
        ("cov_mode", c_int), ("sort_mode", c_int), ("skip_voxelize", c_int),
    ]


class OracleUpdaterParamsC(ctypes.Structure):
    """oracle::UpdaterParams."""
    _fields_ = [
        ("query_voxel_size", c_double), ("map_voxel_size", c_double), ("removal_interval", c_int),
        ("is_large_scale", ctypes.c_bool), ("submap_size", c_double), ("max_range", c_double),
        ("version", c_int), ("lidar2body", c_double * 7),
    ]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "_build", "liberasor_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("erasor_oracle.cpp", "oracle_capi.cpp", "erasor_oracle.hpp")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


_lib = None


def lib(opt: str = "O2"):
    global _lib
    if opt != "O2":
        build()
        return _bind(ctypes.CDLL(os.path.join(_HERE, "_build", "liberasor_oracle_O0.so")))
    if _lib is None:
        _lib = _bind(ctypes.CDLL(build()))
    return _lib


def _bind(L):
    fp = POINTER(c_float)
    L.oracle_create.restype = c_void_p
    L.oracle_create.argtypes = [POINTER(OracleParamsC)]
    L.oracle_destroy.argtypes = [c_void_p]
    L.oracle_run.restype = c_double
    L.oracle_run.argtypes = [c_void_p, fp, c_size_t, fp, c_size_t, c_int]
    L.oracle_get_bin_of_point.argtypes = [c_void_p, c_int, POINTER(c_int32)]
    L.oracle_get_bins.argtypes = [c_void_p, c_int, POINTER(c_double), POINTER(c_double), POINTER(c_uint32), POINTER(c_uint8)]
    L.oracle_get_status.argtypes = [c_void_p, POINTER(c_double), POINTER(c_double)]
    L.oracle_get_negzero_fenced.restype = c_long
    L.oracle_get_negzero_fenced.argtypes = [c_void_p]
    L.oracle_num_planes.restype = c_int
    L.oracle_num_planes.argtypes = [c_void_p]
    L.oracle_get_plane.argtypes = [c_void_p, c_int, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32), POINTER(c_double),
                                   POINTER(c_int32), POINTER(c_double), POINTER(c_int32)]
    L.oracle_cloud_size.restype = c_size_t
    L.oracle_cloud_size.argtypes = [c_void_p, c_int]
    L.oracle_get_cloud.restype = c_size_t
    L.oracle_get_cloud.argtypes = [c_void_p, c_int, fp, POINTER(c_uint32), c_size_t]
    L.oracle_mean_cov.restype = ctypes.c_uint
    L.oracle_mean_cov.argtypes = [fp, c_size_t, c_int, fp, fp]
    L.oracle_jacobi_svd.argtypes = [fp, fp, fp]
    L.oracle_voxelize.restype = c_size_t
    L.oracle_voxelize.argtypes = [fp, c_size_t, c_double, fp, c_size_t]
    L.oracle_mapgen_create.restype = c_void_p
    L.oracle_mapgen_create.argtypes = [ctypes.c_float, c_int]
    L.oracle_mapgen_destroy.argtypes = [c_void_p]
    L.oracle_mapgen_accum.argtypes = [c_void_p, POINTER(c_double), fp, c_size_t]
    L.oracle_mapgen_get.restype = c_size_t
    L.oracle_mapgen_get.argtypes = [c_void_p, c_int, fp, c_size_t]
    L.oracle_transform.argtypes = [fp, c_size_t, fp, fp]
    L.oracle_pose_to_matrix.argtypes = [POINTER(c_double), fp]
    L.oracle_invert4.argtypes = [fp, fp]
    L.oracle_fetch_voi.restype = c_size_t
    L.oracle_fetch_voi.argtypes = [fp, c_size_t, POINTER(c_double), c_double, fp, POINTER(c_uint32), c_size_t]
    L.oracle_extract_ground.argtypes = [c_void_p, fp, c_size_t, POINTER(c_uint8)]
    L.oracle_updater_create.restype = c_void_p
    L.oracle_updater_create.argtypes = [POINTER(OracleUpdaterParamsC), POINTER(OracleParamsC), fp, c_size_t]
    L.oracle_updater_destroy.argtypes = [c_void_p]
    L.oracle_updater_callback_node.restype = c_int
    L.oracle_updater_callback_node.argtypes = [c_void_p, c_int, POINTER(c_double), fp, c_size_t]
    L.oracle_updater_last_erasor_seconds.restype = c_double
    L.oracle_updater_last_erasor_seconds.argtypes = [c_void_p]
    L.oracle_updater_last_voi_seconds.restype = c_double
    L.oracle_updater_last_voi_seconds.argtypes = [c_void_p]
    L.oracle_updater_cloud_size.restype = c_size_t
    L.oracle_updater_cloud_size.argtypes = [c_void_p, c_int]
    L.oracle_updater_get_cloud.restype = c_size_t
    L.oracle_updater_get_cloud.argtypes = [c_void_p, c_int, fp, POINTER(c_uint32), c_size_t]
    L.oracle_updater_save_static_map.restype = c_size_t
    L.oracle_updater_save_static_map.argtypes = [c_void_p, c_float, fp, c_size_t]
    return L


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(POINTER(c_float))


def _as_cloud(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.size == 0:
        return a.reshape(0, 4)
    assert a.ndim == 2 and a.shape[1] == 4
    return a


def params_to_c(p) -> OracleParamsC:
    c = OracleParamsC()
    for name, _ in OracleParamsC._fields_:
        setattr(c, name, getattr(p, name))
    return c


class Oracle:
    """One reference-restated ERASOR instance.  ``p`` is an ``erasor_b200.params.ErasorParams``."""

    ARRANGED, COMPLEMENT, MAP_REJECTED, CURR_REJECTED, GROUND_VIZ = range(5)

    def __init__(self, p, opt: str = "O2"):
        self.L = lib(opt)
        self.p = p
        self._pc = params_to_c(p)
        self.h = self.L.oracle_create(ctypes.byref(self._pc))
        self.n_map = 0
        self.n_query = 0

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, map_voi, query_voi, frame: int = 0) -> float:
        m, q = _as_cloud(map_voi), _as_cloud(query_voi)
        self.n_map, self.n_query = len(m), len(q)
        return self.L.oracle_run(self.h, _fptr(m), len(m), _fptr(q), len(q), frame)

    def bin_of_point(self, which: int) -> np.ndarray:
        n = self.n_map if which == 0 else self.n_query
        out = np.empty(n, dtype=np.int32)
        if n:
            self.L.oracle_get_bin_of_point(self.h, which, out.ctypes.data_as(POINTER(c_int32)))
        return out

    def bins(self, which: int):
        B = self.p.num_rings * self.p.num_sectors
        mn, mx = np.empty(B), np.empty(B)
        cnt, occ = np.empty(B, dtype=np.uint32), np.empty(B, dtype=np.uint8)
        self.L.oracle_get_bins(self.h, which, mn.ctypes.data_as(POINTER(c_double)), mx.ctypes.data_as(POINTER(c_double)),
                               cnt.ctypes.data_as(POINTER(c_uint32)), occ.ctypes.data_as(POINTER(c_uint8)))
        return mn, mx, cnt, occ

    def status(self):
        B = self.p.num_rings * self.p.num_sectors
        st, st1 = np.zeros(B), np.zeros(B)
        self.L.oracle_get_status(self.h, st.ctypes.data_as(POINTER(c_double)), st1.ctypes.data_as(POINTER(c_double)))
        return st, st1

    def negzero_fenced(self) -> int:
        return int(self.L.oracle_get_negzero_fenced(self.h))

    def planes(self):
        out = []
        it = self.p.gf_iter
        for i in range(self.L.oracle_num_planes(self.h)):
            b, npnt, ns, ne = c_int32(), c_int32(), c_int32(), c_int32()
            lpr = c_double()
            nd = np.zeros((it, 4))
            ng = np.zeros(it, dtype=np.int32)
            self.L.oracle_get_plane(self.h, i, ctypes.byref(b), ctypes.byref(npnt), ctypes.byref(ns), ctypes.byref(lpr),
                                    ctypes.byref(ne), nd.ctypes.data_as(POINTER(c_double)), ng.ctypes.data_as(POINTER(c_int32)))
            out.append(dict(bin=b.value, n_points=npnt.value, n_seeds=ns.value, lpr=lpr.value, n_empty=ne.value,
                            normal_d=nd, n_ground=ng))
        return out

    def cloud(self, which: int):
        n = self.L.oracle_cloud_size(self.h, which)
        xyzi = np.empty((n, 4), dtype=np.float32)
        src = np.empty(n, dtype=np.uint32)
        if n:
            self.L.oracle_get_cloud(self.h, which, _fptr(xyzi), src.ctypes.data_as(POINTER(c_uint32)), n)
        return xyzi, src

    def extract_ground(self, cloud) -> np.ndarray:
        c = _as_cloud(cloud)
        g = np.zeros(len(c), dtype=np.uint8)
        self.L.oracle_extract_ground(self.h, _fptr(c), len(c), g.ctypes.data_as(POINTER(c_uint8)))
        return g


def mean_cov(cloud, mode: int = 0):
    c = _as_cloud(cloud)
    cov, mean = np.zeros(9, dtype=np.float32), np.zeros(4, dtype=np.float32)
    n = lib().oracle_mean_cov(_fptr(c), len(c), mode, _fptr(cov), _fptr(mean))
    return n, cov.reshape(3, 3), mean


def jacobi_svd(A):
    A = np.ascontiguousarray(A, dtype=np.float32).reshape(9)
    U, sv = np.zeros(9, dtype=np.float32), np.zeros(3, dtype=np.float32)
    lib().oracle_jacobi_svd(_fptr(A), _fptr(U), _fptr(sv))
    return U.reshape(3, 3), sv


def voxelize(cloud, leaf: float) -> np.ndarray:
    c = _as_cloud(cloud)
    out = np.empty((max(len(c), 1), 4), dtype=np.float32)
    n = lib().oracle_voxelize(_fptr(c), len(c), leaf, _fptr(out), len(out))
    return out[:n].copy()


def transform(cloud, T) -> np.ndarray:
    c = _as_cloud(cloud)
    T = np.ascontiguousarray(T, dtype=np.float32).reshape(16)
    out = np.empty_like(c)
    if len(c):
        lib().oracle_transform(_fptr(c), len(c), _fptr(T), _fptr(out))
    return out


def pose_to_matrix(pose7) -> np.ndarray:
    p = np.ascontiguousarray(pose7, dtype=np.float64)
    T = np.zeros(16, dtype=np.float32)
    lib().oracle_pose_to_matrix(p.ctypes.data_as(POINTER(c_double)), _fptr(T))
    return T.reshape(4, 4)


def invert4(T) -> np.ndarray:
    T = np.ascontiguousarray(T, dtype=np.float32).reshape(16)
    out = np.zeros(16, dtype=np.float32)
    lib().oracle_invert4(_fptr(T), _fptr(out))
    return out.reshape(4, 4)


def fetch_voi(map_cloud, pose7, max_range: float):
    """OfflineMapUpdater::fetch_VoI restated as a free function: (voi in the body frame, index of every VoI point in the map)."""
    m = _as_cloud(map_cloud)
    p = np.ascontiguousarray(pose7, dtype=np.float64)
    voi = np.empty((max(len(m), 1), 4), dtype=np.float32)
    idx = np.empty(max(len(m), 1), dtype=np.uint32)
    n = lib().oracle_fetch_voi(_fptr(m), len(m), p.ctypes.data_as(POINTER(c_double)), float(max_range), _fptr(voi),
                               idx.ctypes.data_as(POINTER(c_uint32)), len(m))
    return voi[:n].copy(), idx[:n].copy()


class OracleUpdater:
    """The restated caller loop (OfflineMapUpdater::callback_node)."""
    MAP_ARRANGED, MAP_VOI, QUERY_VOI, STATIC_EST, EGO_COMPLEMENT, MAP_REJECTED, TOTAL_MAP_REJECTED, OUTSKIRTS, SUBMAP_COMPLEMENT = range(9)

    def __init__(self, up, ep, initial_map):
        self.L = lib()
        self.ep = ep
        upc = OracleUpdaterParamsC()
        upc.query_voxel_size = up.query_voxel_size
        upc.map_voxel_size = up.map_voxel_size
        upc.removal_interval = up.removal_interval
        upc.is_large_scale = up.is_large_scale
        upc.submap_size = up.submap_size
        upc.max_range = up.max_range
        upc.version = up.version
        for i in range(7):
            upc.lidar2body[i] = up.lidar2body[i]
        self._upc, self._epc = upc, params_to_c(ep)
        m = _as_cloud(initial_map)
        self.h = self.L.oracle_updater_create(ctypes.byref(upc), ctypes.byref(self._epc), _fptr(m), len(m))

    def close(self):
        if self.h:
            self.L.oracle_updater_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def callback_node(self, seq: int, odom7, lidar) -> bool:
        o = np.ascontiguousarray(odom7, dtype=np.float64)
        l = _as_cloud(lidar)
        return bool(self.L.oracle_updater_callback_node(self.h, seq, o.ctypes.data_as(POINTER(c_double)), _fptr(l), len(l)))

    def erasor_seconds(self) -> float:
        return self.L.oracle_updater_last_erasor_seconds(self.h)

    def voi_seconds(self) -> float:
        return self.L.oracle_updater_last_voi_seconds(self.h)

    def cloud(self, which: int):
        n = self.L.oracle_updater_cloud_size(self.h, which)
        xyzi = np.empty((n, 4), dtype=np.float32)
        src = np.empty(n, dtype=np.uint32)
        if n:
            self.L.oracle_updater_get_cloud(self.h, which, _fptr(xyzi), src.ctypes.data_as(POINTER(c_uint32)), n)
        return xyzi, src

    def save_static_map(self, voxel_size: float) -> np.ndarray:
        n = self.L.oracle_updater_cloud_size(self.h, 0) + self.L.oracle_updater_cloud_size(self.h, 8) + 16
        out = np.empty((n, 4), dtype=np.float32)
        k = self.L.oracle_updater_save_static_map(self.h, voxel_size, _fptr(out), n)
        return out[:k].copy()


class OracleMapGen:
    """The restated naive map builder (src/mapgen/mapgen.hpp:198-309)."""
    CLOUD_MAP, CLOUD_CURR, SAVED_ORIGINAL, SAVED_VOXELIZED = range(4)

    def __init__(self, leafsize: float, is_large_scale: bool = False):
        self.L = lib()
        self.h = c_void_p(self.L.oracle_mapgen_create(leafsize, 1 if is_large_scale else 0))

    def accum(self, odom7, lidar):
        o = np.ascontiguousarray(odom7, dtype=np.float64)
        c = _as_cloud(lidar)
        self.L.oracle_mapgen_accum(self.h, o.ctypes.data_as(POINTER(c_double)), _fptr(c), len(c))

    def cloud(self, which: int) -> np.ndarray:
        n = self.L.oracle_mapgen_get(self.h, which, None, 0)
        out = np.empty((max(n, 1), 4), dtype=np.float32)
        self.L.oracle_mapgen_get(self.h, which, _fptr(out), n)
        return out[:n].copy()

    def close(self):
        if self.h:
            self.L.oracle_mapgen_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
