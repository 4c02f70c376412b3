"""Parameters of the R-POD -> SRT -> R-GPF path and the reference's config YAML keys.

The key names and defaults are the reference's rosparam names, read once at
construction: ``/erasor/*`` (reference ``include/erasor/erasor.h:47-61``),
``/erasor/version`` and ``/MapUpdater/*`` / ``/large_scale/*`` / ``/tf/lidar2body``
(``src/offline_map_updater/src/OfflineMapUpdater.cpp:63-105``).  A reference
``config/*.yaml`` loads unchanged through :func:`load_yaml`.
"""
from __future__ import annotations

import ctypes
import dataclasses
from dataclasses import dataclass, field
from typing import List


class ErasorParamsC(ctypes.Structure):
    """Mirror of ``erasor_params_t`` in include/erasor_b200.h (and of oracle::Params' head)."""
    _fields_ = [
        ("max_range", ctypes.c_double),
        ("min_h", ctypes.c_double),
        ("max_h", ctypes.c_double),
        ("th_bin_max_h", ctypes.c_double),
        ("scan_ratio_threshold", ctypes.c_double),
        ("rejection_ratio", ctypes.c_double),
        ("gf_dist_thr", ctypes.c_double),
        ("gf_th_seeds_height", ctypes.c_double),
        ("map_voxel_size", ctypes.c_double),
        ("num_rings", ctypes.c_int),
        ("num_sectors", ctypes.c_int),
        ("num_lowest_pts", ctypes.c_int),
        ("minimum_num_pts", ctypes.c_int),
        ("gf_iter", ctypes.c_int),
        ("gf_num_lpr", ctypes.c_int),
        ("version", ctypes.c_int),
        ("cov_mode", ctypes.c_int),
        ("sort_mode", ctypes.c_int),
        ("skip_voxelize", ctypes.c_int),
    ]


@dataclass
class ErasorParams:
    # defaults: erasor.h:47-61 (note max_range's 10.0 there vs 60.0 in OfflineMapUpdater.cpp:78)
    max_range: float = 10.0
    num_rings: int = 20
    num_sectors: int = 60
    max_h: float = 3.0
    min_h: float = 0.0
    th_bin_max_h: float = 0.39
    scan_ratio_threshold: float = 0.22
    num_lowest_pts: int = 5
    minimum_num_pts: int = 4
    rejection_ratio: float = 0.33
    gf_dist_thr: float = 0.05
    gf_iter: int = 3
    gf_num_lpr: int = 10
    gf_th_seeds_height: float = 0.5
    map_voxel_size: float = 0.2
    version: int = 3
    # mode switches (0/1/0 = reference-faithful defaults of this build; see include/erasor_b200.h)
    cov_mode: int = 0
    sort_mode: int = 1
    skip_voxelize: int = 0

    def to_c(self) -> ErasorParamsC:
        c = ErasorParamsC()
        for name, _ in ErasorParamsC._fields_:
            setattr(c, name, getattr(self, name))
        return c

    @property
    def num_bins(self) -> int:
        return self.num_rings * self.num_sectors

    def replace(self, **kw) -> "ErasorParams":
        return dataclasses.replace(self, **kw)


@dataclass
class UpdaterParams:
    # defaults: OfflineMapUpdater.cpp:66-83
    query_voxel_size: float = 0.05
    map_voxel_size: float = 0.05
    voxelization_interval: int = 10
    removal_interval: int = 2
    data_name: str = "00"
    env: str = "outdoor"
    initial_map_path: str = "/"
    save_path: str = "/"
    is_large_scale: bool = False
    submap_size: float = 200.0
    max_range: float = 60.0
    version: int = 3
    verbose: bool = True
    lidar2body: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0])


def from_dict(cfg: dict):
    """Build (ErasorParams, UpdaterParams) from a parsed reference yaml tree."""
    e = cfg.get("erasor", {}) or {}
    m = cfg.get("MapUpdater", {}) or {}
    ls = cfg.get("large_scale", {}) or {}
    ep = ErasorParams()
    for f in dataclasses.fields(ErasorParams):
        if f.name in e:
            setattr(ep, f.name, type(getattr(ep, f.name))(e[f.name]))
    up = UpdaterParams()
    for k in ("query_voxel_size", "map_voxel_size", "voxelization_interval", "removal_interval",
              "data_name", "env", "initial_map_path", "save_path"):
        if k in m:
            setattr(up, k, type(getattr(up, k))(m[k]))
    if "is_large_scale" in ls:
        up.is_large_scale = bool(ls["is_large_scale"])
    if "submap_size" in ls:
        up.submap_size = float(ls["submap_size"])
    # /erasor/max_range is read by BOTH classes with different defaults
    up.max_range = float(e.get("max_range", 60.0))
    up.version = int(e.get("version", 3))
    up.verbose = bool(cfg.get("verbose", True))
    tf = (cfg.get("tf", {}) or {}).get("lidar2body")
    if tf is not None and len(tf) == 7:
        up.lidar2body = [float(v) for v in tf]
    return ep, up


def load_yaml(path: str):
    import yaml
    with open(path, "r") as f:
        cfg = yaml.safe_load(f)
    return from_dict(cfg)


# The shipped KITTI presets (reference config/seq_*.yaml), so that tests and the
# bench do not need /root/reference at run time.
PRESETS = {
    "seq_00": dict(max_range=80.0, num_rings=20, num_sectors=108, min_h=-1.3, max_h=3.0, th_bin_max_h=0.2,
                   scan_ratio_threshold=0.1, minimum_num_pts=6, rejection_ratio=0.0, gf_dist_thr=0.15,
                   gf_iter=3, gf_num_lpr=20, gf_th_seeds_height=0.5, version=3),
    "seq_01": dict(max_range=60.0, num_rings=15, num_sectors=108, min_h=-1.3, max_h=3.0, th_bin_max_h=0.2,
                   scan_ratio_threshold=0.2, minimum_num_pts=6, rejection_ratio=0.0, gf_dist_thr=0.15,
                   gf_iter=3, gf_num_lpr=10, gf_th_seeds_height=0.5, version=3),
    "seq_02": dict(max_range=60.0, num_rings=15, num_sectors=60, min_h=-1.3, max_h=3.2, th_bin_max_h=0.05,
                   scan_ratio_threshold=0.13, minimum_num_pts=20, rejection_ratio=0.0, gf_dist_thr=0.15,
                   gf_iter=3, gf_num_lpr=20, gf_th_seeds_height=0.5, version=3),
    "seq_05": dict(max_range=60.0, num_rings=15, num_sectors=60, min_h=-1.3, max_h=3.2, th_bin_max_h=0.05,
                   scan_ratio_threshold=0.3, minimum_num_pts=10, rejection_ratio=0.0, gf_dist_thr=0.15,
                   gf_iter=3, gf_num_lpr=10, gf_th_seeds_height=0.5, version=3),
    "seq_07": dict(max_range=80.0, num_rings=20, num_sectors=108, min_h=-0.8, max_h=3.1, th_bin_max_h=0.2,
                   scan_ratio_threshold=0.20, num_lowest_pts=1, minimum_num_pts=6, rejection_ratio=0.0,
                   gf_dist_thr=0.125, gf_iter=3, gf_num_lpr=10, gf_th_seeds_height=0.5, version=3),
    "large_scale_05": dict(max_range=80.0, num_rings=20, num_sectors=108, min_h=-1.3, max_h=3.0, th_bin_max_h=0.2,
                           scan_ratio_threshold=0.2, minimum_num_pts=6, rejection_ratio=0.0, gf_dist_thr=0.25,
                           gf_iter=3, gf_num_lpr=20, gf_th_seeds_height=0.5, map_voxel_size=0.2, version=3),
    "synthetic_40x360": dict(max_range=80.0, num_rings=40, num_sectors=360, min_h=-1.3, max_h=3.0, th_bin_max_h=0.2,
                             scan_ratio_threshold=0.2, minimum_num_pts=6, rejection_ratio=0.0, gf_dist_thr=0.15,
                             gf_iter=3, gf_num_lpr=10, gf_th_seeds_height=0.5, version=3),
}
UPDATER_PRESETS = {
    "seq_00": dict(query_voxel_size=0.2, map_voxel_size=0.2, removal_interval=4, data_name="00"),
    "seq_01": dict(query_voxel_size=0.2, map_voxel_size=0.2, removal_interval=1, data_name="01"),
    "seq_02": dict(query_voxel_size=0.2, map_voxel_size=0.2, removal_interval=5, data_name="02"),
    "seq_05": dict(query_voxel_size=0.2, map_voxel_size=0.05, removal_interval=8, data_name="05"),
    "seq_07": dict(query_voxel_size=0.2, map_voxel_size=0.2, removal_interval=5, data_name="07"),
    "large_scale_05": dict(query_voxel_size=0.2, map_voxel_size=0.2, removal_interval=4, data_name="00",
                           is_large_scale=True, submap_size=160.0),
}


def preset(name: str) -> ErasorParams:
    return ErasorParams(**PRESETS[name])


def updater_preset(name: str) -> UpdaterParams:
    up = UpdaterParams(**UPDATER_PRESETS.get(name, {}))
    up.max_range = PRESETS[name]["max_range"]
    up.version = PRESETS[name].get("version", 3)
    up.lidar2body = [0.0, 0.0, 1.73, 0.0, 0.0, 0.0, 1.0]
    return up
