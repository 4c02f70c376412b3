"""Seeded synthetic stand-in for the KITTI / SemanticKITTI inputs of the path.

No KITTI data, rosbags or PCDs exist in the build image or on the GPU box (SURVEY.md "ground
facts"), so every BASELINE.json config runs on a synthetic twin produced here:

* a street scene in the world ("origin") frame: undulating ground, building walls, poles,
  parked cars (static, label 10) and moving objects (labels 252 car / 254 person) whose pose
  depends on the node index -- the things ERASOR exists to erase;
* an HDL-64-like ray caster (64 beams, +2 .. -24.8 deg, configurable azimuth steps) that
  produces one ``erasor/node`` worth of data per node: the pose body->origin (reference
  ``msg/node.msg``: ``odom``) and the scan in the LIDAR frame with the SemanticKITTI label
  numerically cast into ``intensity`` (reference ``scripts/semantickitti2bag/kitti2node.py:324``);
* the naively accumulated map the reference's ``mapgen`` step would hand to the path
  (``src/mapgen/mapgen.hpp:198-263``): scans moved to the world frame and voxelised at 0.2 m,
  emitted in ascending voxel-key order like ``pcl::VoxelGrid`` does (x fastest), which is what
  gives real maps their spatial coherence in memory.

Everything is numpy on the host; it is input generation, not part of the timed path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

LABEL_ROAD, LABEL_BUILDING, LABEL_POLE, LABEL_PARKED_CAR, LABEL_VEGETATION = 40.0, 50.0, 80.0, 10.0, 70.0
LABEL_MOVING_CAR, LABEL_MOVING_PERSON = 252.0, 254.0
SENSOR_HEIGHT = 1.73   # /tf/lidar2body z of every KITTI yaml


@dataclass
class Box:
    lo: np.ndarray          # (3,)
    hi: np.ndarray          # (3,)
    label: float
    vel: np.ndarray         # (3,) metres per node; zero for static boxes


class Scene:
    """A seeded street scene plus a straight-ish trajectory."""

    def __init__(self, seed: int = 5, length: float = 160.0, n_nodes: int = 161, n_dynamic: int = 10,
                 n_static_boxes: int = 40):
        rng = np.random.default_rng(seed)
        self.rng_seed = seed
        self.length = length
        self.n_nodes = n_nodes
        self.boxes: List[Box] = []
        zero = np.zeros(3)
        # building walls both sides of the road, with gaps
        x = -60.0
        while x < length + 60.0:
            w = rng.uniform(8.0, 25.0)
            for side in (-1.0, 1.0):
                if rng.uniform() < 0.8:
                    off = rng.uniform(9.0, 14.0)
                    depth = rng.uniform(6.0, 12.0)
                    h = rng.uniform(4.0, 9.0)
                    y0, y1 = sorted((side * off, side * (off + depth)))
                    self.boxes.append(Box(np.array([x, y0, -0.2]), np.array([x + w, y1, h]), LABEL_BUILDING, zero))
            x += w + rng.uniform(1.0, 6.0)
        # poles, parked cars, shrubs
        for _ in range(n_static_boxes):
            kind = rng.integers(0, 3)
            cx = rng.uniform(-40.0, length + 40.0)
            side = rng.choice([-1.0, 1.0])
            if kind == 0:
                cy = side * rng.uniform(6.0, 8.5)
                self.boxes.append(Box(np.array([cx - 0.1, cy - 0.1, -0.1]), np.array([cx + 0.1, cy + 0.1, 5.0]), LABEL_POLE, zero))
            elif kind == 1:
                cy = side * rng.uniform(4.5, 6.0)
                self.boxes.append(Box(np.array([cx - 2.2, cy - 0.9, -0.05]), np.array([cx + 2.2, cy + 0.9, 1.5]), LABEL_PARKED_CAR, zero))
            else:
                cy = side * rng.uniform(7.0, 9.0)
                s = rng.uniform(0.5, 1.5)
                self.boxes.append(Box(np.array([cx - s, cy - s, -0.1]), np.array([cx + s, cy + s, rng.uniform(0.8, 2.5)]), LABEL_VEGETATION, zero))
        # moving objects: cars in the opposite / same lane, a few pedestrians
        for i in range(n_dynamic):
            if i % 4 != 3:
                lane = rng.choice([-2.5, 2.5])
                v = rng.uniform(0.4, 1.2) * (-1.0 if lane > 0 else 1.0)   # m per node
                cx = rng.uniform(-20.0, length + 20.0)
                self.boxes.append(Box(np.array([cx - 2.2, lane - 0.9, 0.0]), np.array([cx + 2.2, lane + 0.9, 1.6]),
                                      LABEL_MOVING_CAR, np.array([v, 0.0, 0.0])))
            else:
                cy = rng.choice([-1.0, 1.0]) * rng.uniform(3.5, 5.5)
                cx = rng.uniform(0.0, length)
                self.boxes.append(Box(np.array([cx - 0.3, cy - 0.3, 0.0]), np.array([cx + 0.3, cy + 0.3, 1.75]),
                                      LABEL_MOVING_PERSON, np.array([rng.uniform(-0.15, 0.15), rng.uniform(-0.05, 0.05), 0.0])))
        # ground undulation parameters
        self.g_amp = rng.uniform(0.02, 0.06, size=3)
        self.g_freq = rng.uniform(0.02, 0.08, size=3)
        self.g_phase = rng.uniform(0, 2 * np.pi, size=3)
        # trajectory: along +x with a gentle weave and yaw
        s = np.linspace(0.0, length, n_nodes)
        self.traj_x = s
        self.traj_y = 0.6 * np.sin(s * 0.05 + rng.uniform(0, 6.28))
        self.traj_yaw = 0.03 * np.sin(s * 0.04 + rng.uniform(0, 6.28))

    # -- poses ---------------------------------------------------------------------------
    def pose7(self, k: int) -> np.ndarray:
        """body -> origin pose of node k as [x y z qx qy qz qw] (geometry_msgs/Pose)."""
        yaw = self.traj_yaw[k]
        return np.array([self.traj_x[k], self.traj_y[k], 0.0, 0.0, 0.0, np.sin(yaw / 2), np.cos(yaw / 2)])

    def ground_z(self, x, y):
        return (self.g_amp[0] * np.sin(self.g_freq[0] * x + self.g_phase[0])
                + self.g_amp[1] * np.sin(self.g_freq[1] * y + self.g_phase[1])
                + self.g_amp[2] * np.sin(self.g_freq[2] * (x + y) + self.g_phase[2]))

    # -- lidar ---------------------------------------------------------------------------
    def scan(self, k: int, n_beams: int = 64, n_az: int = 1800, max_range: float = 120.0, noise: float = 0.01,
             seed_offset: int = 0) -> np.ndarray:
        """Ray-cast node k.  Returns float32 [n,4] = x,y,z (LIDAR frame), label-as-intensity."""
        rng = np.random.default_rng(self.rng_seed * 1000003 + k * 7919 + seed_offset)
        elev = np.deg2rad(np.linspace(2.0, -24.8, n_beams))
        az = np.linspace(0, 2 * np.pi, n_az, endpoint=False) + rng.uniform(0, 2 * np.pi / n_az)
        ce, se = np.cos(elev)[:, None], np.sin(elev)[:, None]
        d_l = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :], np.broadcast_to(se, (n_beams, n_az))], axis=-1)
        d_l = d_l.reshape(-1, 3)
        yaw = self.traj_yaw[k]
        c, s = np.cos(yaw), np.sin(yaw)
        Rw = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
        d_w = d_l @ Rw.T
        o_w = np.array([self.traj_x[k], self.traj_y[k], SENSOR_HEIGHT])
        n = d_w.shape[0]
        t_best = np.full(n, np.inf)
        lab = np.zeros(n, dtype=np.float32)
        # ground: plane z = 0, undulation added to the hit afterwards
        dz = d_w[:, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            tg = np.where(dz < -1e-6, -o_w[2] / dz, np.inf)
        t_best = tg.copy()
        lab[:] = LABEL_ROAD
        # boxes: slab test, restricted to the azimuth columns each box can subtend
        inv = 1.0 / np.where(np.abs(d_w) < 1e-12, 1e-12, d_w)
        inv = inv.reshape(n_beams, n_az, 3)
        t_best = t_best.reshape(n_beams, n_az)
        lab = lab.reshape(n_beams, n_az)
        az_w = az + yaw                                   # world azimuth of column k (increasing)
        daz = 2 * np.pi / n_az
        for b in self.boxes:
            shift = b.vel * k
            lo, hi = b.lo + shift, b.hi + shift
            ctr = 0.5 * (lo + hi)
            if np.hypot(ctr[0] - o_w[0], ctr[1] - o_w[1]) > max_range + 30.0:
                continue
            if lo[0] <= o_w[0] <= hi[0] and lo[1] <= o_w[1] <= hi[1]:
                cols = np.arange(n_az)
            else:
                cx = np.array([lo[0], lo[0], hi[0], hi[0]]) - o_w[0]
                cy = np.array([lo[1], hi[1], lo[1], hi[1]]) - o_w[1]
                ac = np.arctan2(ctr[1] - o_w[1], ctr[0] - o_w[0])
                rel = np.arctan2(cy, cx) - ac
                rel = (rel + np.pi) % (2 * np.pi) - np.pi
                a0, a1 = ac + rel.min() - daz, ac + rel.max() + daz
                k0 = int(np.floor((a0 - az_w[0]) / daz))
                k1 = int(np.ceil((a1 - az_w[0]) / daz))
                cols = np.arange(k0, k1 + 1) % n_az
                if len(cols) >= n_az:
                    cols = np.arange(n_az)
            iv = inv[:, cols, :]
            t0 = (lo - o_w) * iv
            t1 = (hi - o_w) * iv
            tn = np.minimum(t0, t1).max(axis=2)
            tf = np.maximum(t0, t1).min(axis=2)
            tb = t_best[:, cols]
            hit = (tf >= np.maximum(tn, 0.0)) & (tn > 0.5) & (tn < tb)
            t_best[:, cols] = np.where(hit, tn, tb)
            lab[:, cols] = np.where(hit, np.float32(b.label), lab[:, cols])
        t_best = t_best.reshape(-1)
        lab = lab.reshape(-1)
        ok = np.isfinite(t_best) & (t_best < max_range) & (t_best > 2.7)
        t = t_best[ok] + rng.normal(0.0, noise, size=int(ok.sum()))
        p_w = o_w[None, :] + d_w[ok] * t[:, None]
        is_ground = lab[ok] == LABEL_ROAD
        p_w[is_ground, 2] += self.ground_z(p_w[is_ground, 0], p_w[is_ground, 1])
        # world -> lidar frame
        p_l = (p_w - o_w[None, :]) @ Rw
        out = np.empty((p_l.shape[0], 4), dtype=np.float32)
        out[:, :3] = p_l.astype(np.float32)
        out[:, 3] = lab[ok]
        return out

    def scan_world(self, k: int, **kw) -> np.ndarray:
        """Scan k moved to the world frame with the float affine the path itself uses."""
        sc = self.scan(k, **kw)
        T = pose_matrix(self.pose7(k)) @ pose_matrix(np.array([0, 0, SENSOR_HEIGHT, 0, 0, 0, 1.0]))
        out = sc.copy()
        out[:, :3] = (sc[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
        return out

    # -- map -----------------------------------------------------------------------------
    def build_map(self, nodes, voxel: float = 0.2, **scan_kw) -> np.ndarray:
        """Naive accumulation of the given nodes' scans + 0.2 m voxelisation (mapgen.hpp:198-263)."""
        clouds = [self.scan_world(k, **scan_kw) for k in nodes]
        acc = np.concatenate(clouds, axis=0)
        return voxel_downsample(acc, voxel)


def pose_matrix(pose7: np.ndarray) -> np.ndarray:
    x, y, z, qx, qy, qz, qw = [float(v) for v in pose7]
    d = qx * qx + qy * qy + qz * qz + qw * qw
    s = 2.0 / d
    R = np.array([[1 - s * (qy * qy + qz * qz), s * (qx * qy - qw * qz), s * (qx * qz + qw * qy)],
                  [s * (qx * qy + qw * qz), 1 - s * (qx * qx + qz * qz), s * (qy * qz - qw * qx)],
                  [s * (qx * qz - qw * qy), s * (qy * qz + qw * qx), 1 - s * (qx * qx + qy * qy)]])
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = [x, y, z]
    return T


def voxel_downsample(cloud: np.ndarray, voxel: float) -> np.ndarray:
    """Centroid per voxel, label of the first member, output in ascending voxel key (x fastest)."""
    if len(cloud) == 0:
        return cloud.reshape(0, 4).astype(np.float32)
    ijk = np.floor(cloud[:, :3].astype(np.float64) / voxel).astype(np.int64)
    ijk -= ijk.min(axis=0)
    dims = ijk.max(axis=0) + 1
    key = ijk[:, 0] + dims[0] * (ijk[:, 1] + dims[1] * ijk[:, 2])
    order = np.argsort(key, kind="stable")
    key_s = key[order]
    uniq, first, counts = np.unique(key_s, return_index=True, return_counts=True)
    sums = np.add.reduceat(cloud[order, :3].astype(np.float64), first, axis=0)
    out = np.empty((len(uniq), 4), dtype=np.float32)
    out[:, :3] = (sums / counts[:, None]).astype(np.float32)
    # a voxel touched by any dynamic point keeps the dynamic label (trails must stay labelled)
    lab_s = cloud[order, 3]
    dyn = (lab_s >= 252) & (lab_s <= 259)
    any_dyn = np.add.reduceat(dyn.astype(np.int64), first) > 0
    first_lab = lab_s[first]
    dyn_lab = np.maximum.reduceat(np.where(dyn, lab_s, 0.0), first)
    out[:, 3] = np.where(any_dyn, dyn_lab, first_lab)
    return out


def fetch_voi_numpy(map_world: np.ndarray, pose7: np.ndarray, max_range: float) -> Tuple[np.ndarray, np.ndarray]:
    """Host-side stand-in for OfflineMapUpdater::fetch_VoI (OfflineMapUpdater.cpp:381-438) used only to
    prepare frame-independent benchmark inputs: 2-D radius cut, then world -> body.  Returns (voi_body, index)."""
    dx = map_world[:, 0].astype(np.float64) - pose7[0]
    dy = map_world[:, 1].astype(np.float64) - pose7[1]
    idx = np.nonzero(dx * dx + dy * dy < max_range * max_range)[0]
    Tinv = np.linalg.inv(pose_matrix(pose7))
    sel = map_world[idx]
    out = sel.copy()
    out[:, :3] = (sel[:, :3].astype(np.float64) @ Tinv[:3, :3].T + Tinv[:3, 3]).astype(np.float32)
    return out, idx


def query_body(scene: Scene, k: int, voxel: float = 0.2, **scan_kw) -> np.ndarray:
    """Query VoI of node k as the caller prepares it: voxelise the raw scan, lidar -> body
    (OfflineMapUpdater.cpp:237-241)."""
    sc = scene.scan(k, **scan_kw)
    q = voxel_downsample(sc, voxel) if voxel > 0 else sc
    q = q.copy()
    q[:, 2] += np.float32(SENSOR_HEIGHT)
    return q


def make_frames(seed: int, n_frames: int, preset_max_range: float, n_map_nodes: int = 81, n_beams: int = 64, n_az: int = 1800,
                length: float = 160.0, n_dynamic: int = 10, query_voxel: float = 0.2, map_stride: int = 2):
    """Frame-independent workload: one accumulated map, n_frames (map_voi, query_voi) pairs in the body frame.
    Returns dict(map_world, frames=[(map_voi, query_voi, node_index, voi_index)], scene)."""
    n_nodes = max(n_map_nodes, 2)
    scene = Scene(seed=seed, length=length, n_nodes=n_nodes, n_dynamic=n_dynamic)
    nodes = list(range(0, n_nodes, map_stride))
    map_world = scene.build_map(nodes, voxel=0.2, n_beams=n_beams, n_az=n_az)
    frames = []
    picks = np.linspace(0, n_nodes - 1, n_frames).round().astype(int)
    for k in picks:
        voi, idx = fetch_voi_numpy(map_world, scene.pose7(int(k)), preset_max_range)
        q = query_body(scene, int(k), voxel=query_voxel, n_beams=n_beams, n_az=n_az, seed_offset=17)
        frames.append((voi, q, int(k), idx))
    return dict(map_world=map_world, frames=frames, scene=scene)


def adversarial_points(p, n_random: int = 20000, seed: int = 0) -> np.ndarray:
    """Points that sit on or next to every decision boundary of the binning arithmetic
    (SURVEY section 7 step 4): sector and ring edges at float resolution, the z window, r == max_r,
    the axes, signed zeros (excluding the App. B-1 case, which is fenced and tested on its own),
    denormals and huge values."""
    rng = np.random.default_rng(seed)
    R, S = p.num_rings, p.num_sectors
    ring_size = p.max_range / R
    sector_size = 2 * 3.1415926535 / S
    pts = []
    z_in = np.float32(0.5 * (p.min_h + p.max_h))
    # sector boundaries at several radii, +- a few float ulps in angle
    for k in range(S + 1):
        ang = k * sector_size
        for rad in (0.37, 1.0, 7.3, 0.5 * p.max_range, 0.999 * p.max_range):
            x, y = np.float32(rad * np.cos(ang)), np.float32(rad * np.sin(ang))
            for dx in (-2, -1, 0, 1, 2):
                for dy in (-2, -1, 0, 1, 2):
                    xx = _nudge(x, dx)
                    yy = _nudge(y, dy)
                    pts.append((xx, yy, z_in))
    # ring boundaries along a few directions
    for k in range(R + 2):
        rad = k * ring_size
        for ang in (0.0, 0.3, np.pi / 2, 2.0, np.pi, 4.0, 5.5):
            x, y = np.float32(rad * np.cos(ang)), np.float32(rad * np.sin(ang))
            for dx in (-2, -1, 0, 1, 2):
                pts.append((_nudge(x, dx), _nudge(y, -dx), z_in))
                pts.append((_nudge(x, dx), y, z_in))
    # z window edges
    for zz in (p.min_h, p.max_h):
        zf = np.float32(zz)
        for d in (-2, -1, 0, 1, 2):
            pts.append((np.float32(3.0), np.float32(4.0), _nudge(zf, d)))
    # axes, diagonals, zeros, tiny and huge
    tiny, huge = np.float32(1e-42), np.float32(3e38)
    for x, y in ((0.0, 0.0), (1.0, 0.0), (0.0, 1.0), (-1.0, 0.0), (0.0, -1.0), (1.0, 1.0), (-1.0, 1.0), (-1.0, -1.0), (1.0, -1.0),
                 (tiny, tiny), (-tiny, tiny), (tiny, -tiny), (-tiny, -tiny), (huge, 1.0), (1.0, huge), (-0.0, 0.0), (0.0, 0.0),
                 (1.0, -0.0), (p.max_range, 0.0), (0.0, p.max_range), (-p.max_range, 0.0), (0.0, -p.max_range),
                 (1.0, tiny), (1.0, -tiny), (-1.0, tiny), (-1.0, -tiny)):
        pts.append((np.float32(x), np.float32(y), z_in))
    pts.append((np.float32(np.nan), np.float32(1.0), z_in))
    pts.append((np.float32(1.0), np.float32(np.inf), z_in))
    pts.append((np.float32(1.0), np.float32(1.0), np.float32(np.nan)))
    a = np.array(pts, dtype=np.float32)
    # random fill, including points beyond max_range and outside the z window
    r = rng.uniform(0, 1.2 * p.max_range, n_random)
    th = rng.uniform(0, 2 * np.pi, n_random)
    z = rng.uniform(p.min_h - 0.5, p.max_h + 0.5, n_random)
    b = np.stack([r * np.cos(th), r * np.sin(th), z], axis=1).astype(np.float32)
    xyz = np.concatenate([a, b], axis=0)
    out = np.zeros((len(xyz), 4), dtype=np.float32)
    out[:, :3] = xyz
    out[:, 3] = rng.integers(0, 260, len(xyz)).astype(np.float32)
    # drop the fenced App. B-1 inputs (y == -0.0 and x <= -0.0); they get their own test
    bad = (out[:, 1] == 0) & np.signbit(out[:, 1]) & ((out[:, 0] < 0) | ((out[:, 0] == 0) & np.signbit(out[:, 0])))
    return out[~bad]


def _nudge(v: np.float32, k: int) -> np.float32:
    v = np.float32(v)
    for _ in range(abs(k)):
        v = np.nextafter(v, np.float32(np.inf if k > 0 else -np.inf), dtype=np.float32)
    return v
