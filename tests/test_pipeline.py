"""erasor_b200/pipeline.py (mapgen -> updater -> PR/RR) driven with the oracle's voxeliser and caller loop injected: the
sequencing, the yaml plumbing and the file outputs, on a synthetic SemanticKITTI-layout drive, without a GPU."""
import os

import numpy as np
import pytest

from erasor_b200 import evaluate as E
from erasor_b200 import kitti, pipeline, params as P

YAML = """
erasor:
    max_range: 40.0
    num_rings: 10
    num_sectors: 36
    min_h: -1.0
    max_h: 3.0
    th_bin_max_h: 0.05
    scan_ratio_threshold: 0.3
    minimum_num_pts: 5
    rejection_ratio: 0
    gf_dist_thr: 0.15
    gf_iter: 3
    gf_num_lpr: 10
    gf_th_seeds_height: 0.5
    version: 3
MapUpdater:
    data_name: "99"
    query_voxel_size: 0.2
    map_voxel_size: 0.2
    voxelization_interval: 10
    removal_interval: 2
tf:
    lidar2body: [0.0, 0.0, 1.73, 0.0, 0.0, 0.0, 1.0]
"""


class _OracleUpdaterAdapter:
    def __init__(self, oracle_mod, up, ep, m):
        self.o = oracle_mod.OracleUpdater(up, ep, m)

    def process_node(self, seq, odom, cloud):
        return self.o.callback_node(seq, odom, cloud)

    def save_static_map(self, voxel):
        return self.o.save_static_map(voxel)

    def close(self):
        self.o.close()


def _write_drive(root, n_frames=8, seed=1):
    """a straight drive past a wall, with a 'car' (class 252) that is there only in the first frames"""
    rng = np.random.default_rng(seed)
    seq = root / "sequences" / "99"
    (seq / "velodyne").mkdir(parents=True); (seq / "labels").mkdir()
    with open(seq / "poses.txt", "w") as fh:
        for f in range(n_frames):
            T = np.eye(4); T[2, 3] = 2.0 * f                                      # camera z = forward
            fh.write(" ".join(f"{v:.9e}" for v in T[:3, :].reshape(-1)) + "\n")
            x0 = 2.0 * f                                                          # vehicle position along map x
            gx = rng.uniform(-30, 30, 6000); gy = rng.uniform(-30, 30, 6000)
            ground = np.stack([gx, gy, rng.normal(-1.73, 0.02, 6000)], axis=1)
            wy = rng.uniform(-30, 30, 1500); wz = rng.uniform(-1.7, 1.0, 1500)
            wall = np.stack([np.full(1500, 25.0) - x0, wy, wz], axis=1)           # static wall at map x = 25
            pts, lab = [ground, wall], [np.full(6000, 40), np.full(1500, 50)]
            if f < 3:                                                             # the car: map (12..14, 3..5), only early on
                cx = rng.uniform(12, 14, 400) - x0; cy = rng.uniform(3, 5, 400); cz = rng.uniform(-1.6, -0.2, 400)
                pts.append(np.stack([cx, cy, cz], axis=1)); lab.append(np.full(400, 252))
            xyz = np.concatenate(pts).astype(np.float32)
            scan = np.concatenate([xyz, rng.uniform(0, 1, (len(xyz), 1)).astype(np.float32)], axis=1)
            scan.tofile(seq / "velodyne" / f"{f:06d}.bin")
            np.concatenate(lab).astype(np.uint32).tofile(seq / "labels" / f"{f:06d}.label")
    return str(root)


def test_sequence_pipeline_with_oracle_backends(oracle_mod, tmp_path):
    root = _write_drive(tmp_path)
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text(YAML)
    out = tmp_path / "out"
    res = pipeline.run_semantickitti(root, "99", 0, 8, 1, str(cfg), out_dir=str(out),
                                     voxelize=lambda c, leaf: oracle_mod.voxelize(c, leaf),
                                     make_updater=lambda up, ep, m: _OracleUpdaterAdapter(oracle_mod, up, ep, m))
    assert res["nodes"] == 8                                   # what the C++ subscribers receive: the bag's doubled first frame is dropped
    assert res["processed_scans"] == 4                         # removal_interval 2: nodes 2, 4, 6, 8 of the 8 received (frames 1, 3, 5, 7)
    naive, static = res["naive_map"], res["static_map"]
    assert len(static) > 1000 and len(static) <= len(naive)
    n_dyn_before = int(kitti.is_dynamic(naive[:, 3]).sum())
    n_dyn_after = int(kitti.is_dynamic(static[:, 3]).sum())
    assert n_dyn_before > 20 and n_dyn_after < n_dyn_before    # the parked-then-gone car is (at least partly) erased
    q = res["quality"]
    assert 0.0 <= q["PR"] <= 100.0 and 0.0 <= q["RR"] <= 100.0 and q["RR"] > 0.0
    back = E.read_pcd_ascii(str(out / "99_result.pcd"))
    assert back.shape == static.shape and np.allclose(back, static, atol=1e-5)
    assert os.path.exists(out / "99_naive_map.pcd")
    # the same run through run_offline directly gives the same map (no hidden state in the sequencing)
    ep, up = P.load_yaml(str(cfg))
    nodes = list(kitti.iter_nodes(root, "99", 0, 8, 1))
    again = pipeline.run_offline(nodes, naive, up, ep, make_updater=lambda u, e, m: _OracleUpdaterAdapter(oracle_mod, u, e, m))
    assert np.array_equal(again["static_map"].view(np.uint32), static.view(np.uint32))
