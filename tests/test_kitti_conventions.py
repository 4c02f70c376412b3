"""SemanticKITTI -> node conventions (erasor_b200/kitti.py): frame list, pose composition, label cast and its
decode, on a synthetic dataset written in the SemanticKITTI layout; and the export into the file layout the C++ driver reads."""
import os

import numpy as np
import pytest

from erasor_b200 import evaluate as E
from erasor_b200 import kitti as K


def _rot(axis, ang):
    c, s = np.cos(ang), np.sin(ang)
    R = np.eye(4)
    i, j = [(1, 2), (2, 0), (0, 1)][axis]
    R[i, i], R[i, j], R[j, i], R[j, j] = c, -s, s, c
    return R


def _quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


@pytest.fixture()
def dataset(tmp_path):
    rng = np.random.default_rng(7)
    seq = tmp_path / "sequences" / "05"
    (seq / "velodyne").mkdir(parents=True)
    (seq / "labels").mkdir()
    poses, scans, labels = [], [], []
    for f in range(12):
        T = _rot(1, 0.05 * f) @ _rot(0, 0.01 * f) @ _rot(2, -0.02 * f)
        T[:3, 3] = [0.3 * f, -0.02 * f, 1.1 * f]
        poses.append(T)
        n = 50 + f
        s = rng.normal(0, 10, (n, 4)).astype(np.float32)
        sem = rng.choice(np.array([40, 44, 48, 50, 70, 252, 253, 259], dtype=np.uint32), n)
        inst = rng.integers(0, 300, n).astype(np.uint32)
        lab = sem | (inst << np.uint32(16))
        s.tofile(seq / "velodyne" / f"{f:06d}.bin")
        lab.tofile(seq / "labels" / f"{f:06d}.label")
        scans.append(s); labels.append(lab)
    with open(seq / "poses.txt", "w") as fh:
        for T in poses:
            fh.write(" ".join(f"{v:.12e}" for v in T[:3, :].reshape(-1)) + "\n")
    return str(tmp_path), poses, scans, labels


def test_frame_range_repeats_the_first_frame():
    assert K.frame_range(2350, 2360, 2) == [2350, 2350, 2352, 2354, 2356, 2358]        # kitti2node.py:388
    assert len(K.frame_range(2350, 2670, 2)) == 161                                     # BASELINE config 2: 161 nodes


def test_quaternion_round_trip():
    rng = np.random.default_rng(0)
    for _ in range(200):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        M = np.eye(4); M[:3, :3] = _quat_to_R(q)
        q2 = K.quaternion_from_matrix(M)
        assert abs(np.linalg.norm(q2) - 1) < 1e-12
        assert min(np.abs(q2 - q).max(), np.abs(q2 + q).max()) < 1e-12
        assert np.allclose(_quat_to_R(q2), M[:3, :3], atol=1e-12)


def test_nodes_follow_the_reference_conventions(dataset):
    root, poses, scans, labels = dataset
    nodes = list(K.iter_nodes(root, "05", 2, 11, 3, ros_duplicate_first=True))          # the bag's contents
    assert [s for s, _, _ in nodes] == [2, 2, 5, 8]
    assert [s for s, _, _ in K.iter_nodes(root, "05", 2, 11, 3)] == [2, 5, 8]           # what the C++ nodes receive (first message dropped)
    for seq, odom, cloud in nodes:
        tf = K.TF_ORIGIN @ poses[seq] @ K.CAM2BASE                                       # kitti2node.py:274-275
        assert np.allclose(odom[:3], tf[:3, 3], atol=1e-9)
        assert np.allclose(_quat_to_R(odom[3:]), tf[:3, :3], atol=1e-6)                  # CAM2BASE is orthonormal to ~1e-7 only
        assert cloud.dtype == np.float32 and cloud.shape == (len(scans[seq]), 4)
        assert np.array_equal(cloud[:, :3], scans[seq][:, :3])
        assert np.array_equal(cloud[:, 3], labels[seq].astype(np.float32))              # numeric cast, not a byte view (:324)
        sem, inst = K.decode_label(cloud[:, 3])                                          # erasor_utils.cpp:64-66
        exact = labels[seq] < (1 << 24)
        assert np.array_equal(sem[exact], labels[seq][exact] & 0xFFFF) and np.array_equal(inst[exact], labels[seq][exact] >> 16)
        assert np.array_equal(K.is_dynamic(cloud[:, 3])[exact], np.isin(labels[seq][exact] & 0xFFFF, K.DYNAMIC_CLASSES))


def test_export_layout_for_the_cpp_driver(dataset, tmp_path):
    root, poses, scans, labels = dataset
    nodes = list(K.iter_nodes(root, "05", 0, 6, 2, ros_duplicate_first=True))
    out = tmp_path / "env"
    K.export_env_layout(nodes, str(out))
    rows = [l.strip().split(",") for l in open(out / "poses_lidar2body.csv")][1:]
    assert len(rows) == len(nodes) == 4
    for i, (row, (seq, odom, cloud)) in enumerate(zip(rows, nodes)):
        assert int(row[0]) == i and len(row) == 9
        assert np.allclose(np.array(row[2:], dtype=np.float64), odom, rtol=0, atol=1e-12)   # columns 2..8 are what the driver reads
        back = E.read_pcd_ascii(str(out / "pcds" / f"{i:06d}.pcd"))
        assert back.shape == cloud.shape and np.allclose(back, cloud, rtol=1e-6, atol=1e-6)
