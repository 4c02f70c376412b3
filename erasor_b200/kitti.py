"""SemanticKITTI -> the path's input conventions, restated from the reference's data-prep tooling so that a
user with the raw dataset can feed the device-resident OfflineMapUpdater (erasor_b200.capi.Updater, or the C++ driver
examples/offline_map_updater_main.cpp) without ROS, rosbag or Python 2.

What the reference's scripts/semantickitti2bag/kitti2node.py does, and this module reproduces:
  * frames:  [init] + range(init, end, interval)  -- the first frame twice in the BAG, "since the cpp drops the first data"
    (:386-388); the C++ nodes therefore see range(init, end, interval), and so do the ROS-free callers here (iter_nodes)
  * pose of a node = tf_origin . T_w_cam0[i] . CAM2BASE  as translation + quaternion (x, y, z, w)  (:258-277, :296-309)
  * cloud = velodyne xyz with the FULL 32-bit SemanticKITTI label (semantic | instance << 16) cast NUMERICALLY to float32 in
    `intensity` (:322-324); the C++ side decodes it with static_cast<uint32_t>(intensity), & 0xFFFF for the class,
    >> 16 for the instance (src/offline_map_updater/src/erasor_utils.cpp:64-66).  (The reference README speaks of a byte
    reinterpretation; the code casts, and so does this module.)  Labels above 2^24 are not exactly representable in float32:
    the cast rounds them exactly as numpy's astype does in the reference.
  * header.seq = the dataset frame index (:312)

Not bit-pinned: tf.transformations.quaternion_from_matrix is not installed here; `quaternion_from_matrix` below is the standard
trace / largest-diagonal construction, equal to it up to floating-point rounding (and sign: w >= 0 is not enforced there either).
Dataset layout read: <root>/sequences/<seq>/{velodyne/%06d.bin, labels/%06d.label, poses.txt}.
"""
from __future__ import annotations

import os
from typing import Iterator, List, Sequence, Tuple

import numpy as np

# calibration constants of the reference (kitti2node.py:258-265): camera-0 -> vehicle base, and the axis swap into the map frame
CAM2BASE = np.array([[-1.857739385241e-03, -9.999659513510e-01, -8.039975204516e-03, -4.784029760483e-03],
                     [-6.481465826011e-03, 8.051860151134e-03, -9.999466081774e-01, -7.337429464231e-02],
                     [9.999773098287e-01, -1.805528627661e-03, -6.496203536139e-03, -3.339968064433e-01],
                     [0.0, 0.0, 0.0, 1.0]])
TF_ORIGIN = np.array([[0.0, 0.0, 1.0, 0.0],
                      [-1.0, 0.0, 0.0, 0.0],
                      [0.0, -1.0, 0.0, 0.0],
                      [0.0, 0.0, 0.0, 1.0]])
DYNAMIC_CLASSES = (252, 253, 254, 255, 256, 257, 258, 259)      # erasor_utils.cpp:3, include/erasor/erasor.h:150, scripts/analysis.py:6


def frame_range(init_stamp: int, end_stamp: int, interval: int) -> List[int]:
    """kitti2node.py:388 -- the first frame is written twice because the C++ node drops its first message."""
    return [init_stamp] + list(range(init_stamp, end_stamp, interval))


def read_poses(path: str) -> np.ndarray:
    """poses.txt: one 3x4 row-major T_w_cam0 per line -> (n, 4, 4) float64 (pykitti/odometry.py:99-117)."""
    rows = np.loadtxt(path, dtype=np.float64, ndmin=2)
    if rows.shape[1] != 12:
        raise ValueError(f"{path}: expected 12 numbers per line, got {rows.shape[1]}")
    T = np.zeros((len(rows), 4, 4))
    T[:, :3, :] = rows.reshape(-1, 3, 4)
    T[:, 3, 3] = 1.0
    return T


def quaternion_from_matrix(M: np.ndarray) -> np.ndarray:
    """(x, y, z, w) of the rotation part of a homogeneous matrix (largest-pivot construction; unit norm)."""
    R = np.asarray(M, dtype=np.float64)[:3, :3]
    t = np.trace(R)
    if t > 0.0:
        s = np.sqrt(t + 1.0) * 2.0
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2.0
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    return q / np.linalg.norm(q)


def node_pose(T_w_cam0: np.ndarray) -> np.ndarray:
    """odom of the node message: [x, y, z, qx, qy, qz, qw] of tf_origin . T_w_cam0 . CAM2BASE (kitti2node.py:274-309)."""
    tf = TF_ORIGIN @ (np.asarray(T_w_cam0, dtype=np.float64) @ CAM2BASE)
    return np.concatenate([tf[:3, 3], quaternion_from_matrix(tf)])


def read_scan(path: str) -> np.ndarray:
    """velodyne/%06d.bin: float32 (x, y, z, reflectance) (pykitti/utils.py:90-98)."""
    a = np.fromfile(path, dtype=np.float32)
    if a.size % 4:
        raise ValueError(f"{path}: size is not a multiple of 4 floats")
    return a.reshape(-1, 4)


def read_labels(path: str) -> np.ndarray:
    """labels/%06d.label: uint32 per point, semantic | instance << 16 (pykitti/utils.py:102-121)."""
    return np.fromfile(path, dtype=np.uint32)


def node_cloud(scan: np.ndarray, labels: np.ndarray) -> np.ndarray:
    """the node's lidar cloud: xyz + the full label cast to float32 in `intensity` (kitti2node.py:322-324)."""
    if len(scan) != len(labels):
        raise ValueError(f"scan has {len(scan)} points, label file {len(labels)}")
    out = np.empty((len(scan), 4), dtype=np.float32)
    out[:, :3] = scan[:, :3]
    out[:, 3] = labels.astype(np.float32)
    return out


def decode_label(intensity: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """(semantic class, instance id) as the C++ side reads them back (erasor_utils.cpp:64-66)."""
    v = np.asarray(intensity, dtype=np.float32).astype(np.uint32)
    return v & np.uint32(0xFFFF), v >> np.uint32(16)


def is_dynamic(intensity: np.ndarray) -> np.ndarray:
    sem, _ = decode_label(intensity)
    return np.isin(sem, DYNAMIC_CLASSES)


def iter_nodes(dataset_root: str, sequence: str, init_stamp: int, end_stamp: int, interval: int, ros_duplicate_first: bool = False
               ) -> Iterator[Tuple[int, np.ndarray, np.ndarray]]:
    """Yields (header.seq, odom[7], cloud[n, 4]) as the reference's C++ nodes RECEIVE them: range(init, end, interval).
    The bag itself holds the first frame twice (frame_range) only because the subscribers drop their first message
    (kitti2node.py:386-388); feeding that duplicate to a ROS-free caller would shift `stack_count % removal_interval`
    (OfflineMapUpdater.cpp:206) and mapgen's accumulation by one node.  ros_duplicate_first=True reproduces the bag
    contents (for export / bag parity)."""
    seq_dir = os.path.join(dataset_root, "sequences", sequence)
    poses = read_poses(os.path.join(seq_dir, "poses.txt"))
    frames = frame_range(init_stamp, end_stamp, interval)
    for f in (frames if ros_duplicate_first else frames[1:]):
        if f >= len(poses):
            raise IndexError(f"frame {f} beyond poses.txt ({len(poses)} poses)")
        scan = read_scan(os.path.join(seq_dir, "velodyne", f"{f:06d}.bin"))
        labels = read_labels(os.path.join(seq_dir, "labels", f"{f:06d}.label"))
        yield f, node_pose(poses[f]), node_cloud(scan, labels)


def export_env_layout(nodes: Sequence[Tuple[int, np.ndarray, np.ndarray]], out_dir: str) -> None:
    """Writes the file layout examples/offline_map_updater_main.cpp reads (the reference's main_in_your_env.cpp:79-123):
    <out>/poses_lidar2body.csv (header, then idx,time,x,y,z,qx,qy,qz,qw) and <out>/pcds/%06d.pcd, numbered 0.. in node order."""
    from .evaluate import write_pcd_ascii
    os.makedirs(os.path.join(out_dir, "pcds"), exist_ok=True)
    with open(os.path.join(out_dir, "poses_lidar2body.csv"), "w") as f:
        f.write("index,timestamp,x,y,z,qx,qy,qz,qw\n")
        for i, (seq, odom, cloud) in enumerate(nodes):
            f.write(f"{i},{float(seq):.6f}," + ",".join(f"{v:.17g}" for v in odom) + "\n")
            write_pcd_ascii(os.path.join(out_dir, "pcds", f"{i:06d}.pcd"), cloud)
