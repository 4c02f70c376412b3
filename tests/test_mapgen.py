"""The naive map builder (erasor_b200/mapgen.py, reference src/mapgen/mapgen.hpp:198-309) against the oracle's C++ restatement of
the same code: two implementations of the spec (numpy float32 transforms vs C++), one voxeliser (the oracle's, injected -- on a
GPU box the product uses the device kernel behind erasor_updater_voxelize).  Bit-identical clouds, in order."""
import numpy as np
import pytest

from erasor_b200 import kitti, mapgen, synth


def _nodes(n_nodes=9, seed=4):
    """a short drive: labelled scans in the lidar frame + node poses (x, y, z, qx, qy, qz, qw)"""
    sc = synth.Scene(seed=seed, length=40.0, n_nodes=n_nodes, n_dynamic=3)
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n_nodes):
        scan = sc.scan(k, n_beams=24, n_az=360) if hasattr(sc, "scan") else None
        if scan is None:
            n = 4000
            r = rng.uniform(0.5, 45.0, n); th = rng.uniform(-np.pi, np.pi, n)
            scan = np.stack([r * np.cos(th), r * np.sin(th), rng.normal(-1.7, 0.05, n), rng.choice([40.0, 48.0, 252.0], n)], axis=1)
        yaw = 0.03 * k
        q = np.array([0.0, 0.0, np.sin(yaw / 2), np.cos(yaw / 2)])
        q = q + rng.normal(0, 1e-3, 4); q /= np.linalg.norm(q)             # a general (not axis-aligned) unit quaternion
        odom = np.concatenate([[4.0 * k + rng.normal(0, 0.01), 0.1 * k, 0.02 * k], q])
        out.append((k, odom, np.ascontiguousarray(scan, dtype=np.float32)))
    return out


@pytest.mark.parametrize("large_scale", [False, True])
def test_mapgen_matches_the_oracle(oracle_mod, large_scale, monkeypatch):
    if large_scale:
        monkeypatch.setattr(mapgen, "SUBMAP_EVERY", 500)      # the reference's period; with 9 nodes only the first submap fires
    nodes = _nodes()
    gen = mapgen.NaiveMapGenerator(lambda c, leaf: oracle_mod.voxelize(c, leaf), leafsize=0.2, is_large_scale=large_scale)
    o = oracle_mod.OracleMapGen(0.2, large_scale)
    for seq, odom, cloud in nodes:
        gen.accum_point_cloud(odom, cloud)
        o.accum(odom, cloud)
        oc, om = o.cloud(o.CLOUD_CURR), o.cloud(o.CLOUD_MAP)
        assert gen.cloud_curr.shape == oc.shape and np.array_equal(gen.cloud_curr.view(np.uint32), oc.view(np.uint32)), f"node {seq}: cloud_curr"
        assert gen.cloud_map.shape == om.shape and np.array_equal(gen.cloud_map.view(np.uint32), om.view(np.uint32)), f"node {seq}: cloud_map"
    orig, vox = gen.save_naive_map()
    oo, ov = o.cloud(o.SAVED_ORIGINAL), o.cloud(o.SAVED_VOXELIZED)
    assert orig.shape == oo.shape and np.array_equal(orig.view(np.uint32), oo.view(np.uint32))
    assert vox.shape == ov.shape and np.array_equal(vox.view(np.uint32), ov.view(np.uint32))
    assert len(vox) > 1000 and len(vox) < len(orig)
    if large_scale:
        assert len(gen.cloud_maps) == 1          # cnt_voxel 0 fires at the second node (mapgen.hpp:248)
    o.close()


def test_vehicle_footprint_and_lift(oracle_mod):
    """points inside CAR_BODY_SIZE are dropped (strict <, float threshold vs double distance); the rest are lifted by 1.73 m"""
    r = np.float32(2.7)
    pts = np.array([[r, 0, 0, 1], [np.nextafter(r, np.float32(0)), 0, 0, 2], [np.nextafter(r, np.float32(9)), 0, 0, 3],
                    [1.0, 1.0, 0.5, 4], [30.0, -4.0, -1.2, 5]], dtype=np.float32)
    ident = np.array([0, 0, 0, 0, 0, 0, 1.0])
    seen = {}
    gen = mapgen.NaiveMapGenerator(lambda c, leaf: (seen.setdefault("in", c.copy()), c)[1], leafsize=0.2)
    gen.accum_point_cloud(ident, pts)
    kept = seen["in"]
    d2 = pts[:, 0].astype(np.float64) ** 2 + pts[:, 1].astype(np.float64) ** 2
    expect = pts[~(d2 < np.float64(np.float32(2.7 ** 2)))]
    assert np.array_equal(kept[:, 3], expect[:, 3])
    assert np.array_equal(kept[:, 2], (expect[:, 2] + np.float32(1.73)).astype(np.float32))
    o = oracle_mod.OracleMapGen(0.2, False)
    o.accum(ident, pts)
    assert len(o.cloud(o.CLOUD_CURR)) == len(oracle_mod.voxelize(kept, 0.2))
    o.close()


def test_pose_and_transform_match_the_oracle(oracle_mod):
    rng = np.random.default_rng(2)
    for _ in range(50):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        odom = np.concatenate([rng.normal(0, 100, 3), q])
        T = mapgen.pose_to_matrix(odom)
        assert np.array_equal(T.view(np.uint32), oracle_mod.pose_to_matrix(odom).view(np.uint32))
        c = rng.normal(0, 30, (257, 4)).astype(np.float32)
        assert np.array_equal(mapgen.transform_point_cloud(c, T).view(np.uint32), oracle_mod.transform(c, T).view(np.uint32))


def test_kitti_nodes_feed_mapgen(oracle_mod, tmp_path):
    """end of the producer chain: SemanticKITTI files -> kitti.iter_nodes -> build_map (voxeliser injected)"""
    rng = np.random.default_rng(9)
    seq = tmp_path / "sequences" / "00"
    (seq / "velodyne").mkdir(parents=True); (seq / "labels").mkdir()
    with open(seq / "poses.txt", "w") as fh:
        for f in range(6):
            T = np.eye(4); T[2, 3] = 1.5 * f                       # camera z forward
            fh.write(" ".join(f"{v:.9e}" for v in T[:3, :].reshape(-1)) + "\n")
            n = 3000
            r = rng.uniform(0.5, 40.0, n); th = rng.uniform(-np.pi, np.pi, n)
            s = np.stack([r * np.cos(th), r * np.sin(th), rng.normal(-1.7, 0.05, n), rng.uniform(0, 1, n)], axis=1).astype(np.float32)
            s.tofile(seq / "velodyne" / f"{f:06d}.bin")
            rng.choice(np.array([40, 48, 252], dtype=np.uint32), n).tofile(seq / "labels" / f"{f:06d}.label")
    nodes = list(kitti.iter_nodes(str(tmp_path), "00", 0, 6, 2))
    orig, vox = mapgen.build_map(nodes, leafsize=0.2, voxelize=lambda c, leaf: oracle_mod.voxelize(c, leaf))
    assert len(vox) > 500 and set(np.unique(kitti.decode_label(vox[:, 3])[0])) <= {40, 48, 252}
    # the drive advances along +x of the map frame (tf_origin maps camera z to map x)
    assert vox[:, 0].max() - vox[:, 0].min() > 80.0 - 1.0 + 4.5 - 1.0
