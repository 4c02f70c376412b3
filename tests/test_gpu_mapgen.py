"""SURVEY 8f-4 on the device: the naive map builder with (a) the device voxeliser injected into the numpy per-node step and
(b) the whole per-node step (2.7 m body cut, 1.73 m lift, pose transform, 0.2 m voxelisation) on the device through
erasor_updater_mapgen_node -- both bit-identical to the oracle's C++ restatement of src/mapgen/mapgen.hpp:198-309, cloud by
cloud; then mapgen -> device OfflineMapUpdater -> PR/RR (erasor_b200.pipeline.run_sequence) against the same chain on the oracle."""
import numpy as np
import pytest

from erasor_b200 import mapgen, pipeline
from erasor_b200 import params as P
from test_mapgen import _nodes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("large_scale", [False, True])
@pytest.mark.parametrize("whole_node_on_device", [False, True])
def test_device_mapgen_matches_the_oracle(oracle_mod, large_scale, whole_node_on_device):
    nodes = _nodes()
    vox, prod = mapgen.device_producer()
    gen = mapgen.NaiveMapGenerator(vox, leafsize=0.2, is_large_scale=large_scale, node_producer=prod if whole_node_on_device else None)
    o = oracle_mod.OracleMapGen(0.2, large_scale)
    for seq, odom, cloud in nodes:
        gen.accum_point_cloud(odom, cloud)
        o.accum(odom, cloud)
        oc, om = o.cloud(o.CLOUD_CURR), o.cloud(o.CLOUD_MAP)
        assert gen.cloud_curr.shape == oc.shape and np.array_equal(gen.cloud_curr.view(np.uint32), oc.view(np.uint32)), f"node {seq}: cloud_curr"
        assert gen.cloud_map.shape == om.shape and np.array_equal(gen.cloud_map.view(np.uint32), om.view(np.uint32)), f"node {seq}: cloud_map"
    orig, voxd = gen.save_naive_map()
    oo, ov = o.cloud(o.SAVED_ORIGINAL), o.cloud(o.SAVED_VOXELIZED)
    assert orig.shape == oo.shape and np.array_equal(orig.view(np.uint32), oo.view(np.uint32))
    assert voxd.shape == ov.shape and np.array_equal(voxd.view(np.uint32), ov.view(np.uint32))
    o.close()


def test_vehicle_cut_edges_on_device(oracle_mod):
    """strict < against a FLOAT threshold with a DOUBLE distance (mapgen.hpp:219-224), NaN kept, exactly as the oracle"""
    r = np.float32(2.7)
    pts = np.array([[r, 0, 0, 1], [np.nextafter(r, np.float32(0)), 0, 0, 2], [np.nextafter(r, np.float32(9)), 0, 0, 3],
                    [1.0, 1.0, 0.5, 4], [30.0, -4.0, -1.2, 5], [-1.9091883, 1.9091883, 0.1, 6], [0, 0, 0, 7]], dtype=np.float32)
    ident = np.array([0, 0, 0, 0, 0, 0, 1.0])
    _, prod = mapgen.device_producer()
    got = prod(ident, pts)
    o = oracle_mod.OracleMapGen(0.2, False)
    o.accum(ident, pts)
    exp = o.cloud(o.CLOUD_CURR)
    assert got.shape == exp.shape and np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    o.close()


def test_run_sequence_on_device_equals_oracle_chain(oracle_mod):
    """mapgen -> OfflineMapUpdater -> save_static_map -> PR/RR with everything heavy on the device vs the oracle's objects."""
    from test_pipeline import _OracleUpdaterAdapter
    nodes = _nodes(n_nodes=12, seed=7)
    ep = P.preset("seq_05")
    up = P.updater_preset("seq_05")
    up.removal_interval = 2
    up.map_voxel_size = 0.2
    dev = pipeline.run_sequence(nodes, up, ep)                                   # device voxeliser / producer / updater
    ora = pipeline.run_sequence(nodes, up, ep, voxelize=lambda c, leaf: oracle_mod.voxelize(c, leaf),
                                make_updater=lambda u, e, m: _OracleUpdaterAdapter(oracle_mod, u, e, m))
    assert dev["processed_scans"] == ora["processed_scans"] == 6
    assert np.array_equal(dev["naive_map"].view(np.uint32), ora["naive_map"].view(np.uint32))
    assert dev["static_map"].shape == ora["static_map"].shape and np.array_equal(dev["static_map"].view(np.uint32), ora["static_map"].view(np.uint32))
    assert dev["quality"] == ora["quality"]
