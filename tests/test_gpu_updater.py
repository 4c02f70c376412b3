"""GPU parity of the device-resident caller against the oracle's restated
OfflineMapUpdater::callback_node loop: query voxelisation, fetch_VoI + transform, the path, map reassembly,
save_static_map -- every cloud bit-identical, order included, node after node (the map is the frame-to-frame state)."""
import numpy as np
import pytest

from erasor_b200 import params as P
from erasor_b200 import synth

pytestmark = pytest.mark.gpu


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.fixture(scope="module")
def scene():
    sc = synth.Scene(seed=31, length=50.0, n_nodes=21, n_dynamic=6)
    kw = dict(n_beams=32, n_az=720)
    nodes = list(range(0, 21))
    initial_map = sc.build_map(nodes, voxel=0.2, **kw)
    scans = [sc.scan(k, seed_offset=5, **kw) for k in nodes]
    poses = [sc.pose7(k) for k in nodes]
    return dict(map=initial_map, scans=scans, poses=poses)


def test_voxelize_preserving_labels(oracle_mod, scene):
    from erasor_b200 import capi
    ep, up = P.preset("seq_05"), P.updater_preset("seq_05")
    u = capi.Updater(up, ep, scene["map"][:10])
    rng = np.random.default_rng(3)
    rnd = np.zeros((50000, 4), dtype=np.float32)
    rnd[:, :3] = rng.uniform(-12, 12, size=(50000, 3))
    rnd[:, 3] = rng.integers(0, 260, 50000)
    for cloud, leaf in ((scene["scans"][3], 0.2), (rnd, 0.2), (rnd[:1], 0.2), (rnd[:0], 0.2), (scene["scans"][7], 0.05), (scene["map"], 0.4)):
        got = u.voxelize(cloud, leaf)
        ref = oracle_mod.voxelize(cloud, leaf)
        assert _same(got, ref), f"n={len(cloud)} leaf={leaf}: {got.shape} vs {ref.shape}"
    u.close()


def test_voxelize_large_cloud_and_key_widths(oracle_mod, scene):
    """The fused kernel's radix sort beyond its comfortable sizes: 2.6 M points (segments longer than eight rows, several per
    warp), a 31-bit key space (four 9-bit passes), a 9-bit one (one pass), duplicates across segment boundaries, and the
    "leaf too small" overflow case (keys = cloud index)."""
    from erasor_b200 import capi
    ep, up = P.preset("seq_05"), P.updater_preset("seq_05")
    u = capi.Updater(up, ep, scene["map"][:10])
    rng = np.random.default_rng(11)
    def cloud(n, span, zspan):
        c = np.zeros((n, 4), dtype=np.float32)
        c[:, 0:2] = rng.uniform(-span, span, size=(n, 2))
        c[:, 2] = rng.uniform(-zspan, zspan, size=n)
        c[:, 3] = rng.integers(0, 260, n)
        return c
    big = cloud(2_600_000, 60.0, 3.0)
    big[1_000_000:1_300_000] = big[:300_000]                   # exact duplicates far apart in the cloud: same voxel, cloud-order sums
    cases = ((big, 0.3), (cloud(200_000, 400.0, 19.0), 0.25), (cloud(40_000, 0.7, 0.1), 0.2), (cloud(30_000, 4000.0, 4000.0), 0.01))
    for c, leaf in cases:
        got = u.voxelize(c, leaf)
        ref = oracle_mod.voxelize(c, leaf)
        assert _same(got, ref), f"n={len(c)} leaf={leaf}: {got.shape} vs {ref.shape}"
    u.close()


@pytest.mark.parametrize("name,version,large", [("seq_05", 3, False), ("seq_00", 3, False), ("seq_05", 2, False), ("large_scale_05", 3, True)])
def test_sequence_parity(oracle_mod, scene, name, version, large):
    from erasor_b200 import capi
    ep = P.preset(name).replace(version=version)
    up = P.updater_preset(name)
    up.version = version
    up.removal_interval = 2
    up.is_large_scale = large
    up.submap_size = 25.0 if large else up.submap_size      # small window so that reassign_submap triggers along 50 m
    o = oracle_mod.OracleUpdater(up, ep, scene["map"])
    u = capi.Updater(up, ep, scene["map"])
    n_proc = 0
    for k, (pose, scan) in enumerate(zip(scene["poses"], scene["scans"])):
        a = o.callback_node(k, pose, scan)
        b = u.process_node(k, pose, scan)
        assert a == b
        if not a:
            continue
        n_proc += 1
        assert _same(u.cloud(u.QUERY_VOI), o.cloud(o.QUERY_VOI)[0]), f"node {k}: query_voi"
        assert _same(u.cloud(u.MAP_VOI), o.cloud(o.MAP_VOI)[0]), f"node {k}: map_voi"
        assert _same(u.cloud(u.OUTSKIRTS), o.cloud(o.OUTSKIRTS)[0]), f"node {k}: outskirts"
        assert _same(u.cloud(u.MAP_REJECTED), o.cloud(o.MAP_REJECTED)[0]), f"node {k}: map_rejected"
        assert _same(u.cloud(u.MAP_ARRANGED), o.cloud(o.MAP_ARRANGED)[0]), f"node {k}: map_arranged"
    assert n_proc == 10
    assert len(o.cloud(o.TOTAL_MAP_REJECTED)[0]) > 0, "the sequence must reject something"
    assert _same(u.save_static_map(0.2), o.save_static_map(0.2))
    u.close()


def test_lookahead_gives_the_same_maps(oracle_mod, scene):
    """erasor_updater_prefetch_scan: scans handed in one node ahead (upload + voxelisation on the second stream) -- every cloud
    identical to the oracle's caller loop, also when a look-ahead is never consumed (wrong scan) or two are pending."""
    from erasor_b200 import capi
    ep, up = P.preset("seq_05"), P.updater_preset("seq_05")
    up.removal_interval = 2
    o = oracle_mod.OracleUpdater(up, ep, scene["map"])
    u = capi.Updater(up, ep, scene["map"])
    scans = [np.ascontiguousarray(s, dtype=np.float32) for s in scene["scans"]]
    proc = [k for k in range(len(scans)) if (k + 1) % up.removal_interval == 0]
    nxt = {a: b for a, b in zip(proc, proc[1:])}
    u.prefetch_scan_ptr(scans[proc[0]].ctypes.data, len(scans[proc[0]]), capi.PTR_HOST)
    for k, pose in enumerate(scene["poses"]):
        a = o.callback_node(k, pose, scans[k])
        if k in nxt and k != proc[2]:                      # one node gets no look-ahead: the usual path in between
            u.prefetch_scan_ptr(scans[nxt[k]].ctypes.data, len(scans[nxt[k]]), capi.PTR_HOST)
        if k == proc[4]:                                   # a look-ahead nobody consumes (a scan that is never processed)
            u.prefetch_scan_ptr(scans[0].ctypes.data, len(scans[0]), capi.PTR_HOST)
        b = u.process_node_ptr(k, pose, scans[k].ctypes.data, len(scans[k]), capi.PTR_HOST)
        assert a == b
        if a:
            assert _same(u.cloud(u.QUERY_VOI), o.cloud(o.QUERY_VOI)[0]), f"node {k}: query_voi"
            assert _same(u.cloud(u.MAP_ARRANGED), o.cloud(o.MAP_ARRANGED)[0]), f"node {k}: map_arranged"
    assert _same(u.save_static_map(0.2), o.save_static_map(0.2))
    u.close()
