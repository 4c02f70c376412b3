"""GPU parity of the map-resident frame-independent mode (erasor_process_nodes): the global map stays in HBM, per node only
the pose and the body-frame query cross PCIe; fetch_VoI (OfflineMapUpdater.cpp:381-438) is fused into the polar binning.
Checked against the oracle's fetch_VoI + ERASOR on the same (map, pose, query): per-frame masks over the GLOBAL map indices,
the folded mask, the per-node counters.  Also: asynchronous lanes sharing one map, the accumulate semantics of the fold, the
internal sub-batch split, the bit-packed AND of the exchange step."""
import numpy as np
import pytest

from erasor_b200 import params as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from erasor_b200 import capi as C
    return C


def _nodes(small_workload, picks):
    scene = small_workload["scene"]
    poses = np.stack([scene.pose7(small_workload["frames"][i][2]) for i in picks]).astype(np.float64)
    qs = [small_workload["frames"][i][1] for i in picks]
    return poses, qs


def _oracle_nodes(oracle_mod, p, map_world, poses, qs, voi_range):
    n = len(map_world)
    o = oracle_mod.Oracle(p)
    fk, stats = [], []
    for pose, q in zip(poses, qs):
        voi, idx = oracle_mod.fetch_voi(map_world, pose, voi_range)
        o.run(voi, q)
        _, rej = o.cloud(o.MAP_REJECTED)
        k = np.ones(n, dtype=np.uint8)
        k[idx[rej]] = 0
        fk.append(k)
        stats.append((len(voi), len(o.planes()), len(rej)))
    return np.stack(fk), np.array(stats, dtype=np.int64)


@pytest.mark.parametrize("name,version", [("seq_05", 3), ("seq_05", 2), ("seq_00", 3), ("synthetic_40x360", 3)])
def test_process_nodes_parity(capi, oracle_mod, small_workload, name, version):
    p = P.preset(name).replace(skip_voxelize=1, version=version)
    map_world = small_workload["map_world"]
    poses, qs = _nodes(small_workload, range(6))
    qo = np.cumsum([0] + [len(q) for q in qs]).astype(np.uint64)
    Q = np.concatenate(qs)
    ofk, ost = _oracle_nodes(oracle_mod, p, map_world, poses, qs, p.max_range)
    assert (ofk == 0).any(), "the workload must reject something"
    m = capi.Map(map_world)
    h = capi.Handle(p)
    h.attach_map(m)
    for rep in range(2):                       # second call: descriptors reused (same geometry)
        m.reset_keep()
        keep, fk = h.process_nodes(poses, Q, qo, want_frame_keep=True)
        assert np.array_equal(fk, ofk), f"{name} v{version}: per-frame masks differ in {np.count_nonzero(fk != ofk)} points"
        assert np.array_equal(keep, ofk.min(axis=0))
        assert np.array_equal(m.get_keep(), keep)
    nv, nf, nr = h.node_stats()
    assert np.array_equal(nv, ost[:, 0]) and np.array_equal(nf, ost[:, 1]) and np.array_equal(nr, ost[:, 2])
    # the fold accumulates over batches: two halves onto one mask == the whole
    m.reset_keep()
    h.process_nodes(poses[:2], np.concatenate(qs[:2]), np.cumsum([0] + [len(q) for q in qs[:2]]).astype(np.uint64))
    keep2, _ = h.process_nodes(poses[2:], np.concatenate(qs[2:]), np.cumsum([0] + [len(q) for q in qs[2:]]).astype(np.uint64))
    assert np.array_equal(keep2, ofk.min(axis=0))
    # a different VoI radius than /erasor/max_range
    m.reset_keep()
    ofk2, _ = _oracle_nodes(oracle_mod, p, map_world, poses[:2], qs[:2], 0.8 * p.max_range)
    _, fk2 = h.process_nodes(poses[:2], np.concatenate(qs[:2]), np.cumsum([0] + [len(q) for q in qs[:2]]).astype(np.uint64),
                             voi_max_range=0.8 * p.max_range, want_frame_keep=True)
    assert np.array_equal(fk2, ofk2)
    h.close(); m.close()


def test_packed_xyz_queries(capi, small_workload):
    """ERASOR_PTR_QUERY_XYZ: queries shipped as packed x y z (12 bytes per point) give the masks of the x y z i form -- host
    buffers through the synchronous call, device buffers through the asynchronous one, and a sub-batch split in between."""
    import torch
    p = P.preset("seq_05").replace(skip_voxelize=1)
    map_world = small_workload["map_world"]
    poses, qs = _nodes(small_workload, range(6))
    qo = np.cumsum([0] + [len(q) for q in qs]).astype(np.uint64)
    Q = np.concatenate(qs)
    m = capi.Map(map_world)
    h = capi.Handle(p)
    h.attach_map(m)
    keep_ref, fk_ref = h.process_nodes(poses, Q, qo, want_frame_keep=True)
    assert (fk_ref == 0).any()
    m.reset_keep()
    keep, fk = h.process_nodes(poses, Q, qo, want_frame_keep=True, packed_xyz=True)
    assert np.array_equal(fk, fk_ref) and np.array_equal(keep, keep_ref)
    # device pointers, asynchronous, queries at an address that is only 4-byte aligned
    m.reset_keep()
    dev = torch.device("cuda", 0)
    buf = torch.zeros(3 * len(Q) + 1, dtype=torch.float32, device=dev)
    buf[1:] = torch.from_numpy(np.ascontiguousarray(Q[:, :3])).to(dev).reshape(-1)
    dfk = torch.empty((len(qs), len(map_world)), dtype=torch.uint8, device=dev)
    dkeep = torch.empty(len(map_world), dtype=torch.uint8, device=dev)
    P7 = np.ascontiguousarray(poses, dtype=np.float64)
    h.process_nodes_ptr(P7, buf.data_ptr() + 4, qo, 0.0, dfk.data_ptr(), dkeep.data_ptr(), capi.PTR_DEVICE | capi.PTR_QUERY_XYZ, asynchronous=True)
    h.wait()
    assert np.array_equal(dfk.cpu().numpy(), fk_ref) and np.array_equal(dkeep.cpu().numpy(), keep_ref)
    # cloud mode refuses the flag: its outputs carry the query's intensity
    with pytest.raises(Exception):
        h.L.erasor_set_inputs.restype
        h._ck(h.L.erasor_set_inputs(h.h, Q.ctypes.data, 0, Q.ctypes.data, len(Q), capi.PTR_HOST | capi.PTR_QUERY_XYZ))
    h.close(); m.close()


def test_process_nodes_equals_process_frames(capi, oracle_mod, small_workload):
    """Same nodes through the batch entry point (VoIs cut by the oracle's fetch_VoI, shipped per frame) and through the
    map-resident one: identical rejected sets."""
    p = P.preset("seq_05").replace(skip_voxelize=1)
    map_world = small_workload["map_world"]
    poses, qs = _nodes(small_workload, [0, 2, 3, 5])
    vois, idxs = zip(*[oracle_mod.fetch_voi(map_world, pose, p.max_range) for pose in poses])
    mo = np.cumsum([0] + [len(v) for v in vois]).astype(np.uint64)
    qo = np.cumsum([0] + [len(q) for q in qs]).astype(np.uint64)
    h = capi.Handle(p)
    keep_b = h.process_frames(np.concatenate(vois), mo, np.concatenate(qs), qo)
    m = capi.Map(map_world)
    h.attach_map(m)
    _, fk = h.process_nodes(poses, np.concatenate(qs), qo, want_frame_keep=True)
    for f in range(len(qs)):
        k = np.ones(len(map_world), dtype=np.uint8)
        k[idxs[f][keep_b[int(mo[f]):int(mo[f + 1])] == 0]] = 0
        assert np.array_equal(fk[f], k), f"frame {f}"
    h.close(); m.close()


def test_async_lanes_share_one_map(capi, oracle_mod, small_workload):
    """Three handles fed round-robin with asynchronous submissions against one resident map (the overlapped form bench.py
    times): device buffers and pinned host buffers, replayed CUDA graphs; the shared folded mask equals the oracle's."""
    import torch
    p = P.preset("seq_05").replace(skip_voxelize=1)
    map_world = small_workload["map_world"]
    n = len(map_world)
    groups = [[0, 1], [2, 3], [4, 5]]
    m = capi.Map(map_world)
    lanes = []
    expect = np.ones(n, dtype=np.uint8)
    for g in groups:
        poses, qs = _nodes(small_workload, g)
        ofk, _ = _oracle_nodes(oracle_mod, p, map_world, poses, qs, p.max_range)
        expect &= ofk.min(axis=0)
        Q = np.concatenate(qs)
        qo = np.cumsum([0] + [len(q) for q in qs]).astype(np.uint64)
        h = capi.Handle(p)
        h.attach_map(m)
        lanes.append(dict(h=h, poses=np.ascontiguousarray(poses), qo=qo, dQ=torch.from_numpy(Q).cuda(), hQ=torch.from_numpy(Q).pin_memory(),
                          hK=torch.empty(n, dtype=torch.uint8).pin_memory(), ofk=ofk))
    torch.cuda.synchronize()
    for kind in ("device", "host"):
        for rep in range(3):
            m.reset_keep()
            for L in lanes:
                if kind == "device":
                    L["h"].process_nodes_ptr(L["poses"], L["dQ"].data_ptr(), L["qo"], 0.0, 0, 0, capi.PTR_DEVICE, asynchronous=True)
                else:
                    L["h"].process_nodes_ptr(L["poses"], L["hQ"].data_ptr(), L["qo"], 0.0, 0, L["hK"].data_ptr(), capi.PTR_HOST, asynchronous=True)
            for L in lanes:
                L["h"].wait()
            assert np.array_equal(m.get_keep(), expect), (kind, rep)
    # the last lane to finish saw every verdict only if it ran last; its own copy must at least contain its own frames' zeros
    for L in lanes:
        assert np.all(L["hK"].numpy()[L["ofk"].min(axis=0) == 0] == 0)
        L["h"].close()
    m.close()


def test_fold_accumulates_and_checks_bounds(capi, oracle_mod, small_workload):
    """the fold must accumulate over batches (explicit reset) and ignore indices beyond the mask."""
    import torch
    p = P.preset("seq_05").replace(skip_voxelize=1)
    map_world = small_workload["map_world"]
    n = len(map_world)
    poses, qs = _nodes(small_workload, [1, 4])
    h = capi.Handle(p)
    g = torch.zeros(n, dtype=torch.uint8, device="cuda")
    h.reset_keep_mask(g.data_ptr(), n)
    expect = np.ones(n, dtype=np.uint8)
    for pose, q in zip(poses, qs):                       # two different batches folded onto one mask
        voi, idx = oracle_mod.fetch_voi(map_world, pose, p.max_range)
        o = oracle_mod.Oracle(p); o.run(voi, q)
        _, rej = o.cloud(o.MAP_REJECTED)
        expect[idx[rej]] = 0
        dM, dQ = torch.from_numpy(voi).cuda(), torch.from_numpy(q).cuda()
        dI = torch.from_numpy(idx.view(np.int32)).cuda()
        dK = torch.empty(len(voi), dtype=torch.uint8, device="cuda")
        mo = np.array([0, len(voi)], dtype=np.uint64); qo = np.array([0, len(q)], dtype=np.uint64)
        h.process_frames_ptr(dM.data_ptr(), mo, dQ.data_ptr(), qo, dK.data_ptr(), capi.PTR_DEVICE, (dI.data_ptr(), g.data_ptr(), n))
    assert np.array_equal(g.cpu().numpy(), expect) and (expect == 0).sum() > 0
    # separate fold entry: accumulates too, and drops out-of-range indices
    g2 = torch.zeros(n, dtype=torch.uint8, device="cuda")
    h.reset_keep_mask(g2.data_ptr(), n)
    bad_idx = torch.tensor([0, 5, n + 7, 2**31 + 3], dtype=torch.int64).to(torch.int32).cuda()
    zeros = torch.zeros(4, dtype=torch.uint8, device="cuda")
    h.fold_keep_masks(zeros.data_ptr(), bad_idx.data_ptr(), 4, g2.data_ptr(), n)
    h.fold_keep_masks(zeros.data_ptr(), torch.tensor([9, 11, 9, 0], dtype=torch.int32).cuda().data_ptr(), 4, g2.data_ptr(), n)
    h.synchronize()
    e2 = np.ones(n, dtype=np.uint8); e2[[0, 5, 9, 11]] = 0
    assert np.array_equal(g2.cpu().numpy(), e2)
    h.close()


def test_sub_batch_split_many_small_frames(capi, oracle_mod):
    """More frames than one submission's work queue holds (40 x 360 bins: 2^21 / 14400 = 145 frames): the batch is split
    internally; every frame still equals the oracle and the per-frame counters cover the whole batch."""
    from test_gpu_parity import _crafted_bin_frame
    p = P.preset("synthetic_40x360").replace(skip_voxelize=1)
    rng = np.random.default_rng(77)
    frames = [_crafted_bin_frame(rng, int(rng.integers(30, 120)), "rough") for _ in range(150)]
    mo = np.cumsum([0] + [len(m) for m, _ in frames]).astype(np.uint64)
    qo = np.cumsum([0] + [len(q) for _, q in frames]).astype(np.uint64)
    h = capi.Handle(p)
    keep = h.process_frames(np.concatenate([m for m, _ in frames]), mo, np.concatenate([q for _, q in frames]), qo)
    nf, nr = h.frame_stats()
    o = oracle_mod.Oracle(p)
    tot = 0
    for f, (m, q) in enumerate(frames):
        o.run(m, q)
        _, rej = o.cloud(o.MAP_REJECTED)
        ok = np.ones(len(m), dtype=np.uint8); ok[rej] = 0
        assert np.array_equal(keep[int(mo[f]):int(mo[f + 1])], ok), f"frame {f}"
        assert nf[f] == len(o.planes()) and nr[f] == len(rej), f"frame {f} counters"
        tot += len(rej)
    assert tot > 0
    h.close()


def test_bit_packed_and_of_masks(capi):
    """The local half of the exchange step (pack to bits, AND over ranks, unpack) against numpy, ragged tail included."""
    import torch
    p = P.preset("seq_05")
    h = capi.Handle(p)
    g = torch.Generator().manual_seed(3)
    for n in (1, 31, 32, 33, 100003):
        for r in (1, 2, 8, 40):
            masks = (torch.rand((r, n), generator=g) > 0.1).to(torch.uint8)
            d = masks.cuda()
            out = torch.empty(n, dtype=torch.uint8, device="cuda")
            h.and_keep_masks(d.data_ptr(), r, n, out.data_ptr())
            h.synchronize()
            assert np.array_equal(out.cpu().numpy(), masks.numpy().min(axis=0)), (n, r)
    h.close()


def test_two_gpu_exchange_through_the_c_abi(capi, tmp_path):
    """The collective behind the C ABI (erasor_comm_* + erasor_allgather_and_keep) on 2 GPUs vs numpy.  Needs >= 2 devices
    (gpurun --gpus 2); skipped on a single-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "scripts", "two_gpu_exchange.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29617", script], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "exchange ok" in r.stdout


def test_class_c_bins_turning_up_late(capi, oracle_mod):
    """R-GPF's class C (bins beyond 2560 points) is only launched while such bins are being seen; a submission that meets one
    unannounced gets the class run by erasor_wait.  Batch masks must equal the oracle before, at and after the switch."""
    from test_gpu_parity import _crafted_bin_frame
    p = P.preset("seq_05").replace(skip_voxelize=1)
    rng = np.random.default_rng(5)
    small = _crafted_bin_frame(rng, 200, "rough")
    big = _crafted_bin_frame(rng, 3500, "rough")
    h = capi.Handle(p)
    o = oracle_mod.Oracle(p)
    for tag, (m, q) in (("small", small), ("small again (class C now off)", small), ("big (fix-up)", big), ("big (class C launched)", big), ("small", small)):
        keep = h.process_frames(m, np.array([0, len(m)], dtype=np.uint64), q, np.array([0, len(q)], dtype=np.uint64))
        o.run(m, q)
        _, rej = o.cloud(o.MAP_REJECTED)
        ok = np.ones(len(m), dtype=np.uint8); ok[rej] = 0
        assert len(rej) > 0 and np.array_equal(keep, ok), tag
    h.close()
