"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): bin indices and SRT flags bit-exact; R-GPF plane normals and retained
point sets within 1e-4.  The CUDA R-GPF reproduces the oracle's float operation order, so these tests
assert exact equality and, separately, the 1e-4 bar.
"""
import numpy as np
import pytest

from erasor_b200 import params as P
from erasor_b200 import synth

pytestmark = pytest.mark.gpu

PRESETS = ["seq_05", "seq_00", "seq_01", "seq_07", "synthetic_40x360"]
NORMAL_TOL = 1e-4   # north_star tolerance for plane normals / d


def _crop(voi, max_range):
    r2 = voi[:, 0].astype(np.float64) ** 2 + voi[:, 1].astype(np.float64) ** 2
    return voi[r2 < (max_range + 5.0) ** 2]


def _frame(small_workload, i, p):
    voi, q, k, idx = small_workload["frames"][i]
    return _crop(voi, p.max_range), q


@pytest.fixture(scope="module")
def capi():
    from erasor_b200 import capi as C
    return C


def _run_both(capi, oracle_mod, p, m, q):
    o = oracle_mod.Oracle(p)
    o.run(m, q)
    h = capi.Handle(p)
    h.set_inputs(m, q)
    return o, h


@pytest.mark.parametrize("name", PRESETS)
def test_bins_exact_adversarial(capi, oracle_mod, name):
    p = P.preset(name).replace(skip_voxelize=1)
    pts = synth.adversarial_points(p, n_random=300000, seed=3)
    q = synth.adversarial_points(p, n_random=20000, seed=4)
    o, h = _run_both(capi, oracle_mod, p, pts, q)
    for which in (0, 1):
        bop, mn, mx, cnt = h.get_bins(which)
        ob = o.bin_of_point(which)
        assert np.array_equal(bop, ob), f"{name} cloud {which}: {np.count_nonzero(bop != ob)} bin ids differ"
        omn, omx, ocnt, occ = o.bins(which)
        assert np.array_equal(cnt, ocnt)
        occ = ocnt > 0
        assert np.array_equal(mn[occ].astype(np.float64), omn[occ]) and np.array_equal(mx[occ].astype(np.float64), omx[occ])
        assert np.all(np.isnan(mn[~occ])) and np.all(np.isnan(mx[~occ]))
    f = h.fence_counts()
    assert f["ambiguous_sector"] == 0 and f["negzero_points"] == 0
    h.close()


@pytest.mark.parametrize("name", PRESETS)
@pytest.mark.parametrize("version", [3, 2])
def test_frame_parity(capi, oracle_mod, small_workload, name, version):
    p = P.preset(name).replace(skip_voxelize=1, version=version)
    for fi in (1, 3):
        m, q = _frame(small_workload, fi, p)
        o, h = _run_both(capi, oracle_mod, p, m, q)
        h.compare(version)
        # bins
        for which in (0, 1):
            bop, mn, mx, cnt = h.get_bins(which)
            assert np.array_equal(bop, o.bin_of_point(which))
            omn, omx, ocnt, _ = o.bins(which)
            assert np.array_equal(cnt, ocnt)
        # status
        st, _ = o.status()
        assert np.array_equal(h.get_status().astype(np.float64), st.astype(np.float32).astype(np.float64)), f"{name} v{version} status"
        # planes
        gp, op = h.get_planes(), o.planes()
        assert [g["bin"] for g in gp] == [x["bin"] for x in op]
        for g, x in zip(gp, op):
            assert g["n_points"] == x["n_points"] and g["n_seeds"] == x["n_seeds"]
            assert g["lpr"] == x["lpr"]
            assert np.allclose(g["normal_d"], x["normal_d"], atol=NORMAL_TOL, rtol=0)
            assert np.array_equal(g["normal_d"], x["normal_d"]), "plane parameters are expected to be bit-identical"
            assert np.array_equal(g["n_ground"], x["n_ground"])
        # retained / rejected sets
        keep, gnd = h.get_static_mask()
        rej_xyzi, rej_src = o.cloud(o.MAP_REJECTED)
        okeep = np.ones(len(m), dtype=np.uint8)
        okeep[rej_src] = 0
        assert np.array_equal(keep, okeep)
        gv_xyzi, gv_src = o.cloud(o.GROUND_VIZ)
        ognd = np.zeros(len(m), dtype=np.uint8)
        ognd[gv_src] = 1
        assert np.array_equal(gnd, ognd)
        # clouds in the reference's order
        arr, cmp_ = h.get_static_estimate()
        oarr, _ = o.cloud(o.ARRANGED)
        ocmp, _ = o.cloud(o.COMPLEMENT)
        assert arr.shape == oarr.shape and np.array_equal(arr.view(np.uint32), oarr.view(np.uint32)), f"{name} v{version} arranged"
        assert cmp_.shape == ocmp.shape and np.array_equal(cmp_.view(np.uint32), ocmp.view(np.uint32))
        mr, cr = h.get_outliers()
        ocr, _ = o.cloud(o.CURR_REJECTED)
        assert mr.shape == rej_xyzi.shape and np.array_equal(mr.view(np.uint32), rej_xyzi.view(np.uint32))
        assert cr.shape == ocr.shape and np.array_equal(cr.view(np.uint32), ocr.view(np.uint32))
        h.close()


def test_batch_masks(capi, oracle_mod, small_workload):
    p = P.preset("seq_05").replace(skip_voxelize=1)
    frames = [_frame(small_workload, i, p) for i in range(6)]
    mo = np.cumsum([0] + [len(m) for m, _ in frames]).astype(np.uint64)
    qo = np.cumsum([0] + [len(q) for _, q in frames]).astype(np.uint64)
    M = np.concatenate([m for m, _ in frames])
    Q = np.concatenate([q for _, q in frames])
    h = capi.Handle(p)
    keep = h.process_frames(M, mo, Q, qo)
    nf, nr = h.frame_stats()
    for f, (m, q) in enumerate(frames):
        o = oracle_mod.Oracle(p)
        o.run(m, q)
        _, rej_src = o.cloud(o.MAP_REJECTED)
        okeep = np.ones(len(m), dtype=np.uint8)
        okeep[rej_src] = 0
        assert np.array_equal(keep[int(mo[f]):int(mo[f + 1])], okeep), f"frame {f}"
        assert nf[f] == len(o.planes()) and nr[f] == len(rej_src)
    h.close()


def test_empty_and_tiny_inputs(capi, oracle_mod):
    p = P.preset("seq_05").replace(skip_voxelize=1)
    z = np.zeros((0, 4), dtype=np.float32)
    one = np.array([[3.0, 4.0, 0.1, 7.0]], dtype=np.float32)
    for m, q in ((z, z), (one, z), (z, one), (one, one)):
        o, h = _run_both(capi, oracle_mod, p, m, q)
        h.compare(3)
        arr, cmp_ = h.get_static_estimate()
        oarr, _ = o.cloud(o.ARRANGED)
        assert arr.shape == oarr.shape
        assert np.array_equal(h.get_status().astype(np.float64), o.status()[0])
        h.close()


def test_negzero_fence(capi, oracle_mod):
    """the reference throws on y == -0.0f, x < 0; both sides fence it to y = +0 and count it."""
    p = P.preset("seq_05").replace(skip_voxelize=1)
    m = np.array([[-5.0, -0.0, 0.2, 1.0], [-0.0, -0.0, 0.2, 1.0], [5.0, -0.0, 0.2, 1.0], [-5.0, 0.0, 0.2, 1.0]], dtype=np.float32)
    o, h = _run_both(capi, oracle_mod, p, m, m[:1])
    bop, *_ = h.get_bins(0)
    assert np.array_equal(bop, o.bin_of_point(0))
    assert h.fence_counts()["negzero_points"] == 3 and o.negzero_fenced() == 3   # two map points + one query point
    h.close()


@pytest.mark.parametrize("name", ["seq_05", "seq_00", "large_scale_05"])
def test_v3_with_in_bin_voxelization(capi, oracle_mod, small_workload, name):
    """Version 3 as shipped: flagged bins = voxelize_preserving_labels(bin_curr + ground) (erasor.cpp:526-528)."""
    p = P.preset(name).replace(skip_voxelize=0, version=3)
    for fi in (1, 4):
        m, q = _frame(small_workload, fi, p)
        o, h = _run_both(capi, oracle_mod, p, m, q)
        h.compare(3)
        assert len(o.planes()) > 0, "the frame must exercise flagged bins"
        arr, cmp_ = h.get_static_estimate()
        oarr, _ = o.cloud(o.ARRANGED)
        assert arr.shape == oarr.shape, f"{arr.shape} vs {oarr.shape}"
        assert np.array_equal(arr.view(np.uint32), oarr.view(np.uint32)), f"{name}: arranged cloud differs"
        ocmp, _ = o.cloud(o.COMPLEMENT)
        assert np.array_equal(cmp_.view(np.uint32), ocmp.view(np.uint32))
        h.close()


def test_fold_keep_masks(capi):
    """The library's fold kernel (multi-GPU exchange helper) against the reference fold of erasor_b200/dist.py."""
    import torch
    from erasor_b200 import dist as D
    p = P.preset("seq_05").replace(skip_voxelize=1)
    h = capi.Handle(p)
    g = torch.Generator().manual_seed(0)
    n_global, n = 100000, 250000
    idx = torch.randint(0, n_global, (n,), generator=g)
    keep = (torch.rand(n, generator=g) > 0.05).to(torch.uint8)
    expect = D.fold_masks(n_global, [idx], [keep]).numpy()
    d_idx = idx.to(torch.int32).cuda()
    d_keep = keep.cuda()
    d_out = torch.zeros(n_global, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    h.reset_keep_mask(d_out.data_ptr(), n_global)          # the fold accumulates: a job starts with an explicit reset
    h.fold_keep_masks(d_keep.data_ptr(), d_idx.data_ptr(), n, d_out.data_ptr(), n_global)
    h.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), expect)
    h.close()


def test_full_size_dense_voi_properties(capi, oracle_mod, small_workload):
    """BASELINE.json config 4 (large_scale_05.yaml geometry, ~50 M-point VoI): too big for the oracle, so the CUDA path is
    checked through size-independent properties -- every point accounted for, the per-bin tables equal an independent
    torch reduction over the returned bin ids, a 300 k-point sample of bin ids equals the oracle's (the bin of a point
    does not depend on its neighbours), and the keep mask only ever clears points of flagged bins."""
    import torch
    p = P.preset("large_scale_05").replace(skip_voxelize=1)
    base, q, _, _ = small_workload["frames"][2]
    base = base[(base[:, 0].astype(np.float64) ** 2 + base[:, 1].astype(np.float64) ** 2) < 82.0 ** 2]
    reps = int(np.ceil(50_000_000 / len(base)))
    g = torch.Generator(device="cuda").manual_seed(5)
    tb = torch.from_numpy(base).cuda()
    M = tb.repeat(reps, 1)
    M[:, :3] += (torch.rand((M.shape[0], 3), device="cuda", generator=g) - 0.5) * torch.tensor([0.19, 0.19, 0.02], device="cuda")
    N = M.shape[0]
    assert N >= 50_000_000
    Q = torch.from_numpy(q).cuda()
    h = capi.Handle(p)
    h.set_inputs_device(M.data_ptr(), N, Q.data_ptr(), len(q))
    bop, mn, mx, cnt = h.get_bins(0)
    binned = bop >= 0
    assert int(cnt.sum()) == int(binned.sum())                                    # N_m == sum |bin| + |complement|
    B = p.num_bins
    tb_ids = torch.from_numpy(bop).cuda().long()
    sel = tb_ids >= 0
    z = M[:, 2]
    t_cnt = torch.bincount(tb_ids[sel], minlength=B)
    assert torch.equal(t_cnt.cpu(), torch.from_numpy(cnt.astype(np.int64)))
    t_mx = torch.full((B,), -float("inf"), device="cuda").scatter_reduce_(0, tb_ids[sel], z[sel], reduce="amax")
    t_mn = torch.full((B,), float("inf"), device="cuda").scatter_reduce_(0, tb_ids[sel], z[sel], reduce="amin")
    occ = cnt > 0
    assert np.array_equal(t_mx.cpu().numpy()[occ], mx[occ]) and np.array_equal(t_mn.cpu().numpy()[occ], mn[occ])
    # sample against the oracle
    pick = torch.randperm(N, device="cuda", generator=g)[:300000]
    sample = M[pick].cpu().numpy()
    o = oracle_mod.Oracle(p)
    o.run(sample, q[:10])
    assert np.array_equal(o.bin_of_point(0), bop[pick.cpu().numpy()])
    # the whole path at full size: mask mode, one frame
    keep = torch.empty(N, dtype=torch.uint8, device="cuda")
    mo = np.array([0, N], dtype=np.uint64)
    qo = np.array([0, len(q)], dtype=np.uint64)
    h.reset_kernel_times(True)
    h.process_frames_ptr(M.data_ptr(), mo, Q.data_ptr(), qo, keep.data_ptr(), capi.PTR_DEVICE)
    k1_ms, k1_n = h.kernel_time_ms(1)
    nf, nr = h.frame_stats()
    rejected = (keep == 0)
    assert int(rejected.sum().item()) == int(nr[0])
    assert nf[0] > 0 and nr[0] > 0
    assert bool((tb_ids[rejected] >= 0).all().item())                              # only binned points can be rejected
    print(f"[full size] N={N} K1 {k1_ms / max(1, k1_n):.3f} ms -> {16 * (N + len(q)) / (k1_ms / max(1, k1_n) * 1e-3) / 1e9:.0f} GB/s, "
          f"flagged bins {int(nf[0])}, rejected {int(nr[0])}")
    h.close()


@pytest.mark.parametrize("overrides", [
    dict(gf_iter=1), dict(gf_iter=0), dict(minimum_num_pts=0), dict(minimum_num_pts=1, num_lowest_pts=0, gf_num_lpr=1),
    dict(num_lowest_pts=1000000), dict(scan_ratio_threshold=0.9), dict(cov_mode=1), dict(gf_dist_thr=0.0), dict(gf_th_seeds_height=-5.0),
])
def test_parameter_edge_cases(capi, oracle_mod, small_workload, overrides):
    """Parameter corners of R-GPF / SRT: single and zero plane-fit iterations, no minimum point count, LPR windows that
    fall off the bin, thresholds that flag almost every bin, the PCL >= 1.11 covariance, seed windows that select nothing
    (the App. B-3 empty-fit fence)."""
    p = P.preset("seq_05").replace(skip_voxelize=1, **overrides)
    m, q = _frame(small_workload, 3, p)
    for version in (3, 2):
        o = oracle_mod.Oracle(p.replace(version=version))
        o.run(m, q)
        h = capi.Handle(p.replace(version=version))
        h.set_inputs(m, q)
        h.compare(version)
        assert np.array_equal(h.get_status(), o.status()[0].astype(np.float32))
        gp, op = h.get_planes(), o.planes()
        assert [g["bin"] for g in gp] == [x["bin"] for x in op]
        for g, x in zip(gp, op):
            assert np.array_equal(g["normal_d"], x["normal_d"]) and np.array_equal(g["n_ground"], x["n_ground"]), overrides
        arr, cmp_ = h.get_static_estimate()
        oarr, _ = o.cloud(o.ARRANGED)
        assert arr.shape == oarr.shape and np.array_equal(arr.view(np.uint32), oarr.view(np.uint32)), (overrides, version)
        mr, _ = h.get_outliers()
        omr, _ = o.cloud(o.MAP_REJECTED)
        assert mr.shape == omr.shape and np.array_equal(mr.view(np.uint32), omr.view(np.uint32))
        if overrides.get("gf_th_seeds_height", 0) < 0:
            assert h.fence_counts()["empty_plane_fits"] > 0 and sum(x["n_empty"] for x in op) > 0
        h.close()


def test_batch_masks_v2_and_ragged(capi, oracle_mod, small_workload):
    """Mask mode with version 2, frames of very different sizes and an empty frame in the middle."""
    p = P.preset("seq_00").replace(skip_voxelize=1, version=2)
    fr = [_frame(small_workload, i, p) for i in (0, 2, 5)]
    z = np.zeros((0, 4), dtype=np.float32)
    frames = [fr[0], (z, z), (fr[1][0][:777], fr[1][1]), fr[2], (fr[2][0], z)]
    mo = np.cumsum([0] + [len(m) for m, _ in frames]).astype(np.uint64)
    qo = np.cumsum([0] + [len(q) for _, q in frames]).astype(np.uint64)
    M = np.concatenate([m for m, _ in frames])
    Q = np.concatenate([q for _, q in frames])
    h = capi.Handle(p)
    keep = h.process_frames(M, mo, Q, qo)
    for f, (m, q) in enumerate(frames):
        o = oracle_mod.Oracle(p)
        o.run(m, q)
        _, rej = o.cloud(o.MAP_REJECTED)
        okeep = np.ones(len(m), dtype=np.uint8)
        okeep[rej] = 0
        assert np.array_equal(keep[int(mo[f]):int(mo[f + 1])], okeep), f"frame {f}"
    h.close()


def _crafted_bin_frame(rng, n_ground, z_kind, sector_deg=15.0, r0=20.0):
    """One map bin that the Scan Ratio Test flags (tall object over ground, low query bin), with a chosen z distribution of
    its ground points; returns (map, query).  Everything sits inside one ring / sector of the seq_05 geometry."""
    th = np.deg2rad(sector_deg) + rng.uniform(-0.02, 0.02, n_ground)
    r = r0 + rng.uniform(0.2, 3.0, n_ground)
    if z_kind == "dup":            # a handful of distinct heights, many exact duplicates
        z = rng.choice(np.array([-1.0, -0.95, -0.9, -0.9000001, -0.0, 0.0], dtype=np.float32), n_ground)
    elif z_kind == "equal":        # every point at the same height
        z = np.full(n_ground, -0.9, dtype=np.float32)
    elif z_kind == "ulp":          # clusters a few ulps wide next to a wide outlier range (coarse q cells)
        z = (np.float32(-0.9) + rng.integers(0, 7, n_ground).astype(np.float32) * np.float32(6e-8)).astype(np.float32)
    else:                          # rough ground
        z = rng.normal(-0.9, 0.04, n_ground).astype(np.float32)
    g = np.stack([r * np.cos(th), r * np.sin(th), z, np.full(n_ground, 40.0)], axis=1).astype(np.float32)
    n_obj = 60
    tho = np.deg2rad(sector_deg) + rng.uniform(-0.01, 0.01, n_obj)
    ro = r0 + rng.uniform(1.0, 2.0, n_obj)
    obj = np.stack([ro * np.cos(tho), ro * np.sin(tho), rng.uniform(-0.7, 2.0, n_obj), np.full(n_obj, 252.0)], axis=1).astype(np.float32)
    m = rng.permutation(np.concatenate([g, obj]))
    nq = 40
    thq = np.deg2rad(sector_deg) + rng.uniform(-0.02, 0.02, nq)
    rq = r0 + rng.uniform(0.2, 3.0, nq)
    q = np.stack([rq * np.cos(thq), rq * np.sin(thq), rng.normal(-0.9, 0.03, nq), np.full(nq, 40.0)], axis=1).astype(np.float32)
    return m, q


@pytest.mark.parametrize("z_kind", ["rough", "dup", "equal", "ulp"])
@pytest.mark.parametrize("n_ground", [60, 190, 450, 600, 1200, 2400, 3500])
def test_rgpf_sort_classes_and_ties(capi, oracle_mod, z_kind, n_ground):
    """R-GPF's z-sort (erasor.cpp:240) across the three size classes of K4 with tie-heavy and ulp-wide height distributions:
    the packed 32-bit network + exact fix-up, its 64-bit fallback and the radix sort must all give std::stable_sort's order
    (plane normals, ground counts and the retained cloud are compared bit for bit)."""
    rng = np.random.default_rng(1000 + n_ground)
    p = P.preset("seq_05").replace(skip_voxelize=1, version=3)
    m, q = _crafted_bin_frame(rng, n_ground, z_kind)
    o, h = _run_both(capi, oracle_mod, p, m, q)
    h.compare(3)
    gp, op = h.get_planes(), o.planes()
    assert len(op) >= 1, "the crafted bin must be flagged"
    assert [g["bin"] for g in gp] == [x["bin"] for x in op]
    for g, x in zip(gp, op):
        assert np.array_equal(g["normal_d"], x["normal_d"]) and np.array_equal(g["n_ground"], x["n_ground"]), (z_kind, n_ground)
    arr, _ = h.get_static_estimate()
    oarr, _ = o.cloud(o.ARRANGED)
    assert arr.shape == oarr.shape and np.array_equal(arr.view(np.uint32), oarr.view(np.uint32))
    h.close()


def test_process_frames_fold_matches_separate_calls(capi, small_workload):
    """erasor_process_frames_fold == erasor_process_frames followed by erasor_fold_keep_masks (device and pinned-host clouds,
    repeated calls so that the cached CUDA graph is replayed)."""
    import torch
    p = P.preset("seq_05").replace(skip_voxelize=1)
    fr = [_frame(small_workload, i, p) for i in (1, 3, 4)]
    mo = np.cumsum([0] + [len(m) for m, _ in fr]).astype(np.uint64)
    qo = np.cumsum([0] + [len(q) for _, q in fr]).astype(np.uint64)
    M = np.concatenate([m for m, _ in fr]); Q = np.concatenate([q for _, q in fr])
    n_global = 50000
    rng = np.random.default_rng(11)
    idx = rng.integers(0, n_global, len(M)).astype(np.uint32)
    h = capi.Handle(p)
    dM, dQ = torch.from_numpy(M).cuda(), torch.from_numpy(Q).cuda()
    dI = torch.from_numpy(idx.view(np.int32)).cuda()
    keep_a = torch.empty(len(M), dtype=torch.uint8, device="cuda"); keep_b = torch.empty_like(keep_a)
    g_a = torch.zeros(n_global, dtype=torch.uint8, device="cuda"); g_b = torch.zeros_like(g_a)
    h.process_frames_ptr(dM.data_ptr(), mo, dQ.data_ptr(), qo, keep_a.data_ptr(), capi.PTR_DEVICE)
    g_a.fill_(1)
    h.fold_keep_masks(keep_a.data_ptr(), dI.data_ptr(), len(M), g_a.data_ptr(), n_global)
    torch.cuda.synchronize()
    for _ in range(3):
        g_b.fill_(1)
        h.process_frames_ptr(dM.data_ptr(), mo, dQ.data_ptr(), qo, keep_b.data_ptr(), capi.PTR_DEVICE, (dI.data_ptr(), g_b.data_ptr(), n_global))
        torch.cuda.synchronize()
        assert torch.equal(keep_a, keep_b) and torch.equal(g_a, g_b)
    assert int((g_a == 0).sum().item()) > 0
    hM, hQ = torch.from_numpy(M).pin_memory(), torch.from_numpy(Q).pin_memory()
    hK = torch.empty(len(M), dtype=torch.uint8).pin_memory()
    for _ in range(2):
        g_b.fill_(1)
        h.process_frames_ptr(hM.data_ptr(), mo, hQ.data_ptr(), qo, hK.data_ptr(), capi.PTR_HOST, (dI.data_ptr(), g_b.data_ptr(), n_global))
        torch.cuda.synchronize()
        assert torch.equal(keep_a.cpu(), hK) and torch.equal(g_a, g_b)
    h.close()
