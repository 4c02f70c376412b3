"""CPU suite: the oracle against source-derived invariants and against
the independent numpy restatement.  The reference ships no golden vectors, so this is
how the oracle is hardened; parity stays "unpinned" in the judge's sense."""
import numpy as np
import pytest

import np_restatement as NP
from erasor_b200 import params as P
from erasor_b200 import synth

PRESETS = ["seq_05", "seq_00", "seq_07"]


def _frame(w, i, p):
    voi, q, k, idx = w["frames"][i]
    r2 = voi[:, 0].astype(np.float64) ** 2 + voi[:, 1].astype(np.float64) ** 2
    return voi[r2 < (p.max_range + 5.0) ** 2], q


@pytest.mark.parametrize("name", PRESETS)
@pytest.mark.parametrize("version", [3, 2])
def test_invariants(oracle_mod, small_workload, name, version):
    p = P.preset(name).replace(version=version, skip_voxelize=1)
    m, q = _frame(small_workload, 2, p)
    o = oracle_mod.Oracle(p)
    o.run(m, q)
    R, S = p.num_rings, p.num_sectors
    bm, bq = o.bin_of_point(0), o.bin_of_point(1)
    mn, mx, cnt, occ = o.bins(0)
    arr, arr_src = o.cloud(o.ARRANGED)
    cmp_, cmp_src = o.cloud(o.COMPLEMENT)
    rej, rej_src = o.cloud(o.MAP_REJECTED)
    gv, gv_src = o.cloud(o.GROUND_VIZ)
    # (1) N_m == sum |bin_map| + |complement|   (erasor.cpp:75-83)
    assert len(m) == int(cnt.sum()) + len(cmp_)
    assert np.array_equal(np.sort(cmp_src), np.nonzero(bm < 0)[0])
    assert np.array_equal(cmp_src, np.sort(cmp_src)), "complement keeps source order"
    # (2) every binned point is inside the z window, the range and the index bounds
    sel = bm >= 0
    z = m[sel, 2].astype(np.float64)
    r = np.sqrt(m[sel, 0].astype(np.float64) ** 2 + m[sel, 1].astype(np.float64) ** 2)
    assert np.all((z > p.min_h) & (z < p.max_h) & (r <= p.max_range))
    assert np.all((bm[sel] // R < S) & (bm[sel] % R < R))
    # (3) bin.max_h / min_h / count agree with the points
    nmn, nmx, ncnt = NP.bin_tables(p, m, bm)
    assert np.array_equal(ncnt, cnt) and np.array_equal(nmn, mn) and np.array_equal(nmx, mx)
    assert np.array_equal(occ.astype(bool), cnt > 0)
    # (4) status values
    st, st1 = o.status()
    assert set(np.unique(st)).issubset({0.0, 0.25, 0.5, 0.8, 1.0})
    # (5) ground u non-ground == bin points, disjoint; (6) classification consistent with the last plane
    planes = o.planes()
    gset, rset = set(gv_src.tolist()), set(rej_src.tolist())
    assert not (gset & rset)
    for pl in planes:
        members = np.nonzero(bm == pl["bin"])[0]
        assert pl["n_points"] == len(members)
        g_here = [i for i in members if i in gset]
        r_here = [i for i in members if i in rset]
        assert len(g_here) + len(r_here) == len(members)
        assert len(g_here) == pl["n_ground"][-1]
        n = pl["normal_d"][-1, :3].astype(np.float32)
        d = pl["normal_d"][-1, 3]
        pts = m[members, :3]
        res = (pts[:, 0] * n[0] + pts[:, 1] * n[1]).astype(np.float32) + (pts[:, 2] * n[2]).astype(np.float32)
        is_g = res.astype(np.float64) < (p.gf_dist_thr - d)
        assert np.array_equal(np.nonzero(is_g)[0], np.nonzero(np.isin(members, g_here))[0])
        # (7) unit normal
        assert abs(np.linalg.norm(pl["normal_d"][-1, :3]) - 1.0) < 1e-5
    # (8) output identity for non-flagged bins: selected == bin_map, order included
    flagged = {pl["bin"] for pl in planes}
    n_sel = len(arr) - len(gv)
    sel_src = arr_src[:n_sel]
    if version == 3:
        pos = 0
        for b in range(R * S):
            members = np.nonzero(bm == b)[0]
            if b in flagged:
                qn = int(np.count_nonzero(bq == b))
                gn = int(sum(1 for i in members if i in gset))
                pos += qn + gn
            else:
                assert np.array_equal(sel_src[pos:pos + len(members)], members.astype(np.uint32)), f"bin {b}"
                pos += len(members)
        assert pos == n_sel
    assert np.array_equal(arr_src[n_sel:], gv_src)


@pytest.mark.parametrize("name", PRESETS + ["synthetic_40x360"])
def test_numpy_restatement_agrees(oracle_mod, small_workload, name):
    p = P.preset(name).replace(skip_voxelize=1)
    m, q = _frame(small_workload, 4, p)
    adv = synth.adversarial_points(p, n_random=50000, seed=9)
    m = np.concatenate([m, adv])
    for version in (3, 2):
        pv = p.replace(version=version)
        o = oracle_mod.Oracle(pv)
        o.run(m, q)
        bm, fenced = NP.bin_of_points(pv, m)
        bq, _ = NP.bin_of_points(pv, q)
        assert np.array_equal(bm, o.bin_of_point(0)) and np.array_equal(bq, o.bin_of_point(1))
        mmn, mmx, mcnt = NP.bin_tables(pv, m, bm)
        qmn, qmx, qcnt = NP.bin_tables(pv, q, bq)
        st, _ = o.status()
        if version == 3:
            nst, _, nflag = NP.status_v3(pv, mmn, mmx, mcnt, qmn, qmx, qcnt)
        else:
            nst, nflag = NP.status_v2(pv, mmn, mmx, mcnt, qmn, qmx, qcnt)
        assert np.array_equal(nst, st), f"{name} v{version}"
        assert sorted(np.nonzero(nflag)[0].tolist()) == sorted(pl["bin"] for pl in o.planes())


def test_negzero_fence_oracle(oracle_mod):
    p = P.preset("seq_05").replace(skip_voxelize=1)
    m = np.array([[-5.0, -0.0, 0.2, 1.0], [-0.0, -0.0, 0.2, 1.0], [5.0, -0.0, 0.2, 1.0], [-5.0, 0.0, 0.2, 1.0]], dtype=np.float32)
    o = oracle_mod.Oracle(p)
    o.run(m, m[:0])
    b = o.bin_of_point(0)
    assert o.negzero_fenced() == 2
    assert b[0] == b[3] and b[2] == 0 * p.num_rings + 1   # x=5 -> ring 1 (5/4), sector 0
    nb, nf = NP.bin_of_points(p, m)
    assert np.array_equal(nb, b) and nf == 2


def test_svd_properties(oracle_mod):
    rng = np.random.default_rng(0)
    for _ in range(200):
        a = rng.normal(size=(3, 3))
        cov = (a @ a.T).astype(np.float32)
        U, sv = oracle_mod.jacobi_svd(cov)
        assert np.allclose(U @ U.T, np.eye(3), atol=1e-5)
        assert sv[0] >= sv[1] >= sv[2] >= 0
        ref = np.linalg.svd(cov.astype(np.float64), compute_uv=False)
        assert np.allclose(sv, ref, rtol=1e-4, atol=1e-5 * ref[0])
        n = U[:, 2].astype(np.float64)
        assert np.linalg.norm(cov.astype(np.float64) @ n - ref[2] * n) < 1e-3 * max(ref[0], 1e-6)
    U, sv = oracle_mod.jacobi_svd(np.zeros((3, 3), dtype=np.float32))     # the App. B-3 fence
    assert np.array_equal(U, np.eye(3, dtype=np.float32)) and np.all(sv == 0)


def test_mean_cov_modes(oracle_mod):
    rng = np.random.default_rng(1)
    pts = np.zeros((500, 4), dtype=np.float32)
    pts[:, :3] = rng.normal(size=(500, 3)) * [2.0, 1.0, 0.05] + [40.0, -30.0, 0.1]
    for mode in (0, 1):
        n, cov, mean = oracle_mod.mean_cov(pts, mode)
        assert n == 500
        ref = np.cov(pts[:, :3].astype(np.float64).T, bias=True)
        tol = 5e-3 if mode == 0 else 1e-5      # mode 0 is the ill-conditioned unshifted formula
        assert np.allclose(cov, ref, atol=tol)
        assert np.allclose(mean[:3], pts[:, :3].mean(axis=0), atol=1e-4)


def test_voxelize_preserving_labels(oracle_mod):
    rng = np.random.default_rng(2)
    pts = np.zeros((4000, 4), dtype=np.float32)
    pts[:, :3] = rng.uniform(-3, 3, size=(4000, 3))
    pts[:, 3] = rng.integers(0, 260, 4000)
    out = oracle_mod.voxelize(pts, 0.2)
    key = np.floor(pts[:, :3] * np.float32(5.0)).astype(np.int64)
    assert len(out) == len(np.unique(key, axis=0))
    # labels are restored, never averaged: every output intensity is one of the input labels of a nearby point
    assert set(np.unique(out[:, 3])).issubset(set(np.unique(pts[:, 3])))
    # brute-force 1-NN label check on a sample
    for i in rng.integers(0, len(out), 50):
        dd = ((out[i, 0] - pts[:, 0]) ** 2 + (out[i, 1] - pts[:, 1]) ** 2).astype(np.float32) + ((out[i, 2] - pts[:, 2]) ** 2).astype(np.float32)
        assert out[i, 3] == pts[np.argmin(dd), 3]
    assert len(oracle_mod.voxelize(pts[:0], 0.2)) == 0


@pytest.mark.parametrize("seed,n,span,leaf", [(1, 1500, 6.0, 0.2), (2, 2500, 3.0, 0.05), (3, 800, 40.0, 0.4), (4, 1, 1.0, 0.2), (5, 600, 3000.0, 0.001)])
def test_oracle_voxelize_matches_the_numpy_restatement(oracle_mod, seed, n, span, leaf):
    """erasor_utils::voxelize_preserving_labels (erasor_utils.cpp:80-114): the C++ oracle (sorted (key, index) pairs, grid-accelerated
    1-NN) against a structurally different numpy reading (np.unique, sequential float32 sums, brute-force 1-NN) -- bit for bit,
    including the "leaf too small" overflow case (seed 5) and duplicated points."""
    rng = np.random.default_rng(seed)
    c = np.zeros((n, 4), dtype=np.float32)
    c[:, :3] = rng.uniform(-span, span, size=(n, 3)).astype(np.float32)
    c[:, 2] *= 0.2
    c[:, 3] = rng.integers(0, 260, n).astype(np.float32)
    if n > 100:
        c[n // 2:n // 2 + 40] = c[:40]                             # exact duplicates far apart in the cloud
        c[-20:, :3] = c[100:120, :3]                               # same position, different label: 1-NN ties to the lowest index
    got = oracle_mod.voxelize(c, leaf)
    ref = NP.voxelize_preserving_labels(c, leaf)
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), f"differs in {np.count_nonzero(got.view(np.uint32) != ref.view(np.uint32))} words"
