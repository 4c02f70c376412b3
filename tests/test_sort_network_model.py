"""CPU model of R-GPF's z-sort as built in erasor_b200/csrc/kernels.cu (bitonic_warp_sort / bitonic_group_sort /
group_packed_zsort): the same network, stage for stage, on numpy arrays laid out as (thread, register), plus the packed
32-bit key with its exact fix-up.  It pins the index arithmetic of the network (directions, partners, the cross-warp levels)
and the monotonicity argument of the packed key independently of a GPU; the CUDA code itself is checked against the oracle in
tests/test_gpu_parity.py::test_rgpf_sort_classes_and_ties."""
import numpy as np
import pytest

F32 = np.float32


def _cmpx(v, r, r2, asc):
    a, b = v[:, r].copy(), v[:, r2].copy()
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    v[:, r] = np.where(asc, lo, hi)
    v[:, r2] = np.where(asc, hi, lo)


def _lane_stages(v, E, asc):
    j = E >> 1
    while j > 0:
        for r in range(E):
            if (r & j) == 0:
                _cmpx(v, r, r | j, asc)
        j >>= 1


def _shfl_stages(v, E, dmax, gt, lane, asc):
    d = dmax
    while d > 0:
        keep_min = ((lane & d) == 0) == asc
        for r in range(E):
            o = v[gt ^ d, r]
            v[:, r] = np.where(keep_min, np.minimum(v[:, r], o), np.maximum(v[:, r], o))
        d >>= 1


def _warp_sort(v, E, gt, lane, asc_top):
    k = 2
    while k < E:
        j = k >> 1
        while j > 0:
            for r in range(E):
                if (r & j) == 0:
                    _cmpx(v, r, r | j, np.full(len(gt), (r & k) == 0))
            j >>= 1
        k <<= 1
    _lane_stages(v, E, (lane & 1) == 0)
    for dmax, bit in ((1, 2), (2, 4), (4, 8), (8, 16)):
        _shfl_stages(v, E, dmax, gt, lane, (lane & bit) == 0)
        _lane_stages(v, E, (lane & bit) == 0)
    _shfl_stages(v, E, 16, gt, lane, asc_top)
    _lane_stages(v, E, asc_top)


def group_sort(vals, E, NW):
    """vals[(thread, register)] -> sorted by position p = thread * E + register"""
    T = NW * 32
    v = vals.reshape(T, E).copy()
    gt = np.arange(T)
    lane, warp = gt & 31, gt >> 5
    _warp_sort(v, E, gt, lane, np.full(T, True) if NW == 1 else (warp & 1) == 0)
    lvl = 2
    while lvl <= NW:
        asc = (warp & lvl) == 0
        dw = lvl >> 1
        while dw > 0:
            keep_min = ((warp & dw) == 0) == asc
            partner = ((warp ^ dw) << 5) | lane
            for r in range(E):
                o = v[partner, r]
                v[:, r] = np.where(keep_min, np.minimum(v[:, r], o), np.maximum(v[:, r], o))
            dw >>= 1
        _shfl_stages(v, E, 16, gt, lane, asc)
        _lane_stages(v, E, asc)
        lvl <<= 1
    return v.reshape(-1)


def z_sort_key(z):
    u = z.view(np.uint32).copy()
    u[u == 0x80000000] = 0
    return np.where(u & 0x80000000, ~u, u | np.uint32(0x80000000)).astype(np.uint32)


def _place(words, n, E, G, pad):
    """slot (thread, r) <- element e = r * G + thread, as the kernels load them"""
    vals = np.full((G, E), pad, dtype=words.dtype)
    for r in range(E):
        e = r * G + np.arange(G)
        ok = e < n
        vals[ok, r] = words[e[ok]]
    return vals.reshape(-1)


@pytest.mark.parametrize("E", [4, 8, 16])
@pytest.mark.parametrize("NW", [1, 4, 8])      # 4: the 128-thread class planned in DESIGN.md section 10
def test_network_64bit_is_a_stable_sort(E, NW):
    rng = np.random.default_rng(E * 10 + NW)
    G, N = NW * 32, NW * 32 * E
    for n in (1, 2, N // 3, N - 1, N):
        z = rng.integers(0, 37, size=n).astype(np.uint64)
        words = (z << np.uint64(32)) | np.arange(n, dtype=np.uint64)
        out = group_sort(_place(words, n, E, G, np.uint64(0xFFFFFFFFFFFFFFFF)), E, NW)
        order = (out[:n] & np.uint64(0xFFFFFFFF)).astype(np.int64)
        assert np.array_equal(order, np.argsort(z, kind="stable"))


def packed_sort(z, E, NW, cap):
    """group_packed_zsort: returns (order, collisions) or (None, collisions) when the list would overflow"""
    n, G = len(z), NW * 32
    IB = 9 if NW == 1 else 12
    QMAX = (1 << (32 - IB)) - 1
    mn, mx = z.min(), z.max()
    D = F32(mx - mn)
    with np.errstate(all="ignore"):
        scale = F32(QMAX) / D if D > 0 else F32(0)
    if not (scale <= F32(3.0e38)):
        scale = F32(0)
    t = ((z - mn).astype(F32) * scale).astype(F32)
    q = np.minimum(np.trunc(t).astype(np.uint64), QMAX).astype(np.uint32)
    assert np.all(np.diff(q[np.argsort(z, kind="stable")].astype(np.int64)) >= 0), "q must be monotone in z"
    words = (q << np.uint32(IB)) | np.arange(n, dtype=np.uint32)
    out = group_sort(_place(words, n, E, G, np.uint32(0xFFFFFFFF)), E, NW)[:n]
    order = (out & np.uint32((1 << IB) - 1)).astype(np.int64)
    qs = out >> np.uint32(IB)
    lst = [p for p in range(n - 1) if qs[p] == qs[p + 1]]
    if len(lst) > cap:
        return None, len(lst)
    k = z_sort_key(z)
    while True:
        swapped = False
        for phase in (0, 1):
            for p in lst:
                if (p & 1) == phase:
                    a, b = order[p], order[p + 1]
                    if k[a] > k[b] or (k[a] == k[b] and a > b):
                        order[p], order[p + 1] = b, a
                        swapped = True
        if not swapped:
            return order, len(lst)


@pytest.mark.parametrize("E,NW,cap", [(4, 1, 256), (16, 1, 256), (8, 8, 1024)])
def test_packed_key_with_fixup_is_exact(E, NW, cap):
    rng = np.random.default_rng(E + NW)
    N = NW * 32 * E
    fallbacks = 0
    for t in range(40):
        n = int(rng.integers(2, N + 1))
        kind = t % 5
        if kind == 0:
            z = rng.normal(-0.9, 0.03, n)
        elif kind == 1:
            z = np.concatenate([rng.normal(-0.9, 0.03, n - n // 4), rng.uniform(-1.2, 3.0, n // 4)])
        elif kind == 2:
            z = rng.choice(np.array([-0.0, 0.0, 1.0, -1.0, 1.0000001]), n)
        elif kind == 3:
            z = np.concatenate([np.full(n // 2, -0.9) + rng.integers(0, 5, n // 2) * 6e-8, rng.uniform(-1.2, 3.0, n - n // 2)])
        else:
            z = rng.uniform(-1e-30, 1e-30, n)
        z = rng.permutation(z.astype(F32))
        order, ncol = packed_sort(z, E, NW, cap)
        expect = np.lexsort((np.arange(n), z_sort_key(z)))
        if order is None:
            fallbacks += 1          # the kernel runs the 64-bit network instead (exact by itself, tested above)
            continue
        assert np.array_equal(order, expect), (kind, n, ncol)
    assert fallbacks < 20
