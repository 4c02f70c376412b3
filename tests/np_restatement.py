"""Independent numpy (float64) restatement of R-POD binning + the Scan Ratio Test status logic.

A second, structurally different reading of reference erasor.cpp:11-21, 100-144 (binning) and
:448-486, :493-563, :573-595 (v3 status) / :346-427 (v2 status), used to cross-check the C++ oracle
("an independent second restatement of SRT+binning in numpy").
TEST INFRASTRUCTURE.
"""
import numpy as np

PI_TRUNC = 3.1415926535
INF = 10000000000000.0
LITTLE, MERGE, MAP_HIGH, BLOCKED, CURR_HIGH = 0.0, 0.25, 0.5, 0.8, 1.0


def bin_of_points(p, cloud):
    """bin = sector*R + ring or -1, vectorised; same double arithmetic as the reference."""
    x = cloud[:, 0].astype(np.float64)
    y = cloud[:, 1].astype(np.float64)
    z = cloud[:, 2].astype(np.float64)
    R, S = p.num_rings, p.num_sectors
    ring_size = p.max_range / R
    sector_size = 2 * PI_TRUNC / S
    with np.errstate(invalid="ignore", over="ignore"):
        r = np.sqrt(x * x + y * y)
        ok = (z < p.max_h) & (z > p.min_h) & (r <= p.max_range)
        a = np.arctan2(y, x)
        theta = np.where(y >= 0, a, 2 * PI_TRUNC + a)
        sec = np.minimum((theta / sector_size).astype(np.int64), S - 1)
        # App. B-1 fence: negative sector (y == -0.0, x <= -0) -> y := +0
        neg = ok & (sec < 0)
        if np.any(neg):
            a2 = np.arctan2(np.zeros_like(x), x)
            sec = np.where(neg, np.minimum((a2 / sector_size).astype(np.int64), S - 1), sec)
        ring = np.minimum((r / ring_size).astype(np.int64), R - 1)
    out = np.where(ok, sec * R + ring, -1)
    return out.astype(np.int32), int(np.count_nonzero(neg))


def bin_tables(p, cloud, bins):
    B = p.num_rings * p.num_sectors
    cnt = np.bincount(bins[bins >= 0], minlength=B).astype(np.uint32)
    mx = np.full(B, -INF)
    mn = np.full(B, INF)
    z = cloud[:, 2].astype(np.float64)
    sel = bins >= 0
    np.maximum.at(mx, bins[sel], z[sel])
    np.minimum.at(mn, bins[sel], z[sel])
    return mn, mx, cnt


def status_v3(p, mmn, mmx, mcnt, qmn, qmx, qcnt):
    R, S = p.num_rings, p.num_sectors
    B = R * S
    st1 = np.zeros(B)
    flag = np.zeros(B, dtype=bool)
    for b in range(B):
        if mcnt[b] == 0:
            continue
        if qcnt[b] < p.minimum_num_pts:
            continue
        mdh = mmx[b] - mmn[b]
        cdh = qmx[b] - qmn[b]
        with np.errstate(divide="ignore", invalid="ignore"):
            a, c = np.float64(mdh) / np.float64(cdh), np.float64(cdh) / np.float64(mdh)
        ratio = c if c < a else a
        if qcnt[b] > 0:
            if ratio < p.scan_ratio_threshold:
                if mdh >= cdh:
                    st1[b] = MAP_HIGH
                elif mdh <= cdh:
                    st1[b] = CURR_HIGH
            else:
                st1[b] = MERGE
    st = st1.copy()
    for theta in range(S):
        for r in range(R):
            b = theta * R + r
            if st1[b] == MAP_HIGH:
                if (mmx[b] - mmn[b]) > 0.5:
                    flag[b] = True
                else:
                    st[b] = 0.0
            elif st1[b] == MERGE:
                close = False
                for j in (theta - 1, theta, theta + 1):
                    tj = j + R if j < 0 else (j - R if j >= S else j)     # sic: wraps with num_rings
                    if tj < 0 or tj >= S:
                        continue
                    for rr in range(max(0, r - 1), min(r + 1, R - 1) + 1):
                        if rr == r and tj == theta:
                            continue
                        if st1[tj * R + rr] == CURR_HIGH:
                            close = True
                st[b] = BLOCKED if close else MERGE
    return st, st1, flag


def status_v2(p, mmn, mmx, mcnt, qmn, qmx, qcnt):
    B = p.num_rings * p.num_sectors
    st = np.zeros(B)
    flag = np.zeros(B, dtype=bool)
    for b in range(B):
        if qcnt[b] < p.minimum_num_pts:
            continue
        if qcnt[b] > 0 and mcnt[b] > 0:
            mdh = mmx[b] - mmn[b]
            cdh = qmx[b] - qmn[b]
            with np.errstate(divide="ignore", invalid="ignore"):
                a, c = np.float64(mdh) / np.float64(cdh), np.float64(cdh) / np.float64(mdh)
            ratio = c if c < a else a
            if ratio < p.scan_ratio_threshold:
                if mdh >= cdh:
                    st[b] = MAP_HIGH
                    flag[b] = mmx[b] > p.th_bin_max_h
                elif mdh <= cdh:
                    st[b] = CURR_HIGH
            else:
                st[b] = MERGE
    return st, flag


# ------------------------------------------------------------------------------------------------
# erasor_utils::voxelize_preserving_labels (reference erasor_utils.cpp:80-114): pcl::VoxelGrid<PointXYZI> with a cubic leaf
# (PCL 1.8: downsample_all_data, no minimum points per voxel), then the intensity of the nearest source point (1-NN) copied
# onto every centroid.  A second, structurally different reading -- float32 numpy arrays, np.unique instead of a sort of
# (key, index) pairs, a brute-force distance matrix instead of a search grid -- with the same two pinned choices as the oracle
# (members summed in cloud order; 1-NN ties to the lowest index).  For small clouds (the 1-NN is O(voxels x points)).
# ------------------------------------------------------------------------------------------------
def voxelize_preserving_labels(cloud, leaf_size):
    f32 = np.float32
    c = np.ascontiguousarray(cloud, dtype=f32).reshape(-1, 4)
    n = len(c)
    if n == 0:
        return np.zeros((0, 4), dtype=f32)
    leaf = f32(leaf_size)
    inv = f32(1.0) / leaf
    xyz = c[:, :3]
    mn, mx = xyz.min(axis=0), xyz.max(axis=0)                       # getMinMax3D
    d = ((mx - mn) * inv).astype(np.int64) + 1                      # truncation like the reference's int64 cast
    if int(d[0]) * int(d[1]) * int(d[2]) > 2147483647:             # "Leaf size is too small for the input dataset"
        vox = c.copy()
    else:
        min_b = np.floor(mn * inv).astype(np.int32)
        max_b = np.floor(mx * inv).astype(np.int32)
        div_b = max_b - min_b + 1
        ijk = (np.floor(xyz * inv) - min_b.astype(f32)).astype(np.int32)            # float subtraction, then the int cast
        key = (ijk[:, 0].astype(np.int64) + ijk[:, 1].astype(np.int64) * int(div_b[0]) +
               ijk[:, 2].astype(np.int64) * int(div_b[0]) * int(div_b[1])).astype(np.int32).astype(np.uint32)   # int32 arithmetic, unsigned order
        keys, inverse = np.unique(key, return_inverse=True)          # ascending voxel key = output order
        vox = np.zeros((len(keys), 4), dtype=f32)
        cnt = np.zeros(len(keys), dtype=np.int64)
        for i in range(n):                                           # cloud order: sequential float32 sums per voxel
            v = inverse[i]
            vox[v] += c[i]
            cnt[v] += 1
        vox /= cnt.astype(f32)[:, None]
    out = vox.copy()
    for v in range(len(vox)):                                        # exact 1-NN, ties to the lowest cloud index
        dx = vox[v, 0] - xyz[:, 0]
        dy = vox[v, 1] - xyz[:, 1]
        dz = vox[v, 2] - xyz[:, 2]
        d2 = (dx * dx + dy * dy) + dz * dz                           # float32, the association of pcl's squared distance
        out[v, 3] = c[int(np.argmin(d2)), 3]                         # argmin returns the first minimum
    return out
