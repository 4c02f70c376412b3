"""Blast radius of the choices the reference leaves to its libraries.

Parity is pinned to the builder's oracle, not to a PCL / Eigen build (none exists in this image).  This script measures, on
the five golden inputs, how far the outputs move when each pinned choice is swapped for another LEGAL outcome:
  sort ties      std::sort's order of equal z (erasor.cpp:240): stable order (pinned) vs libstdc++'s introsort vs reverse-stable
  covariance     pcl::computeMeanAndCovarianceMatrix of PCL <= 1.10 (pinned, unshifted) vs PCL >= 1.11 (shifted)
  classification the n x 3 GEMV at erasor.cpp:271 without FMA (pinned: x86-64 baseline build) vs FMA-contracted (-march=native)
  1-NN ties      FLANN's nearest neighbour among equidistant points: lowest index (pinned) vs highest
CPU only (oracle).  Prints a markdown table; DESIGN.md section 2 quotes it.
"""
import ctypes
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from erasor_b200 import params as P  # noqa: E402
from oracle import oracle_py as O  # noqa: E402


def run(p, m, q, study=(0, 0, 0)):
    L = O.lib()
    L.oracle_set_study.argtypes = [ctypes.c_int] * 3
    L.oracle_set_study(*study)
    o = O.Oracle(p)
    o.run(m, q)
    L.oracle_set_study(0, 0, 0)
    _, rej = o.cloud(o.MAP_REJECTED)
    arr, _ = o.cloud(o.ARRANGED)
    pl = o.planes()
    return dict(rej=set(rej.tolist()), arranged=arr, planes=pl)


def diff(a, b):
    d_pts = len(a["rej"] ^ b["rej"])
    dn, dbins = 0.0, 0
    for x, y in zip(a["planes"], b["planes"]):
        if x["bin"] != y["bin"]:
            continue
        e = float(np.abs(x["normal_d"] - y["normal_d"]).max())
        dn = max(dn, e)
        dbins += int(e > 1e-4)
    same_arr = a["arranged"].shape == b["arranged"].shape and np.array_equal(a["arranged"].view(np.uint32), b["arranged"].view(np.uint32))
    return d_pts, dn, dbins, same_arr


def main():
    rows = []
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
        z = np.load(path, allow_pickle=True)
        kw = {k: int(v) for k, v in z["overrides"]}
        p = P.preset(str(z["preset"])).replace(**kw)
        m, q = z["map_voi"], z["query_voi"]
        base = run(p, m, q)
        variants = {
            "sort ties: libstdc++ std::sort": run(p.replace(sort_mode=0), m, q),
            "sort ties: reverse-stable": run(p, m, q, (1, 0, 0)),
            "covariance: other PCL generation": run(p.replace(cov_mode=1 - p.cov_mode), m, q),
            "classification GEMV: FMA-contracted": run(p, m, q, (0, 1, 0)),
            "1-NN ties: highest index": run(p, m, q, (0, 0, 1)),
        }
        for name, v in variants.items():
            d_pts, dn, dbins, same = diff(base, v)
            rows.append((os.path.basename(path)[:-4], name, len(base["planes"]), len(base["rej"]), d_pts, dn, dbins, same))
    # two full-size frames of the bench's seq-05 twin (dup-heavy voxel centroids, 1200+ flagged bins' worth of variety)
    from erasor_b200 import synth
    w = synth.make_frames(seed=5, n_frames=6, preset_max_range=80.0, n_map_nodes=41, n_beams=32, n_az=900)
    for name, preset in (("twin frame 1 (seq_05)", "seq_05"), ("twin frame 4 (seq_00)", "seq_00")):
        p = P.preset(preset).replace(skip_voxelize=0)
        fi = 1 if "frame 1" in name else 4
        voi, idx = O.fetch_voi(w["map_world"], w["scene"].pose7(w["frames"][fi][2]), p.max_range)
        q = w["frames"][fi][1]
        base = run(p, voi, q)
        for vname, v in {"sort ties: libstdc++ std::sort": run(p.replace(sort_mode=0), voi, q), "sort ties: reverse-stable": run(p, voi, q, (1, 0, 0)),
                         "covariance: other PCL generation": run(p.replace(cov_mode=1), voi, q), "classification GEMV: FMA-contracted": run(p, voi, q, (0, 1, 0)),
                         "1-NN ties: highest index": run(p, voi, q, (0, 0, 1))}.items():
            d_pts, dn, dbins, same = diff(base, v)
            rows.append((name, vname, len(base["planes"]), len(base["rej"]), d_pts, dn, dbins, same))
    print("| golden case | variant | flagged bins | rejected pts (pinned) | rejected set differs in | max |normal,d| delta | bins beyond 1e-4 | arranged cloud identical |")
    print("|---|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]} | {r[4]} pts | {r[5]:.3g} | {r[6]} | {'yes' if r[7] else 'no'} |")


if __name__ == "__main__":
    main()
