"""Generates tests/golden/*.npz -- small known-answer vectors for the R-POD -> SRT -> R-GPF path.

The reference ships NO golden vectors, tests or fixtures and cannot be compiled or imported in
this image, so these are produced by the CPU oracle (oracle/), not by the reference: they freeze the oracle's
answers (regression anchor for the oracle itself, and a second target for the CUDA path), they do not pin parity
to the reference.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from erasor_b200 import params as P  # noqa: E402
from erasor_b200 import synth  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    ("seq05_v3", "seq_05", dict(version=3, skip_voxelize=0)),
    ("seq05_v3_novox", "seq_05", dict(version=3, skip_voxelize=1)),
    ("seq05_v2", "seq_05", dict(version=2, skip_voxelize=1)),
    ("seq00_v3", "seq_00", dict(version=3, skip_voxelize=0)),
    ("seq07_v3_shiftedcov", "seq_07", dict(version=3, skip_voxelize=0, cov_mode=1)),
]


def main():
    w = synth.make_frames(seed=23, n_frames=3, preset_max_range=80.0, n_map_nodes=17, n_beams=24, n_az=480, length=60.0, n_dynamic=6)
    voi, q, k, idx = w["frames"][1]
    adv = synth.adversarial_points(P.preset("seq_05"), n_random=500, seed=2)[:3000]
    m = np.concatenate([voi, adv]).astype(np.float32)
    for name, preset, kw in CASES:
        p = P.preset(preset).replace(**kw)
        o = O.Oracle(p)
        o.run(m, q)
        mn, mx, cnt, _ = o.bins(0)
        qmn, qmx, qcnt, _ = o.bins(1)
        st, st1 = o.status()
        pl = o.planes()
        arr, arr_src = o.cloud(o.ARRANGED)
        cmp_, _ = o.cloud(o.COMPLEMENT)
        rej, rej_src = o.cloud(o.MAP_REJECTED)
        crej, _ = o.cloud(o.CURR_REJECTED)
        gv, gv_src = o.cloud(o.GROUND_VIZ)
        np.savez_compressed(
            os.path.join(HERE, f"{name}.npz"), preset=preset, overrides=np.array(sorted(kw.items()), dtype=object).astype(str),
            map_voi=m, query_voi=q, bin_map=o.bin_of_point(0), bin_query=o.bin_of_point(1),
            map_cnt=cnt, map_min=mn, map_max=mx, query_cnt=qcnt, query_min=qmn, query_max=qmx,
            status=st, plane_bins=np.array([x["bin"] for x in pl], dtype=np.int32),
            plane_normal_d=np.array([x["normal_d"] for x in pl]).reshape(len(pl), p.gf_iter, 4),
            plane_n_ground=np.array([x["n_ground"] for x in pl], dtype=np.int32).reshape(len(pl), p.gf_iter),
            plane_n_seeds=np.array([x["n_seeds"] for x in pl], dtype=np.int32), plane_lpr=np.array([x["lpr"] for x in pl]),
            arranged=arr, complement=cmp_, map_rejected=rej, curr_rejected=crej, rejected_src=rej_src, ground_src=gv_src)
        print(name, "map", len(m), "query", len(q), "flagged", len(pl), "rejected", len(rej), "arranged", len(arr))


if __name__ == "__main__":
    main()
