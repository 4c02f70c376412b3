"""Per-phase cycle profile of R-GPF (K4) on the bench workload, node mode (run on the GPU box).
usage: k4_profile.py [config] [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from erasor_b200 import capi  # noqa: E402

config = sys.argv[1] if len(sys.argv) > 1 else "seq05"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 20
p, map_world, poses, qs = bench.load_workload(config, 0, 1, frames)
qo = np.cumsum([0] + [len(q) for q in qs]).astype(np.uint64)
Q = np.concatenate(qs)
m = capi.Map(map_world)
h = capi.Handle(p)
h.attach_map(m)
for _ in range(3):
    h.process_nodes(poses, Q, qo)
npts, prof = h.rgpf_profile()
names = ["load+idxsort", "zsort", "seeds", "accumulate", "svd+plane", "classify+compact", "outputs", "sweeps"]
print("bins", len(npts), "n mean", npts.mean(), "max", npts.max(), "sum", npts.sum())
for lo, hi in ((0, 64), (64, 128), (128, 256), (256, 512), (512, 1024), (1024, 2560), (2560, 1 << 30)):
    sel = (npts > lo) & (npts <= hi)
    if sel.any():
        print(f"n in ({lo},{hi}]: {sel.sum()} bins; mean cycles per phase:", {k: int(prof[sel, i].mean()) for i, k in enumerate(names)},
              "total", int(prof[sel, :7].sum(1).mean()), "max total", int(prof[sel, :7].sum(1).max()))
tot = prof[:, :7].sum(1)
print("critical path (max over bins): %d cycles = %.1f us at 1.965 GHz; sum over bins %.1f M cycles" % (tot.max(), tot.max() / 1965.0, tot.sum() / 1e6))
