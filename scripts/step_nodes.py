"""Run the bench workload's map-resident step (erasor_process_nodes) a few times: driver for ncu captures (no timing, no oracle).
usage: step_nodes.py [n_steps] [config] [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from erasor_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
config = sys.argv[2] if len(sys.argv) > 2 else "seq05"
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 20
p, map_world, poses, qs = bench.load_workload(config, 0, 1, frames)
qo = np.cumsum([0] + [len(q) for q in qs]).astype(np.uint64)
Q = np.concatenate(qs)
m = capi.Map(map_world)
h = capi.Handle(p)
h.attach_map(m)
for _ in range(n):
    keep, _ = h.process_nodes(poses, Q, qo)
nv, nf, nr = h.node_stats()
print("kept", int(keep.sum()), "of", len(keep), "mean voi", int(nv.mean()), "flagged bins", int(nf.sum()), "rejected", int(nr.sum()))
