"""Run the bench workload's mask-mode step a few times (driver for ncu captures; no timing, no oracle)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from erasor_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
p, mw, maps, qs, idxs = bench.load_workload(0, 1, 20)
mo = np.cumsum([0] + [len(m) for m in maps]).astype(np.uint64); qo = np.cumsum([0] + [len(q) for q in qs]).astype(np.uint64)
M = np.concatenate(maps); Q = np.concatenate(qs)
h = capi.Handle(p)
for _ in range(n):
    keep = h.process_frames(M, mo, Q, qo)
print("kept", int(keep.sum()), "of", len(keep))
