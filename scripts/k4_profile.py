"""Per-phase cycle profile of R-GPF (K4) on the bench workload (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from erasor_b200 import capi
p, mw, maps, qs, idxs = bench.load_workload(0, 1, 20)
mo = np.cumsum([0] + [len(m) for m in maps]).astype(np.uint64); qo = np.cumsum([0] + [len(q) for q in qs]).astype(np.uint64)
M = np.concatenate(maps); Q = np.concatenate(qs)
h = capi.Handle(p)
for _ in range(3):
    keep = h.process_frames(M, mo, Q, qo)
npts, prof = h.rgpf_profile()
names = ["load+idxsort", "zsort", "seeds", "accumulate", "svd+plane", "classify+compact", "outputs", "sweeps"]
print("bins", len(npts), "n mean", npts.mean(), "max", npts.max())
for lo, hi in ((0, 128), (128, 512), (512, 1024), (1024, 4096)):
    sel = (npts > lo) & (npts <= hi)
    if sel.any():
        print(f"n in ({lo},{hi}]: {sel.sum()} bins; mean cycles per phase:", {k: int(prof[sel, i].mean()) for i, k in enumerate(names)}, "total", int(prof[sel, :7].sum(1).mean()), "max total", int(prof[sel, :7].sum(1).max()))
print("K3 phase cycles (frame 0): status, chunk prefix, flag scan+base, map offsets, records+queue, query offsets:", h.srt_profile()[:6].tolist())
