"""Run-to-run spread of the overlapped-lanes measurement (same loop as bench.py's `value`), per lane count."""
import os, sys, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from erasor_b200 import capi
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
p, map_world, poses, qs = bench.load_workload("seq05", 0, 1, frames)
dev = torch.device("cuda", 0)
qo = np.cumsum([0] + [len(q) for q in qs]).astype(np.uint64)
Q = np.ascontiguousarray(np.concatenate(qs), dtype=np.float32)
gmap = capi.Map(map_world, device=0)
LMAX = 6
lanes = [capi.Handle(p, device=0) for _ in range(LMAX)]
for h in lanes: h.attach_map(gmap)
n_copies = 8
dQ = [torch.from_numpy(Q).to(dev) for _ in range(n_copies)]
hQ = torch.from_numpy(Q).pin_memory()
hK = [torch.empty(len(map_world), dtype=torch.uint8).pin_memory() for _ in range(LMAX)]
xs = torch.cuda.ExternalStream(lanes[0].stream, device=dev)
def sub_res(i, lane): lanes[lane].process_nodes_ptr(poses, dQ[i % n_copies].data_ptr(), qo, 0.0, 0, 0, capi.PTR_DEVICE, asynchronous=True)
def sub_host(i, lane): lanes[lane].process_nodes_ptr(poses, hQ.data_ptr(), qo, 0.0, 0, hK[lane].data_ptr(), capi.PTR_HOST, asynchronous=True)
def timed(submit, L, warm):
    for i in range(warm):
        lanes[i % L].wait(); submit(i, i % L)
    for h in lanes: h.wait()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(xs)
    for i in range(steps):
        lanes[i % L].wait(); submit(warm + i, i % L)
    for h in lanes[:L]: h.wait()
    e1.record(xs)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1000, (time.perf_counter() - t0) / steps * 1e6
gc.disable()
quick = len(sys.argv) > 3
for L in ((1, 4) if quick else (1, 2, 3, 4, 6)):
    timed(sub_res, L, n_copies * L)
    r = [timed(sub_res, L, 2 * L) for _ in range(4 if quick else 10)]
    print(f"resident lanes={L} us/step (events):", [round(a) for a, _ in r], "wall:", [round(b) for _, b in r], "scans/s median", round(frames / np.median([a for a, _ in r]) * 1e6))
for L in (() if quick else (1, 3, 4, 6)):
    timed(sub_host, L, 2 * L)
    r = [timed(sub_host, L, 2 * L) for _ in range(10)]
    print(f"host lanes={L} us/step:", [round(a) for a, _ in r], "scans/s median", round(frames / np.median([a for a, _ in r]) * 1e6))
# CPU cost of a submit + wait when the GPU is idle
t0 = time.perf_counter()
for i in range(50):
    sub_res(i, 0); lanes[0].wait()
print("one-lane submit+wait wall us", (time.perf_counter() - t0) / 50 * 1e6)
