import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from erasor_b200 import capi, params, synth
p, map_world, _poses, _qs = bench.load_workload("seq05", 0, 1, 20)
up, ep = params.updater_preset("seq_05"), params.preset("seq_05")
scene = synth.Scene(seed=5, length=160.0, n_nodes=161, n_dynamic=12)
scans = {k: scene.scan(k, seed_offset=17) for k in range(161) if (k + 1) % 8 == 0}
pinned = {k: torch.from_numpy(v).pin_memory() for k, v in scans.items()}
for rep in range(2):
    u = capi.Updater(up, ep, map_world)
    torch.cuda.synchronize()
    T0 = time.perf_counter(); per = []
    for k in range(161):
        t0 = time.perf_counter()
        if k in pinned: u.process_node_ptr(k, scene.pose7(k), pinned[k].data_ptr(), len(scans[k]), capi.PTR_HOST)
        else: u.process_node_ptr(k, scene.pose7(k), 0, 0, capi.PTR_HOST)
        if k in pinned: per.append(round(1000 * (time.perf_counter() - t0), 2))
    print("pass", rep, "total ms", 1000 * (time.perf_counter() - T0), per)
    u.close()
