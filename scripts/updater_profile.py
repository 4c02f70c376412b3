"""A few processed nodes through the device-resident OfflineMapUpdater (for ncu launch lists)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from erasor_b200 import capi, params, synth
p, map_world, _poses, _qs = bench.load_workload("seq05", 0, 1, 20)
up, ep = params.updater_preset("seq_05"), params.preset("seq_05")
up.removal_interval = 1
scene = synth.Scene(seed=5, length=160.0, n_nodes=161, n_dynamic=12)
u = capi.Updater(up, ep, map_world)
for k in (7, 15, 23, 31):
    s = scene.scan(k, seed_offset=17)
    t0 = time.perf_counter()
    u.process_node(k, scene.pose7(k), s)
    print(k, len(s), "ms", round(1000 * (time.perf_counter() - t0), 3), "map", u.map_size(), "fused phases us", [None if x is None else round(x / 1000, 1) for x in u.fused_profile()])
u.close()
