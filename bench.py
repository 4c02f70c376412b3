#!/usr/bin/env python
"""bench.py -- LiDAR scans/sec through R-POD + SRT + R-GPF on the synthetic twin of KITTI seq 05.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` (N>1: launched under torch.distributed.run,
one rank per GPU).  One JSON line on rank 0.

Workload (BASELINE.json configs[1]): KITTI seq 05, frames 2350-2670, reference config/seq_05.yaml
(15 rings x 60 sectors @ 60 m, version 3).  161 nodes, every 8th processed (removal_interval 8) = 20 hot-path
frames per offline pass.  No KITTI data exists here, so the twin is a seeded synthetic street scene
(erasor_b200/synth.py): HDL-64-like ray-cast scans, a 0.2 m voxelised accumulated map with moving-object trails.
A "step" = one pass of the hot path over one rank's 20 frames (frame-independent mode: every frame against
the same initial map, the mode north_star shards across GPUs).  Weak scaling: every rank gets its own 20 frames.

value : scans/s, clouds already resident in HBM (C-ABI call with device pointers).
e2e   : scans/s through the same C-ABI call with pinned HOST buffers: clouds H2D and keep-masks D2H inside the
        timed region.
roofline : K1 (polar binning + per-bin min/max/count), the dominant kernel; CUDA-event time per launch on the
        library's stream, algorithmic bytes 16*(N_m+N_q) per frame (one float4 read per input point).
cpu_baseline : the oracle port (oracle/, restated reference path) on one host core, bounded sample.
--impl reference : the same oracle port over all host cores (independent frames in a process pool).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FRAMES_PER_PASS = 20          # 161 nodes / removal_interval 8 (config/seq_05.yaml)
N_INPUT_COPIES = 4            # rotate input copies so that consecutive steps never find their clouds in the 126 MB L2
K1_DRAM_TRAFFIC_BYTES = 56.861952e6 + 2.707712e6   # ncu --set full, one launch of k1_rpod_bin on the default workload (profiles/r01/ncu_step_full_raw.csv)
CACHE_DIR = os.environ.get("ERASOR_B200_CACHE", "/tmp/erasor_b200_cache")


def load_workload(rank: int, world: int, frames_per_rank: int):
    """seq-05 twin: one map, world*frames_per_rank frames along the trajectory; this rank's contiguous share."""
    from erasor_b200 import params, synth
    p = params.preset("seq_05").replace(skip_voxelize=1)
    os.makedirs(CACHE_DIR, exist_ok=True)
    total = world * frames_per_rank
    key = f"seq05_twin_seed5_n161_s2_f{total}.npz"
    path = os.path.join(CACHE_DIR, key)
    if os.path.exists(path):
        z = np.load(path)
        map_world = z["map_world"]
        vois = [z[f"voi_{i}"] for i in range(total)]
        qs = [z[f"q_{i}"] for i in range(total)]
        idxs = [z[f"idx_{i}"] for i in range(total)]
    else:
        w = synth.make_frames(seed=5, n_frames=total, preset_max_range=p.max_range, n_map_nodes=161, n_beams=64, n_az=1800,
                              length=160.0, n_dynamic=12, query_voxel=0.2, map_stride=2)
        map_world = w["map_world"]
        vois = [f[0] for f in w["frames"]]
        qs = [f[1] for f in w["frames"]]
        idxs = [f[3] for f in w["frames"]]
        if rank == 0:
            tmp = path + f".tmp{os.getpid()}.npz"
            np.savez(tmp, map_world=map_world, **{f"voi_{i}": v for i, v in enumerate(vois)}, **{f"q_{i}": q for i, q in enumerate(qs)},
                     **{f"idx_{i}": x for i, x in enumerate(idxs)})
            os.replace(tmp, path)
    lo = rank * frames_per_rank
    return p, map_world, vois[lo:lo + frames_per_rank], qs[lo:lo + frames_per_rank], idxs[lo:lo + frames_per_rank]


def offline_pass_block(p, map_world, device_index):
    """Sequential offline pass through the device-resident OfflineMapUpdater: 161 nodes, every 8th
    processed (config/seq_05.yaml removal_interval 8), raw scans from pinned host memory, map state in HBM.  Wall clock
    around erasor_updater_process_node (it synchronises its stream).  Same pass on the oracle's restated caller loop on
    one host core.  Informational: the driver's headline numbers are `value` / `e2e` above."""
    import torch
    from erasor_b200 import capi, params, synth
    from oracle import oracle_py
    up = params.updater_preset("seq_05")
    ep = params.preset("seq_05")                     # version 3 with in-bin voxelisation, as shipped
    scene = synth.Scene(seed=5, length=160.0, n_nodes=161, n_dynamic=12)
    nodes = list(range(161))
    processed = [k for k in nodes if (k + 1) % up.removal_interval == 0]
    path = os.path.join(CACHE_DIR, "seq05_twin_seed5_scans_ri8.npz")
    if os.path.exists(path):
        z = np.load(path)
        scans = {k: z[f"s_{k}"] for k in processed}
    else:
        scans = {k: scene.scan(k, seed_offset=17) for k in processed}
        tmp = path + f".tmp{os.getpid()}.npz"
        np.savez(tmp, **{f"s_{k}": v for k, v in scans.items()})
        os.replace(tmp, path)
    empty = np.zeros((0, 4), dtype=np.float32)
    pinned = {k: torch.from_numpy(v).pin_memory() for k, v in scans.items()}
    poses = [scene.pose7(k) for k in nodes]
    best = None
    pass_ms = []
    u = capi.Updater(up, ep, map_world, device=device_index)
    for rep in range(5):
        if rep:
            u.reset(map_world)                       # load_global_map again; device buffers are kept
        l0 = u.kernel_launch_count()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in nodes:
            if k in pinned:
                u.process_node_ptr(k, poses[k], pinned[k].data_ptr(), len(scans[k]), capi.PTR_HOST)
            else:
                u.process_node_ptr(k, poses[k], 0, 0, capi.PTR_HOST)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        launches = u.kernel_launch_count() - l0
        n_final = u.map_size()
        gpu_map = u.cloud(u.MAP_ARRANGED) if rep == 0 else None
        pass_ms.append(round(1000 * dt, 2))
        if best is None or dt < best[0]:
            best = (dt, launches, n_final)
        if rep == 0:
            first_map = gpu_map
    u.close()
    o = oracle_py.OracleUpdater(up, ep, map_world)
    t0 = time.perf_counter()
    hot = 0.0
    for k in nodes:
        if o.callback_node(k, poses[k], scans.get(k, empty)):
            hot += o.erasor_seconds()
    cpu_dt = time.perf_counter() - t0
    ref_map, _ = o.cloud(o.MAP_ARRANGED)
    same = bool(first_map.shape == ref_map.shape and np.array_equal(first_map.view(np.uint32), ref_map.view(np.uint32)))
    from erasor_b200 import evaluate
    pr = evaluate.evaluate(map_world, first_map)
    return {"nodes": len(nodes), "processed_scans": len(processed), "scans_per_s": len(processed) / best[0], "ms_per_scan": 1000 * best[0] / len(processed),
            "h2d_bytes_per_scan": int(np.mean([16 * len(v) for v in scans.values()])), "gpu_launches": int(best[1]), "final_map_points": int(best[2]),
            "cpu_oracle_scans_per_s": len(processed) / cpu_dt, "cpu_oracle_hot_path_share": hot / cpu_dt,
            "final_map_bit_identical_to_oracle": same,
            "quality_vs_initial_map": {"PR": round(pr["PR"], 3), "RR": round(pr["RR"], 3), "F1": round(pr["F1"], 4),
                                        "note": "erasor_b200/evaluate.py == reference scripts/analysis_runner.py metric; GT = labelled initial map (synthetic twin)"},
            "pass_ms": pass_ms, "cpu_oracle_pass_ms": round(1000 * cpu_dt, 1),
            "timing": "wall clock around the synchronous C-ABI calls, best of 5 passes (the first includes module load and buffer allocation)"}


def clocks_sampler_start(gpu_index: int):
    q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    try:
        return subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "10", "-i", str(gpu_index)],
                                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return None


def clocks_sampler_stop(proc):
    if proc is None:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    proc.terminate()
    try:
        out, _ = proc.communicate(timeout=5)
    except Exception:
        proc.kill()
        out = ""
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for line in out.strip().splitlines():
        f = [x.strip() for x in line.split(",")]
        if len(f) < 7:
            continue
        try:
            sm.append(float(f[0])); mx.append(float(f[1]))
        except ValueError:
            continue
        for nme, v in zip(names, f[3:7]):
            if v.lower().startswith("active"):
                reasons.add(nme)
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
            "samples": len(sm), "reasons": sorted(reasons)}


def quality(keep, maps):
    """Per-point Preservation / Rejection rate of the frame-independent estimate on this rank's frames
    (labels ride in intensity: 252-259 = moving classes).  Not the NN-matched PR/RR of scripts/analysis_runner.py."""
    lab = np.concatenate([m[:, 3] for m in maps])
    dyn = (lab >= 252) & (lab <= 259)
    k = keep.astype(bool)
    pr = 100.0 * np.count_nonzero(k & ~dyn) / max(1, np.count_nonzero(~dyn))
    rr = 100.0 * np.count_nonzero(~k & dyn) / max(1, np.count_nonzero(dyn))
    return {"PR_pointwise": round(pr, 3), "RR_pointwise": round(rr, 3), "note": "frame-independent mode, per-frame VoI points"}


# ------------------------------------------------------------------------------------------------
# reference arm: the restated reference path (oracle port) on all host cores
# ------------------------------------------------------------------------------------------------
_W = {}


def _ref_init(pdict, maps, qs):
    from erasor_b200 import params
    from oracle import oracle_py
    _W["p"] = params.ErasorParams(**pdict)
    _W["o"] = oracle_py.Oracle(_W["p"])
    _W["maps"], _W["qs"] = maps, qs


def _ref_frame(i):
    return _W["o"].run(_W["maps"][i], _W["qs"][i])


def usable_cores():
    """CPUs this process can actually use: the affinity mask capped by the cgroup CPU quota (a container that sees 128 CPUs
    with cpu.max = 16 CPUs is throttled to 16; more worker processes than that only add contention)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    note = f"{n} CPUs in the affinity mask"
    quota = None
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:                       # cgroup v2: "<quota|max> <period>"
            q, per = f.read().split()
            if q != "max":
                quota = float(q) / float(per)
    except (OSError, ValueError):
        try:
            with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f, open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as g:   # cgroup v1
                q, per = float(f.read()), float(g.read())
                if q > 0:
                    quota = q / per
        except (OSError, ValueError):
            pass
    if quota is not None and quota < n:
        n = max(1, int(quota))
        note += f", cgroup quota {quota:g} CPUs"
    return n, note


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import dataclasses
    import multiprocessing as mp
    from oracle import oracle_py
    oracle_py.build()
    p, map_world, maps, qs, _ = load_workload(0, 1, FRAMES_PER_PASS)
    cores, cores_note = usable_cores()
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_ref_init, initargs=(dataclasses.asdict(p), maps, qs)) as pool:
        # a step = the 20-frame pass repeated until every worker has ~4 frames (balanced waves): the reference is
        # single-threaded, so "all the host threads it can use" means independent frames in parallel processes
        reps = max(1, -(-4 * cores // len(maps)))
        idx = list(range(len(maps))) * reps
        for _ in range(args.warmup):
            pool.map(_ref_frame, idx, chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_ref_frame, idx, chunksize=1)
        dt = time.perf_counter() - t0
    sps = len(idx) * args.steps / dt
    line = {
        "impl": "reference", "metric": "LiDAR scans/sec through R-POD+SRT+R-GPF on KITTI-05 (synthetic twin)",
        "value": sps, "unit": "scans/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 points, f64 index/SRT arithmetic", "data": "synthetic",
        "config": workload_config(p, maps, qs, 1),
        "cpu_baseline": {"value": sps, "unit": "scans/s", "cores": cores, "kind": "port",
                         "sample": f"{len(idx)} frames per step x {args.steps} steps, one oracle process per usable core ({cores_note}) "
                                   "(the reference itself is single-threaded and cannot be compiled here: needs ROS/PCL/Eigen)"},
        "e2e": {"value": sps, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(p, maps, qs, world):
    return {"workload": "KITTI seq 05 (2350-2670) synthetic twin, config/seq_05.yaml, frame-independent pass",
            "frames_per_step_per_gpu": len(maps), "frames_per_step": len(maps) * world,
            "rings_x_sectors": f"{p.num_rings}x{p.num_sectors}", "max_range_m": p.max_range, "version": p.version,
            "mean_map_voi_points": int(np.mean([len(m) for m in maps])), "mean_query_points": int(np.mean([len(q) for q in qs])),
            "in_bin_voxelize": "n/a in mask mode (v3 voxelisation only changes the cloud outputs)",
            "l2": f"{N_INPUT_COPIES} rotating input copies, working set > 126 MB L2", "parallelism": f"frames sharded x{world}"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from erasor_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: erasor_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    p, map_world, maps, qs, idxs = load_workload(rank, world, args.frames)
    F = len(maps)
    mo = np.cumsum([0] + [len(m) for m in maps]).astype(np.uint64)
    qo = np.cumsum([0] + [len(q) for q in qs]).astype(np.uint64)
    M = np.ascontiguousarray(np.concatenate(maps), dtype=np.float32)
    Q = np.ascontiguousarray(np.concatenate(qs), dtype=np.float32)
    NM, NQ = len(M), len(Q)

    h = capi.Handle(p, device=local)
    xs = torch.cuda.ExternalStream(h.stream, device=dev)

    # resident inputs (N_INPUT_COPIES rotating copies) and pinned host inputs
    dM = [torch.from_numpy(M).to(dev) for _ in range(N_INPUT_COPIES)]
    dQ = [torch.from_numpy(Q).to(dev) for _ in range(N_INPUT_COPIES)]
    dK = torch.empty(NM, dtype=torch.uint8, device=dev)
    hM = torch.from_numpy(M).pin_memory()
    hQ = torch.from_numpy(Q).pin_memory()
    hK = torch.empty(NM, dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()

    # The path's one exchange (north_star): every rank folds its frames' keep-masks onto the global map
    # (a point survives if no frame rejected it) and the per-rank masks are all-gathered over NVLink.
    NG = len(map_world)
    gidx = torch.from_numpy(np.concatenate(idxs).astype(np.uint32).view(np.int32)).to(dev)      # uint32 indices for the fold kernel
    final_keep = [None]
    keep_g = torch.ones(NG, dtype=torch.uint8, device=dev)
    gather_buf = torch.empty((world, NG), dtype=torch.uint8, device=dev) if world > 1 else None

    from erasor_b200 import dist as edist

    # Every step folds its frames' masks onto keep_g (library kernel in the step's own submission, no communication).  The job's single collective -- the
    # all-gather of the folded masks -- runs once after the last step of a timed block, inside the timed region.  Steps never
    # contain a collective, so ranks may run different numbers of untimed steps (the clock-sampling warm loop on rank 0).
    def fold(keep_dev):
        h.fold_keep_masks(keep_dev.data_ptr(), gidx.data_ptr(), NM, keep_g.data_ptr(), NG)      # on the handle's stream

    def exchange():
        with torch.cuda.stream(xs):
            final_keep[0] = edist.allgather_and(keep_g, gather_buf)        # the single NCCL collective (no-op at N=1)

    fold_args = (gidx.data_ptr(), keep_g.data_ptr(), NG) if world > 1 else None     # N > 1: erasor_process_frames_fold (one submission)

    def step_resident(i):
        c = i % N_INPUT_COPIES
        h.process_frames_ptr(dM[c].data_ptr(), mo, dQ[c].data_ptr(), qo, dK.data_ptr(), capi.PTR_DEVICE, fold_args)

    def step_host(i):
        h.process_frames_ptr(hM.data_ptr(), mo, hQ.data_ptr(), qo, hK.data_ptr(), capi.PTR_HOST, fold_args)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = h.kernel_launch_count()
        e0.record(xs)
        for i in range(steps):
            fn(warmup + i)
        if world > 1:
            exchange()                        # the job's one collective, inside the timed region
        e1.record(xs)
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), h.kernel_launch_count() - l0

    sampler = clocks_sampler_start(local) if rank == 0 else None
    if sampler is not None:
        t_busy = time.perf_counter()
        i_busy = 0
        while time.perf_counter() - t_busy < 0.3:        # untimed: gives nvidia-smi (10 ms period) samples under this load
            step_resident(i_busy)
            i_busy += 1
    # --- value: resident inputs; K1 timed per launch with CUDA events on the library's stream ---
    h.reset_kernel_times(True)
    ms_res, launches = timed(step_resident, args.steps, args.warmup)
    k1_ms, k1_n = h.kernel_time_ms(1)
    kernel_ms = {name: h.kernel_time_ms(i)[0] / max(1, h.kernel_time_ms(i)[1]) for i, name in ((1, 'k1_rpod_bin'), (2, 'k2_scatter'), (3, 'k3_srt'), (4, 'k4_rgpf_all_classes'))}
    # (the K1 events bracket warm-up launches too; they are the same work, so the per-launch mean is unaffected)
    h.reset_kernel_times(False)
    ms_res_plain, launches = timed(step_resident, args.steps, max(args.warmup, 3))
    # --- e2e: host buffers through the same call ---
    ms_e2e, _ = timed(step_host, args.steps, max(args.warmup, 3))
    clocks = clocks_sampler_stop(sampler) if rank == 0 else None

    step_resident(0)
    fold(dK)
    exchange()
    torch.cuda.synchronize()
    n_static_map = int(final_keep[0].sum().item())

    keep = dK.cpu().numpy()
    if rank == 0:
        from oracle import oracle_py
        oracle_py.build()
        # cpu_baseline: the oracle port on ONE core over a bounded sample of this workload
        o = oracle_py.Oracle(p)
        o.run(maps[0], qs[0])
        t0 = time.perf_counter()
        reps, nfr = 0, 0
        while True:
            for f in range(F):
                o.run(maps[f], qs[f])
                nfr += 1
            reps += 1
            if time.perf_counter() - t0 > 10.0 or reps >= 20:
                break
        cpu_dt = time.perf_counter() - t0
        # parity spot check of the benchmarked output (frame 0) against the oracle
        o.run(maps[0], qs[0])
        _, rej = o.cloud(o.MAP_REJECTED)
        ok0 = np.ones(len(maps[0]), dtype=np.uint8)
        ok0[rej] = 0
        parity_ok = bool(np.array_equal(keep[:len(maps[0])], ok0))

        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        k1_bytes = 16.0 * (NM + NQ)
        k1_avg_ms = k1_ms / max(1, k1_n)
        achieved = k1_bytes / (k1_avg_ms * 1e-3) / 1e9 if k1_avg_ms > 0 else 0.0
        scans = F * world
        value = scans * args.steps / (ms_res_plain * 1e-3)
        # whole-step view: bytes_frame = 16 (N_m + N_q) + N_m + 16 N_F, N_F = points of the bins R-GPF ran on
        npts_flagged, _ = h.rgpf_profile()
        n_f = int(npts_flagged.sum())
        step_bytes = 16.0 * (NM + NQ) + NM + 16.0 * n_f
        step_ms = ms_res_plain / args.steps
        step_gbs = step_bytes / (step_ms * 1e-3) / 1e9
        ev_step = ms_res / args.steps
        line = {
            "metric": "LiDAR scans/sec through R-POD+SRT+R-GPF on KITTI-05 (synthetic twin)",
            "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_res_plain / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 points, f64 index/SRT arithmetic", "data": "synthetic",
            "config": workload_config(p, maps, qs, world),
            "e2e": {"value": scans * args.steps / (ms_e2e * 1e-3), "unit": "scans/s",
                    "h2d_bytes_per_step": int(16 * (NM + NQ)), "d2h_bytes_per_step": int(NM),
                    "note": "pinned host clouds -> erasor_process_frames(PTR_HOST) -> pinned host keep mask"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k1_rpod_bin", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": K1_DRAM_TRAFFIC_BYTES if (world == 1 and args.frames == FRAMES_PER_PASS) else None,
                         "traffic_source": "profiles/r01/ncu_step_full_raw.csv: dram__bytes_read.sum + dram__bytes_write.sum of one k1_rpod_bin launch on this workload",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": k1_bytes, "avg_launch_ms": k1_avg_ms, "launches_timed": int(k1_n),
                         "ms_per_step_with_event_timing": ms_res / args.steps,
                         "avg_ms_per_launch_by_cuda_events": {k: round(v, 5) for k, v in kernel_ms.items()},
                         "share_of_step": round((k1_ms / max(1, k1_n)) / (ms_res / args.steps), 3)},
            "pipeline": {"bytes_per_step": step_bytes, "flagged_bin_points_per_step": n_f, "flagged_bins_per_step": int(len(npts_flagged)),
                         "achieved": step_gbs, "unit": "GB/s", "frac_of_hbm_peak": step_gbs / peak,
                         "kernel_share_of_step": {k: round(v / ev_step, 3) for k, v in kernel_ms.items()},
                         "note": "the largest share is k4_rgpf, which is bound by serial latency, not by HBM: the reference's exact-order "
                                 "float accumulation and Jacobi SVD run on one lane per bin (DESIGN.md section 5); all flagged bins of the "
                                 "step are resident at once, so its time is the slowest bin's chain"},
            "cpu_baseline": {"value": nfr / cpu_dt, "unit": "scans/s", "cores": 1, "kind": "port",
                             "sample": f"{nfr} frames ({reps} passes over this rank's {F} frames), oracle -O2, one core; "
                                       "reference cannot be compiled here (ROS/PCL/Eigen absent)"},
            "clocks": clocks,
            "parity_spot_check": parity_ok,
            "quality": quality(keep, maps),
            "static_map_points": {"kept": n_static_map, "of": NG, "collective": "one all_gather of the folded keep-masks after the K steps, inside the timed region" if world > 1 else "none (1 GPU)"},
        }
        if world == 1 and not args.no_offline_pass:
            try:
                line["offline_pass"] = offline_pass_block(p, map_world, local)
            except Exception as e:      # the headline numbers must survive a failure of the informational block
                line["offline_pass"] = {"error": repr(e)}
        print(json.dumps(line))
    h.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_PASS, help="frames per step per GPU")
    ap.add_argument("--no-offline-pass", action="store_true", help="skip the informational sequential-pass block")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
