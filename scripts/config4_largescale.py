"""BASELINE config 4: config/large_scale_05.yaml geometry (20 x 108 @ 80 m), ~50 M-point map VoI per frame, N_q = one scan.
The VoI is the synthetic twin's 80 m crop jitter-replicated to 50 M points on the device (seed 5), as SURVEY 8d-4 prescribes.
Prints one JSON line: per-kernel CUDA-event times, algorithmic bytes (16 (N_m + N_q) + N_m + 16 N_F) and the
achieved fraction of the measured HBM peak.  Run under `ncu --set full -k regex:k1_rpod_bin ...` for dram__bytes.
usage: config4_largescale.py [n_points] [steps]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from erasor_b200 import capi, params, synth  # noqa: E402

N_TARGET = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 5


def main():
    p = params.preset("large_scale_05").replace(skip_voxelize=1)
    w = synth.make_frames(seed=5, n_frames=6, preset_max_range=80.0, n_map_nodes=41, n_beams=32, n_az=900)
    base, q, _, _ = w["frames"][2]
    base = base[(base[:, 0].astype(np.float64) ** 2 + base[:, 1].astype(np.float64) ** 2) < 82.0 ** 2]
    reps = int(np.ceil(N_TARGET / len(base)))
    g = torch.Generator(device="cuda").manual_seed(5)
    M = torch.from_numpy(base).cuda().repeat(reps, 1)
    M[:, :3] += (torch.rand((M.shape[0], 3), device="cuda", generator=g) - 0.5) * torch.tensor([0.19, 0.19, 0.02], device="cuda")
    N = M.shape[0]
    Q = torch.from_numpy(q).cuda()
    keep = torch.empty(N, dtype=torch.uint8, device="cuda")
    mo, qo = np.array([0, N], dtype=np.uint64), np.array([0, len(q)], dtype=np.uint64)
    h = capi.Handle(p)
    xs = torch.cuda.ExternalStream(h.stream)
    for _ in range(2):
        h.process_frames_ptr(M.data_ptr(), mo, Q.data_ptr(), qo, keep.data_ptr(), capi.PTR_DEVICE)
    h.reset_kernel_times(True)
    for _ in range(STEPS):
        h.process_frames_ptr(M.data_ptr(), mo, Q.data_ptr(), qo, keep.data_ptr(), capi.PTR_DEVICE)
    kt = {name: h.kernel_time_ms(i) for i, name in ((1, "k1_rpod_bin"), (2, "k2_srt_scatter"), (3, "k3_srt"), (4, "k4_rgpf_all_classes"))}
    h.reset_kernel_times(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(xs)
    for _ in range(STEPS):
        h.process_frames_ptr(M.data_ptr(), mo, Q.data_ptr(), qo, keep.data_ptr(), capi.PTR_DEVICE)
    e1.record(xs)
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / STEPS
    nf, nr = h.frame_stats()
    npts, _ = h.rgpf_profile()
    n_f = int(npts.sum())
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    kms = {k: (v[0] / max(1, v[1])) for k, v in kt.items()}
    kbytes = {"k1_rpod_bin": 16.0 * (N + len(q)) + 2.0 * N, "k2_srt_scatter": 2.0 * 2 * N + 36.0 * n_f, "k3_srt": 0.0, "k4_rgpf_all_classes": 20.0 * n_f + float(nr.sum())}
    step_bytes = 16.0 * (N + len(q)) + N + 16.0 * n_f
    line = {"config": "4: large_scale_05.yaml geometry, dense map VoI", "rings_x_sectors": f"{p.num_rings}x{p.num_sectors}", "n_map_voi": int(N), "n_query": int(len(q)),
            "flagged_bins": int(nf[0]), "flagged_bin_points": n_f, "rejected_points": int(nr[0]), "largest_flagged_bin": int(npts.max()) if len(npts) else 0,
            "ms_per_frame": step_ms, "frames_per_s": 1000.0 / step_ms,
            "algorithmic_bytes_per_frame": step_bytes, "achieved_gbs": step_bytes / (step_ms * 1e-3) / 1e9, "frac_of_hbm_peak": step_bytes / (step_ms * 1e-3) / 1e9 / peak,
            "hbm_peak_gbs": peak,
            "by_kernel": {k: {"avg_launch_ms": round(kms[k], 4), "algorithmic_bytes": kbytes[k],
                              "achieved_gbs": round(kbytes[k] / (kms[k] * 1e-3) / 1e9, 1) if kms[k] > 0 else 0.0,
                              "frac": round(kbytes[k] / (kms[k] * 1e-3) / 1e9 / peak, 4) if kms[k] > 0 else 0.0} for k in kms}}
    print(json.dumps(line))
    h.close()


if __name__ == "__main__":
    main()
