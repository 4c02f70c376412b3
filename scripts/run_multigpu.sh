#!/bin/bash
# scripts/run_multigpu.sh N [extra bench args] -- the frame-sharded job on N GPUs of one box (gpurun --gpus N): the C-ABI collective
# test (N >= 2) and bench.py under torch.distributed.run, one rank per GPU.
N=${1:-2}; shift
mkdir -p gpurun_out
nvidia-smi -L | head -$N
if [ "$N" -ge 2 ]; then python -m pytest tests/test_gpu_nodes.py -m gpu -q -k two_gpu > gpurun_out/test_two_gpu.log 2>&1; tail -2 gpurun_out/test_two_gpu.log; fi
python bench.py --steps 3 --warmup 3 --no-offline-pass --no-sweep > /dev/null 2>&1     # warms the workload cache for N = 1 shapes
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 30 --warmup 5 "$@" \
    > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_n$N.json") if l.startswith("{")][-1])
print("N=%d value=%.0f e2e=%.0f ms_per_step=%.4f one_lane=%.0f" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["lanes"]["value_one_lane"]))
print(json.dumps(d["exchange"])[:600])
print(json.dumps(d["quality_final_map"]))
PY
# BASELINE config 5 on the same N GPUs: 40 x 360 bins, 262144-point scans, 2 M-point shared map, 32 nodes per step per rank
if [ "${CONFIG5:-0}" = "1" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --config synthetic40x360 --frames 32 --steps 8 --warmup 3 \
    --no-offline-pass --no-sweep > gpurun_out/config5_n$N.json 2> gpurun_out/config5_n$N.err
echo "config5 rc=$?"; tail -2 gpurun_out/config5_n$N.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/config5_n$N.json") if l.startswith("{")][-1])
print("config5 N=%d value=%.0f e2e=%.0f ms_per_step=%.4f one_lane=%.0f" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["lanes"]["value_one_lane"]))
print(json.dumps(d["exchange"])[:500])
PY
fi
