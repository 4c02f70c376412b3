set -x
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "process_frames_fold or fold_keep or batch_masks" 2>&1 | tail -3
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?; tail -c 300 gpurun_out/bench_n2.err
