set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python scripts/k4_profile.py > gpurun_out/k4_profile.txt 2>&1; tail -12 gpurun_out/k4_profile.txt
timeout 600 python bench.py --steps 30 --warmup 5 --no-offline-pass > gpurun_out/bench_k4q.json 2> gpurun_out/bench_k4q.err; tail -c 2500 gpurun_out/bench_k4q.json
