set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 30 --warmup 5 --no-offline-pass > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; tail -c 3000 gpurun_out/bench_graph.json
ERASOR_B200_NO_GRAPH=1 python bench.py --steps 30 --warmup 5 --no-offline-pass > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err; tail -c 1200 gpurun_out/bench_nograph.json
