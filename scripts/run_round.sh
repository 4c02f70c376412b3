set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python scripts/k4_profile.py > gpurun_out/k4_profile.txt 2>&1; tail -6 gpurun_out/k4_profile.txt
timeout 600 python bench.py --steps 30 --warmup 5 --no-offline-pass > gpurun_out/bench_round.json 2> gpurun_out/bench_round.err; tail -c 600 gpurun_out/bench_round.err
ERASOR_B200_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -s 30 -c 40 --csv --log-file gpurun_out/launches_step.csv python scripts/step_only.py 8 > gpurun_out/launches_step.log 2>&1
tail -2 gpurun_out/launches_step.log
