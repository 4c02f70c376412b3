#!/bin/bash
# scripts/gpu_round.sh -- one GPU session: parity tests, smoke, bench (both arms), ncu launch list + one full capture of every
# kernel of a step, memcheck of the R-GPF size classes (run under gpurun, 1 GPU)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py 2> gpurun_out/bench.err > gpurun_out/bench.json; tail -c 300 gpurun_out/bench.err; head -c 400 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench.err > gpurun_out/bench_ref.json; head -c 600 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-offline-pass > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
ERASOR_B200_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k1_rpod|k2_scatter|k3_srt|k4_rgpf" -s 12 -c 6 -f -o gpurun_out/prof_step_full \
    python scripts/step_only.py 5 > gpurun_out/ncu_step_full.log 2>&1
tail -2 gpurun_out/ncu_step_full.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "sort_classes and (rough or dup) and (190 or 1200 or 3500)" > gpurun_out/memcheck.log 2>&1; echo memcheck rc=$?; tail -4 gpurun_out/memcheck.log
