#!/bin/bash
# scripts/ncu_one.sh <kernel-regex> <out-name> [skip] -- one `ncu --set full` capture of one kernel of bench.py
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$1" -s ${3:-4} -c 1 -f -o gpurun_out/prof_$2 \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_$2.log 2>&1
tail -2 gpurun_out/ncu_$2.log
