#!/bin/bash
# scripts/ncu_full.sh -- one `ncu --set full` capture of each kernel of the path (run under gpurun, 1 GPU)
mkdir -p gpurun_out
for k in k1_rpod_bin k2_scatter k3_srt k4_rgpf; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_$k \
      python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_$k.log 2>&1
  tail -2 gpurun_out/ncu_$k.log
done
ls -la gpurun_out
