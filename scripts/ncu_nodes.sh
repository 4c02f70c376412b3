#!/bin/bash
# scripts/ncu_nodes.sh [config] [frames] -- launch list + one `ncu --set full` capture of each kernel of the node-mode step (gpurun, 1 GPU)
CFG=${1:-seq05}; FR=${2:-20}
mkdir -p gpurun_out
python scripts/step_nodes.py 1 $CFG $FR > gpurun_out/step_nodes_warm.log 2>&1     # builds the workload cache outside ncu
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -s 20 -c 40 --csv --log-file gpurun_out/launches_nodes_$CFG.csv \
    python scripts/step_nodes.py 6 $CFG $FR > gpurun_out/ncu_launches_$CFG.log 2>&1
for k in k1_rpod_bin k2_srt_scatter "k4_rgpf<.int.256, .int.32>" "k4_rgpf<.int.128, .int.128>"; do
  name=$(echo "$k" | tr -d '<> ,.' )
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$k" -s 3 -c 1 -f -o gpurun_out/prof_${CFG}_$name \
      python scripts/step_nodes.py 5 $CFG $FR > gpurun_out/ncu_${CFG}_$name.log 2>&1
  tail -1 gpurun_out/ncu_${CFG}_$name.log
done
ls -la gpurun_out | head -40
