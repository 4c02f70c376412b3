mkdir -p gpurun_out
export ERASOR_B200_NO_GRAPH=1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k2_scatter|k3_srt" -s 6 -c 2 -f -o gpurun_out/prof_k23 \
    python scripts/step_only.py 5 > gpurun_out/ncu_k23.log 2>&1
tail -3 gpurun_out/ncu_k23.log
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k4_rgpf" -s 9 -c 3 -f -o gpurun_out/prof_k4new \
    python scripts/step_only.py 5 > gpurun_out/ncu_k4new.log 2>&1
tail -3 gpurun_out/ncu_k4new.log
ls -la gpurun_out | tail -5
