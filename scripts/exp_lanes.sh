#!/bin/bash
# quick matrix: lanes x K1/K2 CTAs-per-SM x fused SRT (resident + e2e scans/s), one line each
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-offline-pass --no-sweep > /dev/null 2>&1   # builds the workload cache
for fused in 0 1; do for cps in 4 2 1; do for lanes in 1 2 3; do
  ERASOR_B200_UNFUSED_SRT=$((1-fused)) ERASOR_B200_CTAS_PER_SM=$cps python bench.py --steps 30 --warmup 3 --lanes $lanes --no-offline-pass --no-sweep 2> gpurun_out/exp.err | \
   python -c "import sys,json; d=json.load(sys.stdin); k=d['roofline']['by_kernel']; print('fused=$fused cps=$cps lanes=$lanes value=%.0f one_lane=%.0f e2e=%.0f e2e1=%.0f parity=%s k1=%.1f k2=%.1f k3=%.1f k4=%.1f us' % (d['value'], d['lanes']['value_one_lane'], d['e2e']['value'], d['lanes']['e2e_one_lane'], d['parity_spot_check'], *[1000*k[x]['avg_launch_ms'] for x in ('k1_rpod_bin','k2_scatter','k3_srt','k4_rgpf_all_classes')]))"
done; done; done | tee gpurun_out/exp_lanes.txt
