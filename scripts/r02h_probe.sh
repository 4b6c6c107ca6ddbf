#!/bin/bash
# GPU call: ncu captures of the alignment kernel after the round-2b changes (throughput geometry; upfront cluster geometry).
set -u
mkdir -p gpurun_out
export PROBE_REPS=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sia_kernel -c 1 -o gpurun_out/r02h_sia_f2 -f python scripts/probe_geom.py 3552 1:2 > gpurun_out/r02h_ncu_f2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sia_kernel --launch-skip 40 -c 1 -o gpurun_out/r02h_sia_b32 -f python scripts/probe_small_b.py 32 > gpurun_out/r02h_ncu_b32.log 2>&1
ls -la gpurun_out/*.ncu-rep
