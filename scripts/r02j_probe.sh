#!/bin/bash
set -u
mkdir -p gpurun_out
export PROBE_REPS=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sia_kernel -c 1 -o gpurun_out/r02j_sia_f2 -f python scripts/probe_geom.py 3552 1:2 > gpurun_out/r02j_ncu_f2.log 2>&1
ls -la gpurun_out/r02j*.ncu-rep
