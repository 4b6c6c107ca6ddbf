#!/bin/bash
# GPU call: what bounds the full-batch alignment kernel?  ncu captures of two geometries, clock64 section timers, A/B of knobs.
set -u
mkdir -p gpurun_out
export PROBE_REPS=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sia_kernel -c 1 -o gpurun_out/r02c_sia_f2 -f python scripts/probe_geom.py 3552 1:2 > gpurun_out/r02c_ncu_f2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sia_kernel -c 1 -o gpurun_out/r02c_sia_f1 -f python scripts/probe_geom.py 3552 1:1 > gpurun_out/r02c_ncu_f1.log 2>&1
SVO_B200_LIB=build/libsvo_b200_dbg.so SVO_B200_SIA_DEBUG=1 timeout 200 python scripts/probe_geom.py 3552 1:2 1:1 > gpurun_out/r02c_dbg.log 2>&1
export PROBE_REPS=10
timeout 200 python scripts/probe_geom.py 3552 1:2 1:1 -1:0 > gpurun_out/r02c_ab_default.log 2>&1
SVO_B200_SIA_WINDOWS=0 timeout 200 python scripts/probe_geom.py 3552 1:1 > gpurun_out/r02c_ab_nowin.log 2>&1
SVO_B200_SIA_FPT2=2 timeout 200 python scripts/probe_geom.py 3552 -1:0 > gpurun_out/r02c_ab_fpt2_2.log 2>&1
tail -n 3 gpurun_out/r02c_ab_*.log
ls -la gpurun_out | tail; du -sh gpurun_out
