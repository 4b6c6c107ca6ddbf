#!/bin/bash
# GPU call: parity of the upfront / plain-pinhole / pose-from-shared changes, then small-batch latency A/B and section timers.
set -u
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/r02d_gputests.log 2>&1
tail -4 gpurun_out/r02d_gputests.log
timeout 200 python scripts/probe_small_b.py 1 8 32 > gpurun_out/r02d_small_upfront.log 2>&1
SVO_B200_SIA_UPFRONT=0 timeout 200 python scripts/probe_small_b.py 1 8 32 > gpurun_out/r02d_small_perlevel.log 2>&1
PROBE_REPS=10 timeout 200 python scripts/probe_geom.py 3552 1:2 1:1 > gpurun_out/r02d_full_plain.log 2>&1
SVO_B200_SIA_PLAIN=0 PROBE_REPS=10 timeout 200 python scripts/probe_geom.py 3552 1:2 > gpurun_out/r02d_full_general.log 2>&1
SVO_B200_LIB=build/libsvo_b200_dbg.so SVO_B200_SIA_DEBUG=1 timeout 200 python scripts/probe_small_b.py 1 2>&1 | grep "sia dbg" | tail -n 6 > gpurun_out/r02d_dbg_upfront.log
SVO_B200_SIA_UPFRONT=0 SVO_B200_LIB=build/libsvo_b200_dbg.so SVO_B200_SIA_DEBUG=1 timeout 200 python scripts/probe_small_b.py 1 2>&1 | grep "sia dbg" | tail -n 6 > gpurun_out/r02d_dbg_perlevel.log
SVO_B200_LIB=build/libsvo_b200_dbg.so timeout 300 python scripts/side_kernels.py --quick 2>&1 | grep -E "po dbg|\{" | tail -n 8 > gpurun_out/r02d_dbg_side.log
tail -n 2 gpurun_out/r02d_small_*.log gpurun_out/r02d_full_*.log
cat gpurun_out/r02d_dbg_*.log | cut -c1-300
