#!/bin/bash
# GPU call: xyz_ref in shared memory + last-warp factorisation (throughput geometry), pose optimizer changes -- parity and timing.
set -u
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/r02f_gputests.log 2>&1
tail -4 gpurun_out/r02f_gputests.log
PROBE_REPS=20 timeout 200 python scripts/probe_geom.py 3552 1:2 1:1 > gpurun_out/r02f_full.log 2>&1
timeout 300 python scripts/side_kernels.py > gpurun_out/r02f_side.log 2>&1
SVO_B200_LIB=build/libsvo_b200_dbg.so timeout 300 python scripts/side_kernels.py --quick 2>&1 | grep -E "po dbg" | tail -n 3 > gpurun_out/r02f_dbg_side.log
SVO_B200_LIB=build/libsvo_b200_dbg.so SVO_B200_SIA_DEBUG=1 PROBE_REPS=1 timeout 200 python scripts/probe_geom.py 3552 1:2 2>&1 | grep "sia dbg" | tail -n 6 > gpurun_out/r02f_dbg_full.log
tail -n 2 gpurun_out/r02f_full.log gpurun_out/r02f_side.log
cat gpurun_out/r02f_dbg_side.log; cut -c1-330 gpurun_out/r02f_dbg_full.log
