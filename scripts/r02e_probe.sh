#!/bin/bash
# GPU call: st.async exchange of the upfront cluster variant -- parity, latency A/B, section timers.
set -u
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_sia_gpu.py tests/test_properties_gpu.py tests/test_ref_gpu.py tests/test_golden_gpu.py tests/test_camera_models.py -m gpu -x -q) > gpurun_out/r02e_gputests.log 2>&1
tail -4 gpurun_out/r02e_gputests.log
timeout 200 python scripts/probe_small_b.py 1 8 32 > gpurun_out/r02e_small_async.log 2>&1
SVO_B200_SIA_ASYNC=0 timeout 200 python scripts/probe_small_b.py 1 8 32 > gpurun_out/r02e_small_barrier.log 2>&1
SVO_B200_LIB=build/libsvo_b200_dbg.so SVO_B200_SIA_DEBUG=1 timeout 200 python scripts/probe_small_b.py 1 2>&1 | grep "sia dbg" | tail -n 4 > gpurun_out/r02e_dbg_async.log
SVO_B200_SIA_ASYNC=0 SVO_B200_LIB=build/libsvo_b200_dbg.so SVO_B200_SIA_DEBUG=1 timeout 200 python scripts/probe_small_b.py 1 2>&1 | grep "sia dbg" | tail -n 4 > gpurun_out/r02e_dbg_barrier.log
tail -n 2 gpurun_out/r02e_small_*.log
cat gpurun_out/r02e_dbg_*.log | cut -c1-330
