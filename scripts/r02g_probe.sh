#!/bin/bash
# GPU call: two-sweep reference patch precompute -- parity (alignment tests) and full-batch / small-batch timing.
set -u
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_sia_gpu.py tests/test_properties_gpu.py tests/test_ref_gpu.py tests/test_golden_gpu.py tests/test_camera_models.py -m gpu -x -q) > gpurun_out/r02g_gputests.log 2>&1
grep -E "passed|failed" gpurun_out/r02g_gputests.log | tail -2
PROBE_REPS=20 timeout 200 python scripts/probe_geom.py 3552 1:2 1:1 > gpurun_out/r02g_full.log 2>&1
timeout 200 python scripts/probe_small_b.py 1 32 148 296 > gpurun_out/r02g_small.log 2>&1
tail -n 2 gpurun_out/r02g_full.log gpurun_out/r02g_small.log
