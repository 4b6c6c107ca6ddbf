#!/bin/bash
# GPU call: 32-value patch cache at four CTAs per SM (SVO_B200_SIA_BQ=1) -- parity and full-batch timing; filtered launch list.
set -u
mkdir -p gpurun_out
(SVO_B200_SIA_BQ=1 timeout 400 python -m pytest tests/test_sia_gpu.py tests/test_properties_gpu.py tests/test_ref_gpu.py tests/test_camera_models.py -m gpu -x -q) > gpurun_out/r02m_gputests_bq.log 2>&1
grep -E "passed|failed" gpurun_out/r02m_gputests_bq.log | tail -2
PROBE_REPS=20 timeout 200 python scripts/probe_geom.py 3552 1:2 > gpurun_out/r02m_full_48.log 2>&1
SVO_B200_SIA_BQ=1 PROBE_REPS=20 timeout 200 python scripts/probe_geom.py 3552 1:2 > gpurun_out/r02m_full_bq.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'sia_kernel|pyramid_|half_sample' -c 1500 --csv --log-file gpurun_out/r02m_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/r02m_launches_bench.log 2>&1
tail -n 1 gpurun_out/r02m_full_*.log; wc -l gpurun_out/r02m_launches.csv
