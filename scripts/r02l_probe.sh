#!/bin/bash
# GPU call: launch list of one short bench run restricted to the library's kernels (per-launch durations under ncu are cold and
# serialised: for the kernels' SHARE of the legs), and the batch size at which the throughput geometry overtakes the 320x1 one.
set -u
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'svo::' -c 1500 --csv --log-file gpurun_out/r02l_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/r02l_launches_bench.log 2>&1
PROBE_REPS=50 timeout 200 python scripts/probe_geom.py 296 1:1 1:2 > gpurun_out/r02l_b296.log 2>&1
PROBE_REPS=50 timeout 200 python scripts/probe_geom.py 444 1:1 1:2 > gpurun_out/r02l_b444.log 2>&1
PROBE_REPS=50 timeout 200 python scripts/probe_geom.py 148 1:1 1:2 4:0 > gpurun_out/r02l_b148.log 2>&1
tail -n 1 gpurun_out/r02l_b*.log; wc -l gpurun_out/r02l_launches.csv
