#!/bin/bash
# GPU call: the round's final record -- every GPU test, both bench arms, the launch list of one short bench run.
set -u
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/r02k_gputests.log 2>&1
grep -E "passed|failed" gpurun_out/r02k_gputests.log | tail -2
timeout 400 python bench.py > gpurun_out/r02k_bench_main.json 2> gpurun_out/r02k_bench_main.err
tail -c 400 gpurun_out/r02k_bench_main.json; echo
timeout 400 python bench.py --impl reference > gpurun_out/r02k_bench_ref.json 2> gpurun_out/r02k_bench_ref.err
tail -c 300 gpurun_out/r02k_bench_ref.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02k_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02k_launches_bench.log 2>&1
ls -la gpurun_out | tail -8
