#!/bin/bash
# Two host threads (tracking + mapper), each with its own device context, under compute-sanitizer (run on the GPU box):
#   gpurun --timeout 900 -- 'bash scripts/sanitize_pipeline.sh r02'
# Writes gpurun_out/<tag>_pipeline_{racecheck,memcheck}.log
set -u
tag=${1:-rXX}
mkdir -p gpurun_out
rm -rf /tmp/svo_pipe && mkdir -p /tmp/svo_pipe
python -m pytest tests/test_host_cpp_gpu.py -q -k "pipeline_two_threads and 1" --basetemp=/tmp/svo_pipe > gpurun_out/${tag}_pipeline_pytest.log 2>&1
in=$(find /tmp/svo_pipe -name in.bin | head -1)
echo "input: $in" | tee gpurun_out/${tag}_pipeline_racecheck.log
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all rpg_svo_b200/host/host_pipeline_demo pipeline "$in" /tmp/svo_pipe/out_race.bin >> gpurun_out/${tag}_pipeline_racecheck.log 2>&1
echo "input: $in" > gpurun_out/${tag}_pipeline_memcheck.log
timeout 600 compute-sanitizer --tool memcheck rpg_svo_b200/host/host_pipeline_demo pipeline "$in" /tmp/svo_pipe/out_mem.bin >> gpurun_out/${tag}_pipeline_memcheck.log 2>&1
tail -n 3 gpurun_out/${tag}_pipeline_racecheck.log gpurun_out/${tag}_pipeline_memcheck.log
