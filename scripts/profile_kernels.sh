#!/bin/bash
# ncu captures of every kernel of the path at its BASELINE config (run through gpurun from the repository root; what a call
# writes under gpurun_out/ must stay below 64 MiB, hence two parts):
#   gpurun --timeout 1200 -- 'bash scripts/profile_kernels.sh r02 side'
#   gpurun --timeout 1200 -- 'bash scripts/profile_kernels.sh r02 sia'
# then here:  python scripts/kernels_dram_json.py gpurun_out/<tag>_kernels.ncu-rep gpurun_out/<tag>_sia.ncu-rep gpurun_out/<tag>_sia_b32.ncu-rep
# Numbers printed under ncu are never bench values.
set -u
tag=${1:-rXX}
part=${2:-side}
mkdir -p gpurun_out
if [ "$part" = side ]; then
  timeout 900 ncu --set full --clock-control none \
    -k regex:'depth_filter_kernel|align_batch_kernel|find_match_direct_kernel|pose_opt_kernel|point_optimize_kernel|reproject_match_kernel|fast_detect_kernel|pyramid_l0_l1_stream_kernel|pyramid_fused_kernel' \
    --launch-count 26 -o gpurun_out/${tag}_kernels -f python scripts/side_kernels.py --quick > gpurun_out/${tag}_kernels.log 2>&1
else
  # the alignment kernel: the full-batch geometry (bench default window), then the small-batch cluster geometry
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sia_kernel -c 1 -o gpurun_out/${tag}_sia -f \
    python bench.py --steps 1 --warmup 1 --no-extras --no-cpu --no-e2e > /dev/null 2>&1
  timeout 300 ncu --set full --clock-control none -k regex:sia_kernel --launch-skip 40 -c 1 -o gpurun_out/${tag}_sia_b32 -f \
    python scripts/probe_small_b.py 32 > /dev/null 2>&1
  # launch list of one short bench run (per-launch durations, cold and serialised: for the kernel's SHARE of the step)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${tag}_launches_bench.log 2>&1
fi
ls -la gpurun_out/ | tail -8; du -sh gpurun_out
