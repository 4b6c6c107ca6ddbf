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
  # the alignment kernel: the throughput geometry (3552 pairs), then the small-batch cluster geometry (32 pairs)
  PROBE_REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:sia_kernel -c 1 -o gpurun_out/${tag}_sia_f2 -f \
    python scripts/probe_geom.py 3552 1:2 > gpurun_out/${tag}_ncu_f2.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:sia_kernel --launch-skip 40 -c 1 -o gpurun_out/${tag}_sia_b32 -f \
    python scripts/probe_small_b.py 32 > gpurun_out/${tag}_ncu_b32.log 2>&1
  # launch list of one short bench run, restricted to the library's kernels (-k matches the bare function name; the torch kernels
  # that manufacture the synthetic stream would otherwise fill the capture): per-launch durations, cold and serialised, for the
  # kernels' SHARE of the legs.  Summaries: scripts/ncu_lines.py / scripts/ncu_full.py (stall taxonomy, phase buckets).
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'sia_kernel|pyramid_|half_sample' -c 1500 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/${tag}_launches_bench.log 2>&1
fi
ls -la gpurun_out/ | tail -8; du -sh gpurun_out
