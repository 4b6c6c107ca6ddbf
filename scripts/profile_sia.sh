#!/bin/bash
# Capture the alignment kernel once with ncu on the GPU box (run through gpurun from the repository root):
#   gpurun --timeout 600 -- 'bash scripts/profile_sia.sh r02a'
# then, back in the build container:
#   python scripts/ncu_lines.py gpurun_out/<tag>_sia.ncu-rep 30 > profiles/<tag>_sia_kernel_ncu_summary.txt
#   python scripts/update_dram_json.py gpurun_out/<tag>_sia.ncu-rep profiles/<tag>_sia_kernel_ncu_summary.txt
# Numbers printed by bench.py under ncu are never bench values.
set -e
tag=${1:-rXX}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:sia_kernel -c 1 -o gpurun_out/${tag}_sia \
    python bench.py --steps 1 --warmup 1 --no-extras --no-cpu > /dev/null 2>&1
ls -la gpurun_out/${tag}_sia.ncu-rep
