"""Each kernel of the path at its BASELINE config, launched through the C ABI exactly as bench.py's roofline_by_kernel leg does
(for the ncu captures: `bash scripts/profile_kernels.sh <tag>`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import bench
from rpg_svo_b200 import capi

ctx = capi.Context(0)
out = bench.measure_kernels(ctx, capi, 6570.9, quick="--quick" in sys.argv)
print(json.dumps({k: v["kernel_ms"] for k, v in out.items()}))
