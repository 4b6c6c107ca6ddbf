"""Device-time probe of the side kernels through the ABI (run on the GPU box)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rpg_svo_b200 import capi, synth

def timeit(fn, reps=50):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return 1e6 * (time.perf_counter() - t0) / reps

ctx = capi.Context(0)
out = {}
pc = synth.make_pose_opt_case(1005, 1000, 1920, 1080)
pargs = (2.0, 10, pc["cam"].fx, pc["T_init"], pc["f"], pc["pos"], pc["level"], pc["has_point"])
out["pose_opt_C3_e2e_us"] = timeit(lambda: ctx.pose_optimize(*pargs))
B = 64
off = np.arange(B + 1, dtype=np.int32) * 1000
cat = lambda a: np.concatenate([a] * B)
out["pose_opt_C3_batch64_e2e_us_per_frame"] = timeit(lambda: ctx.pose_optimize_batch(2.0, 10, [pc["cam"].fx] * B, np.stack([pc["T_init"]] * B), off, cat(pc["f"]), cat(pc["pos"]), cat(pc["level"]), cat(pc["has_point"])), 10) / B
print(json.dumps(out))
