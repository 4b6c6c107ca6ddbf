import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rpg_svo_b200 import capi, synth
from tests.test_sia_gpu import _border_case
from oracle import binding as ob
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
d = _border_case(5, n)
ctx = capi.Context(0)
ref, cur = ctx.frame(d["ref_pyr"]), ctx.frame(d["cur_pyr"])
T0 = synth.se3_identity()
mx, mn = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (4, 0)
print("launching", n, mx, mn, flush=True)
g = ctx.sparse_img_align(ref, cur, d["cam"], T0, d["px"], d["f"], d["pos"], d["has_point"], d["ref_pos"], mx, mn, want_trace=True)
o = ob.sparse_img_align(d["ref_pyr"], d["cur_pyr"], d["cam"], T0, d["px"], d["f"], d["pos"], d["has_point"], d["ref_pos"], mx, mn)
print("iters", len(g["trace"]), len(o["trace"]), "err", synth.pose_error(g["T"], o["T"]), flush=True)
