"""One call of each side-path kernel (for an ncu launch list): reprojector, FAST detector, depth filter, align2D, pose optimizer."""
import sys
sys.path.insert(0, '.')
from rpg_svo_b200 import synth, capi
ctx = capi.Context(0)
m = synth.make_map_case(4001, n_kfs=10, n_points=1200, n_candidates=150)
kfs, cur = [ctx.frame(p) for p in m["kf_pyr"]], ctx.frame(m["cur_pyr"])
for _ in range(3):
    g = ctx.reproject_map(m["view"], kfs, cur, m["cur_T_f_w"], m["cam"], m["options"], m["cell_order"], m["pt_type"], m["pt_n_failed"], m["pt_n_succeeded"])
    d = ctx.fast_detect(cur, 30, 3, 20.0)
print(g["n_matches"], g["n_speculative"], d["n"])
