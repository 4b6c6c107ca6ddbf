import sys, numpy as np
sys.path.insert(0, '.')
from rpg_svo_b200 import synth, capi
from oracle import binding as ob
ctx = capi.Context(0)
d = synth.make_two_view(3, width=752, height=480, n_levels=4)
pyr = d["ref_pyr"]; fr = ctx.frame(pyr)
g = ctx.fast_detect(fr, 30, 3, 20.0); o = ob.fast_detect(pyr, 3, 30, 20.0, cap=8192)
print(g["n"], len(o["x"]))
go = {(x // 30, y // 30): (x, y, l, s) for x, y, l, s in zip(g["x"], g["y"], g["level"], g["score"])}
oo = {(x // 30, y // 30): (x, y, l, s) for x, y, l, s in zip(o["x"], o["y"], o["level"], o["score"])}
for k in sorted(set(go) | set(oo)):
    if go.get(k) != oo.get(k): print(k, "gpu", go.get(k), "orc", oo.get(k))
for L in range(3):
    gl = ctx.fast_detect(fr, 30, L + 1, 20.0); ol = ob.fast_detect(pyr, L + 1, 30, 20.0, cap=8192)
    print("levels<=", L, "equal", all(np.array_equal(gl[k], ol[k]) for k in ("x", "y", "level")))
for L in range(3):
    print(L, np.array_equal(fr.download_level(L), pyr[L]))
