import sys, time, numpy as np
sys.path.insert(0, '.')
from rpg_svo_b200 import synth, capi
from oracle import binding as ob
ctx = capi.Context(0)
d = synth.make_frame_pair(1000)
ref = ctx.frame(d['ref_pyr']); cur = ctx.frame(d['cur_pyr'])
# pyramid check
f0 = ctx.frame_from_level0(d['ref_pyr'][0], 5)
for l in range(5):
    print('pyr level', l, np.array_equal(f0.download_level(l), d['ref_pyr'][l]))
T0 = synth.se3_identity()
for lv in range(5):
    g = ctx.sparse_residuals(ref, cur, d['cam'], lv, T0, d['px'], d['f'], d['pos'], d['has_point'], d['ref_pos'])
    o = ob.sparse_residuals(d['ref_pyr'][lv], d['cur_pyr'][lv], lv, d['cam'], T0, d['px'], d['f'], d['pos'], d['has_point'], d['ref_pos'])
    v = o['visible'].astype(bool); m = o['in_image'].astype(bool)
    print('lvl', lv, 'vis eq', np.array_equal(g['visible'], o['visible']), 'in eq', np.array_equal(g['in_image'], o['in_image']),
          'refpatch maxdiff', np.abs(g['ref_patch'][v]-o['ref_patch'][v]).max(), 'res maxdiff', np.nanmax(np.abs(g['residuals'][m]-o['residuals'][m])),
          'chi2', g['chi2'], o['chi2'], 'nmeas', g['n_meas'], o['n_meas'])
    print('   H relerr', np.abs(g['H']-o['H']).max()/np.abs(o['H']).max(), 'Jres', np.abs(g['Jres']-o['Jres']).max()/np.abs(o['Jres']).max())
for (mx, mn) in [(4,0),(4,2),(2,0)]:
    t=time.time()
    g = ctx.sparse_img_align(ref, cur, d['cam'], T0, d['px'], d['f'], d['pos'], d['has_point'], d['ref_pos'], mx, mn, want_trace=True)
    tg = time.time()-t
    o = ob.sparse_img_align(d['ref_pyr'], d['cur_pyr'], d['cam'], T0, d['px'], d['f'], d['pos'], d['has_point'], d['ref_pos'], mx, mn)
    print(mx, mn, 'gpu ms %.2f'%(tg*1e3), 'ntr', g['n_tracked'], o['n_tracked'], 'iters', len(g['trace']), len(o['trace']), 'err vs oracle', synth.pose_error(g['T'], o['T']), 'vs gt', synth.pose_error(g['T'], d['T_cur_ref_gt']), g['stats'])
    for a,b in list(zip(g['trace'], o['trace']))[:6]:
        print('   ', a['level'], a['iter'], a['accepted'], b['accepted'], a['n_meas'], b['n_meas'], '%.6f %.6f'%(a['chi2'], b['chi2']), np.abs(a['x']-b['x']).max())
