"""Debug tool: per-phase clock64 breakdown of lh_grid4_kernel (library built with -DUNC_PROFILE)."""
import sys, os, ctypes as C, shutil, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
shutil.copy(os.path.join(ROOT, 'tools', 'libprof.so'), os.path.join(ROOT, 'unconfined_b200', 'libunconfined_b200.so'))
import bench, unconfined_b200 as ub
d, t, r, z = bench.c5a_grid(0, nt=2)
p, tD, sv, rD, zD, lay = bench.derive(d, t, r, z, ub)
prm = ub.Params(p)
ub.eval_grid(prm, tD, sv, rD, zD, lay)
out = (C.c_ulonglong * 16)()
ub.lib().unc_debug_profile(None, 1)
t0 = time.perf_counter()
ub.eval_grid(prm, tD, sv, rD, zD, lay)
dt = time.perf_counter() - t0
ub.lib().unc_debug_profile(out, 0)
v = np.array(out[:11], float)
names = ['round barrier', 'barrier after tables', 'ap_terms(stage)', 'hot loop', 'wynn/phaseB', 'pool exit', 'de Hoog',
         'item fetch + barrier', 'z loads', 'item_tables', 'uniformity check (warp 0)']
print('wall %.1f ms for %d points -> %.3g points/s' % (dt * 1e3, len(tD) * len(rD) * len(zD), len(tD) * len(rD) * len(zD) / dt))
for n, x in zip(names, v):
    print('%-28s %6.2f%%' % (n, 100 * x / v.sum()))
