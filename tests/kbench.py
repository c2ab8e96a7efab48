"""Kernel micro-benchmark for A/B runs of environment-selected kernel variants (not a test): times the kernels of one fused CG
iteration of config 2 with the library's own events.   python tests/kbench.py [n | nx ny nz]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fibergen_b200 as fb
from microstructures import config2_fibres, fiber_list

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dims = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (n, n, n)
Cs, Ds, R, Lc = config2_fibres()
sc = min(dims) / 256.0
fibs, box = fiber_list(dims, Cs * sc, Ds, R * sc, Lc * sc)
s = fb.LSSolver(dims[0], dims[1], dims[2], mode="elasticity", method="cg", error_estimator="residual", tol=1e-300, maxiter=14)
s.add_material("m", "iso", 0.6121, 1.5739)
s.add_material("f", "iso", 30.93, 17.4)
s.init()
s.init_phase(fibs)
s.set_strain([1, 0, 0, 0, 0, 0])
cnt = [0]


def cb():
    cnt[0] += 1
    if cnt[0] == 4:
        s.lib.fgb_profile_enable(s.ctx(), 1)
    return False


s.set_convergence_callback(cb)
s.run()
c = fb.Context.__new__(fb.Context)
c.lib, c.h = s.lib, s.ctx()
prof = fb.Context.profile_results(c)
c.h = None
tot = 0
out = []
for k, (ms, cntk) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    if cntk >= 8:
        out.append("%s %.4f" % (k, ms / cntk))
        tot += ms / cntk
print(os.environ.get("KB_TAG", "default"), "sum %.4f |" % tot, " ".join(out))
