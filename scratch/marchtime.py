import sys, os, ctypes as C, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import fibergen_b200 as fb
from microstructures import sphere_phi
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for n in [(64, 512, 512), (128, 512, 256)]:
    ctx = fb.Context(*n, mode="elasticity", gamma_scheme="staggered")
    phi = sphere_phi(n, R=0.3, sub=1)
    lam1, mu1 = fb.lame(1.0, 0.3)
    lam2, mu2 = fb.lame(25.0, 0.2)
    ctx.set_phases([1 - phi, phi], [("iso", [mu1, lam1]), ("iso", [mu2, lam2])])
    rng = np.random.default_rng(0)
    fr, fp, fp2, fx = ctx.field(rng.standard_normal((6,) + n)), ctx.field(rng.standard_normal((6,) + n)), ctx.field(), ctx.field(np.zeros((6,) + n))
    pAp, delta = C.c_double(), C.c_double()
    for it in range(7):
        if it == 2:
            ctx.profile(True)
        ctx.chk(ctx.lib.fgb_cg_step(ctx.h, -1, fr, 0.3, fp, fp2, -2, 3.0, 0.0, C.byref(pAp)))
        ctx.chk(ctx.lib.fgb_cg_update(ctx.h, fx, fr, fp2, -2, 1e-3, C.byref(delta)))
        fp, fp2 = fp2, fp
    res = ctx.profile_results()
    print(tag, n, " ".join("%s=%.3f" % (k, v[0] / max(v[1], 1)) for k, v in sorted(res.items())), flush=True)
    ctx.close()
