import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fibergen_b200 as fb
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for n, mode, d, scheme in [((256, 256, 256), "elasticity", 6, "staggered"), ((256, 256, 256), "elasticity", 6, "collocated"),
                           ((256, 256, 128), "hyperelasticity", 9, "collocated"), ((128, 128, 128), "elasticity", 6, "collocated"),
                           ((64, 64, 64), "elasticity", 6, "collocated"), ((256, 256, 256), "heat", 3, "staggered")]:
    ctx = fb.Context(*n, 1.0, 1.0, 1.0, mode=mode, gamma_scheme=scheme)
    rng = np.random.default_rng(1)
    f = ctx.field(rng.standard_normal((d,) + n))
    E = np.zeros(d)
    for it in range(8):
        if it == 3:
            ctx.profile(True)
        ctx.gamma(f, E, 1.3, 0.4, -1.0, 0.0)
    res = ctx.profile_results()
    print(tag, n, mode, scheme, " ".join("%s=%.3f" % (k, v[0] / max(v[1], 1)) for k, v in sorted(res.items()) if k.startswith("fft")), flush=True)
    ctx.close()
