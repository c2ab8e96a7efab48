import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fibergen_b200 as fb
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for n, mode, scheme in [((512, 128, 128), "elasticity", "staggered"), ((128, 512, 128), "elasticity", "staggered"),
                        ((128, 128, 512), "elasticity", "staggered"), ((512, 128, 128), "elasticity", "collocated"),
                        ((1024, 64, 128), "elasticity", "staggered"), ((64, 1024, 128), "elasticity", "staggered"),
                        ((128, 64, 1024), "elasticity", "staggered")]:
    ctx = fb.Context(*n, 1.0, 1.0, 1.0, mode=mode, gamma_scheme=scheme)
    d = 6
    rng = np.random.default_rng(1)
    f = ctx.field(rng.standard_normal((d,) + n))
    E = np.zeros(d)
    for it in range(8):
        if it == 3:
            ctx.profile(True)
        ctx.gamma(f, E, 1.3, 0.4, -1.0, 0.0)
    res = ctx.profile_results()
    print(tag, n, scheme, " ".join("%s=%.3f" % (k, v[0] / max(v[1], 1)) for k, v in sorted(res.items()) if k.startswith("fft")), flush=True)
    ctx.close()
