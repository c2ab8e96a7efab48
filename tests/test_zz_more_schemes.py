"""More scheme-level parity cases against the CPU oracle (same bar as test_gpu_schemes.py: iteration count, residual history
<= 1e-10, mean stress <= 1e-9, strain / stress / displacement fields) on paths the headline configuration does not take:
more than three phases (the non-marching fused sweep), the exact-residual CG variant (explicit operator result), and a
non-cubic cell (hx != hy != hz in the fused stencils)."""
import numpy as np
import pytest

from oracle import fg_oracle as fo
import fibergen_b200 as fb
from microstructures import sphere_phi
from test_gpu_schemes import build_pair, compare, el_phases

pytestmark = pytest.mark.gpu


def test_cg_four_phases():
    """four isotropic phases: VoigtMixedMaterialLaw over all of them (fg:12752) in the fused direction/stress/div sweep"""
    n = (16, 16, 16)
    incl = [sphere_phi(n, R=0.2, sub=2, center=c) for c in ((0.25, 0.25, 0.25), (0.75, 0.75, 0.25), (0.5, 0.5, 0.75))]
    assert max(np.max(a + b) for a in incl for b in incl if a is not b) <= 1.0          # disjoint inclusions
    phases = []
    lam, mu = fb.lame(1.0, 0.3)
    phases.append(("matrix", "iso", (mu, lam), fo.LinearIsotropic(mu, lam), 1 - sum(incl)))
    for k, (E, nu) in enumerate(((10.0, 0.3), (40.0, 0.2), (4.0, 0.4))):
        lam, mu = fb.lame(E, nu)
        phases.append(("incl%d" % k, "iso", (mu, lam), fo.LinearIsotropic(mu, lam), incl[k]))
    s, o = build_pair(n, phases=phases, method="cg", gamma_scheme="staggered", error_estimator="residual", tol=1e-8)
    compare(s, o, E=[0.2, 0.1, -0.3, 0.4, 0.0, 0.25])


@pytest.mark.parametrize("reinit", [1, 4])
def test_cg_exact_residual_reinit(reinit):
    """cg_reinit > 0: the residual is recomputed from the current strain every `reinit` iterations (fg:23221-23235)"""
    n = (16, 12, 10)
    s, o = build_pair(n, phases=el_phases(n, contrast=20.0), method="cg", gamma_scheme="staggered", error_estimator="residual",
                      tol=1e-8, cg_reinit=reinit)
    compare(s, o, E=[0.3, -0.1, 0.2, 0.5, 0.1, -0.4])


@pytest.mark.parametrize("method,scheme,ee", [("cg", "staggered", "residual"), ("basic", "collocated", "sigma")])
def test_non_cubic_cell(method, scheme, ee):
    """cell 1 x 2 x 1.5 on a 12 x 16 x 10 grid: three different voxel sizes"""
    n, L = (12, 16, 10), (1.0, 2.0, 1.5)
    s, o = build_pair(n, L=L, phases=el_phases(n, contrast=15.0), method=method, gamma_scheme=scheme, error_estimator=ee, tol=1e-8)
    compare(s, o, E=[0.1, 0.2, 0.3, -0.2, 0.15, 0.05])
