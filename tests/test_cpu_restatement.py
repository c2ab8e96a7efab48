"""The OpenMP C++ restatement of the CG iteration (oracle/fg_cpu.cpp, the CPU baseline of bench.py) against the numpy oracle:
same reference material, residual history and mean stress on the same inputs."""
import numpy as np
import pytest

from oracle import fg_oracle as fo, fg_cpu
from microstructures import sphere_phi


@pytest.mark.parametrize("n,L", [((32, 16, 8), (1.0, 0.7, 0.4)), ((16, 16, 16), (1., 1., 1.)), ((8, 4, 2), (1., 1., 1.))])
def test_cpu_restatement_matches_oracle(n, L):
    phi = sphere_phi(n, R=0.3, sub=3)
    mats = ((0.4, 0.6), (4.0, 6.0))
    E = [0.3, -0.1, 0.2, 0.5, 0.1, -0.4]
    r = fg_cpu.cg_iterations(n, L, phi, mats, E, warm=2, steps=6)
    o = fo.LSSolver(*n, *L, mode="elasticity", method="cg", gamma_scheme="staggered", error_estimator="residual", tol=1e-300, maxiter=7)
    o.add_phase("m", fo.LinearIsotropic(*mats[0]), 1 - phi)
    o.add_phase("f", fo.LinearIsotropic(*mats[1]), phi)
    o.setStrain(E)
    o.run()
    ro = np.array(o.residuals)
    assert len(ro) == 8
    assert abs(r["mu_0"] - o.mu_0) <= 1e-14 * o.mu_0
    assert np.abs(r["residuals"] - ro).max() <= 1e-12
    assert np.abs(r["mean_stress"] - o.calcMeanStress()).max() <= 1e-12
    assert r["threads"] >= 1


def test_cpu_restatement_refuses_other_lengths():
    with pytest.raises(ValueError):
        fg_cpu.cg_iterations((12, 8, 8), (1., 1., 1.), np.zeros((12, 8, 8)), ((1., 1.), (2., 2.)), [1, 0, 0, 0, 0, 0])
