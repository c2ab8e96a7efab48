"""Golden cases: small, fully specified problems of the hot path whose oracle results are committed as fixtures
(tests/golden/golden_r02.npz, written by tests/golden/make_golden.py).  The reference itself cannot run in this image
 and ships no stored vectors, so the fixtures freeze the oracle -- which is pinned to the reference's own
known answers by tests/test_oracle_pinning.py -- and let the device path be checked without executing the oracle.

A case = (key, grid, solver settings, phases, load).  `phases` entries: (name, law id of the C ABI, parameters, phi);
geometry comes from tests/microstructures.py (deterministic)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from microstructures import sphere_phi, sphere_normals  # noqa: E402

FIXTURE = os.path.join(HERE, "golden_r02.npz")
NSAMPLES = 128


def lame(E, nu):
    return E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))


def cases():
    out = []
    lam1, mu1 = lame(1.0, 0.3)
    lam2, mu2 = lame(40.0, 0.3)
    lam3, mu3 = lame(10.0, 0.3)
    n = (24, 20, 18)
    phi = sphere_phi(n, R=0.25, sub=3)
    out.append(dict(key="cg_staggered_24x20x18", n=n, mode="elasticity",
                    settings=dict(method="cg", gamma_scheme="staggered", error_estimator="residual", tol=1e-8),
                    phases=[("matrix", "iso", (mu1, lam1), 1 - phi), ("sphere", "iso", (mu2, lam2), phi)],
                    normals=None, E=[0.3, -0.1, 0.2, 0.5, 0.1, -0.4]))
    n = (16, 16, 16)
    phi = sphere_phi(n, R=0.25, sub=3)
    out.append(dict(key="basic_staggered_16", n=n, mode="elasticity",
                    settings=dict(method="basic", gamma_scheme="staggered", error_estimator="sigma", tol=1e-8),
                    phases=[("matrix", "iso", (mu1, lam1), 1 - phi), ("sphere", "iso", (mu3, lam3), phi)],
                    normals=None, E=[1, 0, 0, 0, 0, 0]))
    out.append(dict(key="cg_collocated_16", n=n, mode="elasticity",
                    settings=dict(method="cg", gamma_scheme="collocated", error_estimator="residual", tol=1e-8),
                    phases=[("matrix", "iso", (mu1, lam1), 1 - phi), ("sphere", "iso", (mu3, lam3), phi)],
                    normals=None, E=[0, 0, 1, 0.5, 0, 0]))
    n = (16, 12, 10)
    phi = sphere_phi(n, R=0.3, sub=3)
    out.append(dict(key="heat_cg_laminate_16x12x10", n=n, mode="heat",
                    settings=dict(method="cg", gamma_scheme="staggered", mixing_rule="laminate", error_estimator="residual", tol=1e-8),
                    phases=[("matrix", "iso", (1.0,), 1 - phi), ("fibre", "iso", (10.0,), phi)],
                    normals=sphere_normals(n), E=[1, 0.3, 0]))
    n = (12, 12, 12)
    phi = sphere_phi(n, R=0.3, sub=1)
    out.append(dict(key="neo_hooke_newton_cg_12", n=n, mode="hyperelasticity",
                    settings=dict(method="cg", error_estimator="residual", outer_error_estimator="sigma", tol=1e-6),
                    phases=[("matrix", "nh", (10.0, 10.0), 1 - phi), ("incl", "nh", (10.0, 100.0), phi)],
                    normals=None, E=[1, 1.1, 1, 0, 0, 0, 0, 0, 0]))
    # round 2: Willot's rotated scheme (needs a reference material with lambda_0 != 0) and the collocated viscosity Delta operator
    n = (16, 12, 10)
    phi = sphere_phi(n, R=0.3, sub=3)
    out.append(dict(key="cg_willot_16x12x10", n=n, mode="elasticity",
                    settings=dict(method="cg", gamma_scheme="willot", error_estimator="residual", tol=1e-8),
                    phases=[("matrix", "iso", (mu1, lam1), 1 - phi), ("sphere", "iso", (mu3, lam3), phi)],
                    normals=None, E=[1, 0, 0, 0, 0.3, 0], ref=(1.0, 2.0)))
    n = (12, 10, 8)
    phi = sphere_phi(n, R=0.3, sub=1)
    out.append(dict(key="viscosity_cg_collocated_12x10x8", n=n, mode="viscosity",
                    settings=dict(method="cg", gamma_scheme="collocated", error_estimator="residual", tol=1e-7),
                    phases=[("fluid", "iso", (1.0,), 1 - phi), ("solid", "iso", (1e-3,), phi)],
                    normals=None, E=[0, 0, 0, 0, 0, 1.0]))
    for c in out:
        c.setdefault("ref", None)
    return out


def oracle_law(fo, mode, law, params):
    if law == "iso" and mode == "viscosity":
        return fo.ScalarLinearIsotropic(0.5 * params[0], 6)          # fluidity, fg:15234-15239
    if law == "iso":
        return fo.ScalarLinearIsotropic(params[0], 3) if mode == "heat" else fo.LinearIsotropic(*params)
    if law == "nh":
        return fo.NeoHooke(*params)
    raise ValueError(law)


def sample_index(n, dim):
    """fixed pseudo-random voxel sample of the solution field (flat indices into one component plane)"""
    rng = np.random.default_rng(20260101)
    return rng.integers(0, n[0] * n[1] * n[2], size=NSAMPLES)


def summarize(eps, n):
    """what the fixture stores of a field: per-component mean and L2 norm, and the sampled voxels"""
    d = eps.shape[0]
    flat = eps.reshape(d, -1)
    idx = sample_index(n, d)
    return flat.mean(axis=1), np.sqrt((flat * flat).mean(axis=1)), flat[:, idx]
