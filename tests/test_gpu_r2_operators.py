"""Round-2 operators through the C ABI against the CPU oracle: the collocated Fourier div / grad of the hyperelastic
post-processing (G0DivOperatorFourierHyper fg:20155, GradOperatorFourierHyper fg:22069) with the reference's own identities
(fg:24518-24583), Willot's rotated-scheme operator (fg:19083), the zero-trace collocated Delta operator of the viscosity mode
(fg:20462-20471), the pressure field (fg:15559-15573) and the load-step extrapolation (fg:21468-21513)."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import fg_oracle as fo
import fibergen_b200 as fb
from fibergen_b200.solver import _dp
from microstructures import sphere_phi
from test_gpu_schemes import build_pair, compare, el_phases

pytestmark = pytest.mark.gpu

GRIDS = [((2, 1, 1), (1., 1., 1.)), ((41, 33, 11), (1., 1., 1.)), ((41, 33, 11), (41., 33., 11.)), ((12, 9, 1), (1., 2., 1.)),
         ((16, 16, 16), (1., 1., 1.)), ((64, 8, 6), (2., 1., 3.)), ((7, 5, 3), (1., 1., 1.))]
MU0, LAM0 = 1324.3, 324.2
TOL = math.sqrt(np.finfo(float).eps)


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("n,L", GRIDS)
def test_g0div_and_grad_hyper(n, L):
    ctx = fb.Context(*n, *L, mode="hyperelasticity", gamma_scheme="staggered")
    o = fo.LSSolver(*n, *L, mode="hyperelasticity")
    rng = np.random.default_rng(11)
    tau = rng.standard_normal((9,) + n)
    for mu0, lam0, alpha in ((MU0, LAM0, 1.0), (0.7, 0.0, -1.0)):
        f = ctx.field(tau)
        ctx.chk(ctx.lib.fgb_g0div_hyper(ctx.h, f, mu0, lam0, alpha))
        want = o.G0DivOperatorHyper(mu0, lam0, tau, alpha)
        assert relerr(ctx.download(f)[:3], want) < 5e-11      # O(p^2) stages at the prime lengths 41, 11 + cancellation in i xi . tau
        ctx.chk(ctx.lib.fgb_field_free(ctx.h, f))
    q = np.zeros((9,) + n)
    q[:3] = rng.standard_normal((3,) + n)
    f = ctx.field(q)
    ctx.chk(ctx.lib.fgb_grad_hyper(ctx.h, f))
    want = o.ifft(o.GradOperatorFourierHyper(o.fft(q[:3])))
    assert relerr(ctx.download(f), want) < 2e-12
    ctx.close()


@pytest.mark.parametrize("n,L", GRIDS[:3])
def test_reference_hyper_identities_on_device(n, L):
    """fg:24518-24555 (grad G0 Div C0 grad u == grad u) and fg:24558-24583 (Gamma_collocated == grad G0 Div) on the device"""
    ctx = fb.Context(*n, *L, mode="hyperelasticity", gamma_scheme="collocated")
    rng = np.random.default_rng(12)
    q = np.zeros((9,) + n)
    q[:3] = rng.random((3,) + n)
    W = ctx.field(q)
    T = ctx.field()
    ctx.chk(ctx.lib.fgb_grad_hyper(ctx.h, W))
    org = ctx.download(W)
    ctx.chk(ctx.lib.fgb_calc_stress_const(ctx.h, W, T, MU0, LAM0))
    ctx.chk(ctx.lib.fgb_g0div_hyper(ctx.h, T, MU0, LAM0, 1.0))
    ctx.chk(ctx.lib.fgb_grad_hyper(ctx.h, T))
    scale = max(1.0, np.abs(org).max())
    assert np.linalg.norm(np.abs(ctx.download(T) - org).reshape(9, -1).max(axis=1)) <= TOL * scale
    ctx.upload(W, rng.random((9,) + n))
    ctx.gamma(W, np.zeros(9), MU0, LAM0, -1.0, 0.0)
    W1 = ctx.download(W)
    ctx.chk(ctx.lib.fgb_calc_stress_const(ctx.h, W, T, MU0, LAM0))
    ctx.chk(ctx.lib.fgb_grad_g0div_hyper(ctx.h, T, MU0, LAM0, 1.0))        # both operators in Fourier space, as fg:24572-24575
    scale = max(1.0, np.abs(W1).max())
    assert np.linalg.norm(np.abs(ctx.download(T) - W1).reshape(9, -1).max(axis=1)) <= TOL * scale
    ctx.close()


@pytest.mark.parametrize("n,L", GRIDS)
def test_willot_operator(n, L):
    """GammaOperatorWillotR (fg:20322) vs the oracle, and the reference's `WillotR epsG0div identity` (fg:24107) on the device"""
    ctx = fb.Context(*n, *L, mode="elasticity", gamma_scheme="willot")
    o = fo.LSSolver(*n, *L, mode="elasticity", gamma_scheme="willot")
    rng = np.random.default_rng(13)
    tau = rng.standard_normal((6,) + n)
    E = rng.standard_normal(6)
    for mu0, lam0, alpha, beta in ((MU0, LAM0, 1.0, 0.0), (2.0, 1.0, -8.0, 1.0)):
        f = ctx.field(tau)
        ctx.gamma(f, E, mu0, lam0, alpha, beta)
        o.set_reference(mu0, lam0)
        o.setBCProjector(fo.Id4(6))
        want = o.GammaOperator(E, mu0, lam0, tau, alpha, beta)
        got = ctx.download(f)
        if np.isnan(want).any():
            # 2x1x1: the rotated scheme's frequency vector vanishes at the Nyquist frequency (tan(pi/2) * (1 + e^{i pi}) = inf * 0)
            assert np.array_equal(np.isnan(got), np.isnan(want))
        else:
            assert relerr(got, want) < 5e-12
        ctx.chk(ctx.lib.fgb_field_free(ctx.h, f))
    ctx.close()


def test_willot_scheme_solve():
    n = (16, 16, 16)
    for method, ee in (("cg", "residual"), ("basic", "sigma")):
        s = fb.LSSolver(*n, mode="elasticity", method=method, gamma_scheme="willot", error_estimator=ee, tol=1e-8)
        o = fo.LSSolver(*n, mode="elasticity", method=method, gamma_scheme="willot", error_estimator=ee, tol=1e-8)
        for m, (name, law, params, olaw, phi) in enumerate(el_phases(n)):
            s.add_material(name, law, *params)
            o.add_phase(name, olaw, phi)
        s.set_reference(1.0, 2.0)
        o.set_reference(1.0, 2.0)
        s.init()
        for m, (name, law, params, olaw, phi) in enumerate(el_phases(n)):
            s.set_phase(m, phi)
        compare(s, o, E=[1, 0, 0, 0, 0.3, 0])


@pytest.mark.parametrize("scheme", ["collocated", "willot", "staggered"])
@pytest.mark.parametrize("method,ee", [("cg", "residual"), ("basic", "epsilon")])
def test_viscosity_all_schemes(scheme, method, ee):
    """DeltaOperatorCollocated (zero-trace transform, fg:20462), DeltaOperatorWillotR (fg:20380), DeltaOperatorStaggered (fg:20422)"""
    n = (12, 10, 8)
    phi = sphere_phi(n, R=0.3, sub=1)
    kw = dict(mode="viscosity", method=method, gamma_scheme=scheme, error_estimator=ee, tol=1e-7)
    if (scheme, method) == ("willot", "cg"):
        # CG on the rotated scheme's Delta operator of this stiff suspension is not a contraction (the residual spikes to 0.9) and
        # amplifies rounding: the oracle run against itself with one parameter changed by 1e-15 differs by 2e-2 after 6 iterations.
        # The first iterations are still a sharp operator-level comparison.
        kw["maxiter"] = 4
    s = fb.LSSolver(*n, **kw)
    o = fo.LSSolver(*n, **kw)
    s.add_material("fluid", "iso", 1.0)
    s.add_material("solid", "iso", 1e-3)
    if scheme == "willot":
        s.set_reference(1.0, 5.0)
        o.set_reference(1.0, 5.0)
    s.init()
    s.set_phase(0, 1 - phi)
    s.set_phase(1, phi)
    o.add_phase("fluid", fo.ScalarLinearIsotropic(0.5 * 1.0, 6), 1 - phi)
    o.add_phase("solid", fo.ScalarLinearIsotropic(0.5 * 1e-3, 6), phi)
    if (scheme, method) == ("willot", "basic"):
        # the fixed-point iteration with the rotated scheme's Delta operator is not a contraction for this suspension: the reference
        # algorithm diverges until the residual is NaN (fg:21202) -- on the device exactly as in the oracle
        s.set_strain([0, 0, 0, 0, 0, 1.0])
        o.setStrain([0, 0, 0, 0, 0, 1.0])
        with pytest.raises(fb.FgbError, match="NaN detected"):
            s.run()
        with pytest.raises(ArithmeticError, match="NaN detected"):
            o.run()
        return
    if scheme == "willot":
        # the rotated scheme leaves the hydrostatic part of the fluid stress to rounding (component 11 drifts at the 1e-8 level while
        # the residual history agrees to 1e-13), so the fields are compared at 1e-6 here
        s.set_strain([0, 0, 0, 0, 0, 1.0])
        o.setStrain([0, 0, 0, 0, 0, 1.0])
        s.run()
        o.run()
        rs, ro = s.get_residuals(), np.array(o.residuals)
        assert len(rs) == len(ro) and np.abs(rs - ro).max() <= 1e-10 * np.abs(ro).max()
        assert np.abs(s.get_mean_stress() - o.calcMeanStress()).max() <= 1e-9 * np.abs(o.calcMeanStress()).max()
        assert np.abs(s.get_field() - o.epsilon).max() <= 1e-6 * np.abs(o.epsilon).max()
    else:
        compare(s, o, E=[0, 0, 0, 0, 0, 1.0])
    p_o = o.calcPressure()
    p_s = s.get_field("p")
    assert p_s.shape == (1,) + n
    assert np.abs(p_s[0] - p_o).max() <= (1e-6 if scheme == "willot" else 1e-8) * max(np.abs(p_o).max(), 1e-300)


@pytest.mark.parametrize("order", [1, 2])
def test_loadstep_extrapolation(order):
    """loadstep_extrapolation_order > 0: polynomial start value from the previous load steps (fg:21634-21650, fg:21468-21513)"""
    n = (10, 10, 10)
    phi = sphere_phi(n, R=0.3, sub=1)
    phases = [("matrix", "nh", (10.0, 10.0), fo.NeoHooke(10.0, 10.0), 1 - phi),
              ("incl", "nh", (10.0, 100.0), fo.NeoHooke(10.0, 100.0), phi)]
    s, o = build_pair(n, mode="hyperelasticity", phases=phases, method="cg", error_estimator="residual",
                      outer_error_estimator="sigma", tol=1e-6, loadsteps=4, loadstep_extrapolation_order=order)
    compare(s, o, E=np.array([1, 1.1, 1, 0, 0, 0, 0, 0, 0], dtype=float))


def test_loads_before_init_and_setting_validation():
    """loads set before init() must neither corrupt memory nor be dropped; malformed values raise"""
    n = (8, 8, 8)
    s = fb.LSSolver(*n, mode="hyperelasticity", method="cg", error_estimator="residual", tol=1e-6)
    F = np.array([1, 1.05, 1, 0, 0.01, 0, 0, 0.02, 0], dtype=float)
    s.set_strain(F)                                   # before init: kept
    s.add_material("m", "nh", 10.0, 10.0)
    s.init()
    s.set_phase(0, np.ones(n))
    s.run()
    assert np.allclose(s.get_mean_strain(), F, atol=1e-12)
    for key, val in (("tol", "abc"), ("maxiter", "ten"), ("loadsteps", "0"), ("tol", "[0.25"), ("freq_hack", "maybe")):
        with pytest.raises(fb.FgbError):
            s.set(key, val)
    s.set("loadsteps", [0.0, 0.25, 1.0])              # sequences are joined with ','


def test_device_phase_init_matches_restated_initphi():
    """fgb_init_phase_capsules (initPhi fg:17489 on the device) against the C restatement oracle/fg_phase.c: the Hashin demo's coated
    sphere (three phases, priority of the last material) and a periodic short-fibre cell with its images"""
    from oracle import fg_phase as fp
    from microstructures import rsa_capsules, fiber_list
    from test_oracle_pinning import HASHIN_FIBERS
    cases = [((64, 64, 64), HASHIN_FIBERS, 3)]
    n = (48, 48, 48)
    Cs, Ds, R, Lc = rsa_capsules(n, seed=3, vol_frac=0.12, diameter_vox=6.0, aspect=4.0, max_tries=400)
    fibs, box = fiber_list(n, Cs, Ds, R, Lc, material=1)
    assert abs(box[0] - 1) < 1e-15
    cases.append((n, fibs, 2))
    for n, fibers, nmat in cases:
        s = fb.LSSolver(*n, mode="elasticity")
        for m in range(nmat):
            s.add_material("m%d" % m, "iso", 1.0 + m, 1.0)
        s.init()
        s.init_phase(fibers, matrix_mat=0, normals=True)
        want, cnt = fp.init_phi(n, (1., 1., 1.), fibers, nmat)
        got = np.stack([s.get_phase(m) for m in range(nmat)])
        assert cnt > 1000
        assert np.abs(got - want).max() <= 1e-12
        assert np.abs(got.sum(axis=0) - 1).max() <= 1e-15
        s.close()


def test_pipelined_cg_is_the_same_iteration():
    """fgb_cgdev_* (gamma, beta, alpha resident on the device, operator application enqueued ahead of the stop test) against the
    host-scalar loop: identical residual histories, bit for bit, and the same solution"""
    n = (32, 24, 20)
    res, eps = [], []
    for pipelined in (True, False):
        s = fb.LSSolver(*n, mode="elasticity", method="cg", gamma_scheme="staggered", error_estimator="residual", tol=1e-9,
                        pipelined_cg=pipelined)
        for name, law, params, olaw, phi in el_phases(n, contrast=30.0):
            s.add_material(name, law, *params)
        s.init()
        for m, (name, law, params, olaw, phi) in enumerate(el_phases(n, contrast=30.0)):
            s.set_phase(m, phi)
        s.set_strain([0.3, -0.1, 0.2, 0.5, 0.1, -0.4])
        s.run()
        res.append(s.get_residuals())
        eps.append(s.get_field())
        s.close()
    assert len(res[0]) == len(res[1]) > 10
    assert np.array_equal(res[0], res[1])
    assert np.array_equal(eps[0], eps[1])
    # a path without the fused sweep (collocated: explicit operator result, generic kernels)
    res = []
    for pipelined in (True, False):
        s = fb.LSSolver(*n, mode="elasticity", method="cg", gamma_scheme="collocated", error_estimator="residual", tol=1e-9,
                        pipelined_cg=pipelined)
        for name, law, params, olaw, phi in el_phases(n, contrast=30.0):
            s.add_material(name, law, *params)
        s.init()
        for m, (name, law, params, olaw, phi) in enumerate(el_phases(n, contrast=30.0)):
            s.set_phase(m, phi)
        s.set_strain([0.3, -0.1, 0.2, 0.5, 0.1, -0.4])
        s.run()
        res.append(s.get_residuals())
        s.close()
    assert np.array_equal(res[0], res[1])


@pytest.mark.parametrize("mode,d", [("elasticity", 6), ("heat", 3), ("hyperelasticity", 9)])
def test_dfg_transfer_and_sweeps(mode, d):
    """prolongate_to_dfg / restrict_from_dfg (fg:14216-14339): the reference's `staggered dfg operator` identity (fg:24491-24515) on
    the device, and the constitutive sweep through the doubly fine grid (fg:18143-18149, fg:18343-18347) against the oracle"""
    n = (10, 6, 7)
    nf = tuple(2 * x for x in n)
    ctx = fb.Context(*n, mode=mode, gamma_scheme="staggered")
    ctx.chk(ctx.lib.fgb_set_dfg(ctx.h, 2))
    rng = np.random.default_rng(21)
    c1 = rng.random((d,) + n)
    f = ctx.field(c1)
    ctx.chk(ctx.lib.fgb_dfg_prolongate(ctx.h, f))
    ctx.upload(f, np.zeros((d,) + n))
    ctx.chk(ctx.lib.fgb_dfg_restrict(ctx.h, f))
    assert np.abs(ctx.download(f) - c1).max() <= 4e-16
    phi = sphere_phi(nf, R=0.3, sub=2)
    o = fo.LSSolver(*n, mode=mode, gamma_scheme="full_staggered")
    if mode == "heat":
        laws = [("scalar", [1.0]), ("scalar", [10.0])]
        olaws = [fo.ScalarLinearIsotropic(1.0, 3), fo.ScalarLinearIsotropic(10.0, 3)]
    elif mode == "elasticity":
        laws = [("iso", [0.4, 0.6]), ("iso", [4.0, 6.0])]
        olaws = [fo.LinearIsotropic(0.4, 0.6), fo.LinearIsotropic(4.0, 6.0)]
    else:
        laws = [("nh", [10.0, 10.0]), ("nh", [10.0, 100.0])]
        olaws = [fo.NeoHooke(10.0, 10.0), fo.NeoHooke(10.0, 100.0)]
    ctx.lnx, ctx.ny, ctx.nz, ctx.nzp = nf[0], nf[1], nf[2], fb.nzp_of(nf[2])        # set_phases pads to the shape it is given
    ctx.set_phases([1 - phi, phi], laws)
    ctx.lnx, ctx.ny, ctx.nz, ctx.nzp = n[0], n[1], n[2], fb.nzp_of(n[2])
    o.add_phase("a", olaws[0], 1 - phi)
    o.add_phase("b", olaws[1], phi)
    eps = 0.05 * rng.standard_normal((d,) + n)
    if d == 9:
        eps[:3] += 1.0
    src, dst = ctx.field(eps), ctx.field()
    ctx.chk(ctx.lib.fgb_calc_stress(ctx.h, src, dst, 0.7, 0.3, 1.0))
    want = o.calcStress(0.7, 0.3, eps, 1.0)
    assert relerr(ctx.download(dst), want) < 1e-13
    assert relerr(ctx.mean_pk1(src), o.calcMeanStress(eps)) < 1e-13
    ctx.close()


@pytest.mark.parametrize("scheme", ["half_staggered", "full_staggered"])
@pytest.mark.parametrize("kind", ["elasticity", "heat_laminate", "neo_hooke"])
def test_dfg_schemes(scheme, kind):
    """gamma_scheme half_staggered / full_staggered (use_dfg fg:14894) end to end against the oracle"""
    n = (12, 10, 8)
    nf = tuple(2 * x for x in n)
    fine = scheme == "full_staggered"
    phi = sphere_phi(nf if fine else n, R=0.3, sub=1 if fine else 2)
    normals = None
    if kind == "elasticity":
        mode, kw = "elasticity", dict(method="cg", error_estimator="residual", tol=1e-8)
        phases = [("m", "iso", (0.4, 0.6), fo.LinearIsotropic(0.4, 0.6), 1 - phi), ("f", "iso", (4.0, 6.0), fo.LinearIsotropic(4.0, 6.0), phi)]
        E = [1, 0, 0, 0, 0, 0.2]
    elif kind == "heat_laminate":
        mode, kw = "heat", dict(method="cg", mixing_rule="laminate", error_estimator="residual", tol=1e-8)
        phases = [("m", "iso", (1.0,), fo.ScalarLinearIsotropic(1.0, 3), 1 - phi), ("f", "iso", (10.0,), fo.ScalarLinearIsotropic(10.0, 3), phi)]
        from microstructures import sphere_normals
        normals = sphere_normals(nf)          # with the doubly fine grid the normals live on the fine grid (fg:14925-14937)
        E = [1, 0.3, 0]
    else:
        mode, kw = "hyperelasticity", dict(method="cg", error_estimator="residual", outer_error_estimator="sigma", tol=1e-6)
        phases = [("m", "nh", (10.0, 10.0), fo.NeoHooke(10.0, 10.0), 1 - phi), ("f", "nh", (10.0, 100.0), fo.NeoHooke(10.0, 100.0), phi)]
        E = np.array([1, 1.1, 1, 0, 0, 0, 0, 0, 0], dtype=float)
    s, o = build_pair(n, mode=mode, phases=phases, normals=normals, gamma_scheme=scheme, **kw)
    compare(s, o, E=E)


def test_nunan_keller_viscosity_demo():
    """demo/python/nunan_keller/project.xml: effective viscosity of a simple cubic lattice of rigid spheres (viscosity mode, CG,
    full_staggered, n = 32, tol 1e-5, smooth_tol 1e-5) against the (alpha, beta) of Nunan & Keller (1984) that the demo prints next
    to its result (project.xml:21-31).  The five traceless load cases and the 5x5 inversion are those of calc_effective_properties
    (fg:26264-26312); alpha = (mu_eff[0][0] - mu_eff[0][1])/2 - 1, beta = mu_eff[3][3] - 1 (project.xml:38-40)."""
    theory = {0.12: (0.46580, 0.28995), 0.28: (2.1459, 0.74379)}
    n = (32, 32, 32)
    for V, (alpha_t, beta_t) in theory.items():
        R = (V / (4 * math.pi / 3)) ** (1 / 3.0)
        s = fb.LSSolver(*n, mode="viscosity", method="cg", gamma_scheme="full_staggered", tol=1e-5, smooth_tol=1e-5)
        s.add_material("matrix", "iso", 1.0)
        s.add_material("fiber", "iso", 0.0)
        s.init()
        s.init_phase([((0.5, 0.5, 0.5), (1, 0, 0), 0.0, R, 1)])
        Ecols = np.zeros((6, 5))
        Ecols[0, 0] = Ecols[1, 1] = 1
        Ecols[1, 0] = Ecols[2, 1] = -1
        Ecols[3, 2] = Ecols[4, 3] = Ecols[5, 4] = 1
        S = np.zeros((6, 5))
        for i in range(5):
            s.set_strain(Ecols[:, i])
            s.run()
            S[:, i] = s.get_mean_stress()
        C55 = Ecols[1:, :] @ np.linalg.inv(S[1:, :])                     # "2*eta" (5x5), fg:26300-26301
        C = np.zeros((6, 6))
        C[1:, 1:] = C55
        for i in range(5):
            if S[0, i] != 0:
                for j in range(1, 6):
                    C[j, 0] = (Ecols[j, i] - C[j, 1:] @ S[1:, i]) / S[0, i]
                break
        C[0, :] = -(C[1, :] + C[2, :])
        C[:, :3] -= C[:, :3].min(axis=1, keepdims=True)
        Cv = C.copy()
        Cv[:, 3:] *= 0.5                                                  # Voigt notation, fg:26343-26349
        alpha = 0.5 * (Cv[0, 0] - Cv[0, 1]) - 1
        beta = Cv[3, 3] - 1
        print("Nunan-Keller V=%.2f: alpha %.4f (theory %.4f), beta %.4f (theory %.4f)" % (V, alpha, alpha_t, beta, beta_t))
        assert abs(alpha - alpha_t) <= 0.05 * alpha_t
        assert abs(beta - beta_t) <= 0.05 * beta_t
        s.close()


@pytest.mark.parametrize("mixing", ["voigt", "laminate", "reuss"])
@pytest.mark.parametrize("method,ee", [("cg", "residual"), ("basic", "sigma")])
@pytest.mark.parametrize("n", [(24, 20, 18), (16, 12, 300), (9, 7, 5)])
def test_fused_heat_path(mixing, method, ee, n):
    """the fused heat sweeps (per-voxel effective conductivity, marching flux/div sweep, implicit gradient) against the oracle, which
    evaluates the mixed law voxel by voxel every iteration; includes axis-aligned interface normals (n_k = 0 -> NaN in the laminate's
    component-wise jump formula, fg:13236-13239 / fg:9375, reproduced)"""
    phi = sphere_phi(n, R=0.3, sub=3)
    phases = [("matrix", "iso", (1.0,), fo.ScalarLinearIsotropic(1.0, 3), 1 - phi),
              ("fibre", "iso", (10.0,), fo.ScalarLinearIsotropic(10.0, 3), phi)]
    from microstructures import sphere_normals
    s, o = build_pair(n, mode="heat", phases=phases, normals=sphere_normals(n) if mixing == "laminate" else None,
                      method=method, gamma_scheme="staggered", mixing_rule=mixing, error_estimator=ee, tol=1e-8)
    if mixing == "laminate" and n == (9, 7, 5):
        # odd grid: the normals of the voxels in the sphere's mid-planes have a component that is exactly zero, the reference's
        # component-wise jump Hessian is singular there and the solution turns NaN (fg:9375, fg:21202) -- identically on the device
        s.set_strain([1, 0.3, -0.2])
        o.setStrain([1, 0.3, -0.2])
        with pytest.raises(fb.FgbError, match="NaN detected"):
            s.run()
        with pytest.raises(ArithmeticError, match="NaN detected"):
            o.run()
        return
    compare(s, o, E=[1, 0.3, -0.2])
    # the same run through the generic kernels gives the same history (A/B of the fusion)
    assert s.lib.fgb_cg_implicit_w_supported(s.ctx()) == 1


@pytest.mark.parametrize("n", [(12, 12, 12), (16, 10, 20)])
@pytest.mark.parametrize("loadsteps", [1, 2])
def test_fused_neo_hooke_newton_cg(n, loadsteps):
    """fused Neo-Hooke inner CG (per-voxel tangent cache, implicit operator result, device-resident scalars) against the oracle, and
    against the generic host-scalar loop of the same library (pipelined_cg = 0)"""
    phi = sphere_phi(n, R=0.3, sub=2)
    phases = [("matrix", "nh", (10.0, 10.0), fo.NeoHooke(10.0, 10.0), 1 - phi),
              ("incl", "nh", (10.0, 100.0), fo.NeoHooke(10.0, 100.0), phi)]
    kw = dict(mode="hyperelasticity", phases=phases, method="cg", error_estimator="residual", outer_error_estimator="sigma", tol=1e-6,
              loadsteps=loadsteps)
    F = np.array([1, 1.1, 1, 0, 0.02, 0, 0, 0, 0.01], dtype=float)
    s, o = build_pair(n, **kw)
    rs = compare(s, o, E=F)
    s2, _ = build_pair(n, **kw)
    s2.set("pipelined_cg", False)
    s2.set_strain(F)
    s2.run()
    r2 = s2.get_residuals()
    assert len(r2) == len(rs) and np.abs(r2 - rs).max() <= 1e-10 * np.abs(rs).max()
    assert np.abs(s2.get_field() - s.get_field()).max() <= 1e-10
