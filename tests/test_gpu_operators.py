"""Operator-level parity: every CUDA kernel reached through the C ABI (fgb_*) against the CPU oracle on the
same seeded inputs, on the reference's own test grids (2x1x1, 41x33x11 with L=1 and L=n, fg:27259-27273),
degenerate 2-D / 1-D grids used by the demos, and power-of-two cubes.  FP64 tolerances are written per test."""
import math

import numpy as np
import pytest

from oracle import fg_oracle as fo
import fibergen_b200 as fb
from microstructures import sphere_phi, sphere_normals

pytestmark = pytest.mark.gpu

GRIDS = [((2, 1, 1), (1., 1., 1.)), ((41, 33, 11), (1., 1., 1.)), ((41, 33, 11), (41., 33., 11.)),
         ((10, 1, 1), (1., 1., 1.)), ((12, 9, 1), (1., 2., 1.)), ((16, 16, 16), (1., 1., 1.)),
         ((32, 8, 20), (2., 1., 3.)), ((7, 5, 3), (1., 1., 1.))]
MODES = [("elasticity", 6), ("heat", 3), ("hyperelasticity", 9)]
MU0, LAM0 = 1324.3, 324.2      # reference material of fibergen --test (fg:24007-24008)


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("n,L", GRIDS)
def test_fft_forward_backward(n, L):
    """FFT3<double>::forward + 1/nxyz scaling and ::backward (fg:7232-7244, fg:18531-18584)"""
    ctx = fb.Context(*n, *L, mode="elasticity", gamma_scheme="collocated")
    rng = np.random.default_rng(0)
    x = rng.standard_normal((6,) + n)
    f = ctx.field(x)
    ctx.chk(ctx.lib.fgb_fft_forward(ctx.h, f))
    got = ctx.download_padded(f)
    got_c = got.reshape(6, n[0], n[1], -1, 2)
    got_c = got_c[..., 0] + 1j * got_c[..., 1]
    o = fo.LSSolver(*n, *L)
    want = o.fft(x)
    assert relerr(got_c, want) < 5e-14
    ctx.chk(ctx.lib.fgb_fft_backward(ctx.h, f))
    assert relerr(ctx.download(f), x) < 5e-14
    ctx.close()


@pytest.mark.parametrize("n,L", GRIDS)
@pytest.mark.parametrize("mode,d", MODES)
def test_div_eps_staggered(n, L, mode, d):
    """divOperatorStaggered* (fg:18853-19071) and epsOperatorStaggered* (fg:18614-18846)"""
    ctx = fb.Context(*n, *L, mode=mode, gamma_scheme="staggered")
    o = fo.LSSolver(*n, *L, mode=mode, gamma_scheme="staggered")
    rng = np.random.default_rng(1)
    tau = rng.standard_normal((d,) + n)
    f = ctx.field(tau)
    ctx.chk(ctx.lib.fgb_div_staggered(ctx.h, f))
    want = o.divOperatorStaggered(tau)
    assert relerr(ctx.u_download(), want) < 1e-13
    u = rng.standard_normal((ctx.udim,) + n)
    E = rng.standard_normal(d)
    ctx.u_upload(u)
    ctx.chk(ctx.lib.fgb_eps_staggered(ctx.h, f, fb.solver._dp(ctx.vec(E))))
    assert relerr(ctx.download(f), o.epsOperatorStaggered(E, u)) < 1e-13
    ctx.close()


@pytest.mark.parametrize("n,L", GRIDS)
@pytest.mark.parametrize("mode,d", MODES)
@pytest.mark.parametrize("scheme", ["staggered", "collocated"])
def test_gamma_operator(n, L, mode, d, scheme):
    """GammaOperator (fg:20488-20531) incl. the Fourier-space Green operators (fg:19302-19927)"""
    ctx = fb.Context(*n, *L, mode=mode, gamma_scheme=scheme)
    o = fo.LSSolver(*n, *L, mode=mode, gamma_scheme=scheme)
    rng = np.random.default_rng(2)
    tau = rng.standard_normal((d,) + n)
    E = rng.standard_normal(d)
    for (mu0, lam0, alpha, beta) in [(MU0, LAM0, 1.0, 0.0), (0.7, 0.0, -1.0, 0.0), (2.0, 1.0, -8.0, 1.0)]:
        f = ctx.field(tau)
        ctx.gamma(f, E, mu0, lam0, alpha, beta)
        o.set_reference(mu0, lam0)
        o.setBCProjector(fo.Id4(d))
        want = o.GammaOperator(E, mu0, lam0, tau, alpha, beta)
        assert relerr(ctx.download(f), want) < 2e-12, (mu0, lam0, alpha, beta)
        ctx.chk(ctx.lib.fgb_field_free(ctx.h, f))
    ctx.close()


def test_gamma_freq_hack():
    n, L = (8, 6, 4), (1., 1., 1.)
    ctx = fb.Context(*n, *L, mode="elasticity", gamma_scheme="collocated")
    ctx.chk(ctx.lib.fgb_set_freq_hack(ctx.h, 1))
    o = fo.LSSolver(*n, *L, mode="elasticity", gamma_scheme="collocated", freq_hack=True)
    o.set_reference(1.3, 0.4)
    o.setBCProjector(fo.Id4(6))
    tau = np.random.default_rng(3).standard_normal((6,) + n)
    f = ctx.field(tau)
    ctx.gamma(f, np.zeros(6), 1.3, 0.4, -1.0, 0.0)
    want = o.GammaOperator(np.zeros(6), 1.3, 0.4, tau, -1.0, 0.0)
    assert relerr(ctx.download(f), want) < 2e-12
    ctx.close()


@pytest.mark.parametrize("n,L", [((2, 1, 1), (1., 1., 1.)), ((41, 33, 11), (1., 1., 1.)), ((41, 33, 11), (41., 33., 11.))])
@pytest.mark.parametrize("mode,d", MODES)
def test_reference_operator_identities(n, L, mode, d):
    """the epsG0div projection identities of `fibergen --test` on the device (fg:23946-23974, fg:24086-24182,
    fg:24460-24488), tolerance sqrt(eps) as check_tol (fg:23502)"""
    tol = math.sqrt(np.finfo(float).eps)
    mu0, lam0 = (1.0, 0.0) if mode == "heat" else (MU0, LAM0)
    z = fb.solver._dp(np.zeros(9))
    rng = np.random.default_rng(4)
    # staggered
    ctx = fb.Context(*n, *L, mode=mode, gamma_scheme="staggered")
    ctx.u_upload(rng.random((ctx.udim,) + n))
    f = ctx.field()
    ctx.chk(ctx.lib.fgb_eps_staggered(ctx.h, f, z))
    org = ctx.download(f)
    ctx.chk(ctx.lib.fgb_calc_stress_const(ctx.h, f, f, mu0, lam0))
    ctx.chk(ctx.lib.fgb_div_staggered(ctx.h, f))
    ctx.chk(ctx.lib.fgb_g0_staggered(ctx.h, mu0, lam0, 1.0))
    ctx.chk(ctx.lib.fgb_eps_staggered(ctx.h, f, z))
    assert np.linalg.norm(np.abs(ctx.download(f) - org).reshape(d, -1).max(axis=1)) <= tol
    ctx.close()
    # collocated
    ctx = fb.Context(*n, *L, mode=mode, gamma_scheme="collocated")
    f = ctx.field(rng.random((d,) + n))
    ctx.gamma(f, np.zeros(d), mu0, lam0, 1.0, 0.0)
    org = ctx.download(f)
    ctx.chk(ctx.lib.fgb_calc_stress_const(ctx.h, f, f, mu0, lam0))
    ctx.gamma(f, np.zeros(d), mu0, lam0, 1.0, 0.0)
    assert np.linalg.norm(np.abs(ctx.download(f) - org).reshape(d, -1).max(axis=1)) <= tol
    ctx.close()


def _two_phase(ctx, o, n, mode, mixing, rng):
    phi = sphere_phi(n, R=0.3, sub=3)
    normals = sphere_normals(n) if mixing == "laminate" else None
    if mode == "elasticity":
        laws = [("iso", [1.0, 1.5]), ("iso", [5.0, 2.0])]
        olaws = [fo.LinearIsotropic(1.0, 1.5), fo.LinearIsotropic(5.0, 2.0)]
    elif mode == "heat":
        laws = [("scalar", [1.0]), ("aniso3", [10.0, 8.0, 6.0, 0.5, 0.2, 0.1])]
        olaws = [fo.ScalarLinearIsotropic(1.0, 3), fo.MatrixLinearAnisotropic(10.0, 8.0, 6.0, 0.5, 0.2, 0.1)]
    else:
        laws = [("nh", [10.0, 10.0]), ("svk", [10.0, 100.0])]
        olaws = [fo.NeoHooke(10.0, 10.0), fo.SaintVenantKirchhoff(10.0, 100.0)]
    ctx.set_phases([1 - phi, phi], laws, mixing=mixing, normals=normals)
    o.add_phase("matrix", olaws[0], 1 - phi)
    o.add_phase("incl", olaws[1], phi)
    if normals is not None:
        o.set_normals(normals)


@pytest.mark.parametrize("mode,d", MODES)
@pytest.mark.parametrize("mixing", ["voigt", "laminate", "reuss"])
def test_constitutive_sweeps(mode, d, mixing):
    """calcStress fg:18134, calcStressDeriv fg:18425, meanPK1 fg:12312, meanW fg:12239, getRefMaterial fg:12153"""
    n = (12, 10, 8)
    rng = np.random.default_rng(5)
    ctx = fb.Context(*n, mode=mode, gamma_scheme="staggered")
    o = fo.LSSolver(*n, mode=mode, gamma_scheme="staggered", mixing_rule=mixing)
    _two_phase(ctx, o, n, mode, mixing, rng)
    eps = 0.05 * rng.standard_normal((d,) + n)
    if d == 9:
        eps[:3] += 1.0
    W = rng.standard_normal((d,) + n)
    fe, fw, fo_ = ctx.field(eps), ctx.field(W), ctx.field()
    # The hyperelastic laminate runs a backtracked Newton iteration per interface voxel whose stopping decisions
    # (Armijo test on an energy difference at rounding level, fg:13378) flip with the last bit; the jump vector is
    # then only converged to ~sqrt(eps_a) either way, so parity is bounded by the Newton tolerance, not by FP64.
    tol = 1e-7 if (d == 9 and mixing == "laminate") else 1e-12
    for (mu0, lam0, alpha) in [(0.0, 0.0, 1.0), (2.5, 0.7, -1.0)]:
        ctx.chk(ctx.lib.fgb_calc_stress(ctx.h, fe, fo_, mu0, lam0, alpha))
        assert relerr(ctx.download(fo_), o.calcStress(mu0, lam0, eps, alpha)) < tol
        ctx.chk(ctx.lib.fgb_calc_stress_deriv(ctx.h, fe, fw, fo_, mu0, lam0, alpha))
        assert relerr(ctx.download(fo_), o.calcStressDeriv(mu0, lam0, eps, W, alpha)) < tol
    assert relerr(ctx.mean_pk1(fe), o.calcMeanStress(eps)) < tol
    if mixing == "reuss":
        with pytest.raises(fb.FgbError, match="energy not implemented"):      # fg:12660
            ctx.mean_energy(fe)
    else:
        assert abs(ctx.mean_energy(fe) - o.calcMeanEnergy(eps)) <= tol * abs(o.calcMeanEnergy(eps))
    lmin, lmax = ctx.ref_material(fe)
    _, olmin, olmax = o.getRefMaterial(eps, False, False)
    if olmin > 0:
        assert abs(lmin - olmin) <= 1e-11 * abs(olmax)
    assert abs(lmax - olmax) <= 1e-11 * abs(olmax)
    ctx.chk(ctx.lib.fgb_check_numeric(ctx.h))
    ctx.close()


@pytest.mark.parametrize("mode,d", [("elasticity", 6), ("heat", 3)])
def test_polarization_map(mode, d):
    """calcPolarizationDim fg:18044-18118 (closed form on pure iso voxels, generic solve elsewhere)"""
    n = (10, 9, 8)
    rng = np.random.default_rng(6)
    ctx = fb.Context(*n, mode=mode, gamma_scheme="collocated")
    o = fo.LSSolver(*n, mode=mode, gamma_scheme="collocated")
    _two_phase(ctx, o, n, mode, "voigt", rng)
    eps = rng.standard_normal((d,) + n)
    fe, fq = ctx.field(eps), ctx.field()
    for inv in (0, 1):
        ctx.chk(ctx.lib.fgb_calc_polarization(ctx.h, fe, fq, 1.7, inv))
        assert relerr(ctx.download(fq), o.calcPolarization(1.7, eps, bool(inv))) < 1e-12
    ctx.close()


@pytest.mark.parametrize("n", [(9, 7, 5), (16, 8, 6), (4, 4, 1)])
@pytest.mark.parametrize("mode,d", MODES)
def test_blas_and_reductions(n, mode, d):
    """TensorField BLAS-1 fg:9799-10066, innerProductL2 fg:20871-21036, component_dot/average fg:10088-10208"""
    rng = np.random.default_rng(7)
    ctx = fb.Context(*n, mode=mode, gamma_scheme="staggered")
    o = fo.LSSolver(*n, mode=mode)
    a, b, c = (rng.standard_normal((d,) + n) for _ in range(3))
    fa, fb_, fc, fr = ctx.field(a), ctx.field(b), ctx.field(c), ctx.field()
    assert abs(ctx.inner(fa, fb_) - o.innerProduct(a, b)) < 1e-13 * d
    assert abs(ctx.inner(fa, fb_, fc) - o.innerProduct(a, b, c)) < 1e-13 * d
    assert relerr(ctx.average(fa), o.average(a)) < 1e-12
    assert relerr(ctx.component_dot(fa, fa), o.component_norm(a) ** 2) < 1e-13
    ctx.chk(ctx.lib.fgb_xpay(ctx.h, fr, fa, 0.37, fb_))
    assert relerr(ctx.download(fr), a + 0.37 * b) < 1e-15
    ctx.chk(ctx.lib.fgb_xpaymz(ctx.h, fr, fa, -0.21, fb_, fc))
    assert relerr(ctx.download(fr), a + (-0.21) * (b - c)) < 1e-15
    E = rng.standard_normal(d)
    ctx.chk(ctx.lib.fgb_copy(ctx.h, fa, fr))
    ctx.chk(ctx.lib.fgb_adjust_residual(ctx.h, fr, fb.solver._dp(ctx.vec(E)), fb_))
    assert relerr(ctx.download(fr), a + (E.reshape(-1, 1, 1, 1) - b)) < 1e-15
    ctx.chk(ctx.lib.fgb_set_constant(ctx.h, fr, fb.solver._dp(ctx.vec(E))))
    ctx.chk(ctx.lib.fgb_add_constant(ctx.h, fr, fb.solver._dp(ctx.vec(E))))
    assert relerr(ctx.download(fr), np.zeros((d,) + n) + 2 * E.reshape(-1, 1, 1, 1)) < 1e-15
    # fused CG sweep
    import ctypes as C
    delta = C.c_double()
    x0, r0 = a.copy(), b.copy()
    ctx.upload(fr, c)
    fw = ctx.field(rng.standard_normal((d,) + n))
    w = ctx.download(fw)
    ctx.chk(ctx.lib.fgb_cg_update(ctx.h, fa, fb_, fr, fw, 0.61, C.byref(delta)))
    assert relerr(ctx.download(fa), x0 + 0.61 * c) < 1e-15
    r1 = r0 + (-0.61) * (c - w)
    assert relerr(ctx.download(fb_), r1) < 1e-15
    assert abs(delta.value - o.innerProduct(r1, r1)) < 1e-12 * max(1.0, delta.value)
    ctx.close()


@pytest.mark.parametrize("n", [(64, 64, 64), (128, 64, 256), (256, 128, 64), (64, 512, 128), (512, 64, 64), (64, 64, 512),
                               (64, 64, 1024), (1024, 64, 64), (64, 1024, 64)])
def test_pow2_fast_paths(n):
    """register-resident radix-8/16/32 passes (fft_pow2.cuh) on every supported axis length, forward/backward transform
    and the fused x pass with the staggered / collocated Green operators"""
    L = (1.0, 2.0, 1.5)
    rng = np.random.default_rng(11)
    # plain 3-D transform on a 3-component (heat) field
    ctx = fb.Context(*n, *L, mode="heat", gamma_scheme="collocated")
    o = fo.LSSolver(*n, *L, mode="heat", gamma_scheme="collocated")
    x = rng.standard_normal((3,) + n)
    f = ctx.field(x)
    ctx.chk(ctx.lib.fgb_fft_forward(ctx.h, f))
    got = ctx.download_padded(f).reshape(3, n[0], n[1], -1, 2)
    assert relerr(got[..., 0] + 1j * got[..., 1], o.fft(x)) < 5e-14
    ctx.chk(ctx.lib.fgb_fft_backward(ctx.h, f))
    assert relerr(ctx.download(f), x) < 5e-14
    o.set_reference(0.7, 0.0)
    o.setBCProjector(fo.Id4(3))
    E = rng.standard_normal(3)
    ctx.gamma(f, E, 0.7, 0.0, -1.0, 0.5)
    assert relerr(ctx.download(f), o.GammaOperator(E, 0.7, 0.0, x, -1.0, 0.5)) < 2e-12
    ctx.close()
    # 6- and 9-component collocated operators run the three-pass x kernel (fft_pow2_3.cuh) at every power of two
    for mode, d, scheme in (("elasticity", 6, "staggered"), ("heat", 3, "staggered"), ("elasticity", 6, "collocated"),
                            ("hyperelasticity", 9, "collocated")):
        ctx = fb.Context(*n, *L, mode=mode, gamma_scheme=scheme)
        o = fo.LSSolver(*n, *L, mode=mode, gamma_scheme=scheme)
        o.set_reference(1.3, 0.4)
        o.setBCProjector(fo.Id4(d))
        tau = rng.standard_normal((d,) + n)
        E = rng.standard_normal(d)
        f = ctx.field(tau)
        ctx.gamma(f, E, 1.3, 0.4, -1.0, 0.0)
        assert relerr(ctx.download(f), o.GammaOperator(E, 1.3, 0.4, tau, -1.0, 0.0)) < 2e-12
        ctx.close()


def test_general_tiso_and_neohooke2_laws():
    """LinearGeneral fg:11233, LinearTransverselyIsotropic fg:11479 (orientation field, constant axis and the voxel-0 tangent
    quirk fg:11582), NeoHooke2 fg:11867 against the oracle"""
    n = (10, 8, 6)
    rng = np.random.default_rng(21)
    phi = sphere_phi(n, R=0.3, sub=2)
    # --- elasticity: general + tiso
    A = rng.standard_normal((6, 6))
    Cg = A @ A.T + 6 * np.eye(6)
    orient = rng.standard_normal((3,) + n)
    orient /= np.sqrt((orient ** 2).sum(axis=0))
    tparams = [2 * 1.3, 0.8, 0.4, 0.9, 2 * 0.5]
    for axis in (None, (0.6, 0.0, 0.8)):
        ctx = fb.Context(*n, mode="elasticity", gamma_scheme="staggered")
        o = fo.LSSolver(*n, mode="elasticity", gamma_scheme="staggered")
        tp = tparams + (list(axis) if axis else [0.0, 0.0, 0.0])
        ctx.set_phases([1 - phi, phi], [("general", Cg.ravel()), ("tiso", tp)], orientation=orient)
        o.add_phase("m", fo.LinearGeneral(Cg), 1 - phi)
        o.add_phase("f", fo.LinearTransverselyIsotropic(*tparams, a=axis), phi)
        o.set_orientation(orient)
        eps, W = rng.standard_normal((6,) + n), rng.standard_normal((6,) + n)
        fe, fw, fd = ctx.field(eps), ctx.field(W), ctx.field()
        ctx.chk(ctx.lib.fgb_calc_stress(ctx.h, fe, fd, 1.1, 0.3, -1.0))
        assert relerr(ctx.download(fd), o.calcStress(1.1, 0.3, eps, -1.0)) < 1e-12
        ctx.chk(ctx.lib.fgb_calc_stress_deriv(ctx.h, fe, fw, fd, 1.1, 0.3, 1.0))
        assert relerr(ctx.download(fd), o.calcStressDeriv(1.1, 0.3, eps, W, 1.0)) < 1e-12
        assert relerr(ctx.mean_pk1(fe), o.calcMeanStress(eps)) < 1e-12
        lmin, lmax = ctx.ref_material(fe)
        _, olmin, olmax = o.getRefMaterial(eps, False, False)
        assert abs(lmax - olmax) <= 1e-11 * abs(olmax) and abs(lmin - max(olmin, 0.0)) <= 1e-11 * abs(olmax)
        ctx.close()
    # --- hyperelasticity: Neo-Hooke variant 2 + Neo-Hooke
    ctx = fb.Context(*n, mode="hyperelasticity", gamma_scheme="staggered")
    o = fo.LSSolver(*n, mode="hyperelasticity", gamma_scheme="staggered")
    ctx.set_phases([1 - phi, phi], [("nh2", [3.0, 7.0]), ("nh", [10.0, 20.0])])
    o.add_phase("m", fo.NeoHooke2(3.0, 7.0), 1 - phi)
    o.add_phase("f", fo.NeoHooke(10.0, 20.0), phi)
    F = 0.05 * rng.standard_normal((9,) + n)
    F[:3] += 1.0
    W = rng.standard_normal((9,) + n)
    fe, fw, fd = ctx.field(F), ctx.field(W), ctx.field()
    ctx.chk(ctx.lib.fgb_calc_stress(ctx.h, fe, fd, 0.0, 0.0, 1.0))
    assert relerr(ctx.download(fd), o.calcStress(0.0, 0.0, F, 1.0)) < 1e-12
    ctx.chk(ctx.lib.fgb_calc_stress_deriv(ctx.h, fe, fw, fd, 2.0, 0.5, 1.0))
    assert relerr(ctx.download(fd), o.calcStressDeriv(2.0, 0.5, F, W, 1.0)) < 1e-11
    assert abs(ctx.mean_energy(fe) - o.calcMeanEnergy(F)) <= 1e-11 * abs(o.calcMeanEnergy(F))
    ctx.chk(ctx.lib.fgb_check_numeric(ctx.h))
    # a non-positive det F must be flagged, not silently propagated (fg:10293, fg:21202)
    Fbad = F.copy()
    Fbad[0, 0, 0, 0] = -1.0
    ctx.upload(fe, Fbad)
    ctx.chk(ctx.lib.fgb_calc_stress(ctx.h, fe, fd, 0.0, 0.0, 1.0))
    with pytest.raises(fb.FgbError) as e:
        ctx.chk(ctx.lib.fgb_check_numeric(ctx.h))
    assert e.value.code == fb.lib.FGB_ENUMERIC
    ctx.close()


@pytest.mark.parametrize("n", [(16, 12, 10), (9, 7, 5), (4, 6, 300)])
def test_cg_step_explicit_and_implicit_w(n):
    """fgb_cg_step / fgb_cg_update (runCGElasticity fg:23206-23246): the fused sweeps against the oracle's unfused sequence
    p = r + beta p (fg:23245), w = krylovOperator(p) (fg:20583), <p, p-w>, x += a p, r -= a (p-w) (fg:23221, fg:23237), and the
    FGB_W_IMPLICIT form (w never stored) against the explicit one: the same <p, p-w>, x and r up to rounding (the update
    re-evaluates w in another kernel, where the compiler may contract multiply-adds differently)."""
    import ctypes as C
    rng = np.random.default_rng(5)
    phi = sphere_phi(n, R=0.3, sub=2)
    lam1, mu1 = fb.lame(1.0, 0.3)
    lam2, mu2 = fb.lame(25.0, 0.2)
    mu0, lam0 = 3.1, 0.0
    o = fo.LSSolver(*n, mode="elasticity", gamma_scheme="staggered")
    o.add_phase("matrix", fo.LinearIsotropic(mu1, lam1), 1 - phi)
    o.add_phase("sphere", fo.LinearIsotropic(mu2, lam2), phi)
    o.set_reference(mu0, lam0)
    o.setBCProjector(fo.Id4(6))
    r0, p0, x0 = (rng.standard_normal((6,) + n) for _ in range(3))
    beta, a = 0.37, 0.81
    p1 = r0 + beta * p0
    w1 = o.krylovOperator(p1)
    pAp_o = o.innerProduct(p1, p1 - w1)
    x1 = x0 + a * p1
    r1 = r0 + (-a) * (p1 - w1)
    res = {}
    for implicit in (False, True):
        ctx = fb.Context(*n, mode="elasticity", gamma_scheme="staggered")
        ctx.set_phases([1 - phi, phi], [("iso", [mu1, lam1]), ("iso", [mu2, lam2])])
        assert ctx.lib.fgb_cg_implicit_w_supported(ctx.h) == 1
        fr, fp, fp2, fx = ctx.field(r0), ctx.field(p0), ctx.field(), ctx.field(x0)
        fw = -2 if implicit else ctx.field()
        pAp, delta = C.c_double(), C.c_double()
        if implicit:
            # nothing pending yet
            assert ctx.lib.fgb_cg_update(ctx.h, fx, fr, fp2, -2, a, C.byref(delta)) < 0
        ctx.chk(ctx.lib.fgb_cg_step(ctx.h, -1, fr, beta, fp, fp2, fw, mu0, lam0, C.byref(pAp)))
        assert relerr(ctx.download(fp2), p1) < 1e-15
        if not implicit:
            assert relerr(ctx.download(fw), w1) < 1e-11
        assert abs(pAp.value - pAp_o) < 1e-11 * abs(pAp_o)
        ctx.chk(ctx.lib.fgb_cg_update(ctx.h, fx, fr, fp2, fw, a, C.byref(delta)))
        res[implicit] = (pAp.value, delta.value, ctx.download(fx), ctx.download(fr))
        assert relerr(res[implicit][2], x1) < 1e-15
        assert relerr(res[implicit][3], r1) < 1e-11
        assert abs(delta.value - o.innerProduct(r1, r1)) < 1e-11 * delta.value
        if implicit:
            # the implicit result dies with the u buffer
            ctx.gamma(fp, np.zeros(6), mu0, lam0)
            assert ctx.lib.fgb_cg_update(ctx.h, fx, fr, fp2, -2, a, C.byref(delta)) < 0
            # and is refused where the fused path does not apply (BC projector active / direction update missing)
            assert ctx.lib.fgb_cg_step(ctx.h, -1, -1, 0.0, fp2, fp2, -2, mu0, lam0, C.byref(pAp)) < 0
        ctx.close()
    assert abs(res[False][0] - res[True][0]) <= 1e-13 * abs(res[False][0])
    assert np.array_equal(res[False][2], res[True][2])
    assert np.abs(res[False][3] - res[True][3]).max() <= 1e-14 * np.abs(res[False][3]).max()
    assert abs(res[False][1] - res[True][1]) <= 1e-13 * res[False][1]
