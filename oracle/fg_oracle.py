"""FP64 CPU restatement of fibergen's Lippmann-Schwinger solve loop (numpy).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Every function cites the
reference lines it follows; ``fg:N`` means ``/root/reference/src/fibergen.cpp:N``.
Fields are stored *unpadded* as arrays of shape ``(d, nx, ny, nz)`` (the reference's
padding ``k in [nz, nzp)`` carries no information, fg:227-232); ``to_padded`` /
``from_padded`` convert to the reference's in-place r2c layout used at the C ABI.

Component order (fg:9103): 0:11 1:22 2:33 3:23 4:13 5:12 6:32 7:31 8:21.
"""
from __future__ import annotations

import math
import numpy as np
import scipy.fft as sfft

EPS = np.finfo(np.float64).eps
SMALL = np.finfo(np.float64).tiny          # boost::numeric::bounds<T>::smallest()

# stored index -> (row, col), fg:9103 / fg:13182-13183
ROW = np.array([0, 1, 2, 1, 0, 0, 2, 2, 1])
COL = np.array([0, 1, 2, 2, 2, 1, 1, 0, 0])
# (row, col) -> stored index, fg:9262-9276
IDX = np.array([[0, 5, 4], [8, 1, 3], [7, 6, 2]])

FFT_WORKERS = -1


# --------------------------------------------------------------------------- doubly fine grid transfer (fg:14216-14339)
_DFG_S = (np.array([0, 0, 0, 0, 1, 1, 0, 1, 1]), np.array([0, 0, 0, 1, 0, 1, 1, 0, 1]), np.array([0, 0, 0, 1, 1, 0, 1, 1, 0]))


def prolongate_to_dfg(c):
    """fg:14216-14268: f[g](i,j,k) = c[g](((i + s_i) mod 2nx)//2, ((j + s_j) mod 2ny)//2, ((k + s_k) mod 2nz)//2), the shift s depending on
    the component g (the staggered positions of the shear components)"""
    d, nx, ny, nz = c.shape
    f = np.empty((d, 2 * nx, 2 * ny, 2 * nz))
    for g in range(d):
        ii = ((np.arange(2 * nx) + _DFG_S[0][g]) % (2 * nx)) // 2
        jj = ((np.arange(2 * ny) + _DFG_S[1][g]) % (2 * ny)) // 2
        kk = ((np.arange(2 * nz) + _DFG_S[2][g]) % (2 * nz)) // 2
        f[g] = c[g][np.ix_(ii, jj, kk)]
    return f


def restrict_from_dfg(f):
    """fg:14273-14339: c[g](i,j,k) = mean of the 8 fine values at ((2i + a - s_i) mod 2nx, ...), a in {0,1}"""
    d, fnx, fny, fnz = f.shape
    c = np.zeros((d, fnx // 2, fny // 2, fnz // 2))
    for g in range(d):
        acc = 0.0
        # summation order of the reference: (i0,j0,k0) (i1,j0,k0) (i0,j1,k0) (i1,j1,k0) (i0,j0,k1) ...
        for ck in (0, 1):
            for cj in (0, 1):
                for ci in (0, 1):
                    ii = (2 * np.arange(fnx // 2) + ci + fnx - _DFG_S[0][g]) % fnx
                    jj = (2 * np.arange(fny // 2) + cj + fny - _DFG_S[1][g]) % fny
                    kk = (2 * np.arange(fnz // 2) + ck + fnz - _DFG_S[2][g]) % fnz
                    acc = acc + f[g][np.ix_(ii, jj, kk)]
        c[g] = 0.125 * acc
    return c


# --------------------------------------------------------------------------- layout
def nzp_of(nz):
    return 2 * (nz // 2 + 1)


def to_padded(a):
    """(d,nx,ny,nz) -> (d,nx,ny,nzp) reference layout (fg:9557-9569); padding = 0."""
    d, nx, ny, nz = a.shape
    out = np.zeros((d, nx, ny, nzp_of(nz)))
    out[..., :nz] = a
    return out


def from_padded(a, nz):
    return np.ascontiguousarray(a[..., :nz])


# --------------------------------------------------------------------------- Voigt helpers (fg:494-598)
def Id4(dim):
    I = np.eye(dim)
    if dim == 6:
        I[3, 3] = I[4, 4] = I[5, 5] = 0.5
    return I


def II4(dim):
    M = np.zeros((dim, dim))
    M[:3, :3] = 1.0
    return M


def vnorm_2(v):
    v = np.asarray(v, dtype=float)
    if v.size in (3, 9):
        return math.sqrt(float(v @ v))
    return math.sqrt(float(v @ v + v[3] * v[3] + v[4] * v[4] + v[5] * v[5]))


def _w(dim):
    w = np.ones(dim)
    if dim == 6:
        w[3:] = 2.0
    return w


def dyad4_mv(M, v):
    v = np.asarray(v, dtype=float)
    return M @ (v * _w(v.size))


def dyad4_mm(A, B):
    return A @ (B * _w(A.shape[0])[:, None])


def fix_dim(t, dim):
    """fg:12115-12125: extend a d-vector to 9 entries."""
    out = np.zeros((9,) + np.shape(t)[1:])
    out[:dim] = t[:dim]
    if dim == 6:
        out[6], out[7], out[8] = out[3], out[4], out[5]
    return out


def fix_sym(t, dim):
    """fg:12128-12138 on a 9-vector (in place semantic, returns new)."""
    t = np.array(t, dtype=float, copy=True)
    if dim == 6:
        for a, b in ((3, 6), (4, 7), (5, 8)):
            m = 0.5 * (t[a] + t[b])
            t[a] = m
            t[b] = m
    elif dim == 3:
        t[3:] = 0
    return t


def mat33(t9):
    """stored 9-vector(s) (9,...) -> (...,3,3)"""
    t9 = np.asarray(t9)
    return np.moveaxis(t9[IDX], (0, 1), (-2, -1))


def vec9(m):
    """(...,3,3) -> (9,...)"""
    return np.stack([m[..., ROW[i], COL[i]] for i in range(9)], axis=0)


# --------------------------------------------------------------------------- material laws (fg:10287-11993)
class Law:
    dim = 6
    linear = True

    def W(self, F):
        return 0.5 * np.sum(self.PK1(F, 1.0) * F * _w(self.dim).reshape((-1,) + (1,) * (F.ndim - 1)), axis=0) \
            if self.dim == 6 else 0.5 * np.sum(self.PK1(F, 1.0) * F[:self.dim], axis=0)

    def PK1(self, F, alpha):
        raise NotImplementedError

    def dPK1(self, F, alpha, W):
        raise NotImplementedError

    def tangent_rows(self, F):
        """matrix C with row m = dPK1(F; e_m)  (fg:10414-10426 usage); shape (d,d,...)"""
        d = self.dim
        rows = []
        for m in range(d):
            e = np.zeros((d,) + F.shape[1:])
            e[m] = 1.0
            rows.append(self.dPK1(F, 1.0, e))
        return np.stack(rows, axis=0)


class LinearIsotropic(Law):
    """fg:11354-11476"""
    dim = 6

    def __init__(self, mu, lam):
        self.mu, self.lam = float(mu), float(lam)

    def PK1(self, E, alpha):
        two_mu = 2 * alpha * self.mu
        ltr = alpha * self.lam * (E[0] + E[1] + E[2])
        S = np.empty_like(E[:6])
        S[0] = E[0] * two_mu + ltr
        S[1] = E[1] * two_mu + ltr
        S[2] = E[2] * two_mu + ltr
        S[3] = E[3] * two_mu
        S[4] = E[4] * two_mu
        S[5] = E[5] * two_mu
        return S

    def dPK1(self, E, alpha, W):
        return self.PK1(W, alpha)

    def calcPolarization(self, mu_0, F, inv):
        """closed form fg:11427-11467"""
        mu, lam = self.mu, self.lam
        m = 2.0 * (mu + mu_0)
        a = 1.0 / m
        b = lam / (m * (3.0 * lam + m))
        trF = F[0] + F[1] + F[2]
        P = np.empty_like(F)
        P[0] = a * F[0] - b * trF
        P[1] = a * F[1] - b * trF
        P[2] = a * F[2] - b * trF
        P[3:] = a * F[3:]
        if not inv:
            m = 2.0 * (mu - mu_0)
            trP = P[0] + P[1] + P[2]
            Q = np.empty_like(P)
            Q[0] = m * P[0] + lam * trP
            Q[1] = m * P[1] + lam * trP
            Q[2] = m * P[2] + lam * trP
            Q[3:] = m * P[3:]
            P = Q
        return P


class LinearGeneral(Law):
    """fg:11233-11349; C is the 6x6 'tensor Voigt' matrix (default Id4)."""
    dim = 6

    def __init__(self, C):
        self.C = np.array(C, dtype=float).reshape(6, 6)

    def PK1(self, E, alpha):
        C = self.C
        S = np.empty_like(E[:6])
        for i in range(6):
            S[i] = alpha * (E[0] * C[i, 0] + E[1] * C[i, 1] + E[2] * C[i, 2]
                            + 2.0 * (E[3] * C[i, 3] + E[4] * C[i, 4] + E[5] * C[i, 5]))
        return S

    def dPK1(self, E, alpha, W):
        return self.PK1(W, alpha)


class LinearTransverselyIsotropic(Law):
    """fg:11479-11595.  Direction of anisotropy: the constant ``a`` if given (have_a, fg:11516), else the per-voxel
    orientation field.  Quirk kept for parity: dPK1 passes the loop index m (= 0) instead of the voxel
    index to PK1 (fg:11582), so tangents read the orientation of voxel 0."""
    dim = 6

    def __init__(self, two_mu, lam, alpha_t, beta_t, two_dmu, a=None):
        self.two_mu, self.lam, self.alpha_t, self.beta_t, self.two_dmu = map(float, (two_mu, lam, alpha_t, beta_t, two_dmu))
        self.a = None if a is None or not np.any(np.asarray(a) != 0) else np.asarray(a, dtype=float)
        self.orientation = None   # (3,nx,ny,nz), set by solver
        self.sel = None

    def _apply(self, E, alpha, a):
        # A = a (x) a ; P = 2 mu e + (lam tr e + alpha_t a'ea) I + (alpha_t tr e + beta_t a'ea) A + 2 dmu (Ae + eA)
        e = mat33_sym(E)
        A = a_outer(a)
        tr = E[0] + E[1] + E[2]
        av = np.moveaxis(a, 0, -1)
        aea = np.einsum('...i,...ij,...j->...', av, e, av)
        I = np.eye(3)
        al = np.asarray(alpha, dtype=float)
        ex = (lambda x: x[..., None, None]) if (al.ndim or np.ndim(tr)) else (lambda x: x)
        c_I = al * (self.lam * tr + self.alpha_t * aea)
        c_E = al * self.two_mu
        c_A = al * (self.alpha_t * tr + self.beta_t * aea)
        c_AE = al * self.two_dmu
        bc = lambda x: ex(np.broadcast_to(x, np.shape(tr)))
        P = bc(c_E) * e + bc(c_I) * I + bc(c_A) * A + bc(c_AE) * (A @ e + e @ A)
        return np.stack([P[..., ROW[i], COL[i]] for i in range(6)], axis=0)

    def _axis(self, E, voxel0):
        if self.a is not None:
            return self.a.reshape((3,) + (1,) * (E.ndim - 1)) * np.ones((1,) + E.shape[1:])
        o = self.orientation
        if voxel0:
            return o[:, 0, 0, 0].reshape((3,) + (1,) * (E.ndim - 1)) * np.ones((1,) + E.shape[1:])
        if self.sel is not None:
            o = o[:, self.sel]
        return o.reshape((3,) + E.shape[1:])

    def PK1(self, E, alpha):
        return self._apply(E, alpha, self._axis(E, False))

    def dPK1(self, E, alpha, W):
        return self._apply(W, alpha, self._axis(W, True))


def mat33_sym(E6):
    e = np.empty(E6.shape[1:] + (3, 3))
    for i in range(6):
        e[..., ROW[i], COL[i]] = E6[i]
        e[..., COL[i], ROW[i]] = E6[i]
    return e


def a_outer(a):
    av = np.moveaxis(a, 0, -1)
    return av[..., :, None] * av[..., None, :]


class ScalarLinearIsotropic(Law):
    """fg:11161-11228 (heat/porous dim 3, viscosity dim 6)"""

    def __init__(self, mu, dim=3):
        self.mu, self.dim = float(mu), int(dim)

    def W(self, E):
        return 0.5 * np.sum(self.PK1(E, 1.0)[:3] * E[:3], axis=0)   # Tensor3 dot, fg:11176-11180

    def PK1(self, E, alpha):
        return E[:self.dim] * (alpha * self.mu)

    def dPK1(self, E, alpha, W):
        return W[:self.dim] * (alpha * self.mu)


class MatrixLinearAnisotropic(Law):
    """fg:11089-11155 (heat, d=3)"""
    dim = 3

    def __init__(self, c11=1., c22=1., c33=1., c23=0., c13=0., c12=0.):
        self.c = tuple(map(float, (c11, c22, c33, c23, c13, c12)))

    def W(self, E):
        return 0.5 * np.sum(self.PK1(E, 1.0) * E[:3], axis=0)

    def PK1(self, E, alpha):
        c11, c22, c33, c23, c13, c12 = self.c
        S = np.empty_like(E[:3])
        S[0] = alpha * (c11 * E[0] + c12 * E[1] + c13 * E[2])
        S[1] = alpha * (c12 * E[0] + c22 * E[1] + c23 * E[2])
        S[2] = alpha * (c13 * E[0] + c23 * E[1] + c33 * E[2])
        return S

    def dPK1(self, E, alpha, W):
        return self.PK1(W, alpha)


class NeoHooke(Law):
    """fg:11729-11862: S = mu (I - C^-1) + lam ln J C^-1, P = F S."""
    dim = 9
    linear = False

    def __init__(self, mu, lam):
        self.mu, self.lam = float(mu), float(lam)

    def W(self, F):
        Fm = mat33(F)
        trC = np.sum(F * F, axis=0)
        J = np.linalg.det(Fm)
        logJ = np.log(J)
        return 0.5 * (self.mu * ((trC - 3.0) - 2.0 * logJ) + self.lam * logJ * logJ)

    def PK1(self, F, alpha):
        Fm = mat33(F)
        C = np.swapaxes(Fm, -1, -2) @ Fm
        Cinv = np.linalg.inv(C)
        J = np.linalg.det(Fm)
        al = np.asarray(alpha)[..., None, None] if np.ndim(alpha) else alpha
        all_lnJ = (alpha * self.lam * np.log(J))
        all_lnJ = all_lnJ[..., None, None] if np.ndim(all_lnJ) else all_lnJ
        S = (al * self.mu) * (np.eye(3) - Cinv) + all_lnJ * Cinv
        return vec9(Fm @ S)

    def dPK1(self, F, alpha, W):
        Fm = mat33(F)
        Wm = mat33(W)
        J = np.linalg.det(Fm)
        Finv = np.linalg.inv(Fm)
        FinvT = np.swapaxes(Finv, -1, -2)
        WT = np.swapaxes(Wm, -1, -2)
        FinvTWT = FinvT @ WT
        FinvTWTFinvT = FinvTWT @ FinvT
        tr = np.trace(FinvTWT, axis1=-2, axis2=-1)
        a = np.asarray(alpha)
        ex = (lambda x: x[..., None, None]) if (a.ndim or np.ndim(J)) else (lambda x: x)
        c_mu = a * self.mu
        c_tr = a * self.lam * tr
        c_m = a * (self.mu - self.lam * np.log(J))
        c_mu = np.broadcast_to(c_mu, tr.shape)
        dP = ex(c_mu) * Wm + ex(c_tr) * FinvT + ex(np.broadcast_to(c_m, tr.shape)) * FinvTWTFinvT
        return vec9(dP)


class NeoHooke2(Law):
    """fg:11867-11993: W = 1/2 [mu (J^-2/3 tr C - 3) + K (J-1)^2]"""
    dim = 9
    linear = False

    def __init__(self, mu, K):
        self.mu, self.K = float(mu), float(K)

    def W(self, F):
        J = np.linalg.det(mat33(F))
        return 0.5 * (self.mu * (J ** (-2.0 / 3.0) * np.sum(F * F, axis=0) - 3) + self.K * (J - 1) ** 2)

    def PK1(self, F, alpha):
        Fm = mat33(F)
        FinvT = np.swapaxes(np.linalg.inv(Fm), -1, -2)
        trC = np.sum(F * F, axis=0)
        J = np.linalg.det(Fm)
        al = np.asarray(alpha, dtype=float)
        muJ23 = al * self.mu * J ** (-2.0 / 3.0)
        D = al * self.K * J * (J - 1) - muJ23 * (1.0 / 3.0) * trC
        ex = (lambda x: np.broadcast_to(x, J.shape)[..., None, None])
        return vec9(ex(muJ23) * Fm + ex(D) * FinvT)

    def dPK1(self, F, alpha, W):
        Fm, Wm = mat33(F), mat33(W)
        Finv = np.linalg.inv(Fm)
        FinvT = np.swapaxes(Finv, -1, -2)
        J = np.linalg.det(Fm)
        trC3 = (1.0 / 3.0) * np.sum(F * F, axis=0)
        al = np.asarray(alpha, dtype=float)
        a_muJ23 = al * self.mu * J ** (-2.0 / 3.0)
        a_KJ = al * self.K * J
        a_KJJ1 = a_KJ * (J - 1)
        a_KJJ = a_KJ * (2 * J - 1)
        A = FinvT @ np.swapaxes(Wm, -1, -2)
        B = A @ FinvT
        tr = np.trace(A, axis1=-2, axis2=-1)
        FW23 = (2.0 / 3.0) * np.sum(F * W, axis=0)
        FiTW = np.sum(FinvT * Wm, axis=(-1, -2))
        ex = (lambda x: np.broadcast_to(x, J.shape)[..., None, None])
        dP = ex(a_muJ23) * (ex(-2.0 / 3.0 * tr) * (Fm - ex(trC3) * FinvT) + Wm - ex(FW23) * FinvT + ex(trC3) * B) \
            + ex(a_KJJ * FiTW) * FinvT - ex(a_KJJ1) * B
        return vec9(dP)


class SaintVenantKirchhoff(Law):
    """fg:11598-11725"""
    dim = 9
    linear = False

    def __init__(self, mu, lam):
        self.mu, self.lam = float(mu), float(lam)

    def _S(self, Fm, alpha):
        E = 0.5 * (np.swapaxes(Fm, -1, -2) @ Fm - np.eye(3))
        trE = np.trace(E, axis1=-2, axis2=-1)
        al = np.asarray(alpha)
        two_mu = 2 * al * self.mu
        ltr = al * self.lam * trE
        if np.ndim(two_mu):
            two_mu = two_mu[..., None, None]
        return two_mu * E + ltr[..., None, None] * np.eye(3)

    def W(self, F):
        Fm = mat33(F)
        E = 0.5 * (np.swapaxes(Fm, -1, -2) @ Fm - np.eye(3))
        trE = np.trace(E, axis1=-2, axis2=-1)
        return 0.5 * self.lam * trE * trE + self.mu * np.sum(E * E, axis=(-1, -2))

    def PK1(self, F, alpha):
        Fm = mat33(F)
        return vec9(Fm @ self._S(Fm, alpha))

    def dPK1(self, F, alpha, W):
        Fm, Wm = mat33(F), mat33(W)
        S = self._S(Fm, alpha)
        dE = 0.5 * (np.swapaxes(Wm, -1, -2) @ Fm + np.swapaxes(Fm, -1, -2) @ Wm)
        trdE = np.trace(dE, axis1=-2, axis2=-1)
        al = np.asarray(alpha)
        two_mu = 2 * al * self.mu
        if np.ndim(two_mu):
            two_mu = two_mu[..., None, None]
        dS = two_mu * dE + (al * self.lam * trdE)[..., None, None] * np.eye(3)
        return vec9(Fm @ dS + Wm @ S)


# --------------------------------------------------------------------------- mixing rules (fg:12004-13732)
class Phase:
    def __init__(self, name, law, phi=None):
        self.name, self.law, self.phi = name, law, phi


class MixedBase:
    def __init__(self, dim):
        self.dim = dim
        self.phases = []

    # d x d tangent "rows" matrix per voxel: C[m] = dPK1(F; e_m), shape (d,d,...)
    def tangent_rows(self, F, sel=None):
        d = self.dim
        rows = []
        for m in range(d):
            e = np.zeros((d,) + F.shape[1:])
            e[m] = 1.0
            rows.append(self.dPK1(F, 1.0, e, sel=sel))
        return np.stack(rows, axis=0)

    def calcPolarization(self, mu_0, F, inv):
        """fg:12087-12099 dispatcher + generic fg:10414-10445."""
        d = self.dim
        shp = F.shape[1:]
        out = np.empty_like(F)
        done = np.zeros(shp, dtype=bool)
        for ph in self.phases:
            pure = (ph.phi == 1) & ~done
            if pure.any() and hasattr(ph.law, "calcPolarization"):
                out[:, pure] = ph.law.calcPolarization(mu_0, F[:, pure], inv)
                done |= pure
        rest = ~done
        if rest.any():
            Fr = F[:, rest]
            C = self.tangent_rows(Fr, sel=rest)                    # (d,d,n) rows
            Cm = np.moveaxis(C, -1, 0)                              # (n,d,d) row-major as in ublas
            C2 = Cm + 2.0 * mu_0 * np.eye(d)
            # lapack::gesv on the row-major storage solves C2^T x = F (column-major view)
            Pm = np.linalg.solve(np.swapaxes(C2, -1, -2), np.moveaxis(Fr, 0, -1)[..., None])[..., 0]
            if not inv:
                Pm = (Cm @ Pm[..., None])[..., 0] - 2.0 * mu_0 * Pm
            out[:, rest] = np.moveaxis(Pm, -1, 0)
        return out


class VoigtMixed(MixedBase):
    """fg:12729-12780"""
    threshold = 10 * EPS

    def PK1(self, F, alpha, sel=None):
        P = None
        for ph in self.phases:
            phi = ph.phi if sel is None else ph.phi[sel]
            use = phi > self.threshold
            if not use.any():
                continue
            if hasattr(ph.law, "sel"):
                ph.law.sel = sel
            c = ph.law.PK1(F, phi * alpha)
            c = np.where(use, c, 0.0)
            P = c if P is None else P + c
        return P

    def dPK1(self, F, alpha, W, sel=None):
        P = None
        for ph in self.phases:
            phi = ph.phi if sel is None else ph.phi[sel]
            use = phi > self.threshold
            if not use.any():
                continue
            c = ph.law.dPK1(F, phi * alpha, W)
            c = np.where(use, c, 0.0)
            P = c if P is None else P + c
        return P

    def W(self, F):
        W = 0.0
        for ph in self.phases:
            use = ph.phi > self.threshold
            W = W + np.where(use, ph.phi * ph.law.W(F), 0.0)
        return W


class ReussMixed(MixedBase):
    """fg:12653-12724 (InvertMatrix = LU inverse, fg:1143)."""

    def _Ceff(self, F, sel):
        d = self.dim
        n = F[0].size
        Ff = F.reshape(d, n)
        Sum = np.zeros((n, d, d))
        pure_law = np.full(n, -1)
        for ip, ph in enumerate(self.phases):
            phi = (ph.phi if sel is None else ph.phi[sel]).reshape(n)
            pure_law[(phi == 1) & (pure_law < 0)] = ip
            mixed = (phi != 0) & (phi != 1)
            if mixed.any():
                C = np.moveaxis(ph.law.tangent_rows(Ff[:, mixed]), -1, 0)
                Sum[mixed] += phi[mixed, None, None] * np.linalg.inv(C)
        return Sum, pure_law

    def _apply(self, F, alpha, X, sel, deriv):
        d = self.dim
        shp = F.shape[1:]
        n = F[0].size
        Ff, Xf = F.reshape(d, n), X.reshape(d, n)
        Sum, pure_law = self._Ceff(F, sel)
        out = np.empty((d, n))
        for ip, ph in enumerate(self.phases):
            m = pure_law == ip
            if m.any():
                out[:, m] = ph.law.dPK1(Ff[:, m], alpha, Xf[:, m]) if deriv else ph.law.PK1(Ff[:, m], alpha)
        m = pure_law < 0
        if m.any():
            C = np.linalg.inv(Sum[m])
            out[:, m] = alpha * np.einsum('nkj,jn->kn', C, Xf[:, m])
        return out.reshape((d,) + shp)

    def PK1(self, F, alpha, sel=None):
        return self._apply(F, alpha, F, sel, False)

    def dPK1(self, F, alpha, W, sel=None):
        return self._apply(F, alpha, W, sel, True)


class LaminateMixed(MixedBase):
    """fg:13086-13732 with tangent="approx" (default).  Linear modes (d=3,6) take
    exactly one full Newton step from a=0 (fg:13367-13370); d=9 runs the projected,
    backtracked Newton iteration per interface voxel (python loop: small cases only)."""

    def __init__(self, dim, normals=None):
        super().__init__(dim)
        self.normals = normals      # (3,nx,ny,nz)
        self.eps_t = 4 * EPS
        self.eps_a = EPS ** (2.0 / 3.0)
        self.eps_g = EPS
        self.alpha = 0.001
        self.beta = 0.1
        self.delta = 1 - 1024 * EPS
        self.maxiter = 32
        self.fixed_c1 = -1.0
        self.backtrack = True
        self.project_t = True

    # --- phase selection (fg:13456-13525) ------------------------------------
    def _classify(self, sel):
        phis = [ph.phi if sel is None else ph.phi[sel] for ph in self.phases]
        shp = phis[0].shape
        n = phis[0].size
        phis = [p.reshape(n) for p in phis]
        p1 = np.full(n, -1)
        p2 = np.full(n, -1)
        c1 = np.zeros(n)
        c2 = np.zeros(n)
        pure = np.zeros(n, dtype=bool)
        for ip, phi in enumerate(phis):
            active = ~pure
            nz = active & (phi != 0)
            one = nz & (phi == 1)
            # phi == 1: becomes (only) phase 1, stop
            p1[one] = ip
            c1[one] = 1.0
            p2[one] = -1
            pure |= one
            rest = nz & ~one
            first = rest & (p1 < 0)
            p1[first] = ip
            c1[first] = phi[first]
            second = rest & ~first & (p2 < 0)
            p2[second] = ip
            c2[second] = phi[second]
            third = rest & ~first & ~second
            if third.any():
                raise RuntimeError("The laminate mixing rule supports only two phase mixtures")
        return p1, p2, c1, c2, shp, n

    def _split(self, F, sel):
        """returns list of (mask, law1, law2, c1, c2, F1, F2) groups over flattened voxels"""
        d = self.dim
        p1, p2, c1, c2, shp, n = self._classify(sel)
        Ff = F.reshape(d, n)
        nrm = (self.normals if sel is None else self.normals[:, sel]).reshape(3, n)
        groups = []
        for i1 in range(len(self.phases)):
            m = (p1 == i1) & (p2 < 0)
            if m.any():
                groups.append((m, self.phases[i1].law, None, c1[m], None, Ff[:, m], None))
            for i2 in range(len(self.phases)):
                m = (p1 == i1) & (p2 == i2)
                if not m.any():
                    continue
                cc1 = c1[m].copy()
                if self.fixed_c1 > 0:
                    cc1[:] = self.fixed_c1
                cc2 = 1.0 - cc1
                F1, F2 = self.solve_newton(self.phases[i1].law, self.phases[i2].law, cc1, cc2, nrm[:, m], Ff[:, m])
                groups.append((m, self.phases[i1].law, self.phases[i2].law, cc1, cc2, F1, F2))
        return groups, shp, n

    # --- Newton for the jump vector (fg:13157-13454) -------------------------
    def solve_newton(self, law1, law2, c1, c2, n, Fbar_d):
        d = self.dim
        nv = c1.size
        Fbar = fix_dim(Fbar_d, d)                       # (9,nv)
        if d != 9:
            return self._newton_linear(law1, law2, c1, c2, n, Fbar)
        F1 = np.empty((9, nv))
        F2 = np.empty((9, nv))
        for v in range(nv):
            f1, f2 = self._newton_hyper(law1, law2, c1[v], c2[v], n[:, v], Fbar[:, v])
            F1[:, v], F2[:, v] = f1, f2
        return F1, F2

    def _dFda(self, c1, c2, n):
        """dF1/da_k, dF2/da_k as (3, 9, nv) (fg:13229-13243)"""
        d = self.dim
        nv = np.size(c1)
        dF1 = np.zeros((3, 9, nv))
        dF2 = np.zeros((3, 9, nv))
        for k in range(3):
            for i in range(9):
                if ROW[i] == k:
                    dF1[k, i] = -c2 * n[COL[i]]
                    dF2[k, i] = c1 * n[COL[i]]
            dF1[k] = fix_sym(dF1[k], d)
            dF2[k] = fix_sym(dF2[k], d)
        return dF1, dF2

    @staticmethod
    def _dot9(a, b):
        return np.sum(a * b, axis=0)

    def _gH(self, law1, law2, c1, c2, F1, F2, dF1, dF2):
        d = self.dim
        P1 = fix_dim(law1.PK1(F1[:d], 1.0), d)
        P2 = fix_dim(law2.PK1(F2[:d], 1.0), d)
        g = np.stack([c1 * self._dot9(P1, dF1[k]) + c2 * self._dot9(P2, dF2[k]) for k in range(3)])
        H = []
        for i in range(6):
            k, l = ROW[i], COL[i]
            dP1 = fix_dim(law1.dPK1(F1[:d], 1.0, dF1[l][:d]), d)
            dP2 = fix_dim(law2.dPK1(F2[:d], 1.0, dF2[l][:d]), d)
            H.append(c1 * self._dot9(dP1, dF1[k]) + c2 * self._dot9(dP2, dF2[k]))
        return g, np.stack(H)

    @staticmethod
    def _sym_inv(H):
        """SymTensor3x3::inv (fg:9375-9392 adjugate/det) on (6,nv)"""
        a, b, c, d_, e, f = H[0], H[1], H[2], H[3], H[4], H[5]   # 11 22 33 23 13 12
        det = a * (b * c - d_ * d_) - f * (f * c - d_ * e) + e * (f * d_ - b * e)
        with np.errstate(divide='ignore', invalid='ignore'):
            inv = np.stack([(b * c - d_ * d_), (a * c - e * e), (a * b - f * f),
                            (e * f - a * d_), (f * d_ - e * b), (e * d_ - f * c)]) / det
        return inv

    @staticmethod
    def _sym_mv(S, v):
        return np.stack([S[0] * v[0] + S[5] * v[1] + S[4] * v[2],
                         S[5] * v[0] + S[1] * v[1] + S[3] * v[2],
                         S[4] * v[0] + S[3] * v[1] + S[2] * v[2]])

    def _F12(self, Fbar, c1, c2, a, n):
        d = self.dim
        F1 = Fbar.copy()
        F2 = Fbar.copy()
        for i in range(9):
            F1[i] = F1[i] - c2 * a[ROW[i]] * n[COL[i]]
            F2[i] = F2[i] + c1 * a[ROW[i]] * n[COL[i]]
        return fix_sym(F1, d), fix_sym(F2, d)

    def _newton_linear(self, law1, law2, c1, c2, n, Fbar):
        nv = c1.size
        dF1, dF2 = self._dFda(c1, c2, n)
        F1 = Fbar.copy()
        F2 = Fbar.copy()
        g, H = self._gH(law1, law2, c1, c2, F1, F2, dF1, dF2)
        g_norm = np.sqrt(np.sum(g * g, axis=0))
        Hinvg = self._sym_mv(self._sym_inv(H), g)
        da_norm = np.sqrt(np.sum(Hinvg * Hinvg, axis=0))
        a_next = -Hinvg
        G1, G2 = self._F12(Fbar, c1, c2, a_next, n)
        # voxels that "converged" before the step keep F1=F2=Fbar (fg:13262, fg:13306)
        stop = (g_norm <= self.eps_g) | (da_norm <= self.eps_a)
        F1 = np.where(stop, F1, G1)
        F2 = np.where(stop, F2, G2)
        return F1, F2

    def _newton_hyper(self, law1, law2, c1, c2, n, Fbar):
        ex = lambda x: np.asarray(x, dtype=float).reshape(-1, 1)
        c1a, c2a = np.array([c1]), np.array([c2])
        na = ex(n)
        Fb = ex(Fbar)
        Fbarinv = np.linalg.inv(mat33(Fbar))
        a = np.zeros((3, 1))
        F1, F2 = Fb.copy(), Fb.copy()
        Wv = c1 * law1.W(F1)[0] + c2 * law2.W(F2)[0]
        dF1, dF2 = self._dFda(c1a, c2a, na)
        it = 0
        while True:
            g, H = self._gH(law1, law2, c1a, c2a, F1, F2, dF1, dF2)
            if math.sqrt(float(np.sum(g * g))) <= self.eps_g:
                break
            Hinvg = self._sym_mv(self._sym_inv(H), g)
            da = Hinvg
            gTda = float(np.sum(da * g))
            if math.sqrt(float(np.sum(Hinvg * Hinvg))) <= self.eps_a:
                break
            t = 1.0
            if self.project_t:
                w = float((Fbarinv @ Hinvg[:, 0]) @ n)
                x = float((Fbarinv @ a[:, 0]) @ n)
                if w > 0:
                    t = min(1.0, (x + self.delta / c1) / w)
                elif w < 0:
                    t = min(1.0, (x - self.delta / c2) / w)
            while True:
                a_next = a - t * da
                F1, F2 = self._F12(Fb, c1a, c2a, a_next, na)
                W_next = c1 * law1.W(F1)[0] + c2 * law2.W(F2)[0]
                if not self.backtrack:
                    break
                if W_next < Wv - self.alpha * t * gTda:
                    break
                t *= self.beta
                if t <= self.eps_t:
                    break
            if t <= self.eps_t:
                break
            a, Wv = a_next, W_next
            if it >= self.maxiter:
                break
            it += 1
        return F1[:, 0], F2[:, 0]

    # --- law interface ---------------------------------------------------------
    def PK1(self, F, alpha, sel=None):
        d = self.dim
        groups, shp, n = self._split(F, sel)
        out = np.empty((d, n))
        for m, l1, l2, c1, c2, F1, F2 in groups:
            P = l1.PK1(F1[:d], c1 * alpha)
            if l2 is not None:
                P = P + l2.PK1(F2[:d], c2 * alpha)
            out[:, m] = P
        return out.reshape((d,) + shp)

    def dPK1(self, F, alpha, W, sel=None):
        d = self.dim
        groups, shp, n = self._split(F, sel)
        Wf = W.reshape(d, n)
        out = np.empty((d, n))
        for m, l1, l2, c1, c2, F1, F2 in groups:
            P = l1.dPK1(F1[:d], c1 * alpha, Wf[:, m])
            if l2 is not None:
                P = P + l2.dPK1(F2[:d], c2 * alpha, Wf[:, m])
            out[:, m] = P
        return out.reshape((d,) + shp)

    def W(self, F):
        d = self.dim
        groups, shp, n = self._split(F, None)
        out = np.empty(n)
        for m, l1, l2, c1, c2, F1, F2 in groups:
            w = c1 * l1.W(F1[:d] if d != 9 else F1)
            if l2 is not None:
                w = w + c2 * l2.W(F2[:d] if d != 9 else F2)
            out[m] = w
        return out.reshape(shp)


# --------------------------------------------------------------------------- error estimators (fg:14344-14637)
class _EE:
    abs_err = math.inf
    rel_err = 1.0

    def update(self):
        raise RuntimeError("Selected error estimator is not compatible with the selected solution method")

    def update_cg(self, gamma, gamma0):
        raise RuntimeError("Selected error estimator is not compatible with the selected solution method")


class NoneEE(_EE):
    abs_err = 1.0

    def update(self):
        pass

    def update_cg(self, gamma, gamma0):
        pass


class ResidualEE(_EE):
    def update_cg(self, gamma, gamma0):
        self.abs_err = math.sqrt(gamma)
        self.rel_err = math.sqrt(gamma / gamma0)


class EpsilonEE(_EE):
    def __init__(self, solver):
        self.s = solver
        self.prev = self._mean()

    def _mean(self):
        return fix_dim(self.s.component_norm(self.s.epsilon), self.s.dim)

    def update(self):
        cur = self._mean()
        self.abs_err = abs(np.linalg.norm(self.prev) - np.linalg.norm(cur))
        self.rel_err = self.abs_err / (SMALL + np.linalg.norm(cur))
        self.prev = cur

    def update_cg(self, gamma, gamma0):
        self.update()


class SigmaEE(_EE):
    def __init__(self, solver):
        self.s = solver
        self.cur = self._mean()
        self.prev = self.cur.copy()
        self.prev_prev = self.cur.copy()
        self.it = 0

    def _mean(self):
        return fix_dim(self.s.calcMeanStress(), self.s.dim)

    def update(self):
        self.cur = self._mean()
        if self.it > 1:
            self.abs_err = 0.5 * (np.linalg.norm(self.prev_prev - self.cur) + np.linalg.norm(self.prev - self.cur))
        else:
            self.abs_err = np.linalg.norm(self.prev - self.cur)
        self.rel_err = self.abs_err / (SMALL + np.linalg.norm(self.cur))
        self.prev_prev = self.prev.copy()
        self.prev = self.cur.copy()
        self.it += 1

    def update_cg(self, gamma, gamma0):
        self.update()


class EnergyEE(_EE):
    def __init__(self, solver):
        self.s = solver
        self.prev = self.s.calcMeanEnergy()

    def update(self):
        cur = self.s.calcMeanEnergy()
        self.abs_err = abs(self.prev - cur)
        self.rel_err = self.abs_err / (SMALL + abs(cur))
        self.prev = cur

    def update_cg(self, gamma, gamma0):
        self.update()


# --------------------------------------------------------------------------- the solver
class LSSolver:
    """Restatement of ``LSSolver<double,double,3>`` (fg:14643-24736), hot path only."""

    def __init__(self, nx, ny, nz, dx=1.0, dy=1.0, dz=1.0, mode="elasticity", method="cg",
                 gamma_scheme="auto", mixing_rule="voigt", error_estimator="epsilon",
                 outer_error_estimator="epsilon", tol=1e-4, abs_tol=EPS, bc_tol=1e-3, maxiter=10000,
                 ref_scale=1.0, bc_relax=1.0, newton_relax=1.0, update_ref="loadstep", freq_hack=False,
                 cg_reinit=0, loadsteps=1, loadstep_extrapolation_order=0):
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.dx, self.dy, self.dz = float(dx), float(dy), float(dz)
        self.nzc = self.nz // 2 + 1
        self.nzp = 2 * self.nzc
        self.nxyz = self.nx * self.ny * self.nz
        self.mode, self.method = mode, method
        # fg:15066-15079
        if gamma_scheme == "auto":
            gamma_scheme = "collocated" if method == "polarization" else "staggered"
        if method == "polarization":
            gamma_scheme = "collocated"
        if gamma_scheme in ("half-staggered", "full-staggered"):
            gamma_scheme = gamma_scheme.replace("-", "_")
        # use_dfg fg:14894: half_staggered / full_staggered evaluate the material on the doubly fine grid; the operators are the
        # staggered ones (fg:20480, fg:20496)
        self.dfg = {"half_staggered": 1, "full_staggered": 2}.get(gamma_scheme, 0)
        if self.dfg:
            gamma_scheme = "staggered"
        self.gamma_scheme = gamma_scheme
        self.error_estimator = error_estimator
        self.outer_error_estimator = outer_error_estimator
        self.tol, self.abs_tol, self.bc_tol, self.maxiter = tol, abs_tol, bc_tol, maxiter
        self.ref_scale, self.bc_relax, self.newton_relax = ref_scale, bc_relax, newton_relax
        self.update_ref, self.freq_hack, self.cg_reinit = update_ref, freq_hack, cg_reinit
        self.dim = {"elasticity": 6, "viscosity": 6, "heat": 3, "porous": 3, "hyperelasticity": 9}[mode]
        d = self.dim
        self.mixing_rule = mixing_rule
        self.normals = None
        self.orientation = None
        if mixing_rule == "voigt":
            self.mat = VoigtMixed(d)
        elif mixing_rule == "reuss":
            self.mat = ReussMixed(d)
        elif mixing_rule == "laminate":
            self.mat = LaminateMixed(d)
        else:
            raise RuntimeError("Unknown material mixing rule '%s'" % mixing_rule)
        self.mu_0 = math.nan
        self.lambda_0 = 0.0
        self.reference_set = False
        self.loadsteps = [i / loadsteps for i in range(loadsteps + 1)]
        self.loadstep_extrapolation_order = loadstep_extrapolation_order
        self.E = np.zeros(d)
        self.S = np.zeros(d)
        self.Id = np.zeros(d)
        self.Id[:3] = 1
        self.F0 = np.zeros(d)
        self.F00 = np.zeros(d)
        self.epsilon = np.zeros((d, self.nx, self.ny, self.nz))
        self.residuals = []
        self.callback = None
        self.n_operator_applications = 0
        self.setBCProjector(Id4(d), check=False)

    # -- setup ----------------------------------------------------------------
    def add_phase(self, name, law, phi):
        """phi on the solver grid; full_staggered: on the doubly fine grid (fg:17154-17156); half_staggered: on the coarse grid, continued
        piecewise constant to the fine grid (initFullStageredRawPhases fg:17648-17680)"""
        phi = np.ascontiguousarray(phi, dtype=float)
        if self.dfg == 1:
            phi = np.repeat(np.repeat(np.repeat(phi, 2, axis=0), 2, axis=1), 2, axis=2)
        self.mat.phases.append(Phase(name, law, phi))

    def _mat_in(self, eps):
        return prolongate_to_dfg(eps) if self.dfg else eps

    def _mat_out(self, tau):
        return restrict_from_dfg(tau) if self.dfg else tau

    @property
    def nxyz_mat(self):
        return self.nxyz * (8 if self.dfg else 1)

    def set_reference(self, mu, lam):
        self.mu_0, self.lambda_0, self.reference_set = float(mu), float(lam), True

    def set_normals(self, n):
        self.normals = np.ascontiguousarray(n, dtype=float)
        if isinstance(self.mat, LaminateMixed):
            self.mat.normals = self.normals

    def set_orientation(self, a):
        self.orientation = np.ascontiguousarray(a, dtype=float)
        for ph in self.mat.phases:
            if isinstance(ph.law, LinearTransverselyIsotropic):
                ph.law.orientation = self.orientation

    def setStrain(self, E):
        self.E = np.array(E, dtype=float)

    def setStress(self, S):
        self.S = np.array(S, dtype=float)

    # -- BC projector (fg:20599-20665) ---------------------------------------------
    def setBCProjector(self, P, check=True):
        P = np.array(P, dtype=float)
        dim = P.shape[0]
        eps = math.sqrt(EPS)
        if check:
            if np.linalg.norm(P - P.T) > eps:
                raise RuntimeError("Projector is not symmetric")
            if np.linalg.norm(P - dyad4_mm(P, P)) > eps:
                raise RuntimeError("Specified Projector is not a projector")
        C0 = 2 * self.mu_0 * Id4(dim) + self.lambda_0 * II4(dim)
        self.BC_P = P
        self.BC_Q = Id4(dim) - P
        self.BC_QC0 = dyad4_mm(self.BC_Q, C0)
        QC0Q = dyad4_mm(self.BC_QC0, self.BC_Q)
        edim = 9 if dim == 6 else dim
        if dim == 6:
            A = np.zeros((9, 9))
            for i in range(9):
                for j in range(i, 9):
                    A[j, i] = A[i, j] = QC0Q[i if i < 6 else i - 3, j if j < 6 else j - 3]
            QC0Q = A
        if np.all(np.isfinite(QC0Q)):
            U, s, VT = np.linalg.svd(QC0Q)
            cut = math.sqrt(EPS) * np.linalg.norm(s)
            Sinv = np.zeros((edim, edim))
            for i in range(edim):
                if abs(s[i]) > cut:
                    Sinv[i, i] = 1.0 / s[i]
            M = VT.T @ Sinv @ U.T
        else:
            M = np.full((edim, edim), math.nan)
        if dim == 6:
            M = M.copy()
            for i in range(3):
                for j in range(6):
                    M[j, 3 + i] = 0.5 * (M[j, 3 + i] + M[j, 6 + i])
                    M[3 + i, j] = 0.5 * (M[3 + i, j] + M[6 + i, j])
            M = M[:6, :6].copy()
        self.BC_M = M
        self.BC_MQ = dyad4_mm(M, self.BC_Q)

    def calcBCMean(self, E, S):
        return E + self.bc_relax * dyad4_mv(self.BC_M, S - dyad4_mv(self.BC_QC0, E))       # fg:20242

    def calcBCProjector(self):
        return self.bc_relax * dyad4_mv(self.BC_MQ, self.F0) \
            - (1 - self.bc_relax) * dyad4_mv(self.BC_M, dyad4_mv(self.BC_QC0, self.F00))   # fg:20258

    def _initBCProjector_real(self, tau):
        if np.linalg.norm(self.BC_MQ) < EPS:                                                # fg:20233
            self.F0 = np.zeros(self.dim)
            return
        self.F0 = self.average(tau)

    # -- reductions (fg:10088-10208, fg:20871-21036) ---------------------------------
    def average(self, a):
        return a.reshape(a.shape[0], -1).sum(axis=1) / self.nxyz

    def component_norm(self, a):
        return np.sqrt((a * a).reshape(a.shape[0], -1).sum(axis=1) / self.nxyz)

    def innerProduct(self, a, b, c=None):
        y = b if c is None else b - c
        if self.dim == 6:
            s = np.sum(a[0] * y[0] + a[1] * y[1] + a[2] * y[2] + 2 * (a[3] * y[3] + a[4] * y[4] + a[5] * y[5]))
        else:
            s = np.sum(a * y)
        return float(s) / self.nxyz

    # -- constitutive sweeps (fg:18134-18184, fg:18425-18478, fg:18044-18118) ---------
    def calcStress(self, mu_0, lambda_0, eps, alpha=1.0):
        beta = -alpha * 2 * mu_0
        gamma = -alpha * lambda_0
        eps = self._mat_in(eps)                                       # fg:18143-18149
        P = self.mat.PK1(eps, alpha)
        if beta != 0:
            P = P + beta * eps
        if gamma != 0:
            trF = eps[0] + eps[1] + eps[2]
            P[:3] = P[:3] + gamma * trF
        return self._mat_out(P)                                       # fg:18343-18347

    def calcStressDiff(self, eps, alpha=1.0):
        return self.calcStress(self.mu_0, self.lambda_0, eps, alpha)

    def calcStressDeriv(self, mu_0, lambda_0, F, W, alpha=1.0):
        beta = -alpha * 2 * mu_0
        gamma = -alpha * lambda_0
        F, W = self._mat_in(F), self._mat_in(W)                       # fg:18435-18441
        dP = self.mat.dPK1(F, alpha, W)
        if beta != 0:
            dP = dP + beta * W
        if gamma != 0:
            trW = W[0] + W[1] + W[2]
            dP[:3] = dP[:3] + gamma * trW
        return self._mat_out(dP)

    def calcStressConst(self, mu_0, lambda_0, eps):
        """fg:17973-18020: tau = C0:eps"""
        two_mu = 2 * mu_0
        ltr = lambda_0 * (eps[0] + eps[1] + eps[2])
        tau = eps * two_mu
        tau[:3] = tau[:3] + ltr
        return tau

    def calcPolarization(self, mu_0, eps, inv=False):
        return self._mat_out(self.mat.calcPolarization(mu_0, self._mat_in(eps), inv))      # fg:18051, fg:18113

    def calcMeanStress(self, eps=None):
        eps = self._mat_in(self.epsilon if eps is None else eps)                             # fg:17797-17803
        P = self.mat.PK1(eps, 1.0 / self.nxyz_mat)                                           # fg:12318 alpha /= nxyz
        return P.reshape(self.dim, -1).sum(axis=1)

    def calcMeanEnergy(self, eps=None):
        eps = self._mat_in(self.epsilon if eps is None else eps)                             # fg:17769-17774
        return float(np.sum(self.mat.W(eps))) / self.nxyz_mat

    def calcMeanCauchyStress(self, eps=None):
        """fg:17920-17941 -> meanCauchy fg:12268-12308 -> Cauchy fg:10326-10346: per voxel sigma = P(F) F^T / det F with the mixed
        law's PK1, averaged over the voxels (hyperelasticity, 9 components)"""
        eps = self.epsilon if eps is None else eps
        if self.dim != 9:
            raise RuntimeError("oracle: Cauchy stress needs the 9-component deformation gradient")
        eps = self._mat_in(eps)                                         # fg:17926-17932
        F = mat33(eps.reshape(9, -1))                                   # (n, 3, 3)
        c = 1.0 / np.linalg.det(F)
        P = mat33(self.mat.PK1(eps, 1.0).reshape(9, -1)) * (c / self.nxyz_mat)[:, None, None]
        return vec9(np.einsum('nik,njk->nij', P, F)).sum(axis=1)

    def calcDisplacement(self, eps=None):
        """get_raw_field('u') fg:15509-15557: the displacement fluctuation of the converged field with alpha = 1.
        elasticity / heat: tau = C0:eps (calcStressConst fg:17973), staggered div_h and G0 (fg:15519-15521, fg:15539-15542);
        viscosity: tau = calcStressDiff, staggered operators with the dual reference 1/(4 mu0), lambda0 = inf, alpha = 1/(2 mu0)
        (fg:15530-15537); hyperelasticity: tau = calcStressDiff, then G0DivOperatorHyper -- the COLLOCATED Fourier divergence
        and G0 with xi = 2 pi m / L (fg:15524-15527 -> fg:20281 -> fg:20155)."""
        eps = self.epsilon if eps is None else eps
        if self.mode == "hyperelasticity":
            return self.G0DivOperatorHyper(self.mu_0, self.lambda_0, self.calcStressDiff(eps), 1.0)
        if self.mode == "viscosity":
            tau = self.calcStressDiff(eps)
            m, l, a = 1 / (4 * self.mu_0), math.inf, 1 / (2 * self.mu_0)
        else:
            tau = self.calcStressConst(self.mu_0, self.lambda_0, eps)
            m, l, a = self.mu_0, self.lambda_0, 1.0
        f = self.divOperatorStaggered(tau)
        return self.ifft(self.G0OperatorFourierStaggered(m, l, self.fft(f), a))

    def calcPressure(self, eps=None):
        """get_raw_field('p') fg:15559-15573: calcStressDiff, divOperatorStaggered, divVector(alpha = 1/(2 mu0)), poisson_solve"""
        eps = self.epsilon if eps is None else eps
        f = self.divOperatorStaggered(self.calcStressDiff(eps))
        return self.poisson_solve(self.divVector(f, 1 / (2 * self.mu_0)))

    def divVector(self, tau, alpha=1.0):
        """fg:19983-20003: b = sum_a (tau_a(x) - tau_a(x + e_a)) * alpha * n_a / L_a (forward neighbour, periodic)"""
        hx, hy, hz = self._h()
        return ((tau[0] - np.roll(tau[0], -1, 0)) * (alpha * hx) + (tau[1] - np.roll(tau[1], -1, 1)) * (alpha * hy)
                + (tau[2] - np.roll(tau[2], -1, 2)) * (alpha * hz))

    def poisson_solve(self, f):
        """fg:23454-23499: unscaled forward transform, division by 2 nxyz sum_a (n_a/L_a)^2 (cos(2 pi i_a / n_a) - 1), zero mean,
        unscaled backward transform"""
        uc = sfft.rfftn(f, axes=(-3, -2, -1), workers=FFT_WORKERS)
        c = 2.0 * self.nxyz
        terms = []
        for n, L, cnt, shape in ((self.nx, self.dx, self.nx, (-1, 1, 1)), (self.ny, self.dy, self.ny, (1, -1, 1)),
                                 (self.nz, self.dz, self.nzc, (1, 1, -1))):
            xi = (2.0 * math.pi / n) * np.arange(cnt, dtype=float)
            terms.append(((n * n / (L * L)) * (np.cos(xi) - 1.0)).reshape(shape))
        with np.errstate(divide='ignore', invalid='ignore'):
            uc = uc / (c * (terms[0] + terms[1] + terms[2]))
        uc[0, 0, 0] = 0
        return sfft.irfftn(uc, s=(self.nx, self.ny, self.nz), axes=(-3, -2, -1), norm="forward", workers=FFT_WORKERS)

    # -- collocated Fourier div / grad for hyperelasticity (fg:20155-20218, fg:22069-22116) ------------
    def _xi2pi(self):
        x0 = (2 * math.pi / self.dx) * self._freq(self.nx)
        x1 = (2 * math.pi / self.dy) * self._freq(self.ny)
        x2 = (2 * math.pi / self.dz) * self._freq(self.nz, self.nzc)
        return x0[:, None, None], x1[None, :, None], x2[None, None, :]

    def G0DivOperatorFourierHyper(self, mu_0, lambda_0, tau_hat, alpha=1.0):
        """fg:20155-20218: f^ = i xi . tau^ (rows of the non-symmetric 9-component tensor), u^ = c1 f^ + c2 xi (xi . f^), u^(0) = 0"""
        x = self._xi2pi()
        c10 = -alpha / (2 * mu_0)
        with np.errstate(divide='ignore'):
            c20 = alpha / (2 * mu_0 * (1 + np.float64(2 * mu_0) / np.float64(lambda_0)))
        norm2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2]
        with np.errstate(divide='ignore', invalid='ignore'):
            c1 = c10 / norm2
            c2 = c20 / (norm2 * norm2)
            f1 = 1j * (x[0] * tau_hat[0] + x[1] * tau_hat[5] + x[2] * tau_hat[4])
            f2 = 1j * (x[0] * tau_hat[8] + x[1] * tau_hat[1] + x[2] * tau_hat[3])
            f3 = 1j * (x[0] * tau_hat[7] + x[1] * tau_hat[6] + x[2] * tau_hat[2])
            eta = np.stack([c1 * f1 + c2 * (x[0] * x[0] * f1 + x[0] * x[1] * f2 + x[0] * x[2] * f3),
                            c1 * f2 + c2 * (x[1] * x[0] * f1 + x[1] * x[1] * f2 + x[1] * x[2] * f3),
                            c1 * f3 + c2 * (x[2] * x[0] * f1 + x[2] * x[1] * f2 + x[2] * x[2] * f3)])
        eta[:, 0, 0, 0] = 0
        return eta

    def G0DivOperatorHyper(self, mu_0, lambda_0, tau, alpha=-1.0):
        """fg:20281-20286: fftTensor, G0DivOperatorFourierHyper, fftInvVector"""
        return self.ifft(self.G0DivOperatorFourierHyper(mu_0, lambda_0, self.fft(tau), alpha))

    def GradOperatorFourierHyper(self, q_hat):
        """fg:22069-22116: W^ = i xi (x) q^ in the stored component order 11,22,33,23,13,12,32,31,21"""
        x0, x1, x2 = self._xi2pi()
        q0, q1, q2 = q_hat[0], q_hat[1], q_hat[2]
        return np.stack([1j * x0 * q0, 1j * x1 * q1, 1j * x2 * q2, 1j * x2 * q1, 1j * x2 * q0, 1j * x1 * q0,
                         1j * x1 * q2, 1j * x0 * q2, 1j * x0 * q1])

    # -- FFT wrappers (fg:18481-18584): forward scaled 1/nxyz, backward unscaled ---------
    def fft(self, x):
        return sfft.rfftn(x, axes=(-3, -2, -1), norm="forward", workers=FFT_WORKERS)

    def ifft(self, y):
        return sfft.irfftn(y, s=(self.nx, self.ny, self.nz), axes=(-3, -2, -1), norm="forward", workers=FFT_WORKERS)

    # -- staggered stencils (fg:18614-19071) -----------------------------------------
    def _h(self):
        return self.nx / self.dx, self.ny / self.dy, self.nz / self.dz

    @staticmethod
    def _fwd(a, axis, h):          # D+ : (a(i+1) - a(i)) * h
        return (np.roll(a, -1, axis) - a) * h

    @staticmethod
    def _bwd(a, axis, h):          # D- : (a(i) - a(i-1)) * h
        return (a - np.roll(a, 1, axis)) * h

    def divOperatorStaggered(self, x):
        hx, hy, hz = self._h()
        f, b = self._fwd, self._bwd
        if self.dim == 3:          # fg:18914-18968
            return (b(x[0], 0, hx) + b(x[1], 1, hy) + b(x[2], 2, hz))[None]
        if self.dim == 6:          # fg:18853-18908
            return np.stack([b(x[0], 0, hx) + f(x[5], 1, hy) + f(x[4], 2, hz),
                             f(x[5], 0, hx) + b(x[1], 1, hy) + f(x[3], 2, hz),
                             f(x[4], 0, hx) + f(x[3], 1, hy) + b(x[2], 2, hz)])
        return np.stack([b(x[0], 0, hx) + f(x[5], 1, hy) + f(x[4], 2, hz),      # fg:19016-19071
                         f(x[8], 0, hx) + b(x[1], 1, hy) + f(x[3], 2, hz),
                         f(x[7], 0, hx) + f(x[6], 1, hy) + b(x[2], 2, hz)])

    def epsOperatorStaggered(self, E, u):
        hx, hy, hz = self._h()
        f, b = self._fwd, self._bwd
        if self.dim == 3:          # fg:18697-18757
            return np.stack([E[0] + f(u[0], 0, hx), E[1] + f(u[0], 1, hy), E[2] + f(u[0], 2, hz)])
        if self.dim == 6:          # fg:18614-18692
            return np.stack([E[0] + f(u[0], 0, hx), E[1] + f(u[1], 1, hy), E[2] + f(u[2], 2, hz),
                             E[3] + 0.5 * (b(u[2], 1, hy) + b(u[1], 2, hz)),
                             E[4] + 0.5 * (b(u[2], 0, hx) + b(u[0], 2, hz)),
                             E[5] + 0.5 * (b(u[1], 0, hx) + b(u[0], 1, hy))])
        return np.stack([E[0] + f(u[0], 0, hx), E[1] + f(u[1], 1, hy), E[2] + f(u[2], 2, hz),   # fg:18763-18846
                         E[3] + b(u[1], 2, hz), E[4] + b(u[0], 2, hz), E[5] + b(u[0], 1, hy),
                         E[6] + b(u[2], 1, hy), E[7] + b(u[2], 0, hx), E[8] + b(u[1], 0, hx)])

    # -- frequency helpers (fg:19393-19406) ---------------------------------------------
    @staticmethod
    def _freq(n, count=None):
        half = (n // 2 - 1) if n % 2 == 0 else n // 2
        i = np.arange(n if count is None else count, dtype=float)
        return np.where(i <= half, i, i - n)

    # -- G0 staggered in Fourier space (fg:19749-19927) -----------------------------------
    def _kpm(self):
        out = []
        for n, L, cnt in ((self.nx, self.dx, None), (self.ny, self.dy, None), (self.nz, self.dz, self.nzc)):
            h = L / (2 * n)
            xi0 = 2 * math.pi * h / L
            xi = xi0 * self._freq(n, cnt)
            kpm = np.sin(xi) / h
            kp = kpm * np.exp(1j * xi)
            km = -np.conj(kp)
            out.append((kpm, kp, km))
        return out

    def G0OperatorFourierStaggered(self, mu_0, lambda_0, tau_hat, alpha=-1.0):
        (s0, kp0, km0), (s1, kp1, km1), (s2, kp2, km2) = self._kpm()
        norm2 = (s0 * s0)[:, None, None] + (s1 * s1)[None, :, None] + (s2 * s2)[None, None, :]
        with np.errstate(divide='ignore', invalid='ignore'):
            if self.dim == 3:                                       # fg:19758-19763, fg:19778-19830
                c10 = -alpha / (2 * mu_0)
                eta = (c10 / norm2) * tau_hat[:1]
                eta[0, 0, 0, 0] = 0
                return eta
            if self.dim == 6:                                       # fg:19749-19755
                c10 = -alpha / mu_0
                c20 = -alpha / (mu_0 * (1 + mu_0 / (lambda_0 + mu_0)))
            else:                                                   # fg:19768-19774
                c10 = -alpha / (2 * mu_0)
                c20 = -alpha / (2 * mu_0 * (1 + np.float64(2 * mu_0) / np.float64(lambda_0)))
            c1 = c10 / norm2
            c2 = c20 / (norm2 * norm2)
            KP = (kp0[:, None, None], kp1[None, :, None], kp2[None, None, :])
            KM = (km0[:, None, None], km1[None, :, None], km2[None, None, :])
            c2_fkp = c2 * (tau_hat[0] * KP[0] + tau_hat[1] * KP[1] + tau_hat[2] * KP[2])
            eta = np.stack([c1 * tau_hat[j] + c2_fkp * KM[j] for j in range(3)])
        eta[:, 0, 0, 0] = 0
        return eta

    # -- collocated Gamma in Fourier space (fg:19302-19745) ----------------------------------
    def _xi(self):
        x0 = (1 / self.dx) * self._freq(self.nx)
        x1 = (1 / self.dy) * self._freq(self.ny)
        x2 = (1 / self.dz) * self._freq(self.nz, self.nzc)
        return x0[:, None, None], x1[None, :, None], x2[None, None, :]

    def GammaOperatorFourierCollocated(self, E, mu_0, lambda_0, tau_hat, alpha=-1.0, beta=0.0):
        xi = self._xi()
        shp = tau_hat.shape[1:]
        X = [np.broadcast_to(x, shp) for x in xi]
        norm2 = X[0] * X[0] + X[1] * X[1] + X[2] * X[2]
        d = self.dim
        with np.errstate(divide='ignore', invalid='ignore'):
            if d == 3:                                              # fg:19302-19377
                c1 = (alpha / (2 * mu_0)) / norm2
                ey = []
                for i in range(3):
                    c = 0
                    for j in range(3):
                        c = c + (c1 * X[i] * X[j]) * tau_hat[j]
                    ey.append(c)
                eta = np.stack(ey) + beta * tau_hat
            elif d == 9:                                            # fg:19619-19745
                c10 = alpha / (2 * mu_0)
                c20 = -alpha / (2 * mu_0 * (1 + np.float64(2 * mu_0) / np.float64(lambda_0)))
                c1 = c10 / norm2
                c2 = c20 / (norm2 * norm2)
                ey = []
                for i in range(9):
                    c = 0
                    for j in range(9):
                        g = c2 * (X[ROW[i]] * X[COL[i]] * X[ROW[j]] * X[COL[j]])
                        if ROW[i] == ROW[j]:
                            g = c1 * (X[COL[i]] * X[COL[j]]) + g
                        c = c + g * tau_hat[j]
                    ey.append(c)
                eta = np.stack(ey) + beta * tau_hat
            else:                                                   # fg:19381-19608
                eta = self._gamma_colloc_el(X, norm2, mu_0, lambda_0, tau_hat, alpha, beta)
        eta[:, 0, 0, 0] = E
        return eta

    def GammaOperatorFourierWillotR(self, E, mu_0, lambda_0, tau_hat, alpha=-1.0, beta=0.0):
        """fg:19083-19298 (the "safe" branch, WILLOT_ALLOW_NONZERO_LAMBDA): Willot's rotated-scheme Green operator for linear
        elasticity.  Discrete frequency vector k_a = (i/4) tan(q_a/2) (1+e^{i q_0})(1+e^{i q_1})(1+e^{i q_2}) / h_a with
        q_a = 2 pi m_a / n_a, r = k/|k|; the Hermitian 6x6 matrix gamma(iv,jv) of fg:19237-19247 is applied with the Voigt
        factor 2 on the shear columns (fg:19266-19273).  Needs lambda_0 != 0 (mu_0/lambda_0 appears explicitly), as in the
        reference; eta_hat(0) = E."""
        if self.dim != 6:
            raise RuntimeError("Unknown gamma scheme 'willot' for mode '%s'" % self.mode)       # fg:20488-20531: elasticity only
        small = np.finfo(float).tiny
        with np.errstate(divide='ignore', invalid='ignore'):
            mu_lambda_0 = np.float64(mu_0) / np.float64(lambda_0)
        n = (self.nx, self.ny, self.nz)
        L = (self.dx, self.dy, self.dz)
        cnt = (None, None, self.nzc)
        shape = [(-1, 1, 1), (1, -1, 1), (1, 1, -1)]
        q, w = [], []
        for a in range(3):
            xi = (2 * math.pi / L[a]) * self._freq(n[a], cnt[a])
            wa = L[a] / n[a]
            q.append((xi * wa).reshape(shape[a]))
            w.append(wa)
        ex = [1.0 + (np.cos(qa) + 1j * np.sin(qa)) for qa in q]                               # 1 + polar(1, q) fg:19131
        exp012 = ex[0] * ex[1] * ex[2]
        k = [(1j * (0.25 * np.tan(0.5 * q[a]))) * exp012 / w[a] for a in range(3)]
        k = [np.broadcast_to(ka, tau_hat.shape[1:]) for ka in k]
        mag = np.sqrt(sum(ka.real ** 2 + ka.imag ** 2 for ka in k)) + small
        r = [ka / mag for ka in k]
        rc = [np.conj(ra) for ra in r]
        rr = r[0] * r[0] + r[1] * r[1] + r[2] * r[2]
        r2 = rr.real ** 2 + rr.imag ** 2
        vi = (0, 1, 2, 1, 0, 0)
        vj = (0, 1, 2, 2, 2, 1)
        dl = np.eye(3)

        def S(a, b, c):
            # the s.. terms of fg:19178-19215: index a == b gives 4 Im(r_c conj r_a)^2, else -4 Im(r_a conj r_b) Im(r_a conj r_c)
            if a == b:
                t = (r[c] * rc[a]).imag
                return 4.0 * t * t
            return -4.0 * (r[a] * rc[b]).imag * (r[a] * rc[c]).imag

        g = {}
        with np.errstate(divide='ignore', invalid='ignore'):
            for iv in range(6):
                for jv in range(iv, 6):
                    i, j, kk, l = vi[iv], vj[iv], vi[jv], vj[jv]
                    sjk, sjl, sik, sil = S(kk, j, i), S(l, j, i), S(kk, i, j), S(l, i, j)
                    num = ((1 + 2 * mu_lambda_0) * 0.25 * (r[i] * rc[l] * dl[j, kk] + r[j] * rc[l] * dl[i, kk]
                                                           + r[i] * rc[kk] * dl[j, l] + r[j] * rc[kk] * dl[i, l])
                           + (0.25 * (r[i] * rc[l] * sjk + r[j] * rc[l] * sik + r[i] * rc[kk] * sjl + r[j] * rc[kk] * sil)
                              - (r[i] * rc[j]).real * (r[kk] * rc[l]).real)
                           - mu_lambda_0 * r[i] * r[j] * rc[kk] * rc[l])
                    g[iv, jv] = num / (mu_0 * (2 * (1 + mu_lambda_0) - r2))
                    g[jv, iv] = np.conj(g[iv, jv])
            ey = []
            for iv in range(6):
                c = 0
                for j in range(3, 6):
                    c = c + g[iv, j] * tau_hat[j]
                c = c * 2
                for j in range(3):
                    c = c + g[iv, j] * tau_hat[j]
                ey.append(c)
            eta = alpha * np.stack(ey) + beta * tau_hat
        eta[:, 0, 0, 0] = E
        return eta

    def _gamma_colloc_el(self, X, norm2, mu_0, lambda_0, tau_hat, alpha, beta):
        c10 = alpha / (4 * mu_0)
        c20 = -alpha / (mu_0 * (1 + mu_0 / (lambda_0 + mu_0)))
        xi0, xi1, xi2 = X
        xi00, xi11, xi22 = xi0 * xi0, xi1 * xi1, xi2 * xi2
        xi01, xi02, xi12 = xi0 * xi1, xi0 * xi2, xi1 * xi2
        c1 = c10 / norm2
        c12 = c1 * 2
        c2 = c20 / (norm2 * norm2)
        c3, c4, c5 = (c12 + c2 * xi00), (c12 + c2 * xi11), (c12 + c2 * xi22)

        def calc_g(S0, S1, S2):
            g = {}
            g[0, 0] = (c12 + c3) * xi00
            g[1, 0] = c2 * xi00 * xi11
            g[2, 0] = c2 * xi00 * xi22
            g[3, 0] = c2 * xi00 * xi12 * S1 * S2
            g[4, 0] = c3 * xi02 * S0 * S2
            g[5, 0] = c3 * xi01 * S0 * S1
            g[1, 1] = (c12 + c4) * xi11
            g[2, 1] = c2 * xi11 * xi22
            g[3, 1] = c4 * xi12 * S1 * S2
            g[4, 1] = c2 * xi11 * xi02 * S0 * S2
            g[5, 1] = c4 * xi01 * S0 * S1
            g[2, 2] = (c12 + c5) * xi22
            g[3, 2] = c5 * xi12 * S1 * S2
            g[4, 2] = c5 * xi02 * S0 * S2
            g[5, 2] = c2 * xi22 * xi01 * S0 * S1
            g[3, 3] = (c1 * (xi11 + xi22) + c2 * xi11 * xi22)
            g[4, 3] = (c1 + c2 * xi22) * xi01 * S0 * S1
            g[5, 3] = (c1 + c2 * xi11) * xi02 * S0 * S2
            g[4, 4] = (c1 * (xi00 + xi22) + c2 * xi00 * xi22)
            g[5, 4] = (c1 + c2 * xi00) * xi12 * S1 * S2
            g[5, 5] = (c1 * (xi00 + xi11) + c2 * xi00 * xi11)
            return g

        g = calc_g(1, 1, 1)
        if self.freq_hack:                                           # fg:19458-19471
            shp = norm2.shape
            fi = np.zeros(shp, bool)
            fj = np.zeros(shp, bool)
            fk = np.zeros(shp, bool)
            if self.nx % 2 == 0:
                fi[self.nx // 2, :, :] = True
            if self.ny % 2 == 0:
                fj[:, self.ny // 2, :] = True
            if self.nz % 2 == 0 and self.nz // 2 < self.nzc:
                fk[:, :, self.nz // 2] = True
            filt = fi | fj | fk
            s = np.where(fi, 0.5, 1.0) * np.where(fj, 0.5, 1.0) * np.where(fk, 0.5, 1.0)
            acc = {k: np.zeros(shp) for k in g}
            for i in (1, -1):
                for j in (1, -1):
                    for k2 in (1, -1):
                        # loop bounds: sign -1 only visited where that axis is filtered
                        active = (fi | (i == 1)) & (fj | (j == 1)) & (fk | (k2 == 1))
                        gg = calc_g(i, j, k2)
                        for k in g:
                            acc[k] = acc[k] + np.where(active, s * gg[k], 0.0)
            g = {k: np.where(filt, acc[k], g[k]) for k in g}
        G = lambda i, j: g[(i, j)] if i >= j else g[(j, i)]
        ey = []
        for i in range(6):
            ey.append(tau_hat[0] * G(i, 0) + tau_hat[1] * G(i, 1) + tau_hat[2] * G(i, 2)
                      + (tau_hat[3] * G(i, 3) + tau_hat[4] * G(i, 4) + tau_hat[5] * G(i, 5)) * 2.0)
        return np.stack(ey) + beta * tau_hat

    # -- Gamma operator compositions (fg:20288-20531) -------------------------------------
    def GammaOperator(self, E, mu_0, lambda_0, tau, alpha=-1.0, beta=0.0):
        self.n_operator_applications += 1
        if self.mode == "viscosity":
            return self.DeltaOperator(E, mu_0, lambda_0, tau, alpha)
        if self.gamma_scheme == "collocated":                        # fg:20302-20340
            tau_hat = self.fft(tau)
            self.F0 = tau_hat[:, 0, 0, 0].real.copy()                # fg:20220
            eta_hat = self.GammaOperatorFourierCollocated(E, mu_0, lambda_0, tau_hat, alpha, beta)
            eta_hat[:, 0, 0, 0] += alpha * self.calcBCProjector()    # fg:20272
            return self.ifft(eta_hat)
        if self.gamma_scheme == "staggered":                         # fg:20288-20300, 20342-20378
            return self._GammaOperatorStaggered(E, mu_0, lambda_0, tau, alpha)
        if self.gamma_scheme == "willot" and self.mode == "elasticity":   # GammaOperatorWillotR fg:20322-20330
            return self._GammaOperatorWillotR(E, mu_0, lambda_0, tau, alpha, beta)
        raise RuntimeError("Unknown gamma scheme '%s'" % self.gamma_scheme)

    def _GammaOperatorStaggered(self, E, mu_0, lambda_0, tau, alpha):
        self._initBCProjector_real(tau)
        f = self.divOperatorStaggered(tau)
        u = self.ifft(self.G0OperatorFourierStaggered(mu_0, lambda_0, self.fft(f), alpha))
        eta = self.epsOperatorStaggered(E, u)
        R = alpha * self.calcBCProjector()                           # fg:20263-20270
        return eta + R.reshape((-1, 1, 1, 1))

    def DeltaOperator(self, E, mu_0, lambda_0, tau, alpha=-1.0):
        """fg:20474-20486 -- viscosity dual formulation: DeltaOperatorStaggered fg:20422-20460, DeltaOperatorWillotR fg:20380-20418,
        DeltaOperatorCollocated fg:20462-20471"""
        if self.gamma_scheme == "collocated":
            # fftTensor(zero_trace) fg:18531-18559: component 0 is not transformed, tau^_0 = -(tau^_1 + tau^_2)
            tau_hat = self.fft(tau)
            tau_hat[0] = -(tau_hat[1] + tau_hat[2])
            m = 1 / (4 * mu_0)
            self.F0 = tau_hat[:, 0, 0, 0].real.copy()                # applyDeltaFourier fg:19075-19080
            eta_hat = self.GammaOperatorFourierCollocated(E, -1.0 / (4 * m), math.inf, tau_hat, alpha, 2 * alpha * m)
            eta_hat[:, 0, 0, 0] += alpha * self.calcBCProjector()
            eta = self.ifft(eta_hat)                                 # fftInvTensor(zero_trace) fg:18563-18584
            eta[0] = -(eta[1] + eta[2])
            return eta
        mu_0 = 1 / (4 * mu_0)
        adj = E - 2 * alpha * mu_0 * self.average(tau)
        if self.gamma_scheme == "staggered":
            eta = self._GammaOperatorStaggered(adj, -1.0 / (4 * mu_0), math.inf, tau, alpha)
        elif self.gamma_scheme == "willot":
            eta = self._GammaOperatorWillotR(adj, -1.0 / (4 * mu_0), math.inf, tau, alpha, 0.0)
        else:
            raise RuntimeError("Unknown gamma scheme '%s'" % self.gamma_scheme)
        return eta + (2 * alpha * mu_0) * tau

    def _GammaOperatorWillotR(self, E, mu_0, lambda_0, tau, alpha, beta):
        """GammaOperatorWillotR fg:20322-20330"""
        tau_hat = self.fft(tau)
        self.F0 = tau_hat[:, 0, 0, 0].real.copy()
        eta_hat = self.GammaOperatorFourierWillotR(E, mu_0, lambda_0, tau_hat, alpha, beta)
        eta_hat[:, 0, 0, 0] += alpha * self.calcBCProjector()
        return self.ifft(eta_hat)

    # -- schemes (fg:20536-20590) ------------------------------------------------------
    def basicScheme(self, E, eps):
        if self.bc_relax != 1.0:
            self.F00 = self.average(eps)
        tau = self.calcStressDiff(eps)
        return self.GammaOperator(E, self.mu_0, self.lambda_0, tau, -1.0)

    def krylovOperator(self, eps):
        return self.basicScheme(np.zeros(self.dim), eps)

    def polarizationScheme(self, P0, eps):
        if self.bc_relax != 1.0:
            self.F00 = self.average(eps)
        tau = self.calcPolarization(self.mu_0, eps)
        P00 = self.average(tau)
        return self.GammaOperator(P00 + P0, self.mu_0, self.lambda_0, tau, -4 * self.mu_0, 1.0)

    # -- reference material (fg:22283-22313, fg:12153-12236, fg:12472-12559) ------------------
    def getRefMaterial(self, F, zero_trace, polarization):
        d = self.dim
        F = self._mat_in(F)                                          # fg:22288-22294
        n = self.nxyz_mat
        Ff = F.reshape(d, n)
        linear = all(ph.law.linear for ph in self.mat.phases)
        if linear:
            # tangent independent of F: evaluate once per distinct phi tuple (+orientation-free laws only)
            key = np.stack([ph.phi.reshape(n) for ph in self.mat.phases], axis=1)
            tiso = any(isinstance(ph.law, LinearTransverselyIsotropic) for ph in self.mat.phases)
            if tiso:
                idx = np.arange(n)
            else:
                _, idx = np.unique(key, axis=0, return_index=True)
        else:
            idx = np.arange(n)
        sel = np.zeros(n, dtype=bool)
        sel[idx] = True
        sel = sel.reshape(F.shape[1:])
        C = self.mat.tangent_rows(F[:, sel], sel=sel)               # (d,d,m) row m = dPK1(e_m)
        Cm = np.moveaxis(C, -1, 0)
        if zero_trace:
            Cm = Cm[:, 1:, 1:]
        ev = np.linalg.eigvalsh(Cm, UPLO='L')                       # syev 'U' on row-major data == lower
        lmin, lmax = float(ev.min()), float(ev.max())
        if lmin < 0:
            lmin = 0.0
        mu_0 = math.sqrt(lmin * lmax) if polarization else 0.5 * (lmin + lmax)
        return mu_0, lmin, lmax

    def calcRefMaterial(self, F):
        mu_0, lmin, lmax = self.getRefMaterial(F, self.mode == "viscosity", self.method == "polarization")
        self.mu_0 = mu_0 * 0.5 * self.ref_scale
        self.ref_eigs = (lmin, lmax)
        self.setBCProjector(self.BC_P)

    # -- convergence (fg:21129-21244) ------------------------------------------------------
    def bc_error(self):
        Emean = self.average(self.epsilon)
        Smean = self.calcMeanStress()
        P_Emean = dyad4_mv(self.BC_P, Emean)
        Q_Smean = dyad4_mv(self.BC_Q, Smean)
        PE = dyad4_mv(self.BC_P, self.current_E)
        if self.dim == 9:
            PE = PE - dyad4_mv(self.BC_P, self.Id)
        norm_E = vnorm_2(PE)
        err_F = vnorm_2(P_Emean - self.current_E) / (1 if norm_E < self.bc_tol else norm_E)
        norm_S = vnorm_2(self.current_S)
        err_S = vnorm_2(Q_Smean - self.current_S) / (1 if norm_S < self.bc_tol else norm_S)
        return max(err_F, err_S)

    def converged(self, it, abs_err, rel_err, check_bc=True):
        """returns (stop, next_iter)"""
        if math.isnan(rel_err):
            raise FloatingPointError("NaN detected in solution. Aborting.")
        self.residuals.append(rel_err)
        if self.callback is not None and self.callback():
            return True, it
        if it >= self.maxiter:
            return True, it
        if rel_err <= self.tol or abs_err <= self.abs_tol:
            bc_err = self.bc_error() if check_bc else 0.0
            self.last_bc_err = bc_err
            if bc_err <= self.bc_tol:
                return True, it
        return False, it + 1

    def create_error_estimator(self, name=None):
        name = name or self.error_estimator
        return {"sigma": lambda: SigmaEE(self), "epsilon": lambda: EpsilonEE(self),
                "energy": lambda: EnergyEE(self), "residual": ResidualEE, "none": NoneEE}[name]()

    # -- drivers ------------------------------------------------------------------------------
    def run(self):
        """fg:21247-21399 + runLoadsteppingSolver fg:21584-21686"""
        self.residuals = []
        self.setBCProjector(self.BC_P)
        eps = math.sqrt(EPS)
        if np.linalg.norm(dyad4_mv(self.BC_P, self.S)) > eps * np.linalg.norm(self.S):
            raise RuntimeError("Incompatible stress boundary condition specified")
        if np.linalg.norm(dyad4_mv(self.BC_Q, self.E)) > eps * np.linalg.norm(self.E):
            raise RuntimeError("Incompatible strain boundary condition specified")
        d = self.dim
        self.epsilon = np.zeros((d, self.nx, self.ny, self.nz))
        if self.mode == "hyperelasticity":
            self.epsilon += self.Id.reshape(-1, 1, 1, 1)
        first = 0 if len(self.loadsteps) > 2 else 1
        last = []                                                    # (t, eps) of the previous load steps, fg:21444-21447
        for istep in range(first, len(self.loadsteps)):
            t = self.loadsteps[istep]
            E = t * self.E
            S = t * self.S
            if self.mode == "hyperelasticity":
                E = E + (1 - t) * dyad4_mv(self.BC_P, self.Id)
            if self.loadstep_extrapolation_order > 0 and istep > first:      # fg:21634-21650
                while len(last) > self.loadstep_extrapolation_order:
                    last.pop(0)
                last.append((self.loadsteps[istep - 1], self.epsilon.copy()))
                if len(last) >= 2:
                    self.extrapolateLoadstepPolynomial(last, t)
            self.runSolver(E, S)
        return False

    def extrapolateLoadstepPolynomial(self, last, t):
        """fg:21468-21513: per voxel the polynomial through the stored load steps (Vandermonde inverse by gesv), evaluated at t"""
        import scipy.linalg as sla
        n = len(last)
        V = np.array([[ls[0] ** j for j in range(n)] for ls in last], dtype=float)
        tp = np.array([t ** i for i in range(n)], dtype=float)
        Vinv = sla.solve(V, np.eye(n))
        f = np.stack([ls[1] for ls in last])                         # (n, d, nx, ny, nz)
        pcoef = np.tensordot(Vinv, f, axes=(1, 0))
        self.epsilon = np.tensordot(tp, pcoef, axes=(0, 0))

    def runSolver(self, E, S):
        self.current_E, self.current_S = E, S
        if self.method == "basic":
            self.runBasic(E, S)
        elif self.method == "polarization":
            self.runPolarization(E, S)
        elif self.method == "cg":
            if self.mode == "hyperelasticity":
                self.runCGHyper(E, S)
            else:
                self.runCGElasticity(E, S)
        else:
            raise RuntimeError("Unknown solver method '%s'" % self.method)

    def runBasic(self, E0, S0):
        """fg:21716-21805"""
        it = 1
        ee = self.create_error_estimator()
        update_ref = self.update_ref != "never"
        E = None
        while True:
            if update_ref:
                self.calcRefMaterial(self.epsilon)
                E = self.calcBCMean(E0, S0)
                update_ref = False
            elif E is None:
                E = self.calcBCMean(E0, S0)
            self.epsilon = self.basicScheme(E, self.epsilon)
            ee.update()
            stop, it = self.converged(it, ee.abs_err, ee.rel_err)
            if stop:
                break

    def runPolarization(self, E0, S0):
        """fg:21808-21851"""
        it = 1
        ee = self.create_error_estimator()
        if self.update_ref != "never":
            self.calcRefMaterial(self.epsilon)
        E = self.calcBCMean(E0, S0)
        self.epsilon = np.zeros_like(self.epsilon) + (4 * self.mu_0 * E).reshape(-1, 1, 1, 1)
        while True:
            P0 = 4 * self.mu_0 * E
            self.epsilon = self.polarizationScheme(P0, self.epsilon)
            ee.update()
            stop, it = self.converged(it, ee.abs_err, ee.rel_err, False)
            if stop:
                break
        self.epsilon = self.calcPolarization(self.mu_0, self.epsilon, True)

    def runCGElasticity(self, E0, S0):
        """fg:23153-23247"""
        if self.update_ref != "never":
            self.calcRefMaterial(self.epsilon)
        E = self.calcBCMean(E0, S0)
        ee = self.create_error_estimator()
        Ecol = E.reshape(-1, 1, 1, 1)
        self.epsilon = np.zeros_like(self.epsilon) + Ecol
        r = self.krylovOperator(self.epsilon)
        r = r + (Ecol - self.epsilon)                                # adjustResidual fg:10012
        gamma = self.innerProduct(r, r) + SMALL
        gamma0 = gamma
        p = r.copy()
        it = 0
        self.cg_scalars = []
        while True:
            w = self.krylovOperator(p)
            alpha = self.innerProduct(p, p, w) + SMALL
            self.cg_scalars.append((gamma, alpha))
            alpha = gamma / alpha
            self.epsilon = self.epsilon + alpha * p
            ee.update_cg(gamma, gamma0)
            stop, it = self.converged(it, ee.abs_err, ee.rel_err)
            if stop:
                break
            if self.cg_reinit > 0 and (it % self.cg_reinit) == 0:
                r = self.krylovOperator(self.epsilon)
                r = r + (Ecol - self.epsilon)
            else:
                r = r + (-alpha) * (p - w)
            delta = self.innerProduct(r, r) + SMALL
            beta = delta / gamma
            gamma = delta
            p = r + beta * p

    def ApplyOperator(self, F, Q):
        """fg:23132-23150"""
        W = self.calcStressDeriv(self.mu_0, self.lambda_0, F, Q, 1.0)
        return self.GammaOperator(np.zeros(self.dim), self.mu_0, self.lambda_0, W, -1.0)

    def runCGHyper(self, E0, S0):
        """fg:22699-23130"""
        dE = E0 - dyad4_mv(self.BC_P, self.average(self.epsilon))
        self.epsilon = self.epsilon + dE.reshape(-1, 1, 1, 1)
        ee_outer = self.create_error_estimator(self.outer_error_estimator)
        it_outer = 0
        gamma0 = -1.0
        while True:
            if gamma0 < 0 or self.update_ref == "always":
                self.calcRefMaterial(self.epsilon)
            F = self.epsilon.copy()
            X = self.calcStress(0.0, 0.0, F, 1.0)
            X0 = dyad4_mv(self.BC_M, S0)
            X = self.GammaOperator(X0, self.mu_0, self.lambda_0, X, -1.0)
            R = self.ApplyOperator(F, X)
            Q = R.copy()
            gamma = self.innerProduct(R, R) + SMALL
            if gamma0 < 0:
                gamma0 = gamma
            ee = self.create_error_estimator()
            it = 0
            while True:
                W = self.ApplyOperator(F, Q)
                alpha = self.innerProduct(Q, Q, W) + SMALL
                if alpha <= 0:
                    raise ArithmeticError("indefinite operator (alpha=%g) canceling CG!" % alpha)
                alpha = gamma / alpha
                X = X + alpha * Q
                self.epsilon = F + self.newton_relax * X
                ee.update_cg(gamma, gamma0)
                stop, it = self.converged(it, ee.abs_err, ee.rel_err, False)
                if stop:
                    break
                R = R + (-alpha) * (Q - W)
                delta = self.innerProduct(R, R) + SMALL
                beta = delta / gamma
                gamma = delta
                Q = R + beta * Q
            ee_outer.update()
            stop, it_outer = self.converged(it_outer, ee_outer.abs_err, ee_outer.rel_err)
            if stop:
                break

    # -- effective properties (fg:26030-26160) -------------------------------------------------
    def calc_effective_properties(self):
        d = self.dim
        S = np.zeros((d, d))
        hist = []
        for i in range(d):
            E = np.zeros(d)
            E[i] = 1.0
            self.setStrain(E)
            self.run()
            hist.append(list(self.residuals))
            S[:, i] = self.calcMeanStress()
        Ceff = S.copy()                                              # E = identity
        Cv = Ceff.copy()
        if d == 6:
            Cv[:, 3:] *= 0.5
        self.effective_histories = hist
        return Ceff, Cv


# --------------------------------------------------------------------------- analytic laminate (fg:26405-26474)
def calc_isotropic_laminate(layers):
    """Closed-form effective stiffness of a laminate of isotropic layers [(phi, lambda, mu), ...],
    lamination direction x, as the reference's ``calc_isotropic_laminate`` action computes it
    (fg:26405-26446; Milton, Theory of Composites, eq. 9.9).  Returned in the same convention as
    ``Ceff_voigt`` of ``calc_effective_properties`` (fg:26083-26088), i.e. tensor-Voigt entries."""
    c1 = c2 = c3 = c4 = c5 = c6 = 0.0
    for phi, lam, mu in layers:
        c1 += phi / (lam + 2 * mu)
        c2 += phi / mu
        c3 += phi * mu
        c4 += phi * lam / (lam + 2 * mu)
        c5 += phi * 4 * mu * (lam + mu) / (lam + 2 * mu)
        c6 += phi * 2 * mu * lam / (lam + 2 * mu)
    C = np.zeros((6, 6))
    C[0, 0] = 1 / c1
    C[1, 1] = C[2, 2] = c5 + c4 * c4 / c1
    C[3, 3] = c3
    C[4, 4] = C[5, 5] = 1 / c2
    C[0, 1] = C[1, 0] = C[0, 2] = C[2, 0] = c4 / c1
    C[1, 2] = C[2, 1] = c6 + c4 * c4 / c1
    return C
