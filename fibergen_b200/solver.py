"""Python view of the host-side solver (fgb::LSSolver, include/fgb200_lssolver.h) and of the raw
device context (fgb_ctx, include/fgb200.h).  Both are thin: numpy arrays in the reference's padded
plane layout go in and out, all computation happens in libfgb200.so on the GPU.

Field arrays exchanged with this module have shape ``(dim, local_nx, ny, nzp)`` with
``nzp = 2*(nz//2+1)`` (fg:9557-9569); ``pad``/``unpad`` convert from/to ``(dim, nx, ny, nz)``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as L


def nzp_of(nz):
    return 2 * (nz // 2 + 1)


def pad(a):
    a = np.asarray(a, dtype=np.float64)
    out = np.zeros(a.shape[:-1] + (nzp_of(a.shape[-1]),))
    out[..., :a.shape[-1]] = a
    return out


def unpad(a, nz):
    return np.ascontiguousarray(a[..., :nz])


def lame(E, nu):
    """(E, nu) -> (lambda, mu), Material conversions fg:7294-7455"""
    return E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))


def _dp(a):
    return a.ctypes.data_as(L.c_dp)


def _planes(arr):
    """array (n, ...) contiguous -> (ctypes array of n double*, keepalive)"""
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    n = arr.shape[0]
    ptrs = (L.c_dp * n)()
    for i in range(n):
        ptrs[i] = arr[i].ctypes.data_as(L.c_dp)
    return ptrs, arr


def capsule_array(fibers):
    """fibers: iterable of (centre, axis, L0, R, material) as the <place_fiber> action takes them (fg:25789-25823)"""
    fibers = list(fibers)
    arr = (L.Capsule * max(len(fibers), 1))()
    for i, (c, a, L0, R, mat) in enumerate(fibers):
        arr[i].c[:] = [float(x) for x in c]
        arr[i].a[:] = [float(x) for x in a]
        arr[i].L0, arr[i].R, arr[i].material = float(L0), float(R), int(mat)
    return arr, len(fibers)


class Context:
    """Raw device context (fgb_create ... fgb_destroy) with numpy conveniences -- operator-level tests."""

    def __init__(self, nx, ny, nz, Lx=1.0, Ly=1.0, Lz=1.0, mode="elasticity", gamma_scheme="staggered",
                 device=-1, rank=0, nranks=1):
        self.lib = L.load()
        h = C.c_void_p()
        rc = self.lib.fgb_create(C.byref(h), nx, ny, nz, Lx, Ly, Lz, L.MODES[mode], L.SCHEMES[gamma_scheme], device, rank, nranks)
        if rc:
            raise L.FgbError(rc, self.lib.fgb_last_error(None).decode())
        self.h = h
        self.nx, self.ny, self.nz = nx, ny, nz
        self.nzp = nzp_of(nz)
        self.dim = self.lib.fgb_dim(h)
        self.lnx = self.lib.fgb_local_nx(h)
        self.udim = 1 if self.dim == 3 else 3

    def close(self):
        if getattr(self, "h", None):
            self.lib.fgb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def chk(self, rc):
        if rc < 0:
            raise L.FgbError(rc, self.lib.fgb_last_error(self.h).decode())
        return rc

    # fields -----------------------------------------------------------------
    def field(self, data=None):
        f = self.chk(self.lib.fgb_field_alloc(self.h))
        if data is not None:
            self.upload(f, data)
        return f

    def upload(self, f, data):
        """data: (dim, lnx, ny, nz) unpadded"""
        ptrs, keep = _planes(pad(data))
        self.chk(self.lib.fgb_field_upload(self.h, f, ptrs))

    def download(self, f):
        out = np.empty((self.dim, self.lnx, self.ny, self.nzp))
        ptrs, keep = _planes(out)
        self.chk(self.lib.fgb_field_download(self.h, f, ptrs))
        return unpad(keep, self.nz)

    def download_padded(self, f):
        out = np.empty((self.dim, self.lnx, self.ny, self.nzp))
        ptrs, keep = _planes(out)
        self.chk(self.lib.fgb_field_download(self.h, f, ptrs))
        return keep

    def upload_padded(self, f, data):
        ptrs, keep = _planes(data)
        self.chk(self.lib.fgb_field_upload(self.h, f, ptrs))

    def u_upload(self, data):
        ptrs, keep = _planes(pad(data))
        self.chk(self.lib.fgb_u_upload(self.h, ptrs, self.udim))

    def u_download(self):
        out = np.empty((self.udim, self.lnx, self.ny, self.nzp))
        ptrs, keep = _planes(out)
        self.chk(self.lib.fgb_u_download(self.h, ptrs, self.udim))
        return unpad(keep, self.nz)

    # setup --------------------------------------------------------------------
    def set_phases(self, phis, laws, mixing="voigt", normals=None, orientation=None):
        """phis: list of (lnx,ny,nz) arrays; laws: list of (law_name, params)"""
        self.chk(self.lib.fgb_set_num_phases(self.h, len(phis)))
        for p, (phi, (name, params)) in enumerate(zip(phis, laws)):
            pp = np.ascontiguousarray(pad(phi))
            self.chk(self.lib.fgb_set_phase(self.h, p, _dp(pp)))
            pr = np.ascontiguousarray(params, dtype=np.float64)
            self.chk(self.lib.fgb_set_law(self.h, p, L.LAWS[name], _dp(pr), pr.size))
        if normals is not None:
            ptrs, keep = _planes(pad(normals))
            self.chk(self.lib.fgb_set_normals(self.h, ptrs))
        if orientation is not None:
            ptrs, keep = _planes(pad(orientation))
            self.chk(self.lib.fgb_set_orientation(self.h, ptrs))
        self.chk(self.lib.fgb_set_mixing(self.h, L.MIXING[mixing], None, 0))

    def vec(self, v):
        a = np.zeros(9)
        v = np.asarray(v, dtype=np.float64)
        a[:v.size] = v
        return a

    # scalar-returning helpers ---------------------------------------------------
    def inner(self, a, b, c=-1):
        out = C.c_double()
        self.chk(self.lib.fgb_inner(self.h, a, b, c, C.byref(out)))
        return out.value

    def average(self, f):
        out = np.zeros(9)
        self.chk(self.lib.fgb_average(self.h, f, _dp(out)))
        return out[:self.dim].copy()

    def component_dot(self, a, b):
        out = np.zeros(9)
        self.chk(self.lib.fgb_component_dot(self.h, a, b, _dp(out)))
        return out[:self.dim].copy()

    def mean_pk1(self, f, alpha=1.0):
        out = np.zeros(9)
        self.chk(self.lib.fgb_mean_pk1(self.h, f, alpha, _dp(out)))
        return out[:self.dim].copy()

    def mean_energy(self, f):
        out = C.c_double()
        self.chk(self.lib.fgb_mean_energy(self.h, f, C.byref(out)))
        return out.value

    def ref_material(self, f, zero_trace=False):
        a, b = C.c_double(), C.c_double()
        self.chk(self.lib.fgb_ref_material(self.h, f, int(zero_trace), C.byref(a), C.byref(b)))
        return a.value, b.value

    def gamma(self, f, E, mu0, lambda0, alpha=-1.0, beta=0.0):
        e = self.vec(E)
        self.chk(self.lib.fgb_gamma(self.h, f, _dp(e), mu0, lambda0, alpha, beta))

    def profile(self, on=True):
        self.chk(self.lib.fgb_profile_enable(self.h, int(on)))

    def profile_results(self):
        buf = C.create_string_buffer(4096)
        self.chk(self.lib.fgb_profile_names(self.h, buf, 4096))
        res = {}
        for name in filter(None, buf.value.decode().split(",")):
            ms, n = C.c_double(), C.c_uint64()
            self.chk(self.lib.fgb_profile_get(self.h, name.encode(), C.byref(ms), C.byref(n)))
            res[name] = (ms.value, n.value)
        return res


class LSSolver:
    """Python handle on fgb::LSSolver -- the reference-facing object (same setting keys as the XML
    ``<solver>`` element / ``fibergen.FG.set('solver..key', v)``, fg:15044-15094)."""

    def __init__(self, nx, ny, nz, dx=1.0, dy=1.0, dz=1.0, rank=0, nranks=1, device=-1, **settings):
        self.lib = L.load()
        h = C.c_void_p()
        rc = self.lib.fgls_create(C.byref(h), nx, ny, nz, dx, dy, dz, rank, nranks, device)
        if rc:
            raise L.FgbError(rc, "fgls_create failed")
        self.h = h
        self.nx, self.ny, self.nz = nx, ny, nz
        self.nzp = nzp_of(nz)
        self._cb = None
        self._initialised = False
        self.lnx = nx // nranks
        for k, v in settings.items():
            self.set(k, v)

    def close(self):
        if getattr(self, "h", None):
            self.lib.fgls_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def chk(self, rc):
        if rc < 0:
            raise L.FgbError(rc, self.lib.fgls_last_error(self.h).decode())
        return rc

    def set(self, key, value):
        if isinstance(value, bool):
            value = "1" if value else "0"
        elif isinstance(value, float):
            value = repr(value)
        elif isinstance(value, (list, tuple, np.ndarray)):
            value = ",".join(repr(float(x)) for x in value)
        self.chk(self.lib.fgls_set(self.h, key.encode(), str(value).encode()))

    def add_material(self, name, law, *params):
        p = np.ascontiguousarray(np.asarray(params, dtype=np.float64).ravel())
        self.chk(self.lib.fgls_add_material(self.h, name.encode(), law.encode(), _dp(p), p.size))

    def set_reference(self, mu, lam):
        self.chk(self.lib.fgls_set_reference(self.h, mu, lam))

    def init(self):
        self.chk(self.lib.fgls_init(self.h))
        self.dim = self.lib.fgls_dim(self.h)
        self.lnx = self.lib.fgls_local_nx(self.h)
        self._initialised = True

    def init_comm(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self.chk(self.lib.fgls_init_comm(self.h, buf))

    def set_phase(self, m, phi, padded=False):
        pp = np.ascontiguousarray(phi if padded else pad(phi), dtype=np.float64)
        self.chk(self.lib.fgls_set_phase(self.h, m, _dp(pp)))

    def init_phase(self, fibers, matrix_mat=0, normals=False, orientation=False):
        """initPhi on the device from a fibre list [(centre, axis, L0, R, material), ...] (fg:17489)"""
        arr, n = capsule_array(fibers)
        self.chk(self.lib.fgls_init_phase_capsules(self.h, n, C.cast(arr, C.c_void_p), matrix_mat, int(normals), int(orientation)))

    def get_phase(self, m, padded=False):
        out = np.empty((self.lnx, self.ny, self.nzp))
        self.chk(self.lib.fgls_get_phase(self.h, m, _dp(out)))
        return out if padded else unpad(out, self.nz)

    def set_normals(self, n):
        ptrs, keep = _planes(pad(n))
        self.chk(self.lib.fgls_set_normals(self.h, ptrs))

    def set_orientation(self, a):
        ptrs, keep = _planes(pad(a))
        self.chk(self.lib.fgls_set_orientation(self.h, ptrs))

    def set_strain(self, E):
        e = np.zeros(9)
        E = np.asarray(E, dtype=np.float64)
        e[:E.size] = E
        self.chk(self.lib.fgls_set_strain(self.h, _dp(e)))

    def set_stress(self, S):
        s = np.zeros(9)
        S = np.asarray(S, dtype=np.float64)
        s[:S.size] = S
        self.chk(self.lib.fgls_set_stress(self.h, _dp(s)))

    def set_bc_projector(self, P):
        p = np.ascontiguousarray(P, dtype=np.float64)
        self.chk(self.lib.fgls_set_bc_projector(self.h, _dp(p)))

    def set_convergence_callback(self, fn):
        """fn() -> bool (True stops the iteration), reference: set_convergence_callback fg:27160"""
        if fn is None:
            self._cb = None
            self.chk(self.lib.fgls_set_callback(self.h, L.CALLBACK(0), None))
            return
        self._cb = L.CALLBACK(lambda user: 1 if fn() else 0)
        self.chk(self.lib.fgls_set_callback(self.h, self._cb, None))

    def run(self):
        self.chk(self.lib.fgls_run(self.h))

    def get_residuals(self):
        n = self.lib.fgls_num_residuals(self.h)
        out = np.zeros(max(n, 1))
        self.chk(self.lib.fgls_get_residuals(self.h, _dp(out), n))
        return out[:n].copy()

    def get_mean_stress(self):
        out = np.zeros(9)
        self.chk(self.lib.fgls_mean_stress(self.h, _dp(out)))
        return out[:self.dim].copy()

    def get_mean_strain(self):
        out = np.zeros(9)
        self.chk(self.lib.fgls_mean_strain(self.h, _dp(out)))
        return out[:self.dim].copy()

    def get_mean_energy(self):
        out = C.c_double()
        self.chk(self.lib.fgls_mean_energy(self.h, C.byref(out)))
        return out.value

    def get_effective_property(self):
        out = np.zeros(self.dim * self.dim)
        self.chk(self.lib.fgls_effective_properties(self.h, _dp(out)))
        return out.reshape(self.dim, self.dim)

    def get_field(self, name="epsilon", padded=False, out=None):
        """get_raw_field fg:15396: "epsilon", "sigma" (dim planes) or "u" (3 planes, 1 for heat)"""
        ncomp = self.lib.fgls_field_components(self.h, name.encode())
        if ncomp < 0:
            raise L.FgbError(-1, "Unknown field '%s'" % name)
        if out is None:
            out = np.empty((ncomp, self.lnx, self.ny, self.nzp))
        ptrs, keep = _planes(out)
        self.chk(self.lib.fgls_get_field(self.h, name.encode(), ptrs))
        return keep if padded else unpad(keep, self.nz)

    def get_mean_cauchy_stress(self):
        out = np.zeros(9)
        self.chk(self.lib.fgls_mean_cauchy_stress(self.h, _dp(out)))
        return out

    def ref_material(self):
        a, b = C.c_double(), C.c_double()
        self.chk(self.lib.fgls_ref_material(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def calc_ref_material(self):
        self.chk(self.lib.fgls_calc_ref_material(self.h))
        return self.ref_material()

    def bc_matrices(self):
        M = np.zeros(self.dim * self.dim)
        MQ = np.zeros(self.dim * self.dim)
        self.chk(self.lib.fgls_bc_matrices(self.h, _dp(M), _dp(MQ)))
        return M.reshape(self.dim, self.dim), MQ.reshape(self.dim, self.dim)

    def get_solve_time(self):
        return self.lib.fgls_solve_time(self.h)

    def launches(self):
        return int(self.lib.fgls_launches(self.h))

    def ctx(self):
        return C.c_void_p(self.lib.fgls_ctx(self.h))
