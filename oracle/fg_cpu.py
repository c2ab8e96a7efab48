"""ctypes view of oracle/fg_cpu.cpp: OpenMP C++ restatement of the CG iteration (CPU baseline).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import fg_phase as _fp


def set_threads(n):
    """OpenMP threads used by libfgoracle (phase initialisation and CG restatement)"""
    lib = _fp.load()
    lib.fgcpu_set_threads.argtypes = [C.c_int]
    lib.fgcpu_set_threads.restype = None
    lib.fgcpu_set_threads(int(n))


def cg_iterations(n, L, phi_fibre, materials, E, warm=1, steps=2):
    """two-phase problem: matrix fraction 1 - phi_fibre, materials = ((mu_m, lam_m), (mu_f, lam_f)); returns a dict with the
    residual history of warm + steps iterations, the seconds of the last `steps`, the thread count, mean stress and mu_0"""
    lib = _fp.load()
    dp = C.POINTER(C.c_double)
    lib.fgcpu_cg_iterations.restype = C.c_int
    lib.fgcpu_cg_iterations.argtypes = [C.c_int, C.c_int, C.c_int, dp, C.c_int, C.POINTER(dp), dp, dp, dp, C.c_int, C.c_int, dp, dp,
                                        C.POINTER(C.c_int), dp, dp]
    phi1 = np.ascontiguousarray(phi_fibre, dtype=np.float64)
    phi0 = np.ascontiguousarray(1 - phi1)
    ptrs = (dp * 2)(phi0.ctypes.data_as(dp), phi1.ctypes.data_as(dp))
    mu = np.array([materials[0][0], materials[1][0]], dtype=np.float64)
    lam = np.array([materials[0][1], materials[1][1]], dtype=np.float64)
    Lv = np.array(L, dtype=np.float64)
    Ev = np.array(E, dtype=np.float64)
    res = np.zeros(warm + steps)
    sec, mu0 = C.c_double(), C.c_double()
    thr = C.c_int()
    ms = np.zeros(6)
    rc = lib.fgcpu_cg_iterations(n[0], n[1], n[2], Lv.ctypes.data_as(dp), 2, ptrs, mu.ctypes.data_as(dp), lam.ctypes.data_as(dp),
                                 Ev.ctypes.data_as(dp), warm, steps, res.ctypes.data_as(dp), C.byref(sec), C.byref(thr), ms.ctypes.data_as(dp),
                                 C.byref(mu0))
    if rc:
        raise ValueError("fg_cpu: axis lengths must be powers of two")
    return {"residuals": res, "seconds": sec.value, "threads": thr.value, "mean_stress": ms, "mu_0": mu0.value}
