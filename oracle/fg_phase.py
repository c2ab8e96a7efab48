"""ctypes view of oracle/fg_phase.c: fibergen's composite-voxel phase initialisation (initPhi fg:17489, integratePhiVoxel
fg:16622, halfspace_box_cut_volume fg:1385, CapsuleFiber fg:5237), restated in C.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfgoracle.so")
_lib = None


class Capsule(C.Structure):
    _fields_ = [("c1", C.c_double * 3), ("a", C.c_double * 3), ("r", C.c_double * 3), ("R", C.c_double), ("L", C.c_double),
                ("mat", C.c_int)]


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        lib.fgo_capsule_init.argtypes = [C.POINTER(Capsule), dp, dp, C.c_double, C.c_double, C.c_int]
        lib.fgo_capsule_init.restype = None
        lib.fgo_capsule_distance.argtypes = [C.POINTER(Capsule), dp, dp]
        lib.fgo_capsule_distance.restype = C.c_double
        lib.fgo_halfspace_box_cut_volume.argtypes = [dp, dp, dp, C.c_double, C.c_double, C.c_double]
        lib.fgo_halfspace_box_cut_volume.restype = C.c_double
        lib.fgo_init_phi.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, C.POINTER(Capsule), C.c_int, C.c_int, C.c_int, C.c_double,
                                     C.c_int, C.c_int, C.POINTER(dp)]
        lib.fgo_init_phi.restype = C.c_long
        _lib = lib
    return _lib


def _v3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def halfspace_box_cut_volume(x, n, x0, dx, dy, dz):
    return load().fgo_halfspace_box_cut_volume(_v3(x), _v3(n), _v3(x0), dx, dy, dz)


def capsules(fibers):
    """fibers: iterable of (centre, axis, L0, R, material) as the <place_fiber> action takes them (fg:25789-25823)"""
    lib = load()
    fibers = list(fibers)
    arr = (Capsule * max(len(fibers), 1))()
    for i, (c, a, L0, R, mat) in enumerate(fibers):
        lib.fgo_capsule_init(C.byref(arr[i]), _v3(c), _v3(a), float(L0), float(R), int(mat))
    return arr, len(fibers)


def init_phi(n, L, fibers, nmat, matrix_mat=0, smooth_levels=-1, smooth_tol=0.001, x0=(0., 0., 0.), rows=None):
    """LSSolver::initPhi for capsule fibres; returns phi[nmat, rows, ny, nz] (rows = (i0, i1), default the whole grid) and the
    number of interface voxels.  Defaults smooth_levels = -1, smooth_tol = 1e-3 are the reference's (fg:14842-14843)."""
    lib = load()
    arr, nf = capsules(fibers)
    i0, i1 = (0, n[0]) if rows is None else rows
    phi = np.zeros((nmat, i1 - i0, n[1], n[2]))
    dp = C.POINTER(C.c_double)
    ptrs = (dp * nmat)(*[phi[m].ctypes.data_as(dp) for m in range(nmat)])
    cnt = lib.fgo_init_phi(n[0], n[1], n[2], _v3(L), _v3(x0), nf, arr, nmat, matrix_mat, smooth_levels, smooth_tol, i0, i1, ptrs)
    return phi, int(cnt)
