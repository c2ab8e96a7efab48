"""x-slab partition plan of the slab-decomposed solve loop -- the host-side index arithmetic shared by
bench.py, the multi-GPU checks and the CPU (gloo) tests.  It mirrors what csrc/comm.cu does on the device:

    rank r owns x planes [r*nx/P, (r+1)*nx/P) of every component (x is the slowest index, fg:230);
    the half spectra are exchanged as chunks (component c, peer q) of shape [lnx][lny][nzcs]:
        staging  S[c][q][il][jl][k]   <- y pass of the local slab, j = q*lny + jl
        y-slab   R[c][ii][jl][k]      <- chunk (c, q) of rank q lands at ii = q*lnx + il
"""
from __future__ import annotations

import numpy as np


def slab(nx, rank, nranks):
    if nx % nranks:
        raise ValueError("nx must be divisible by the number of ranks")
    lnx = nx // nranks
    return rank * lnx, (rank + 1) * lnx


def split_x(a, nranks, axis=-3):
    """split a (..., nx, ny, nz) array into the per-rank slabs"""
    return np.split(a, nranks, axis=axis)


def to_staging(yhat_local, nranks):
    """(C, lnx, ny, nzc) slab spectrum after the y pass -> staging layout (C, P, lnx, lny, nzc)"""
    C, lnx, ny, nzc = yhat_local.shape
    lny = ny // nranks
    return np.ascontiguousarray(yhat_local.reshape(C, lnx, nranks, lny, nzc).transpose(0, 2, 1, 3, 4))


def from_staging(stg):
    C, P, lnx, lny, nzc = stg.shape
    return np.ascontiguousarray(stg.transpose(0, 2, 1, 3, 4).reshape(C, lnx, P * lny, nzc))


def alltoall_numpy(stagings):
    """reference all-to-all over a list of per-rank staging arrays -> list of y-slab arrays R (C, nx, lny, nzc)"""
    P = len(stagings)
    out = []
    for me in range(P):
        C, _, lnx, lny, nzc = stagings[me].shape
        R = np.empty((C, P, lnx, lny, nzc), dtype=stagings[me].dtype)
        for q in range(P):
            R[:, q] = stagings[q][:, me]
        out.append(R.reshape(C, P * lnx, lny, nzc))
    return out
