"""Seeded synthetic microstructures shared by the tests, smoke() and bench.py."""
import numpy as np


def sphere_phi(n, R=0.25, sub=1, center=(0.5, 0.5, 0.5)):
    """volume fraction of a centred sphere per voxel of an n[0] x n[1] x n[2] grid of the unit cell;
    sub > 1 gives composite voxels by sub-sampling (phi in [0,1]), sub == 1 a binary field."""
    nx, ny, nz = n
    phi = np.zeros(n)
    for a in range(sub):
        for b in range(sub):
            for c in range(sub):
                x = (np.arange(nx)[:, None, None] + (a + 0.5) / sub) / nx - center[0]
                y = (np.arange(ny)[None, :, None] + (b + 0.5) / sub) / ny - center[1]
                z = (np.arange(nz)[None, None, :] + (c + 0.5) / sub) / nz - center[2]
                phi += (x * x + y * y + z * z <= R * R)
    return phi / sub ** 3


def sphere_normals(n, center=(0.5, 0.5, 0.5)):
    nx, ny, nz = n
    x = (np.arange(nx)[:, None, None] + 0.5) / nx - center[0]
    y = (np.arange(ny)[None, :, None] + 0.5) / ny - center[1]
    z = (np.arange(nz)[None, None, :] + 0.5) / nz - center[2]
    v = np.stack(np.broadcast_arrays(x, y, z)).astype(float)
    r = np.sqrt((v * v).sum(axis=0))
    r[r == 0] = 1.0
    return v / r


def rsa_capsules(n, seed=0, vol_frac=0.15, diameter_vox=8.0, aspect=10.0, acg=(0.7, 0.2, 0.1), max_tries=20000):
    """Periodic, non-overlapping capsules by seeded random sequential addition (PCG64); orientations from an angular central
    Gaussian with diag(acg).  Voxel units.  Returns (centres (m,3), unit axes (m,3), radius, length of the cylindrical part)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nx, ny, nz = n
    N = np.array(n, dtype=float)
    R = 0.5 * diameter_vox
    Lc = max(aspect * diameter_vox - 2 * R, 0.0)          # length of the cylindrical part (voxels)
    Lc = min(Lc, 0.9 * min(n))
    vol_one = np.pi * R * R * Lc + 4.0 / 3.0 * np.pi * R ** 3
    target = vol_frac * nx * ny * nz
    A = np.sqrt(np.asarray(acg, dtype=float))
    placed = 0.0
    tries = 0
    t9 = np.linspace(-0.5 * Lc, 0.5 * Lc, 9)
    Cs = np.zeros((0, 3))
    Ds = np.zeros((0, 3))
    while placed < target and tries < max_tries:
        tries += 1
        c = rng.random(3) * N
        d = rng.standard_normal(3) * A
        d /= np.linalg.norm(d)
        if len(Cs):
            # sampled segment-segment distance (9 x 9 points), minimum image, against all placed fibres at once
            delta = Cs - c
            delta -= N * np.round(delta / N)
            a = t9[:, None] * d[None, :]                                   # (9,3) relative to c
            b = delta[:, None, :] + t9[None, :, None] * Ds[:, None, :]      # (m,9,3)
            dd = a[None, :, None, :] - b[:, None, :, :]
            if (dd * dd).sum(axis=3).min() < (2 * R + 1.0) ** 2:
                continue
        Cs = np.vstack([Cs, c])
        Ds = np.vstack([Ds, d])
        placed += vol_one
    return Cs, Ds, R, Lc


def fiber_list(n, Cs, Ds, R, Lc, material=1, tile=(1, 1, 1)):
    """<place_fiber> style list [(centre, axis, L0, R, material)] in the unit cell [0,1)^3 * tile/max(tile) ... of a grid with n voxels
    per cell, including the periodic images that reach into the (tiled) cell; L0 = Lc + 4R/3 is CapsuleFiber's total length
    (fg:5258).  tile repeats the cell (the microstructure of a tiled grid is the periodic continuation of the cell)."""
    N = np.array(n, dtype=float)
    T = np.array(tile, dtype=int)
    h = 1.0 / (N[0] * T[0])                      # voxel size: the tiled grid spans a unit length along x
    box = N * T * h
    out = []
    reach = 0.5 * Lc + R + 2.0                   # voxels
    for c, d in zip(Cs, Ds):
        for ti in range(-1, T[0] + 1):
            for tj in range(-1, T[1] + 1):
                for tk in range(-1, T[2] + 1):
                    cc = c + N * np.array([ti, tj, tk])
                    if np.any(cc < -reach) or np.any(cc > N * T + reach):
                        continue
                    out.append((tuple(cc * h), tuple(d), (Lc + 4.0 / 3.0 * R) * h, R * h, material))
    return out, tuple(box)


def capsule_fibers(n, seed=0, vol_frac=0.15, diameter_vox=8.0, aspect=10.0, acg=(0.7, 0.2, 0.1), max_tries=20000,
                   binary=True):
    """rsa_capsules rasterised by the voxel-centre test in a local window per fibre.  Returns (phi_fibre, n_fibres)."""
    nx, ny, nz = n
    Cs, Ds, R, Lc = rsa_capsules(n, seed, vol_frac, diameter_vox, aspect, acg, max_tries)
    phi = np.zeros(n, dtype=np.float64)
    for c, d in zip(Cs, Ds):
        half = 0.5 * Lc * np.abs(d) + R + 1
        lo = np.floor(c - half).astype(int)
        hi = np.ceil(c + half).astype(int) + 1
        ii = np.arange(lo[0], hi[0])
        jj = np.arange(lo[1], hi[1])
        kk = np.arange(lo[2], hi[2])
        X = (ii[:, None, None] + 0.5) - c[0]
        Y = (jj[None, :, None] + 0.5) - c[1]
        Z = (kk[None, None, :] + 0.5) - c[2]
        t = np.clip(X * d[0] + Y * d[1] + Z * d[2], -0.5 * Lc, 0.5 * Lc)
        dist2 = (X - t * d[0]) ** 2 + (Y - t * d[1]) ** 2 + (Z - t * d[2]) ** 2
        inside = dist2 <= R * R
        I, J, K = np.nonzero(inside)
        phi[ii[I] % nx, jj[J] % ny, kk[K] % nz] = 1.0
    return phi, len(Cs)


def config2_fibres():
    """the fibre list of BASELINE config 2 (256^3 cell, D = 8 voxels, L/D = 10, 15 vol-%, ACG diag(.7,.2,.1), PCG64 seed 0):
    committed fixture tests/golden/fibres_c2_seed0.npz (written by tests/golden/make_fibres.py, 100 s of RSA)"""
    import os
    f = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fibres_c2_seed0.npz"))
    return f["centres"], f["axes"], float(f["R"]), float(f["Lc"])
