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


def capsule_fibers(n, seed=0, vol_frac=0.15, diameter_vox=8.0, aspect=10.0, acg=(0.7, 0.2, 0.1), max_tries=20000,
                   binary=True):
    """Periodic, non-overlapping capsules by seeded random sequential addition (PCG64); orientations
    from an angular central Gaussian with diag(acg).  Returns (phi_fibre, n_fibres).
    Rasterised by the voxel-centre test in a local window per fibre (fast enough for 256^3)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nx, ny, nz = n
    N = np.array(n, dtype=float)
    R = 0.5 * diameter_vox
    Lc = max(aspect * diameter_vox - 2 * R, 0.0)          # length of the cylindrical part (voxels)
    Lc = min(Lc, 0.9 * min(n))
    vol_one = np.pi * R * R * Lc + 4.0 / 3.0 * np.pi * R ** 3
    target = vol_frac * nx * ny * nz
    phi = np.zeros(n, dtype=np.float64)
    segs = []
    A = np.sqrt(np.asarray(acg, dtype=float))
    placed = 0.0
    tries = 0

    t9 = None
    Cs = np.zeros((0, 3))
    Ds = np.zeros((0, 3))
    while placed < target and tries < max_tries:
        tries += 1
        c = rng.random(3) * N
        d = rng.standard_normal(3) * A
        d /= np.linalg.norm(d)
        if len(segs):
            # sampled segment-segment distance (9 x 9 points), minimum image, against all placed fibres at once
            if t9 is None:
                t9 = np.linspace(-0.5 * Lc, 0.5 * Lc, 9)
            delta = Cs - c
            delta -= N * np.round(delta / N)
            a = t9[:, None] * d[None, :]                                   # (9,3) relative to c
            b = delta[:, None, :] + t9[None, :, None] * Ds[:, None, :]      # (m,9,3)
            dd = a[None, :, None, :] - b[:, None, :, :]
            if (dd * dd).sum(axis=3).min() < (2 * R + 1.0) ** 2:
                continue
        Cs = np.vstack([Cs, c])
        Ds = np.vstack([Ds, d])
        segs.append((c, d))
        placed += vol_one
        # rasterise in the bounding window
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
    return phi, len(segs)
