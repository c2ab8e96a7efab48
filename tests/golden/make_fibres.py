"""Writes tests/golden/fibres_c2_seed0.npz: the capsule list of BASELINE config 2 (periodic RSA, PCG64 seed 0, 256^3
cell, D = 8 voxels, L/D = 10, 15 vol-%, ACG diag(0.7, 0.2, 0.1)).   python tests/golden/make_fibres.py   (about 100 s)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from microstructures import rsa_capsules  # noqa: E402

if __name__ == "__main__":
    Cs, Ds, R, Lc = rsa_capsules((256, 256, 256), seed=0, vol_frac=0.15, diameter_vox=8.0, aspect=10.0, acg=(0.7, 0.2, 0.1), max_tries=60000)
    vol = (np.pi * R * R * Lc + 4.0 / 3.0 * np.pi * R ** 3) * len(Cs) / 256.0 ** 3
    np.savez_compressed(os.path.join(HERE, "fibres_c2_seed0.npz"), centres=Cs, axes=Ds, R=R, Lc=Lc, n=256)
    print("%d capsules, volume fraction %.4f" % (len(Cs), vol))
