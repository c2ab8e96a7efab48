"""Multi-GPU parity check, launched by torchrun (one rank per GPU).  Every rank solves its x-slab of the SAME global
problem with the slab-partitioned library path; rank 0 additionally solves the whole problem on one GPU and both are
compared with each other (and, for the small grid, with the CPU oracle)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C

import fibergen_b200 as fb
from fibergen_b200.partition import slab
from microstructures import sphere_phi, sphere_normals


def unique_id(s, rank):
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = C.create_string_buffer(128)
        rc = s.lib.fgb_comm_unique_id(raw)
        assert rc == 0, rc
        buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())


def solve(n, rank, world, device, phis, mats, normals=None, E=None, **kw):
    s = fb.LSSolver(*n, rank=rank, nranks=world, device=device, **kw)
    for name, law, params in mats:
        s.add_material(name, law, *params)
    if kw.get("gamma_scheme") == "willot":
        s.set_reference(1.0, 2.0)
    s.init()
    if world > 1:
        s.init_comm(unique_id(s, rank))
    x0, x1 = slab(n[0], rank, world)
    for m, phi in enumerate(phis):
        s.set_phase(m, phi[x0:x1])
    if normals is not None:
        s.set_normals(normals[:, x0:x1])
    s.set_strain(E)
    s.run()
    return s, (x0, x1)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    cases = [
        ("elasticity cg staggered 32^3", (32, 32, 32), dict(mode="elasticity", method="cg", error_estimator="residual", tol=1e-8), "el"),
        ("elasticity cg staggered 64x64x32 (pow2 fast path)", (64, 64, 32), dict(mode="elasticity", method="cg", error_estimator="residual", tol=1e-8), "el"),
        ("elasticity cg staggered 512x16x16 / 16x512x16 (three-pass FFT + peer stores)", (512, 16, 16), dict(mode="elasticity", method="cg", error_estimator="residual", tol=1e-8), "el"),
        ("elasticity cg staggered 16x512x16", (16, 512, 16), dict(mode="elasticity", method="cg", error_estimator="residual", tol=1e-8), "el"),
        ("elasticity cg staggered 16x8x300 (512-thread sweep with halo planes)", (16, 8, 300), dict(mode="elasticity", method="cg", error_estimator="residual", tol=1e-8), "el"),
        ("elasticity cg staggered 16x8x512 (half-length z transform)", (16, 8, 512), dict(mode="elasticity", method="cg", error_estimator="residual", tol=1e-8), "el"),
        ("elasticity cg collocated 64x32x16 (6-component three-pass x pass)", (64, 32, 16), dict(mode="elasticity", method="cg", gamma_scheme="collocated", error_estimator="residual", tol=1e-8), "el"),
        ("elasticity basic collocated 24x16x10", (24, 16, 10), dict(mode="elasticity", method="basic", gamma_scheme="collocated", error_estimator="sigma", tol=1e-7), "el"),
        ("elasticity cg staggered 1024x16x16 (64-byte peer tiles of the fused x pass)", (1024, 16, 16), dict(mode="elasticity", method="cg", error_estimator="residual", tol=1e-8), "el"),
        ("elasticity cg staggered 16x1024x16 (three-pass y transform with peer stores)", (16, 1024, 16), dict(mode="elasticity", method="cg", error_estimator="residual", tol=1e-8), "el"),
        ("elasticity cg willot 32x16x16", (32, 16, 16), dict(mode="elasticity", method="cg", gamma_scheme="willot", error_estimator="residual", tol=1e-8), "el"),
        ("viscosity cg collocated 16x12x10 (zero-trace transform)", (16, 12, 10), dict(mode="viscosity", method="cg", gamma_scheme="collocated", error_estimator="residual", tol=1e-7), "visc"),
        ("heat cg laminate 32x24x16", (32, 24, 16), dict(mode="heat", method="cg", mixing_rule="laminate", error_estimator="residual", tol=1e-8), "heat"),
        ("neo-hooke newton-cg 16^3", (16, 16, 16), dict(mode="hyperelasticity", method="cg", error_estimator="residual", outer_error_estimator="sigma", tol=1e-6), "nh"),
    ]
    for name, n, kw, kind in cases:
        if n[0] % world or n[1] % world:
            continue                               # the slab partition needs nx and ny divisible by the number of ranks
        phi = sphere_phi(n, R=0.3, sub=2)
        normals = None
        if kind == "el":
            lam1, mu1 = fb.lame(1.0, 0.3)
            lam2, mu2 = fb.lame(20.0, 0.3)
            mats = [("matrix", "iso", (mu1, lam1)), ("incl", "iso", (mu2, lam2))]
            E = [1, 0, 0, 0.5, 0, 0.2]
        elif kind == "visc":
            phi = sphere_phi(n, R=0.3, sub=1)
            mats = [("fluid", "iso", (1.0,)), ("solid", "iso", (1e-3,))]
            E = [0, 0, 0, 0, 0, 1.0]
        elif kind == "heat":
            mats = [("matrix", "iso", (1.0,)), ("incl", "iso", (10.0,))]
            normals = sphere_normals(n)
            E = [1, 0.3, 0]
        else:
            phi = sphere_phi(n, R=0.3, sub=1)
            mats = [("matrix", "nh", (10.0, 10.0)), ("incl", "nh", (10.0, 100.0))]
            E = [1, 1.1, 1, 0, 0, 0, 0, 0, 0]
        phis = [1 - phi, phi]
        s, (x0, x1) = solve(n, rank, world, local, phis, mats, normals, E, **kw)
        res = s.get_residuals()
        sm = s.get_mean_stress()
        eps = s.get_field()
        u = s.get_field("u")                     # derived field: staggered operators (hyperelasticity: the 9-component G0-div) on the slab
        if rank == 0:
            s1, _ = solve(n, 0, 1, local, phis, mats, normals, E, **kw)
            r1, m1, e1, u1 = s1.get_residuals(), s1.get_mean_stress(), s1.get_field(), s1.get_field("u")
            good = (len(res) == len(r1) and np.abs(res - r1).max() <= 1e-10 * np.abs(r1).max()
                    and np.abs(sm - m1).max() <= 1e-9 * np.abs(m1).max()
                    and np.abs(eps - e1[:, x0:x1]).max() <= 1e-9 * np.abs(e1).max()
                    and np.abs(u - u1[:, x0:x1]).max() <= 1e-8 * max(np.abs(u1).max(), 1e-300))
            print("%-55s ranks=%d iters %d/%d  |dres| %.2e  |dstress| %.2e  %s" % (
                name, world, len(res), len(r1), np.abs(res[:len(r1)] - r1[:len(res)]).max() if len(res) and len(r1) else -1,
                np.abs(sm - m1).max() / np.abs(m1).max(), "OK" if good else "MISMATCH"), flush=True)
            ok = ok and good
            s1.close()
        s.close()
        dist.barrier()
    # device phase initialisation on the slab against the single-GPU one
    from microstructures import rsa_capsules, fiber_list
    n = (32, 32, 32)
    Cs, Ds, R, Lc = rsa_capsules(n, seed=5, vol_frac=0.12, diameter_vox=5.0, aspect=3.0, max_tries=300)
    fibs, box = fiber_list(n, Cs, Ds, R, Lc, material=1)
    s = fb.LSSolver(*n, rank=rank, nranks=world, device=local, mode="heat", mixing_rule="laminate")
    s.add_material("m", "iso", 1.0)
    s.add_material("f", "iso", 10.0)
    s.init()
    s.init_comm(unique_id(s, rank))
    s.init_phase(fibs, normals=True)
    x0, x1 = slab(n[0], rank, world)
    mine = np.stack([s.get_phase(m) for m in range(2)])
    if rank == 0:
        s1 = fb.LSSolver(*n, device=local, mode="heat", mixing_rule="laminate")
        s1.add_material("m", "iso", 1.0)
        s1.add_material("f", "iso", 10.0)
        s1.init()
        s1.init_phase(fibs, normals=True)
        full = np.stack([s1.get_phase(m) for m in range(2)])
        good = np.array_equal(mine, full[:, x0:x1]) and 0.05 < full[1].mean() < 0.2
        print("%-55s ranks=%d  %s" % ("device phase initialisation on the slab", world, "OK" if good else "MISMATCH"), flush=True)
        ok = ok and good
        s1.close()
    s.close()
    dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
