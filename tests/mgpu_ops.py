"""Operator-level multi-GPU checks against the CPU oracle (torchrun, one rank per GPU)."""
import os
import sys
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fibergen_b200 as fb
from fibergen_b200.partition import slab
from oracle import fg_oracle as fo


def comm_init(ctx, rank):
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = C.create_string_buffer(128)
        assert ctx.lib.fgb_comm_unique_id(raw) == 0
        buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    idb = C.create_string_buffer(bytes(buf.cpu().numpy().tobytes()), 128)
    ctx.chk(ctx.lib.fgb_comm_init(ctx.h, idb))


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bad = 0
    for n, L in [((8, 4, 6), (1., 1., 1.)), ((16, 12, 10), (1., 2., 1.5)), ((64, 64, 32), (1., 1., 1.))]:
        x0, x1 = slab(n[0], rank, world)
        for mode, d in (("heat", 3), ("elasticity", 6), ("hyperelasticity", 9)):
            for scheme in ("collocated", "staggered"):
                rng = np.random.default_rng(7)
                ctx = fb.Context(*n, *L, mode=mode, gamma_scheme=scheme, device=local, rank=rank, nranks=world)
                comm_init(ctx, rank)
                o = fo.LSSolver(*n, *L, mode=mode, gamma_scheme=scheme)
                o.set_reference(1.3, 0.4)
                o.setBCProjector(fo.Id4(d))
                tau = rng.standard_normal((d,) + n)
                E = rng.standard_normal(d)
                f = ctx.field(tau[:, x0:x1])
                errs = {}
                errs["avg"] = relerr(ctx.average(f), o.average(tau))
                errs["inner"] = abs(ctx.inner(f, f) - o.innerProduct(tau, tau)) / o.innerProduct(tau, tau)
                if scheme == "staggered":
                    ctx.chk(ctx.lib.fgb_div_staggered(ctx.h, f))
                    errs["div"] = relerr(ctx.u_download(), o.divOperatorStaggered(tau)[:, x0:x1])
                    u = rng.standard_normal((ctx.udim,) + n)
                    ctx.u_upload(u[:, x0:x1])
                    f2 = ctx.field()
                    ctx.chk(ctx.lib.fgb_eps_staggered(ctx.h, f2, fb.solver._dp(ctx.vec(E))))
                    errs["eps"] = relerr(ctx.download(f2), o.epsOperatorStaggered(E, u)[:, x0:x1])
                ctx.gamma(f, E, 1.3, 0.4, -1.0, 0.0)
                errs["gamma"] = relerr(ctx.download(f), o.GammaOperator(E, 1.3, 0.4, tau, -1.0, 0.0)[:, x0:x1])
                worst = max(errs.values())
                if worst > 2e-12:
                    bad += 1
                print("rank %d %s %-16s %-10s " % (rank, n, mode, scheme) + " ".join("%s=%.1e" % kv for kv in errs.items()), flush=True)
                ctx.close()
                dist.barrier()
    t = torch.tensor([bad], device="cuda")
    dist.all_reduce(t)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 0 else 1)


if __name__ == "__main__":
    main()
