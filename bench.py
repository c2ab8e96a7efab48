#!/usr/bin/env python
"""bench.py -- voxel-iterations/s of the Lippmann-Schwinger CG iteration on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on the host cores

A "step" is one CG iteration (runCGElasticity hot loop, fg:23206-23246) = one pass of the hot path
(material law -> div -> 3-D FFT -> G0 -> inverse FFT -> sym-grad -> dots -> vector updates) over the grid.
Workload at N=1: BASELINE config 2 -- 256^3 short-fibre composite, linear elasticity, CG, staggered grid,
Voigt mixing, residual estimator.  N>1: weak scaling, 256^3 voxels per GPU, global grid 256x512x256 / 256x512x512 /
512x512x512 for N = 2 / 4 / 8 (periodic tiling of the same cell), x-slab partition.
Timing: CUDA events on the launching stream, W warm-up iterations, K timed, barrier + synchronize on both
sides, max over ranks.  Inputs (3.5 GB of fields) are far larger than the 126 MB L2, so no explicit flush.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

METRIC = "voxel-iterations/sec"
UNIT = "voxel-iterations/s"
B_ALG_CG_STAGGERED = 728.0        # algorithmic bytes per voxel-iteration, SURVEY.md 8(d) / BASELINE.md section 2

# matrix / fibre of demo/elasticity/sfrp_parameter_fit
E_M, NU_M, E_F, NU_F = 1.665, 0.36, 73.0, 0.18

# algorithmic bytes per voxel of each kernel of the unfused iteration (d=6 tensor comps, u=3 vector comps, 2 phases)
KERNEL_BYTES_PER_VOXEL = {
    "calc_stress": (6 + 1 + 6) * 8, "div_staggered": (6 + 3) * 8, "fft_z_r2c": 2 * 3 * 8, "fft_y_fwd": 2 * 3 * 8,
    "fft_x_green": 2 * 3 * 8, "fft_y_bwd": 2 * 3 * 8, "fft_z_c2r": 2 * 3 * 8, "eps_staggered": (3 + 6) * 8,
    "inner_product": 2 * 6 * 8 + 6 * 8, "cg_update": 6 * 6 * 8, "xpay": 3 * 6 * 8,
    "cg_direction_stress_div": (12 + 1 + 6 + 3) * 8, "stress_div": (6 + 1 + 3) * 8, "eps_dot": (3 + 6 + 6) * 8,
    # implicit operator result (FGB_W_IMPLICIT): the sum reads u and p only, the update reads x, r, p, u and writes x, r
    "eps_dot_implicit": (3 + 6) * 8, "cg_update_implicit": (3 * 6 + 3 + 2 * 6) * 8,
}


def lame(E, nu):
    return E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.gpu = gpu
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def microstructure(n, seed=0):
    from microstructures import capsule_fibers
    phi, nf = capsule_fibers(n, seed=seed, vol_frac=0.15, diameter_vox=8.0, aspect=10.0, acg=(0.7, 0.2, 0.1), max_tries=6000)
    return phi, nf


def tile_x(phi, reps):
    return np.concatenate([phi] * reps, axis=0) if reps > 1 else phi


# ------------------------------------------------------------------------------------------------ CUDA arm
def run_cuda(args):
    import torch
    import fibergen_b200 as fb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA arm has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    base = args.grid
    # weak scaling: every rank owns base^3 voxels; the global grid grows along y, z and x in turn so that no axis exceeds
    # 2*base (x-slabs of nx/P planes; nx and ny must be divisible by P)
    mult = {1: (1, 1, 1), 2: (1, 2, 1), 4: (1, 2, 2), 8: (2, 2, 2)}.get(world, (world, 1, 1))
    if base * mult[0] % world or base * mult[1] % world:
        mult = (world, 1, 1)
    n = (base * mult[0], base * mult[1], base * mult[2])
    nxyz = n[0] * n[1] * n[2]
    K, W = args.steps, max(args.warmup, 3)

    phi_cell, nfib = microstructure((base, base, base))
    lnx = n[0] // world
    # the global microstructure is the periodic tiling of the cell; this rank's x-slab of it
    x0 = rank * lnx
    reps = (mult[0], mult[1], mult[2])
    phi_local = np.tile(phi_cell, reps)[x0:x0 + lnx]
    vf = float(phi_cell.mean())
    lam_m, mu_m = lame(E_M, NU_M)
    lam_f, mu_f = lame(E_F, NU_F)

    def make_solver(tol, maxiter):
        s = fb.LSSolver(*n, rank=rank, nranks=world, device=local_rank, mode="elasticity", method="cg", gamma_scheme="staggered",
                        mixing_rule="voigt", error_estimator="residual", tol=tol, maxiter=maxiter)
        s.add_material("matrix", "iso", mu_m, lam_m)
        s.add_material("fibre", "iso", mu_f, lam_f)
        s.init()
        if world > 1:
            idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                import ctypes as C
                raw = C.create_string_buffer(128)
                rc = s.lib.fgb_comm_unique_id(raw)
                if rc:
                    raise SystemExit("fgb_comm_unique_id failed: %d" % rc)
                idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
            dist.broadcast(idbuf, 0)
            s.init_comm(bytes(idbuf.cpu().numpy().tobytes()))
        return s

    # pinned host staging of the phase planes (padded reference layout)
    nzp = fb.nzp_of(n[2])
    host_phi = torch.empty((2, lnx, n[1], nzp), dtype=torch.float64).pin_memory()
    hp = host_phi.numpy()
    hp[...] = 0
    hp[0, :, :, :n[2]] = 1 - phi_local
    hp[1, :, :, :n[2]] = phi_local
    host_eps = torch.empty((6, lnx, n[1], nzp), dtype=torch.float64).pin_memory()

    stream = torch.cuda.current_stream()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing: W warm-up + K timed CG iterations ----------------
    s = make_solver(tol=1e-300, maxiter=10 ** 6)
    s.lib.fgb_set_stream(s.ctx(), stream.cuda_stream)
    s.set_phase(0, hp[0], padded=True)
    s.set_phase(1, hp[1], padded=True)
    s.set_strain([1, 0, 0, 0, 0, 0])
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    state = {"n": 0, "launch0": 0, "launch1": 0}
    sampler = ClockSampler(local_rank)
    ctxp = s.ctx()

    def cb():
        state["n"] += 1
        if state["n"] == W:
            barrier()
            if rank == 0:
                sampler.start()
            s.lib.fgb_profile_enable(ctxp, 1)
            state["launch0"] = s.launches()
            ev0.record(stream)
        if state["n"] == W + K:
            ev1.record(stream)
            barrier()
            state["launch1"] = s.launches()
            return True
        return False

    s.set_convergence_callback(cb)
    s.run()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = nxyz * K / (ms * 1e-3)
    launches = state["launch1"] - state["launch0"]

    # per-kernel durations from the library's own events (same stream, same timed region)
    ctxobj = fb.Context.__new__(fb.Context)
    ctxobj.lib, ctxobj.h = s.lib, ctxp
    prof = fb.Context.profile_results(ctxobj)
    ctxobj.h = None          # borrowed handle: the solver owns the context
    s.lib.fgb_profile_enable(ctxp, 0)
    peak, peak_src = peaks()
    nloc = nxyz // world
    kernels = {}
    for name, (tot_ms, cnt) in prof.items():
        if cnt == 0:
            continue
        avg = tot_ms / cnt
        bpv = KERNEL_BYTES_PER_VOXEL.get(name)
        kernels[name] = {"launches": int(cnt), "avg_ms": avg, "share": tot_ms / ms,
                         "gbs": (bpv * nloc / (avg * 1e-3) / 1e9) if bpv else None}
    top = max(kernels, key=lambda k: kernels[k]["avg_ms"] * kernels[k]["launches"]) if kernels else None
    roofline = None
    if top and kernels[top]["gbs"]:
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(top)
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "kernel": top, "achieved": kernels[top]["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": kernels[top]["gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": KERNEL_BYTES_PER_VOXEL[top] * nloc}
    iter_frac = B_ALG_CG_STAGGERED * value / (world * peak * 1e9)
    s.set_convergence_callback(None)
    s.close()

    # ---------------- end to end: host phase planes in, strain field out, complete solve ----------------
    # A fresh solver pays one-time costs on its first run (field allocation, peer-memory mapping, NCCL connection set-up) that a
    # user amortises over the load cases of one job (calc_effective_properties runs 6): warm the solver with a 3-iteration run,
    # then time a complete cold-data solve: phase planes from pinned host memory in, converged strain field out.
    s2 = make_solver(tol=1e-6, maxiter=3)
    s2.lib.fgb_set_stream(s2.ctx(), stream.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.set_strain([1, 0, 0, 0, 0, 0])
    s2.set_phase(0, hp[0], padded=True)
    s2.set_phase(1, hp[1], padded=True)
    s2.run()
    s2.get_field("epsilon", padded=True, out=host_eps.numpy())
    s2.set("maxiter", args.e2e_maxiter)
    barrier()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    s2.set_phase(0, hp[0], padded=True)            # H2D from pinned memory
    s2.set_phase(1, hp[1], padded=True)
    ea.record(stream)
    s2.run()
    eb.record(stream)
    s2.get_field("epsilon", padded=True, out=host_eps.numpy())      # D2H of the solution
    sm = s2.get_mean_stress()
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    e2e_parts = {"h2d_ms": e0.elapsed_time(ea), "solve_ms": ea.elapsed_time(eb), "d2h_and_mean_stress_ms": eb.elapsed_time(e1)}
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    iters = len(s2.get_residuals())
    res_last = float(s2.get_residuals()[-1])
    e2e = {"value": nxyz * iters / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(host_phi.numel() * 8 * world / iters),
           "d2h_bytes_per_step": int(host_eps.numel() * 8 * world / iters), "iterations": iters, "ms_total": e2e_ms,
           "final_residual": res_last, "mean_stress_11": float(sm[0]), "parts": e2e_parts,
           "what": "fgls (warm solver): set_phase (H2D, pinned) + run() to tol 1e-6 + get_field('epsilon') (D2H) + mean stress"}
    s2.close()

    # ---------------- CPU baseline: the oracle port on this host, bounded sample ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(base, phi_cell, iters=2)

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": "config 2: short-fibre composite %dx%dx%d, linear elasticity, CG, staggered grid, Voigt mixing, "
                                      "residual estimator" % n, "grid": list(n), "fibres_per_cell": nfib, "fibre_volume_fraction": vf,
                          "partition": "x-slabs, %d rank(s)" % world, "l2": "inputs larger than L2 (3.5 GB of fields vs 126 MB), no flush",
                          "step": "one CG iteration"},
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
               "iteration_hbm": {"algorithmic_bytes_per_voxel_iteration": B_ALG_CG_STAGGERED, "achieved_gbs_per_gpu": B_ALG_CG_STAGGERED * value / world / 1e9,
                                 "frac_of_peak": iter_frac, "peak_gbs": peak, "peak_source": peak_src},
               "kernels": kernels, "cpu_baseline": cpu}
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ CPU arm
def oracle_solver(n, phi, tol=1e-300, maxiter=10 ** 6):
    from oracle import fg_oracle as fo
    lam_m, mu_m = lame(E_M, NU_M)
    lam_f, mu_f = lame(E_F, NU_F)
    o = fo.LSSolver(*n, mode="elasticity", method="cg", gamma_scheme="staggered", mixing_rule="voigt",
                    error_estimator="residual", tol=tol, maxiter=maxiter)
    o.add_phase("matrix", fo.LinearIsotropic(mu_m, lam_m), 1 - phi)
    o.add_phase("fibre", fo.LinearIsotropic(mu_f, lam_f), phi)
    o.setStrain([1, 0, 0, 0, 0, 0])
    return o


def time_oracle_iterations(n, phi, warm, steps):
    """times `steps` CG iterations of the oracle after `warm` untimed ones (callback = reference's convergence callback)"""
    o = oracle_solver(n, phi)
    st = {"n": 0, "t0": None, "t1": None}

    def cb():
        st["n"] += 1
        if st["n"] == warm:
            st["t0"] = time.perf_counter()
        if st["n"] == warm + steps:
            st["t1"] = time.perf_counter()
            return True
        return False
    o.callback = cb
    # the reference-material scan is excluded from the loop timing
    o.run()
    return st["t1"] - st["t0"]


def cpu_baseline(base, phi, iters=2):
    cores = os.cpu_count() or 1
    n = (base, base, base)
    dt = time_oracle_iterations(n, phi, 1, iters)
    return {"value": n[0] * n[1] * n[2] * iters / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d CG iterations of the same %d^3 problem with the numpy/pocketfft oracle (oracle/fg_oracle.py), "
                      "FFT on all %d host threads, elementwise work single-threaded numpy" % (iters, base, cores)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    base = args.ref_grid
    n = (base, base, base)
    phi, nfib = microstructure(n)
    K, W = args.steps, max(args.warmup, 1)
    cores = os.cpu_count() or 1
    dt = time_oracle_iterations(n, phi, W, K)
    value = n[0] * n[1] * n[2] * K / dt
    sample = ("each step = one CG iteration of the same workload family on a %d^3 cell (bounded sample of config 2), numpy/pocketfft "
              "oracle port of the reference algorithm, FFT on %d host threads" % (base, cores))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
           "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "config 2: short-fibre composite, linear elasticity, CG, staggered grid, Voigt mixing, residual estimator",
                      "grid": list(n), "step": "one CG iteration"},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "the reference (Boost/FFTW3/LAPACK C++) cannot be built in this image; this is the repo's oracle port, not fibergen's own binary"}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--grid", type=int, default=256, help="cell edge per GPU (BASELINE config 2: 256)")
    ap.add_argument("--ref-grid", type=int, default=128, help="cell edge of the CPU arm's bounded sample")
    ap.add_argument("--e2e-maxiter", type=int, default=400)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
