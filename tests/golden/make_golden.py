"""Writes tests/golden/golden_r02.npz from the CPU oracle:  python tests/golden/make_golden.py
(run in the build container; a few seconds).  See cases.py for what a case is."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from oracle import fg_oracle as fo  # noqa: E402
import cases as gc  # noqa: E402


def solve(case):
    o = fo.LSSolver(*case["n"], mode=case["mode"], **case["settings"])
    for name, law, params, phi in case["phases"]:
        o.add_phase(name, gc.oracle_law(fo, case["mode"], law, params), phi)
    if case["normals"] is not None:
        o.set_normals(case["normals"])
    if case["ref"] is not None:
        o.set_reference(*case["ref"])
    o.setStrain(case["E"])
    o.run()
    return o


def main():
    out = {}
    for case in gc.cases():
        o = solve(case)
        k = case["key"]
        mean, rms, samples = gc.summarize(o.epsilon, case["n"])
        out[k + "/residuals"] = np.array(o.residuals)
        out[k + "/mean_stress"] = np.asarray(o.calcMeanStress())
        out[k + "/mu0"] = np.array([o.mu_0, o.lambda_0])
        out[k + "/eps_mean"] = mean
        out[k + "/eps_rms"] = rms
        out[k + "/eps_samples"] = samples
        sig = o.calcStress(0.0, 0.0, o.epsilon)
        out[k + "/sigma_rms"] = np.sqrt((sig.reshape(sig.shape[0], -1) ** 2).mean(axis=1))
        u = o.calcDisplacement()
        out[k + "/u_rms"] = np.sqrt((u.reshape(u.shape[0], -1) ** 2).mean(axis=1))
        print("%-28s iterations %3d  last residual %.3e  <sigma> %s" % (k, len(o.residuals), o.residuals[-1], out[k + "/mean_stress"][:3]))
    np.savez_compressed(gc.FIXTURE, **out)
    print("wrote", gc.FIXTURE, os.path.getsize(gc.FIXTURE), "bytes")


if __name__ == "__main__":
    main()
