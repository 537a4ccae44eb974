"""Emulated solve + simulation of the fixtures against the golden vectors (kernel-logic check without a GPU)."""
import os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
from build import build
from egdst_b200 import capi
from tests import goldens
from tests.parity import solution_errors

def run(name, solve=True):
    m = goldens.model_for(name); m.prepare()
    g = goldens.load(name)
    lib = capi.ModelLibrary(build(m))
    out = {}
    if solve:
        t = time.time(); sol = lib.solve(m); out["solve_s"] = round(time.time() - t, 1)
        e = solution_errors(sol.M, sol.D, g["M"], g["D"])
        out["solve"] = {k: (("%.1e" % v) if isinstance(v, float) else v) for k, v in e.items() if k != "where"}
        out["status"] = sol.status()
    else:
        sol = lib.import_solution(m, g["M"], g["D"])
    sims = lib.simulate(m, sol, g["init"], g["randstream"], 0)
    out["sim"] = goldens.sims_errors(sims, g["sims"], g["skipcols"])
    s2, mom = lib.simulate_philox(m, sol, g["init"], 7, want_sims=True, want_moments=True)
    import numpy as np
    ref1 = np.nansum(s2, axis=0).T  # [nso, nt]
    refn = (~np.isnan(s2)).sum(axis=0).T
    with np.errstate(invalid="ignore"):
        fin = np.isfinite(ref1) & np.isfinite(mom[0])
        out["mom_err"] = float(np.max(np.abs(mom[0][fin] - ref1[fin])) if fin.any() else 0), float(np.abs(mom[2] - refn).max())
    print(name, out, flush=True)

if __name__ == "__main__":
    args = sys.argv[1:] or ["retirement2"]
    solve = "--nosolve" not in args
    for n in [a for a in args if not a.startswith("--")]:
        run(n, solve)
