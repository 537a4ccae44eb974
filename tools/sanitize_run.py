"""Small end-to-end run for compute-sanitizer: solve + simulate three fixtures (incl. folds and multiple states)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from tests import goldens
from tests.parity import solution_errors
for name in ("retirement1", "occ3", "model2"):
    m = goldens.model_for(name); m.compile(); m.solve()
    g = goldens.load(name)
    e = solution_errors(m.M, m.D, g["M"], g["D"])
    m.sim(g["init"], "own_shocks", randstream=g["randstream"])
    lib = m._capi()
    s, mom = lib.simulate_philox(m, m._solution, g["init"], 5, want_sims=True, want_moments=True)
    print(name, m._solution.status(), "%.1e %.1e" % (e["C"], e["V"]), goldens.sims_errors(m.sims, g["sims"])["max"], float(np.nansum(mom[2])))
