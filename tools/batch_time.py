"""S3: batched estimation sweep -- nvec deaton2 parameter vectors solved in one pass and simulated (moments only)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from egdst_b200 import examples
from tests.oracles import oracle_for

nvec = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nsim = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
m = examples.deaton2(); m.compile()
lib = m._capi()
rng = np.random.default_rng(4096)
params = np.column_stack([rng.uniform(0.0, 0.05, nvec), rng.uniform(0.75, 1.75, nvec)])
t = time.perf_counter(); sol = lib.solve_batch(m, params); dt = time.perf_counter() - t
bad = [v for v in range(nvec) if sol.status(v)[0]]
print("solve_batch(%d) first call %.1f ms; vectors with status != 0: %d %s" % (nvec, dt * 1e3, len(bad), [sol.status(v) for v in bad[:5]]))
for _ in range(2):
    lib.resolve(sol, m, params)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5):
    lib.resolve(sol, m, params)
torch.cuda.synchronize()
ms = (time.perf_counter() - t) / 5 * 1e3
units = nvec * m.nt * m.nst * m.nd * m.ngridm
print("batched solve: %.2f ms for %d vectors = %.1f us/vector, %.3e grid-point-periods/s" % (ms, nvec, ms * 1e3 / nvec, units / ms * 1e3))
lib.profile_enable(True); lib.resolve(sol, m, params); torch.cuda.synchronize(); prof = lib.profile_read(); lib.profile_enable(False)
print("  ", {k: round(v[0], 2) for k, v in prof.items() if v[1]})
print("   phases of vector 0 (ms):", {k: round(v, 3) for k, v in sol.phase_ms().items()})
init = np.column_stack([np.ones(nsim), np.full(nsim, 0.25)])
mom = lib.sim_moments(m, sol, init, 7)
t = time.perf_counter()
for _ in range(3):
    mom = lib.sim_moments(m, sol, init, 7)
ms2 = (time.perf_counter() - t) / 3 * 1e3
lib.profile_enable(True); lib.sim_moments(m, sol, init, 7); prof = lib.profile_read(); lib.profile_enable(False)
print("   sim kernel device time:", {k: round(v[0], 2) for k, v in prof.items() if v[1]})
print("batched sim moments (host API): %.2f ms for %d x %d agents x %d periods = %.3e agent-periods/s" % (ms2, nvec, nsim, m.nt, nvec * nsim * m.nt / ms2 * 1e3))
# CPU reference for a few vectors
secs = []
for i in range(8):
    mi = examples.deaton2(interest=params[i, 0], income=params[i, 1])
    o = oracle_for(mi); Mr, Dr = o.solve(); s1 = o.seconds
    rs = np.random.default_rng(i).random(4 * nsim * m.nt)
    o.simulate(Mr, Dr, init, rs, 0); secs.append((s1, o.seconds))
print("reference CPU per vector: solve %.2f ms, sim(%d agents) %.2f ms" % (np.mean([a for a, b in secs]) * 1e3, nsim, np.mean([b for a, b in secs]) * 1e3))
