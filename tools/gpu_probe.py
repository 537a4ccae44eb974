"""First-contact GPU probe: fixture parity summary, S1 solve timing + parity, simulation timing."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from egdst_b200 import examples
from tests import goldens
from tests.parity import solution_errors
from tests.oracles import oracle_for

def fmt(e):
    return {k: (("%.2e" % v) if isinstance(v, float) else v) for k, v in e.items() if k != "where"}

out = {}
for name in goldens.CASES:
    try:
        m = goldens.model_for(name); m.compile()
        t = time.perf_counter(); m.solve(); dt = time.perf_counter() - t
        g = goldens.load(name)
        e = solution_errors(m.M, m.D, g["M"], g["D"])
        m.sim(g["init"], "own_shocks", randstream=g["randstream"])
        se = goldens.sims_errors(m.sims, g["sims"])
        print(name, "solve %.1f ms" % (dt * 1e3), "status", m._solution.status(), fmt(e), se, flush=True)
        out[name] = {"solve": fmt(e), "sim": se}
    except Exception as ex:
        print(name, "FAILED", repr(ex), flush=True)

which = sys.argv[1:] or ["s1"]
if "s1" in which:
    m = examples.retirement2_scaled(); m.compile()
    lib = m._capi()
    t = time.perf_counter(); m.solve(); dt = time.perf_counter() - t
    sol = m._solution
    print("S1 first solve %.1f ms status %s units %d" % (dt * 1e3, sol.status(), sol.units()), flush=True)
    import ctypes
    cudart = ctypes.CDLL("libcudart.so.12") if False else None
    for rep in range(3):
        t = time.perf_counter(); lib.resolve(sol, m); sol.sizes(); dt = time.perf_counter() - t
        print("S1 resolve+sizes %.1f ms" % (dt * 1e3), flush=True)
    orc = oracle_for(m)
    t = time.perf_counter(); Mr, Dr = orc.solve(); dt = time.perf_counter() - t
    print("oracle (%s) S1 solve %.2f s (gateway %.2f s)" % (orc.kind, dt, orc.seconds), flush=True)
    e = solution_errors(sol.M, sol.D, Mr, Dr)
    print("S1 parity", fmt(e), e["where"], flush=True)
    rows = sum(Mr[0][it].shape[0] for it in range(m.nt))
    print("oracle rows", rows, "ours", sum(sol.M[0][it].shape[0] for it in range(m.nt)))
    # simulation timing (Philox), 1e6 agents
    nsim = 1_000_000
    rng = np.random.default_rng(20141)
    init = np.column_stack([np.ones(nsim), m.a0 + 0.5 * (m.mmax - m.a0) * rng.random(nsim)])
    for rep in range(2):
        t = time.perf_counter(); s, mom = lib.simulate_philox(m, sol, init, 12345, want_sims=False, want_moments=True); dt = time.perf_counter() - t
        print("sim 1e6 agents moments-only (host API) %.1f ms -> %.3e agent-periods/s" % (dt * 1e3, nsim * m.nt / dt), flush=True)
    # parity sub-run with a host randstream
    nsim = 10000
    init = init[:nsim]
    rs = rng.random(4 * nsim * m.nt)
    sims = lib.simulate(m, sol, init, rs, 0)
    sr = orc.simulate(Mr, Dr, init, rs, 0)
    print("S2 parity (1e4 agents, reference tables vs ours):", goldens.sims_errors(sims, sr), "oracle sim %.2f s" % orc.seconds, flush=True)
