"""Debug driver: solve a model with the host-emulated kernels and compare with oracle/_ref."""
import os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
import numpy as np
from build import build
from egdst_b200 import examples, capi
from oracle.ref import Reference
from tests.parity import solution_errors

def run(name, **kw):
    m = examples.ALL[name](**kw); m.prepare()
    lib = capi.ModelLibrary(build(m))
    t = time.time(); sol = lib.solve(m); te = time.time() - t
    r = Reference(m); Mr, Dr = r.solve()
    M, D = sol.M, sol.D
    e = solution_errors(M, D, Mr, Dr)
    print(name, kw, "emu %.1fs" % te, "status", sol.status(), {k: (("%.2e" % v) if isinstance(v, float) else v) for k, v in e.items()})
    return m, lib, sol, (M, D), (Mr, Dr)

if __name__ == "__main__":
    names = sys.argv[1:] or ["cake1"]
    for n in names:
        run(n)
