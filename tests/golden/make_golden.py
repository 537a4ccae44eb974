"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference C (oracle/_ref).

Run in the development container (where /root/reference exists):
    python tests/golden/make_golden.py
The reference repo holds no test vectors of its own, so these files -- outputs of the reference
itself on its own example models -- are what pins the oracle restatement and the CUDA path on machines
without /root/reference.  Each .npz holds, for one shipped example model: every M{ist,it} and D{ist,it}
cell, and a 16-agent simulation on a fixed random stream.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from egdst_b200 import examples  # noqa: E402
from oracle.ref import Reference  # noqa: E402

CASES = {
    "cake1": {}, "cake2": {}, "deaton1": {}, "deaton2": {}, "retirement1": {}, "retirement2": {}, "occ3": {},
    "model2": dict(T=8, sigma=0.25, duw=float(np.log(5.0)), ngridm=60, nquad=8),
    "humancapital": {},
}
MODELS = dict(examples.ALL, **examples.EXTRA)


def sim_inputs(m, nsim=16, seed=2014):
    rng = np.random.default_rng(seed)
    ist0 = np.full(nsim, float(m.nst))  # last state (model2: 'working')
    if any(v["continuous"] for v in m.s):
        # the reference's simulator addresses the solution correctly only from initial cells at the FIRST grid point of
        # every continuous state (egdst_simulator.c:313 adds the corner offset to the initial cell index)
        ist0 = np.full(nsim, 1.0)
    init = np.column_stack([ist0, m.a0 + 0.25 * (m.mmax - m.a0) * (0.05 + rng.random(nsim))])
    rs = rng.random(4 * nsim * m.nt)
    return init, rs


def main():
    for name, kw in CASES.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        m = MODELS[name](**kw)
        r = Reference(m)
        M, D = r.solve()
        init, rs = sim_inputs(m)
        sims = r.simulate(M, D, init, rs, 0)
        out = {"init": init, "randstream": rs, "sims": sims, "nst": m.nst, "nt": m.nt}
        if any(v["continuous"] for v in m.s):
            sims[:, :, 3] = np.nan  # never assigned by the reference on the continuous branch (uninitialised stack)
            out["skipcols"] = np.array([3])
        for ist in range(m.nst):
            for it in range(m.nt):
                if M[ist][it] is not None and M[ist][it].size:
                    out["M_%d_%d" % (ist, it)] = M[ist][it]
                    out["D_%d_%d" % (ist, it)] = D[ist][it]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "cells", sum(1 for k in out if k.startswith("M_")), "sims", sims.shape)


if __name__ == "__main__":
    main()
