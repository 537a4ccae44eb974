"""Odd-shape probe: solve + simulate unusual grid sizes / horizons on the GPU and compare with the reference oracle."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from egdst_b200 import examples
from tests.oracles import oracle_for
from tests.parity import solution_errors
from tests.goldens import sims_errors

cases = [
    ("retirement2", dict(ngridm=7, ngridmax=100, T=3, ny=3)),
    ("retirement2", dict(ngridm=33, ngridmax=100, T=2, ny=1)),
    ("retirement2", dict(ngridm=101, ngridmax=300, T=9, ny=7)),
    ("retirement2", dict(ngridm=2049, ngridmax=5000, T=6, ny=5, nthrhmax=2049)),
    ("retirement2", dict(ngridm=50000, ngridmax=100000, T=5, ny=20, nthrhmax=50000, interest=0.02)),
    ("deaton2", dict(ngridm=5, ngridmax=100, T=4)),
    ("model2", dict(T=2, ngridm=11, nquad=2, sigma=0.2)),
    ("occ3", dict(ngridm=17, ngridmax=100, ny=3, T=6)),
]
bad = 0
for name, kw in cases:
    try:
        m = examples.ALL[name](**kw); m.compile()
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always"); m.solve()
        orc = oracle_for(m)
        try:
            Mr, Dr = orc.solve()
        except Exception as e:
            print(name, kw, "reference failed:", str(e)[:80], "| ours status", m._solution.status(), [str(x.message)[:60] for x in w]); continue
        e = solution_errors(m.M, m.D, Mr, Dr)
        rng = np.random.default_rng(1); nsim = 200
        init = np.column_stack([np.full(nsim, float(m.nst)), m.a0 + (m.mmax - m.a0) * rng.random(nsim) * 0.9])
        rs = rng.random(4 * nsim * m.nt)
        m.sim(init, "own_shocks", randstream=rs)
        se = sims_errors(m.sims, orc.simulate(Mr, Dr, init, rs, 0))
        ok = e["C"] < 1e-9 and e["V"] < 1e-9 and e["TH"] < 1e-8 and e["Dseq"] and se["nan_mismatch"] == 0 and se["discrete_mismatch"] == 0 and se["max"] < 1e-9
        bad += not ok
        print("OK " if ok else "BAD", name, kw, "status", m._solution.status(), {k: (("%.1e" % v) if isinstance(v, float) else v) for k, v in e.items() if k != "where"}, se)
    except Exception as ex:
        bad += 1
        print("EXC", name, kw, repr(ex)[:200])
print("bad =", bad)
