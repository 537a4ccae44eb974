"""S1b probe: status and parity of the shipped-parameter retirement model (next to the reference's instability) and of
its stable sibling (interest=0.02) under the current EGM launch configuration."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from egdst_b200 import examples
from oracle import ref
from tests.parity import solution_errors, cell_errors
for label, kw in (("shipped", {}), ("stable", dict(interest=0.02))):
    m = examples.retirement(T=40, ngridm=2000, ngridmax=4000, nthrhmax=2000, ny=20, **kw)
    m.compile(); m.solve()
    st = m._solution.status()
    Mb, Db = ref.Reference(m).solve()
    errs = []
    for it in range(m.nt):
        if m.M[0][it] is None or Mb[0][it] is None or m.M[0][it].size == 0: errs.append(None); continue
        e = cell_errors(m.M[0][it], m.D[0][it], Mb[0][it], Db[0][it]); errs.append(max(e["C"], e["V"]))
    print(label, "parts", os.environ.get("EGDST_EGM_PARTS", "auto"), "status", st, "err by period (last 8 .. first 8):",
          ["%.1e" % x if x is not None else None for x in errs[-8:]], ["%.1e" % x if x is not None else None for x in errs[:8]], flush=True)
