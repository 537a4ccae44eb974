"""Small fixed workload for ncu captures of the solve kernel: S1 solve + 1 resolve."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from egdst_b200 import examples
m = examples.retirement2_scaled(); m.compile()
lib = m._capi()
sol = lib.solve(m, strict=True)
lib.resolve(sol, m)
print("ok", sol.status(), sol.units())
