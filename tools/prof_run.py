"""Small fixed workload for ncu captures: S1 solve (+1 resolve) and two simulation launches of NSIM agents."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from egdst_b200 import examples, capi

nsim = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
mode = sys.argv[2] if len(sys.argv) > 2 else "both"
m = examples.retirement2_scaled(); m.compile()
lib = m._capi()
sol = lib.solve(m, strict=True)
lib.resolve(sol, m)
dev = torch.device("cuda", 0)
nt, nso = m.nt, m.nsimout()
d_init = torch.empty(2 * nsim, dtype=torch.float64, device=dev)
d_init[:nsim] = 1.0
d_init[nsim:] = m.a0 + 0.5 * (m.mmax - m.a0) * torch.rand(nsim, dtype=torch.float64, device=dev)
d_sims = torch.empty(nso * nt * nsim, dtype=torch.float64, device=dev)
d_mom = torch.zeros(3 * nso * nt, dtype=torch.float64, device=dev)
for _ in range(2):
    lib.simulate_device(m, sol, d_init.data_ptr(), nsim, 0, 12345, d_sims.data_ptr() if mode != "moments" else 0, d_mom.data_ptr() if mode != "sims" else 0)
torch.cuda.synchronize()
print("ok", sol.status(), float(d_mom.view(nt, nso, 3)[:, 0, 2].sum().item()))
