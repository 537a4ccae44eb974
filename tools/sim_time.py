"""Time the simulation kernel on the S1 solution: sims+moments, sims only, moments only (CUDA events)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from egdst_b200 import examples, capi

nsim = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
m = examples.retirement2_scaled(); m.compile()
lib = m._capi()
sol = lib.solve(m, strict=True)
dev = torch.device("cuda", 0)
nt, nso = m.nt, m.nsimout()
d_init = torch.empty(2 * nsim, dtype=torch.float64, device=dev)
d_init[:nsim] = 1.0
d_init[nsim:] = m.a0 + 0.5 * (m.mmax - m.a0) * torch.rand(nsim, dtype=torch.float64, device=dev)
d_sims = torch.empty(nso * nt * nsim, dtype=torch.float64, device=dev)
d_mom = torch.zeros(3 * nso * nt, dtype=torch.float64, device=dev)
desc = capi.Desc(m)
for label, ps, pm in (("sims+moments", d_sims.data_ptr(), d_mom.data_ptr()), ("sims only", d_sims.data_ptr(), 0), ("moments only", 0, d_mom.data_ptr())):
    for _ in range(2):
        lib.simulate_device(m, sol, d_init.data_ptr(), nsim, 0, 12345, ps, pm, desc=desc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lib.simulate_device(m, sol, d_init.data_ptr(), nsim, 0, 12345, ps, pm, desc=desc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gb = 8.0 * nso * nt * nsim / 1e9
    print("%-14s %8.2f ms  %.3e agent-periods/s  %7.1f GB/s (of sims bytes)" % (label, ms, nsim * nt / ms * 1e3, gb / ms * 1e3), flush=True)
