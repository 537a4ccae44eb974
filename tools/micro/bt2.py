import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from egdst_b200 import examples, capi
nvec, nsim = 4096, 1024
m = examples.deaton2(); m.compile(); lib = m._capi()
rng = np.random.default_rng(4096)
params = np.column_stack([rng.uniform(0.0, 0.05, nvec), rng.uniform(0.75, 1.75, nvec)])
sol = lib.solve_batch(m, params)
init = np.column_stack([np.ones(nsim), np.full(nsim, 0.25)])
dev = torch.device("cuda", 0)
d_init = torch.tensor(init.ravel(order="F"), device=dev)
nso, nt = m.nsimout(), m.nt
d_mom = torch.zeros(nvec * 3 * nso * nt, dtype=torch.float64, device=dev)
desc = capi.Desc(m)
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    d_mom.zero_(); lib.sim_moments_device(m, sol, d_init.data_ptr(), nsim, 0, 7, d_mom.data_ptr(), desc=desc)
    torch.cuda.synchronize(); print("device path %.2f ms" % ((time.perf_counter() - t) * 1e3))
for rep in range(3):
    t = time.perf_counter(); mom = lib.sim_moments(m, sol, init, 7); print("host path %.2f ms" % ((time.perf_counter() - t) * 1e3))
import ctypes as C
dp = C.POINTER(C.c_double)
initf = np.ascontiguousarray(init.ravel(order="F")); momh = np.zeros(nvec * 3 * nso * nt)
for rep in range(3):
    t = time.perf_counter(); rc = lib.L.egdst_sim_moments(C.byref(desc.c), sol.handle, 0, nvec, initf.ctypes.data_as(dp), nsim, 0, 7, momh.ctypes.data_as(dp)); print("raw C host path %.2f ms rc=%d" % ((time.perf_counter() - t) * 1e3, rc))
print(torch.cuda.mem_get_info())
