"""oracle/port -- plain-C restatement of the reference's shared numerics, terminal period and simulator
(oracle/port/egdst_port.c), compiled per model image with gcc.

TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
The backward-induction solver is not restated: its checker is the unmodified reference (oracle/ref.py).
Pinned against the golden vectors by tests/test_cpu_port.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

from egdst_b200 import codegen
from egdst_b200.capi import Desc, EgdstDesc

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_build")
SRC = os.path.join(HERE, "port", "egdst_port.c")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(model, force: bool = False) -> str:
    model.prepare()
    key = codegen.model_key(model)
    outdir = os.path.join(OUT, key)
    path = os.path.join(outdir, "libegdst_port.so")
    if os.path.isfile(path) and not force and os.path.getmtime(path) >= os.path.getmtime(SRC):
        return path
    os.makedirs(outdir, exist_ok=True)
    with open(os.path.join(outdir, "modelspec_dev.h"), "w") as f:
        f.write(codegen.emit_devspec(model))
    cmd = ["gcc", "-std=gnu99", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I" + outdir, "-I" + os.path.join(ROOT, "include"),
           SRC, "-o", path, "-lm"]
    subprocess.run(cmd, check=True)
    return path


class Port:
    kind = "port"

    def __init__(self, model):
        self.model = model
        self.lib = L = C.CDLL(build(model))
        L.port_cdfni.restype = C.c_double
        L.port_cdfni.argtypes = [C.c_double]
        L.port_bracket.argtypes = [C.c_double, _dp, C.c_int, C.c_int]
        L.port_linter.restype = C.c_double
        L.port_linter.argtypes = [C.c_double, C.c_int, _dp, _dp]
        L.port_terminal.argtypes = [C.POINTER(EgdstDesc), C.c_int, C.c_int, _dp, _dp, _dp]
        L.port_simulate.argtypes = [C.POINTER(EgdstDesc), _ip, _ip, _dp, _dp, _dp, C.c_int, _dp, C.c_int, _dp]
        self.last_seconds = None

    def cdfni(self, p: float) -> float:
        return float(self.lib.port_cdfni(p))

    def bracket(self, x, grid, type_=0) -> int:
        g = np.ascontiguousarray(grid, dtype=np.float64)
        return int(self.lib.port_bracket(float(x), g.ctypes.data_as(_dp), g.size, type_))

    def linter(self, x, grid, fun) -> float:
        g = np.ascontiguousarray(grid, dtype=np.float64)
        f = np.ascontiguousarray(fun, dtype=np.float64)
        return float(self.lib.port_linter(float(x), g.size, g.ctypes.data_as(_dp), f.ctypes.data_as(_dp)))

    def terminal(self, ist: int, id_: int):
        d = Desc(self.model)
        n = self.model.ngridm
        M, Cc, V = np.empty(n), np.empty(n), np.empty(n)
        rc = self.lib.port_terminal(C.byref(d.c), ist, id_, M.ctypes.data_as(_dp), Cc.ctypes.data_as(_dp), V.ctypes.data_as(_dp))
        return None if rc else (M, Cc, V)

    def solve(self):
        raise NotImplementedError("the solver is not restated: its checker is the compiled reference (oracle/ref.py)")

    def simulate(self, M, D, init, randstream, rndtype: int = 0) -> np.ndarray:
        """[nsim, nt, nsimout], like egdstmodel.m:1270 permutes the gateway's output."""
        m = self.model
        d = Desc(m)
        nst, nt = m.nst, m.nt
        mlen = np.zeros(nt * nst, dtype=np.int32)
        thlen = np.zeros(nt * nst, dtype=np.int32)
        mb, db = [], []
        for it in range(nt):
            for ist in range(nst):
                c = it * nst + ist
                if M[ist][it] is not None and M[ist][it].size:
                    mlen[c], thlen[c] = M[ist][it].shape[0], D[ist][it].shape[0]
                    mb.append(np.asarray(M[ist][it], dtype=np.float64).ravel(order="F"))
                    db.append(np.asarray(D[ist][it], dtype=np.float64).ravel(order="F"))
        Mbuf = np.ascontiguousarray(np.concatenate(mb)) if mb else np.zeros(1)
        Dbuf = np.ascontiguousarray(np.concatenate(db)) if db else np.zeros(1)
        init = np.atleast_2d(np.asarray(init, dtype=np.float64))
        nsim = init.shape[0]
        initf = np.ascontiguousarray(init.ravel(order="F"))
        rs = np.ascontiguousarray(np.asarray(randstream, dtype=np.float64).ravel())
        nso = m.nsimout()
        sims = np.empty(nso * nt * nsim, dtype=np.float64)
        t = time.perf_counter()
        self.lib.port_simulate(C.byref(d.c), mlen.ctypes.data_as(_ip), thlen.ctypes.data_as(_ip), Mbuf.ctypes.data_as(_dp),
                               Dbuf.ctypes.data_as(_dp), initf.ctypes.data_as(_dp), nsim, rs.ctypes.data_as(_dp), rndtype,
                               sims.ctypes.data_as(_dp))
        self.last_seconds = time.perf_counter() - t
        return np.transpose(sims.reshape((nso, nt, nsim), order="F"), (2, 1, 0))
