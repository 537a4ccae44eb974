"""oracle/_ref -- the UNMODIFIED reference C, compiled where it lies, as the parity oracle.

TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product (egdst_b200) never does.

Recipe:
    gcc -std=gnu99 -O2 -fno-inline -ffp-contract=off -include stdbool.h
        -DDISTRIB={1|2} -DTOLERANCE=.. -DZEROCONSUMPTION=.. -DDOUBLEPOINT_DELTA=.. -DVERBOSE=0
        -DmexFunction=ref_<gateway>_gateway   (one per gateway translation unit)
        /root/reference/@egdstmodel/{egdst_lib,egdst_solver,egdst_simulator,egdst_call}.c
        <generated>/modelspec.c  oracle/shim/mexshim.c   -> oracle/_ref/<key>/libegdst_ref.so

* ``-O2 -fno-inline`` (never plain -O2): the reference reads ``evf`` uninitialised during the adraw
  seed phase (egdst_solver.c:495,572-583) and gcc >= -O1 with inlining miscompiles it.
* No reference source is copied into the repo: the sources are compiled from /root/reference and only
  the shared object (plus the generated modelspec.c/.h) lands in oracle/_ref/, which is git-ignored but
  travels to the GPU box.  On a box without /root/reference the prebuilt .so is used as is.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from egdst_b200 import codegen
from egdst_b200.quadrature import model_quadrature

HERE = os.path.dirname(os.path.abspath(__file__))
REFROOT = os.environ.get("EGDST_REFERENCE_ROOT", "/root/reference")
REFSRC = os.path.join(REFROOT, "@egdstmodel")
OUTROOT = os.path.join(HERE, "_ref")

BASE_FLAGS = ["-std=gnu99", "-O2", "-fno-inline", "-ffp-contract=off", "-include", "stdbool.h", "-fPIC", "-w"]
NOISE_FLAGS = ["-std=gnu99", "-O2", "-fno-inline", "-march=x86-64-v3", "-ffp-contract=fast", "-include", "stdbool.h", "-fPIC", "-w"]


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFSRC, "egdst_solver.c"))


def lib_path(model, variant: str = "base") -> str:
    model.prepare()
    return os.path.join(OUTROOT, codegen.model_key(model) + ("" if variant == "base" else "_" + variant), "libegdst_ref.so")


def build(model, variant: str = "base", force: bool = False) -> Optional[str]:
    """Build (or find) the reference library for this model structure.  Returns None when neither the
    prebuilt library nor /root/reference is present."""
    path = lib_path(model, variant)
    if os.path.isfile(path) and not force:
        return path
    if not reference_available():
        return None
    outdir = os.path.dirname(path)
    os.makedirs(outdir, exist_ok=True)
    c_src, h_src = codegen.emit_refspec(model)
    with open(os.path.join(outdir, "modelspec.c"), "w") as f:
        f.write(c_src)
    with open(os.path.join(outdir, "modelspec.h"), "w") as f:
        f.write(h_src)
    flags = list(BASE_FLAGS if variant == "base" else NOISE_FLAGS)
    flags += ["-DDISTRIB=%d" % (1 if model.shock["type"] == "lognormal" else 2)]
    for k, v in model.cflags.items():
        flags.append("-D%s=%s" % (k, v))
    inc = ["-I" + outdir, "-I" + os.path.join(HERE, "shim"), "-I" + REFSRC]
    objs = []
    units = [("egdst_lib.c", None), ("egdst_solver.c", "ref_solver_gateway"),
             ("egdst_simulator.c", "ref_simulator_gateway"), ("egdst_call.c", "ref_call_gateway")]
    for src, gate in units:
        obj = os.path.join(outdir, src.replace(".c", ".o"))
        cmd = ["gcc"] + flags + inc + (["-DmexFunction=" + gate] if gate else []) + ["-c", os.path.join(REFSRC, src), "-o", obj]
        subprocess.run(cmd, check=True)
        objs.append(obj)
    for src in (os.path.join(outdir, "modelspec.c"), os.path.join(HERE, "shim", "mexshim.c")):
        obj = os.path.join(outdir, os.path.basename(src).replace(".c", ".o"))
        subprocess.run(["gcc"] + flags + inc + ["-c", src, "-o", obj], check=True)
        objs.append(obj)
    # -Bsymbolic: the reference defines globals named `err` and `error`, which glibc also exports
    subprocess.run(["gcc", "-shared", "-Wl,-Bsymbolic", "-o", path] + objs + ["-lm"], check=True)
    for o in objs:
        os.remove(o)
    return path


def build_harness(model, force: bool = False) -> Optional[str]:
    """oracle/_ref/<key>_env/libegdst_refenv.so: the reference's envelope2()/envelop() callable on given points
    (oracle/shim/envharness.c includes the unmodified egdst_solver.c from /root/reference)."""
    model.prepare()
    outdir = os.path.join(OUTROOT, codegen.model_key(model) + "_env")
    path = os.path.join(outdir, "libegdst_refenv.so")
    if os.path.isfile(path) and not force:
        return path
    if not reference_available():
        return None
    os.makedirs(outdir, exist_ok=True)
    c_src, h_src = codegen.emit_refspec(model)
    with open(os.path.join(outdir, "modelspec.c"), "w") as f:
        f.write(c_src)
    with open(os.path.join(outdir, "modelspec.h"), "w") as f:
        f.write(h_src)
    flags = list(BASE_FLAGS) + ["-DDISTRIB=%d" % (1 if model.shock["type"] == "lognormal" else 2)]
    for k, v in model.cflags.items():
        flags.append("-D%s=%s" % (k, v))
    inc = ["-I" + outdir, "-I" + os.path.join(HERE, "shim"), "-I" + REFSRC]
    srcs = [os.path.join(REFSRC, "egdst_lib.c"), os.path.join(outdir, "modelspec.c"),
            os.path.join(HERE, "shim", "mexshim.c"), os.path.join(HERE, "shim", "envharness.c")]
    subprocess.run(["gcc"] + flags + inc + ["-shared", "-Wl,-Bsymbolic", "-o", path] + srcs + ["-lm"], check=True)
    return path


class _MX(C.Structure):
    pass


_MX._fields_ = [("cls", C.c_int), ("m", C.c_size_t), ("n", C.c_size_t), ("pr", C.POINTER(C.c_double)),
                ("cells", C.POINTER(C.POINTER(_MX))), ("nfields", C.c_int), ("fnames", C.POINTER(C.c_char_p)),
                ("fvals", C.POINTER(C.POINTER(_MX))), ("logical_val", C.c_int)]
_PMX = C.POINTER(_MX)


class RefError(RuntimeError):
    pass


class Reference:
    """Runs the reference gateways on an ``EgdstModel`` through the shim."""

    def __init__(self, model, variant: str = "base", libpath: Optional[str] = None):
        # libpath: any library exporting the three gateways + the shim harness (tests/mexharness.py drives the
        # product's own MEX gateways through the same fake model object)
        path = libpath or build(model, variant)
        if path is None:
            raise FileNotFoundError("oracle/_ref library for this model is not built and /root/reference is absent")
        self.lib = L = C.CDLL(path)
        L.shim_double.restype = _PMX
        L.shim_double.argtypes = [C.c_size_t, C.c_size_t, C.c_void_p]
        L.shim_logical.restype = _PMX
        L.shim_logical.argtypes = [C.c_int]
        L.shim_cell.restype = _PMX
        L.shim_cell.argtypes = [C.c_size_t, C.c_size_t]
        L.shim_struct.restype = _PMX
        L.shim_struct.argtypes = [C.c_size_t]
        L.shim_setfield.argtypes = [_PMX, C.c_size_t, C.c_char_p, _PMX]
        L.shim_free.argtypes = [_PMX]
        L.mxSetCell.argtypes = [_PMX, C.c_size_t, _PMX]
        L.mxGetCell.restype = _PMX
        L.mxGetCell.argtypes = [_PMX, C.c_size_t]
        L.shim_call.restype = C.c_double
        L.shim_call.argtypes = [C.c_int, C.c_int, C.POINTER(_PMX), C.c_int, C.POINTER(_PMX)]
        L.shim_errmsg.restype = C.c_char_p
        L.shim_warnings.restype = C.c_char_p
        L.shim_warncount.restype = C.c_int
        self.model = model
        self.last_seconds = None
        self.warnings = ""

    # -- fake mxArray construction
    def _dbl(self, arr, shape=None):
        a = np.asarray(arr, dtype=np.float64)
        if shape is None:
            a2 = np.atleast_2d(a)
            if a.ndim <= 1:
                a2 = a2.reshape(1, -1) if a.size else a2.reshape(0, 0)
            shape = a2.shape
        flat = np.asfortranarray(a.reshape(shape)).ravel(order="F") if a.size else a.ravel()
        return self.lib.shim_double(shape[0], shape[1], flat.ctypes.data if flat.size else None)

    def _model_object(self, quadrature=True, init=None, randstream=None, M=None, D=None):
        m, L = self.model, self.lib
        obj = L.shim_struct(1)
        sf = lambda k, v: L.shim_setfield(obj, 0, k.encode(), v)  # noqa: E731
        for k in ("t0", "T", "ngridm", "ngridmax", "nthrhmax", "ny", "nd", "nnd", "nst", "nnst", "mmax", "a0"):
            sf(k, self._dbl([float(getattr(m, k))]))
        sf("stm", self._dbl(np.asarray(m.stm, dtype=np.float64)))
        sf("states", self._dbl(m.states, m.states.shape))
        sf("decisions", self._dbl(m.decisions, m.decisions.shape))
        opt = L.shim_struct(1)
        for k, v in m.optim.items():
            L.shim_setfield(opt, 0, k.encode(), L.shim_logical(1 if v else 0))
        sf("optim", opt)
        par = L.shim_struct(len(m.param))
        for i, p in enumerate(m.param):
            L.shim_setfield(par, i, b"value", self._dbl([p["value"]]))
        sf("param", par)
        sv = L.shim_struct(len(m.s))
        for i, s in enumerate(m.s):
            L.shim_setfield(sv, i, b"discrete", L.shim_logical(1 if s["discrete"] else 0))
            L.shim_setfield(sv, i, b"grid", self._dbl(np.asarray(s.get("grid", []), dtype=np.float64)))
        sf("s", sv)
        sf("eq", L.shim_struct(len(m.eq)))
        if quadrature:  # (ny == 1 included: the gateway reads the property unconditionally, egdst_solver.c:162)
            q = model_quadrature(max(m.ny, 1))  # fresh copy each call: the gateway overwrites the abscissas (egdst_solver.c:164)
            sf("quadrature", self._dbl(q.reshape(2, max(m.ny, 1)).T, (max(m.ny, 1), 2)))
        if init is not None:
            sf("init", self._dbl(init, init.shape))
        if randstream is not None:
            sf("randstream", self._dbl(randstream.reshape(-1, 1), (randstream.size, 1)))
        if M is not None:
            nst, nt = m.nst, m.nt
            cm, cd = L.shim_cell(nst, nt), L.shim_cell(nst, nt)
            for it in range(nt):
                for ist in range(nst):
                    if M[ist][it] is not None:
                        L.mxSetCell(cm, ist + it * nst, self._dbl(M[ist][it], M[ist][it].shape))
                        L.mxSetCell(cd, ist + it * nst, self._dbl(D[ist][it], D[ist][it].shape))
            sf("M", cm)
            sf("D", cd)
        return obj

    @staticmethod
    def _mat(p) -> Optional[np.ndarray]:
        if not p:
            return None
        a = p.contents
        n = a.m * a.n
        if n == 0:
            return np.zeros((a.m, a.n))
        return np.ctypeslib.as_array(a.pr, shape=(n,)).copy().reshape((a.m, a.n), order="F")

    def _run(self, which, nlhs, args, strict):
        L = self.lib
        L.shim_reset()
        plhs = (_PMX * max(nlhs, 1))()
        prhs = (_PMX * len(args))(*args)
        sec = L.shim_call(which, nlhs, plhs, len(args), prhs)
        self.warnings = L.shim_warnings().decode(errors="replace")
        if sec < 0:
            raise RefError("reference raised: " + L.shim_errmsg().decode(errors="replace"))
        if strict and L.shim_warncount() > 0:
            raise RefError("reference warned (run invalid, BASELINE.md section 4): " + self.warnings)
        self.last_seconds = sec
        return plhs

    # -- gateways
    def solve(self, strict: bool = True):
        """[M, D, dbgout] = egdst_solver(model); returns (M, D) as nst x nt nested lists of arrays."""
        m = self.model
        m.prepare()
        obj = self._model_object()
        plhs = self._run(0, 3, [obj], strict)
        nst, nt = m.nst, m.nt
        M = [[None] * nt for _ in range(nst)]
        D = [[None] * nt for _ in range(nst)]
        for it in range(nt):
            for ist in range(nst):
                M[ist][it] = self._mat(self.lib.mxGetCell(plhs[0], ist + it * nst))
                D[ist][it] = self._mat(self.lib.mxGetCell(plhs[1], ist + it * nst))
        for k in range(3):
            self.lib.shim_free(plhs[k])
        self.lib.shim_free(obj)
        return M, D

    def simulate(self, M, D, init, randstream, rndtype: int = 0, strict: bool = False) -> np.ndarray:
        """sims = egdst_simulator(model, rndtype) permuted to [nsim, nt, nsimout] as egdstmodel.m:1270 does."""
        m = self.model
        init = np.atleast_2d(np.asarray(init, dtype=np.float64))
        rs = np.asarray(randstream, dtype=np.float64).ravel()
        obj = self._model_object(quadrature=False, init=init, randstream=rs, M=M, D=D)
        rt = self._dbl([float(rndtype)])
        plhs = self._run(1, 1, [obj, rt], strict)
        a = plhs[0].contents
        nsimout, nt, nsim = m.nsimout(), m.nt, init.shape[0]
        sims = np.ctypeslib.as_array(a.pr, shape=(nsimout * nt * nsim,)).copy().reshape((nsimout, nt, nsim), order="F")
        self.lib.shim_free(plhs[0])
        self.lib.shim_free(obj)
        self.lib.shim_free(rt)
        return np.transpose(sims, (2, 1, 0))

    def call(self, M, D, sw: int, args) -> np.ndarray:
        args = np.atleast_2d(np.asarray(args, dtype=np.float64))
        obj = self._model_object(quadrature=False, M=M, D=D)
        a = self._dbl(args, args.shape)
        s = self._dbl([float(sw)])
        plhs = self._run(2, 1, [obj, s, a], False)
        res = self._mat(plhs[0]).ravel()
        self.lib.shim_free(plhs[0])
        self.lib.shim_free(obj)
        return res


class EnvelopeHarness(Reference):
    """The reference's static envelope routines on given points (differential tests of the envelope kernels)."""

    def __init__(self, model):
        path = build_harness(model)
        if path is None:
            raise FileNotFoundError("oracle/_ref envelope harness is not built and /root/reference is absent")
        super().__init__(model, libpath=path)
        L = self.lib
        dp = C.POINTER(C.c_double)
        L.ref_envelope2.restype = C.c_int
        L.ref_envelope2.argtypes = [_PMX, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_double, C.c_char_p, C.c_int]
        L.ref_envelop.restype = C.c_int
        L.ref_envelop.argtypes = [_PMX, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_int]

    def envelope2(self, it, ist, idd, X, Cc, V, evfa0):
        """Secondary envelope of one decision's EGM points (in generation order).  Returns (X, C, V) of the kept points."""
        m = self.model
        n = len(X)
        buf = np.zeros(4 * (m.ngridmax + n + 16))
        buf[0:4 * n:4] = X; buf[1:4 * n:4] = Cc; buf[2:4 * n:4] = V; buf[3:4 * n:4] = idd
        obj = self._model_object()
        err = C.create_string_buffer(400)
        k = self.lib.ref_envelope2(obj, it, ist, idd, buf.ctypes.data_as(C.POINTER(C.c_double)), n, float(evfa0), err, 400)
        self.lib.shim_free(obj)
        if k < 0:
            raise RefError("reference envelope2: " + err.value.decode(errors="replace"))
        return buf[0:4 * k:4].copy(), buf[1:4 * k:4].copy(), buf[2:4 * k:4].copy()

    def envelop(self, it, ist, funcs, evfa0):
        """Primary envelope over functions [(X, C, V), ...] (index = position).  Returns (grid, V, C, thresholds, indices)."""
        m = self.model
        quads = []
        for j, (X, Cc, V) in enumerate(funcs):
            for x, c, v in zip(X, Cc, V):
                quads += [x, c, v, float(j)]
        g = np.array(quads, dtype=np.float64)
        dim0 = g.size // 4
        ev = np.array(evfa0, dtype=np.float64)
        og, of, of2 = np.zeros(m.ngridmax + 8), np.zeros(m.ngridmax + 8), np.zeros(m.ngridmax + 8)
        oth, oix = np.zeros(m.nthrhmax + 8), np.zeros(m.nthrhmax + 8)
        n, mm = C.c_int(0), C.c_int(0)
        dp = C.POINTER(C.c_double)
        obj = self._model_object()
        err = C.create_string_buffer(400)
        rc = self.lib.ref_envelop(obj, it, ist, len(funcs), dim0, g.ctypes.data_as(dp), ev.ctypes.data_as(dp), og.ctypes.data_as(dp),
                                  of.ctypes.data_as(dp), of2.ctypes.data_as(dp), oth.ctypes.data_as(dp), oix.ctypes.data_as(dp),
                                  C.byref(n), C.byref(mm), err, 400)
        self.lib.shim_free(obj)
        if rc < 0:
            raise RefError("reference envelop: " + err.value.decode(errors="replace"))
        return og[:n.value].copy(), of[:n.value].copy(), of2[:n.value].copy(), oth[:mm.value].copy(), oix[:mm.value].copy()
