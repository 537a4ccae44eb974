"""Drive the product's MEX gateways (mex/egdst_solver.c, egdst_simulator.c, egdst_call.c) without MATLAB:
they are compiled against the functional MEX shim of oracle/shim (test infrastructure), linked with the
per-model CUDA library, and called with the same fake model object the oracle uses for the reference's
gateways.  This is the drop-in boundary test: same names, arities, layouts and error behaviour."""
import os
import subprocess

from egdst_b200 import build, codegen
from oracle.ref import Reference

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_mexbuild")


def build_mex(model) -> str:
    model.prepare()
    lib = build.build_model_library(model)
    key = codegen.model_key(model)
    outdir = os.path.join(OUT, key)
    path = os.path.join(outdir, "libegdst_mex.so")
    srcs = [os.path.join(ROOT, "mex", f) for f in ("egdst_solver.c", "egdst_simulator.c", "egdst_call.c", "egdst_mex_common.h")]
    srcs += [os.path.join(ROOT, "oracle", "shim", "mexshim.c"), lib]
    if os.path.isfile(path) and all(os.path.getmtime(path) >= os.path.getmtime(s) for s in srcs):
        return path
    os.makedirs(outdir, exist_ok=True)
    flags = ["-std=gnu99", "-O2", "-fPIC", "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(ROOT, "include")]
    for k in ("TOLERANCE", "ZEROCONSUMPTION", "DOUBLEPOINT_DELTA"):
        flags.append("-D%s=%s" % (k, model.cflags[k]))  # as compile.m:763-774 passes cflags to mex
    objs = []
    for src, gate in (("egdst_solver.c", "ref_solver_gateway"), ("egdst_simulator.c", "ref_simulator_gateway"), ("egdst_call.c", "ref_call_gateway")):
        o = os.path.join(outdir, src[:-2] + ".o")
        subprocess.run(["gcc"] + flags + ["-DmexFunction=" + gate, "-c", os.path.join(ROOT, "mex", src), "-o", o], check=True)
        objs.append(o)
    o = os.path.join(outdir, "mexshim.o")
    subprocess.run(["gcc"] + flags + ["-c", os.path.join(ROOT, "oracle", "shim", "mexshim.c"), "-o", o], check=True)
    objs.append(o)
    libdir = os.path.dirname(lib)
    subprocess.run(["gcc", "-shared", "-Wl,-Bsymbolic", "-o", path] + objs +
                   ["-L" + libdir, "-l:" + os.path.basename(lib), "-Wl,-rpath," + libdir, "-lm"], check=True)
    return path


class MexDropIn(Reference):
    """Same driver as the oracle's, pointed at the product's gateways."""

    def __init__(self, model):
        super().__init__(model, libpath=build_mex(model))
