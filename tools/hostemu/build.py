"""Build the kernels for the host emulator (development aid, see cuda_emul.h).

    python tools/hostemu/build.py retirement2      -> tools/hostemu/_build/<key>/libemu.so

The result exports the same C ABI as the nvcc-built library, so `egdst_b200.capi.ModelLibrary(path)`
can drive it from a debugging session.  It is never used by the package, the tests' checker or bench.py.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from egdst_b200 import codegen, examples  # noqa: E402


def build(model, opt="-O1"):
    model.prepare()
    key = codegen.model_key(model)
    out = os.path.join(HERE, "_build", key)
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "modelspec_dev.h"), "w") as f:
        f.write(codegen.emit_devspec(model))
    asan = os.environ.get("EGDST_EMU_ASAN") == "1"  # AddressSanitizer build: out-of-bounds accesses of the kernels' logic
    lib = os.path.join(out, "libemu_asan.so" if asan else "libemu.so")
    csrc = os.path.join(ROOT, "egdst_b200", "csrc")
    cmd = ["g++", "-std=c++20", opt, "-g", "-DEGDST_HOSTEMU", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-w",
           "-I" + out, "-I" + HERE, "-I" + os.path.join(ROOT, "include"), "-I" + csrc,
           '-DEGDST_MODEL_KEY="%s"' % key, "-x", "c++", os.path.join(csrc, "egdst_capi.cu"), "-o", lib]
    if asan:
        cmd[3:3] = ["-fsanitize=address", "-fno-omit-frame-pointer"]
    cmd[3:3] = os.environ.get("EGDST_EMU_DEFS", "").split()  # extra -D switches (debug prints)
    subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "retirement2"
    print(build(examples.ALL[name]()))
