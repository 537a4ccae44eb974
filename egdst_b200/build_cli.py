"""Command-line entry of the model compiler: ``python -m egdst_b200.build_cli model.json --out DIR``.

Reads the model definition as a JSON dump of the ``@egdstmodel`` object's public properties (what MATLAB's
``jsonencode(struct(model))`` writes, or ``EgdstModel.to_dict()``), generates ``modelspec_dev.h`` and compiles the
per-model CUDA library for sm_100a into DIR.  This is the step that replaces the reference's three ``mex`` calls
(@egdstmodel/compile.m:754-819); INTEGRATION.md shows the edited compile.m.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import sys

from . import build, codegen
from .model import EgdstModel


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("model_json")
    ap.add_argument("--out", default=None, help="directory that receives libegdst_b200_<key>.so (default: in-tree _lib/<key>)")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args(argv)
    with open(a.model_json) as f:
        m = EgdstModel.from_dict(json.load(f))
    m.prepare()
    path = build.build_model_library(m, force=a.force)
    if a.out:
        os.makedirs(a.out, exist_ok=True)
        dst = os.path.join(a.out, os.path.basename(path))
        shutil.copy2(path, dst)
        path = dst
    print(json.dumps({"key": codegen.model_key(m), "library": path, "optim": m.optim}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
