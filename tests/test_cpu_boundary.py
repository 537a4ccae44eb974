"""CPU-side checks of the drop-in boundary: every model image builds for sm_100a, loads, and exports each
symbol include/egdst_b200.h declares; compute entry points fail loudly without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from egdst_b200 import build, capi, examples

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "egdst_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(egdst_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    m = examples.retirement2()
    m.compile()
    return m, m._capi()


def test_header_symbols_are_exported(lib):
    m, L = lib
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L.L, s), "library does not export %s" % s
    assert set(capi.ModelLibrary.EXPORTS) <= set(syms)


def test_model_image_identity(lib):
    m, L = lib
    assert L.L.egdst_abi_version() == capi.ABI_VERSION
    assert L.L.egdst_model_nparam() == len(m.param)
    assert L.L.egdst_model_neq() == len(m.eq)
    assert L.L.egdst_model_key().decode() in L.path


def test_no_cpu_path(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m, L = lib
    with pytest.raises(capi.EgdstError) as e:
        L.solve(m)
    assert e.value.code == 2 and "no CUDA device" in str(e.value)


def test_missing_library_raises(tmp_path):
    with pytest.raises(FileNotFoundError):
        capi.ModelLibrary(str(tmp_path / "libegdst_b200_none.so"))


def test_descriptor_layout(lib):
    m, L = lib
    d = capi.Desc(m)
    assert d.c.ngridm == 100 and d.c.nd == 2 and d.c.nst == 1 and d.c.ny == 10 and d.c.T == 25
    assert d.c.optim_MUnoD == 1 and d.c.nparam == 3
    assert np.isclose(np.ctypeslib.as_array(d.c.quadrature, shape=(20,))[:10].sum(), 1.0)


def test_all_fixture_images_share_sources():
    # models differing only in run-time properties share one image (keyed on generated source)
    a = examples.retirement2(); a.prepare()
    b = examples.retirement2_scaled(); b.prepare()
    assert build.library_path(a) == build.library_path(b)
