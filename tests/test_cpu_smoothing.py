"""Taste-shock smoothing (sigma_eps > 0, an extension without a reference oracle: parity UNPINNED) on the host
emulator of the kernels: closed form, equation residuals, and the sigma_eps -> 0 limit against the reference."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "hostemu"))

from egdst_b200 import capi, examples  # noqa: E402
from tests import smoothing_checks as sc  # noqa: E402
from tests.oracles import oracle_for, ref_available  # noqa: E402
from tests.parity import solution_errors  # noqa: E402


def _emulated(model):
    from build import build  # tools/hostemu/build.py
    model.prepare()
    return capi.ModelLibrary(build(model))


def test_two_period_closed_form_with_taste_shocks():
    m = sc.two_period_model(sigma_eps=0.5, ngridm=2000)
    lib = _emulated(m)
    sol = lib.solve(m)
    assert sol.status(0)[0] == 0, sol.status(0)
    w = sc.two_period_closed_form(sol, m)
    assert w["C"] < 1e-10 and w["V"] < 2e-6, w
    # the smoothing term is far above the tolerance: a hard max would be off by sigma*log(1+exp(-duw/sigma)) = 0.157
    # the mode is a property of the compiled image: the same image refuses a model without taste shocks
    m0 = sc.two_period_model(sigma_eps=0.0, ngridm=2000)
    m0.prepare()
    with pytest.raises(capi.EgdstError):
        lib.solve(m0)


def test_euler_and_bellman_equations_hold_with_taste_shocks():
    m = examples.retirement2(T=5, ngridm=300, ngridmax=900, ny=5)
    m.sigma_eps = 0.2
    lib = _emulated(m)
    sol = lib.solve(m)
    assert sol.status(0)[0] == 0, sol.status(0)
    for it in (3, 2, 0):
        w = sc.euler_bellman_residuals(sol, m, it)
        assert w["euler"] < 1e-8 and w["bellman"] < 1e-7 and w["points"] > 100, (it, w)  # the kernels use Acklam's quantile (1e-9), numpy the exact one


def test_vanishing_taste_shocks_reproduce_the_reference():
    m = examples.retirement2(ngridm=100, ngridmax=600, ny=4)  # 25 periods: the retirement threshold reaches the top of the grid
    if not ref_available(m):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    Mr, Dr = oracle_for(m).solve()
    m.sigma_eps = 1e-8
    lib = _emulated(m)
    sol = lib.solve(m)
    assert sol.status(0)[0] == 0, sol.status(0)
    e = solution_errors(sol.M, sol.D, Mr, Dr)
    assert e["C"] < 1e-6 and e["V"] < 1e-6 and e["TH"] < 1e-6 and e["Dseq"], e
