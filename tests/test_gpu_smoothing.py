"""Taste-shock smoothing (model.sigma_eps > 0) on the GPU.  An EXTENSION without a reference oracle -- parity is
UNPINNED for sigma_eps > 0; pinned are a closed form, the model's own equations re-evaluated in numpy from the exported
choice-specific cells, and the sigma_eps -> 0 limit against the reference (tests/smoothing_checks.py)."""
import numpy as np
import pytest

from egdst_b200 import capi, examples
from tests import smoothing_checks as sc
from tests.oracles import oracle_for
from tests.parity import solution_errors

pytestmark = pytest.mark.gpu


def test_two_period_closed_form_with_taste_shocks():
    for sig in (0.5, 0.05):
        m = sc.two_period_model(sigma_eps=sig, ngridm=2000)
        m.compile()
        sol = m._capi().solve(m)
        assert sol.status(0)[0] == 0, sol.status(0)
        w = sc.two_period_closed_form(sol, m)
        assert w["C"] < 1e-10 and w["V"] < 2e-6, (sig, w)


def test_choice_cells_need_the_smoothing_mode():
    m = examples.retirement2()
    m.compile()
    sol = m._capi().solve(m)
    with pytest.raises(capi.EgdstError):
        sol.choice_cell(0, 0, 0)


@pytest.mark.parametrize("sig", [0.2, 0.02])
def test_euler_and_bellman_equations_hold_with_taste_shocks(sig):
    m = examples.retirement2(ngridm=500, ngridmax=1500)
    m.sigma_eps = sig
    m.compile()
    sol = m._capi().solve(m)
    assert sol.status(0)[0] == 0, sol.status(0)
    for it in (23, 20, 12, 5, 0):
        w = sc.euler_bellman_residuals(sol, m, it)
        assert w["euler"] < 1e-8 and w["bellman"] < 1e-7 and w["points"] > 100, (it, w)  # Acklam's quantile vs the exact one


@pytest.mark.parametrize("kw", [dict(), dict(ngridm=500, ngridmax=2000)])
def test_vanishing_taste_shocks_reproduce_the_reference(kw):
    m = examples.retirement2(**kw)
    Mr, Dr = oracle_for(m).solve()
    m.sigma_eps = 1e-8
    m.compile()
    sol = m._capi().solve(m)
    assert sol.status(0)[0] == 0, sol.status(0)
    e = solution_errors(sol.M, sol.D, Mr, Dr)
    assert e["C"] < 1e-6 and e["V"] < 1e-6 and e["TH"] < 1e-6 and e["Dseq"], e


def test_smoothing_makes_the_choice_specific_values_smooth():
    """With a sizeable sigma_eps the secondary envelope has nothing left to remove (Iskhakov et al. 2017, Theorem 3: the
    kinks of the value functions are smoothed out): no decision cell of retirement2 loses grid points."""
    m = examples.retirement2()
    m.sigma_eps = 0.5
    m.compile()
    sol = m._capi().solve(m)
    assert sol.status(0)[0] == 0
    for it in range(m.nt - 1):
        for d in (0, 1):
            c = sol.choice_cell(it, 0, d)
            assert c.shape[0] >= 0.9 * m.ngridm and np.all(np.diff(c[1:, 0]) > 0), (it, d, c.shape)  # (the longer list is cut at the unified grid's bound)


def test_batched_solves_in_the_smoothing_mode(monkeypatch):
    m = examples.retirement2(ngridm=200, ngridmax=600)
    m.sigma_eps = 0.1
    m.compile()
    lib = m._capi()
    one = lib.solve(m)
    pv = np.array([list(m.param_vector())] * 12)
    for scope in ("cta", "warp"):
        monkeypatch.setenv("EGDST_SOLVE_SCOPE", scope)
        monkeypatch.setenv("EGDST_WARP_G", "5")
        sol = lib.solve_batch(m, pv)
        for v in (0, 7, 11):
            assert sol.status(v)[0] == 0
            Mb, Db = sol.cells(v)
            e = solution_errors(Mb, Db, one.M, one.D)
            assert e["C"] < 1e-12 and e["V"] < 1e-12 and e["rowdiff"] == 0, (scope, v, e)
            a, b = sol.choice_cell(3, 0, 1, ivec=v), one.choice_cell(3, 0, 1)
            assert a.shape == b.shape and np.max(np.abs(a[1:] - b[1:])) < 1e-12
