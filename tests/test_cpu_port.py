"""Pins the plain-C restatement (oracle/port) against the golden vectors -- outputs of the reference itself -- and
against closed forms: simulator on the reference's tables, terminal period, bracket search, interpolation, cdfni."""
import numpy as np
import pytest

from oracle.port import Port
from tests import goldens


@pytest.mark.parametrize("name", list(goldens.CASES))
def test_port_simulator_reproduces_reference_paths(name):
    m = goldens.model_for(name)
    m.prepare()
    g = goldens.load(name)
    sims = Port(m).simulate(g["M"], g["D"], g["init"], g["randstream"], 0)
    e = goldens.sims_errors(sims, g["sims"], g["skipcols"])
    assert e["nan_mismatch"] == 0 and e["inf_mismatch"] == 0 and e["discrete_mismatch"] == 0 and e["max"] < 1e-13, (name, e)


@pytest.mark.parametrize("name", ["deaton1", "deaton2", "cake1", "cake2"])
def test_port_terminal_period_matches_reference(name):
    # one decision => the primary envelope is the identity: rows 1.. of the last period are the END2 grid
    m = goldens.model_for(name)
    m.prepare()
    g = goldens.load(name)
    M, C, V = Port(m).terminal(0, 0)
    ref = g["M"][0][m.nt - 1]
    assert ref.shape[0] == m.ngridm + 1
    assert np.allclose(ref[1:, 0], M, rtol=1e-14, atol=0) and np.allclose(ref[1:, 1], C, rtol=1e-14, atol=0)
    assert np.allclose(ref[1:, 3], V, rtol=1e-13, atol=1e-13)


def test_port_numerics():
    m = goldens.model_for("retirement2")
    m.prepare()
    p = Port(m)
    # Acklam's approximation: stated relative error 1.15e-9 (egdst_lib.c:422-424)
    from scipy.stats import norm
    for q in (1e-6, 0.001, 0.02, 0.02425, 0.3, 0.5, 0.77, 0.97575, 0.99, 1 - 1e-7):
        z = norm.ppf(q)
        assert abs(p.cdfni(q) - z) <= 1.2e-9 * max(1.0, abs(z)), q
    assert p.cdfni(0.0) == -np.inf and p.cdfni(1.0) == np.inf and p.cdfni(0.5) == 0.0
    grid = np.array([0.0, 1.0, 2.0, 4.0, 8.0])
    assert [p.bracket(x, grid) for x in (-1, 0.5, 1.0, 1.5, 3.9, 4.0, 7.0, 100.0)] == [0, 0, 1, 1, 2, 3, 3, 3]
    # type 1 (threshold lookup): the reference's bisection never tests index n-2, so [grid[n-2], grid[n-1]) maps to n-3
    # (egdst_lib.c:150-161) -- a quirk that the restatement and the CUDA path keep for parity
    assert [p.bracket(x, grid, 1) for x in (7.0, 8.0, 9.0)] == [2, 4, 4]
    fun = grid ** 2
    assert p.linter(3.0, grid, fun) == pytest.approx(10.0) and p.linter(-1.0, grid, fun) == pytest.approx(-1.0)
    assert p.linter(10.0, grid, fun) == pytest.approx(64 + 12 * 2)  # linear extrapolation with the last slope
