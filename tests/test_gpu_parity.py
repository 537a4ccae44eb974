"""GPU parity tests: the CUDA solver and simulator, called through the C ABI, against
(1) the committed golden vectors (outputs of the unmodified reference C) and (2) the oracle
(oracle/_ref when its prebuilt library travelled with the snapshot) on seeded inputs."""
import numpy as np
import pytest

from tests import goldens
from tests.parity import solution_errors

pytestmark = pytest.mark.gpu

TOL = 1e-9


@pytest.fixture(scope="module", params=list(goldens.CASES))
def solved(request):
    name = request.param
    m = goldens.model_for(name)
    m.compile()
    m.solve()
    assert m._solution.status()[0] == 0, m._solution.status()
    return name, m, goldens.load(name)


def test_solution_matches_golden(solved):
    name, m, g = solved
    e = solution_errors(m.M, m.D, g["M"], g["D"])
    assert e["cells"] > 0
    assert e["C"] < TOL and e["V"] < TOL and e["evf"] < TOL and e["TH"] < TOL and e["Dseq"], (name, e)


def test_simulation_matches_golden(solved):
    name, m, g = solved
    m.sim(g["init"], "own_shocks", randstream=g["randstream"])
    e = goldens.sims_errors(m.sims, g["sims"], g["skipcols"])
    assert e["nan_mismatch"] == 0 and e["inf_mismatch"] == 0 and e["discrete_mismatch"] == 0 and e["max"] < TOL, (name, e)


def test_simulation_on_imported_reference_solution(solved):
    """The MEX simulator receives model.M / model.D from the host: import the reference's cells and simulate."""
    name, m, g = solved
    lib = m._capi()
    sol = lib.import_solution(m, g["M"], g["D"])
    sims = lib.simulate(m, sol, g["init"], g["randstream"], 0)
    e = goldens.sims_errors(sims, g["sims"], g["skipcols"])
    assert e["nan_mismatch"] == 0 and e["inf_mismatch"] == 0 and e["discrete_mismatch"] == 0 and e["max"] < 1e-12, (name, e)
