"""The product's MEX gateways, driven through the MEX shim with a fake @egdstmodel object, against the golden
vectors of the reference's gateways: [M,D,dbg]=egdst_solver(model), sims=egdst_simulator(model,rnd),
res=egdst_call(model,sw,args)."""
import numpy as np
import pytest

from tests import goldens
from tests.mexharness import MexDropIn
from tests.parity import solution_errors

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["deaton2", "retirement2", "occ3", "model2", "humancapital"])
def test_mex_gateways_match_reference_golden(name):
    m = goldens.model_for(name)
    g = goldens.load(name)
    mex = MexDropIn(m)
    M, D = mex.solve()
    e = solution_errors(M, D, g["M"], g["D"])
    assert e["C"] < 1e-9 and e["V"] < 1e-9 and e["TH"] < 1e-9 and e["Dseq"] and e["evf"] < 1e-9, (name, e)
    # infeasible cells stay empty, like the reference's cell arrays
    for ist in range(m.nst):
        for it in range(m.nt):
            assert (M[ist][it] is None or M[ist][it].size == 0) == (g["M"][ist][it] is None)
    sims = mex.simulate(g["M"], g["D"], g["init"], g["randstream"], 0)
    se = goldens.sims_errors(sims, g["sims"], g["skipcols"])
    assert se["nan_mismatch"] == 0 and se["discrete_mismatch"] == 0 and se["max"] < 1e-12, (name, se)


def test_mex_simulator_rejects_short_randstream():
    from oracle.ref import RefError
    m = goldens.model_for("retirement2")
    g = goldens.load("retirement2")
    mex = MexDropIn(m)
    with pytest.raises(RefError, match="randstream is too short"):
        mex.simulate(g["M"], g["D"], g["init"], g["randstream"][:10], 0)


def test_mex_call_matches_direct_evaluation():
    m = goldens.model_for("retirement2")
    g = goldens.load("retirement2")
    mex = MexDropIn(m)
    # utility(it=1, ist=1, id=1 (retire), c): log(c)+duw ; id=2 (work): log(c)
    args = np.array([[1, 1, 1, 2.0], [1, 1, 2, 2.0], [3, 1, 2, 0.5]])
    res = mex.call(g["M"], g["D"], 1, args)
    assert np.allclose(res, [np.log(2.0) + 0.5, np.log(2.0), np.log(0.5)], rtol=0, atol=1e-14)
    # value function at grid nodes of period it=5 equals column 4 of M
    M5 = g["M"][0][4]
    rows = [10, 40, 70]
    args = np.array([[5, 1, M5[r, 0]] for r in rows])
    res = mex.call(g["M"], g["D"], 6, args)
    assert np.allclose(res, M5[rows, 3], rtol=0, atol=1e-10)
