"""egdst_call on the device solution against the reference's egdst_call gateway, and the error paths:
soft errors keep the partial result and carry the reference's message, hard errors raise."""
import warnings

import numpy as np
import pytest

from egdst_b200 import capi, examples
from oracle import ref
from tests import goldens

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def solved():
    m = goldens.model_for("retirement2")
    m.compile()
    m.solve()
    return m, goldens.load("retirement2")


def test_call_matches_reference_gateway(solved):
    m, g = solved
    r = ref.Reference(m)
    rng = np.random.default_rng(5)
    n = 64
    it = rng.integers(1, m.T + 1, n).astype(float)   # it in [t0, T]
    ist = np.ones(n)
    idd = rng.integers(1, m.nd + 1, n).astype(float)
    cons = rng.uniform(0.05, 8.0, n)
    for sw, name, args in (
        (1, "utility", np.column_stack([it, ist, idd, cons])),
        (2, "mutility", np.column_stack([it, ist, idd, cons])),
        (3, "discount", np.column_stack([it, ist])),
        (4, "budget", np.column_stack([np.minimum(it, m.T - 1), ist, idd, rng.uniform(m.a0, 6.0, n), ist, rng.uniform(0.5, 1.5, n)])),
        (5, "mbudget", np.column_stack([np.minimum(it, m.T - 1), ist, idd, rng.uniform(m.a0, 6.0, n), ist, rng.uniform(0.5, 1.5, n)])),
        (6, "vf", np.column_stack([it, ist, rng.uniform(m.a0, m.mmax, n)])),
    ):
        ours = m.call(name, args)
        theirs = r.call(g["M"], g["D"], sw, args)
        fin = np.isfinite(theirs)
        if sw == 6:
            # terminal-period rows: the reference evaluates utility with the decision index left over from the previous
            # row (egdst_call.c never sets curr.id for sw=6); ours takes the terminal cell's decision -- not comparable
            fin &= args[:, 0] < m.T
        assert np.array_equal(np.isnan(ours), np.isnan(theirs)), name
        assert np.allclose(ours[fin], theirs[fin], rtol=1e-9, atol=1e-9), (name, np.abs(ours[fin] - theirs[fin]).max())
    # terminal period: utility of consuming everything under the terminal cell's decision (retire: +duw)
    vT = m.call("vf", np.array([[m.T, 1, 2.0]]))
    assert vT[0] == pytest.approx(np.log(2.0) + 0.5, abs=1e-14)
    # out-of-range arguments give NaN like the reference (egdst_call.c:45-58)
    bad = m.call("utility", np.array([[m.T + 5, 1, 1, 1.0], [1, 7, 1, 1.0], [1, 1, 9, 1.0]]))
    assert np.all(np.isnan(bad))
    with pytest.raises(ValueError):
        m.call("nonsense", [[1, 1]])


def test_soft_error_keeps_partial_result_with_reference_message():
    # too few threshold slots: the reference warns "Not enough space for thresholds..." and returns what it has
    m = examples.retirement2(nthrhmax=2)
    m.compile()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m.solve()
    assert any(issubclass(x.category, capi.EgdstWarning) and "Not enough space for thresholds" in str(x.message) for x in w)
    code, it, ist, idd = m._solution.status()
    assert code == 9 and 0 <= it < m.nt
    assert m.M[0][m.nt - 1] is not None  # the terminal period is there
    r = ref.Reference(m)
    with pytest.raises(ref.RefError, match="Not enough space for thresholds"):
        r.solve(strict=True)


def test_hard_errors_raise():
    m = examples.retirement2()
    m.compile()
    lib = m._capi()
    d = capi.Desc(m)
    d.c.nparam = 7
    h = capi.C.c_void_p()
    assert lib.L.egdst_solve(capi.C.byref(d.c), capi.C.byref(h)) == 2 and "parameters" in lib.last_error()
    d = capi.Desc(m)
    d.c.abi_version = 99
    assert lib.L.egdst_solve(capi.C.byref(d.c), capi.C.byref(h)) == 2 and "abi_version" in lib.last_error()
    d = capi.Desc(m)
    d.c.device = 99
    assert lib.L.egdst_solve(capi.C.byref(d.c), capi.C.byref(h)) == 2 and "device" in lib.last_error()
    m.solve()
    with pytest.raises(capi.EgdstError, match="randstream is too short"):
        lib.simulate(m, m._solution, [[1, 0.5]], np.zeros(3), 0)
    with pytest.raises(RuntimeError):
        examples.deaton2().solve()  # not compiled


def test_mmax_too_small_reports_adraw_failure_like_reference():
    # M(a0) > mmax: "Could not complete initial stage in adraw().. Seems like M(a0)>mmax! Increase mmax!"
    m = examples.deaton2(mmax=2)
    m.compile()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m.solve()
    msgs = " ".join(str(x.message) for x in w)
    r = ref.Reference(m)
    try:
        r.solve(strict=True)
        refmsg = ""
    except ref.RefError as e:
        refmsg = str(e)
    if refmsg:
        assert ("adraw" in refmsg) == ("adraw" in msgs) or ("savings" in refmsg and "savings" in msgs), (refmsg, msgs)
