"""The CUDA kernels' logic without a GPU: tools/hostemu compiles the same .cu sources with g++ (one std::thread per
CUDA thread, a team of one CTA) and the C ABI is driven exactly as on the device.  Small models only -- this is a
logic check (chained scans, ticket loops, envelope chains, lookup tables, continuous-state branch), not the product
path: the emulator lives under tools/ and is never loaded by egdst_b200."""
import os
import shutil
import sys

import numpy as np
import pytest

from egdst_b200 import capi, examples
from tests import goldens
from tests.oracles import oracle_for, ref_available
from tests.parity import solution_errors

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools", "hostemu"))

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")


def _emulated(model):
    from build import build  # tools/hostemu/build.py
    model.prepare()
    return capi.ModelLibrary(build(model))


@pytest.mark.parametrize("make,kw", [
    (examples.retirement2, dict(T=6, ngridm=40, ngridmax=200, ny=4)),
    (examples.humancapital, dict(T=5, ngridm=30, ny=3)),
])
def test_emulated_kernels_match_reference(make, kw):
    m = make(**kw)
    if not ref_available(m):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    lib = _emulated(m)
    sol = lib.solve(m)
    assert sol.status()[0] == 0, sol.status()
    orc = oracle_for(m)
    Mr, Dr = orc.solve()
    e = solution_errors(sol.M, sol.D, Mr, Dr)
    assert e["C"] < 1e-9 and e["V"] < 1e-9 and e["TH"] < 1e-8 and e["Dseq"] and e["rowdiff"] == 0, e
    rng = np.random.default_rng(11)
    nsim = 96
    cont = any(v["continuous"] for v in m.s)
    ist0 = np.ones(nsim) if cont else np.full(nsim, float(m.nst))
    init = np.column_stack([ist0, m.a0 + (m.mmax - m.a0) * (0.05 + 0.5 * rng.random(nsim))])
    rs = rng.random(4 * nsim * m.nt)
    sims = lib.simulate(m, lib.import_solution(m, Mr, Dr), init, rs, 0)
    se = goldens.sims_errors(sims, orc.simulate(Mr, Dr, init, rs, 0), skipcols=[3] if cont else [])
    assert se["nan_mismatch"] == 0 and se["discrete_mismatch"] == 0 and se["max"] < 1e-9, se
    # counter-based draws: moments accumulated in the kernel equal the sums over the returned paths
    s2, mom = lib.simulate_philox(m, sol, init, 7, want_sims=True, want_moments=True)
    alive = ~np.isnan(s2)
    assert np.array_equal(mom[2], alive.sum(axis=0).T.astype(float))
    assert np.allclose(mom[0], np.where(alive, s2, 0.0).sum(axis=0).T, rtol=1e-12, atol=1e-9)


def test_emulated_phases_with_a_single_striding_cta(monkeypatch):
    """A team of one CTA (the emulator's only shape) on a grid of 700 points: every phase of the solve kernel strides
    over many work items -- EGM items chained by the look-back scan (EGDST_EGM_P: 24 points per item), several
    rank blocks and merge chunks per envelope job."""
    monkeypatch.setenv("EGDST_EGM_P", "24")
    m = examples.retirement2(T=3, ngridm=700, ngridmax=1500, ny=2, nthrhmax=700)
    if not ref_available(m):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    lib = _emulated(m)
    sol = lib.solve(m)
    assert sol.status()[0] == 0, sol.status()
    Mr, Dr = oracle_for(m).solve()
    e = solution_errors(sol.M, sol.D, Mr, Dr)
    assert e["C"] < 1e-9 and e["V"] < 1e-9 and e["TH"] < 1e-8 and e["Dseq"] and e["rowdiff"] == 0, e


def test_emulated_table_free_path_for_oversized_cells(monkeypatch):
    """EGDST_TABCAP: cells with more rows than the table capacity keep the plain columns and the reference's bisection
    in the EGM step and in the simulator (normally reached only when ngridmax far above 2*ngridm is actually used)."""
    monkeypatch.setenv("EGDST_TABCAP", "12")
    m = examples.retirement2(T=5, ngridm=40, ngridmax=200, ny=3)
    if not ref_available(m):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    lib = _emulated(m)
    sol = lib.solve(m)
    assert sol.status()[0] == 0, sol.status()
    orc = oracle_for(m)
    Mr, Dr = orc.solve()
    e = solution_errors(sol.M, sol.D, Mr, Dr)
    assert e["C"] < 1e-9 and e["V"] < 1e-9 and e["TH"] < 1e-8 and e["Dseq"] and e["rowdiff"] == 0, e
    rng = np.random.default_rng(5)
    nsim = 64
    init = np.column_stack([np.ones(nsim), m.a0 + (m.mmax - m.a0) * (0.05 + 0.5 * rng.random(nsim))])
    rs = rng.random(4 * nsim * m.nt)
    se = goldens.sims_errors(lib.simulate(m, sol, init, rs, 0), orc.simulate(Mr, Dr, init, rs, 0))
    assert se["nan_mismatch"] == 0 and se["discrete_mismatch"] == 0 and se["max"] < 1e-9, se


def test_emulated_cta_scope_matches_grid_scope(monkeypatch):
    """The vector-per-CTA scope of the solve kernel (sweeps of small models) runs the same phases separated by CTA
    barriers: a batched solve in that scope equals the single solves of its vectors."""
    m = examples.deaton2(T=8, ngridm=40, ngridmax=200, ny=4)
    lib = _emulated(m)
    params = np.array([[0.01, 0.9], [0.04, 1.5], [0.025, 1.2]])
    monkeypatch.setenv("EGDST_SOLVE_SCOPE", "cta")
    batch = lib.solve_batch(m, params)
    monkeypatch.setenv("EGDST_SOLVE_SCOPE", "grid")
    for v in range(params.shape[0]):
        assert batch.status(v)[0] == 0
        mv = examples.deaton2(T=8, ngridm=40, ngridmax=200, ny=4, interest=float(params[v, 0]), income=float(params[v, 1]))
        mv.prepare()
        one = lib.solve(mv)
        Mb, Db = batch.cells(v)
        e = solution_errors(Mb, Db, one.M, one.D)
        assert e["C"] < 1e-12 and e["V"] < 1e-12 and e["rowdiff"] == 0, (v, e)


def test_emulated_warp_scope_matches_grid_scope(monkeypatch):
    """The WARP scope (a warp per vector, the warps of a CTA starting every phase together): 6 vectors on CTAs of 4
    warps -- a second round with two idle warps -- and a model with two decisions (secondary envelope, primary
    envelope over decisions) next to the one-decision sweep model."""
    monkeypatch.setenv("EGDST_WARP_G", "4")
    m = examples.deaton2(T=8, ngridm=40, ngridmax=200, ny=4)
    lib = _emulated(m)
    rng = np.random.default_rng(6)
    params = np.column_stack([rng.uniform(0.0, 0.05, 6), rng.uniform(0.75, 1.75, 6)])
    monkeypatch.setenv("EGDST_SOLVE_SCOPE", "warp")
    batch = lib.solve_batch(m, params)
    monkeypatch.setenv("EGDST_SOLVE_SCOPE", "grid")
    for v in range(params.shape[0]):
        assert batch.status(v)[0] == 0
        mv = examples.deaton2(T=8, ngridm=40, ngridmax=200, ny=4, interest=float(params[v, 0]), income=float(params[v, 1]))
        mv.prepare()
        one = lib.solve(mv)
        Mb, Db = batch.cells(v)
        e = solution_errors(Mb, Db, one.M, one.D)
        assert e["C"] < 1e-12 and e["V"] < 1e-12 and e["rowdiff"] == 0, (v, e)
    r = examples.retirement2(ngridm=60, ngridmax=300, T=6, ny=3)
    r.prepare()
    rl = _emulated(r)
    pv = np.array([list(r.param_vector())] * 3)
    monkeypatch.setenv("EGDST_SOLVE_SCOPE", "warp")
    monkeypatch.setenv("EGDST_WARP_G", "3")
    rb = rl.solve_batch(r, pv)
    monkeypatch.setenv("EGDST_SOLVE_SCOPE", "grid")
    one = rl.solve(r)
    for v in range(3):
        assert rb.status(v)[0] == 0
        Mb, Db = rb.cells(v)
        e = solution_errors(Mb, Db, one.M, one.D)
        assert e["C"] < 1e-12 and e["V"] < 1e-12 and e["TH"] < 1e-12 and e["rowdiff"] == 0, (v, e)


def test_emulated_late_zero_consumption_resends_match_reference():
    """egdst_solver.c:1080-1099: a zero-consumption signal AFTER the seed stage rebuilds the rest of the savings grid.
    The means-tested budget of examples.deaton_meanstest fires it in three of six periods (26 re-sends); its grids also
    fold back with a single decision, and in period 1 the reference's envelope2 keeps a double point (M, M+1e-10) whose
    second consumption is -inf (a crossing outside its bracket, on an analytic segment) -- with one decision no primary
    envelope re-sorts the list, so the pair reaches the solution cell and the next period reads it.  All six periods
    must equal the reference row for row."""
    m = examples.deaton_meanstest()
    if not ref_available(m):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    Mr, Dr = oracle_for(m).solve()
    lib = _emulated(m)
    sol = lib.solve(m)
    assert sol.status(0)[0] == 0, sol.status(0)
    assert sol.resends() >= 20, sol.resends()
    from tests.parity import cell_errors
    for it in range(m.nt - 1, -1, -1):
        e = cell_errors(sol.M[0][it], sol.D[0][it], Mr[0][it], Dr[0][it])
        assert e["C"] < 1e-9 and e["V"] < 1e-9 and e["rows"][0] == e["rows"][1], (it, e)
    assert np.isneginf(Mr[0][1][:, 1]).sum() == 1 and np.isneginf(sol.M[0][1][:, 1]).sum() == 1  # the -inf row is there, on both sides


def test_emulated_reused_solution_object_carries_nothing_over():
    """the library keeps one released solution object for re-use by the next solve of the same
    shape.  Solve A, free, solve B (other parameter values), free, import A's cells, simulate -- the MEX simulator's
    flow -- must give A's paths, not paths under B's parameters."""
    kw = dict(T=6, ngridm=40, ngridmax=200, ny=4)
    a = examples.deaton2(income=1.25, **kw)
    b = examples.deaton2(income=3.0, **kw)
    lib = _emulated(a)
    b.prepare()
    sa = lib.solve(a)
    Ma, Da = sa.M, sa.D
    rng = np.random.default_rng(11)
    nsim = 24
    init = np.column_stack([np.ones(nsim), a.a0 + 0.5 * (a.mmax - a.a0) * rng.random(nsim)])
    rs = rng.random(4 * nsim * a.nt)
    live = lib.simulate(a, sa, init, rs, 0)
    del sa                      # released: kept for re-use
    sb = lib.solve(b)           # same shape: the cached object, now solved under income=3.0
    assert sb.status(0)[0] == 0
    del sb
    imported = lib.import_solution(a, Ma, Da)   # the cached object again
    again = lib.simulate(a, imported, init, rs, 0)
    se = goldens.sims_errors(again, live)
    assert se["nan_mismatch"] == 0 and se["discrete_mismatch"] == 0 and se["max"] < 1e-12, se


def test_emulated_call_and_simulate_use_the_parameters_of_the_call():
    """the reference reloads the parameters on every MEX call (egdst_call.c:29-30,
    egdst_simulator.c:59-60), so setparam after a solve changes what call() and sim() compute without a new solve."""
    m = examples.deaton2(T=6, ngridm=40, ngridmax=200, ny=4, interest=0.01)
    lib = _emulated(m)
    sol = lib.solve(m)
    args = np.array([[1.0, 1.0]])
    assert lib.call(m, sol, 3, args)[0] == pytest.approx(1 / 1.01, abs=1e-15)
    rng = np.random.default_rng(12)
    nsim = 16
    init = np.column_stack([np.ones(nsim), m.a0 + 0.5 * (m.mmax - m.a0) * rng.random(nsim)])
    rs = rng.random(4 * nsim * m.nt)
    before = lib.simulate(m, sol, init, rs, 0)
    m.setparam("interest", 0.10)
    assert lib.call(m, sol, 3, args)[0] == pytest.approx(1 / 1.10, abs=1e-15)     # discount = 1/(1+interest), new value
    after = lib.simulate(m, sol, init, rs, 0)
    # the policy tables are those of the old solve, the budget between periods uses the new interest rate: the reference's
    # own simulator, given the same tables and the new parameters, is the checker
    if ref_available(m):
        Mr, Dr = sol.M, sol.D
        theirs = oracle_for(m).simulate(Mr, Dr, init, rs, 0)
        se = goldens.sims_errors(after, theirs)
        assert se["nan_mismatch"] == 0 and se["discrete_mismatch"] == 0 and se["max"] < 1e-9, se
    both = np.isfinite(after) & np.isfinite(before)
    assert np.max(np.abs(after[both] - before[both])) > 1e-3


def test_emulated_means_tested_family_agreement_is_partial_and_tracked():
    """KNOWN PARITY GAP, tracked as a number.  Random run-time configurations of examples.deaton_meanstest (seed 1, 40
    draws; the reference completes without a warning on 19 of them) exercise re-sends after the seed stage, folds with
    a single decision and -- the unsolved part -- flat stretches where runs of the secondary envelope have EXACTLY equal
    values.  There the reference's sequential sweep switches between equal-valued runs (or does not) by its processing
    order and emits, or omits, a double point whose consumption is -inf; the parallel restatement decides such ties by
    "lower index wins" and differs -- that was the working hypothesis; tracing one case bit by bit against the reference's
    captured point lists showed instead (1) cells whose grid steps back need the reference's
    bisection (fixed), (2) the seed point's node sum is partitioned over threads where the reference's is sequential, and
    (3) the value at an envelope crossing is x*slope + intercept here and (x*(f1-f0))/(g1-g0) + intercept in the reference:
    last-bit differences that a run split amplifies.  (3) in the reference's form makes the emulator bit-identical and
    this count 12, but two of 108 fresh GPU draws of the multi-state families then miss 1e-9 (by exactly 2^-24 at a double
    point) where all pass with the slope form, so it is NOT adopted.  Observed at the end of round 2: 10 configurations
    equal the reference row for row and to 1e-9, 1 more to 1e-9 with a different row count, 6 carry a -inf row or extra
    rows on one side only, in 2 this implementation ends in a soft error the reference does not raise.
    The assertion is a ratchet: the number of exact matches must not go down."""
    import warnings
    base = examples.deaton_meanstest()
    if not ref_available(base):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    lib = _emulated(base)
    from tests.parity import cell_errors
    from oracle.ref import RefError
    rng = np.random.default_rng(1)
    valid = exact = close = 0
    report = []
    for trial in range(40):
        kw = dict(T=int(rng.integers(3, 9)), ngridm=int(rng.integers(40, 200)), ny=int(rng.integers(2, 8)), mmax=float(rng.uniform(12, 30)),
                  interest=float(rng.uniform(0.0, 0.04)), income=float(rng.uniform(0.9, 1.6)))
        kw["ngridmax"] = 5 * kw["ngridm"]
        m = examples.deaton_meanstest(**kw)
        m.prepare()
        try:
            Mr, Dr = oracle_for(m).solve()
        except RefError:
            continue  # the reference itself aborts or warns: not a parity case
        valid += 1
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sol = lib.solve(m)
        worst, rows_equal = float("inf"), False
        if sol.status(0)[0] == 0:
            worst, rows_equal = 0.0, True
            for it in range(m.nt):
                e = cell_errors(sol.M[0][it], sol.D[0][it], Mr[0][it], Dr[0][it])
                worst = max(worst, e["C"], e["V"])
                rows_equal = rows_equal and e["rows"][0] == e["rows"][1]
        if worst < 1e-9 and rows_equal:
            exact += 1
        elif worst < 1e-9:
            close += 1
        else:
            report.append((trial, sol.status(0)[0], worst))
    print("means-tested family: %d valid draws, %d exact, %d more within 1e-9, %d differ: %s" % (valid, exact, close, len(report), report))
    assert valid >= 15
    assert exact >= 10, (valid, exact, close, report)


def test_emulated_cells_whose_grid_steps_back_take_the_reference_bisection(monkeypatch):
    """The reference's bisection (bxsearch, egdst_lib.c:138-165) is defined on any array; the direct index of the lookup
    tables only on increasing grids.  Degenerate models produce cells that step back -- here period 1 of a means-tested
    configuration: row 0 is (a0 = 0, 0), row 1 has M = -0.96 -- and the two then pick different intervals for cash between
    the rows.  Such cells are flagged by their table build and looked up by bisection: the simulator on the REFERENCE's
    cells equals the reference's simulator, and the solver's period-0 savings points right above the means-test cut-off,
    whose next-period cash falls into that region, are regular Euler points as in the reference (its adraw trace under
    VERBOSE=4 prints M = 2.241582, 2.434090, 2.604051, ...), not the near-zero-consumption points the direct index gave."""
    kw = {"T": 3, "ngridm": 107, "ny": 5, "mmax": 25.980296058161365, "interest": 0.02452013204212162, "income": 1.542108393353632, "ngridmax": 535}
    m = examples.deaton_meanstest(**kw)
    if not ref_available(m):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    orc = oracle_for(m)
    Mr, Dr = orc.solve()
    assert Mr[0][1][1, 0] < Mr[0][1][0, 0]  # the reference's own period-1 cell steps back
    lib = _emulated(m)
    rng = np.random.default_rng(3)
    nsim = 48
    init = np.column_stack([np.ones(nsim), m.a0 + (m.mmax - m.a0) * rng.random(nsim) * 0.3])
    rs = rng.random(4 * nsim * m.nt)
    ours = lib.simulate(m, lib.import_solution(m, Mr, Dr), init, rs, 0)
    theirs = orc.simulate(Mr, Dr, init, rs, 0)
    se = goldens.sims_errors(ours, theirs)
    assert se["nan_mismatch"] == 0 and se["discrete_mismatch"] == 0 and se["max"] < 1e-9, se
    # the solver's raw period-0 points (debug dump of the host emulator; the final cell does not show them: the secondary
    # envelope drops this stretch, which lies below the first run)
    import struct
    monkeypatch.setenv("EGDST_DEBUG_DUMP_IT", "0")
    sol = lib.solve(m)
    assert sol.status(0)[0] == 0
    blob = open("/tmp/egdst_pt_raw_sd0.bin", "rb").read()
    n = struct.unpack("ii", blob[:8])[0]
    raw = np.frombuffer(blob[16:16 + 24 * n]).reshape(3, n)
    expect = np.array([2.241582, 2.434090, 2.604051, 2.769047, 2.933914, 3.100650, 3.270306, 3.443532, 3.620782])
    assert all(np.min(np.abs(raw[0] - x)) < 1e-6 for x in expect), raw[0, 36:50]
    k = int(np.argmin(np.abs(raw[0] - expect[0])))
    assert raw[1, k] == pytest.approx(0.167442, abs=1e-6)  # consumption there: 0.17, not 4e-10
