"""Host-side logic that needs no GPU: code generation rules, quadrature, state enumeration, optim_* inference,
model (de)serialisation, and the oracle (compiled reference) against the committed golden vectors."""
import json
import os

import numpy as np
import pytest

from egdst_b200 import codegen, examples
from egdst_b200.model import EgdstModel
from egdst_b200.quadrature import model_quadrature, quadpoints
from tests import goldens
from tests.parity import solution_errors


def test_quadrature_gauss_legendre_on_unit_interval():
    for n in (1, 2, 10, 100):
        x, w = quadpoints(n, 0.0, 1.0)
        assert abs(w.sum() - 1.0) < 1e-13 and np.all((x > 0) & (x < 1)) and np.all(np.diff(x) > 0)
        # exact for polynomials up to degree 2n-1
        for k in range(0, min(2 * n, 12)):
            assert abs((w * x ** k).sum() - 1.0 / (k + 1)) < 1e-12
    q = model_quadrature(10)
    assert q.shape == (20,) and abs(q[:10].sum() - 1) < 1e-13


def test_state_enumeration_first_variable_slowest():
    m = EgdstModel("t")
    m.s = ("a", [0, "a0", 1, "a1"])
    m.s = ("b", [10, "b0", 20, "b1", 30, "b2"])
    assert m.nst == 6 and m.nnst == 2 and m.stm == [2, 3, 3, 1]
    assert m.states.tolist() == [[0, 10], [0, 20], [0, 30], [1, 10], [1, 20], [1, 30]]


def test_std_convert_rules_and_order():
    m = examples.retirement2()
    s = codegen.std_convert(m, "min(savings,cash)+wage_income*(id!=0)+age+dc1+st1")
    assert "MIN(next->savings,curr->cash)" in s and "wage_income(curr,next)" in s and "(curr->it+t0)" in s
    assert "decisions[curr->id+0*nd]" in s and "states[curr->ist+0*nst]" in s


def test_optim_inference_matches_the_shipped_models():
    want = {"retirement2": dict(optim_MUnoD=True, optim_UnoD=False), "deaton2": dict(optim_MUnoD=True, optim_UnoD=True),
            "occ3": dict(optim_MUnoD=True, optim_UnoD=False)}
    for name, w in want.items():
        m = examples.ALL[name](); m.prepare()
        for k, v in w.items():
            assert bool(m.optim[k]) == v, (name, k, m.optim)


def test_model_dict_roundtrip_preserves_the_generated_source():
    for name in examples.ALL:
        m = goldens.model_for(name); m.prepare()
        m2 = EgdstModel.from_dict(json.loads(json.dumps(m.to_dict()))); m2.prepare()
        assert codegen.emit_devspec(m) == codegen.emit_devspec(m2), name
        assert codegen.model_key(m) == codegen.model_key(m2)
        assert (m2.nst, m2.nd, m2.ngridm, m2.ngridmax, m2.ny) == (m.nst, m.nd, m.ngridm, m.ngridmax, m.ny)
        assert np.array_equal(m.param_vector(), m2.param_vector())


def test_duplicate_refs_and_reserved_words_are_rejected():
    m = EgdstModel("t")
    m.param = ("x", "", 1.0)
    with pytest.raises(ValueError):
        m.param = ("x", "", 2.0)
    with pytest.raises(ValueError):
        m.param = ("cash", "", 2.0)


def test_setparam_getparam():
    m = examples.deaton2()
    m.setparam("interest", 0.03, 2, 1.5)
    assert m.getparam("interest") == 0.03 and m.getparam(2) == 1.5
    m.setparam([0.01, 1.0])
    assert m.getparam().tolist() == [0.01, 1.0]
    with pytest.raises(ValueError):
        m.setparam([1.0])


@pytest.mark.parametrize("name", ["cake1", "cake2", "deaton2", "retirement2", "model2", "humancapital"])
def test_oracle_reproduces_golden_vectors(name):
    """Pins the checker: the compiled reference (prebuilt oracle/_ref or built from /root/reference) against the
    committed outputs of the reference on its own example models."""
    from tests.oracles import oracle_for, ref_available
    m = goldens.model_for(name)
    if not ref_available(m):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    g = goldens.load(name)
    orc = oracle_for(m)
    Mr, Dr = orc.solve()
    e = solution_errors(Mr, Dr, g["M"], g["D"])
    assert e["C"] < 1e-12 and e["V"] < 1e-12 and e["TH"] < 1e-12 and e["Dseq"] and e["rowdiff"] == 0, e
    sims = orc.simulate(Mr, Dr, g["init"], g["randstream"], 0)
    se = goldens.sims_errors(sims, g["sims"], g["skipcols"])
    assert se["nan_mismatch"] == 0 and se["discrete_mismatch"] == 0 and se["max"] < 1e-12


def test_cake_closed_forms_in_golden_vectors():
    # cake1: c_t(M) = M/(T-t+1) with T=25 periods; cake2: c_t(M) = M(1-b)/(1-b^(T-t+1)), b=.75
    g1, g2 = goldens.load("cake1"), goldens.load("cake2")
    for it in (0, 5, 10, 24):
        M = g1["M"][0][it]
        ok = M[:, 0] > 1e-6
        assert np.allclose(M[ok, 1], M[ok, 0] / (25 - it), rtol=1e-10)
        M = g2["M"][0][it]
        ok = M[:, 0] > 1e-6
        b = 0.75
        assert np.allclose(M[ok, 1], M[ok, 0] * (1 - b) / (1 - b ** (25 - it)), rtol=1e-10)


def test_continuous_state_spec_motion_rule_and_generated_code():
    # egdstmodel.m:629-648 (name, limits, points -> uniform grid doubling as the values) and :1000-1004 (motion rule)
    m = EgdstModel("c")
    m.s = ("regime", [0, "low", 1, "high"])
    m.s = ("z", [0.5, 1.5], 5)
    assert m.nst == 10 and m.stm == [2, 5, 5, 1]
    z = m.s[1]
    assert z["continuous"] and not z["discrete"] and z["gridpoints"] == 5 and z["grid"] == [0.5, 0.75, 1.0, 1.25, 1.5]
    assert [v["value"] for v in z["values"]] == z["grid"] and m.states[:, 1].tolist() == z["grid"] * 2
    m.trpr = (1, "true", [[0.9, 0.1], [0.2, 0.8]])
    m.trpr = (2, "true", "0.9*st2+0.1")
    assert m.trpr[1]["cases"][0]["prob"] == "0.9*st2+0.1"
    with pytest.raises(ValueError):
        m.d = ("x", [0.0, 1.0], 3)  # only states can be continuous
    with pytest.raises(ValueError):
        m.s = ("w", [2.0, 1.0], 4)  # decreasing grid
    hc = examples.humancapital()
    src = codegen.emit_devspec(hc)
    assert "#define EGDST_NCONT 1" in src and "st1grid[5]" in src and "egdst_gridcell(nval,st1grid" in src
    assert "next->st[0]=" in src  # trpr_cont: the exact value for the simulator (compile.m:556-575)
    c, h = codegen.emit_refspec(hc)
    assert 'st1grid = (double *) mxGetPr(mxGetField(mxGetProperty(Model,0,"s"),0,"grid"));' in c
    assert "bxsearch(nval,(double*)st1grid,(int)stm[0])" in c and "extern double *st1grid;" in h
    # a motion rule may not address the next period's cell (compile.m:531)
    bad = examples.humancapital()
    bad.trpr = (1, "false", "st1+ist1")
    with pytest.raises(ValueError):
        codegen.emit_devspec(bad)
    # the discrete models carry no continuous-state code
    assert "#define EGDST_NCONT 0" in codegen.emit_devspec(examples.retirement2())
    # round trip through plain data
    again = EgdstModel.from_dict(json.loads(json.dumps(hc.to_dict())))
    assert codegen.emit_devspec(again) == src and again.s[0]["grid"] == hc.s[0]["grid"]


# What MATLAB's jsonencode(struct(model)) writes for egdst_examples/model_retirement2.m (hand-written here: neither
# MATLAB nor Octave exist in this image).  jsonencode turns struct arrays into lists of objects, 1x1 struct arrays
# into a single object, cell arrays into lists, empty struct arrays into [], numeric scalars into numbers and logicals
# into true/false; the class stores the trpr matrix [1] as a cell of strings (egdstmodel.m:985-999).
MATLAB_RETIREMENT2_JSON = r'''
{"label":"retire2","t0":1,"T":25,
 "s":{"index":1,"name":"Singleton state","type":"double","discrete":true,"continuous":false,
      "values":{"value":0,"description":"dummy state"},"gridlimits":[],"gridpoints":[],"grid":[]},
 "d":{"index":1,"name":"Labour supply","type":"double","discrete":true,"continuous":false,
      "values":[{"value":0,"description":"retire"},{"value":1,"description":"work"}],"gridlimits":[],"gridpoints":[],"grid":[]},
 "mmax":10,"ngridm":100,"ngridmax":1000,"nthrhmax":10,"ny":10,"a0":-5,
 "discount":"1/(1+interest)","survival":"1.0",
 "u":{"utility":"log(consumption)+duw*(id==0)","marginal":"1/consumption","marginalinverse":"1/mutility","extrap":"log(x)"},
 "transform":{"direct":"log(x+1)","inverse":"exp(x)-1"},
 "budget":{"cashinhand":"savings+wage_income*(id!=0)","marginal":"1+interest"},
 "shock":{"type":"lognormal","mu":"-0.5*sigma*sigma","sigma":"0.25"},
 "trpr":{"varindex":1,"cases":{"condition":"true","prob":[["1.0000000000"]]}},
 "choiceset":{"defaultallow":true,"rules":[]},
 "feasible":{"defaultfeasible":true,"rules":[]},
 "eq":{"ref":"wage_income","type":"next","expression":"wage*shock","description":"Realized wage income"},
 "coef":[],
 "param":[{"ref":"duw","description":"disutility of work","value":0.5},
          {"ref":"interest","description":"return on savings","value":0.045},
          {"ref":"wage","description":"wage (times multiplicator shock)","value":1.05}],
 "cflags":{"TOLERANCE":"1e-10","ZEROCONSUMPTION":"1e-10","DOUBLEPOINT_DELTA":"1e-10","VERBOSE":0},
 "quiet":false,"needtocompile":true,"dir":"tmp_retire2","id":"abc123",
 "optim":{"optim_UasD":true,"optim_MUnoD":true,"optim_UnoD":false,"optim_TRPRnoSH":true},
 "init":[1,0],"nst":1,"nd":2,"nnst":1,"nnd":1,"stm":[1,1],"states":0,"decisions":[0,1]}
'''


def test_build_cli_reads_a_matlab_shaped_dump(tmp_path, capsys):
    """matlab/compile_b200.m writes jsonencode(struct(model)) and runs `python -m egdst_b200.build_cli` on it
    (the replacement of compile.m:754-819): the dump of model_retirement2.m must give the model image of the
    transcribed example, with the switches compile.m:669-747 infers."""
    from egdst_b200 import build, build_cli
    d = json.loads(MATLAB_RETIREMENT2_JSON)
    m = EgdstModel.from_dict(d)
    ref_m = examples.retirement2()
    assert codegen.model_key(m) == codegen.model_key(ref_m)
    assert codegen.emit_devspec(m) == codegen.emit_devspec(ref_m)
    assert codegen.emit_refspec(m) == codegen.emit_refspec(ref_m)
    m.prepare()
    assert {k: bool(v) for k, v in m.optim.items()} == d["optim"]
    assert (m.t0, m.T, m.ngridm, m.ngridmax, m.nthrhmax, m.ny, m.a0, m.mmax) == (1, 25, 100, 1000, 10, 10, -5, 10)
    assert list(m.param_vector()) == [0.5, 0.045, 1.05] and m.cflags["VERBOSE"] == "0"
    # the command compile_b200.m runs (the image of this key is built by __graft_entry__.build(); nothing to compile here)
    if not os.path.isfile(build.library_path(ref_m)):
        pytest.skip("model image not built")
    jf = tmp_path / "model.json"
    jf.write_text(MATLAB_RETIREMENT2_JSON)
    assert build_cli.main([str(jf), "--out", str(tmp_path)]) == 0
    info = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert info["key"] == codegen.model_key(ref_m) and os.path.isfile(info["library"])
    assert {k: bool(v) for k, v in info["optim"].items()} == d["optim"]
