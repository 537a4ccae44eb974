"""Model code generator: user exec strings -> C / CUDA model functions.

Restates the translation rules of the reference's ``compile`` method
(@egdstmodel/compile.m).  The reference regex-rewrites the user's C-syntax strings
(``StdConvertN``, compile.m:12-64) and emits ``modelspec.c/.h`` (compile.m:183-655),
then infers the ``optim_*`` switches by regex (compile.m:669-747).

Two flavours are produced from one model:

* ``emit_refspec``  -- ``modelspec.c`` / ``modelspec.h`` that the *unmodified* reference C
  sources compile against (MEX API calls included).  Used only to build ``oracle/_ref``.
* ``emit_devspec``  -- one self-contained header, ``modelspec_dev.h``, whose functions take an
  explicit ``const egdst_ctx *cx`` (parameters per instance, needed for batched sweeps) and are
  qualified ``EGDST_FN`` (``__device__ __forceinline__`` under nvcc, ``static inline`` under gcc).
  It is consumed by the CUDA kernels in ``egdst_b200/csrc`` and by the CPU restatement in
  ``oracle/``; it contains only the *model* (user strings), never solver logic.
"""
from __future__ import annotations

import hashlib
import re
from typing import List

RESERVED_GLOBALS = ["t0", "T", "ngridm", "ngridmax", "nthrhmax", "ny", "nd", "nnd", "nst", "nnst", "mmax", "a0"]


def _w(word: str) -> str:
    """MATLAB ``\\<word\\>`` == python ``\\bword\\b``."""
    return r"\b" + word + r"\b"


def std_convert(model, execstr: str) -> str:
    """StdConvertN (compile.m:12-64): same substitutions, same order."""
    s = execstr
    s = re.sub(_w("min"), "MIN", s)
    s = re.sub(_w("max"), "MAX", s)
    s = re.sub(_w("it"), "curr->it", s)
    s = re.sub(_w("age"), "(curr->it+t0)", s)
    s = re.sub(_w("id"), "curr->id", s)
    for i in range(1, len(model.d) + 1):
        s = re.sub(_w("dc%d" % i), "(byval>0?curr->dc[%d]:decisions[curr->id+%d*nd])" % (i - 1, i - 1), s)
    s = re.sub(_w("ist"), "curr->ist", s)
    s = re.sub(_w("ist1"), "next->ist", s)
    for i in range(1, len(model.s) + 1):
        s = re.sub(_w("st%d" % i), "(byval>0?curr->st[%d]:states[curr->ist+%d*nst])" % (i - 1, i - 1), s)
        s = re.sub(_w("st%dn" % i), "(byval>0?next->st[%d]:states[next->ist+%d*nst])" % (i - 1, i - 1), s)
    for eq in model.eq:
        if eq["type"] == "current":
            s = re.sub(_w(re.escape(eq["ref"])), eq["ref"] + "(curr)", s)
        else:
            s = re.sub(_w(re.escape(eq["ref"])), eq["ref"] + "(curr,next)", s)
    s = re.sub(_w("cash"), "curr->cash", s)
    s = re.sub(_w("savings"), "next->savings", s)
    s = re.sub(_w("shock"), "next->shock", s)
    s = re.sub(_w("sigma"), "sigma_param(curr,next)", s)
    s = re.sub(_w("mu"), "mu_param(curr,next)", s)
    s = re.sub(_w("discount"), "discount(curr)", s)
    s = re.sub(_w("survival"), "survival(curr)", s)
    return s


def prohibit(inputstring: str, banned, where: str) -> str:
    """ProhibitString (compile.m:65-85): error if a banned identifier (with optional digits) is used."""
    if isinstance(banned, str):
        banned = [banned]
    for b in banned:
        if re.search(r"\b" + b + r"\d*\b", inputstring, flags=re.IGNORECASE):
            raise ValueError("Error in %s: use of `%s` is not allowed in %s! String: %s" % (where, b, where, inputstring))
    return inputstring


def _lines(expr) -> List[str]:
    return [expr] if isinstance(expr, str) else list(expr)


def _coef_rows(arr) -> List[str]:
    """Coefficient arrays with base-1 padding and ``%18.15f`` rounding (compile.m:199-220)."""
    nr, nc = len(arr), len(arr[0])
    rows = ["{" + ", ".join([" 0.0"] * (nc + 1)).strip() + "},"]
    for i in range(nr):
        body = "{0.0, " + ", ".join("%18.15f" % float(v) for v in arr[i])
        rows.append(body + ("}};" if i == nr - 1 else "},"))
    return rows


# --------------------------------------------------------------------------------------------
# optim_* inference (compile.m:669-747)
# --------------------------------------------------------------------------------------------
def infer_optim(model) -> dict:
    nnd = model.nnd
    u = model.u
    marg = "".join(_lines(u["marginal"]))
    util = "".join(_lines(u["utility"]))
    cls = "dc[" + "  ".join(str(i) for i in range(1, nnd + 1)) + "]"
    # MATLAB: regexp(tmp,['dc[' num2str(1:nnd) ']']) -- a character class of the digits (and spaces)
    dcre = "dc[" + re.escape("".join(str(i) for i in range(1, nnd + 1)) + " ") + "]"
    optim = {}
    optim["optim_MUnoD"] = not (re.search(dcre, marg) or re.search("id", marg))
    optim["optim_UnoD"] = not (re.search(dcre, util) or re.search("id", util))
    # additive separability in consumption and d (compile.m:693-727)
    uasd = True
    dcre0 = "dc[" + re.escape("".join(str(i) for i in range(0, nnd)) + " ") + "]"
    for ln in _lines(u["utility"]):
        tmp = ln
        while re.search(r"\([^+\-()]*\)", tmp):
            tmp = re.sub(r"\(([^+\-()]*)\)", r"[\1]", tmp)
        while re.search(r"\([^()]*(\+|\-)[^()]*\)", tmp):
            tmp = re.sub(r"(\([^()]*)(\+|\-)([^()]*\))", r"\1#\3", tmp)
        for sub in re.split(r"\+|-", tmp):
            if (re.search(dcre0, sub) or re.search("id", sub)) and re.search("consumption", sub):
                uasd = False
                break
        if not uasd:
            break
    if not uasd:
        raise ValueError("Utility is not additively separable in consumption and discrete choices. "
                         "This case is not yet implemented!")
    optim["optim_UasD"] = True
    allpr = ""
    for tr in model.trpr:
        for case in tr["cases"]:
            pr = case["prob"]
            if isinstance(pr, str):
                allpr += "#" + pr
            else:
                allpr += "#" + "".join("".join(r) for r in pr)
    optim["optim_TRPRnoSH"] = not re.search("shock", allpr)
    del cls
    return optim


# --------------------------------------------------------------------------------------------
# flavour (a): reference-compatible modelspec.c / modelspec.h  (oracle/_ref only)
# --------------------------------------------------------------------------------------------
def emit_refspec(model):
    """Return (modelspec_c, modelspec_h) as the reference's compile step would write them
    (compile.m:183-655).  They reference the MEX API and egdst_lib.h, so they only build together
    with the reference sources -- which is exactly their purpose (oracle/_ref)."""
    C: List[str] = []
    H: List[str] = []
    cv = lambda s: std_convert(model, s)  # noqa: E731
    ns, ndv = len(model.s), len(model.d)
    C += ["/*Model specific code for the model '%s'*/" % model.label, '#include "egdst_lib.h"', ""]
    H += ["/*Model specific h for the model '%s'*/" % model.label, "#ifndef MODELSPECguard", "#define  MODELSPECguard", ""]
    H += ["typedef struct curr_variables {int it; int ist; double st[%d]; int id; double dc[%d]; double cash; "
          "double savings; double shock;} PeriodVars;" % (ns, ndv), ""]
    H += ["#define NREQ %d" % len(model.eq), ""]
    for cf in model.coef:
        arr = cf["array"]
        C.append("const double %s[%d][%d] = {" % (cf["ref"], len(arr) + 1, len(arr[0]) + 1))
        C += _coef_rows(arr)
    C.append("")
    for p in model.param:
        H.append("extern double %s; /*Parameter:%s*/" % (p["ref"], p["description"]))
        C.append("double %s; /*Parameter:%s*/" % (p["ref"], p["description"]))
    for i, sv in enumerate(model.s):  # grids of the continuous states (compile.m:229-252)
        if sv["continuous"]:
            H.append("extern double *st%dgrid;" % (i + 1))
            C.append("double *st%dgrid;" % (i + 1))
    H.append("extern double *stgrids[%d]; /*pointers to grids of continous states*/" % ns)
    C += ["double *stgrids[%d];" % ns, ""]
    C += ["void loadcontinuousgrid() {"]
    for i, sv in enumerate(model.s):
        if sv["continuous"]:
            C.append('st%dgrid = (double *) mxGetPr(mxGetField(mxGetProperty(Model,0,"s"),%d,"grid"));' % (i + 1, i))
            C.append("stgrids[%d] = (double *) st%dgrid;" % (i, i + 1))
        else:
            C.append("/* stgrids[%d] is never used */" % i)
    C += ["}", ""]
    H += ["", "void loadcontinuousgrid();"]

    def fn(proto, body_lines):
        C.append(proto + " {")
        H.append(proto + ";")
        C.extend(body_lines)
        C.append("")

    def expr_body(expr, banned, where):
        ls = _lines(expr)
        if isinstance(expr, str):
            return ["return " + cv(prohibit(expr, banned, where)) + ";}"]
        return [cv(prohibit(ln, banned, where)) for ln in ls] + ["}"]

    fn("double discount(PeriodVars *curr)", expr_body(model.discount, ["id", "dc", "cash"], "discount"))
    fn("double survival(PeriodVars *curr)", expr_body(model.survival, ["id", "dc", "cash"], "survival"))
    fn("double utility(PeriodVars *curr,double consumption)",
       ['if (consumption<0) {printf("it=%d ist=%d id=%d consumption=%f ",curr->it,curr->ist,curr->id,consumption);'
        'mexWarnMsgTxt ("Utility function called with negative consumption..");}']
       + expr_body(model.u["utility"], "cash", "utility"))
    fn("double utility_marginal(PeriodVars *curr,double consumption)", expr_body(model.u["marginal"], "cash", "marginal utility"))
    fn("double utility_marginal_inverse(PeriodVars *curr,double mutility)",
       expr_body(model.u["marginalinverse"], "cash", "marginal utility inverse"))
    banned_tr = ["id", "dc", "cash", "savings", "shock"]
    fn("double tr(PeriodVars *curr,double x)", expr_body(model.transform["direct"], banned_tr, "extrapolation function"))
    fn("double trinv(PeriodVars *curr,double x)", expr_body(model.transform["inverse"], banned_tr, "extrapolation function"))
    fn("double cashinhand(PeriodVars *curr,PeriodVars *next)", expr_body(model.budget["cashinhand"], "cash", "cashinhand"))
    fn("double cashinhand_marginal(PeriodVars *curr,PeriodVars *next)",
       expr_body(model.budget["marginal"], "cash", "cashinhand marginal"))
    fn("double mu_param(PeriodVars *curr,PeriodVars *next)", expr_body(model.shock["mu"], "shock", "mu parameter"))
    fn("double sigma_param(PeriodVars *curr,PeriodVars *next)", expr_body(model.shock["sigma"], "shock", "sigma parameter"))
    # choiceset / feasible (compile.m:412-445)
    body = ["int res = %d;" % (1 if model.choiceset["defaultallow"] else 0)]
    for r in model.choiceset["rules"]:
        body.append("if (%s) res = %d; /*%s*/" % (cv(prohibit(r["condition"], "cash", ".choiceset")),
                                                0 if model.choiceset["defaultallow"] else 1, r["description"]))
    fn("int inchoiceset(PeriodVars *curr)", body + ["return res;}"])
    body = ["int res = %d;" % (1 if model.feasible["defaultfeasible"] else 0)]
    for r in model.feasible["rules"]:
        body.append("if (%s) res = %d; /*%s*/" % (cv(prohibit(r["condition"], ["id", "dc", "cash"], ".feasible")),
                                                0 if model.feasible["defaultfeasible"] else 1, r["description"]))
    fn("int feasible(PeriodVars *curr)", body + ["return res;}"])
    for eq in model.eq:
        proto = ("double %s(PeriodVars *curr)" if eq["type"] == "current" else "double %s(PeriodVars *curr,PeriodVars *next)") % eq["ref"]
        ex = eq["expression"]
        fn(proto, ["return " + cv(ex) + ";}"] if isinstance(ex, str) else [cv(ln) for ln in ex] + ["}"])
    fn("void loadparameters ()",
       ['%s=mxGetScalar(mxGetField(mxGetProperty(Model,0,"param"),%d,"value"));' % (p["ref"], i)
        for i, p in enumerate(model.param)] + ["}"])
    # trpr (compile.m:476-551)
    body = ["double nval,res=1.0;", "int varindex, varindex1;"]
    for tr in model.trpr:
        v = tr["varindex"] - 1
        nv = len(model.s[v]["values"])
        body.append("varindex =(curr->ist/(int)stm[nnst+%d])%%(int)stm[%d];" % (v, v))
        body.append("varindex1=(next->ist/(int)stm[nnst+%d])%%(int)stm[%d];" % (v, v))
        first = True
        for case in tr["cases"]:
            body.append(("if (%s) {" if first else "else if (%s) {") % cv(case["condition"]))
            first = False
            if model.s[v]["continuous"]:
                # deterministic motion rule: linear interpolation weights onto the two adjacent grid points
                # (compile.m:527-537)
                g = "st%dgrid" % (v + 1)
                body.append("if (all==1) {")
                body.append("    nval=%s;" % cv(prohibit(case["prob"], "ist1", "motion rules")))
                body.append("    varindex = bxsearch(nval,(double*)%s,(int)stm[%d]);" % (g, v))
                body.append("    if (varindex==varindex1) res*=(%s[varindex+1]-nval)/(%s[varindex+1]-%s[varindex]);" % (g, g, g))
                body.append("    else if (varindex+1==varindex1) res*=(nval-%s[varindex])/(%s[varindex+1]-%s[varindex]);" % (g, g, g))
                body.append("    else return 0.0;")
                body.append("}")
                body.append("}")
                continue
            body.append("  switch (varindex) {")
            for ii in range(nv):
                body.append("  case %d:" % ii)
                body.append("    switch (varindex1) {")
                for jj in range(nv):
                    body.append("      case %d:" % jj)
                    body.append("          res*=%s;" % cv(case["prob"][ii][jj]))
                    body.append("          break;")
                body.append("      default:")
                body.append('          mexErrMsgTxt("Error in trpr: unknown index of the next period state variable");')
                body.append("          break;")
                body.append("    }")
                body.append("    break;")
            body.append("  default:")
            body.append('    mexErrMsgTxt("Error in trpr: unknown index of the current period state variable");')
            body.append("    break;")
            body.append("  }")
            body.append("}")
        body.append("else {")
        body.append('mexErrMsgTxt("Error in trpr: unknown combination of current state and decision (the set of cases is not complete)!");')
        body.append("}")
        body.append("if (res==0.0) return 0.0;")
    fn("double trpr(PeriodVars *curr,PeriodVars *next,int all)", body + ["return res;}"])
    body = ["int varindex;"]
    for tr in model.trpr:  # compile.m:556-575
        v = tr["varindex"] - 1
        if model.s[v]["continuous"]:
            for k, case in enumerate(tr["cases"]):
                body.append(("if (%s) {" if k == 0 else "else if (%s) {") % cv(case["condition"]))
                body.append(" next->st[%d]=%s;" % (v, cv(case["prob"])))
                body.append("}")
    fn("void trpr_cont(PeriodVars *curr,PeriodVars *next)", body + ["}"])
    body = ["int i=0;"]
    for eq in model.eq:
        if eq["type"] == "next":
            body.append("if (next==NULL) out[i++]=mxGetNaN();")
            body.append("else            out[i++]=%s(curr,next);" % eq["ref"])
        else:
            body.append("out[i++]=%s(curr);" % eq["ref"])
    fn("void eqs_sim(PeriodVars *curr,PeriodVars *next,double *out)", body + ["}"])
    H += ["", "#endif", ""]
    return "\n".join(C) + "\n", "\n".join(H) + "\n"


# --------------------------------------------------------------------------------------------
# flavour (b): ctx-explicit header for the CUDA kernels and the CPU restatement
# --------------------------------------------------------------------------------------------
def _ctxify(model, s: str) -> str:
    """After StdConvertN: bind globals, tables and parameters to the per-instance context ``cx``
    and thread ``cx`` through calls to other generated functions."""
    for i, p in enumerate(model.param):
        s = re.sub(_w(re.escape(p["ref"])), "cx->param[%d]" % i, s)
    for g in RESERVED_GLOBALS:
        s = re.sub(r"(?<![>\w.])" + g + r"\b", "cx->" + g, s)
    s = re.sub(r"(?<![>\w.])byval\b", "cx->byval", s)
    s = re.sub(r"(?<![>\w.])decisions\[", "cx->decisions[", s)
    s = re.sub(r"(?<![>\w.])states\[", "cx->states[", s)
    s = re.sub(r"(?<![>\w.])stm\[", "cx->stm[", s)
    s = s.replace("sigma_param(curr,next)", "sigma_param(cx,curr,next)")
    s = s.replace("mu_param(curr,next)", "mu_param(cx,curr,next)")
    s = s.replace("discount(curr)", "discount(cx,curr)")
    s = s.replace("survival(curr)", "survival(cx,curr)")
    for eq in model.eq:
        s = s.replace(eq["ref"] + "(curr,next)", eq["ref"] + "(cx,curr,next)")
        s = s.replace(eq["ref"] + "(curr)", eq["ref"] + "(cx,curr)")
    return s


def shock_independent_of_savings(model) -> bool:
    """True when neither the shock parameters nor the transition probabilities can depend on end-of-period
    savings (the strings, after equation references are expanded one level, mention neither ``savings`` nor ``cash``
    nor ``shock``).  Then the quadrature shocks and their probabilities are the same for every point of the savings
    grid and the EGM kernel computes them once per CTA instead of once per node."""
    import re as _re
    eqrefs = {e["ref"]: (e["expression"] if isinstance(e["expression"], str) else " ".join(e["expression"])) for e in model.eq}

    def text(x):
        return x if isinstance(x, str) else " ".join(str(v) for v in x)

    def depends(expr, depth=0):
        expr = text(expr)
        if _re.search(r"\b(savings|cash|shock)\b", expr):
            return True
        if depth < 4:
            for ref, body in eqrefs.items():
                if _re.search(r"\b" + _re.escape(ref) + r"\b", expr) and depends(body, depth + 1):
                    return True
        return False

    if depends(model.shock["mu"]) or depends(model.shock["sigma"]):
        return False
    for t in model.trpr:
        for c in t["cases"]:
            if depends(c["condition"]):
                return False
            if isinstance(c["prob"], str):  # motion rule of a continuous state
                if depends(c["prob"]):
                    return False
            elif any(depends(x) for row in c["prob"] for x in row):
                return False
    return True


def emit_devspec(model) -> str:
    """One header with every model function of the reference's modelspec (compile.m:254-655),
    ctx-explicit.  ``EGDST_FN`` / ``EGDST_CONST`` / ``egdst_ctx`` come from ``egdst_modelctx.h``."""
    cv = lambda s: _ctxify(model, std_convert(model, s))  # noqa: E731
    ns, ndv = max(len(model.s), 1), max(len(model.d), 1)
    cont = [i for i, sv in enumerate(model.s) if sv["continuous"]]
    L: List[str] = []
    L += ["/* generated by egdst_b200.codegen for model '%s' -- do not edit */" % model.label,
          "#ifndef EGDST_MODELSPEC_DEV_H", "#define EGDST_MODELSPEC_DEV_H",
          "#define EGDST_NNST %d" % ns, "#define EGDST_NND %d" % ndv,
          "#define EGDST_NREQ %d" % len(model.eq), "#define EGDST_NPARAM %d" % len(model.param),
          "#define EGDST_DISTRIB %d" % (1 if model.shock["type"] == "lognormal" else 2),
          "#define EGDST_SHOCK_INDEP_A %d" % (1 if shock_independent_of_savings(model) else 0),
          "#define EGDST_NCONT %d" % len(cont)]
    # the optim_* switches are functions of the exec strings (compile.m:669-747), hence constants of the image: the
    # kernels specialise on them at compile time and the host checks the descriptor against them
    opt = infer_optim(model)
    L += ["#define EGDST_OPT_%s %d" % (k[len("optim_"):].upper(), 1 if opt[k] else 0)
          for k in ("optim_UasD", "optim_MUnoD", "optim_UnoD", "optim_TRPRnoSH")]
    if float(getattr(model, "sigma_eps", 0.0) or 0.0) > 0.0:
        # taste-shock smoothing (an extension without a reference counterpart) is compiled into its own image, so that the
        # reference-parity images carry none of its code; the VALUE of sigma_eps stays a run-time property
        L += ["#define EGDST_SMOOTHING 1"]
    L += ['#include "egdst_modelctx.h"', ""]
    if cont:
        # grids of the continuous states (the reference loads them from model.s(i).grid at run time, compile.m:239-247;
        # here they are part of the compiled image: re-defining the state variable re-generates it)
        for i in cont:
            g = model.s[i]["grid"]
            L.append("EGDST_CONST double st%dgrid[%d] = {%s};" % (i + 1, len(g), ",".join("%.17g" % x for x in g)))
        L.append("EGDST_CONST int egdst_contvar[EGDST_NCONT] = {%s};  /* position in st[] of each continuous state */"
                 % ",".join(str(i) for i in cont))
        L.append("EGDST_FN const double *egdst_contgrid(int k) {")
        L.append("  switch (k) {")
        for k, i in enumerate(cont[:-1]):
            L.append("  case %d: return st%dgrid;" % (k, i + 1))
        L.append("  default: return st%dgrid;" % (cont[-1] + 1))
        L.append("  }")
        L.append("}")
        L.append("")
    for cf in model.coef:
        arr = cf["array"]
        L.append("EGDST_CONST double %s[%d][%d] = {" % (cf["ref"], len(arr) + 1, len(arr[0]) + 1))
        L += _coef_rows(arr)
    P1 = "const egdst_ctx *cx,const PeriodVars *curr"
    P2 = P1 + ",const PeriodVars *next"
    protos = [
        ("double", "discount", P1), ("double", "survival", P1),
        ("double", "utility", P1 + ",double consumption"), ("double", "utility_marginal", P1 + ",double consumption"),
        ("double", "utility_marginal_inverse", P1 + ",double mutility"),
        ("double", "tr", P1 + ",double x"), ("double", "trinv", P1 + ",double x"),
        ("double", "cashinhand", P2), ("double", "cashinhand_marginal", P2),
        ("double", "mu_param", P2), ("double", "sigma_param", P2),
        ("int", "inchoiceset", P1), ("int", "feasible", P1),
    ]
    for eq in model.eq:
        protos.append(("double", eq["ref"], P1 if eq["type"] == "current" else P2))
    protos.append(("double", "trpr", P2 + ",int all"))
    protos.append(("void", "trpr_cont", P1 + ",PeriodVars *next"))
    protos.append(("void", "eqs_sim", P2 + ",double *out"))
    for rt, nm, ar in protos:
        L.append("EGDST_FN %s %s(%s);" % (rt, nm, ar))
    L.append("")

    def fn(rt, nm, ar, body):
        L.append("EGDST_FN %s %s(%s) {" % (rt, nm, ar))
        L.extend(body)
        L.append("")

    def eb(expr, banned, where):
        if isinstance(expr, str):
            return ["return " + cv(prohibit(expr, banned, where)) + ";}"]
        return [cv(prohibit(ln, banned, where)) for ln in expr] + ["}"]

    fn("double", "discount", P1, eb(model.discount, ["id", "dc", "cash"], "discount"))
    fn("double", "survival", P1, eb(model.survival, ["id", "dc", "cash"], "survival"))
    fn("double", "utility", P1 + ",double consumption", eb(model.u["utility"], "cash", "utility"))
    fn("double", "utility_marginal", P1 + ",double consumption", eb(model.u["marginal"], "cash", "marginal utility"))
    fn("double", "utility_marginal_inverse", P1 + ",double mutility", eb(model.u["marginalinverse"], "cash", "marginal utility inverse"))
    banned_tr = ["id", "dc", "cash", "savings", "shock"]
    fn("double", "tr", P1 + ",double x", eb(model.transform["direct"], banned_tr, "extrapolation function"))
    fn("double", "trinv", P1 + ",double x", eb(model.transform["inverse"], banned_tr, "extrapolation function"))
    fn("double", "cashinhand", P2, eb(model.budget["cashinhand"], "cash", "cashinhand"))
    fn("double", "cashinhand_marginal", P2, eb(model.budget["marginal"], "cash", "cashinhand marginal"))
    fn("double", "mu_param", P2, eb(model.shock["mu"], "shock", "mu parameter"))
    fn("double", "sigma_param", P2, eb(model.shock["sigma"], "shock", "sigma parameter"))
    body = ["int res = %d;" % (1 if model.choiceset["defaultallow"] else 0)]
    for r in model.choiceset["rules"]:
        body.append("if (%s) res = %d;" % (cv(prohibit(r["condition"], "cash", ".choiceset")),
                                           0 if model.choiceset["defaultallow"] else 1))
    fn("int", "inchoiceset", P1, body + ["return res;}"])
    body = ["int res = %d;" % (1 if model.feasible["defaultfeasible"] else 0)]
    for r in model.feasible["rules"]:
        body.append("if (%s) res = %d;" % (cv(prohibit(r["condition"], ["id", "dc", "cash"], ".feasible")),
                                           0 if model.feasible["defaultfeasible"] else 1))
    fn("int", "feasible", P1, body + ["return res;}"])
    for eq in model.eq:
        ex = eq["expression"]
        fn("double", eq["ref"], P1 if eq["type"] == "current" else P2,
           ["return " + cv(ex) + ";}"] if isinstance(ex, str) else [cv(ln) for ln in ex] + ["}"])
    body = ["double nval,res=1.0;", "int varindex, varindex1;", "(void)all; (void)nval;"]
    for tr in model.trpr:
        v = tr["varindex"] - 1
        nv = len(model.s[v]["values"])
        if ns == 1:  # a single state variable: its index is the state index (stride 1, size nst)
            body.append("varindex =curr->ist;")
            body.append("varindex1=next->ist;")
        else:
            body.append("varindex =(curr->ist/(int)cx->stm[cx->nnst+%d])%%(int)cx->stm[%d];" % (v, v))
            body.append("varindex1=(next->ist/(int)cx->stm[cx->nnst+%d])%%(int)cx->stm[%d];" % (v, v))
        first = True
        for case in tr["cases"]:
            body.append(("if (%s) {" if first else "else if (%s) {") % cv(case["condition"]))
            first = False
            if model.s[v]["continuous"]:  # compile.m:527-537
                g = "st%dgrid" % (v + 1)
                body.append("  if (all==1) {")
                body.append("    nval=%s;" % cv(prohibit(case["prob"], "ist1", "motion rules")))
                body.append("    varindex = egdst_gridcell(nval,%s,(int)cx->stm[%d]);" % (g, v))
                body.append("    if (varindex==varindex1) res*=(%s[varindex+1]-nval)/(%s[varindex+1]-%s[varindex]);" % (g, g, g))
                body.append("    else if (varindex+1==varindex1) res*=(nval-%s[varindex])/(%s[varindex+1]-%s[varindex]);" % (g, g, g))
                body.append("    else return 0.0;")
                body.append("  }")
                body.append("}")
                continue
            body.append("  switch (varindex) {")
            for ii in range(nv):
                body.append("  case %d:" % ii)
                body.append("    switch (varindex1) {")
                for jj in range(nv):
                    body.append("      case %d: res*=%s; break;" % (jj, cv(case["prob"][ii][jj])))
                body.append("      default: EGDST_MODEL_FAIL(cx,EGDST_ERR_TRPR_INDEX); break;")
                body.append("    }")
                body.append("    break;")
            body.append("  default: EGDST_MODEL_FAIL(cx,EGDST_ERR_TRPR_INDEX); break;")
            body.append("  }")
            body.append("}")
        body.append("else { EGDST_MODEL_FAIL(cx,EGDST_ERR_TRPR_CASES); }")
        body.append("if (res==0.0) return 0.0;")
    fn("double", "trpr", P2 + ",int all", body + ["return res;}"])
    body = ["(void)cx; (void)curr; (void)next;"]
    for tr in model.trpr:  # compile.m:556-575: the exact next-period values of the continuous states (simulator)
        v = tr["varindex"] - 1
        if model.s[v]["continuous"]:
            for k, case in enumerate(tr["cases"]):
                body.append(("if (%s) {" if k == 0 else "else if (%s) {") % cv(case["condition"]))
                body.append("  next->st[%d]=%s;" % (v, cv(case["prob"])))
                body.append("}")
    fn("void", "trpr_cont", P1 + ",PeriodVars *next", body + ["}"])
    body = ["int i=0;", "(void)i;"]
    for eq in model.eq:
        if eq["type"] == "next":
            body.append("if (next==0) out[i++]=EGDST_NAN; else out[i++]=%s(cx,curr,next);" % eq["ref"])
        else:
            body.append("out[i++]=%s(cx,curr);" % eq["ref"])
    fn("void", "eqs_sim", P2 + ",double *out", body + ["}"])
    L += ["#endif", ""]
    return "\n".join(L)


def model_key(model) -> str:
    """Stable key of the *compiled image*: depends on the generated source (exec strings, structure)
    but not on runtime properties (grids, horizon, parameter values)."""
    src = emit_devspec(model)
    return re.sub(r"\W", "", model.label)[:16] + "_" + hashlib.sha1(src.encode()).hexdigest()[:10]
