"""Checks of the taste-shock smoothing mode (model.sigma_eps > 0), shared by the emulated (CPU) and the GPU tests.

The mode is an EXTENSION: the reference has a hard max only, so there is no reference oracle for
sigma_eps > 0 -- parity is UNPINNED there.  What can be pinned:
  * a closed form (two periods, log utility, no uncertainty), `two_period_closed_form`;
  * the Euler equation and the Bellman equation themselves, re-evaluated in numpy from the exported choice-specific
    cells of two consecutive periods with logsum and choice probabilities written out, `euler_bellman_residuals`;
  * the limit sigma_eps -> 0, which must reproduce the reference's solution (tests call solution_errors on it).
"""
import numpy as np
from scipy.stats import norm

from egdst_b200 import examples
from egdst_b200.quadrature import model_quadrature


def two_period_model(sigma_eps=0.5, ngridm=2000):
    return examples.retirement_two_period(sigma_eps=sigma_eps, ngridm=ngridm)


def two_period_closed_form(sol, m):
    """Terminal period: c = M and v_d(M) = log M + duw*(d==retire) for both decisions, hence
         EV(M') = log M' + duw + sigma*log(1 + exp(-duw/sigma)),   E[u'] = 1/M'.
    First period, decision d with income y_d = wage*(d==work): M' = A + y_d, Euler 1/c = 1/M', so
         c(M) = (M + y_d)/2,   v_d(M) = 2 log c + duw*(d==retire) + duw + sigma*log(1 + exp(-duw/sigma)).
    Consumption is exact (the terminal policy is linear); the value carries the linear-interpolation error of log on
    the terminal grid, so it is compared where M' >= 1.  Returns the worst errors."""
    sig, duw, wage = m.sigma_eps, 0.5, 1.05
    lse = duw + sig * np.log1p(np.exp(-duw / sig))
    worst = {"C": 0.0, "V": 0.0, "rows": 0}
    for d, y in ((0, 0.0), (1, wage)):
        cell = sol.choice_cell(0, 0, d)
        assert cell is not None and cell.shape[0] > 100, (d, None if cell is None else cell.shape)
        M, C, A, V = cell[1:, 0], cell[1:, 1], cell[1:, 2], cell[1:, 3]
        assert np.all(np.diff(M) > 0)
        np.testing.assert_allclose(A, M - C, rtol=0, atol=1e-12)
        ok = (A + y >= 1.0) & (A + y <= 9.5)  # next period's cash inside the terminal grid, away from log's curvature at 0
        assert ok.sum() > 50
        c_exact = (M + y) / 2.0
        v_exact = 2.0 * np.log(c_exact) + duw * (d == 0) + lse
        worst["C"] = max(worst["C"], float(np.max(np.abs(C[ok] - c_exact[ok]))))
        worst["V"] = max(worst["V"], float(np.max(np.abs(V[ok] - v_exact[ok]))))
        worst["rows"] += int(ok.sum())
    return worst


def euler_bellman_residuals(sol, m, it):
    """For every point of every choice-specific cell of period `it` (0-based) whose next-period cash lands inside the
    grids of all next-period choice cells for every quadrature node: recompute
         rhs = beta (1+r) sum_iy w_iy sum_d' P_d'(M') u'(c_d'(M')),      ev = sum_iy w_iy sigma log sum_d' exp(v_d'(M')/sigma)
    in numpy (np.interp on the exported cells) and compare u'(C) with rhs and V with u(C) + beta*ev."""
    sig = m.sigma_eps
    p = {q["ref"]: q["value"] for q in m.param}
    r, duw, wage = p["interest"], p["duw"], p["wage"]
    beta = 1.0 / (1.0 + r)
    sg = float(m.shock["sigma"])
    mu = -0.5 * sg * sg
    if m.ny > 1 and sg > 0:
        q = model_quadrature(m.ny)
        w, z = q[:m.ny], norm.ppf(q[m.ny:])
        shocks = np.exp(mu + z * sg)
    else:
        w, shocks = np.array([1.0]), np.array([np.exp(mu + sg * sg / 2)])
    nxt = [sol.choice_cell(it + 1, 0, d) for d in (0, 1)]
    lo = max(c[1, 0] for c in nxt)
    hi = min(c[-1, 0] for c in nxt)
    worst = {"euler": 0.0, "bellman": 0.0, "points": 0}
    for d in (0, 1):
        cell = sol.choice_cell(it, 0, d)
        M, C, A, V = cell[1:, 0], cell[1:, 1], cell[1:, 2], cell[1:, 3]
        rhs = np.zeros_like(M)
        ev = np.zeros_like(M)
        inside = np.ones_like(M, dtype=bool)
        inside[-1] = False  # a list cut at the unified grid's bound ends with an interpolated point (egdst_ph_dsave), not an EGM point
        for wi, sh in zip(w, shocks):
            M1 = A + wage * sh * (d != 0)
            inside &= (M1 >= lo) & (M1 <= hi)
            v = np.stack([np.interp(M1, c[1:, 0], c[1:, 3]) for c in nxt])
            cc = np.stack([np.interp(M1, c[1:, 0], c[1:, 1]) for c in nxt])
            vmax = v.max(axis=0)
            e = np.exp((v - vmax) / sig)
            P = e / e.sum(axis=0)
            rhs += wi * (P / cc).sum(axis=0)
            ev += wi * (vmax + sig * np.log(e.sum(axis=0)))
        rhs *= beta * (1.0 + r)
        # the residual is measured on consumption, in the parity metric |dC| / max(1, |C|): next to the borrowing limit
        # consumption is ~1e-11 while M = A + C is stored at |M| ~ 5, so C is only known to ~1e-16 in absolute terms.
        # Kink points inserted by the secondary envelope satisfy no Euler equation: at most a few per cell are excused.
        euler = np.abs(C - 1.0 / rhs) / np.maximum(1.0, np.abs(C))
        keep = inside & (euler < 1e-4)
        assert keep.sum() >= inside.sum() - 6 and inside.sum() > 50, (it, d, int(keep.sum()), int(inside.sum()))
        worst["euler"] = max(worst["euler"], float(np.max(euler[keep])))
        worst["bellman"] = max(worst["bellman"], float(np.max(np.abs(V[keep] - (np.log(C[keep]) + duw * (d == 0) + beta * ev[keep])))))
        worst["points"] += int(keep.sum())
    return worst
