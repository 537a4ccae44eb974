/* egdst_port.c -- plain-C, single-threaded RESTATEMENT of the parts of the reference whose algorithm fits on a
 * page: the shared numerics, the terminal period and the forward simulator.
 *
 * TEST INFRASTRUCTURE ONLY (oracle "kind: port"): only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this.  The product (egdst_b200) never does.  Written from the behavioural
 * description of the reference, function by function:
 *   port_bracket        @egdstmodel/egdst_lib.c:138-165   (bxsearch_common / bxsearch)
 *   port_linter         @egdstmodel/egdst_lib.c:168-176
 *   port_cdfni          @egdstmodel/egdst_lib.c:435-519   (Acklam's rational approximation)
 *   port_cdfinv         @egdstmodel/egdst_lib.c:66-101    (DISTRIB 1 lognormal, 2 normal)
 *   port_terminal       @egdstmodel/egdst_solver.c:452-475 (END2 closed-form grid)
 *   port_policy         @egdstmodel/egdst_simulator.c:145-199
 *   port_simulate       @egdstmodel/egdst_simulator.c:47-117, 204-383, output rows :122-143
 * The backward-induction solver is NOT restated here: its checker is the unmodified reference itself, compiled
 * from /root/reference (oracle/ref.py, "kind: reference").  Pinned by tests/test_cpu_port.py against the golden
 * vectors of tests/golden/ (outputs of the reference on its own example models).
 * The model functions come from the generated modelspec_dev.h (user strings only, no solver logic).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "modelspec_dev.h"
#include "egdst_b200.h"

/* index of the interpolation interval: 0 below grid[1]; the last interval (type 0) or last point (type 1) above;
 * otherwise the largest interior i with grid[i] <= x, found by bisection over [1, n-2] */
int port_bracket(double x, const double *grid, int n, int type) {
    int lo = 1, hi = n - 2;
    if (x < grid[1]) return 0;
    if (type == 0 && x >= grid[n - 2]) return n - 2;
    if (type == 1 && x >= grid[n - 1]) return n - 1;
    while (hi - lo > 1) {
        int mid = (hi + lo) / 2;
        if (grid[mid] > x) hi = mid; else lo = mid;
    }
    return lo;
}

/* two-point formula on the bracketing interval; extrapolates linearly outside the grid */
double port_linter(double x, int n, const double *grid, const double *fun) {
    int i = port_bracket(x, grid, n, 0);
    double w = grid[i + 1] - grid[i];
    return fun[i + 1] * (x - grid[i]) / w + fun[i] * (grid[i + 1] - x) / w;
}

double port_cdfni(double p) {
    static const double a[6] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                                1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[5] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                                6.680131188771972e+01, -1.328068155288572e+01};
    static const double c[6] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                                -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
    static const double d[4] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00, 3.754408661907416e+00};
    const double low = 0.02425, high = 1 - 0.02425;
    double q, r;
    if (p < 0 || p > 1) return 0.0;
    if (p == 0) return -INFINITY;
    if (p == 1) return INFINITY;
    if (p < low) {
        q = sqrt(-2 * log(p));
        return (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    }
    if (p > high) {
        q = sqrt(-2 * log(1 - p));
        return -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    }
    q = p - 0.5;
    r = q * q;
    return (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
           (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
}

double port_cdfinv(double p, double mu, double sigma) {
#if EGDST_DISTRIB == 1
    return exp(sigma * port_cdfni(p) + mu);
#else
    return sigma * port_cdfni(p) + mu;
#endif
}

static double port_expectation(const egdst_ctx *cx, const PeriodVars *curr, const PeriodVars *next) {
#if EGDST_DISTRIB == 1
    double s = sigma_param(cx, curr, next);
    return exp(mu_param(cx, curr, next) + s * s / 2);
#else
    return mu_param(cx, curr, next);
#endif
}

static void port_ctx(const egdst_desc *d, egdst_ctx *cx) {
    int i;
    memset(cx, 0, sizeof(*cx));
    cx->t0 = d->t0; cx->T = d->T; cx->ngridm = d->ngridm; cx->ngridmax = d->ngridmax; cx->nthrhmax = d->nthrhmax;
    cx->ny = d->ny; cx->nd = d->nd; cx->nnd = d->nnd; cx->nst = d->nst; cx->nnst = d->nnst;
    cx->optim_UasD = d->optim_UasD; cx->optim_MUnoD = d->optim_MUnoD; cx->optim_UnoD = d->optim_UnoD; cx->optim_TRPRnoSH = d->optim_TRPRnoSH;
    cx->mmax = d->mmax; cx->a0 = d->a0;
    cx->tolerance = d->tolerance; cx->zeroconsumption = d->zeroconsumption; cx->doublepoint_delta = d->doublepoint_delta;
    cx->stm = d->stm; cx->states = d->states; cx->decisions = d->decisions;
    for (i = 0; i < EGDST_NPARAM; i++) cx->param[i] = d->params[i];
}

/* terminal period of decision id in state ist: grid equally spaced in tr() space between ZEROCONSUMPTION and mmax,
 * everything is consumed, value = utility.  M, C, V hold ngridm doubles.  Returns 0, or 1 if (ist,id) is not admissible. */
int port_terminal(const egdst_desc *d, int ist, int id, double *M, double *C, double *V) {
    egdst_ctx cx;
    PeriodVars cur;
    int i, n = d->ngridm;
    double m1, m2;
    port_ctx(d, &cx);
    memset(&cur, 0, sizeof(cur));
    cur.it = d->T - d->t0; cur.ist = ist; cur.id = id;
    if (feasible(&cx, &cur) != 1 || inchoiceset(&cx, &cur) != 1) return 1;
    m1 = tr(&cx, &cur, cx.zeroconsumption);
    m2 = tr(&cx, &cur, cx.mmax);
    for (i = 0; i < n; i++) {
        M[i] = trinv(&cx, &cur, m1 + i * (m2 - m1) / (n - 1));
        C[i] = M[i];
        V[i] = utility(&cx, &cur, C[i]);
    }
    return 0;
}

/* one cell of the solution in the packed export layout (rows x 4 column-major: M, C, A, V; thresholds rows x 2) */
typedef struct { const double *M, *C, *V, *dec, *th; int n, nth; } port_cell;

/* consumption, savings, decision and value at `cash` from the period's tables */
static void port_policy(const egdst_ctx *cx, const port_cell *cell, PeriodVars *cur, double *c, double *vf) {
    int k = 0;
    *c = port_linter(cur->cash, cell->n, cell->M, cell->C);
    cur->savings = cur->cash - *c;
    while (k < cell->nth && cur->cash >= cell->th[k]) k++;   /* last threshold not above cash */
    cur->id = (int)cell->dec[k > 0 ? k - 1 : 0];
    for (int i = 0; i < cx->nnd; i++) cur->dc[i] = cx->decisions[i * cx->nd + cur->id];
    if (cur->cash < cell->M[1] && cell->V[0] > -INFINITY)
        *vf = utility(cx, cur, *c) + discount(cx, cur) * cell->V[0];   /* credit-constrained branch: exact */
    else
        *vf = port_linter(cur->cash, cell->n, cell->M, cell->V);
}

/* sims: [nsimout, nt, nsim] column-major, NaN where the agent is dead or was skipped.  Returns the number of
 * skipped agents (bad initial state / cash).  With continuous state variables (EGDST_NCONT > 0) the exact values are
 * carried in st[] and the policy is the multilinear mix of the surrounding grid cells (egdst_simulator.c:309-365). */
int port_simulate(const egdst_desc *d, const int *mlen, const int *thlen, const double *Mbuf, const double *Dbuf,
                  const double *init, int nsim, const double *rs, int rndtype, double *sims) {
    egdst_ctx cx;
    const int nt = d->T - d->t0 + 1, nst = d->nst, nso = 11 + d->nnst + d->nnd + d->neq;
    port_cell *cells = (port_cell *)calloc((size_t)nt * nst, sizeof(port_cell));
    size_t om = 0, ot = 0;
    int skipped = 0;
    double eqs[EGDST_NREQ > 0 ? EGDST_NREQ : 1];
    port_ctx(d, &cx);
    cx.byval = EGDST_NCONT > 0 ? 1 : 0;   /* by index when all state variables are discrete, else by value (egdst_simulator.c:91-92) */
    for (int c = 0; c < nt * nst; c++) {
        port_cell *pc = cells + c;
        pc->n = mlen[c]; pc->nth = thlen[c];
        pc->M = Mbuf + om; pc->C = pc->M + pc->n; pc->V = pc->M + 3 * (size_t)pc->n;
        pc->dec = Dbuf + ot; pc->th = pc->dec + pc->nth;
        om += 4 * (size_t)pc->n; ot += 2 * (size_t)pc->nth;
    }
    for (size_t i = 0; i < (size_t)nso * nt * nsim; i++) sims[i] = NAN;
    for (int isim = 0; isim < nsim; isim++) {
        const double *u = rs + (rndtype == 1 ? 0 : (size_t)4 * nt * isim);
        const int ist0 = (int)init[isim] - 1;
        const double m0 = init[nsim + isim];
        PeriodVars cur;
        double mu = NAN, sigma = NAN, c = 0, vf = 0;
        int k = 0;
        memset(&cur, 0, sizeof(cur));
        if (ist0 < 0 || ist0 >= nst || m0 < cx.a0 || m0 > cx.mmax) { skipped++; continue; }
        cur.ist = ist0; cur.cash = m0; cur.shock = NAN;
        for (int i = 0; i < cx.nnst; i++) cur.st[i] = cx.states[i * nst + ist0];
        if (!feasible(&cx, &cur)) { skipped++; continue; }
        for (int it = 0; it < nt; it++) {
            if (it == 0) {
                eqs_sim(&cx, &cur, (const PeriodVars *)0, eqs);
            } else {
                PeriodVars nx = cur;
                double r_state = u[k++], r_shock = u[k++], r_alive = u[k++];
                int ist1, last = 0;
                nx.it = it;
                if (r_alive > survival(&cx, &cur)) break;                 /* death: the remaining rows stay NaN */
                for (ist1 = 0; ist1 < nst; ist1++) {                       /* inverse-CDF sampling of the next state */
                    double pr;
#if EGDST_NCONT > 0
                    {   /* only cells at the first grid point of every continuous state (egdst_simulator.c:268-271) */
                        int off = 0;
                        for (int q = 0; q < EGDST_NCONT; q++) {
                            const int j0 = egdst_contvar[q];
                            if ((ist1 / (int)cx.stm[cx.nnst + j0]) % (int)cx.stm[j0] != 0) off = 1;
                        }
                        if (off) continue;
                    }
#endif
                    nx.ist = ist1;
                    for (int i = 0; i < cx.nnst; i++) nx.st[i] = cx.states[i * nst + ist1];
                    trpr_cont(&cx, &cur, &nx);                              /* continuous states move exactly */
                    if (!feasible(&cx, &nx)) continue;
                    last = ist1;
                    if (cx.optim_TRPRnoSH != 1) {                          /* probabilities may depend on the shock */
                        mu = mu_param(&cx, &cur, &nx); sigma = sigma_param(&cx, &cur, &nx);
                        nx.shock = sigma <= 0 ? port_expectation(&cx, &cur, &nx) : port_cdfinv(r_shock, mu, sigma);
                    }
                    pr = trpr(&cx, &cur, &nx, 0);
                    r_state -= pr;
                    if (r_state <= 0) break;
                }
                if (ist1 >= nst) {                                          /* ran off the end: keep the last feasible state */
                    nx.ist = last;
                    for (int i = 0; i < cx.nnst; i++) nx.st[i] = cx.states[i * nst + last];
                    trpr_cont(&cx, &cur, &nx);
                }
                if (cx.optim_TRPRnoSH == 1) {
                    mu = mu_param(&cx, &cur, &nx); sigma = sigma_param(&cx, &cur, &nx);
                    nx.shock = sigma <= 0 ? port_expectation(&cx, &cur, &nx) : port_cdfinv(r_shock, mu, sigma);
                }
                nx.cash = cashinhand(&cx, &cur, &nx);
                eqs_sim(&cx, &cur, &nx, eqs);
                cur = nx;
            }
            {
                double *out = sims + ((size_t)isim * nt + it) * nso;
                int j = 11;
#if EGDST_NCONT > 0
                /* egdst_simulator.c:309-365.  Corner q of the 2^NCONT cells around the exact state takes grid point
                 * j1 (bit clear) or j1+1 (bit set) of each continuous state; non-positive weights are skipped; the
                 * recorded cell/decision is the corner where the running weight passes one half.  Column 3 (value)
                 * is left unassigned by the reference on this branch; the mix of the corner values is written here.
                 * The grid part of an initial cell index is removed first (the reference adds the corner offset on
                 * top of it, :313, which only works for initial cells at the first grid point). */
                {
                    int j1[EGDST_NCONT], stride[EGDST_NCONT], base = cur.ist, ist_pick = -1, id_pick = 0, bad = 0;
                    double wlo[EGDST_NCONT], whi[EGDST_NCONT], wc = 0, wvf = 0, half = .5;
                    for (int q = 0; q < EGDST_NCONT; q++) {
                        const int j0 = egdst_contvar[q], n = (int)cx.stm[j0];
                        const double *g = egdst_contgrid(q), x = cur.st[j0];
                        stride[q] = (int)cx.stm[cx.nnst + j0];
                        base -= ((cur.ist / stride[q]) % n) * stride[q];
                        j1[q] = port_bracket(x, g, n, 0);
                        wlo[q] = (g[j1[q] + 1] - x) / (g[j1[q] + 1] - g[j1[q]]);
                        whi[q] = (x - g[j1[q]]) / (g[j1[q] + 1] - g[j1[q]]);
                    }
                    for (int q = 0; q < (1 << EGDST_NCONT); q++) {
                        double wt = 1, cc, vv;
                        PeriodVars pv = cur;
                        pv.ist = base;
                        for (int b = 0; b < EGDST_NCONT; b++) {
                            wt *= ((q >> b) & 1) ? whi[b] : wlo[b];
                            pv.ist += stride[b] * (j1[b] + ((q >> b) & 1));
                        }
                        if (!(wt > 0)) continue;
                        if (cells[(size_t)it * nst + pv.ist].n < 2) { bad = 1; break; }
                        port_policy(&cx, cells + (size_t)it * nst + pv.ist, &pv, &cc, &vv);
                        wc += cc * wt; wvf += vv * wt;
                        half -= wt;
                        if (ist_pick < 0 && half < 0) { ist_pick = pv.ist; id_pick = pv.id; }
                    }
                    if (bad || ist_pick < 0) break;
                    c = MIN(wc, cur.cash - cx.a0);
                    cur.savings = cur.cash - c;
                    vf = wvf;
                    cur.ist = ist_pick; cur.id = id_pick;
                    for (int i = 0; i < cx.nnd; i++) cur.dc[i] = cx.decisions[i * cx.nd + cur.id];
                }
#else
                const port_cell *pc = cells + (size_t)it * nst + cur.ist;
                if (pc->n < 2) break;
                port_policy(&cx, pc, &cur, &c, &vf);
#endif
                out[0] = cur.cash; out[1] = c; out[2] = cur.savings; out[3] = vf; out[4] = cur.id; out[5] = cur.ist;
                out[6] = mu; out[7] = sigma; out[8] = cur.shock; out[9] = utility(&cx, &cur, c); out[10] = discount(&cx, &cur);
                for (int i = 0; i < cx.nnst; i++) out[j++] = cur.st[i];
                for (int i = 0; i < cx.nnd; i++) out[j++] = cur.dc[i];
                for (int i = 0; i < d->neq; i++) out[j++] = eqs[i];
            }
        }
    }
    free(cells);
    return skipped;
}
