// egdst_solver.cuh -- backward-induction kernels for one period (all states, all decisions, all
// parameter vectors of a batch at once).
//
// Restates the reference's egmbellman (egdst_solver.c:370-752) in parallel form:
//   k_terminal   terminal-period closed-form grid                     egdst_solver.c:452-475
//   k_seed       adraw stage 0/1 + ZEROCONSUMPTION feedback            egdst_solver.c:955-1099, 583-627
//   k_egm        expectation over (ist1, iy) + Euler inversion         egdst_solver.c:490-665
//   k_compact    adraw stop rule, drop rules, fold detection           egdst_solver.c:1100-1152, 640-664, 819
// The upper envelopes are in egdst_envelope.cuh.
//
// Parallel decomposition (hard part 1): the A-grid is sequential in the reference only
// through (i) the stage-0 bisection, whose candidates mmax, (mmax+a0)/2, ... are known in advance and
// are evaluated concurrently, (ii) the a0 point and its re-sends (serial, a handful of evaluations),
// and (iii) the stop rule, which is applied after all N-1 closed-form grid points were evaluated.
#pragma once

#include "egdst_tables.cuh"

// next-period tables of one state ist1
struct EgdstNext {
    const double *M, *C, *V;  // rows 0..n1 (row 0 = a0 row)
    const double *th, *dd;
    int n1, nth;              // n1 = egdims (rows excluding the a0 row)
    double evf;
    int cell;
    const EgdstRow *ivl;      // row table of the cell (egdst_tables.cuh)
    double M1, Mlast, Clast;  // M[1], M[n1], C[n1]
};

EGDST_DEV EgdstNext egdst_next_tables(const EgdstDev &P, int ivec, int it1, int ist1) {
    int cell = egdst_cell(P, ivec, it1, ist1);
    EgdstNext t;
    t.M = egdst_colM(P, cell);
    t.C = egdst_colC(P, cell);
    t.V = egdst_colV(P, cell);
    t.n1 = P.mlen[cell] - 1;
    t.th = P.thTH + (size_t)cell * P.cx.nthrhmax;
    t.dd = P.thD + (size_t)cell * P.cx.nthrhmax;
    t.nth = P.thlen[cell];
    t.evf = P.evf[cell];
    t.cell = cell;
    t.ivl = egdst_cell_rows(P, cell);
    t.M1 = t.M[1]; t.Mlast = t.M[t.n1]; t.Clast = t.C[t.n1];
    return t;
}

// accumulator of one Euler evaluation (possibly a partial one: a strided subset of the nodes)
struct EgdstAcc {
    double rhs, evf, checksum;
    int badq;        // smallest flattened node index ist1*ny+iy that aborted the evaluation, or INT_MAX
    int badtype;     // EGDST_PT_C1NEG or EGDST_PT_EVFINF
    double badcash, badshock;
};
#define EGDST_NOBAD 0x7fffffff

// Evaluate nodes q = part, part+nparts, ... of the expectation at end-of-period savings A
// (egdst_solver.c:494-574).  keep==0 skips the value function (adraw seed phase).
// Quadrature shocks and node probabilities of one (it, ist, id) for every (ist1, iy): shk/shp [nst*ny].
// Valid when the model image says they cannot depend on savings (EGDST_SHOCK_INDEP_A, codegen): computed once
// per CTA instead of once per node (one exp per node saved).  shp == 0 marks nodes the reference skips
// (infeasible ist1, zero transition probability, iy >= niy).
EGDST_DEV void egdst_fill_shocktab(const egdst_ctx *cx, const EgdstDev &P, const PeriodVars *curr, double *shk, double *shp, int tid, int nthreads) {
    const int ny = cx->ny, nst = cx->nst;
    for (int q = tid; q < nst * ny; q += nthreads) {
        const int ist1 = q / ny, iy = q - ist1 * ny;
        PeriodVars next;
        next.it = curr->it + 1; next.savings = 0.0; next.id = 0; next.cash = 0.0; next.shock = 0.0; next.ist = ist1;
        double p = 0.0, sh = 0.0;
        if (feasible(cx, &next) == 1) {
            const int niy = (sigma_param(cx, curr, &next) <= 0 || ny == 1) ? 1 : ny;
            if (iy < niy) {
                sh = (niy == 1) ? egdst_expectation(cx, curr, &next) : egdst_rescale(cx, curr, &next, P.qz[iy]);
                next.shock = sh;
                p = trpr(cx, curr, &next, 1);
                if (niy > 1) p *= P.qw[iy];
            }
        }
        shk[q] = sh; shp[q] = p;
    }
}

EGDST_DEV void egdst_eval_nodes(const egdst_ctx *cx, const EgdstDev &P, int ivec, const PeriodVars *curr, double A, int keep,
                                int part, int nparts, EgdstAcc &acc, const double *shk = 0, const double *shp = 0) {
    const int ny = cx->ny, nst = cx->nst;
    acc.rhs = 0.0; acc.evf = 0.0; acc.checksum = 0.0; acc.badq = EGDST_NOBAD; acc.badtype = 0; acc.badcash = 0.0; acc.badshock = 0.0;
    PeriodVars next;
    next.it = curr->it + 1;
    next.savings = A;
    next.id = 0;
    next.cash = 0.0;
    next.shock = 0.0;
    for (int ist1 = 0; ist1 < nst; ist1++) {
        next.ist = ist1;
        if (feasible(cx, &next) != 1) continue;
        double pr1pre = 0.0;
        int niy = ny;
        if (!shk) {
            if (cx->optim_TRPRnoSH == 1) {
                pr1pre = trpr(cx, curr, &next, 1);
                if (pr1pre == 0.0) continue;
            }
            niy = (sigma_param(cx, curr, &next) <= 0 || ny == 1) ? 1 : ny;
        }
        const EgdstNext t = egdst_next_tables(P, ivec, next.it, ist1);
        for (int iy = part; iy < niy; iy += nparts) {
            double pr1;
            if (shk) {
                pr1 = shp[ist1 * ny + iy];
                next.shock = shk[ist1 * ny + iy];
            } else if (niy == 1) {
                next.shock = egdst_expectation(cx, curr, &next);
                pr1 = (cx->optim_TRPRnoSH != 1) ? trpr(cx, curr, &next, 1) : pr1pre;
            } else {
                next.shock = egdst_rescale(cx, curr, &next, P.qz[iy]);
                pr1 = (cx->optim_TRPRnoSH != 1) ? trpr(cx, curr, &next, 1) : pr1pre;
                pr1 *= P.qw[iy];
            }
            if (pr1 == 0.0) continue;
            const int q = ist1 * ny + iy;
            if (q > acc.badq) break;  // the reference would have stopped before this node
            acc.checksum += pr1;
            next.cash = cashinhand(cx, curr, &next);
            // one bracket lookup serves consumption (rows 0..n1) and value (rows 1..n1): the second bracket of the
            // reference is max(i,1) of the first (same strictly increasing grid); rows i, i+1 come as one record
            const bool tab = egdst_cell_has_tab(P, t.n1 + 1);
            EgdstInterval iv;
            int i;
            if (tab) i = egdst_lookup_tab(P, t.cell, t.ivl, next.cash, t.n1 + 1, iv);
            else {
                i = egdst_bracket(next.cash, t.M, t.n1 + 1, 0); iv.g0 = t.M[i]; iv.g1 = t.M[i + 1]; iv.c0 = t.C[i]; iv.c1 = t.C[i + 1]; iv.v0 = t.V[i]; iv.v1 = t.V[i + 1];
                iv.y = egdst_div_safe(iv.g1 - iv.g0) ? 1.0 / (iv.g1 - iv.g0) : 0.0;
            }
            // the reference's quotients are kept bit for bit (two divisions per interpolation, egdst_lib.c:175): next to its
            // instability boundary a plain reciprocal-multiply variant drifted 1e-2 away in C, so the
            // shared-reciprocal form below is the exactly rounded one (egdst_div_by)
            double w = iv.g1 - iv.g0;
            double y = iv.y;  // shared correctly rounded reciprocal RN(1/w), from the record (0: degenerate interval, plain divisions)
            double c1 = y != 0.0 ? egdst_lerp_y(next.cash, iv.g0, iv.g1, iv.c0, iv.c1, w, y) : egdst_lerp(next.cash, iv.g0, iv.g1, iv.c0, iv.c1);
            if (next.cash > t.Mlast) c1 = MAX(c1, t.Clast);  // constant extrapolation guard (egdst_solver.c:554)
            if (c1 <= 0) {
                acc.badq = q; acc.badtype = EGDST_PT_C1NEG; acc.badcash = next.cash; acc.badshock = next.shock;
                break;
            }
            if (cx->optim_MUnoD != 1 || (cx->optim_UnoD != 1 && keep == 1 && next.cash < t.M1))
                next.id = egdst_optimd(next.cash, t.th, t.dd, t.nth);
            else
                next.id = 0;
            acc.rhs += pr1 * utility_marginal(cx, &next, c1) * cashinhand_marginal(cx, curr, &next);
            if (keep == 1) {
                double v1;
                if (next.cash < t.M1 && t.evf > -EGDST_INF) {
                    v1 = utility(cx, &next, next.cash - cx->a0) + discount(cx, &next) * t.evf;  // egdst_solver.c:763
                } else {
                    if (i < 1) {  // value table starts at row 1 (row 0 of V is evf(a0), not a value)
                        if (tab) iv = egdst_load_interval(t.ivl + 1);
                        else { iv.g0 = t.M[1]; iv.g1 = t.M[2]; iv.v0 = t.V[1]; iv.v1 = t.V[2]; iv.y = egdst_div_safe(iv.g1 - iv.g0) ? 1.0 / (iv.g1 - iv.g0) : 0.0; }
                        w = iv.g1 - iv.g0;
                        y = iv.y;
                    }
                    v1 = egdst_linter_extrap_iv(cx, &next, next.cash, iv.g0, iv.g1, iv.v0, iv.v1, t.M1, t.Mlast, w, y);
                }
                const double term = pr1 * v1;
                acc.evf += term;
                if (term == -EGDST_INF) {
                    acc.badq = q; acc.badtype = EGDST_PT_EVFINF; acc.badcash = next.cash; acc.badshock = next.shock;
                    break;
                }
            }
        }
    }
}

// combine partial accumulators across a warp; every lane ends with the full result
EGDST_DEV void egdst_warp_combine(EgdstAcc &a) {
    const int minq = __shfl_sync(EGDST_FULL, egdst_warp_min(a.badq), 0);
    const unsigned owner = __ballot_sync(EGDST_FULL, a.badq == minq && minq != EGDST_NOBAD);
    a.rhs = __shfl_sync(EGDST_FULL, egdst_warp_sum(a.rhs), 0);
    a.evf = __shfl_sync(EGDST_FULL, egdst_warp_sum(a.evf), 0);
    a.checksum = __shfl_sync(EGDST_FULL, egdst_warp_sum(a.checksum), 0);
    if (minq != EGDST_NOBAD) {
        const int src = __ffs(owner) - 1;
        a.badtype = __shfl_sync(EGDST_FULL, a.badtype, src);
        a.badcash = __shfl_sync(EGDST_FULL, a.badcash, src);
        a.badshock = __shfl_sync(EGDST_FULL, a.badshock, src);
    }
    a.badq = minq;
}

// ---------------------------------------------------------------------------------------------
// terminal period: M_i = trinv(m1 + i (m2-m1)/(N-1)), C = M, V = u(C); evfa0 = -inf   (END2, A0T = 0)
// grid (ceil(N/B), nst*nd, nvec)
// ---------------------------------------------------------------------------------------------
__global__ void egdst_k_terminal(EgdstDev P, int it) {
    const int ivec = blockIdx.z, ist = blockIdx.y / P.cx.nd, id = blockIdx.y % P.cx.nd;
    egdst_ctx cx; egdst_load_ctx(P, ivec, cx);
    const int sd = egdst_sd(P, ivec, ist, id);
    PeriodVars curr; curr.it = it; curr.ist = ist; curr.id = id; curr.cash = 0; curr.savings = 0; curr.shock = 0;
    const int act = (feasible(&cx, &curr) == 1) && (inchoiceset(&cx, &curr) == 1);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        P.active[sd] = act;
        P.evfa0[sd] = -EGDST_INF;
        P.ptN[sd] = act ? P.N : 0;
        P.nfold[sd] = 0;
        if (act) atomicAdd(P.units + ivec, (unsigned long long)P.N);
    }
    if (!act || i >= P.N) return;
    const double m1 = tr(&cx, &curr, cx.zeroconsumption - 0.0), m2 = tr(&cx, &curr, cx.mmax - 0.0);
    const double m = trinv(&cx, &curr, m1 + i * (m2 - m1) / (P.N - 1)) + 0.0;
    const double c = m - 0.0;
    P.ptX[(size_t)sd * P.gcap + i] = m;
    P.ptC[(size_t)sd * P.gcap + i] = c;
    P.ptV[(size_t)sd * P.gcap + i] = utility(&cx, &curr, c);
}

// ---------------------------------------------------------------------------------------------
// seed: one CTA per (ist,id,ivec).  Warps evaluate the stage-0 candidates concurrently; the whole CTA
// evaluates A=a0 and any re-sent point; thread 0 runs the adraw state machine.
// ---------------------------------------------------------------------------------------------
struct EgdstSeedShared {
    double candA[EGDST_MAXCAND], candM[EGDST_MAXCAND];
    int candBad[EGDST_MAXCAND];
    double wr[32], we[32], wc[32], wcash[32], wshock[32];
    int wq[32], wt[32];
    double A, rhs, evf, checksum, badcash, badshock;
    int ncand, badq, badtype, go, nresend;
};

EGDST_DEV void egdst_block_eval(const egdst_ctx *cx, const EgdstDev &P, int ivec, const PeriodVars *curr, double A, int keep,
                                EgdstSeedShared &S, const double *shk, const double *shp) {
    EgdstAcc a;
    egdst_eval_nodes(cx, P, ivec, curr, A, keep, threadIdx.x, blockDim.x, a, shk, shp);
    egdst_warp_combine(a);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) { S.wr[w] = a.rhs; S.we[w] = a.evf; S.wc[w] = a.checksum; S.wq[w] = a.badq; S.wt[w] = a.badtype; S.wcash[w] = a.badcash; S.wshock[w] = a.badshock; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0, e = 0, c = 0; int bq = EGDST_NOBAD, bw = 0;
        for (int k = 0; k < nw; k++) { r += S.wr[k]; e += S.we[k]; c += S.wc[k]; if (S.wq[k] < bq) { bq = S.wq[k]; bw = k; } }
        S.rhs = r; S.evf = e; S.checksum = c; S.badq = bq; S.badtype = S.wt[bw]; S.badcash = S.wcash[bw]; S.badshock = S.wshock[bw];
    }
    __syncthreads();
}

__global__ void egdst_k_seed(EgdstDev P, int it, int useTab) {
    __shared__ EgdstSeedShared S;
    EGDST_DYN_SMEM(double, shsm);
    const int ivec = blockIdx.z, ist = blockIdx.y, id = blockIdx.x;
    egdst_ctx cx; egdst_load_ctx(P, ivec, cx);
    const int sd = egdst_sd(P, ivec, ist, id);
    PeriodVars curr; curr.it = it; curr.ist = ist; curr.id = id; curr.cash = 0; curr.savings = 0; curr.shock = 0;
    const int act = (feasible(&cx, &curr) == 1) && (inchoiceset(&cx, &curr) == 1);
    const int N = P.N;
    double *rawM = P.rawM + (size_t)sd * N, *rawC = P.rawC + (size_t)sd * N, *rawV = P.rawV + (size_t)sd * N, *rawStop = P.rawStop + (size_t)sd * N;
    int *rawFlag = P.rawFlag + (size_t)sd * N;
    double *seed = P.seed + (size_t)sd * 8;
    if (threadIdx.x == 0) {
        P.active[sd] = act;
        P.evfa0[sd] = 0.0;
        P.ptN[sd] = 0;
        P.nfold[sd] = 0;
        rawFlag[0] = EGDST_PT_NONE;
        rawStop[0] = EGDST_INF;  // "stop": no grid points unless the seed succeeds
        // stage-0 candidates (egdst_solver.c:979-1027): mmax, then halfway to a0 until A-a0<TOLERANCE
        int n = 0; double A = cx.mmax;
        while (n < EGDST_MAXCAND) { S.candA[n++] = A; if (A - cx.a0 < cx.tolerance) break; A = (A + cx.a0) / 2; }
        S.ncand = n;
    }
    __syncthreads();
    if (!act) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const double beta = discount(&cx, &curr);
    const double *shk = 0, *shp = 0;
    if (useTab) {  // shocks and node probabilities of this (it, ist, id), once per CTA
        egdst_fill_shocktab(&cx, P, &curr, shsm, shsm + cx.nst * cx.ny, threadIdx.x, blockDim.x);
        shk = shsm; shp = shsm + cx.nst * cx.ny;
        __syncthreads();
    }
    // stage 0 in waves of one candidate per warp: the base point is almost always among the first few
    // candidates (mmax, (mmax+a0)/2, ...), so later waves rarely run
    double baseA = 0, baseM = 0;
    if (threadIdx.x == 0) S.go = 0;
    __syncthreads();
    for (int k0 = 0; k0 < S.ncand; k0 += nw) {
        const int kk = k0 + w;
        if (kk < S.ncand) {
            EgdstAcc a;
            egdst_eval_nodes(&cx, P, ivec, &curr, S.candA[kk], 0, lane, 32, a, shk, shp);
            egdst_warp_combine(a);
            if (lane == 0) {
                int bad = 0;
                if (a.badq != EGDST_NOBAD) bad = EGDST_PT_C1NEG;
                else if (fabs(a.checksum - 1) > cx.tolerance) bad = EGDST_PT_CHECKSUM;
                S.candBad[kk] = bad;
                S.candM[kk] = S.candA[kk] + utility_marginal_inverse(&cx, &curr, beta * a.rhs);
            }
        }
        __syncthreads();
        // thread 0: first candidate with M<=mmax is the base point (sequential semantics of adraw stage 0)
        if (threadIdx.x == 0) {
            const int kend = k0 + nw < S.ncand ? k0 + nw : S.ncand;
            for (int k = k0; k < kend; k++) {
                if (S.candBad[k] == EGDST_PT_C1NEG) { egdst_fail(P, ivec, EGDST_ERR_NOSAVINGS, it, ist, id); S.go = -1; break; }
                if (S.candBad[k] == EGDST_PT_CHECKSUM) { egdst_fail(P, ivec, EGDST_ERR_CHECKSUM, it, ist, id); S.go = -1; break; }
                if (S.candM[k] <= cx.mmax) { baseA = S.candA[k]; baseM = S.candM[k]; S.go = 1; break; }
                if (S.candA[k] - cx.a0 < cx.tolerance) { egdst_fail(P, ivec, EGDST_ERR_ADRAW_INIT, it, ist, id); S.go = -1; break; }
            }
            if (kend == S.ncand && S.go == 0) { egdst_fail(P, ivec, EGDST_ERR_ADRAW_INIT, it, ist, id); S.go = -1; }
        }
        __syncthreads();
        if (S.go != 0) break;
    }
    if (threadIdx.x == 0) {
        S.A = cx.a0;
        S.nresend = S.ncand;  // used as the adraw call counter (loop guard, egdst_solver.c:963)
    }
    __syncthreads();
    if (S.go != 1) return;
    // stage 1 (+ re-sends): serial in A, parallel over nodes
    double lim1 = 0, lim2 = 0, lim3 = 0, lim2p = 0, lim3p = 0, k3 = 0, lastA = cx.a0, aM = 0, evfa0 = 0.0;
    int stored = 0;
    while (true) {
        egdst_block_eval(&cx, P, ivec, &curr, S.A, 1, S, shk, shp);
        if (threadIdx.x == 0) {
            lastA = S.A;
            int resend = 0, fatal = 0;
            if (S.badq == EGDST_NOBAD && fabs(S.checksum - 1) > cx.tolerance) {
                egdst_fail(P, ivec, EGDST_ERR_CHECKSUM, it, ist, id); fatal = 1;
            } else if (S.badq != EGDST_NOBAD) {
                evfa0 = -EGDST_INF;
                aM = S.badcash;
                if (S.badtype == EGDST_PT_C1NEG) {
                    aM = cx.a0 - 1;
                    PeriodVars next; next.it = it + 1; next.ist = S.badq / cx.ny; next.id = 0; next.savings = lastA; next.shock = S.badshock; next.cash = S.badcash;
                    const EgdstNext t = egdst_next_tables(P, ivec, it + 1, next.ist);
                    const double target = (t.evf > -EGDST_INF) ? cx.a0 : t.M[1];
                    int fail = 0;
                    lastA = egdst_cashinhandinverse(&cx, &curr, next, target, &fail) + cx.zeroconsumption;
                    if (fail) { egdst_fail(P, ivec, EGDST_ERR_CASHINVERSE, it, ist, id); fatal = 1; }
                }
            } else {
                aM = lastA + utility_marginal_inverse(&cx, &curr, beta * S.rhs);
                if (isfinite(aM)) {
                    if (fabs(lastA - cx.a0) < cx.tolerance && evfa0 > -EGDST_INF) evfa0 = S.evf;
                    rawM[0] = aM; rawC[0] = aM - lastA; rawV[0] = utility(&cx, &curr, aM - lastA) + beta * S.evf;
                    rawFlag[0] = EGDST_PT_OK; stored = 1;
                }
            }
            if (!fatal) {
                // adraw, ngenerated==1 branch (egdst_solver.c:1032-1099)
                double aa = (aM - baseM) / (lastA - baseA), bb = baseM - aa * baseA;
                if (k3 == 0) {
                    lim2p = MIN(cx.mmax, (cx.mmax - bb) / aa);
                    lim3p = -bb / aa;
                    if (cx.a0 < 0 && cx.a0 < lim3p) k3 = MAX(floor(N * (lim3p - cx.a0) / (lim2p - cx.a0)), 2.0);
                    else { lim3p = cx.a0; k3 = 1.0; }
                    lim1 = tr(&cx, &curr, lim3p - cx.a0); lim2 = tr(&cx, &curr, lim2p - lim3p); lim3 = tr(&cx, &curr, 0);
                }
                if (aM <= cx.a0 - 1 + cx.tolerance) {
                    aa = (cx.a0 - baseM) / (cx.a0 - baseA); bb = baseM - aa * baseA;
                    lim2p = MIN(cx.mmax, (cx.mmax - bb) / aa);
                    lim3p = lastA - cx.zeroconsumption;
                    k3 = 1.0;
                    lim1 = tr(&cx, &curr, lim3p - cx.a0); lim2 = tr(&cx, &curr, lim2p - lim3p); lim3 = tr(&cx, &curr, 0);
                    resend = 1;
                    if (++S.nresend + 1 >= cx.ngridmax) { egdst_fail(P, ivec, EGDST_ERR_ADRAW_LOOP, it, ist, id); resend = 0; fatal = 1; }
                }
            }
            S.A = lastA;
            S.go = fatal ? -1 : resend;
        }
        __syncthreads();
        if (S.go != 1) break;
    }
    if (threadIdx.x == 0) {
        P.evfa0[sd] = evfa0;
        if (S.go == 0) {
            seed[0] = lim1; seed[1] = lim2; seed[2] = lim3; seed[3] = lim3p; seed[4] = k3; seed[5] = lastA;
            rawStop[0] = aM;
            if (!stored) rawFlag[0] = EGDST_PT_EVFINF;
        }
    }
}

// closed-form A-grid after the seed (egdst_solver.c:1104-1136); n = 1..N-1
EGDST_DEV double egdst_agrid_target(const egdst_ctx *cx, const PeriodVars *curr, const double *seed, int n, int N) {
    const double lim1 = seed[0], lim2 = seed[1], lim3 = seed[2], lim3p = seed[3], k3 = seed[4];
    if (n < (int)k3 - 1) return -trinv(cx, curr, lim3 + (k3 - 1 - n) * (lim1 - lim3) / (k3 - 1)) + lim3p;
    return trinv(cx, curr, lim3 + (n - k3 + 1) * (lim2 - lim3) / (N - k3)) + lim3p;
}
EGDST_DEV double egdst_agrid(const egdst_ctx *cx, const PeriodVars *curr, const double *seed, int n, int N) {
    const double t = egdst_agrid_target(cx, curr, seed, n, N);
    const double prev = (n == 1) ? seed[5] : egdst_agrid_target(cx, curr, seed, n - 1, N);
    return (t - prev < 0) ? prev + 1e-5 : t;  // astep<0 rule (egdst_solver.c:1120-1133)
}

// ---------------------------------------------------------------------------------------------
// EGM step for grid points n=1..N-1.  blockDim = (32, parts <= SPLIT): lanes are 32 consecutive A points (their
// next-period cash values are neighbours, so the table searches of a warp stay coherent), the `parts`
// warps of a CTA share the quadrature nodes of the same 32 points and combine through shared memory.
// The host picks `parts` so that the nodes divide evenly (10 nodes: 5 warps of 2, not 8 warps of 2 and 1).
// grid (ceil((N-1)/32), nst*nd, nvec)
// ---------------------------------------------------------------------------------------------
#ifndef EGDST_EGM_SPLIT
#ifdef EGDST_HOSTEMU
#define EGDST_EGM_SPLIT 2
#else
#define EGDST_EGM_SPLIT 8
#endif
#endif

#ifndef EGDST_EGM_MINB
#define EGDST_EGM_MINB 3
#endif
__global__ void __launch_bounds__(32 * EGDST_EGM_SPLIT, EGDST_EGM_MINB) egdst_k_egm(EgdstDev P, int it, int useTab) {
    EGDST_DYN_SMEM(double, shsm);
    __shared__ double s_rhs[EGDST_EGM_SPLIT][32], s_evf[EGDST_EGM_SPLIT][32], s_chk[EGDST_EGM_SPLIT][32], s_cash[EGDST_EGM_SPLIT][32];
    __shared__ int s_q[EGDST_EGM_SPLIT][32], s_t[EGDST_EGM_SPLIT][32];
    const int ivec = blockIdx.z, ist = blockIdx.y / P.cx.nd, id = blockIdx.y % P.cx.nd;
    const int sd = egdst_sd(P, ivec, ist, id);
    const int lane = threadIdx.x, part = threadIdx.y, nparts = blockDim.y;
    const int N = P.N;
    if (!P.active[sd] || P.rawFlag[(size_t)sd * N] == EGDST_PT_NONE) return;  // uniform per CTA
    egdst_ctx cx; egdst_load_ctx(P, ivec, cx);
    PeriodVars curr; curr.it = it; curr.ist = ist; curr.id = id; curr.cash = 0; curr.savings = 0; curr.shock = 0;
    const double *seed = P.seed + (size_t)sd * 8;
    const double *shk = 0, *shp = 0;
    if (useTab) {  // shocks and node probabilities of this (it, ist, id), once per CTA
        egdst_fill_shocktab(&cx, P, &curr, shsm, shsm + cx.nst * cx.ny, threadIdx.y * 32 + threadIdx.x, 32 * nparts);
        shk = shsm; shp = shsm + cx.nst * cx.ny;
        __syncthreads();
    }
    // gridDim.x CTAs share the ceil((N-1)/32) blocks of 32 points of this (ist,id): one block each for a single model,
    // all of them in turn for the short grids of a batched sweep (context and shock table set up once)
    for (int xb = blockIdx.x; xb * 32 < N - 1; xb += gridDim.x) {
        const int n = 1 + xb * 32 + lane;
        double A = 0.0;
        EgdstAcc a; a.rhs = 0; a.evf = 0; a.checksum = 0; a.badq = EGDST_NOBAD; a.badtype = 0; a.badcash = 0; a.badshock = 0;
        if (n < N) {
            A = egdst_agrid(&cx, &curr, seed, n, N);
            egdst_eval_nodes(&cx, P, ivec, &curr, A, 1, part, nparts, a, shk, shp);
        }
        s_rhs[part][lane] = a.rhs; s_evf[part][lane] = a.evf; s_chk[part][lane] = a.checksum;
        s_q[part][lane] = a.badq; s_t[part][lane] = a.badtype; s_cash[part][lane] = a.badcash;
        __syncthreads();
        if (part == 0 && n < N) {
            double rhs = 0, evf = 0, chk = 0, badcash = 0; int bq = EGDST_NOBAD, bt = 0;
            for (int k = 0; k < nparts; k++) {
                rhs += s_rhs[k][lane]; evf += s_evf[k][lane]; chk += s_chk[k][lane];
                if (s_q[k][lane] < bq) { bq = s_q[k][lane]; bt = s_t[k][lane]; badcash = s_cash[k][lane]; }
            }
            const size_t o = (size_t)sd * N + n;
            if (bq != EGDST_NOBAD) {
                P.rawFlag[o] = bt;
                P.rawStop[o] = (bt == EGDST_PT_C1NEG) ? cx.a0 - 1 : badcash;
            } else if (fabs(chk - 1) > cx.tolerance) {
                P.rawFlag[o] = EGDST_PT_CHECKSUM; P.rawStop[o] = EGDST_INF;
            } else {
                const double beta = discount(&cx, &curr);
                const double M = A + utility_marginal_inverse(&cx, &curr, beta * rhs);
                P.rawStop[o] = M;
                if (!isfinite(M)) P.rawFlag[o] = EGDST_PT_NONFINITE;
                else {
                    const double c = M - A;
                    P.rawM[o] = M; P.rawC[o] = c; P.rawV[o] = utility(&cx, &curr, c) + beta * evf;
                    P.rawFlag[o] = EGDST_PT_OK;
                }
            }
        }
        if (xb + (int)gridDim.x < (N - 1 + 31) / 32) __syncthreads();  // the partial sums are reused by the next block
    }
}

// ---------------------------------------------------------------------------------------------
// stop rule + compaction + fold detection: one CTA per (ist,id,ivec).
// The reference generates point n only while every earlier returned M was < mmax
// (egdst_solver.c:1100); stored points are those with a finite M and no abort (:640-664).
// Folds (M or V decreasing, :819) split the list into runs for the secondary envelope.
// ---------------------------------------------------------------------------------------------
#define EGDST_CMP_IPT 8
#ifdef EGDST_HOSTEMU
#define EGDST_CMP_THREADS 64
#else
#define EGDST_CMP_THREADS 256
#endif
// IPT raw points per thread (8; 2 for a single large model: more CTAs share the latency of the passes).
// grid (chC, nst*nd, nvec), blockDim.x = P.cmpW <= EGDST_CMP_THREADS: the CTAs of one (ist,id) list are chained by a decoupled look-back scan whose state
// carries (points kept so far, "stop rule fired").  Folds inside a CTA's own output are appended to an unordered
// list; the last CTA to finish adds the folds on chunk boundaries, orders the list and publishes the counts.
template <int IPT>
__global__ void __launch_bounds__(EGDST_CMP_THREADS) egdst_k_compact(EgdstDev P, int it) {
    __shared__ int sh[40];
    __shared__ int s_chunk, s_last;
    __shared__ unsigned long long s_excl;
    const int ivec = blockIdx.z, ist = blockIdx.y / P.cx.nd, id = blockIdx.y % P.cx.nd;
    const int sd = egdst_sd(P, ivec, ist, id);
    if (!P.active[sd]) return;
    const int N = P.N;
    const double *rawM = P.rawM + (size_t)sd * N, *rawC = P.rawC + (size_t)sd * N, *rawV = P.rawV + (size_t)sd * N, *rawStop = P.rawStop + (size_t)sd * N;
    const int *rawFlag = P.rawFlag + (size_t)sd * N;
    double *X = P.ptX + (size_t)sd * P.gcap, *Cc = P.ptC + (size_t)sd * P.gcap, *V = P.ptV + (size_t)sd * P.gcap;
    int *runStart = P.runStart + (size_t)sd * (P.gcap + 1);
    int *foldList = P.foldList + (size_t)sd * (P.gcap + 1);
    volatile unsigned long long *st = P.scanC + (size_t)sd * P.chC;
    if (rawFlag[0] == EGDST_PT_NONE) { if (blockIdx.x == 0 && threadIdx.x == 0) { P.ptN[sd] = 0; P.nfold[sd] = 0; } return; }
    const int chunkw = blockDim.x * IPT;  // raw points per CTA (P.cmpW threads: narrow CTAs for short grids)
    const int nch = (N + chunkw - 1) / chunkw;      // chunks that hold raw points == gridDim.x
    if (threadIdx.x == 0) s_chunk = atomicAdd(P.tickC + 2 * sd, 1);
    __syncthreads();
    const int chunk = s_chunk;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const unsigned ltmask = (1u << lane) - 1u;
    int err = 0;
    if (chunk < nch) {
        const int wbase = chunk * chunkw + w * (32 * IPT);
        // the stop rule inside this chunk: first n whose returned M fails "M<mmax" (that point itself is kept)
        int flag[IPT], mystop = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            const int n = wbase + j * 32 + lane;
            flag[j] = EGDST_PT_NONE;
            if (n < N) {
                flag[j] = rawFlag[n];
                if (!(rawStop[n] < P.cx.mmax) && n < mystop) mystop = n;
            }
        }
        const int ls = egdst_block_min(mystop, sh);
        unsigned bal[IPT];
        int wtotal = 0, late = 0x7fffffff, badsum = 0;
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            const int n = wbase + j * 32 + lane;
            const bool in = n < N && n <= ls;
            if (in && flag[j] == EGDST_PT_C1NEG && n > 0 && n < late) late = n;
            if (in && flag[j] == EGDST_PT_CHECKSUM) badsum = 1;
            bal[j] = __ballot_sync(EGDST_FULL, in && flag[j] == EGDST_PT_OK);
            wtotal += __popc(bal[j]);
        }
        int total;
        int woff = egdst_block_excl_scan(lane == 0 ? wtotal : 0, sh, &total);
        woff = __shfl_sync(EGDST_FULL, woff, 0);
        if (w == 0) {
            const unsigned long long e = egdst_lookback<1>(st, chunk, egdst_scan_pack(total, ls != 0x7fffffff ? 1 : 0), &err);
            if (lane == 0) s_excl = e;
        }
        __syncthreads();
        const unsigned long long excl = s_excl;
        if (!egdst_scan_hi(excl)) {  // the grid did not stop in an earlier chunk: this chunk's points count
            if (late != 0x7fffffff) egdst_fail(P, ivec, EGDST_ERR_RESEND_LATE, it, ist, id);
            if (badsum) egdst_fail(P, ivec, EGDST_ERR_CHECKSUM, it, ist, id);
            const int first = egdst_scan_lo(excl);
            int pos = first + woff;
#pragma unroll
            for (int j = 0; j < IPT; j++) {
                const int n = wbase + j * 32 + lane;
                if (bal[j] & (1u << lane)) {
                    const int dst = pos + __popc(bal[j] & ltmask);
                    if (dst < P.gcap) { X[dst] = rawM[n]; Cc[dst] = rawC[n]; V[dst] = rawV[n]; }
                }
                pos += __popc(bal[j]);
            }
            __syncthreads();
            // folds between neighbours that this CTA wrote itself (M or V decreasing, egdst_solver.c:819)
            const int lim = first + total < P.gcap ? first + total : P.gcap;
            for (int p = first + 1 + threadIdx.x; p < lim; p += blockDim.x)
                if (X[p - 1] > X[p] || V[p - 1] > V[p]) { const int k = atomicAdd(P.foldCnt + sd, 1); if (k <= P.gcap) foldList[k] = p; }
        }
    }
    if (err) egdst_fail(P, ivec, EGDST_ERR_ENV2SPACE, it, ist, id);
    // last CTA of this (ist,id): boundary folds, ordering, counts
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(P.tickC + 2 * sd + 1, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const unsigned long long tot = egdst_scan_inclusive(st, nch - 1);
    const int kept = egdst_scan_lo(tot);
    const int nvd = kept < P.gcap ? kept : P.gcap;
    for (int c = 1 + threadIdx.x; c < nch; c += blockDim.x) {
        const unsigned long long pe = egdst_scan_inclusive(st, c - 1), pi = egdst_scan_inclusive(st, c);
        const int p = egdst_scan_lo(pe);
        if (!egdst_scan_hi(pe) && egdst_scan_lo(pi) > p && p > 0 && p < nvd)
            if (EGDST_LDCG(X + p - 1) > EGDST_LDCG(X + p) || EGDST_LDCG(V + p - 1) > EGDST_LDCG(V + p)) {
                const int k = atomicAdd(P.foldCnt + sd, 1); if (k <= P.gcap) foldList[k] = p;
            }
    }
    __syncthreads();
    int nf = *((volatile int *)(P.foldCnt + sd));
    if (nf > P.gcap - 1) nf = P.gcap - 1;
    for (int i = threadIdx.x; i < nf; i += blockDim.x) {  // rank sort (folds are rare)
        const int v = EGDST_LDCG(foldList + i);
        int r = 0;
        for (int j = 0; j < nf; j++) r += EGDST_LDCG(foldList + j) < v ? 1 : 0;
        runStart[r + 1] = v;
    }
    if (threadIdx.x == 0) {
        if (kept >= P.cx.ngridmax) egdst_fail(P, ivec, EGDST_ERR_GRIDSPACE, it, ist, id);
        runStart[0] = 0;
        runStart[nf + 1] = nvd;
        P.ptN[sd] = nvd;
        P.nfold[sd] = nf;
        atomicAdd(P.units + ivec, (unsigned long long)nvd);
    }
}
