// egdst_envelope.cuh -- upper envelope of tabulated value functions, thresholds and double points.
//
// Parallel restatement of the reference's serial sweep:
//   envelop      egdst_solver.c:1165-1550   (qsort + sweep with cases 0/1/2)
//   funcvalue    egdst_solver.c:1553-1567   linter2 :1585-1593   comp1 :1570-1582
//   thresholds   egdst_solver.c:1596-1915   brsolve :1918-1968
//   envelope2    egdst_solver.c:776-913     (same routine, the "functions" are monotone runs of one id)
//
// The serial sweep visits the union of all points in (x asc, V desc, index asc) order, keeps a point
// when its own function is the maximum there, and inserts the crossing(s) of the previous and the new
// maximal function whenever the argmax changes.  Between two consecutive points of the union every
// function is a single linear piece (or its analytic credit-constrained branch), therefore
//   (A) every point can find its rank in the union and the argmax at its abscissa independently
//       (binary searches into the other functions' lists -- no sort, no sweep state), and
//   (B) the crossings belong to the boundaries r-1|r of the union where argmax(r-1) != argmax(r); the
//       chain of switches inside one boundary is found with the reference's own recursion rule.
// Output order is restored with block-wide prefix sums (warp shuffles).  In generic position this
// produces the same grid, values, thresholds and double points as the sweep; in exact ties it may
// differ by which of two coincident points is kept (the functions agree).
#pragma once

#include "egdst_common.cuh"

// ---- views over the input functions ---------------------------------------------------------
// MODE 0: primary envelope, function f = decision id f of state ist (lists ptX/ptC/ptV of sd0+f)
// MODE 1: secondary envelope of one (ist,id): function f = f-th monotone run; every run but the last is
//         extended by the constant-extrapolation sentinel (1.5*mmax, C_last, V_last) (egdst_solver.c:824-827)
template <int MODE>
struct EgdstEnvView {
    static const int kMode = MODE;
    int F;
    const double *X, *C, *V;  // MODE 0: base of decision 0 (stride gcap); MODE 1: the id's list
    const int *n;             // MODE 0: ptN + sd0 ; MODE 1: runStart
    const double *evfa0;      // MODE 0: evfa0 + sd0 ; MODE 1: &evfa0[sd]
    int gcap, id;
    double sentinel;

    EGDST_DEV_M int npts(int f) const {
        if (MODE == 0) return n[f];
        return n[f + 1] - n[f] + (f < F - 1 ? 1 : 0);
    }
    EGDST_DEV_M int pstart(int f) const {  // offset of function f in the flattened point order
        if (MODE == 0) { int s = 0; for (int g = 0; g < f; g++) s += n[g]; return s; }
        return n[f] + f;
    }
    EGDST_DEV_M double x(int f, int k) const {
        if (MODE == 0) return X[(size_t)f * gcap + k];
        const int real = n[f + 1] - n[f];
        return k < real ? X[n[f] + k] : sentinel;
    }
    EGDST_DEV_M double c(int f, int k) const {
        if (MODE == 0) return C[(size_t)f * gcap + k];
        const int real = n[f + 1] - n[f];
        return C[n[f] + (k < real ? k : real - 1)];
    }
    EGDST_DEV_M double v(int f, int k) const {
        if (MODE == 0) return V[(size_t)f * gcap + k];
        const int real = n[f + 1] - n[f];
        return V[n[f] + (k < real ? k : real - 1)];
    }
    EGDST_DEV_M double evf(int f) const {
        if (MODE == 0) return evfa0[f];
        return f == 0 ? evfa0[0] : -EGDST_INF;
    }
    EGDST_DEV_M int fid(int f) const { return MODE == 0 ? f : id; }
};

// linter2 (egdst_solver.c:1585-1593): no extrapolation
EGDST_DEV double egdst_linter2(double x, double g0, double g1, double f0, double f1) {
    if (x == g0) return f0;
    if (x < g0) return -EGDST_INF;
    if (x > g1) return -EGDST_INF;
    return f1 * (x - g0) / (g1 - g0) + f0 * (g1 - x) / (g1 - g0);
}

// number of points of function g that precede the point (x, v, f, k) in the sweep order
// (x asc, V desc, function asc -- comp1, egdst_solver.c:1570-1582; k breaks ties inside one function)
template <class View>
EGDST_DEV int egdst_env_count_before(const View &E, int g, double x, double v, int f, int k) {
    const int ng = E.npts(g);
    // Lists that do not straddle x need no search: three independent probes instead of a chain of dependent ones.
    // The secondary envelope of a zig-zagging grid has ~10^2 short runs of which two or three contain a given x;
    // every run but the last ends in the far-away sentinel, so "x between the last two points" is the common case.
    if (ng > 0) {
        if (x < E.x(g, 0)) return 0;
        const double xl = E.x(g, ng - 1);
        if (x > xl) return ng;
        if (ng >= 2 && x < xl && x > E.x(g, ng - 2)) return ng - 1;
    }
    int lo = 0, hi = ng;  // lower bound: first index with x_g >= x
    while (lo < hi) { int mid = (lo + hi) >> 1; if (E.x(g, mid) < x) lo = mid + 1; else hi = mid; }
    int cnt = lo;
    for (int j = lo; j < ng && E.x(g, j) == x; j++) {
        const double vj = E.v(g, j);
        if (vj > v || (vj == v && (g < f || (g == f && j < k)))) cnt++;
    }
    return cnt;
}

// segment index of function g "current" at a sweep position: min(#processed-1, n-2) (egdst_solver.c:1524)
EGDST_DEV int egdst_env_cur(int cnt, int ng) { int c = cnt - 1; return c < ng - 2 ? c : ng - 2; }

template <class View>
EGDST_DEV double egdst_env_analytic(const egdst_ctx *cx, const View &E, int it, int ist, int g, double x) {
    PeriodVars cu; cu.it = it; cu.ist = ist; cu.id = E.fid(g); cu.cash = 0; cu.savings = 0; cu.shock = 0;
    return utility(cx, &cu, x - cx->a0) + discount(cx, &cu) * E.evf(g);
}

// funcvalue (egdst_solver.c:1553-1567)
template <class View>
EGDST_DEV double egdst_env_value(const egdst_ctx *cx, const View &E, int it, int ist, int g, int cur, double x) {
    if (cur >= 0) return egdst_linter2(x, E.x(g, cur), E.x(g, cur + 1), E.v(g, cur), E.v(g, cur + 1));
    if (E.evf(g) == -EGDST_INF) return -EGDST_INF;
    return egdst_env_analytic(cx, E, it, ist, g, x);
}
// the second tabulated function (consumption) of g at x, with the credit-constrained branch
template <class View>
EGDST_DEV double egdst_env_value2(const egdst_ctx *cx, const View &E, int g, int cur, double x) {
    if (cur >= 0) return egdst_linter2(x, E.x(g, cur), E.x(g, cur + 1), E.c(g, cur), E.c(g, cur + 1));
    if (E.evf(g) == -EGDST_INF) return cx->zeroconsumption;
    return x - cx->a0;
}

// ---------------------------------------------------------------------------------------------
// Step A: one thread per input point: rank in the union, argmax at its abscissa.
// grid (ceil(maxP/B), njobs_y, nvec)
// ---------------------------------------------------------------------------------------------
template <int MODE>
EGDST_DEV bool egdst_env_job(const EgdstDev &P, int ivec, int jy, int &ist, int &id, int &slot, int &sslot, EgdstEnvView<MODE> &E) {
    // slot: the job's index in the data arrays (mg*, out*); sslot: its index in the synchronisation arrays (scanE, tickE,
    // envNact), which the vector-per-CTA scope keeps in shared memory
    if (MODE == 0) {
        ist = jy; id = 0;
        const int sd0 = egdst_sd(P, ivec, ist, 0);
        int any = 0;
        for (int f = 0; f < P.cx.nd; f++) any |= P.active[sd0 + f];
        if (!any) return false;
        E.F = P.cx.nd; E.X = P.ptX + (size_t)sd0 * P.gcap; E.C = P.ptC + (size_t)sd0 * P.gcap; E.V = P.ptV + (size_t)sd0 * P.gcap;
        E.n = P.ptN + sd0; E.evfa0 = P.evfa0 + sd0; E.gcap = P.gcap; E.id = 0; E.sentinel = 0;
        slot = P.nvec * P.cx.nst * P.cx.nd + ivec * P.cx.nst + ist;
        sslot = P.priSync0 + ivec * P.cx.nst + ist;
        return true;
    } else {
        ist = jy / P.cx.nd; id = jy % P.cx.nd;
        const int sd = egdst_sd(P, ivec, ist, id);
        if (!P.active[sd] || P.nfold[sd] == 0) return false;
        E.F = P.nfold[sd] + 1; E.X = P.ptX + (size_t)sd * P.gcap; E.C = P.ptC + (size_t)sd * P.gcap; E.V = P.ptV + (size_t)sd * P.gcap;
        E.n = P.runStart + (size_t)sd * (P.gcap + 1); E.evfa0 = P.evfa0 + sd; E.gcap = P.gcap; E.id = id; E.sentinel = 1.5 * P.cx.mmax;
        slot = sd;
        sslot = sd;
        return true;
    }
}

// unified grid bound = min over functions of the last abscissa (egdst_solver.c:1266-1271), computed once per CTA
// (the secondary envelope can have ~10^2 functions).  sh: 33 doubles.  Contains two egdst_cta_sync().
template <class View>
EGDST_DEV double egdst_env_grb_block(const View &E, double *sh) {
    double g = EGDST_INF;
    for (int f = threadIdx.x; f < E.F; f += blockDim.x) { const int n = E.npts(f); if (n > 0) { const double xl = E.x(f, n - 1); if (xl < g) g = xl; } }
    for (int o = 16; o > 0; o >>= 1) { const double w = __shfl_xor_sync(EGDST_FULL, g, o); if (w < g) g = w; }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) sh[w] = g;
    egdst_cta_sync();
    double r = sh[0];
    for (int k = 1; k < nw; k++) if (sh[k] < r) r = sh[k];
    egdst_cta_sync();
    return r;
}

// One input point p of the flattened lists: its (f,k), abscissa and value, and -- over the functions g = part,
// part + nparts, ... -- the number of points that precede it in the union and the best function at its abscissa
// (highest value, lowest index among equals: the sweep's "first strictly greater wins"; the point's own function takes
// part with its own value).
template <int MODE>
EGDST_DEV void egdst_env_point_partial(const egdst_ctx &cx, const EgdstEnvView<MODE> &E, int it, int ist, int p, int part, int nparts,
                                       int &f, int &k, double &x, double &v, int &rank, int &best, double &bestv) {
    f = 0;
    if (MODE == 0) { int s = 0; while (f < E.F - 1 && p >= s + E.npts(f)) { s += E.npts(f); f++; } }
    else { int lo = 0, hi = E.F - 1; while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (E.pstart(mid) <= p) lo = mid; else hi = mid - 1; } f = lo; }
    k = p - E.pstart(f);
    x = E.x(f, k); v = E.v(f, k);
    rank = 0; best = 0x7fffffff; bestv = -EGDST_INF;
    // inside its own (increasing) list a point is preceded by exactly k points unless a neighbour shares its abscissa
    const bool tie = (k > 0 && E.x(f, k - 1) == x) || (k + 1 < E.npts(f) && E.x(f, k + 1) == x);
    for (int g = part; g < E.F; g += nparts) {
        const int ng = E.npts(g);
        if (ng <= 0) continue;
        const int cnt = (g == f && !tie) ? k : egdst_env_count_before(E, g, x, v, f, k);
        rank += cnt;
        const double val = (g == f) ? v : egdst_env_value(&cx, E, it, ist, g, egdst_env_cur(cnt, ng), x);
        if (val > bestv || (val == bestv && g < best)) { bestv = val; best = g; }
    }
}
// record the point at its position of the union; the active positions are the prefix with x <= grb
EGDST_DEV void egdst_env_point_commit(const EgdstDev &P, int slot, int sslot, double grb, int f, int k, double x, int rank, int best) {
    if (best == 0x7fffffff) best = f;  // every value -inf (cannot happen: the own value is finite or the maximum)
    const size_t o = (size_t)slot * P.envcap + rank;
    P.mgX[o] = x; P.mgF[o] = f; P.mgK[o] = k; P.mgA[o] = best;
}
// length of the active prefix of the union = 1 + the largest rank of a point with x <= grb.  One atomic per warp, not per
// point: at S1 the rank step of a period would otherwise send 2*10^4 atomics to one word (they serialise in L2).
// Called by all threads of the item (cand = 0 for threads without a point).
EGDST_DEV void egdst_env_active_prefix(const EgdstDev &P, int sslot, int cand) {
    for (int o = 16; o > 0; o >>= 1) { const int w = __shfl_xor_sync(EGDST_FULL, cand, o); cand = w > cand ? w : cand; }
    if ((threadIdx.x & 31) == 0 && cand > 0) atomicMax(P.envNact + sslot, cand);
}
// "all choices produced empty grids" (egdst_solver.c:704-710), checked where the per-decision lists are final
EGDST_DEV void egdst_env_check_allinf(const EgdstDev &P, int ivec, int it, int ist) {
    const int sd0 = egdst_sd(P, ivec, ist, 0);
    int any = 0, tot = 0;
    for (int d_ = 0; d_ < P.cx.nd; d_++) { any |= P.active[sd0 + d_]; tot += P.ptN[sd0 + d_]; }
    if (any && tot == 0) egdst_fail(P, ivec, EGDST_ERR_ALLINF, it, ist, -1);
}

// Step A as a phase: work item = (job, virtual block of npt points); nvb virtual blocks per job stride over its points.
// `nparts` threads share the functions of one point and combine rank and argmax through shared memory -- the
// secondary envelope of a zig-zagging grid has ~10^2 runs, and a point that walks them alone is a chain of ~10^2
// dependent loads.
template <int BS>
struct EgdstRankShared { double shg[33]; double bv[BS]; int rank[BS], best[BS]; };

template <int MODE, int BS>
EGDST_DEV void egdst_ph_envA(const EgdstDev &P, int it, const EgdstTeam &T, int nvb, void *scratch) {
    EgdstRankShared<BS> &R = *reinterpret_cast<EgdstRankShared<BS> *>(scratch);
    double *shg = R.shg, *s_bv = R.bv;
    int *s_rank = R.rank, *s_best = R.best;
    const int nparts = (MODE == 1) ? P.envA1parts : 1, npt = blockDim.x / nparts;
    const int lane = threadIdx.x % npt, part = threadIdx.x / npt;
    const int jpv = (MODE == 0) ? P.cx.nst : P.cx.nst * P.cx.nd;
    const int nwork = T.nv * jpv * nvb;
    for (int w = T.rank; w < nwork; w += T.size) {
        int ivec, jy, vb;
        egdst_item(T, w, jpv, ivec, jy, vb);
        int ist, id, slot, sslot;
        EgdstEnvView<MODE> E;
        if (MODE == 0 && vb == 0 && threadIdx.x == 0) egdst_env_check_allinf(P, ivec, it, jy);
        if (!egdst_env_job<MODE>(P, ivec, jy, ist, id, slot, sslot, E)) continue;
        const int Ptot = E.pstart(E.F - 1) + E.npts(E.F - 1);
        if (vb * npt >= Ptot) continue;  // CTA-uniform
        egdst_ctx cx; egdst_load_ctx(P, ivec, cx);
        egdst_cta_sync();  // scratch of the previous item is free
        const double grb = egdst_env_grb_block(E, shg);
        for (int base = vb * npt; base < Ptot; base += nvb * npt) {
            const int p = base + lane;
            const bool valid = p < Ptot;
            int f = 0, k = 0, rank = 0, best = 0x7fffffff;
            double x = 0, v = 0, bestv = -EGDST_INF;
            if (valid) egdst_env_point_partial<MODE>(cx, E, it, ist, p, part, nparts, f, k, x, v, rank, best, bestv);
            if (nparts > 1) {
                const int slotx = part * npt + lane;
                s_rank[slotx] = rank; s_best[slotx] = best; s_bv[slotx] = bestv;
                egdst_cta_sync();
                if (part == 0 && valid) {
                    for (int q = 1; q < nparts; q++) {
                        const int o = q * npt + lane;
                        rank += s_rank[o];
                        if (s_bv[o] > bestv || (s_bv[o] == bestv && s_best[o] < best)) { bestv = s_bv[o]; best = s_best[o]; }
                    }
                }
            }
            if (part == 0 && valid) egdst_env_point_commit(P, slot, sslot, grb, f, k, x, rank, best);
            egdst_env_active_prefix(P, sslot, (part == 0 && valid && x <= grb) ? rank + 1 : 0);
            if (nparts > 1 && base + nvb * npt < Ptot) egdst_cta_sync();  // scratch reused by the next stride
        }
    }
}

// ---------------------------------------------------------------------------------------------
// crossing chain on the boundary before sweep position (xr,vr,fr,kr): thresholds() restated with an
// explicit stack.  Counts (write==false) or writes the grid points / thresholds it produces.
// WARP-COOPERATIVE: must be called by all 32 lanes of a warp with identical arguments (lane 0 writes).
// ---------------------------------------------------------------------------------------------
template <class View>
EGDST_DEV double egdst_env_brsolve(const egdst_ctx *cx, const View &E, int it, int ist, int ga, double br0, double br1,
                                   double g0, double g1, double f0, double f1, int *err) {
    const double dd = cx->doublepoint_delta;
    for (int iter = 0; iter < 400; iter++) {
        const double fa0 = egdst_env_analytic(cx, E, it, ist, ga, br0), fa1 = egdst_env_analytic(cx, E, it, ist, ga, br1);
        const double s0 = (fa0 - egdst_linter2(br0, g0, g1, f0, f1)) > 0 ? 1.0 : -1.0;
        const double s1 = (fa1 - egdst_linter2(br1, g0, g1, f0, f1)) > 0 ? 1.0 : -1.0;
        if (s0 == s1 || br0 > br1) { *err = EGDST_ERR_BRACKET; return br0; }
        if (fabs(br0 - br1) < 2 * dd || fabs(fa0 - fa1) < dd) return (br0 + br1) / 2;
        const double mid = (br0 + br1) / 2;
        const double sm = (egdst_env_analytic(cx, E, it, ist, ga, mid) - egdst_linter2(mid, g0, g1, f0, f1)) > 0 ? 1.0 : -1.0;
        if (s0 == sm) br0 = mid; else if (s1 == sm) br1 = mid; else return br0;
    }
    return (br0 + br1) / 2;
}

template <class View>
EGDST_DEV void egdst_env_chain(const egdst_ctx *cx, const View &E, int it, int ist, double xl, double xr, double vr, int fr, int kr,
                               int pri0, int nwi0, bool write, double *gx, double *gv, double *gc, double *ga, int gleft,
                               double *tth, double *tdd, int tleft, int &ng, int &nt, int *err) {
    unsigned marks[EGDST_ENV_MARKW];
    for (int w = 0; w < EGDST_ENV_MARKW; w++) marks[w] = 0u;
    if (E.F > 32 * EGDST_ENV_MARKW) { *err = EGDST_ERR_ENV2SPACE; ng = 0; nt = 0; return; }
    short sp_p[EGDST_ENV_STACK], sp_q[EGDST_ENV_STACK];
    int sp = 0;
    sp_p[0] = (short)pri0; sp_q[0] = (short)nwi0; sp = 1;
    ng = 0; nt = 0;
    while (sp > 0) {
        sp--;
        const int p = sp_p[sp], q = sp_q[sp];
        marks[p >> 5] |= 1u << (p & 31);
        marks[q >> 5] |= 1u << (q & 31);
        const int np = E.npts(p), nq = E.npts(q);
        const int cp = egdst_env_cur(egdst_env_count_before(E, p, xr, vr, fr, kr), np);
        const int cq = egdst_env_cur(egdst_env_count_before(E, q, xr, vr, fr, kr), nq);
        double newpoint, cmax;
        if (cp == -1 && cq != -1) {  // pri analytic, nwi linear
            const double g0 = E.x(q, cq), g1 = E.x(q, cq + 1), f0 = E.v(q, cq), f1 = E.v(q, cq + 1);
            if (E.evf(p) == -EGDST_INF) newpoint = E.x(p, 0);
            else newpoint = egdst_env_brsolve(cx, E, it, ist, p, g0, MIN(E.x(p, 0), g1), g0, g1, f0, f1, err);
            cmax = egdst_linter2(newpoint, g0, g1, f0, f1);
        } else if (cp != -1 && cq == -1) {  // nwi analytic, pri linear
            const double g0 = E.x(p, cp), g1 = E.x(p, cp + 1), f0 = E.v(p, cp), f1 = E.v(p, cp + 1);
            if (E.evf(q) == -EGDST_INF) newpoint = E.x(q, 0);
            else newpoint = egdst_env_brsolve(cx, E, it, ist, q, g0, MIN(E.x(q, 0), g1), g0, g1, f0, f1, err);
            cmax = egdst_linter2(newpoint, g0, g1, f0, f1);
        } else if (cp == -1 && cq == -1) {
            *err = EGDST_ERR_TWO_ANALYTIC;
            return;
        } else {
            const double pg0 = E.x(p, cp), pg1 = E.x(p, cp + 1), pf0 = E.v(p, cp), pf1 = E.v(p, cp + 1);
            const double qg0 = E.x(q, cq), qg1 = E.x(q, cq + 1), qf0 = E.v(q, cq), qf1 = E.v(q, cq + 1);
            const double sq = (qf1 - qf0) / (qg1 - qg0), iq = (qf0 * qg1 - qf1 * qg0) / (qg1 - qg0);
            const double spp = (pf1 - pf0) / (pg1 - pg0), ip = (pf0 * pg1 - pf1 * pg0) / (pg1 - pg0);
            if (pg1 == pg0) { newpoint = pg0; cmax = newpoint * sq + iq; }
            else if (qg1 == qg0) { newpoint = qg0; cmax = newpoint * spp + ip; }
            else if (sq == spp) {
                // Parallel pieces.  Between two consecutive abscissas of the union both functions are single linear
                // pieces, so the maximum can pass from one to the other only where they coincide: the constant
                // extrapolations of two runs of a flat stretch (egdst_solver.c:824-827), told apart by the last bit of
                // an interpolation.  The reference places a double point at the mean of the four end points here
                // (:1737-1753) -- an abscissa far outside the bracket.  Here nothing is emitted for EXACTLY parallel pieces of
                // the secondary envelope: a deliberate difference from the reference, seen only in constructed ties
                // (tests/golden/tie_env2.npz, where the reference's own list comes out unsorted); DESIGN.md section 5.
                if (View::kMode == 1) continue;
                newpoint = (pg0 + pg1 + qg0 + qg1) / 4; cmax = newpoint * sq + iq;
            }
            else { newpoint = (ip - iq) / (sq - spp); cmax = newpoint * sq + iq; }
            // With slopes that differ only in the last bit the "intersection" of two all but coincident pieces lands anywhere,
            // usually far outside the bracket of the boundary it belongs to.  The reference emits it regardless (envelope2
            // copies back every point of envelop(), egdst_solver.c:888-895) and its list comes out UNSORTED.  What happens
            // next depends on the model: with several decisions the primary envelope qsorts the merged quadruples, and in
            // every case observed (S1b at BASELINE size, periods 8-6) the stray pair does not reach the solution cell; with
            // a single decision nothing re-sorts and it does (examples.deaton_meanstest, period 1, a row with C = -inf).
            // The rank step of the primary envelope here relies on sorted per-decision lists, so: a single decision emits
            // the pair like the reference; several decisions drop it.  This is an EMPIRICAL rule pinned by those two cases,
            // not an equivalence -- DESIGN.md section 5 states the gap.
            if (View::kMode == 1 && cx->nd > 1 && !(newpoint >= xl && newpoint <= xr)) continue;
        }
        // is a third, not yet visited function above at the crossing? (egdst_solver.c:1807-1845, mode 1)
        // The candidates are split over the lanes of the warp (the secondary envelope can have ~10^2 runs);
        // the serial rule "first strictly greater value wins" = maximum value, lowest index among ties.
        int optk = -1;
        {
            const int lane_ = threadIdx.x & 31;
            double lv = -EGDST_INF; int lk = 0x7fffffff;
            for (int k = lane_; k < E.F; k += 32) {
                if (marks[k >> 5] & (1u << (k & 31))) continue;
                const int nk = E.npts(k);
                if (nk <= 0) continue;
                const int ck = egdst_env_cur(egdst_env_count_before(E, k, xr, vr, fr, kr), nk);
                const double tmax = (ck >= 0) ? egdst_linter2(newpoint, E.x(k, ck), E.x(k, ck + 1), E.v(k, ck), E.v(k, ck + 1))
                                              : egdst_env_analytic(cx, E, it, ist, k, newpoint);
                if (tmax > lv) { lv = tmax; lk = k; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double ov_ = __shfl_xor_sync(EGDST_FULL, lv, o);
                const int ok_ = __shfl_xor_sync(EGDST_FULL, lk, o);
                if (ov_ > lv || (ov_ == lv && ok_ < lk)) { lv = ov_; lk = ok_; }
            }
            if (cmax < lv) { cmax = lv; optk = lk; }
        }
        if (optk != -1) {
            if (sp + 2 > EGDST_ENV_STACK) { *err = EGDST_ERR_ENV2SPACE; return; }
            sp_p[sp] = (short)optk; sp_q[sp] = (short)q; sp++;  // right part, processed second
            sp_p[sp] = (short)p; sp_q[sp] = (short)optk; sp++;  // left part, processed first
            continue;
        }
        const double c_left = egdst_env_value2(cx, E, p, cp, newpoint), c_right = egdst_env_value2(cx, E, q, cq, newpoint);
        const bool single = (E.evf(q) == -EGDST_INF && cq == -1);  // egdst_solver.c:1892-1896
        if (write && (threadIdx.x & 31) == 0) {
            if (ng < gleft) {
                const double xs = single ? newpoint - cx->tolerance : newpoint;
                gx[ng] = xs; gv[ng] = cmax; gc[ng] = c_left;
                if (ga) ga[ng] = xs - c_left;
            }
            if (nt < tleft) { tth[nt] = newpoint; tdd[nt] = (double)q; }
            if (!single && ng + 1 < gleft) {
                const double xd = newpoint + cx->doublepoint_delta;
                gx[ng + 1] = xd; gv[ng + 1] = cmax; gc[ng + 1] = c_right;
                if (ga) ga[ng + 1] = xd - c_right;
            }
        }
        ng += single ? 1 : 2;
        nt += 1;
    }
}

// one position r of the union: its own contribution (kept point, first threshold) and whether a crossing chain
// sits on the boundary before it.  The chains themselves are queued and run one per warp (egdst_ph_envBC).
struct EgdstEnvPos { double x, v; int f, k, a, aprev; bool newx, chain; };
template <int MODE>
EGDST_DEV EgdstEnvPos egdst_env_pos(const EgdstEnvView<MODE> &E, const double *mgX, const int *mgF, const int *mgK, const int *mgA, int r) {
    EgdstEnvPos q;
    q.x = mgX[r]; q.f = mgF[r]; q.k = mgK[r]; q.a = mgA[r];
    q.v = E.v(q.f, q.k);
    q.newx = (r == 0) || (mgX[r - 1] < q.x);
    q.aprev = 0; q.chain = false;
    if (r > 0 && q.newx) { q.aprev = mgA[r - 1]; q.chain = q.aprev != q.a; }
    return q;
}

// Step B/C as a phase.  The union of a job is cut into chunks of blockDim.x positions (one per thread); the nvb work
// items of a job take chunks by ticket until none is left, and the chunks are chained by a decoupled look-back scan
// over (grid points, thresholds) emitted so far.  The rare crossing chains are queued per chunk and run one per warp.
// The last item of a job to finish writes the cell header (MODE 0) or copies the staged result back over the
// decision's point list (MODE 1).
// MODE 0 writes the period's solution cell (rows 1.., thresholds, evf, row 0); MODE 1 rewrites the id's list.
template <int BS>
struct EgdstEnvShared {
    long long sh[40];
    double grb[33];
    int chunk, last, qn;
    unsigned long long excl;
    int qr[BS], qg[BS], qt[BS], qgpos[BS], qtpos[BS];
};

template <int MODE, int BS>
EGDST_DEV void egdst_ph_envBC(const EgdstDev &P, int it, const EgdstTeam &T, int nvb, void *scratch) {
    EgdstEnvShared<BS> &Sh = *reinterpret_cast<EgdstEnvShared<BS> *>(scratch);
    const int jpv = (MODE == 0) ? P.cx.nst : P.cx.nst * P.cx.nd;
    const int nitems = nvb;
    const int nwork = T.nv * jpv * nitems;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int w = T.rank; w < nwork; w += T.size) {
        int ivec, jy, vb;
        egdst_item(T, w, jpv, ivec, jy, vb);
        int ist, id, slot, sslot;
        EgdstEnvView<MODE> E;
        if (!egdst_env_job<MODE>(P, ivec, jy, ist, id, slot, sslot, E)) continue;
        egdst_ctx cx; egdst_load_ctx(P, ivec, cx);
        const double *mgX = P.mgX + (size_t)slot * P.envcap;
        const int *mgF = P.mgF + (size_t)slot * P.envcap, *mgK = P.mgK + (size_t)slot * P.envcap, *mgA = P.mgA + (size_t)slot * P.envcap;
        volatile unsigned long long *st = P.scanE + (size_t)sslot * P.chE;
        double *ox, *oc, *ov, *oa = 0, *oth = 0, *odd = 0;
        int gcapacity, tcapacity = 0, cell = 0;
        if (MODE == 0) {
            cell = egdst_cell(P, ivec, it, ist);
            ox = egdst_colM(P, cell) + 1; oc = egdst_colC(P, cell) + 1; ov = egdst_colV(P, cell) + 1; oa = egdst_colA(P, cell) + 1;
            oth = P.thTH + (size_t)cell * cx.nthrhmax; odd = P.thD + (size_t)cell * cx.nthrhmax;
            gcapacity = P.rowcap - 1; tcapacity = cx.nthrhmax;
        } else {
            ox = P.outX + (size_t)slot * P.envcap; oc = P.outC + (size_t)slot * P.envcap; ov = P.outV + (size_t)slot * P.envcap;
            gcapacity = P.envcap;
        }
        // the active positions of the union are the prefix with x<=grb (length recorded by egdst_ph_envA)
        const int nact = *((volatile int *)(P.envNact + sslot));
        const int chunkw = blockDim.x;
        const int nch = (nact + chunkw - 1) / chunkw;
        int err = 0, serr = 0;
        const double grb = nch > 0 ? egdst_env_grb_block(E, Sh.grb) : 0.0;  // nch is CTA-uniform
        while (true) {
            egdst_cta_sync();  // the previous chunk's queue and scan scratch are free again
            if (threadIdx.x == 0) { Sh.chunk = atomicAdd(P.tickE + 2 * sslot, 1); Sh.qn = 0; }
            egdst_cta_sync();
            const int chunk = Sh.chunk;
            if (chunk >= nch) break;
            const int r = chunk * chunkw + threadIdx.x;
            // pass 1: own contribution of every position; crossing chains go to the item's queue
            int ngj = 0, ntj = 0, qj = -1;
            EgdstEnvPos q; q.x = 0; q.v = 0; q.f = 0; q.k = 0; q.a = 0; q.aprev = 0; q.newx = false; q.chain = false;
            if (r < nact) {
                q = egdst_env_pos<MODE>(E, mgX, mgF, mgK, mgA, r);
                if (r == 0) ntj = 1;  // (a0, argmax at the first point)  egdst_solver.c:1321-1325
                if (q.newx && (q.a == q.f || q.x == grb)) ngj = 1;
                if (q.chain) { const int c = atomicAdd(&Sh.qn, 1); Sh.qr[c] = r; qj = c; }
            }
            egdst_cta_sync();
            const int qn = Sh.qn;
            // pass 2: one chain per warp, counting
            for (int c = warp; c < qn; c += nwarps) {
                const int rr = Sh.qr[c];
                const EgdstEnvPos qq = egdst_env_pos<MODE>(E, mgX, mgF, mgK, mgA, rr);
                int cg, ct;
                egdst_env_chain(&cx, E, it, ist, mgX[rr - 1], qq.x, qq.v, qq.f, qq.k, qq.aprev, qq.a, false, (double *)0, (double *)0, (double *)0, (double *)0, 0, (double *)0, (double *)0, 0, cg, ct, &err);
                if (lane == 0) { Sh.qg[c] = cg; Sh.qt[c] = ct; }
            }
            egdst_cta_sync();
            if (qj >= 0) { ngj += Sh.qg[qj]; ntj += Sh.qt[qj]; }
            long long tot;
            const long long off = egdst_block_excl_scan64(((long long)ntj << 32) | (long long)ngj, Sh.sh, &tot);
            if (threadIdx.x < 32) {
                const unsigned long long e = egdst_lookback<0>(st, chunk, egdst_scan_pack((int)(tot & 0xffffffffLL), (int)(tot >> 32)), &serr);
                if (threadIdx.x == 0) Sh.excl = e;
            }
            egdst_cta_sync();
            const int gpos = egdst_scan_lo(Sh.excl) + (int)(off & 0xffffffffLL), tpos = egdst_scan_hi(Sh.excl) + (int)(off >> 32);
            // pass 3: write the kept points; chains get their output offsets
            if (r < nact && (ngj | ntj)) {
                int g = gpos, t = tpos;
                if (r == 0) { if (MODE == 0 && t < tcapacity) { oth[t] = cx.a0; odd[t] = (double)q.a; } t += 1; }
                if (qj >= 0) { Sh.qgpos[qj] = g; Sh.qtpos[qj] = t; g += Sh.qg[qj]; }
                if (q.newx && (q.a == q.f || q.x == grb)) {
                    double v = q.v, c;
                    if (q.a == q.f) c = E.c(q.f, q.k);
                    else {  // last abscissa of the unified grid: keep the interpolated maximum
                        const int na = E.npts(q.a);
                        const int ca = egdst_env_cur(egdst_env_count_before(E, q.a, q.x, q.v, q.f, q.k), na);
                        v = egdst_env_value(&cx, E, it, ist, q.a, ca, q.x);
                        c = egdst_env_value2(&cx, E, q.a, ca, q.x);
                    }
                    if (g < gcapacity) { ox[g] = q.x; ov[g] = v; oc[g] = c; if (oa) oa[g] = q.x - c; }
                }
            }
            egdst_cta_sync();
            // pass 4: one chain per warp, writing
            for (int c = warp; c < qn; c += nwarps) {
                const int rr = Sh.qr[c];
                const EgdstEnvPos qq = egdst_env_pos<MODE>(E, mgX, mgF, mgK, mgA, rr);
                const int bg = Sh.qgpos[c], bt = Sh.qtpos[c];
                int cg, ct;
                egdst_env_chain(&cx, E, it, ist, mgX[rr - 1], qq.x, qq.v, qq.f, qq.k, qq.aprev, qq.a, true, ox + bg, ov + bg, oc + bg, oa ? oa + bg : (double *)0, gcapacity - bg,
                                MODE == 0 ? oth + bt : (double *)0, MODE == 0 ? odd + bt : (double *)0, MODE == 0 ? tcapacity - bt : 0, cg, ct, &err);
            }
        }
        if (err) egdst_fail(P, ivec, err, it, ist, id);
        if (serr) egdst_fail(P, ivec, EGDST_ERR_ENV2SPACE, it, ist, id);
        // last item of the job: totals and epilogue
        __threadfence();
        egdst_cta_sync();
        if (threadIdx.x == 0) Sh.last = (atomicAdd(P.tickE + 2 * sslot + 1, 1) == nitems - 1);
        egdst_cta_sync();
        if (!Sh.last) continue;
        __threadfence();
        const unsigned long long totals = nch > 0 ? egdst_scan_inclusive(st, nch - 1) : 0ULL;
        const int nout = egdst_scan_lo(totals), nth = egdst_scan_hi(totals);
        if (MODE == 0) {
            if (threadIdx.x == 0) {
                if (nout >= cx.ngridmax) egdst_fail(P, ivec, EGDST_ERR_GRIDSPACE, it, ist, -1);
                if (nth >= cx.nthrhmax) egdst_fail(P, ivec, EGDST_ERR_THRSPACE, it, ist, -1);
                if (nout == 0 || nth == 0) egdst_fail(P, ivec, EGDST_ERR_ENVELOPE, it, ist, -1);
                const int n = nout < gcapacity ? nout : gcapacity;
                const int d0 = nact > 0 ? EGDST_LDCG(mgA) : 0;
                const double e = E.evf(d0);  // egdst_solver.c:730
                P.evf[cell] = e;
                P.mlen[cell] = n + 1;
                P.thlen[cell] = nth < tcapacity ? nth : tcapacity;
                egdst_colM(P, cell)[0] = cx.a0; egdst_colC(P, cell)[0] = 0.0; egdst_colV(P, cell)[0] = e;  // saveoutput :931-941
                egdst_colA(P, cell)[0] = cx.a0 - 0.0;
            }
        } else {
            const int sd = slot;
            if (threadIdx.x == 0 && nout >= cx.ngridmax) egdst_fail(P, ivec, EGDST_ERR_ENV2SPACE, it, ist, id);
            const int n = nout < P.gcap ? nout : P.gcap;
            double *X = P.ptX + (size_t)sd * P.gcap, *Cc = P.ptC + (size_t)sd * P.gcap, *V = P.ptV + (size_t)sd * P.gcap;
            // copy-back by one CTA: batches of independent loads (a plain loop is a chain of n/blockDim L2 round trips)
            for (int i0 = threadIdx.x; i0 < n; i0 += 8 * blockDim.x) {
                double bx[8], bc[8], bv[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int i = i0 + u * blockDim.x;
                    if (i < n) { bx[u] = EGDST_LDCG(ox + i); bc[u] = EGDST_LDCG(oc + i); bv[u] = EGDST_LDCG(ov + i); }
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int i = i0 + u * blockDim.x;
                    if (i < n) { X[i] = bx[u]; Cc[i] = bc[u]; V[i] = bv[u]; }
                }
            }
            if (threadIdx.x == 0) P.ptN[sd] = n;
        }
    }
}
