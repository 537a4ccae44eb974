// egdst_simulator.cuh -- forward Monte-Carlo simulation of agents on a solved model.
//
// Restates the reference's simulator (@egdstmodel/egdst_simulator.c):
//   simulator()   :204-383   one agent's path over it = 0..T-t0 (survival, state transition by inverse-CDF
//                            sampling over trpr, shock = cdfinv(u) or expectation, budget, equations)
//   policy()      :145-199   c = linter(cash; M,C), id by threshold scan, vf exact below M(a0) else linter
//   simsoutput()  :122-143   the nsimout output columns
// One thread per agent, all agents of a CTA in the same period (the period's policy table is shared
// through L1/L2).  The [nsimout, nt, nsim] output is written through a per-warp shared-memory tile so
// that each store instruction covers contiguous runs of an agent's record (the per-thread pattern of
// the reference has stride nsimout*nt).  Uniforms come either from the reference's randstream layout
// (parity mode) or from counter-based Philox4x32-10 keyed by (seed; global agent id, period).
// Continuous states (egdst_simulator.c:310-373) are outside the hot-path scope.
#pragma once

#include "egdst_common.cuh"

#define EGDST_NSIMOUT_MAX (11 + EGDST_NNST + EGDST_NND + EGDST_NREQ)
#ifdef EGDST_HOSTEMU
#define EGDST_SIM_BLOCK 64
#else
#define EGDST_SIM_BLOCK 128
#endif

// Philox4x32-10 (Salmon et al. 2011); counter = (c0,c1,c2,c3), key = (k0,k1)
EGDST_DEV void egdst_philox4x32(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned out[4]) {
    for (int r = 0; r < 10; r++) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
        const unsigned hi0 = (unsigned)(p0 >> 32), lo0 = (unsigned)p0, hi1 = (unsigned)(p1 >> 32), lo1 = (unsigned)p1;
        const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
EGDST_DEV double egdst_u01(unsigned x) { return ((double)x + 0.5) * (1.0 / 4294967296.0); }

struct EgdstSimArgs {
    const double *init;        // [nsim*2] column-major: 1-based ist0, m0
    int nsim;
    int ivec;
    const double *randstream;  // reference layout, or null => Philox
    int rndtype;               // 1 = same shocks for all agents
    long long agent0;
    unsigned long long seed;
    double *sims;              // [nsimout, nt, nsim] or null
    double *moments;           // [3, nsimout, nt] or null
    int nsimout;
};

__global__ void egdst_k_simulate(EgdstDev P, EgdstSimArgs S) {
    __shared__ double tile[EGDST_SIM_BLOCK / 32][32 * (EGDST_NSIMOUT_MAX | 1)];
    __shared__ double mom[3][EGDST_NSIMOUT_MAX];
    egdst_ctx cx; egdst_load_ctx(P, S.ivec, cx);
    cx.status = 0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nt = P.NT, nso = S.nsimout, tstride = nso | 1;  // odd stride: conflict-free record rows
    const double NaN = EGDST_NAN;
    const int ntiles = (S.nsim + 31) / 32;
    const int wpb = blockDim.x >> 5;
    // persistent: every warp strides over tiles of 32 agents; all warps of the CTA run the same number of
    // rounds so that the per-period moment flush can use block barriers
    const int rounds = (ntiles + gridDim.x * wpb - 1) / (gridDim.x * wpb);
    for (int round = 0; round < rounds; round++) {
        const int tileidx = (round * gridDim.x + blockIdx.x) * wpb + w;
        const int isim = tileidx * 32 + lane;
        const bool live_lane = tileidx < ntiles && isim < S.nsim;
        PeriodVars cur; cur.it = 0; cur.ist = 0; cur.id = 0; cur.cash = 0; cur.savings = 0; cur.shock = NaN;
        for (int i = 0; i < EGDST_NNST; i++) cur.st[i] = 0;
        for (int i = 0; i < EGDST_NND; i++) cur.dc[i] = 0;
        double mu = NaN, sigma = NaN, c = 0, vf = 0;
        double eqs[EGDST_NREQ > 0 ? EGDST_NREQ : 1];
        int state = live_lane ? 0 : 2;  // 0 alive, 1 dead/skipped (NaN rows), 2 no agent
        if (live_lane) {
            const int ist0 = (int)S.init[isim] - 1;
            const double m0 = S.init[S.nsim + isim];
            if (ist0 < 0 || ist0 >= cx.nst || m0 < cx.a0 || m0 > cx.mmax) state = 1;  // egdst_simulator.c:215-216
            else { cur.ist = ist0; cur.cash = m0; egdst_fill_state(&cx, &cur); if (!feasible(&cx, &cur)) state = 1; }
        }
        for (int it = 0; it < nt; it++) {
            if (state == 0 && it > 0) {
                PeriodVars nx = cur;
                nx.it = it;
                nx.savings = cur.savings;
                double rrr, rrr1, rrr2;
                if (S.randstream) {
                    const double *rs = S.randstream + (S.rndtype == 1 ? 0 : (size_t)4 * nt * isim) + (size_t)3 * (it - 1);
                    rrr = rs[0]; rrr1 = rs[1]; rrr2 = rs[2];
                } else {
                    const unsigned long long g = (unsigned long long)(S.agent0 + isim);
                    unsigned r4[4];
                    egdst_philox4x32((unsigned)g, (unsigned)(g >> 32), (unsigned)it, 0u, (unsigned)S.seed, (unsigned)(S.seed >> 32), r4);
                    rrr = egdst_u01(r4[0]); rrr1 = egdst_u01(r4[1]); rrr2 = egdst_u01(r4[2]);
                }
                if (rrr2 > survival(&cx, &cur)) {
                    state = 1;  // death: the rest of the record stays NaN
                } else {
                    int chosen = -1, lastfeas = -1;
                    for (int ist1 = 0; ist1 < cx.nst; ist1++) {
                        nx.ist = ist1;
                        egdst_fill_state(&cx, &nx);
                        if (!feasible(&cx, &nx)) continue;
                        lastfeas = ist1;
                        double pr;
                        if (cx.optim_TRPRnoSH == 1) pr = trpr(&cx, &cur, &nx, 0);
                        else {
                            mu = mu_param(&cx, &cur, &nx); sigma = sigma_param(&cx, &cur, &nx);
                            nx.shock = (sigma <= 0) ? egdst_expectation(&cx, &cur, &nx) : egdst_cdfinv(rrr1, mu, sigma);
                            pr = trpr(&cx, &cur, &nx, 0);
                        }
                        rrr -= pr;
                        if (rrr <= 0) { chosen = ist1; break; }
                    }
                    if (chosen < 0) chosen = lastfeas < 0 ? 0 : lastfeas;  // the reference runs off the end here
                    nx.ist = chosen;
                    egdst_fill_state(&cx, &nx);
                    if (cx.optim_TRPRnoSH == 1) {  // shocks for the shock-independent case (egdst_simulator.c:292-298)
                        mu = mu_param(&cx, &cur, &nx); sigma = sigma_param(&cx, &cur, &nx);
                        nx.shock = (sigma <= 0) ? egdst_expectation(&cx, &cur, &nx) : egdst_cdfinv(rrr1, mu, sigma);
                    }
                    nx.cash = cashinhand(&cx, &cur, &nx);
                    eqs_sim(&cx, &cur, &nx, eqs);
                    cur = nx;
                }
            } else if (state == 0) {
                eqs_sim(&cx, &cur, (const PeriodVars *)0, eqs);
            }
            if (state == 0) {
                // policy (egdst_simulator.c:145-199)
                const int cell = egdst_cell(P, S.ivec, it, cur.ist);
                const int nm = P.mlen[cell];
                if (nm < 2) { state = 1; }
                else {
                    const double *Mg = egdst_colM(P, cell), *Cg = egdst_colC(P, cell), *Vg = egdst_colV(P, cell);
                    const int i = egdst_bracket(cur.cash, Mg, nm, 0);
                    c = egdst_lerp(cur.cash, Mg[i], Mg[i + 1], Cg[i], Cg[i + 1]);
                    cur.savings = cur.cash - c;
                    const int nth = P.thlen[cell];
                    const double *th = P.thTH + (size_t)cell * cx.nthrhmax, *dd = P.thD + (size_t)cell * cx.nthrhmax;
                    int ith = 0;
                    while (ith < nth && cur.cash >= th[ith]) ith++;
                    cur.id = (int)dd[ith > 0 ? ith - 1 : 0];
                    egdst_fill_decision(&cx, &cur);
                    const double evf = Vg[0];
                    if (cur.cash < Mg[1] && evf > -EGDST_INF) vf = utility(&cx, &cur, c) + discount(&cx, &cur) * evf;
                    else vf = egdst_lerp(cur.cash, Mg[i], Mg[i + 1], Vg[i], Vg[i + 1]);
                }
            }
            // stage the record of this period
            double *rec = tile[w] + lane * tstride;
            if (state == 0) {
                rec[0] = cur.cash; rec[1] = c; rec[2] = cur.savings; rec[3] = vf; rec[4] = (double)cur.id; rec[5] = (double)cur.ist;
                rec[6] = mu; rec[7] = sigma; rec[8] = cur.shock; rec[9] = utility(&cx, &cur, c); rec[10] = discount(&cx, &cur);
                for (int i = 0; i < cx.nnst; i++) rec[11 + i] = cur.st[i];
                for (int i = 0; i < cx.nnd; i++) rec[11 + cx.nnst + i] = cur.dc[i];
                for (int i = 0; i < nso - 11 - cx.nnst - cx.nnd; i++) rec[11 + cx.nnst + cx.nnd + i] = eqs[i];
            } else {
                for (int j = 0; j < nso; j++) rec[j] = NaN;
            }
            __syncwarp();
            if (S.sims && tileidx < ntiles) {
                // cooperative write: element e of the tile belongs to agent e/nso, column e%nso
                const int nvalid = (S.nsim - tileidx * 32 < 32 ? S.nsim - tileidx * 32 : 32) * nso;
                for (int e = lane; e < nvalid; e += 32) {
                    const int a = e / nso, j = e - a * nso;
                    S.sims[((size_t)(tileidx * 32 + a) * nt + it) * nso + j] = tile[w][a * tstride + j];
                }
            }
            if (S.moments) {
                if (threadIdx.x < 3 * nso) (&mom[0][0])[(threadIdx.x / nso) * EGDST_NSIMOUT_MAX + threadIdx.x % nso] = 0.0;
                __syncthreads();
                // each lane reduces one column over the 32 agents of the tile (bank-conflict-free column walk)
                for (int j = lane; j < nso; j += 32) {
                    double s1 = 0, s2 = 0, n = 0;
                    for (int a = 0; a < 32; a++) { const double x = tile[w][a * tstride + j]; if (x == x) { s1 += x; s2 += x * x; n += 1; } }
                    atomicAdd(&mom[0][j], s1); atomicAdd(&mom[1][j], s2); atomicAdd(&mom[2][j], n);
                }
                __syncthreads();
                if (threadIdx.x < 3 * nso) {
                    const int k = threadIdx.x / nso, j = threadIdx.x % nso;
                    const double val = mom[k][j];
                    if (val != 0.0) atomicAdd(&S.moments[((size_t)it * nso + j) * 3 + k], val);
                }
                __syncthreads();
            }
            __syncwarp();
        }
    }
}
