// egdst_tables.cuh -- per-cell lookup tables shared by the solver (period t reads the tables of t+1) and the
// simulator.  Built by egdst_k_tab right after a cell's primary envelope is final.
//
// Every table access of the hot loops is a per-thread gather, so the layout is chosen to make one policy/value
// lookup cost four 16-byte loads (neighbouring A points / agents with neighbouring cash hit the same lines):
//   lut [ncell][lutcap+1] EgdstLutEntry  direct index by the leading bits of the IEEE representation of
//        (x - a0 + 1): key = exponent and the top `mbits` mantissa bits, a piecewise-linear log2 -- the endogenous
//        grids are (sym-)log spaced (egdst_solver.c:1104-1136), so buckets are evenly filled.  Entry b holds
//        l = #rows with key < b, the number of rows in the bucket and the abscissa of its first row: with about
//        half a row per bucket, #rows <= x is l + (m <= x) without touching the grid (crowded buckets bisect).
//   ivl [ncell][tabcap] EgdstInterval    (M_i, M_i+1, C_i, C_i+1, V_i, V_i+1): everything the interpolations of the
//        EGM step (egdst_solver.c:552-567, 755-772) and of policy() (egdst_simulator.c:178-197) need, contiguous.
// This replaces the 14-step bisections of bxsearch (egdst_lib.c:138-165), 2 per quadrature node and agent-period.
// Cells with more rows than tabcap+1 (possible only when ngridmax >> 2*ngridm is actually used) keep the bisection.
#pragma once

#include "egdst_common.cuh"

EGDST_DEV int egdst_lut_key(double x, double a0, int mbits) {
    const double y = x - a0 + 1.0;
#ifdef EGDST_HOSTEMU
    long long bits; memcpy(&bits, &y, 8);
    const int hi = (int)(bits >> 32);
#else
    const int hi = __double2hiint(y);
#endif
    return (hi >> (20 - mbits)) - (0x3FF00000 >> (20 - mbits));
}

EGDST_DEV const EgdstInterval *egdst_cell_ivl(const EgdstDev &P, int cell) { return P.tabIvl + (size_t)cell * P.tabcap; }
EGDST_DEV const EgdstLutEntry *egdst_cell_lut(const EgdstDev &P, int cell) { return P.tabLut + (size_t)cell * (P.lutcap + 1); }
EGDST_DEV bool egdst_cell_has_tab(const EgdstDev &P, int n) { return n - 1 <= P.tabcap; }

// L2 eviction-priority hints (PTX createpolicy / ld.global.L2::cache_hint): the simulator streams tens of GB of
// output through L2; table lines loaded with evict_last survive that stream, the output is written evict_first.
#ifndef EGDST_HOSTEMU
EGDST_DEV unsigned long long egdst_policy_evict_last() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
EGDST_DEV double2 egdst_ld16_hint(const void *p, unsigned long long pol) {
    double2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
}
#endif

EGDST_DEV EgdstInterval egdst_load_interval(const EgdstInterval *p) {
#ifdef EGDST_HOSTEMU
    return *p;
#else
    const double2 *q = reinterpret_cast<const double2 *>(p);  // three 16-byte loads
    const double2 q0 = q[0], q1 = q[1], q2 = q[2];
    EgdstInterval iv; iv.g0 = q0.x; iv.g1 = q0.y; iv.c0 = q1.x; iv.c1 = q1.y; iv.v0 = q2.x; iv.v1 = q2.y;
    return iv;
#endif
}

// Start-of-period housekeeping for parameter vector ivec (all threads of one CTA): reset the chained-scan state of
// its jobs, mark infeasible (it,ist) cells as empty and detect empty choice sets (egdst_solver.c:294-300, 694-702).
EGDST_DEV void egdst_cells_body(const EgdstDev &P, int ivec, int it) {
    const int nsdv = P.cx.nst * P.cx.nd, sd0 = ivec * nsdv;
    for (int i = threadIdx.x; i < nsdv * P.chC; i += blockDim.x) P.scanC[(size_t)sd0 * P.chC + i] = 0ULL;
    for (int i = threadIdx.x; i < nsdv * 2; i += blockDim.x) P.tickC[2 * sd0 + i] = 0;
    for (int i = threadIdx.x; i < nsdv; i += blockDim.x) P.foldCnt[sd0 + i] = 0;
    for (int i = threadIdx.x; i < nsdv * P.chE; i += blockDim.x) P.scanE[(size_t)sd0 * P.chE + i] = 0ULL;          // secondary slots
    for (int i = threadIdx.x; i < nsdv * 2; i += blockDim.x) P.tickE[2 * sd0 + i] = 0;
    for (int i = threadIdx.x; i < nsdv; i += blockDim.x) P.envNact[sd0 + i] = 0;
    const int ps0 = P.nvec * nsdv + ivec * P.cx.nst;                                                                // primary slots
    for (int i = threadIdx.x; i < P.cx.nst * P.chE; i += blockDim.x) P.scanE[(size_t)ps0 * P.chE + i] = 0ULL;
    for (int i = threadIdx.x; i < P.cx.nst * 2; i += blockDim.x) P.tickE[2 * ps0 + i] = 0;
    for (int i = threadIdx.x; i < P.cx.nst; i += blockDim.x) P.envNact[ps0 + i] = 0;
    for (int ist = threadIdx.x; ist < P.cx.nst; ist += blockDim.x) {
        egdst_ctx cx; egdst_load_ctx(P, ivec, cx);
        PeriodVars curr; curr.it = it; curr.ist = ist; curr.id = 0; curr.cash = 0; curr.savings = 0; curr.shock = 0;
        const int cell = egdst_cell(P, ivec, it, ist);
        if (feasible(&cx, &curr) != 1) { P.mlen[cell] = 0; P.thlen[cell] = 0; continue; }
        int any = 0;
        for (curr.id = 0; curr.id < cx.nd; curr.id++) any |= (inchoiceset(&cx, &curr) == 1);
        if (!any) { P.mlen[cell] = 0; P.thlen[cell] = 0; egdst_fail(P, ivec, EGDST_ERR_EMPTYCHOICE, it, ist, -1); }
    }
}

// build the tables of the cells (ivec, it, all ist): grid (nblk, nst, nvec)
__global__ void egdst_k_tab(EgdstDev P, int it) {
    const int ivec = blockIdx.z, ist = blockIdx.y;
    // the last kernel of period `it` also opens period it-1 (saves a launch per period)
    if (blockIdx.x == 0 && blockIdx.y == 0 && it > 0) egdst_cells_body(P, ivec, it - 1);
    const int cell = egdst_cell(P, ivec, it, ist);
    const int n = P.mlen[cell];
    if (n < 2 || !egdst_cell_has_tab(P, n)) return;
    const double a0 = P.cx.a0;
    const double *M = egdst_colM(P, cell), *C = egdst_colC(P, cell), *V = egdst_colV(P, cell);
    EgdstInterval *r = P.tabIvl + (size_t)cell * P.tabcap;
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = t0; i + 1 < n; i += stride) {
        EgdstInterval v; v.g0 = M[i]; v.g1 = M[i + 1]; v.c0 = C[i]; v.c1 = C[i + 1]; v.v0 = V[i]; v.v1 = V[i + 1];
        r[i] = v;
    }
    // direct index, row-centric (no searches): row r is the first row with key >= b for every bucket b in
    // (key(r-1), key(r)]; the bucket key(r) also gets the number of rows that share the key
    EgdstLutEntry *L = P.tabLut + (size_t)cell * (P.lutcap + 1);
    const int kmax = P.lutcap - 1;
    const int lane = threadIdx.x & 31;
    for (int base = blockIdx.x * blockDim.x; base < n; base += stride) {  // warp-uniform trip count
        const int r = base + threadIdx.x;
        int gap0 = 0, gap1 = 0;
        EgdstLutEntry e; e.l = r; e.cnt = 0; e.m = 0.0;
        if (r < n) {
            const double m = M[r];
            int kr = egdst_lut_key(m, a0, P.mbits); kr = kr < 0 ? 0 : (kr > kmax ? kmax : kr);
            int kp = -1;
            if (r > 0) { kp = egdst_lut_key(M[r - 1], a0, P.mbits); kp = kp < 0 ? 0 : (kp > kmax ? kmax : kp); }
            if (kr > kp) {
                int cnt = 1;
                while (r + cnt < n) { int kn = egdst_lut_key(M[r + cnt], a0, P.mbits); kn = kn < 0 ? 0 : (kn > kmax ? kmax : kn); if (kn != kr) break; cnt++; }
                e.m = m; e.cnt = cnt;
                L[kr] = e;
                e.cnt = 0;
                gap0 = kp + 1; gap1 = kr;  // empty buckets [gap0, gap1) point at row r too
            }
        }
        if (gap1 - gap0 <= 8) { for (int bb = gap0; bb < gap1; bb++) L[bb] = e; gap1 = gap0; }
        // long runs of empty buckets (coarse stretches of a sym-log grid) are filled by the whole warp
        unsigned longm = __ballot_sync(EGDST_FULL, gap1 > gap0);
        while (longm) {
            const int src = __ffs(longm) - 1;
            longm &= longm - 1;
            const int b0 = __shfl_sync(EGDST_FULL, gap0, src), b1 = __shfl_sync(EGDST_FULL, gap1, src);
            EgdstLutEntry f; f.l = __shfl_sync(EGDST_FULL, e.l, src); f.cnt = 0; f.m = __shfl_sync(EGDST_FULL, e.m, src);
            for (int bb = b0 + lane; bb < b1; bb += 32) L[bb] = f;
        }
    }
    {   // buckets above the last row's key
        int kl = egdst_lut_key(M[n - 1], a0, P.mbits); kl = kl < 0 ? 0 : (kl > kmax ? kmax : kl);
        EgdstLutEntry e; e.l = n; e.cnt = 0; e.m = EGDST_INF;
        for (int bb = kl + 1 + t0; bb <= P.lutcap; bb += stride) L[bb] = e;
    }
}

// the simulator's versions: same lookups with the evict_last hint
EGDST_DEV EgdstInterval egdst_load_interval_keep(const EgdstInterval *p, unsigned long long pol) {
#ifdef EGDST_HOSTEMU
    return *p;
#else
    const double2 q0 = egdst_ld16_hint(p, pol), q1 = egdst_ld16_hint((const double2 *)p + 1, pol), q2 = egdst_ld16_hint((const double2 *)p + 2, pol);
    EgdstInterval iv; iv.g0 = q0.x; iv.g1 = q0.y; iv.c0 = q1.x; iv.c1 = q1.y; iv.v0 = q2.x; iv.v1 = q2.y;
    return iv;
#endif
}

// Bracket AND interval record of x in a cell that has tables (egdst_cell_has_tab).  Same bracket as
// egdst_bracket(x, M, n, 0) on a strictly increasing grid: (#rows <= x) - 1 clamped to [0, n-2].  KEEP: load with the
// evict_last hint `pol`.
// A bucket that holds two rows (about a quarter of the lookups at two buckets per row) is resolved by fetching the
// records of both rows at once -- two independent gathers -- instead of a dependent read of the grid column followed
// by the record: the second row's abscissa is g1 of the first row's record.  Three and more rows bisect as before.
template <bool KEEP = false>
EGDST_DEV int egdst_lookup_tab(const EgdstDev &P, int cell, const EgdstInterval *ivl, double x, int n, EgdstInterval &iv,
                               unsigned long long pol = 0ULL) {
    int b = egdst_lut_key(x, P.cx.a0, P.mbits);
    b = b < 0 ? 0 : (b > P.lutcap - 1 ? P.lutcap - 1 : b);
    const EgdstLutEntry *lut = egdst_cell_lut(P, cell);
#ifdef EGDST_HOSTEMU
    const EgdstLutEntry e = lut[b];
#else
    EgdstLutEntry e;
    if (KEEP) {
        const double2 raw = egdst_ld16_hint(lut + b, pol);
        const long long lo = __double_as_longlong(raw.x);
        e.l = (int)(lo & 0xffffffffLL); e.cnt = (int)(lo >> 32); e.m = raw.y;
    } else {
        const int4 raw = *reinterpret_cast<const int4 *>(lut + b);
        e.l = raw.x; e.cnt = raw.y; e.m = __hiloint2double(raw.w, raw.z);
    }
#endif
    const bool in = e.cnt > 0 && e.m <= x;
    int i = e.l - 1 + (in ? 1 : 0);
    if (in && e.cnt > 2) {
        const double *M = egdst_colM(P, cell);
        int l = e.l + 1, h = e.l + e.cnt;
        while (l < h) { const int mid = (l + h) >> 1; if (M[mid] <= x) l = mid + 1; else h = mid; }
        i = l - 1;
    }
    i = i > n - 2 ? n - 2 : i;
    i = i < 0 ? 0 : i;
    const bool two = in && e.cnt == 2 && i + 1 <= n - 2;
    const int i2 = two ? i + 1 : i;
    EgdstInterval iv2;
    if (KEEP) { iv = egdst_load_interval_keep(ivl + i, pol); iv2 = two ? egdst_load_interval_keep(ivl + i2, pol) : iv; }
    else { iv = egdst_load_interval(ivl + i); iv2 = two ? egdst_load_interval(ivl + i2) : iv; }
    if (two && x >= iv.g1) { iv = iv2; i = i2; }
    return i;
}
