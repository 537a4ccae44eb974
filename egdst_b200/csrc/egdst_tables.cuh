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

// build the tables of the cells (ivec, it, all ist): grid (nblk, nst, nvec)
__global__ void egdst_k_tab(EgdstDev P, int it) {
    const int ivec = blockIdx.z, ist = blockIdx.y;
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
    EgdstLutEntry *L = P.tabLut + (size_t)cell * (P.lutcap + 1);
    for (int b = t0; b <= P.lutcap; b += stride) {
        // first row whose key is >= b, and >= b+1 (keys are non-decreasing along the grid)
        int lo = 0, hi = n, lo2 = 0, hi2 = n;
        if (b == P.lutcap) { lo = n; lo2 = n; }
        else {
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (egdst_lut_key(M[mid], a0, P.mbits) < b) lo = mid + 1; else hi = mid; }
            if (b + 1 == P.lutcap) lo2 = n;
            else { lo2 = lo; while (lo2 < hi2) { const int mid = (lo2 + hi2) >> 1; if (egdst_lut_key(M[mid], a0, P.mbits) < b + 1) lo2 = mid + 1; else hi2 = mid; } }
        }
        EgdstLutEntry e; e.l = lo; e.cnt = lo2 - lo; e.m = lo < n ? M[lo] : EGDST_INF;
        L[b] = e;
    }
}

// Bracket of x in the cell's grid.  Same result as egdst_bracket(x, M, n, 0) on a strictly increasing grid:
// (#rows <= x) - 1 clamped to [0, n-2].
EGDST_DEV int egdst_bracket_tab(const EgdstDev &P, int cell, double x, int n) {
    const double *M = egdst_colM(P, cell);
    if (!egdst_cell_has_tab(P, n)) return egdst_bracket(x, M, n, 0);
    int b = egdst_lut_key(x, P.cx.a0, P.mbits);
    b = b < 0 ? 0 : (b > P.lutcap - 1 ? P.lutcap - 1 : b);
    const EgdstLutEntry *lut = egdst_cell_lut(P, cell);
#ifdef EGDST_HOSTEMU
    const EgdstLutEntry e = lut[b];
#else
    const int4 raw = *reinterpret_cast<const int4 *>(lut + b);  // one 16-byte load
    EgdstLutEntry e; e.l = raw.x; e.cnt = raw.y; e.m = __hiloint2double(raw.w, raw.z);
#endif
    int cnt = e.l + ((e.cnt > 0 && e.m <= x) ? 1 : 0);
    if (e.cnt > 1 && e.m <= x) {  // crowded bucket (double points, coarse tables): bisect its remaining rows
        int l = e.l + 1, h = e.l + e.cnt;
        while (l < h) { const int mid = (l + h) >> 1; if (M[mid] <= x) l = mid + 1; else h = mid; }
        cnt = l;
    }
    int i = cnt - 1;
    if (i > n - 2) i = n - 2;
    return i < 0 ? 0 : i;
}
