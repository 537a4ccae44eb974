/* egdst_mex_common.h -- shared part of the three thin MEX gateways of the B200 build.
 *
 * The gateways keep the reference's MEX names, arities and output layouts
 *   [M, D, dbgout] = egdst_solver(model)         @egdstmodel/egdst_solver.c:143-239
 *   sims = egdst_simulator(model, rndtype)       @egdstmodel/egdst_simulator.c:47-117
 *   res  = egdst_call(model, sw, args)           @egdstmodel/egdst_call.c:17-125
 * so that @egdstmodel/egdstmodel.m runs unchanged (it calls them at :1170, :1268, :1190-1200).  All they
 * do is flatten the model object into an egdst_desc (what parseModel/loadparameters did,
 * egdst_lib.c:34-62, compile.m:469-474) and move arrays between mxArrays and the C ABI of
 * include/egdst_b200.h; the numerical work happens in the per-model CUDA library they link against
 * (libegdst_b200_<key>.so, built by nvcc from the same exec strings compile.m used to turn into modelspec.c).
 *
 * TOLERANCE / ZEROCONSUMPTION / DOUBLEPOINT_DELTA arrive as -D flags exactly as compile.m:757-777 passes
 * them to mex; the defaults are those of egdstmodel.m:420-424.
 */
#ifndef EGDST_MEX_COMMON_H
#define EGDST_MEX_COMMON_H

#include <string.h>
#include "mex.h"
#include "egdst_b200.h"

#ifndef TOLERANCE
#define TOLERANCE 1e-10
#endif
#ifndef ZEROCONSUMPTION
#define ZEROCONSUMPTION 1e-10
#endif
#ifndef DOUBLEPOINT_DELTA
#define DOUBLEPOINT_DELTA 1e-10
#endif
#ifndef EGDST_DEVICE
#define EGDST_DEVICE 0
#endif

#define EGDST_MEX_MAXPARAM 256

static double egdst_mex_params[EGDST_MEX_MAXPARAM];

static double egdst_mex_scalar(const mxArray *model, const char *name) {
    const mxArray *p = mxGetProperty(model, 0, name);
    if (!p) { mexErrMsgTxt("egdst: the model object lacks a required property"); return 0.0; }
    return mxGetScalar(p);
}

/* model object -> POD descriptor (the properties parseModel reads, egdst_lib.c:37-55) */
static void egdst_mex_desc(const mxArray *model, egdst_desc *d, int need_quadrature) {
    const mxArray *optim, *param, *q;
    int i;
    memset(d, 0, sizeof(*d));
    d->abi_version = EGDST_ABI_VERSION;
    d->t0 = (int)egdst_mex_scalar(model, "t0");
    d->T = (int)egdst_mex_scalar(model, "T");
    d->ngridm = (int)egdst_mex_scalar(model, "ngridm");
    d->ngridmax = (int)egdst_mex_scalar(model, "ngridmax");
    d->nthrhmax = (int)egdst_mex_scalar(model, "nthrhmax");
    d->ny = (int)egdst_mex_scalar(model, "ny");
    d->nd = (int)egdst_mex_scalar(model, "nd");
    d->nnd = (int)egdst_mex_scalar(model, "nnd");
    d->nst = (int)egdst_mex_scalar(model, "nst");
    d->nnst = (int)egdst_mex_scalar(model, "nnst");
    d->mmax = egdst_mex_scalar(model, "mmax");
    d->a0 = egdst_mex_scalar(model, "a0");
    optim = mxGetProperty(model, 0, "optim");
    d->optim_UasD = mxIsLogicalScalarTrue(mxGetField(optim, 0, "optim_UasD"));
    d->optim_MUnoD = mxIsLogicalScalarTrue(mxGetField(optim, 0, "optim_MUnoD"));
    d->optim_UnoD = mxIsLogicalScalarTrue(mxGetField(optim, 0, "optim_UnoD"));
    d->optim_TRPRnoSH = mxIsLogicalScalarTrue(mxGetField(optim, 0, "optim_TRPRnoSH"));
    d->tolerance = TOLERANCE;
    d->zeroconsumption = ZEROCONSUMPTION;
    d->doublepoint_delta = DOUBLEPOINT_DELTA;
    d->stm = mxGetPr(mxGetProperty(model, 0, "stm"));
    d->states = mxGetPr(mxGetProperty(model, 0, "states"));
    d->decisions = mxGetPr(mxGetProperty(model, 0, "decisions"));
    param = mxGetProperty(model, 0, "param");
    d->nparam = (int)mxGetNumberOfElements(param);
    if (d->nparam > EGDST_MEX_MAXPARAM) mexErrMsgTxt("egdst: too many parameters");
    for (i = 0; i < d->nparam; i++) egdst_mex_params[i] = mxGetScalar(mxGetField(param, i, "value")); /* compile.m:472 */
    d->params = egdst_mex_params;
    d->quadrature = NULL;
    if (need_quadrature && d->ny > 1) {
        q = mxGetProperty(model, 0, "quadrature"); /* [ny x 2]: weights, abscissas in (0,1) (egdstmodel.m:1157-1160) */
        if (!q || mxGetNumberOfElements(q) < (size_t)(2 * d->ny)) mexErrMsgTxt("egdst: model.quadrature is missing or too short");
        d->quadrature = mxGetPr(q);
    }
    d->neq = (int)mxGetNumberOfElements(mxGetProperty(model, 0, "eq"));
    d->device = EGDST_DEVICE;
}

/* model.M / model.D cells -> device solution (the simulator and call gateways receive them from MATLAB) */
static egdst_solution *egdst_mex_import(const mxArray *model, const egdst_desc *d) {
    const mxArray *M = mxGetProperty(model, 0, "M"), *D = mxGetProperty(model, 0, "D");
    const int nst = d->nst, nt = d->T - d->t0 + 1;
    int *mlen, *thlen, c, rc;
    size_t nm = 0, nth = 0, om = 0, oth = 0;
    double *Mbuf, *Dbuf;
    egdst_solution *sol = NULL;
    if (!M || !D || mxGetNumberOfElements(M) < (size_t)(nst * nt)) { mexErrMsgTxt("egdst: the model needs to be solved first (M, D are empty)"); return NULL; }
    mlen = (int *)calloc((size_t)nst * nt, sizeof(int));
    thlen = (int *)calloc((size_t)nst * nt, sizeof(int));
    for (c = 0; c < nst * nt; c++) {           /* C-ABI cell order is it*nst+ist == MATLAB's column-major {ist,it} */
        const mxArray *cm = mxGetCell(M, c), *cd = mxGetCell(D, c);
        mlen[c] = cm ? (int)mxGetM(cm) : 0;
        thlen[c] = (cd && mlen[c]) ? (int)mxGetM(cd) : 0;
        nm += 4 * (size_t)mlen[c]; nth += 2 * (size_t)thlen[c];
    }
    Mbuf = (double *)malloc((nm ? nm : 1) * sizeof(double));
    Dbuf = (double *)malloc((nth ? nth : 1) * sizeof(double));
    for (c = 0; c < nst * nt; c++) {
        if (!mlen[c]) continue;
        memcpy(Mbuf + om, mxGetPr(mxGetCell(M, c)), 4 * (size_t)mlen[c] * sizeof(double)); om += 4 * (size_t)mlen[c];
        memcpy(Dbuf + oth, mxGetPr(mxGetCell(D, c)), 2 * (size_t)thlen[c] * sizeof(double)); oth += 2 * (size_t)thlen[c];
    }
    rc = egdst_solution_import(d, mlen, thlen, Mbuf, Dbuf, &sol);
    free(mlen); free(thlen); free(Mbuf); free(Dbuf);
    if (rc) { mexErrMsgTxt(egdst_last_error()); return NULL; }
    return sol;
}

#endif
