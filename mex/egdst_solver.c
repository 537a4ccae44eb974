/* egdst_solver.c (B200 build) -- MEX gateway  [M, D, dbgout] = egdst_solver(model).
 * Drop-in for @egdstmodel/egdst_solver.c:143-239: same name, arity checks and output layouts
 * (M{ist,it}: rows x 4 = M,C,A,V with the a0 row first; D{ist,it}: rows x 2 = decision index, threshold;
 * saveoutput, egdst_solver.c:917-951); empty cells for infeasible (it,ist).  Soft errors become a warning
 * with the partial result returned (egdst_solver.c:237), hard errors mexErrMsgTxt. */
#include "egdst_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    egdst_desc d;
    egdst_solution *sol = NULL;
    int rc, nst, nt, c, *mlen, *thlen;
    size_t nm = 0, nth = 0, om = 0, oth = 0;
    double *Mbuf, *Dbuf;
    mwSize dims[2];
    if (nrhs != 1) mexErrMsgTxt("Error: wrong number of input arguments!");   /* egdst_solver.c:149 */
    if (nlhs != 3) mexErrMsgTxt("Error: wrong number of output arguments!");  /* egdst_solver.c:150-154 */
    egdst_mex_desc(prhs[0], &d, 1);
    nst = d.nst; nt = d.T - d.t0 + 1;
    dims[0] = (mwSize)nst; dims[1] = (mwSize)nt;
    plhs[0] = mxCreateCellArray(2, dims);
    plhs[1] = mxCreateCellArray(2, dims);
    plhs[2] = mxCreateDoubleMatrix(0, 7, mxREAL); /* dbgout: only filled by DEBUGOUT builds of the reference (:176-182) */
    rc = egdst_solve(&d, &sol);
    if (rc == 2 || !sol) { mexErrMsgTxt(egdst_last_error()); return; }
    if (rc == 1) mexWarnMsgTxt(egdst_last_error());
    mlen = (int *)calloc((size_t)nst * nt, sizeof(int));
    thlen = (int *)calloc((size_t)nst * nt, sizeof(int));
    egdst_solution_sizes(sol, mlen, thlen);
    for (c = 0; c < nst * nt; c++) { nm += 4 * (size_t)mlen[c]; nth += 2 * (size_t)thlen[c]; }
    Mbuf = (double *)malloc((nm ? nm : 1) * sizeof(double));
    Dbuf = (double *)malloc((nth ? nth : 1) * sizeof(double));
    if (egdst_solution_export(sol, Mbuf, Dbuf)) { egdst_free_solution(sol); mexErrMsgTxt(egdst_last_error()); return; }
    for (c = 0; c < nst * nt; c++) {
        mxArray *cm, *cd;
        if (!mlen[c]) continue;
        cm = mxCreateDoubleMatrix((mwSize)mlen[c], 4, mxREAL);
        cd = mxCreateDoubleMatrix((mwSize)thlen[c], 2, mxREAL);
        memcpy(mxGetPr(cm), Mbuf + om, 4 * (size_t)mlen[c] * sizeof(double)); om += 4 * (size_t)mlen[c];
        memcpy(mxGetPr(cd), Dbuf + oth, 2 * (size_t)thlen[c] * sizeof(double)); oth += 2 * (size_t)thlen[c];
        mxSetCell(plhs[0], c, cm);
        mxSetCell(plhs[1], c, cd);
    }
    free(mlen); free(thlen); free(Mbuf); free(Dbuf);
    egdst_free_solution(sol);
}
