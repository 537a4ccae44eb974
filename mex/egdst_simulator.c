/* egdst_simulator.c (B200 build) -- MEX gateway  sims = egdst_simulator(model, rndtype).
 * Drop-in for @egdstmodel/egdst_simulator.c:47-117: reads init, randstream, M, D from the object, checks the
 * randstream length (:71-75) and returns sims[nsimout, nt, nsim], NaN after death (:95-105). */
#include "egdst_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    egdst_desc d;
    egdst_solution *sol;
    const mxArray *init, *rs;
    int rndtype, nsim, nt, nsimout, rc;
    mwSize dims[3];
    if (nrhs != 2) mexErrMsgTxt("Error: wrong number of input arguments!");   /* egdst_simulator.c:54 */
    if (nlhs != 1) mexErrMsgTxt("Error: wrong number of output arguments!");
    rndtype = (int)mxGetScalar(prhs[1]);
    egdst_mex_desc(prhs[0], &d, 0);
    init = mxGetProperty(prhs[0], 0, "init");
    rs = mxGetProperty(prhs[0], 0, "randstream");
    if (!init || !rs) mexErrMsgTxt("Error: init or randstream is missing!");
    nsim = (int)mxGetM(init);
    nt = d.T - d.t0 + 1;
    nsimout = 11 + d.nnst + d.nnd + d.neq;                                     /* egdst_simulator.c:95 */
    dims[0] = (mwSize)nsimout; dims[1] = (mwSize)nt; dims[2] = (mwSize)nsim;
    plhs[0] = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
    sol = egdst_mex_import(prhs[0], &d);
    if (!sol) return;
    rc = egdst_simulate(&d, sol, 0, mxGetPr(init), nsim, mxGetPr(rs), (long long)mxGetNumberOfElements(rs), rndtype, mxGetPr(plhs[0]));
    egdst_free_solution(sol);
    if (rc == 2) mexErrMsgTxt(egdst_last_error());
    if (rc == 1) mexWarnMsgTxt(egdst_last_error());
}
