/* egdst_call.c (B200 build) -- MEX gateway  res = egdst_call(model, sw, args).
 * Drop-in for @egdstmodel/egdst_call.c:17-125: sw = 1 utility, 2 marginal utility, 3 discount, 4 budget,
 * 5 marginal budget, 6 value function; args is narg x k (k = 4,4,2,6,6,3); res is narg x 1. */
#include "egdst_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    egdst_desc d;
    egdst_solution *sol;
    int sw, narg, k, rc;
    if (nrhs != 3) mexErrMsgTxt("Error: wrong number of input arguments!");
    if (nlhs > 1) mexErrMsgTxt("Error: wrong number of output arguments!");
    sw = (int)mxGetScalar(prhs[1]);
    narg = (int)mxGetM(prhs[2]);
    k = (int)mxGetN(prhs[2]);
    egdst_mex_desc(prhs[0], &d, 0);
    plhs[0] = mxCreateDoubleMatrix((mwSize)narg, 1, mxREAL);
    sol = egdst_mex_import(prhs[0], &d);
    if (!sol) return;
    rc = egdst_call(&d, sol, sw, mxGetPr(prhs[2]), narg, k, mxGetPr(plhs[0]));
    egdst_free_solution(sol);
    if (rc == 2) mexErrMsgTxt(egdst_last_error());
    if (rc == 1) mexWarnMsgTxt(egdst_last_error());
}
