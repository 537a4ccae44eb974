/* mex.h -- functional stand-in for MATLAB's MEX API (TEST INFRASTRUCTURE ONLY).
 *
 * MATLAB is not available in this image.  This shim implements just the MEX symbols that the
 * reference's egdst_lib.c / egdst_solver.c / egdst_simulator.c / egdst_call.c and the generated
 * modelspec.c use, so that those sources compile UNMODIFIED from /root/reference and
 * run as the parity oracle (oracle/_ref).  Nothing here is part of the product.
 */
#ifndef EGDST_SHIM_MEX_H
#define EGDST_SHIM_MEX_H

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxUNKNOWN_CLASS = 0, mxCELL_CLASS, mxSTRUCT_CLASS, mxLOGICAL_CLASS, mxDOUBLE_CLASS = 6 } mxClassID;

typedef struct mxArray_tag {
    int cls;                    /* mxClassID */
    size_t m, n;                /* rows, product of the remaining dims */
    double *pr;                 /* numeric payload (NULL when empty) */
    struct mxArray_tag **cells; /* cell payload, m*n entries */
    int nfields;                /* struct/object payload */
    char **fnames;
    struct mxArray_tag **fvals; /* fvals[idx*nfields + f] */
    int logical_val;
} mxArray;

mxArray *mxGetProperty(const mxArray *a, mwIndex idx, const char *name);
mxArray *mxGetField(const mxArray *a, mwIndex idx, const char *name);
double *mxGetPr(const mxArray *a);
void *mxGetData(const mxArray *a);
mxArray *mxGetCell(const mxArray *a, mwIndex idx);
void mxSetCell(mxArray *a, mwIndex idx, mxArray *v);
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a);
size_t mxGetNumberOfElements(const mxArray *a);
double mxGetScalar(const mxArray *a);
int mxIsLogicalScalarTrue(const mxArray *a);
double mxGetNaN(void);
double mxGetInf(void);
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray *mxCreateCellArray(mwSize ndim, const mwSize *dims);
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity c);
void mexWarnMsgTxt(const char *msg);
void mexErrMsgTxt(const char *msg);
int mexEvalString(const char *cmd);
int mexCallMATLAB(int nlhs, mxArray *plhs[], int nrhs, mxArray *prhs[], const char *fn);

#endif
