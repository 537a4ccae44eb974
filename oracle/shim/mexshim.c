/* mexshim.c -- implementation of the MEX shim + a ctypes-friendly harness (TEST INFRASTRUCTURE ONLY).
 *
 * The three reference gateways are each called `mexFunction`; oracle/build_ref.py compiles them with
 * -DmexFunction=ref_solver_gateway / ref_simulator_gateway / ref_call_gateway and links them with this
 * file into oracle/_ref/<model>/libegdst_ref.so.  Python builds the fake model object with the
 * shim_* constructors, runs a gateway through shim_call() (which times it with clock_gettime, as
 * BASELINE.md section 4 prescribes) and reads the outputs back through the mx* accessors.
 */
#include <math.h>
#include <setjmp.h>
#include <time.h>
#include "mex.h"

void ref_solver_gateway(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
void ref_simulator_gateway(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
void ref_call_gateway(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);

static jmp_buf shim_jmp;
static int shim_jmp_armed = 0;
static char shim_err[1024];
static char shim_warn[4096];
static int shim_nwarn = 0;

static mxArray *newarr(int cls, size_t m, size_t n) {
    mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
    a->cls = cls; a->m = m; a->n = n;
    return a;
}

/* ---- constructors used from Python ---- */
mxArray *shim_double(size_t m, size_t n, const double *data) {
    mxArray *a = newarr(mxDOUBLE_CLASS, m, n);
    if (m * n > 0) {
        a->pr = (double *)malloc(m * n * sizeof(double));
        if (data) memcpy(a->pr, data, m * n * sizeof(double)); else memset(a->pr, 0, m * n * sizeof(double));
    }
    return a;
}
mxArray *shim_logical(int v) { mxArray *a = newarr(mxLOGICAL_CLASS, 1, 1); a->logical_val = v ? 1 : 0; return a; }
mxArray *shim_cell(size_t m, size_t n) {
    mxArray *a = newarr(mxCELL_CLASS, m, n);
    a->cells = (mxArray **)calloc(m * n ? m * n : 1, sizeof(mxArray *));
    return a;
}
mxArray *shim_struct(size_t nelem) { return newarr(mxSTRUCT_CLASS, nelem ? 1 : 0, nelem); }
void shim_setfield(mxArray *a, size_t idx, const char *name, mxArray *v) {
    int f; size_t nel = a->m * a->n;
    for (f = 0; f < a->nfields; f++) if (!strcmp(a->fnames[f], name)) break;
    if (f == a->nfields) {
        size_t e; int nf = a->nfields + 1;
        mxArray **nv = (mxArray **)calloc((nel ? nel : 1) * nf, sizeof(mxArray *));
        for (e = 0; e < nel; e++) memcpy(nv + e * nf, a->fvals + e * a->nfields, a->nfields * sizeof(mxArray *));
        free(a->fvals); a->fvals = nv;
        a->fnames = (char **)realloc(a->fnames, nf * sizeof(char *));
        a->fnames[f] = strdup(name); a->nfields = nf;
    }
    a->fvals[idx * a->nfields + f] = v;
}
void shim_free(mxArray *a) {
    size_t i;
    if (!a) return;
    if (a->cells) { for (i = 0; i < a->m * a->n; i++) shim_free(a->cells[i]); free(a->cells); }
    if (a->fvals) { for (i = 0; i < a->m * a->n * (size_t)a->nfields; i++) shim_free(a->fvals[i]); free(a->fvals); }
    if (a->fnames) { for (i = 0; i < (size_t)a->nfields; i++) free(a->fnames[i]); free(a->fnames); }
    free(a->pr); free(a);
}

/* ---- MEX API ---- */
mxArray *mxGetField(const mxArray *a, mwIndex idx, const char *name) {
    int f;
    if (!a) return NULL;
    for (f = 0; f < a->nfields; f++) if (!strcmp(a->fnames[f], name)) return a->fvals[idx * a->nfields + f];
    return NULL;
}
mxArray *mxGetProperty(const mxArray *a, mwIndex idx, const char *name) { return mxGetField(a, idx, name); }
double *mxGetPr(const mxArray *a) { return a ? a->pr : NULL; }
void *mxGetData(const mxArray *a) { return a ? (void *)a->pr : NULL; }
mxArray *mxGetCell(const mxArray *a, mwIndex idx) { return (a && a->cells && idx < a->m * a->n) ? a->cells[idx] : NULL; }
void mxSetCell(mxArray *a, mwIndex idx, mxArray *v) { a->cells[idx] = v; }
size_t mxGetM(const mxArray *a) { return a ? a->m : 0; }
size_t mxGetN(const mxArray *a) { return a ? a->n : 0; }
size_t mxGetNumberOfElements(const mxArray *a) { return a ? a->m * a->n : 0; }
double mxGetScalar(const mxArray *a) {
    if (!a) return 0.0;
    if (a->cls == mxLOGICAL_CLASS) return (double)a->logical_val;
    return a->pr ? a->pr[0] : 0.0;
}
int mxIsLogicalScalarTrue(const mxArray *a) { return a && a->cls == mxLOGICAL_CLASS && a->logical_val; }
double mxGetNaN(void) { return NAN; }
double mxGetInf(void) { return INFINITY; }
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) { (void)c; return shim_double(m, n, NULL); }
mxArray *mxCreateCellArray(mwSize ndim, const mwSize *dims) {
    size_t n = 1, i; for (i = 1; i < ndim; i++) n *= dims[i];
    return shim_cell(dims[0], n);
}
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity c) {
    size_t n = 1, i; (void)cls; (void)c; for (i = 1; i < ndim; i++) n *= dims[i];
    return shim_double(dims[0], n, NULL);
}
void mexWarnMsgTxt(const char *msg) {
    size_t used = strlen(shim_warn);
    shim_nwarn++;
    if (used + strlen(msg) + 2 < sizeof(shim_warn)) { strcat(shim_warn, msg); strcat(shim_warn, "\n"); }
}
void mexErrMsgTxt(const char *msg) {
    strncpy(shim_err, msg, sizeof(shim_err) - 1);
    if (shim_jmp_armed) longjmp(shim_jmp, 1);
    fprintf(stderr, "mexErrMsgTxt outside shim_call: %s\n", msg); abort();
}
int mexEvalString(const char *cmd) { (void)cmd; return 0; }
int mexCallMATLAB(int nlhs, mxArray *plhs[], int nrhs, mxArray *prhs[], const char *fn) {
    (void)nrhs; (void)prhs; (void)fn;
    if (nlhs > 0) plhs[0] = shim_double(1, 1, NULL); /* tic/toc only (VERBOSE builds) */
    return 0;
}

/* ---- harness ---- */
const char *shim_errmsg(void) { return shim_err; }
const char *shim_warnings(void) { return shim_warn; }
int shim_warncount(void) { return shim_nwarn; }
void shim_reset(void) { shim_err[0] = 0; shim_warn[0] = 0; shim_nwarn = 0; }

/* which: 0 solver, 1 simulator, 2 call.  Returns elapsed seconds around the gateway, <0 on mexErrMsgTxt. */
double shim_call(int which, int nlhs, mxArray **plhs, int nrhs, mxArray **prhs) {
    struct timespec t0, t1;
    shim_jmp_armed = 1;
    if (setjmp(shim_jmp)) { shim_jmp_armed = 0; return -1.0; }
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (which == 0) ref_solver_gateway(nlhs, plhs, nrhs, (const mxArray **)prhs);
    else if (which == 1) ref_simulator_gateway(nlhs, plhs, nrhs, (const mxArray **)prhs);
    else ref_call_gateway(nlhs, plhs, nrhs, (const mxArray **)prhs);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    shim_jmp_armed = 0;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
