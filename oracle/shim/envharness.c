/* envharness.c -- TEST INFRASTRUCTURE ONLY: the reference's envelope routines, callable on given points.
 *
 * envelope2() and envelop() (@egdstmodel/egdst_solver.c:776-913, 1165-1550) are `static` inside the solver's
 * translation unit, so this file INCLUDES the unmodified reference source from where it lies under /root/reference
 * (-I<reference>/@egdstmodel; nothing is copied into the repo) and adds two entry points that set the reference's
 * globals from a model object exactly as its gateway does (egdst_solver.c:157-159) and then call the routines.
 * Built by oracle/ref.py:build_harness into oracle/_ref/<key>_env/libegdst_refenv.so; used by the differential tests
 * of the upper-envelope kernels (exact ties, flat stretches, many runs).
 */
#define mexFunction ref_solver_gateway
#include "egdst_solver.c"

/* the shim's dispatcher references the other two gateways; they are not part of this library */
void ref_simulator_gateway(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) { (void)nlhs; (void)plhs; (void)nrhs; (void)prhs; }
void ref_call_gateway(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) { (void)nlhs; (void)plhs; (void)nrhs; (void)prhs; }

static void harness_setup(const mxArray *model) {
    Model = (mxArray *)model;
    parseModel();
    loadparameters();
    err[0] = 0;
}

/* secondary envelope of one decision's points: pts = quadruples (M, C, V, id) with room for 4*ngridmax doubles;
 * returns the number of points kept (the first *nout quadruples of pts), -1 on error (message in errbuf) */
int ref_envelope2(const mxArray *model, int it, int ist, int id, double *pts, int nvd, double evfa0_id, char *errbuf, int errlen) {
    harness_setup(model);
    PeriodVars curr;
    memset(&curr, 0, sizeof(curr));
    curr.it = it; curr.ist = ist; curr.id = id;
    double ev = evfa0_id;
    int skip = envelope2(pts, nvd, &ev, &curr, NULL);
    if (err[0]) { strncpy(errbuf, err, errlen - 1); errbuf[errlen - 1] = 0; return -1; }
    return nvd - skip;
}

/* primary envelope over nfun functions: gridvecs = dim0 quadruples (x, C, V, function index), the layout of mgridvecs
 * (egdst_solver.c:1250-1253: the 3rd element is the compared function, the 2nd rides along); outputs sized by the
 * caller (ngridmax / nthrhmax): outfunc = V, outfunc2 = C */
int ref_envelop(const mxArray *model, int it, int ist, int nfun, int dim0, double *gridvecs, double *evfa0,
                double *outgrid, double *outfunc, double *outfunc2, double *outthrh, double *outindx, int *outn, int *outm,
                char *errbuf, int errlen) {
    harness_setup(model);
    envelop(it, ist, nfun, dim0, gridvecs, evfa0, outgrid, outfunc, outfunc2, outthrh, outindx, outn, outm, NULL);
    if (err[0]) { strncpy(errbuf, err, errlen - 1); errbuf[errlen - 1] = 0; return -1; }
    return 0;
}
