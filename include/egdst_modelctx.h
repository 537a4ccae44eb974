/* egdst_modelctx.h -- per-instance model context shared by the generated model header
 * (modelspec_dev.h), the CUDA kernels (egdst_b200/csrc) and the CPU restatement (oracle/).
 *
 * Replaces the reference's process globals (egdst_lib.c:10-31: t0,T,ngridm,...,states,decisions,
 * byval) and the parameter globals that compile.m emits (compile.m:223-228): everything a generated
 * model function may read lives in one POD that is passed explicitly, so that several parameter
 * vectors can be solved concurrently (batched sweeps) and the code is re-entrant.
 *
 * EGDST_NNST / EGDST_NND / EGDST_NPARAM are defined by the generated header before this include.
 */
#ifndef EGDST_MODELCTX_H
#define EGDST_MODELCTX_H

#include <math.h>

#ifdef __CUDACC__
#define EGDST_FN static __device__ __forceinline__
#define EGDST_CONST static __device__ const
#define EGDST_HD __host__ __device__
#else
#define EGDST_FN static inline
#define EGDST_CONST static const
#define EGDST_HD
#endif

#ifndef MAX
#define MAX(X, Y) (((X) > (Y)) ? (X) : (Y))
#endif
#ifndef MIN
#define MIN(X, Y) (((X) < (Y)) ? (X) : (Y))
#endif
#ifndef __cplusplus
#ifndef true
#define true 1
#define false 0
#endif
#endif

#define EGDST_NAN (nan(""))
#define EGDST_INF (INFINITY)

#ifndef EGDST_NPARAM
#define EGDST_NPARAM 0
#endif
#define EGDST_NPARAM_ (EGDST_NPARAM > 0 ? EGDST_NPARAM : 1)

/* status codes written through cx->status (soft = partial result kept, as mexWarnMsgTxt;
 * hard = as mexErrMsgTxt).  Texts are in egdst_capi.cu / oracle, same wording as the reference. */
enum {
    EGDST_OK = 0,
    EGDST_ERR_TRPR_INDEX = 101, /* compile.m:515,521 */
    EGDST_ERR_TRPR_CASES = 102, /* compile.m:543 */
    EGDST_ERR_CHECKSUM = 1,     /* egdst_solver.c:577-582 */
    EGDST_ERR_NOSAVINGS = 2,    /* egdst_solver.c:589 */
    EGDST_ERR_GRIDSPACE = 3,    /* egdst_solver.c:662 */
    EGDST_ERR_EMPTYCHOICE = 4,  /* egdst_solver.c:694-702 */
    EGDST_ERR_ALLINF = 5,       /* egdst_solver.c:704-710 */
    EGDST_ERR_ENVELOPE = 6,     /* egdst_solver.c:723-728 */
    EGDST_ERR_ADRAW_INIT = 7,   /* egdst_solver.c:1012-1016 */
    EGDST_ERR_ADRAW_LOOP = 8,   /* egdst_solver.c:963-978 (warning in the reference) */
    EGDST_ERR_THRSPACE = 9,     /* egdst_solver.c:1327 */
    EGDST_ERR_TWO_ANALYTIC = 10,/* egdst_solver.c:1689-1693 */
    EGDST_ERR_BRACKET = 11,     /* egdst_solver.c:1934-1945 */
    EGDST_ERR_CASHINVERSE = 12, /* egdst_lib.c:293 */
    EGDST_ERR_INTERP2PT = 13,   /* egdst_lib.c:171 */
    EGDST_ERR_ENV2SPACE = 14,   /* egdst_solver.c:824,839,877 */
    EGDST_ERR_BARRIER = 103     /* internal: a phase barrier of the solve kernel was abandoned (hard error, no result) */
};

typedef struct curr_variables {
    int it;
    int ist;
    double st[EGDST_NNST];
    int id;
    double dc[EGDST_NND];
    double cash;
    double savings;
    double shock;
} PeriodVars;

typedef struct egdst_ctx {
    int t0, T, ngridm, ngridmax, nthrhmax, ny, nd, nnd, nst, nnst;
    int optim_UasD, optim_MUnoD, optim_UnoD, optim_TRPRnoSH;
    int byval;
    double mmax, a0;
    double tolerance, zeroconsumption, doublepoint_delta; /* compile.m:757-777 -D flags, here run-time */
    const double *stm;       /* [2*nnst] sizes then strides            (egdstmodel.m:651,1432) */
    const double *states;    /* [nst*nnst] column-major                (egdst_lib.c:241-245)   */
    const double *decisions; /* [nd*nnd] column-major                  (egdst_lib.c:248-252)   */
    int *status;             /* optional status word                                            */
    double param[EGDST_NPARAM_];
} egdst_ctx;

/* left grid point of the interval a value of a continuous state falls into, end intervals extended outwards
 * (what bxsearch, egdst_lib.c:123-165, returns on an increasing grid). */
EGDST_FN int egdst_gridcell(double x, const double *g, int n) {
    int i = 0;
    while (i < n - 2 && x >= g[i + 1]) i++;
    return i;
}

#ifdef __CUDACC__
#define EGDST_MODEL_FAIL(cx, code) do { if ((cx)->status) *((cx)->status) = (code); } while (0)
#else
#define EGDST_MODEL_FAIL(cx, code) do { if ((cx)->status) *((cx)->status) = (code); } while (0)
#endif

#endif
