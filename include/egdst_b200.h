/* egdst_b200.h -- C ABI of the B200-native egdst solver / simulator library.
 *
 * One shared library is built per generated model image (libegdst_b200_<key>.so: the user's exec
 * strings compiled into the sm_100a kernels).  Its entry points are what the reference's three MEX
 * gateways bind (INTEGRATION.md shows the MEX stubs):
 *
 *   egdst_solve        replaces  [M,D,dbg] = egdst_solver(model)      @egdstmodel/egdst_solver.c:143-239
 *   egdst_simulate     replaces  sims = egdst_simulator(model,rnd)    @egdstmodel/egdst_simulator.c:47-117
 *   egdst_call         replaces  res  = egdst_call(model,sw,args)     @egdstmodel/egdst_call.c:17-125
 *   egdst_desc         replaces  parseModel()/loadparameters()        @egdstmodel/egdst_lib.c:34-62, compile.m:469-474
 *   egdst_last_error   replaces  err[300] / mexWarnMsgTxt / mexErrMsgTxt  egdst_lib.c:299-302, egdst_solver.c:237
 *
 * plus the multi-GPU / batched entry points the reference does not have (egdst_solve_batch,
 * egdst_sim_moments, Philox-driven simulation).
 *
 * Conventions: plain pointers and sizes, caller owns every input and output buffer, the library owns
 * egdst_solution objects until egdst_free_solution.  Return codes: 0 ok; 1 soft error (partial result
 * kept, message available -- the reference warns and returns partial M,D, egdst_solver.c:237); 2 hard
 * error (mexErrMsgTxt in the reference).  All device work runs on the stream set with egdst_set_stream
 * (default: the legacy default stream).  There is no CPU path: without a CUDA device every compute
 * entry point returns 2 with "no CUDA device".
 */
#ifndef EGDST_B200_H
#define EGDST_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define EGDST_ABI_VERSION 2

/* The model object flattened to a POD: exactly the properties the reference reads across the MEX
 * boundary (egdst_lib.c:37-55, compile.m:472, egdst_solver.c:162) plus the -D flags of compile.m:757-777. */
typedef struct egdst_desc {
    int abi_version;
    int t0, T, ngridm, ngridmax, nthrhmax, ny, nd, nnd, nst, nnst;
    double mmax, a0;
    int optim_UasD, optim_MUnoD, optim_UnoD, optim_TRPRnoSH;
    double tolerance, zeroconsumption, doublepoint_delta; /* cflags TOLERANCE, ZEROCONSUMPTION, DOUBLEPOINT_DELTA */
    const double *stm;        /* [2*nnst]                                       */
    const double *states;     /* [nst*nnst] column-major                        */
    const double *decisions;  /* [nd*nnd]  column-major                         */
    const double *params;     /* [nparam] param(i).value                        */
    int nparam;
    const double *quadrature; /* [2*ny] weights then abscissas in (0,1), as model.quadrature (may be NULL if ny==1) */
    int neq;                  /* numel(model.eq)                                 */
    int device;               /* CUDA device ordinal                             */
    double sigma_eps;         /* EXTENSION (no reference counterpart; 0 = off = the reference's hard max): scale of additive
                               * extreme-value taste shocks on the discrete choice.  > 0: the expectation uses the logsum and the
                               * choice probabilities of the choice-specific value functions (egdst_solution_choice_cell)   */
} egdst_desc;

typedef struct egdst_solution egdst_solution;

/* library / model-image introspection */
int egdst_abi_version(void);
const char *egdst_model_key(void);   /* key of the model image compiled into this library */
int egdst_model_nparam(void);
int egdst_model_neq(void);
const char *egdst_last_error(void);  /* per-thread message of the last non-zero return */
void egdst_set_stream(void *cuda_stream);
/* measurement aids: number of kernels launched by this library so far; optional CUDA-event timing of
 * every launch, accumulated per kernel class (setup, solve, tables, simulate, other).  Profiling adds two event
 * records per launch and switches the solve kernel's phase timers on -- never enable it in a timed region. */
long long egdst_launch_count(void);
int egdst_profile_classes(void);
const char *egdst_profile_class_name(int cls);
void egdst_profile_enable(int on);
int egdst_profile_read(double *ms, long long *count); /* arrays of egdst_profile_classes() entries */

/* ---- solve ------------------------------------------------------------------------------- */
/* Backward induction for one model.  *out receives a new solution object (device resident). */
int egdst_solve(const egdst_desc *d, egdst_solution **out);
/* nvec parameter vectors of the same model solved in one pass (params: [nvec*nparam], row = vector). */
int egdst_solve_batch(const egdst_desc *d, const double *params, int nvec, egdst_solution **out);
/* Re-run the solve into an existing solution object (same dimensions; new parameter values allowed).
 * Asynchronous on the library stream: no host synchronisation, no allocation. */
int egdst_resolve(egdst_solution *s, const egdst_desc *d, const double *params);
/* Rows per cell.  mlen/thlen: [nvec*NT*nst], index (ivec*NT+it)*nst+ist; 0 rows = infeasible (it,ist). */
int egdst_solution_sizes(egdst_solution *s, int *mlen, int *thlen);
/* Packed copy-out in cell order: M cell = mlen x 4 column-major (M,C,A,V; row 0 = a0,0,a0,evf(a0)),
 * D cell = thlen x 2 column-major (decision index, threshold) -- the layouts of saveoutput,
 * egdst_solver.c:917-951.  Mbuf holds 4*sum(mlen) doubles, Dbuf 2*sum(thlen). */
int egdst_solution_export(egdst_solution *s, double *Mbuf, double *Dbuf);
/* EXTENSION, smoothing mode only (desc.sigma_eps > 0): the choice-specific cell of decision id in (it, ist) -- what the
 * reference discards after envelop() (egdst_solver.c:720-730).  rows x 4 column-major (M,C,A,V; row 0 = a0,0,a0,
 * evf_d(a0)); *rows = 0: decision not available.  M may be NULL to query the row count. */
int egdst_solution_choice_cell(egdst_solution *s, int ivec, int it, int ist, int id, double *M, int cap, int *rows);
/* status of the last (re)solve of vector ivec: returns the code, fills it/ist/id of the first failure */
int egdst_solution_status(egdst_solution *s, int ivec, int *it, int *ist, int *id);
int egdst_solution_nvec(const egdst_solution *s);
/* total number of EGM grid points kept over all (it,ist,id) of the last solve (the solve work unit) */
long long egdst_solution_units(egdst_solution *s);
/* diagnostic: how many zero-consumption re-sends requested by grid points AFTER the seed stage of the savings grid
 * (egdst_solver.c:1080-1099) the last solve went through */
long long egdst_solution_resends(egdst_solution *s);
/* measurement aid: device time in ms per phase of the solve kernel -- terminal, seed, egm, resend, envelope2, rank,
 * merge, tables, then six steps of the last work item of the EGM phase (16 entries) -- accumulated over this object's solves while egdst_profile_enable(1) was in effect;
 * reading resets the counters.  Returns the number of entries. */
int egdst_solution_phase_ms(egdst_solution *s, double *ms);
/* diagnostic: the secondary upper envelope (envelope2, egdst_solver.c:776-913) of n EGM points (M, C, V_d) of decision
 * id (state 0, period it) in generation order, run by the solve kernel's own phases.  out*: room for ngridmax
 * doubles each; *nout receives the number of points kept.  The tests compare it with the reference's envelope2. */
int egdst_test_envelope2(const egdst_desc *d, int it, int ist, int id, const double *X, const double *C, const double *V, int n,
                         double evfa0, double *outX, double *outC, double *outV, int *nout);
void egdst_free_solution(egdst_solution *s);
/* Frees what the library keeps between calls: the one released solution object it caches for re-use by the next
 * solve of the same shape, and the calling thread's simulation workspace.  (The reference leaks on error paths and
 * relies on MATLAB clearing the MEX file, egdst_solver.c:19; a long-lived host calls this when it unloads a model.) */
void egdst_shutdown(void);
/* Build a solution object from host M/D cells (the MEX simulator receives model.M, model.D). */
int egdst_solution_import(const egdst_desc *d, const int *mlen, const int *thlen, const double *Mbuf, const double *Dbuf,
                          egdst_solution **out);

/* ---- simulate ---------------------------------------------------------------------------- */
/* sims: [nsimout, nt, nsim] column-major doubles, nsimout = 11+nnst+nnd+neq, NaN after death
 * (egdst_simulator.c:95-143).  init: [nsim*2] column-major (1-based ist0, m0).  randstream: U(0,1),
 * 4 slots per agent-period; rndtype 0 = own shocks, 1 = same shocks (egdst_simulator.c:71-75,109-114). */
int egdst_simulate(const egdst_desc *d, egdst_solution *s, int ivec, const double *init, int nsim,
                   const double *randstream, long long nrand, int rndtype, double *sims);
/* Same simulation with counter-based Philox4x32-10 uniforms generated on the device (no randstream).
 * agent0 = global index of the first agent (results do not depend on how agents are sharded).
 * sims may be NULL (moments only).  moments: [3, nsimout, nt] = sum x, sum x^2, alive count. */
int egdst_simulate_philox(const egdst_desc *d, egdst_solution *s, int ivec, const double *init, int nsim,
                          long long agent0, unsigned long long seed, double *sims, double *moments);
/* device-resident variants for sharded multi-GPU runs: d_init [nsim*2], d_sims / d_moments device pointers
 * (either may be NULL); asynchronous on the library stream. */
int egdst_simulate_device(const egdst_desc *d, egdst_solution *s, int ivec, const double *d_init, int nsim,
                          long long agent0, unsigned long long seed, const double *d_randstream, int rndtype,
                          double *d_sims, double *d_moments);

/* Estimation sweeps: moments-only simulation of the same agents (and the same Philox shocks) under parameter
 * vectors ivec0..ivec0+nvec-1 of a batched solution, in one launch.  moments: [nvec][3*nsimout*nt].
 * The _device variant takes device pointers, is asynchronous on the library stream and accumulates into
 * d_moments (the caller zeroes it), so a rank can all-reduce the buffer right after. */
int egdst_sim_moments(const egdst_desc *d, egdst_solution *s, int ivec0, int nvec, const double *init, int nsim,
                      long long agent0, unsigned long long seed, double *moments);
int egdst_sim_moments_device(const egdst_desc *d, egdst_solution *s, int ivec0, int nvec, const double *d_init, int nsim,
                             long long agent0, unsigned long long seed, double *d_moments);

/* ---- call -------------------------------------------------------------------------------- */
/* sw: 1 utility, 2 marginal utility, 3 discount, 4 budget, 5 marginal budget, 6 value function;
 * args: [narg*k] column-major with k = 4,4,2,6,6,3 (egdst_call.c:45-58); res: [narg]. */
int egdst_call(const egdst_desc *d, egdst_solution *s, int sw, const double *args, int narg, int k, double *res);

#ifdef __cplusplus
}
#endif
#endif
