/* dd_b200.h -- C ABI of libdd_b200.so: B200-native time stepping for the
 * nonlinear temperature-enhanced drug-diffusion model.
 *
 * The reference (phao/NA-nonlinear-temperature-enhanced-diffusion-model-DD) has
 * no FFI of its own: its boundary for this path is the Python class API of
 * src/prob1base.py.  Each entry point below names the reference call(s) it
 * replaces (file:line relative to the reference root); INTEGRATION.md shows the
 * ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *  - plain C: opaque handles, pointers and sizes only.  Every function returns
 *    0 on success or a negative dd_status; dd_last_error() gives the message.
 *    Nothing throws across the ABI.
 *  - fields are float64, C order, shape (nrows, M+1) per member, exactly the
 *    reference's (N+1, M+1) arrays (src/prob1base.py:242-248) when row0 = 0 and
 *    nrows = N+1.  Host pointers unless the name says `_dev`.
 *  - all device work is ordered on the context's stream; one host thread per
 *    context.  The library owns device memory; the caller owns the handles.
 *  - a "batch" is B independent trajectories (members) on one grid; B = 1 is a
 *    single trajectory.  A batch may hold only the row slab [row0, row0+nrows)
 *    of the global grid (domain decomposition); the caller then fills the
 *    non-owned rows through the *_dev pointers (halo exchange).
 */
#ifndef DD_B200_H
#define DD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dd_ctx dd_ctx;
typedef struct dd_batch dd_batch;

enum dd_status {
    DD_OK = 0,
    DD_ERR_INVALID = -1,       /* bad argument (the reference raises AssertionError, e.g. dt <= 0: src/prob1base.py:3118) */
    DD_ERR_CUDA = -2,          /* CUDA runtime error */
    DD_ERR_NOT_CONVERGED = -3, /* a Newton linear solve missed its residual bound */
    DD_ERR_NO_DEVICE = -4,     /* no CUDA device: there is no CPU fallback */
    DD_ERR_DOMAIN = -5         /* HCsTriple cs corrector: 2 - dt Kd (Sd - cd1)(1 + cl1) below its positivity threshold
                                  (the reference raises ValueError, src/prob1base.py:3410-3413) */
};

enum dd_var { DD_VAR_CP = 0, DD_VAR_T = 1, DD_VAR_CL = 2, DD_VAR_CD = 3, DD_VAR_CS = 4 };

/* forcing (MMS source) modes */
enum dd_forcing_mode {
    DD_MODE_NONE = 0,      /* NoForcingTerms, src/prob1base.py:852-869 */
    DD_MODE_ARRAYS = 1,    /* caller evaluates fcp..fcs and uploads them (any ForcingTermsBase) */
    DD_MODE_SEPARABLE = 2, /* u_v = phi_v(t) X_v(x) Y_v(y): fused evaluation from 1-D tables */
    DD_MODE_EXPSIN = 3,    /* MMSCaseExpSin closed form, src/prob1_mms_cases.py:296-337 */
    DD_MODE_PROGRAM = 4    /* any MMSCaseSymbolic: sources and exact solution from a generated kernel
                              (dd_b200_program.h), src/prob1base.py:1226-1280 */
};

/* time profiles phi_v(t) of DD_MODE_SEPARABLE */
enum dd_phi_kind {
    DD_PHI_K_INV1PT = 0, /* p0 / (1 + t) */
    DD_PHI_K_EXP = 1,    /* p0 exp(-p1 t) */
    DD_PHI_K_LINEAR = 2, /* p0 - p1 t */
    DD_PHI_K_OSC = 3,    /* p0 (1 + p1 sin(p2 t)) */
    DD_PHI_K_CONST = 4,  /* p0 */
    DD_PHI_K_HOST = 5    /* p = phi(t0), phi'(t0), phi(t1), phi'(t1): refreshed by the caller before each step */
};

/* dd_model, dd_program_member, dd_program_args: shared with generated forcing programs */
#include "dd_b200_program.h"

/* options of the predictor-corrector step: the constructor arguments of
 * P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_RegHCsTriple (src/prob1base.py:3612-3629) */
typedef struct dd_pc_options {
    int num_pc_steps;          /* default 1 */
    int num_newton_steps;      /* default 1 */
    int num_newton_iterations; /* cs corrector cap, default 5 */
    int cd_band_swap;          /* 1 reproduces the W/S band exchange of src/prob1base.py:3094-3100 (default) */
    double consec_xs_rtol;     /* default 1e-6; 0 disables the exit test */
    double solve_tol;          /* bound on |x - x_exact|_inf / |v_new|_inf per linear solve, default 1e-14 */
    int max_sweeps;            /* give up (DD_ERR_NOT_CONVERGED) beyond this many SOR sweeps, default 20000 */
    int fixed_sweeps;          /* > 0: use exactly this many sweeps and skip the adaptive plan */
    int extrapolate_guess;     /* 1: when consecutive steps ping-pong between two slots, start the SOR iteration
                                  from the previous step's increment instead of zero.  Default 0: measured on the
                                  bench mesh it saves one sweep per solve and costs as much in the first load
                                  (the zeroed-T-boundary layer keeps early increments from being smooth in time) */
    int _pad;
} dd_pc_options;

/* what one dd_step_pc call did (worst member of the batch) */
typedef struct dd_step_stats {
    int sweeps[3];        /* SOR sweeps used for the T, cl, cd solves (last Newton step) */
    int passes[3];        /* kernel passes per solve */
    int retries;          /* how often the step was redone with more sweeps */
    int cs_newton_iters;  /* iterations of the cs corrector (max over members, last pc step) */
    double rho[3];        /* Gershgorin ratio of the three matrices */
    double resid[3];      /* max scaled residual |bb - (I-G)x| */
    double bound[3];      /* resulting bound on |x - x_exact| / |v_new| */
} dd_step_stats;

/* ---- context ----------------------------------------------------------- */
int dd_ctx_create(int device, void* cuda_stream /* cudaStream_t or NULL */, dd_ctx** out);
int dd_ctx_destroy(dd_ctx* ctx);
int dd_ctx_synchronize(dd_ctx* ctx);
const char* dd_last_error(const dd_ctx* ctx);
const char* dd_version(void);

/* ---- batch: grid + members (replaces Grid / make_uniform_grid, src/prob1base.py:220-490,
 *      and holds what StateVars holds, 1913-2085) ---------------------------------- */
int dd_batch_create(dd_ctx* ctx, int N, int M, const double* x /* N+1 */, const double* y /* M+1 */,
                    int nmembers, int row0, int nrows, int own0, int own1, int nslots, dd_batch** out);
int dd_batch_destroy(dd_batch* b);
int dd_batch_set_models(dd_batch* b, int first, int count, const dd_model* models);
int dd_batch_set_active(dd_batch* b, int first, int count, const int* active);

/* forcing selection (replaces the ForcingTerms_RegHCsTriple object, src/prob1base.py:3468-3551) */
int dd_forcing_none(dd_batch* b);
/* u_v = phi_v(t) sum_{r<nterms} X_{v,r}(x) Y_{v,r}(y).  tables: X[v][d] holds nterms rows of N+1 doubles
 * (d = 0,1,2: value, 1st, 2nd derivative), Y likewise (M+1); XQ[q] (q: cp, T, cl) holds nterms rows of
 * 3 values per node i (Gauss abscissae of cell i, src/prob1base.py:508-570);
 * phi_kind[v], phi_p[v][4] select phi_v(t).  per-member phi via dd_forcing_set_phi. */
int dd_forcing_separable(dd_batch* b, int nterms, const double* const X[5][3], const double* const Y[5][3],
                         const double* const XQ[3], const double* const YQ[3], const int phi_kind[5],
                         const double phi_p[5][4]);
int dd_forcing_set_phi(dd_batch* b, int first, int count, const int* phi_kind /* count*5 */,
                       const double* phi_p /* count*5*4 */);
/* sx = sin(pi x), cx = cos(pi x) at the nodes, sxq = sin(pi x) at the abscissae (3 per node) */
int dd_forcing_expsin(dd_batch* b, const double* sx, const double* cx, const double* sy, const double* cy,
                      const double* sxq, const double* syq);
/* host-evaluated sources for the next step: f[v][slot] (slot 0: t0, slot 1: t0+dt), NULL = zero */
int dd_forcing_arrays(dd_batch* b, int member, const double* const f[5][2]);
/* generated forcing: `image` is a cubin or PTX image (NVRTC output for sm_100a) that defines the kernel
 * `dd_program` of dd_b200_program.h; xq / yq: Gauss abscissae of the dual cells, 3 per node ((N+1)*3, (M+1)*3
 * doubles, src/prob1base.py:538-570).  Replaces the per-step host evaluation of MMSCaseSymbolic's lambdified
 * expressions (src/prob1base.py:1226-1280) and of the ForcingTerms object built on them. */
int dd_forcing_program(dd_batch* b, const void* image, unsigned long long image_bytes, const double* xq,
                       const double* yq);

/* ---- state slots ------------------------------------------------------- */
int dd_state_upload(dd_batch* b, int slot, int member, const double* const fields[5] /* NULL entries skipped */);
int dd_state_download(dd_batch* b, int slot, int member, double* const fields[5]);
int dd_state_fill_exact(dd_batch* b, int slot, const double* t, int n_t); /* state_from_mms_when, src/prob1base.py:3433-3449 */
int dd_state_dev_ptr(dd_batch* b, int slot, int var, void** ptr, long long* member_stride, int* ld);
int dd_work_dev_ptr(dd_batch* b, const char* name, void** ptr); /* "cp1p","cs1p","YT","Ycl","Ycd","bb","aW",... */
int dd_work_upload(dd_batch* b, const char* name, int member, const double* host);
int dd_work_download(dd_batch* b, const char* name, int member, double* host);

/* ---- the hot path ------------------------------------------------------ */
/* ForwardEulerIntegrator.step, src/prob1base.py:2889-2903.  t0/dt: n_t = 1 (broadcast) or nmembers values */
int dd_step_feuler(dd_batch* b, int slot_in, int slot_out, const double* t0, const double* dt, int n_t);
/* P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_RegHCsTriple.step, src/prob1base.py:3117-3149 */
int dd_step_pc(dd_batch* b, int slot_in, int slot_out, const double* t0, const double* dt, int n_t,
               const dd_pc_options* opt, dd_step_stats* stats /* may be NULL */);
/* The same step with deferred verification: the step is enqueued and the call returns after reading the
 * convergence summaries of the PREVIOUS deferred step (have_prev = 1 and prev_stats filled then), so the host
 * never drains the stream between steps.  A rejected step is redone with more sweeps, together with the step
 * enqueued after it; this needs the input of the rejected step, so rotate three slots (a->b, b->c, c->a).
 * Every other entry point that reads or writes the batch first settles a pending step (dd_step_pc_flush). */
int dd_step_pc_deferred(dd_batch* b, int slot_in, int slot_out, const double* t0, const double* dt, int n_t,
                        const dd_pc_options* opt, dd_step_stats* prev_stats /* nullable */, int* have_prev /* nullable */);
int dd_step_pc_flush(dd_batch* b, dd_step_stats* stats /* nullable */, int* have_stats /* nullable */);
/* nsteps PC steps with times advanced on the device (current_t += dt, src/mms_trial_utils.py:128);
 * result ends in slot_a if nsteps is even else slot_b; optional per-step error norms into
 * norms_out[(nsteps+1)][nmembers][8] (index 0 = initial state) */
int dd_run_pc(dd_batch* b, int slot_a, int slot_b, const double* t0, const double* dt, int n_t, int nsteps,
              const dd_pc_options* opt, double* norms_out, dd_step_stats* stats);
int dd_run_feuler(dd_batch* b, int slot_a, int slot_b, const double* t0, const double* dt, int n_t, int nsteps,
                  double* norms_out);
/* the same run loops with the per-step norms kept on the device and combined there: combined[m][6] = the
 * combined max-integral error norm of member m over all variables (calculate_combined_error_norm,
 * src/mms_trial_utils.py:15-53) and the per-variable figures cp, T, cl, cd, cs of NumericalErrorSummary
 * (src/mms_trial_utils.py:150-190), bit for bit what the reference's Python arithmetic gives on the same norms
 * (builtin sum(), trapezoid, `max(0.0, nan)` keeps 0.0).  Only 6 doubles per member cross PCIe. */
int dd_run_pc_errors(dd_batch* b, int slot_a, int slot_b, const double* t0, const double* dt, int n_t, int nsteps,
                     const dd_pc_options* opt, double* combined /* nmembers*6 */, dd_step_stats* stats);
int dd_run_feuler_errors(dd_batch* b, int slot_a, int slot_b, const double* t0, const double* dt, int n_t, int nsteps,
                         double* combined /* nmembers*6 */);
void dd_pc_options_default(dd_pc_options* opt);

/* ---- pieces of the step, for the class-level API and its tests ---------- */
/* SemiDiscreteField.Fcp/FT/Fcl/Fcd/Fcs(state, t), src/prob1base.py:2599-2672: slot_out[v] = F_v */
int dd_eval_fields(dd_batch* b, int slot_in, int slot_out, const double* t, int n_t);
/* initial_cp_pred / initial_cs_pred + Y_T, Y_cl, Y_cd (src/prob1base.py:2953-2965, 3631-3645, 3122-3124)
 * results in work buffers "cp1p","cs1p","YT","Ycl","Ycd" */
int dd_pc_predict(dd_batch* b, int slot_in, const double* t0, const double* dt, int n_t);
/* newton_step_T / _cl / _cd (src/prob1base.py:2998-3115): linearise at slot_star (cp, T, cl, cd, cs all from
 * that slot), Y from work buffer "YT"/"Ycl"/"Ycd"; T1 / cl1 are read from slot_new for var = CL / CD;
 * result written to slot_new[var] */
int dd_pc_newton(dd_batch* b, int var, int slot_star, int slot_new, const double* t0, const double* dt, int n_t,
                 const dd_pc_options* opt, dd_step_stats* stats);
/* corrector_cp_step + corrector_cs_step (src/prob1base.py:2967-2996, 3665-3702): reads state0 from slot0 and
 * T1, cl1, cd1 from slot_new; writes cp, cs of slot_new */
int dd_pc_correct(dd_batch* b, int slot0, int slot_new, const double* t0, const double* dt, int n_t,
                  const dd_pc_options* opt, int* cs_iters_out /* nmembers or NULL */);
/* last_residual[var] = 2 v - dt F_v(state, t1) - Y_v (src/prob1base.py:3041-3043, 3076-3078, 3111-3113) */
int dd_pc_residual(dd_batch* b, int var, int slot_state, const double* t0, const double* dt, int n_t,
                   double* out_host /* member 0.. all members, nrows*(M+1) each */);

/* ---- error norms (Grid.norm_H / grad_H / norm_p as used by collect_errors,
 *      src/prob1base.py:387-433, src/mms_trial_utils.py:81-110) ------------------ */
/* out[member][8] = H2[cp,T,cl,cd,cs], P2[T,cl,cd] of (state - exact(t)); exact from the forcing tables,
 * or from slot_exact when >= 0 */
int dd_error_norms(dd_batch* b, int slot, int slot_exact, const double* t, int n_t, double* out);

/* ---- slab-decomposed meshes: the PC step in phases, halo exchange by the caller in between -----
 * phase 0: times + predict (needs halo rows of the five input fields); 1/2/3: Newton solve of T/cl/cd into
 * slot_out on the owned rows (exchange that field's halo afterwards); 4: correctors on all local rows;
 * 5: cs exit decision (after reducing the work buffers "cs_it_max" (max) / "cs_it_min" (min) over ranks) and
 * summary[3][4] = rho, ratio (<= 1 means the residual bound is met), resid, bound of the T, cl, cd solves;
 * 6: as 5 without the read-back (nothing is waited for): the caller reduces the device buffers "summary"
 *    (3 x 4 doubles) and "cs_used" (int per member) itself, e.g. while the next step is already running.
 * Phases 21/22/23 only assemble the T/cl/cd system and 31/32/33 only solve it, so that the caller can
 * max-reduce the Gershgorin ratio ("solve_stats" work buffer, first double of each 5-double record) over
 * ranks in between: the SOR relaxation factor is then the same on every rank.
 * All ranks must use the same sweep plan (dd_batch_set_plan) for results independent of the decomposition. */
int dd_step_pc_phase(dd_batch* b, int phase, int slot_in, int slot_out, const double* t0, const double* dt, int n_t,
                     const dd_pc_options* opt, double* summary, int* cs_iters);
int dd_batch_set_plan(dd_batch* b, const int sweeps[3]);
/* Gershgorin ratios (T, cl, cd) from which the solves derive their SOR relaxation factor instead of the ratio
 * reduced on this device; NULL or values outside [0, 1) restore the default.  A slab mesh passes the all-reduced
 * ratios of an earlier step here: every rank then relaxes identically without a collective before each solve. */
int dd_batch_set_relax_rho(dd_batch* b, const double rho[3]);
int dd_batch_get_plan(dd_batch* b, int sweeps[3]);
int dd_sweeps_for_rho(double rho, int max_sweeps); /* SOR sweeps the planner uses for a Gershgorin ratio rho */
int dd_next_plan(int cur, double rho, double ratio, int max_sweeps); /* next step's sweeps from this step's verified ratio */

/* One segment of a Newton solve of the phased step: `sweeps` red-black SOR sweeps of variable var (1 = T, 2 = cl,
 * 3 = cd) continuing from the iterate of the previous segment (first = 1: from zero); unless last = 1 the iterate
 * stays in the work array "x_cur" (dd_work_dev_ptr), whose halo rows a slab driver exchanges before the next
 * segment.  Replaces the part of SuperLU's single direct solve (reference src/prob1base.py:2103, 2130) that a row
 * slab cannot do alone when the matrix needs more sweeps than its halo supports between two exchanges. */
int dd_pc_solve_segment(dd_batch* b, int var, int slot_in, int slot_out, const dd_pc_options* opt, int sweeps,
                        int first, int last);

/* ---- halo exchange of a row slab by direct NVLink stores ------------------------------------------------------
 * One process per GPU on one node: each rank exports the allocations its neighbours write into (its state fields,
 * its flag block) as 64-byte CUDA IPC handles, imports the neighbours' and from then on exchanges halo rows with ONE
 * kernel per exchange: dd_halo_push copies the rank's first / last `count` doubles of owned rows (src_top, src_bot)
 * into the up / down neighbour's halo rows (dst_up, dst_down: imported pointers + offsets) and handshakes through the
 * flag blocks (ready-to-receive, delivered), exchange number `seq` strictly increasing and equal on all ranks; null
 * src pointers on the mesh's first / last rank.  The reference has no counterpart (it is single-process); in the
 * slab driver this replaces the NCCL send / recv pairs of ddmesh.exchange_halos. */
int dd_ipc_export(dd_ctx* ctx, const void* dev_ptr, unsigned char* handle64); /* dev_ptr: base of a cudaMalloc'ed block */
int dd_ipc_import(dd_ctx* ctx, const unsigned char* handle64, void** dev_ptr);
int dd_ipc_close(dd_ctx* ctx, void* dev_ptr);
int dd_halo_flags_create(dd_ctx* ctx, void** flags);
int dd_halo_flags_destroy(dd_ctx* ctx, void* flags);
int dd_halo_push(dd_ctx* ctx, const double* src_top, double* dst_up, const double* src_bot, double* dst_down,
                 long long count, void* my_flags, void* up_flags, void* down_flags, unsigned seq);
int dd_halo_status(dd_ctx* ctx, void* my_flags, int* status); /* 1: a handshake timed out */

/* ---- instrumentation (bench.py) ------------------------------------------ */
long long dd_launch_count(void);                /* kernels launched by the library since it was loaded */
int dd_profile_enable(int on);                  /* bracket every launch group with CUDA events */
int dd_profile_read(const char** names, double* ms, long long* count, int reset); /* returns #classes (<= 16) */

int dd_probe_math(dd_ctx* ctx, int n, const double* in, double* out_exp, double* out_rcp); /* accuracy probe of the inline device exp / 1/x */
/* fp64 roof of the device, measured: every SM runs chains of independent double-precision fused multiply-adds for
   about `ms_target` milliseconds (CUDA events on the context's stream); *tflops = 2 * FMAs / time.  The PC step is
   bound by fp64 issue rather than by HBM, bench.py reports it against this number. */
int dd_probe_fp64(dd_ctx* ctx, double ms_target, double* tflops);
/* name of the CUDA kernel the last solve of variable `var` (1 = T, 2 = cl, 3 = cd) was launched with */
const char* dd_solver_kernel_name(int var);

#ifdef __cplusplus
}
#endif
#endif /* DD_B200_H */
