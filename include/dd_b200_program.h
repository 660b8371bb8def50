/* dd_b200_program.h -- the contract between libdd_b200.so and a GENERATED forcing program (DD_MODE_PROGRAM).
 *
 * Replaces, for manufactured solutions that are arbitrary SymPy expressions, what the reference does on the
 * host: `MMSCaseSymbolic` lambdifies the five expressions and their derivatives (src/prob1base.py:1226-1280,
 * 1283-1487) and `ForcingTerms_RegHCsTriple` composes the five sources from them at every step
 * (src/prob1base.py:2313-2378, 3503-3551).  Here the host side prints the same expressions as CUDA C, compiles
 * them with NVRTC for sm_100a and hands the image to dd_forcing_program(); the library launches the program's
 * kernel once per time level into its staged source arrays, so that the step kernels run exactly as they do for
 * the built-in cases.
 *
 * This file is plain C with no #include: it is also fed verbatim to NVRTC as the program's only header.
 *
 * The image (cubin or PTX) must define
 *     extern "C" __global__ void dd_program(dd_program_args a);
 * launched with block (128,1,1) and grid (ceil((M+1)/128), nrows, nmembers) -- hence at most 65535 rows and 65535
 * members per batch in this mode (dd_forcing_program batches beyond that are refused at launch): thread (j, r, m) owns node
 * (row0 + r, j) of member m and writes a.out[v][m * mstride + r * ld + j] for v = cp, T, cl, cd, cs:
 *   what == DD_PROGRAM_SOURCES: the five MMS sources at time members[m].t[tslot] (fcp: 3x3 Gauss average over
 *                               the dual cell at interior nodes, 0 on the boundary, src/prob1base.py:493-598);
 *   what == DD_PROGRAM_EXACT:   the exact solution at that time.
 * Members with active == 0 are skipped.
 */
#ifndef DD_B200_PROGRAM_H
#define DD_B200_PROGRAM_H

/* ModelConsts (src/prob1base.py:28-45) + model kind + eta; one per member */
typedef struct dd_model {
    double K1, K2, K3, K4, DT, Dl_max, phi_l, gamma_T, Kd, Sd, Dd_max, phi_d, phi_T, r_sp, T_ref;
    double eta;   /* regularisation factor of H_eta, src/prob1base.py:3452-3466 */
    int kind;     /* 1 = DefaultModel01, 2 = DefaultModel02 (Dd uses T + T_ref), src/prob1base.py:71-217 */
    int reaction; /* cs/cd interaction Kd (Sd - cd)(1 + cl) F2(cs): 0 = RegHCsTriple, F2 = H_eta(cs) (3553-3593);
                     1 = CsTriple, F2 = cs (2842-2876); 2 = HCsTriple, F2 = (cs > 0) (3303-3340).  The cs
                     predictor / corrector follow the matching integrator class (3152-3219, 3343-3430, 3596-3702);
                     for 1 and 2 the corrector is a closed form: pass num_newton_iterations = 0 */
} dd_model;

enum dd_program_what { DD_PROGRAM_SOURCES = 0, DD_PROGRAM_EXACT = 1 };

/* what a program sees of one member (kind is reported as 2 with T_ref = 0 for a DefaultModel01 member) */
typedef struct dd_program_member {
    dd_model model;
    double t[2]; /* t0 and t0 + dt of the step being prepared */
    int active;
    int _pad;
} dd_program_member;

typedef struct dd_program_args {
    const double* x;  /* node coordinates (device), N+1 */
    const double* y;  /* M+1 */
    const double* xq; /* Gauss abscissae of the dual cells, 3 per node: xq[3 i + a] (rows 0 and N unused) */
    const double* yq;
    const dd_program_member* members;
    double* out[5];
    long long mstride; /* doubles between members */
    int N, M;          /* global grid: nodes 0..N x 0..M */
    int row0, nrows;   /* the rows this batch holds */
    int ld;            /* row pitch in doubles */
    int nmembers;
    int what;  /* dd_program_what */
    int tslot; /* 0: members[m].t[0], 1: members[m].t[1] */
} dd_program_args;

#endif
