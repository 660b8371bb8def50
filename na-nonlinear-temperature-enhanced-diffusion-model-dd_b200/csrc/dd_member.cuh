// dd_member.cuh -- launch interface of the one-CTA-per-member trajectory kernel (dd_member.cu)
#pragma once

#include <cuda_runtime.h>

#include "dd_nodeprog.cuh"

struct DDMemberArgs {
    DDGeom g;              // the batch's geometry (global arrays); the kernel works on a shared-memory copy of a member
    const DDMember* mem;   // member parameters, start time and step as set by k_time_coefs
    DDMember* mem_rw;      // the same array: the advanced time and its coefficients are written back
    DDForcing F;           // MMS tables
    DDStateC in;           // state at the start (HBM)
    DDState out;           // state after nsteps steps (may be the same slot: every member is read before it is written)
    int nmembers, nsteps, pitch;
    int cap;               // cs-Newton iterations (num_newton_iterations)
    double rtol;           // consec_xs_rtol of the reference's exit test (0: no test)
    int cd_swap, fixed_sweeps, max_sweeps;
    double solve_tol;
    double* combined;      // [member][6] overall, cp, T, cl, cd, cs combined error norms, or null (no norms taken)
    double* stats;         // [member][16]: sweeps[3], rho[3], resid[3], bound[3], cs iterations used, failed
};

bool dd_member_fits(int N, int M);
size_t dd_member_smem_bytes(int N, int M, int* pitch);
cudaError_t dd_launch_member_run(cudaStream_t stream, int mode, DDMemberArgs A, int sm_count);
