// dd_kernels.cuh -- launch interface of the sm_100a kernels (dd_kernels.cu,
// dd_solver.cu), used by the C ABI layer (dd_capi.cu).
#pragma once

#include <cuda_runtime.h>

#include "dd_nodeprog.cuh"

// all launchers return cudaGetLastError() of the launch
struct DDLaunch {
    cudaStream_t stream;
    int nmembers;
    int own0, own1;  // local rows computed by this launch: [own0, own1)
    int vr0, vr1;    // local rows whose Newton rows are valid (readable by the tile solver)
};

cudaError_t dd_launch_time_coefs(const DDLaunch& L, int mode, DDMember* mem, const double* t0_dev,
                                 const double* dt_dev, int n_t, int advance);

cudaError_t dd_launch_feuler(const DDLaunch& L, int mode, const DDGeom& g, const DDMember* mem, const DDForcing& F,
                             const DDStateC& in, const DDState& out);

cudaError_t dd_launch_fields(const DDLaunch& L, int mode, const DDGeom& g, const DDMember* mem, const DDForcing& F,
                             const DDStateC& in, const DDState& out, int slot);

cudaError_t dd_launch_predict(const DDLaunch& L, int mode, const DDGeom& g, const DDMember* mem,
                              const DDForcing& F, const DDStateC& in, const DDPredictOut& out);

// var = DD_T / DD_CL / DD_CD.  Resets then accumulates stats[member].rho.
cudaError_t dd_launch_assemble(const DDLaunch& L, int mode, int var, const DDGeom& g, const DDMember* mem,
                               const DDForcing& F, const DDStateC& ustar, const double* T1, const double* cl1,
                               const double* Y, int cd_swap, const DDRows& R, DDSolveStats* stats);

struct DDSolvePlan {
    int sweeps;      // red-black SOR sweeps in this pass
    int tile_i, tile_j;
    int halo;        // = 2*sweeps (+1 on the last pass), 0 when one tile covers the member
    int threads;
    int last_pass;   // compute residual stats and write v_new
    int const_band;  // T system: rows are (bb, dinv) + grid geometry instead of five stored bands
    int rpw;         // > 0: register-resident kernel with this many rows per warp (staged 16*rpw x 64)
    size_t smem_bytes;
    double rho_fix;  // >= 0: Gershgorin ratio to derive the relaxation factor from (instead of the device statistic)
};

cudaError_t dd_solver_configure();  // opt-in to large dynamic shared memory (once per process)

// wavefront solver (dd_wave.cu): one pass of `sweeps` red-black SOR sweeps marching down column strips; wide grids;
// used for the cl solve by default, the register-tile kernels for T and cd (as fast or faster there)
// halo exchange by direct peer stores (dd_halo.cu)
cudaError_t dd_launch_halo_push(cudaStream_t stream, const double* src_top, double* dst_up, const double* src_bot,
                                double* dst_down, long long count, unsigned* my_flags, unsigned* up_flags,
                                unsigned* down_flags, unsigned seq, unsigned* done_blocks, int* status);
// lane-private marching solver (dd_lane.cu): DD_LANE = 0 | 1 | list of T,cl,cd
cudaError_t dd_lane_configure();
bool dd_lane_ok(const DDGeom& g, const DDLaunch& L, int var);
int dd_lane_max_sweeps();
int dd_lane_pass_sweeps(int left);
cudaError_t dd_launch_solve_lane(const DDLaunch& L, const DDGeom& g, const DDMember* mem, const DDRows& R,
                                 const double* xin, double* xout, const double* vstar, double* vnew,
                                 int zero_boundary, DDSolveStats* stats, int const_band, int sweeps, int last_pass,
                                 double rho_fix);
cudaError_t dd_wave_configure();
bool dd_wave_ok(const DDGeom& g, const DDLaunch& L, int var);  // DD_WAVE = 0 | 1 | list of T,cl,cd (default: 0)
void dd_wave_max_sweeps(int const_band, int* any, int* wide);  // per pass: any variant / the widest one
cudaError_t dd_launch_solve_wave(const DDLaunch& L, const DDGeom& g, const DDMember* mem, const DDRows& R,
                                 const double* xin, double* xout, const double* vstar, double* vnew,
                                 int zero_boundary, DDSolveStats* stats, int const_band, int sweeps, int last_pass,
                                 double rho_fix);

// marching form of the predictor (sources as arrays or none, wide grids); fuse_T also assembles the
// constant-band T system of the first Newton step into R.bb / R.aW and the Gershgorin ratio into stats
bool dd_predict_march_ok(const DDGeom& g, const DDLaunch& L, int mode);
cudaError_t dd_launch_predict_march(const DDLaunch& L, int mode, const DDGeom& g, const DDMember* mem,
                                    const DDForcing& F, const DDStateC& in, const DDPredictOut& out, bool fuse_T,
                                    const DDRows& R, DDSolveStats* stats, bool store_YT);
cudaError_t dd_launch_solve_pass(const DDLaunch& L, const DDGeom& g, const DDMember* mem, const DDRows& R,
                                 const double* xin /* nullable: zero initial iterate */,
                                 const double* vold /* nullable, register kernel, xin null: start from vstar - vold */,
                                 double* xout, const double* vstar, double* vnew, int zero_boundary,
                                 DDSolveStats* stats, const DDSolvePlan& P);
// x0 = vstar - vold (0 off the interior) on local rows [L.vr0, L.vr1), row pitch of R
cudaError_t dd_launch_make_guess(const DDLaunch& L, const DDGeom& g, const DDRows& R, const double* vstar,
                                 const double* vold, double* x0);

// correctors.  cs Newton: `cap` iterations; when rtol > 0 the per-iteration global
// statistics go to it_max / it_min ([member][cap]) and dd_launch_cs_finish applies
// the reference's exit test (src/prob1base.py:3661).
cudaError_t dd_launch_correct(const DDLaunch& L, int mode, const DDGeom& g, const DDMember* mem,
                              const DDForcing& F, const DDStateC& s0, const double* T1, const double* cl1,
                              const double* cd1, double* cp_out, double* cs_out, int cap, double rtol,
                              double* it_max, double* it_min,
                              int* flags /* null: every member is RegHCsTriple; else per-member error flags */);

cudaError_t dd_launch_cs_finish(const DDLaunch& L, int mode, const DDGeom& g, const DDMember* mem,
                                const DDForcing& F, const DDStateC& s0, const double* cl1, const double* cd1,
                                double* cs_out, int cap, double rtol, double* it_max, double* it_min,
                                int* used_out);

// error norms of (state - exact): out[member][8] = H2[cp,T,cl,cd,cs], P2[T,cl,cd]
// exact == nullptr -> exact solution evaluated from the MMS tables (slot 0 time).
cudaError_t dd_launch_error_norms(const DDLaunch& L, int mode, const DDGeom& g, const DDMember* mem,
                                  const DDForcing& F, const DDStateC& s, const DDStateC* exact, double* partial,
                                  int nblocks_per_member, double* out);

cudaError_t dd_launch_fill_exact(const DDLaunch& L, int mode, const DDGeom& g, const DDMember* mem,
                                 const DDForcing& F, const DDState& out);

// residual of a Newton step: res = 2 v - dt F_v(state, t1) - Y   (reference last_residual)
cudaError_t dd_launch_residual(const DDLaunch& L, int mode, int var, const DDGeom& g, const DDMember* mem,
                               const DDForcing& F, const DDStateC& s, const double* Y, double* res);

int dd_norm_blocks_per_member(const DDGeom& g);

// all five MMS sources of time slot `slot` on the launch's rows -> out (fcp, fT, fcl, fcd, fcs)
cudaError_t dd_launch_eval_sources(const DDLaunch& L, int mode, const DDGeom& g, const DDMember* mem,
                                   const DDForcing& F, const DDState& out, int slot);

cudaError_t dd_launch_probe_math(cudaStream_t st, const double* in, double* out_exp, double* out_rcp, int n);
// fp64 roof probe: `iters` rounds of 8 independent FMA chains x 16 per thread, one 1024-thread CTA per SM slot
cudaError_t dd_launch_probe_fp64(cudaStream_t st, double* sink, int blocks, int iters);

// name of the kernel variant the last solve of a variable used (bench.py's roofline names the real kernel)
void dd_note_solver_kernel(int var, const char* name);
void dd_set_last_solver_kernel(const char* fmt, int a, int b, int c);
const char* dd_last_solver_kernel();
