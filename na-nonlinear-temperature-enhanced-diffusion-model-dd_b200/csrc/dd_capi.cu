// dd_capi.cu -- the C ABI of libdd_b200.so (declared in include/dd_b200.h):
// context / batch management, forcing tables, state transfer, and the host-side
// orchestration of the predictor-corrector step.  No Python.h, no torch types.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/dd_b200.h"
#include "dd_kernels.cuh"
#include "dd_member.cuh"
#include "dd_combine.cuh"
#include "dd_tables_host.h"

#define DD_VERSION_STR "dd_b200 0.1 (sm_100a)"

struct dd_ctx {
    int device;
    cudaStream_t stream;
    bool own_stream;
    std::string err;
    int sm_count;
    int nbatches = 0;   // batches alive on this context
    bool dead = false;  // dd_ctx_destroy was called while batches were alive: freed with the last batch
};

struct SolveSummary {  // reduced over members on the device
    double rho, ratio, resid, bound;
};

struct dd_batch {
    dd_ctx* ctx;
    int N, M, B, row0, nrows, own0, own1, nslots;
    size_t field_elems;  // B * nrows * ld
    DDGeom g;
    std::vector<double*> geo_dev;  // owned 1-D arrays
    DDMember* d_mem;
    std::vector<DDMember> h_mem;
    int mode;
    DDForcing F;
    std::vector<double*> table_dev;
    // DD_FORCING_PROGRAM: the loaded image, its kernel, what it is launched with
    cudaLibrary_t prog_lib = nullptr;
    cudaKernel_t prog_kernel = nullptr;
    dd_program_member* d_prog_mem = nullptr;
    const double *d_prog_xq = nullptr, *d_prog_yq = nullptr;
    std::vector<std::vector<double*>> slots;  // [slot][var]
    std::map<std::string, double*> work;
    double *d_t0, *d_dt;
    DDSolveStats* d_stats;  // [nsolve_cap][B]
    int nsolve_cap;
    SolveSummary* d_summary;  // [nsolve_cap]
    double *d_itmax, *d_itmin;
    int* d_used;
    int* d_flags;  // per member: bit 0 = HCsTriple corrector hit its positivity threshold
    bool has_variants;  // some member uses the CsTriple / HCsTriple reaction
    int cs_cap_alloc;
    double *d_norm_partial, *d_norm_out;
    int norm_bpm;
    double* d_combine_state = nullptr;  // dd_run_*_errors: running best / integral / last integrand, [B][18]
    double* d_member_stats = nullptr;   // one-CTA-per-member runs: [B][16] solve statistics of the last step
    const double* x_cur = nullptr;      // iterate left by an unfinished solve segment (dd_pc_solve_segment)
    double* d_combined = nullptr;       // [B][6]
    // staged sources: the MMS sources of a time level are evaluated once (k_eval_sources) into one of two
    // sets of five arrays and the step kernels read them in ARRAYS mode; the t1 set of a step is the t0
    // set of the next one.  smode / sF is what the step kernels are launched with.
    int smode;
    DDForcing sF;
    bool fused_sources;        // DD_FUSED_SOURCES=1: evaluate the sources inside every kernel instead
    double* src_set[2][DD_NVAR];
    bool src_has[2];
    double src_time[2];
    // sweep controller, one set per kind of initial iterate (0: zero, 1: extrapolated; the latter needs
    // far fewer sweeps); floor: one more than the last count that failed
    struct Ctl { int sweeps[3], extra[3], floor[3]; } ctl[2];
    int cm;  // set in use by the solves being issued
    // extrapolated initial iterate: when a step reads slot a and writes slot b right after a step that read
    // b and wrote a (ping-pong), slot b still holds the state of two steps ago and v_n - v_{n-1} starts the solves
    // step records: summaries come back through pinned host memory; with deferred verification the record
    // of step n is read while step n + 1 is already running (two records, used alternately)
    struct StepRec {
        bool active = false;      // enqueued, summary not read yet
        int slot_in = 0, slot_out = 0, nsolves = 0, n_t = 0;
        bool guess = false;
        dd_pc_options opt;
        std::vector<double> t0, dt;
        int sweeps[3] = {0, 0, 0}, passes[3] = {0, 0, 0};
        std::vector<int> solve_sweeps;   // sweeps used by each solve of the step
        SolveSummary* h_sums = nullptr;  // pinned [cap_sums]
        int* h_used = nullptr;           // pinned [B]
        int* h_flags = nullptr;          // pinned [B]
        int cap_sums = 0;
        cudaEvent_t done = nullptr;
    } rec[2];
    int rec_cur;
    double relax_rho[3];  // >= 0: ratio the solver derives omega from (slab meshes: the all-reduced one)
    bool gs_only[3];      // the solve of this variable has fallen back to plain Gauss-Seidel (see gs_fallback)
    bool prev_valid, use_guess, phase_fused_T;
    int prev_in, prev_out;
    double prev_dt, cur_dt;  // first member's step size (the increment scales with it)
    bool is_slab;
    int cmp0, cmp1;  // local rows with a complete stencil (slabs: everything but the outermost halo row)
    int asm0, asm1;  // rows whose Newton rows are exact: their stencil reads predictor output, itself only
                     // defined on [cmp0, cmp1) -> one more row is lost on every interior side
};

static int flush_pending(dd_batch* b, dd_step_stats* stats, int* have);  // verifies a deferred step
static void free_tables(dd_batch* b);

static void reset_ctl(dd_batch* b, bool all) {
    for (int m = 0; m < 2; ++m)
        for (int q = 0; q < 3; ++q) {
            b->ctl[m].sweeps[q] = 0;
            if (all) {
                b->ctl[m].extra[q] = 0;
                b->ctl[m].floor[q] = 1;
            }
        }
    if (all) {
        b->cm = 0;
        b->gs_only[0] = b->gs_only[1] = b->gs_only[2] = false;
    }
}


// ---------------------------------------------------------------------------
// launch accounting: every kernel launched by the library is counted; with profiling on,
// each launch group is bracketed by CUDA events on the context's stream and its device time is
// accumulated per kernel class (bench.py reads this for the roofline of the dominant kernel).
enum ProfClass {
    PC_TIME = 0, PC_PREDICT, PC_ASM_T, PC_ASM_CL, PC_ASM_CD, PC_SOLVE_T, PC_SOLVE_CL, PC_SOLVE_CD, PC_CORRECT,
    PC_CS_FINISH, PC_SUMMARISE, PC_FEULER, PC_NORMS, PC_OTHER, PC_SOURCES, PC_NCLASS
};
static const char* kProfNames[PC_NCLASS] = {"k_time_coefs", "k_predict", "k_assemble<T>", "k_assemble<cl>",
                                            "k_assemble<cd>", "k_rbsor_tile<T>", "k_rbsor_tile<cl>",
                                            "k_rbsor_tile<cd>", "k_correct", "k_cs_decide+k_cs_redo",
                                            "k_summarise", "k_feuler", "k_error_norms", "other", "k_eval_sources"};
// shared by every context of the process; contexts may be driven from different host threads
struct Prof {
    bool on = false;
    std::atomic<long long> launches{0};
    std::mutex mu;  // guards pool / recs while profiling is on
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    struct Rec { int cls; cudaEvent_t a, b; int n; };
    std::vector<Rec> recs;
    double ms[PC_NCLASS] = {0};
    long long count[PC_NCLASS] = {0};
    cudaEvent_t get() {
        if (used == pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            pool.push_back(e);
        }
        return pool[used++];
    }
};
static Prof g_prof;

struct ProfScope {
    cudaStream_t st;
    int cls, n;
    cudaEvent_t a;
    bool on;
    ProfScope(cudaStream_t s, int c, int nl) : st(s), cls(c), n(nl), on(g_prof.on) {
        g_prof.launches += nl;
        if (on) {
            std::lock_guard<std::mutex> lk(g_prof.mu);
            a = g_prof.get();
            cudaEventRecord(a, st);
        }
    }
    ~ProfScope() {
        if (on) {
            std::lock_guard<std::mutex> lk(g_prof.mu);
            cudaEvent_t b = g_prof.get();
            cudaEventRecord(b, st);
            g_prof.recs.push_back({cls, a, b, n});
        }
    }
};
#define CKP(cls, nl, call) do { ProfScope ps_(ctx->stream, cls, nl); CK(call); } while (0)

extern "C" long long dd_launch_count(void) { return g_prof.launches.load(); }

extern "C" int dd_profile_enable(int on) {
    g_prof.on = on != 0;
    return DD_OK;
}

// drains the recorded event pairs (synchronises the events) and returns the accumulated device time
// per kernel class; names/ms/count arrays of length >= 16; resets the accumulators when `reset`
extern "C" int dd_profile_read(const char** names, double* ms, long long* count, int reset) {
    std::lock_guard<std::mutex> lk(g_prof.mu);
    for (auto& r : g_prof.recs) {
        cudaEventSynchronize(r.b);
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        g_prof.ms[r.cls] += t;
        g_prof.count[r.cls] += r.n;
    }
    g_prof.recs.clear();
    g_prof.used = 0;
    for (int c = 0; c < PC_NCLASS; ++c) {
        if (names) names[c] = kProfNames[c];
        if (ms) ms[c] = g_prof.ms[c];
        if (count) count[c] = g_prof.count[c];
        if (reset) { g_prof.ms[c] = 0; g_prof.count[c] = 0; }
    }
    return PC_NCLASS;
}

// ---------------------------------------------------------------------------
static int fail(dd_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(ctx, DD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));           \
    } while (0)

extern "C" const char* dd_version(void) { return DD_VERSION_STR; }

extern "C" const char* dd_last_error(const dd_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int dd_ctx_create(int device, void* cuda_stream, dd_ctx** out) {
    if (!out) return DD_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return DD_ERR_NO_DEVICE;
    if (device < 0 || device >= n) return DD_ERR_INVALID;
    dd_ctx* ctx = new dd_ctx();
    ctx->device = device;
    ctx->own_stream = false;
    if (cudaSetDevice(device) != cudaSuccess) {
        delete ctx;
        return DD_ERR_CUDA;
    }
    if (cuda_stream) {
        ctx->stream = (cudaStream_t)cuda_stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return DD_ERR_CUDA;
        }
        ctx->own_stream = true;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    ctx->sm_count = prop.multiProcessorCount;
    if (dd_solver_configure() != cudaSuccess || dd_wave_configure() != cudaSuccess ||
        dd_lane_configure() != cudaSuccess) {
        cudaGetLastError();
        delete ctx;
        return DD_ERR_CUDA;
    }
    *out = ctx;
    return DD_OK;
}

static void ctx_free(dd_ctx* ctx) {
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// Handles may be released in any order (garbage-collected hosts do): a context outlives its batches.
extern "C" int dd_ctx_destroy(dd_ctx* ctx) {
    if (!ctx) return DD_OK;
    if (ctx->nbatches > 0) {
        ctx->dead = true;
        return DD_OK;
    }
    ctx_free(ctx);
    return DD_OK;
}

extern "C" int dd_ctx_synchronize(dd_ctx* ctx) {
    if (!ctx) return DD_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return DD_OK;
}

// ---------------------------------------------------------------------------
static int upload_vec(dd_ctx* ctx, const std::vector<double>& h, double** d, std::vector<double*>& keep) {
    CK(cudaMalloc((void**)d, h.size() * sizeof(double)));
    keep.push_back(*d);
    CK(cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DD_OK;
}

static int get_work(dd_batch* b, const char* name, double** out) {
    dd_ctx* ctx = b->ctx;
    auto it = b->work.find(name);
    if (it != b->work.end()) {
        *out = it->second;
        return DD_OK;
    }
    double* p = nullptr;
    // work fields have room for the even row pitch of the Newton rows / solver iterates (ld + 1 when ld is odd)
    const size_t padded = (size_t)b->B * b->nrows * (size_t)(b->g.ld + (b->g.ld & 1));
    CK(cudaMalloc((void**)&p, padded * sizeof(double)));
    CK(cudaMemsetAsync(p, 0, padded * sizeof(double), ctx->stream));
    b->work[name] = p;
    *out = p;
    return DD_OK;
}

static void model_to_dev(const dd_model& s, DDModel* d) {
    d->K1 = s.K1; d->K2 = s.K2; d->K3 = s.K3; d->K4 = s.K4; d->DT = s.DT; d->Dl_max = s.Dl_max;
    d->phi_l = s.phi_l; d->gamma_T = s.gamma_T; d->Kd = s.Kd; d->Sd = s.Sd; d->Dd_max = s.Dd_max;
    d->phi_d = s.phi_d; d->phi_T = s.phi_T; d->r_sp = s.r_sp;
    d->T_shift = (s.kind == 2) ? s.T_ref : 0.0;
    d->eta = s.eta;
    d->react = (s.reaction == 1) ? DD_REACT_CS : (s.reaction == 2 ? DD_REACT_H : DD_REACT_REGH);
    d->_pad = 0;
}

extern "C" int dd_batch_create(dd_ctx* ctx, int N, int M, const double* x, const double* y, int nmembers, int row0,
                               int nrows, int own0, int own1, int nslots, dd_batch** out) {
    if (!ctx || !out || !x || !y) return DD_ERR_INVALID;
    *out = nullptr;
    if (N < 2 || M < 2 || nmembers < 1 || nslots < 2) return fail(ctx, DD_ERR_INVALID, "need N, M >= 2, members >= 1, slots >= 2");
    if (row0 < 0 || nrows < 1 || row0 + nrows > N + 1) return fail(ctx, DD_ERR_INVALID, "row slab outside the grid");
    if (own0 < 0 || own1 > nrows || own0 >= own1) return fail(ctx, DD_ERR_INVALID, "bad owned row range");
    CK(cudaSetDevice(ctx->device));
    dd_batch* b = new dd_batch();
    b->ctx = ctx;
    b->N = N; b->M = M; b->B = nmembers; b->row0 = row0; b->nrows = nrows; b->own0 = own0; b->own1 = own1;
    b->nslots = nslots;
    b->is_slab = !(row0 == 0 && nrows == N + 1);
    b->cmp0 = (row0 == 0) ? 0 : 1;
    b->cmp1 = (row0 + nrows == N + 1) ? nrows : nrows - 1;
    b->asm0 = (row0 == 0) ? 0 : 2;
    b->asm1 = (row0 + nrows == N + 1) ? nrows : nrows - 2;
    reset_ctl(b, true);
    b->prev_valid = b->use_guess = b->phase_fused_T = false;
    b->relax_rho[0] = b->relax_rho[1] = b->relax_rho[2] = -1.0;
    b->gs_only[0] = b->gs_only[1] = b->gs_only[2] = false;
    b->prev_in = b->prev_out = -1;
    b->prev_dt = b->cur_dt = 0.0;
    b->rec_cur = 0;
    const int ld = M + 1;
    b->field_elems = (size_t)nmembers * nrows * ld;
    // geometry (reference Grid.__init__, src/prob1base.py:287-304)
    std::vector<double> hx(N + 1), hk(M + 1), hp(N + 1), kp(M + 1), rh(N + 1), rk(M + 1), rhp(N + 1), rkp(M + 1);
    std::vector<double> xs(x, x + N + 1), ys(y, y + M + 1);
    hx[0] = INFINITY; hk[0] = INFINITY;
    for (int i = 1; i <= N; ++i) hx[i] = x[i] - x[i - 1];
    for (int j = 1; j <= M; ++j) hk[j] = y[j] - y[j - 1];
    for (int i = 0; i < N; ++i) hp[i] = (hx[i] + hx[i + 1]) * 0.5;
    hp[N] = INFINITY;
    for (int j = 0; j < M; ++j) kp[j] = (hk[j] + hk[j + 1]) * 0.5;
    kp[M] = INFINITY;
    for (int i = 0; i <= N; ++i) { rh[i] = (i == 0) ? 0.0 : 1.0 / hx[i]; rhp[i] = (i == 0 || i == N) ? 0.0 : 1.0 / hp[i]; }
    for (int j = 0; j <= M; ++j) { rk[j] = (j == 0) ? 0.0 : 1.0 / hk[j]; rkp[j] = (j == 0 || j == M) ? 0.0 : 1.0 / kp[j]; }
    DDGeom& g = b->g;
    g.N = N; g.M = M; g.row0 = row0; g.nrows = nrows; g.ld = ld; g.mstride = (long long)nrows * ld;
    double* d;
    int rc;
#define UP(vec, field) if ((rc = upload_vec(ctx, vec, &d, b->geo_dev)) != DD_OK) return rc; g.field = d;
    UP(xs, x) UP(ys, y) UP(hx, h) UP(hk, k) UP(hp, hp) UP(kp, kp) UP(rh, rh) UP(rk, rk) UP(rhp, rhp) UP(rkp, rkp)
#undef UP
    // members
    b->h_mem.resize(nmembers);
    memset(b->h_mem.data(), 0, sizeof(DDMember) * nmembers);
    for (auto& mb : b->h_mem) mb.active = 1;
    CK(cudaMalloc((void**)&b->d_mem, sizeof(DDMember) * nmembers));
    CK(cudaMemcpyAsync(b->d_mem, b->h_mem.data(), sizeof(DDMember) * nmembers, cudaMemcpyHostToDevice, ctx->stream));
    b->mode = DD_FORCING_NONE;
    memset(&b->F, 0, sizeof(b->F));
    b->smode = DD_FORCING_NONE;
    memset(&b->sF, 0, sizeof(b->sF));
    {
        const char* e = getenv("DD_FUSED_SOURCES");
        b->fused_sources = e && atoi(e) != 0;
    }
    memset(b->src_set, 0, sizeof(b->src_set));
    b->src_has[0] = b->src_has[1] = false;
    b->slots.resize(nslots);
    for (int s = 0; s < nslots; ++s) {
        b->slots[s].resize(DD_NVAR);
        for (int v = 0; v < DD_NVAR; ++v) {
            CK(cudaMalloc((void**)&b->slots[s][v], b->field_elems * sizeof(double)));
            CK(cudaMemsetAsync(b->slots[s][v], 0, b->field_elems * sizeof(double), ctx->stream));
        }
    }
    CK(cudaMalloc((void**)&b->d_t0, sizeof(double) * nmembers));
    CK(cudaMalloc((void**)&b->d_dt, sizeof(double) * nmembers));
    b->nsolve_cap = 0; b->d_stats = nullptr; b->d_summary = nullptr;
    b->d_itmax = b->d_itmin = nullptr; b->d_used = nullptr; b->cs_cap_alloc = 0;
    b->has_variants = false;
    CK(cudaMalloc((void**)&b->d_flags, sizeof(int) * nmembers));
    CK(cudaMemsetAsync(b->d_flags, 0, sizeof(int) * nmembers, ctx->stream));
    b->norm_bpm = dd_norm_blocks_per_member(g);
    CK(cudaMalloc((void**)&b->d_norm_partial, sizeof(double) * 8 * (size_t)b->norm_bpm * nmembers));
    CK(cudaMalloc((void**)&b->d_norm_out, sizeof(double) * 8 * nmembers));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->nbatches += 1;
    *out = b;
    return DD_OK;
}

extern "C" int dd_batch_destroy(dd_batch* b) {
    if (!b) return DD_OK;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    for (double* p : b->geo_dev) cudaFree(p);
    free_tables(b);
    cudaFree(b->d_prog_mem);
    for (auto& s : b->slots) for (double* p : s) cudaFree(p);
    for (auto& kv : b->work) cudaFree(kv.second);
    for (auto& r : b->rec) {
        if (r.h_sums) cudaFreeHost(r.h_sums);
        if (r.h_used) cudaFreeHost(r.h_used);
        if (r.h_flags) cudaFreeHost(r.h_flags);
        if (r.done) cudaEventDestroy(r.done);
    }
    cudaFree(b->d_mem); cudaFree(b->d_t0); cudaFree(b->d_dt); cudaFree(b->d_stats); cudaFree(b->d_summary);
    cudaFree(b->d_itmax); cudaFree(b->d_itmin); cudaFree(b->d_used); cudaFree(b->d_norm_partial);
    cudaFree(b->d_norm_out);
    cudaFree(b->d_combine_state); cudaFree(b->d_combined); cudaFree(b->d_member_stats);
    cudaFree(b->d_flags);
    dd_ctx* ctx = b->ctx;
    delete b;
    if (--ctx->nbatches <= 0 && ctx->dead) ctx_free(ctx);
    return DD_OK;
}

static int push_members(dd_batch* b, int first, int count) {
    dd_ctx* ctx = b->ctx;
    b->src_has[0] = b->src_has[1] = false;  // models / time profiles changed: staged sources are stale
    b->prev_valid = false;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(b->d_mem + first, b->h_mem.data() + first, sizeof(DDMember) * count, cudaMemcpyHostToDevice,
                       ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DD_OK;
}

// the device copy carries t0/dt/tc written by kernels; refresh the host mirror's
// static part only (model, phi, active) -- kernels never change those.
extern "C" int dd_batch_set_models(dd_batch* b, int first, int count, const dd_model* models) {
    if (!b || !models || first < 0 || count < 1 || first + count > b->B) return DD_ERR_INVALID;
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    for (int k = 0; k < count; ++k) model_to_dev(models[k], &b->h_mem[first + k].m);
    b->has_variants = false;
    for (int m = 0; m < b->B; ++m)
        if (b->h_mem[m].m.react != DD_REACT_REGH) b->has_variants = true;
    reset_ctl(b, true);
    return push_members(b, first, count);
}

extern "C" int dd_batch_set_active(dd_batch* b, int first, int count, const int* active) {
    if (!b || !active || first < 0 || count < 1 || first + count > b->B) return DD_ERR_INVALID;
    for (int k = 0; k < count; ++k) b->h_mem[first + k].active = active[k] ? 1 : 0;
    return push_members(b, first, count);
}

// ---------------------------------------------------------------------------
// forcing
// ---------------------------------------------------------------------------
static void free_tables(dd_batch* b) {
    b->src_has[0] = b->src_has[1] = false;
    for (double* p : b->table_dev) cudaFree(p);
    b->table_dev.clear();
    memset(&b->F.tab, 0, sizeof(b->F.tab));
    if (b->prog_lib) cudaLibraryUnload(b->prog_lib);
    b->prog_lib = nullptr;
    b->prog_kernel = nullptr;
    b->d_prog_xq = b->d_prog_yq = nullptr;
}

static int up_table(dd_batch* b, const double* h, size_t n, const double** dst) {
    dd_ctx* ctx = b->ctx;
    double* d = nullptr;
    CK(cudaMalloc((void**)&d, n * sizeof(double)));
    b->table_dev.push_back(d);
    CK(cudaMemcpyAsync(d, h, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    *dst = d;
    return DD_OK;
}

extern "C" int dd_forcing_none(dd_batch* b) {
    if (!b) return DD_ERR_INVALID;
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    free_tables(b);
    b->mode = DD_FORCING_NONE;
    return DD_OK;
}

extern "C" int dd_forcing_separable(dd_batch* b, int nterms, const double* const X[5][3],
                                    const double* const Y[5][3], const double* const XQ[3],
                                    const double* const YQ[3], const int phi_kind[5], const double phi_p[5][4]) {
    if (!b || !X || !Y || !XQ || !YQ || !phi_kind || !phi_p || nterms < 1 || nterms > 16) return DD_ERR_INVALID;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    CK(cudaStreamSynchronize(ctx->stream));
    free_tables(b);
    int rc;
    for (int v = 0; v < 5; ++v)
        for (int d = 0; d < 3; ++d)
            if (!X[v][d] || !Y[v][d]) return fail(ctx, DD_ERR_INVALID, "null separable table");
    for (int q = 0; q < 3; ++q)
        if (!XQ[q] || !YQ[q]) return fail(ctx, DD_ERR_INVALID, "null quadrature table");
    DDHostTables ht;
    dd_prepare_tables(nterms, b->N, b->M, X, Y, XQ, YQ, &ht);
    DDTables& tb = b->F.tab;
    for (int p = 0; p < ht.nprof; ++p)
        for (int d = 0; d < 3; ++d) {
            if ((rc = up_table(b, ht.X[p][d].data(), ht.X[p][d].size(), &tb.X[p][d])) != DD_OK) return rc;
            if ((rc = up_table(b, ht.Y[p][d].data(), ht.Y[p][d].size(), &tb.Y[p][d])) != DD_OK) return rc;
        }
    if ((rc = up_table(b, ht.QX1.data(), ht.QX1.size(), &tb.QX1)) != DD_OK) return rc;
    if ((rc = up_table(b, ht.QY1.data(), ht.QY1.size(), &tb.QY1)) != DD_OK) return rc;
    if ((rc = up_table(b, ht.QX2.data(), ht.QX2.size(), &tb.QX2)) != DD_OK) return rc;
    if ((rc = up_table(b, ht.QY2.data(), ht.QY2.size(), &tb.QY2)) != DD_OK) return rc;
    if ((rc = up_table(b, ht.QX3.data(), ht.QX3.size(), &tb.QX3)) != DD_OK) return rc;
    if ((rc = up_table(b, ht.QY3.data(), ht.QY3.size(), &tb.QY3)) != DD_OK) return rc;
    CK(cudaStreamSynchronize(ctx->stream));  // ht goes out of scope: the copies must have left the host
    tb.nterms = nterms;
    tb.nx = b->N + 1;
    tb.ny = b->M + 1;
    tb.nprof = ht.nprof;
    for (int v = 0; v < 5; ++v) tb.var_prof[v] = ht.var_prof[v];
    for (auto& mb : b->h_mem)
        for (int v = 0; v < 5; ++v) {
            mb.phi_kind[v] = phi_kind[v];
            for (int k = 0; k < 4; ++k) mb.phi_p[v][k] = phi_p[v][k];
        }
    b->mode = DD_FORCING_SEPARABLE;
    return push_members(b, 0, b->B);
}

extern "C" int dd_forcing_set_phi(dd_batch* b, int first, int count, const int* phi_kind, const double* phi_p) {
    if (!b || !phi_kind || !phi_p || first < 0 || count < 1 || first + count > b->B) return DD_ERR_INVALID;
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    for (int k = 0; k < count; ++k)
        for (int v = 0; v < 5; ++v) {
            b->h_mem[first + k].phi_kind[v] = phi_kind[k * 5 + v];
            for (int q = 0; q < 4; ++q) b->h_mem[first + k].phi_p[v][q] = phi_p[(k * 5 + v) * 4 + q];
        }
    return push_members(b, first, count);
}

extern "C" int dd_forcing_expsin(dd_batch* b, const double* sx, const double* cx, const double* sy, const double* cy,
                                 const double* sxq, const double* syq) {
    if (!b || !sx || !cx || !sy || !cy || !sxq || !syq) return DD_ERR_INVALID;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    CK(cudaStreamSynchronize(ctx->stream));
    free_tables(b);
    int rc;
    if ((rc = up_table(b, sx, b->N + 1, &b->F.tab.X[0][0])) != DD_OK) return rc;
    if ((rc = up_table(b, cx, b->N + 1, &b->F.tab.X[0][1])) != DD_OK) return rc;
    if ((rc = up_table(b, sy, b->M + 1, &b->F.tab.Y[0][0])) != DD_OK) return rc;
    if ((rc = up_table(b, cy, b->M + 1, &b->F.tab.Y[0][1])) != DD_OK) return rc;
    if ((rc = up_table(b, sxq, 3 * (size_t)(b->N + 1), &b->F.tab.XQ0)) != DD_OK) return rc;
    if ((rc = up_table(b, syq, 3 * (size_t)(b->M + 1), &b->F.tab.YQ0)) != DD_OK) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    b->F.tab.nterms = 1;
    b->F.tab.nx = b->N + 1;
    b->F.tab.ny = b->M + 1;
    b->mode = DD_FORCING_EXPSIN;
    return DD_OK;
}

// ---------------------------------------------------------------------------
// generated forcing (include/dd_b200_program.h)
// ---------------------------------------------------------------------------
__global__ void k_pack_program_members(const DDMember* mem, dd_program_member* out, int nmem) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmem) return;
    const DDMember& mb = mem[m];
    dd_program_member o;
    o.model.K1 = mb.m.K1; o.model.K2 = mb.m.K2; o.model.K3 = mb.m.K3; o.model.K4 = mb.m.K4;
    o.model.DT = mb.m.DT; o.model.Dl_max = mb.m.Dl_max; o.model.phi_l = mb.m.phi_l; o.model.gamma_T = mb.m.gamma_T;
    o.model.Kd = mb.m.Kd; o.model.Sd = mb.m.Sd; o.model.Dd_max = mb.m.Dd_max; o.model.phi_d = mb.m.phi_d;
    o.model.phi_T = mb.m.phi_T; o.model.r_sp = mb.m.r_sp; o.model.T_ref = mb.m.T_shift;
    o.model.eta = mb.m.eta; o.model.kind = 2; o.model.reaction = mb.m.react;
    o.t[0] = mb.t0;
    o.t[1] = mb.t0 + mb.dt;
    o.active = mb.active;
    o._pad = 0;
    out[m] = o;
}

// one launch of the program: `what` into the five arrays of `out`, at the members' current t0 (tslot 0) or
// t0 + dt (tslot 1); rows: all local rows
static int launch_program(dd_batch* b, int what, int tslot, const DDState& out) {
    dd_ctx* ctx = b->ctx;
    if (!b->prog_kernel) return fail(ctx, DD_ERR_INVALID, "no forcing program loaded");
    if (b->B > 65535 || b->nrows > 65535) return fail(ctx, DD_ERR_INVALID, "forcing program: more than 65535 members or rows");
    k_pack_program_members<<<(b->B + 127) / 128, 128, 0, ctx->stream>>>(b->d_mem, b->d_prog_mem, b->B);
    CK(cudaGetLastError());
    dd_program_args a;
    memset(&a, 0, sizeof(a));
    a.x = b->g.x; a.y = b->g.y; a.xq = b->d_prog_xq; a.yq = b->d_prog_yq;
    a.members = b->d_prog_mem;
    for (int v = 0; v < DD_NVAR; ++v) a.out[v] = out.v[v];
    a.mstride = b->g.mstride;
    a.N = b->N; a.M = b->M; a.row0 = b->row0; a.nrows = b->nrows; a.ld = b->g.ld; a.nmembers = b->B;
    a.what = what; a.tslot = tslot;
    void* params[1] = {&a};
    const dim3 grid((unsigned)((b->M + 1 + 127) / 128), (unsigned)b->nrows, (unsigned)b->B);
    const int phase = what == DD_PROGRAM_SOURCES ? PC_SOURCES : PC_OTHER;
    CKP(phase, 2, cudaLaunchKernel((const void*)b->prog_kernel, grid, dim3(128, 1, 1), params, 0, ctx->stream));
    return DD_OK;
}

extern "C" int dd_forcing_program(dd_batch* b, const void* image, unsigned long long image_bytes, const double* xq,
                                  const double* yq) {
    if (!b || !image || image_bytes == 0 || !xq || !yq) return DD_ERR_INVALID;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    CK(cudaStreamSynchronize(ctx->stream));
    free_tables(b);
    b->mode = DD_FORCING_NONE;
    // the loader keeps no reference to `image` after the call (it is copied into the library object)
    cudaLibrary_t lib = nullptr;
    cudaError_t e = cudaLibraryLoadData(&lib, image, nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, DD_ERR_INVALID, (std::string("forcing program: image rejected: ") + cudaGetErrorString(e)).c_str());
    }
    cudaKernel_t k = nullptr;
    e = cudaLibraryGetKernel(&k, lib, "dd_program");
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaLibraryUnload(lib);
        return fail(ctx, DD_ERR_INVALID, "forcing program: the image does not define the kernel dd_program");
    }
    b->prog_lib = lib;
    b->prog_kernel = k;
    int rc;
    if ((rc = up_table(b, xq, (size_t)(b->N + 1) * 3, &b->d_prog_xq)) != DD_OK) return rc;
    if ((rc = up_table(b, yq, (size_t)(b->M + 1) * 3, &b->d_prog_yq)) != DD_OK) return rc;
    if (!b->d_prog_mem) CK(cudaMalloc((void**)&b->d_prog_mem, sizeof(dd_program_member) * b->B));
    CK(cudaStreamSynchronize(ctx->stream));
    b->mode = DD_FORCING_PROGRAM;
    return DD_OK;
}

static const char* kForcingNames[5][2] = {{"f_cp0", "f_cp1"}, {"f_T0", "f_T1"}, {"f_cl0", "f_cl1"},
                                          {"f_cd0", "f_cd1"}, {"f_cs0", "f_cs1"}};

extern "C" int dd_forcing_arrays(dd_batch* b, int member, const double* const f[5][2]) {
    if (!b || !f || member < 0 || member >= b->B) return DD_ERR_INVALID;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    if (b->mode != DD_FORCING_ARRAYS) {
        CK(cudaStreamSynchronize(ctx->stream));
        free_tables(b);
        b->mode = DD_FORCING_ARRAYS;
    }
    const size_t per = (size_t)b->nrows * b->g.ld;
    for (int v = 0; v < 5; ++v)
        for (int s = 0; s < 2; ++s) {
            double* d;
            int rc = get_work(b, kForcingNames[v][s], &d);
            if (rc != DD_OK) return rc;
            b->F.arr.f[v][s] = d;
            if (f[v][s])
                CK(cudaMemcpyAsync(d + member * per, f[v][s], per * sizeof(double), cudaMemcpyHostToDevice,
                                   ctx->stream));
            else
                CK(cudaMemsetAsync(d + member * per, 0, per * sizeof(double), ctx->stream));
        }
    CK(cudaStreamSynchronize(ctx->stream));
    return DD_OK;
}

// ---------------------------------------------------------------------------
// state
// ---------------------------------------------------------------------------
static bool slot_ok(const dd_batch* b, int s) { return s >= 0 && s < b->nslots; }

extern "C" int dd_state_upload(dd_batch* b, int slot, int member, const double* const fields[5]) {
    if (!b || !fields || !slot_ok(b, slot) || member < 0 || member >= b->B) return DD_ERR_INVALID;
    b->prev_valid = false;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    const size_t per = (size_t)b->nrows * b->g.ld;
    for (int v = 0; v < 5; ++v)
        if (fields[v])
            CK(cudaMemcpyAsync(b->slots[slot][v] + member * per, fields[v], per * sizeof(double),
                               cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DD_OK;
}

extern "C" int dd_state_download(dd_batch* b, int slot, int member, double* const fields[5]) {
    if (!b || !fields || !slot_ok(b, slot) || member < 0 || member >= b->B) return DD_ERR_INVALID;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    const size_t per = (size_t)b->nrows * b->g.ld;
    for (int v = 0; v < 5; ++v)
        if (fields[v])
            CK(cudaMemcpyAsync(fields[v], b->slots[slot][v] + member * per, per * sizeof(double),
                               cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DD_OK;
}

extern "C" int dd_state_dev_ptr(dd_batch* b, int slot, int var, void** ptr, long long* member_stride, int* ld) {
    if (!b || !ptr || !slot_ok(b, slot) || var < 0 || var > 4) return DD_ERR_INVALID;
    *ptr = b->slots[slot][var];
    if (member_stride) *member_stride = b->g.mstride;
    if (ld) *ld = b->g.ld;
    return DD_OK;
}

static int ensure_cs_buffers(dd_batch* b, int cap);
static int ensure_solve_slots(dd_batch* b, int n);

extern "C" int dd_work_dev_ptr(dd_batch* b, const char* name, void** ptr) {
    if (!b || !name || !ptr) return DD_ERR_INVALID;
    if (!strcmp(name, "solve_stats")) {
        // DDSolveStats[3][nmembers] of the phased step: 5 doubles each, rho first (all-reduced by the slab driver)
        int rc0 = ensure_solve_slots(b, 4);
        if (rc0 != DD_OK) return rc0;
        *ptr = b->d_stats;
        return DD_OK;
    }
    if (!strcmp(name, "summary")) {
        // SolveSummary[4] of the phased step: rho, ratio, resid, bound per solve, then a row whose first entry is the
        // corrector's domain-error flag (all-reduced by the slab driver)
        int rc0 = ensure_solve_slots(b, 4);
        if (rc0 != DD_OK) return rc0;
        *ptr = b->d_summary;
        return DD_OK;
    }
    if (!strcmp(name, "x_cur")) {
        // the iterate a solve segment (dd_pc_solve_segment, last = 0) left behind: [member][nrows][even pitch]
        if (!b->x_cur) return fail(b->ctx, DD_ERR_INVALID, "no solve segment is pending");
        *ptr = const_cast<double*>(b->x_cur);
        return DD_OK;
    }
    if (!strcmp(name, "cs_used")) {
        // int[nmembers]: cs-Newton iterations used (phase 5 / 6)
        if (!b->d_used) return fail(b->ctx, DD_ERR_INVALID, "cs statistics not allocated yet");
        *ptr = b->d_used;
        return DD_OK;
    }
    if (!strcmp(name, "cs_it_max") || !strcmp(name, "cs_it_min")) {
        // [member][cap] accumulators of the cs-Newton exit test (allreduced by the slab driver)
        if (b->cs_cap_alloc <= 0) return fail(b->ctx, DD_ERR_INVALID, "cs statistics not allocated yet");
        *ptr = !strcmp(name, "cs_it_max") ? b->d_itmax : b->d_itmin;
        return DD_OK;
    }
    double* p;
    int rc = get_work(b, name, &p);
    if (rc != DD_OK) return rc;
    *ptr = p;
    return DD_OK;
}

extern "C" int dd_work_upload(dd_batch* b, const char* name, int member, const double* host) {
    if (!b || !name || !host || member < 0 || member >= b->B) return DD_ERR_INVALID;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    double* p;
    int rc = get_work(b, name, &p);
    if (rc != DD_OK) return rc;
    const size_t per = (size_t)b->nrows * b->g.ld;
    CK(cudaMemcpyAsync(p + member * per, host, per * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DD_OK;
}

extern "C" int dd_work_download(dd_batch* b, const char* name, int member, double* host) {
    if (!b || !name || !host || member < 0 || member >= b->B) return DD_ERR_INVALID;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    double* p;
    int rc = get_work(b, name, &p);
    if (rc != DD_OK) return rc;
    const size_t per = (size_t)b->nrows * b->g.ld;
    CK(cudaMemcpyAsync(host, p + member * per, per * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DD_OK;
}

static DDStateC cstate(const dd_batch* b, int slot) {
    DDStateC s;
    for (int v = 0; v < DD_NVAR; ++v) s.v[v] = b->slots[slot][v];
    return s;
}
static DDState mstate(const dd_batch* b, int slot) {
    DDState s;
    for (int v = 0; v < DD_NVAR; ++v) s.v[v] = b->slots[slot][v];
    return s;
}
// row ranges: OWNED rows (tile solver), STENCIL rows (every row with a complete stencil: on a slab
// the halo rows are recomputed redundantly), ALL rows (pointwise work)
enum RowRange { ROWS_OWNED = 0, ROWS_ALL = 1, ROWS_STENCIL = 2, ROWS_ASM = 3 };
static DDLaunch launch_of(const dd_batch* b, int range = ROWS_STENCIL) {
    DDLaunch L;
    L.stream = b->ctx->stream;
    L.nmembers = b->B;
    if (range == ROWS_ALL) { L.own0 = 0; L.own1 = b->nrows; }
    else if (range == ROWS_STENCIL) { L.own0 = b->cmp0; L.own1 = b->cmp1; }
    else if (range == ROWS_ASM) { L.own0 = b->asm0; L.own1 = b->asm1; }
    else { L.own0 = b->own0; L.own1 = b->own1; }
    L.vr0 = b->asm0;
    L.vr1 = b->asm1;
    return L;
}

static int set_times(dd_batch* b, const double* t0, const double* dt, int n_t) {
    dd_ctx* ctx = b->ctx;
    if (!t0 || !dt || (n_t != 1 && n_t != b->B)) return fail(ctx, DD_ERR_INVALID, "t0/dt: need 1 or nmembers values");
    for (int k = 0; k < n_t; ++k)
        if (!(dt[k] > 0.0)) return fail(ctx, DD_ERR_INVALID, "dt must be > 0");
    b->cur_dt = dt[0];
    CK(cudaMemcpyAsync(b->d_t0, t0, sizeof(double) * n_t, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(b->d_dt, dt, sizeof(double) * n_t, cudaMemcpyHostToDevice, ctx->stream));
    CKP(PC_TIME, 1, dd_launch_time_coefs(launch_of(b), b->mode, b->d_mem, b->d_t0, b->d_dt, n_t, 0));
    return DD_OK;
}

// Select what the step kernels are launched with.  Fused modes (SEPARABLE / EXPSIN) are staged: the sources
// at t0 and t1 are evaluated into two array sets unless a set already holds that time level (uniform times
// only).  Must be called after the member times are on the device (set_times / advance).
// `carry` (run loops with per-member times, where the time-keyed cache cannot be used): in: the set that
// already holds every member's sources at this step's t0 (-1: none), out: the set holding them at t1
static int stage_sources(dd_batch* b, double t0, double dt, bool times_uniform, bool need_slot1,
                         int* carry = nullptr) {
    dd_ctx* ctx = b->ctx;
    const bool program = b->mode == DD_FORCING_PROGRAM;
    if (!program && ((b->mode != DD_FORCING_SEPARABLE && b->mode != DD_FORCING_EXPSIN) || b->fused_sources)) {
        b->smode = b->mode;
        b->sF = b->F;
        return DD_OK;
    }
    static const char* names[2][DD_NVAR] = {{"src0_cp", "src0_T", "src0_cl", "src0_cd", "src0_cs"},
                                            {"src1_cp", "src1_T", "src1_cl", "src1_cd", "src1_cs"}};
    int rc;
    for (int s = 0; s < 2; ++s)
        for (int v = 0; v < DD_NVAR; ++v)
            if (!b->src_set[s][v] && (rc = get_work(b, names[s][v], &b->src_set[s][v])) != DD_OK) return rc;
    if (!times_uniform) b->src_has[0] = b->src_has[1] = false;
    const double t1 = t0 + dt;
    int s0 = -1;
    for (int s = 0; s < 2; ++s)
        if (b->src_has[s] && b->src_time[s] == t0) s0 = s;
    if (s0 < 0 && carry && *carry >= 0) {
        s0 = *carry;
    } else if (s0 < 0) {
        s0 = 0;
        if (need_slot1 && b->src_has[0] && b->src_time[0] == t1) s0 = 1;  // keep a set that already holds t1
        DDState out;
        for (int v = 0; v < DD_NVAR; ++v) out.v[v] = b->src_set[s0][v];
        if (program) {
            if ((rc = launch_program(b, DD_PROGRAM_SOURCES, 0, out)) != DD_OK) return rc;
        } else {
            CKP(PC_SOURCES, 1, dd_launch_eval_sources(launch_of(b, ROWS_ALL), b->mode, b->g, b->d_mem, b->F, out, 0));
        }
        b->src_has[s0] = times_uniform;
        b->src_time[s0] = t0;
    }
    const int s1 = 1 - s0;
    if (need_slot1 && !(b->src_has[s1] && b->src_time[s1] == t1)) {
        DDState out;
        for (int v = 0; v < DD_NVAR; ++v) out.v[v] = b->src_set[s1][v];
        if (program) {
            if ((rc = launch_program(b, DD_PROGRAM_SOURCES, 1, out)) != DD_OK) return rc;
        } else {
            CKP(PC_SOURCES, 1, dd_launch_eval_sources(launch_of(b, ROWS_ALL), b->mode, b->g, b->d_mem, b->F, out, 1));
        }
        b->src_has[s1] = times_uniform;
        b->src_time[s1] = t1;
    }
    if (carry) *carry = need_slot1 ? s1 : -1;
    b->smode = DD_FORCING_ARRAYS;
    memset(&b->sF, 0, sizeof(b->sF));
    for (int v = 0; v < DD_NVAR; ++v) {
        b->sF.arr.f[v][0] = b->src_set[s0][v];
        b->sF.arr.f[v][1] = b->src_set[s1][v];
    }
    return DD_OK;
}

extern "C" int dd_state_fill_exact(dd_batch* b, int slot, const double* t, int n_t) {
    if (!b || !slot_ok(b, slot) || !t) return DD_ERR_INVALID;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    std::vector<double> one(n_t, 1.0);
    int rc = set_times(b, t, one.data(), n_t);
    if (rc != DD_OK) return rc;
    b->prev_valid = false;
    if (b->mode == DD_FORCING_PROGRAM) {
        if ((rc = launch_program(b, DD_PROGRAM_EXACT, 0, mstate(b, slot))) != DD_OK) return rc;
    } else {
        CKP(PC_OTHER, 1, dd_launch_fill_exact(launch_of(b, ROWS_ALL), b->mode, b->g, b->d_mem, b->F, mstate(b, slot)));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return DD_OK;
}

// ---------------------------------------------------------------------------
// forward Euler, field evaluation, norms
// ---------------------------------------------------------------------------
extern "C" int dd_step_feuler(dd_batch* b, int slot_in, int slot_out, const double* t0, const double* dt, int n_t) {
    if (!b || !slot_ok(b, slot_in) || !slot_ok(b, slot_out) || slot_in == slot_out) return DD_ERR_INVALID;
    b->prev_valid = false;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    int rc = set_times(b, t0, dt, n_t);
    if (rc != DD_OK) return rc;
    if ((rc = stage_sources(b, t0[0], dt[0], n_t == 1, false)) != DD_OK) return rc;
    CKP(PC_FEULER, 1, dd_launch_feuler(launch_of(b), b->smode, b->g, b->d_mem, b->sF, cstate(b, slot_in), mstate(b, slot_out)));
    CK(cudaStreamSynchronize(ctx->stream));
    return DD_OK;
}

extern "C" int dd_eval_fields(dd_batch* b, int slot_in, int slot_out, const double* t, int n_t) {
    if (!b || !slot_ok(b, slot_in) || !slot_ok(b, slot_out) || slot_in == slot_out || !t) return DD_ERR_INVALID;
    b->prev_valid = false;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    std::vector<double> one(n_t, 1.0);
    int rc = set_times(b, t, one.data(), n_t);
    if (rc != DD_OK) return rc;
    if (b->mode == DD_FORCING_PROGRAM) {
        if ((rc = stage_sources(b, t[0], 1.0, n_t == 1, false)) != DD_OK) return rc;
        CKP(PC_OTHER, 1, dd_launch_fields(launch_of(b), b->smode, b->g, b->d_mem, b->sF, cstate(b, slot_in), mstate(b, slot_out), 0));
    } else {
        CKP(PC_OTHER, 1, dd_launch_fields(launch_of(b), b->mode, b->g, b->d_mem, b->F, cstate(b, slot_in), mstate(b, slot_out), 0));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return DD_OK;
}

// error norms of `slot` into out_host (through b->d_norm_out) or, when out_dev is given, into that device buffer
static int norms_async(dd_batch* b, int slot, int slot_exact, double* out_host, double* out_dev = nullptr) {
    dd_ctx* ctx = b->ctx;
    DDStateC ex;
    if (slot_exact >= 0) ex = cstate(b, slot_exact);
    bool have_exact = slot_exact >= 0;
    if (!have_exact && b->mode == DD_FORCING_PROGRAM) {
        // the program writes the exact solution at the members' current time into scratch fields
        static const char* names[DD_NVAR] = {"exact_cp", "exact_T", "exact_cl", "exact_cd", "exact_cs"};
        DDState scratch;
        int rc;
        for (int v = 0; v < DD_NVAR; ++v) {
            if ((rc = get_work(b, names[v], &scratch.v[v])) != DD_OK) return rc;
            ex.v[v] = scratch.v[v];
        }
        if ((rc = launch_program(b, DD_PROGRAM_EXACT, 0, scratch)) != DD_OK) return rc;
        have_exact = true;
    }
    CKP(PC_NORMS, 2, dd_launch_error_norms(launch_of(b, ROWS_OWNED), b->mode, b->g, b->d_mem, b->F, cstate(b, slot),
                                           have_exact ? &ex : nullptr, b->d_norm_partial, b->norm_bpm,
                                           out_dev ? out_dev : b->d_norm_out));
    if (!out_dev)
        CK(cudaMemcpyAsync(out_host, b->d_norm_out, sizeof(double) * 8 * b->B, cudaMemcpyDeviceToHost, ctx->stream));
    return DD_OK;
}

extern "C" int dd_error_norms(dd_batch* b, int slot, int slot_exact, const double* t, int n_t, double* out) {
    if (!b || !slot_ok(b, slot) || !out || (slot_exact >= 0 && !slot_ok(b, slot_exact))) return DD_ERR_INVALID;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    if (slot_exact < 0) {
        if (!t) return DD_ERR_INVALID;
        std::vector<double> one(n_t, 1.0);
        int rc = set_times(b, t, one.data(), n_t);
        if (rc != DD_OK) return rc;
    }
    int rc = norms_async(b, slot, slot_exact, out);
    if (rc != DD_OK) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return DD_OK;
}

// ---------------------------------------------------------------------------
// predictor-corrector step
// ---------------------------------------------------------------------------
extern "C" void dd_pc_options_default(dd_pc_options* o) {
    if (!o) return;
    o->num_pc_steps = 1;
    o->num_newton_steps = 1;
    o->num_newton_iterations = 5;
    o->cd_band_swap = 1;
    o->consec_xs_rtol = 1e-6;
    o->solve_tol = 1e-14;
    o->max_sweeps = 20000;
    o->fixed_sweeps = 0;
    o->extrapolate_guess = 0;
}

static const size_t kSmemMax = 227 * 1024;
static const int kSmemArrays = 6;

// SOR sweeps needed for spectral-radius bound rho of the Jacobi matrix
static int sweeps_for_rho(double rho, int max_sweeps) {
    if (!(rho >= 0.0)) return max_sweeps;
    if (rho < 1e-300) return 2;
    double lam;
    if (rho < 1.0) {
        const double om = 2.0 / (1.0 + sqrt(1.0 - rho * rho));
        lam = om - 1.0;
    } else {
        lam = 0.999;  // not diagonally dominant: plain Gauss-Seidel, rely on the residual check
    }
    int k = 2;
    while (k < max_sweeps && (1.0 + k) * pow(lam, k) > 1e-17) ++k;
    return k;
}

// Sweep plan for the next step from this step's verified residual.  ratio = resid / allowed (<= 1 passed).
// One SOR sweep multiplies the error by about lam = omega - 1; drop a sweep while the predicted ratio keeps
// a 10x margin, add one when the margin is gone.  The bound itself is always re-verified after the solve.
static int next_plan(int cur, double rho, double ratio, int max_sweeps) {
    double lam = 0.999;
    if (rho >= 0.0 && rho < 1.0) lam = 2.0 / (1.0 + sqrt(1.0 - rho * rho)) - 1.0;
    if (lam < 1e-12) lam = 1e-12;
    int want = cur;
    if (!(ratio <= 0.1)) {
        want = cur + 1;
    } else if (cur > 1 && ratio * 10.0 < lam) {
        // a residual at rounding level (ratio ~ 0) says nothing about how many sweeps were in excess
        const double r = ratio > 1e-6 ? ratio : 1e-6;
        // never assume more than a 20x error reduction per sweep when shortening the plan
        const double lam_eff = lam > 0.05 ? lam : 0.05;
        int drop = (int)floor(log(r * 10.0) / log(lam_eff));
        if (drop < 1) drop = 1;
        if (drop > cur / 2) drop = cur / 2;  // a failed step costs a whole retry: shorten the plan gradually
        want = cur - drop;
    }
    if (want > max_sweeps) want = max_sweeps;
    if (want < 1) want = 1;
    return want;
}

// The over-relaxation factor is the optimal one of a symmetric, consistently ordered matrix.  The cd system of a
// very large time step is far from symmetric (its T-Jacobian term), and SOR can then DIVERGE although the matrix
// is strictly diagonally dominant (seen at dt D / h^2 ~ 9).  Gauss-Seidel (omega = 1) converges for every strictly
// diagonally dominant matrix, with an error factor <= rho per sweep: once three times the theoretical SOR count
// has failed the residual bound, the variable's solves use it for the rest of the batch's life.
static int sweeps_for_gs(double rho, int max_sweeps) {
    if (!(rho >= 0.0) || rho >= 1.0) return max_sweeps;
    if (rho < 1e-300) return 2;
    int k = 2;
    while (k < max_sweeps && pow(rho, k) > 1e-17) ++k;
    return k;
}
static bool gs_fallback(dd_batch* b, int vi, int used, double rho, int max_sweeps, int* next) {
    if (b->gs_only[vi] || !(rho < 1.0)) return false;
    if (used < 3 * sweeps_for_rho(rho * 1.02 + 1e-12, max_sweeps) + 8) return false;
    b->gs_only[vi] = true;
    *next = sweeps_for_gs(rho * 1.02 + 1e-12, max_sweeps);
    return true;
}

// choose tile shape and sweeps per pass
static void plan_pass(const dd_batch* b, int sweeps_left, bool allow_last, bool const_band, DDSolvePlan* P) {
    const int rows = b->own1 - b->own0, cols = b->M + 1;
    // const-band (T) systems stage x, bb, dinv (+ four short 1-D arrays); general systems x, bb and 4 bands
    const size_t cell_bytes = (const_band ? 3 : kSmemArrays) * sizeof(double);
    const size_t extra_bytes = const_band ? 4096 : 0;
    const size_t cells_max = (kSmemMax - extra_bytes) / cell_bytes;
    P->const_band = const_band ? 1 : 0;
    P->rpw = 0;
    // ---- register-resident kernel (default): staged region 16*rpw rows x 64 columns, 512 threads ----
    {
        const char* e_solver = getenv("DD_SOLVER");
        const bool want_reg = !(e_solver && !strcmp(e_solver, "smem"));
        static const int rpw_gen[] = {2, 3, 4}, rpw_cb[] = {2, 3, 4, 6, 8};
        const int* rp = const_band ? rpw_cb : rpw_gen;
        const int nrp = const_band ? 5 : 3;
        const char* e_rpw = getenv("DD_RPW");
        if (e_rpw && !*e_rpw) e_rpw = nullptr;
        const char* e_spp = getenv("DD_SWEEPS_PER_PASS");
        if (e_spp && !*e_spp) e_spp = nullptr;  // empty = unset
        if (want_reg) {
            // whole member inside one staged region: no halo at all
            // (halo 1 instead of 0 keeps the staged origin on an even column for the 16-byte loads)
            if (!b->is_slab && cols <= 60) {
                for (int q = 0; q < nrp; ++q)
                    if (rows <= 16 * rp[q] - 4) {
                        P->rpw = rp[q];
                        P->sweeps = sweeps_left;
                        P->tile_i = rows;
                        P->tile_j = cols + (cols & 1);
                        P->halo = 1;
                        P->last_pass = 1;
                        P->threads = 512;
                        P->smem_bytes = 0;
                        return;
                    }
            }
            // tiled: cost per pass ~ staged cells x (load + sweeps + epilogue) in sweep units; loads dominate
            const double c_load = const_band ? 3.0 : 6.0, c_epi = 1.5;
            double best = 1e300;
            DDSolvePlan bp = *P;
            bp.sweeps = 0;
            for (int S = sweeps_left; S >= 1; --S) {
                if (e_spp && atoi(e_spp) < sweeps_left && S != atoi(e_spp)) continue;
                const int last = (S == sweeps_left) && allow_last;
                const int H = 2 * S + 1;  // odd on every pass: staged origin on an even column
                const int tj = 62 - 2 * H;
                if (tj < 8) continue;
                for (int q = 0; q < nrp; ++q) {
                    if (e_rpw && rp[q] != atoi(e_rpw)) continue;
                    const int SI = 16 * rp[q], ti = SI - 2 * H - 2;
                    if (ti < 4) continue;
                    const double tiles = (double)((rows + ti - 1) / ti) * (double)((cols + tj - 1) / tj) * b->B;
                    const int npass = (sweeps_left + S - 1) / S;
                    const double cost = tiles * (double)SI * 64.0 * (c_load + S + c_epi) * npass;
                    if (cost < best) {
                        best = cost;
                        bp.rpw = rp[q];
                        bp.sweeps = S;
                        bp.tile_i = ti;
                        bp.tile_j = tj;
                        bp.halo = H;
                        bp.last_pass = last;
                        bp.threads = 512;
                        bp.smem_bytes = 0;
                    }
                }
            }
            if (bp.sweeps > 0) {
                *P = bp;
                return;
            }
        }
    }
    // whole member in one tile (no halo needed because every edge is a physical boundary)
    if (!b->is_slab && (size_t)(rows + 2) * ((cols + 3) & ~1) <= cells_max) {
        P->sweeps = sweeps_left;
        P->tile_i = rows;
        P->tile_j = cols;
        P->halo = 0;
        P->last_pass = 1;
        const size_t cells = (size_t)(rows + 2) * ((cols + 3) & ~1);
        P->smem_bytes = cells * cell_bytes + (const_band ? 2 * sizeof(double) * (rows + 2 + ((cols + 3) & ~1)) : 0);
        P->threads = cells >= 2048 ? 512 : (cells >= 512 ? 256 : 128);
        return;
    }
    const int sm = b->ctx->sm_count > 0 ? b->ctx->sm_count : 148;
    static const int cand_i[] = {8, 12, 16, 24, 32, 40, 48, 56, 64, 80, 96};
    static const int cand_sj[] = {64, 96, 128, 160, 192, 256};  // staged width: packed width multiple of 16
    // development overrides
    const char* e_ti = getenv("DD_TILE_I");
    const char* e_sj = getenv("DD_STAGE_J");
    const char* e_sp = getenv("DD_SWEEPS_PER_PASS");
    if (e_sp && !*e_sp) e_sp = nullptr;
    const char* e_th = getenv("DD_THREADS");
    double best = 1e300;
    DDSolvePlan bp = *P;
    bp.sweeps = 0;
    // sweeps per pass: try everything that fits; cost ~ (waves of CTAs) x (staged cells) x (staging + sweeps)
    for (int S = sweeps_left; S >= 1; --S) {
        if (e_sp && S != atoi(e_sp) && S != sweeps_left && atoi(e_sp) < sweeps_left) continue;
        const int last = (S == sweeps_left) && allow_last;
        const int H = 2 * S + (last ? 1 : 0);
        for (int ti : cand_i)
            for (int sj : cand_sj) {
                if (e_ti && ti != atoi(e_ti)) continue;
                if (e_sj && sj != atoi(e_sj)) continue;
                const int tj = sj - 2 * H - 2;
                if (tj < 8) continue;
                const size_t cells = (size_t)(ti + 2 * H + 2) * sj;
                if (cells > cells_max) continue;
                const long long tiles = (long long)((rows + ti - 1) / ti) * ((cols + tj - 1) / tj) * b->B;
                int per_sm = (int)(kSmemMax / (cells * cell_bytes + extra_bytes));
                if (per_sm > 4) per_sm = 4;
                // measured on B200 (profiles/r01_tile_sweep.log): a pass costs about (4 + S) "sweep units" per
                // staged cell (staging + epilogue ~ 4 sweeps), CTAs fill the machine in waves, and a single
                // resident CTA per SM cannot overlap its staging with another CTA's sweeps
                const double waves = ceil((double)tiles / ((double)sm * per_sm));
                const double per_cta = (double)cells * (4.0 + S) * (per_sm > 1 ? 1.0 : 1.3);
                const int npass = (sweeps_left + S - 1) / S;
                const double cost = waves * per_sm * per_cta * npass;
                if (cost < best) {
                    best = cost;
                    bp.sweeps = S;
                    bp.tile_i = ti;
                    bp.tile_j = tj;
                    bp.halo = H;
                    bp.last_pass = last;
                    bp.smem_bytes = cells * cell_bytes + (const_band ? 2 * sizeof(double) * ((ti + 2 * H + 2) + sj) : 0);
                    bp.threads = e_th ? atoi(e_th) : (cells >= 2048 ? 512 : 256);
                }
            }
    }
    *P = bp;
}

__global__ void k_summarise(const DDSolveStats* st, int nmem, const DDMember* mem, double tol, SolveSummary* out) {
    // one block per solve (blockIdx.x): statistics st[solve][member] -> out[solve]
    __shared__ double s_rho[256], s_ratio[256], s_res[256], s_bound[256];
    st += (size_t)blockIdx.x * nmem;
    out += blockIdx.x;
    double rho = 0.0, ratio = -1.0, res = 0.0, bound = 0.0;
    for (int m = threadIdx.x; m < nmem; m += blockDim.x) {
        if (!mem[m].active) continue;
        const DDSolveStats s = st[m];
        const double gap = s.rho < 1.0 ? 1.0 - s.rho : 1e-4;
        // |x - x*|_inf <= resid / (1 - rho); allowed: tol * |v_new| plus the rounding floor of the residual
        const double eps = 2.220446049250313e-16;
        const double allowed = tol * gap * s.vmax + 16.0 * eps * (s.bmax + s.xmax);
        double r = (allowed > 0.0) ? s.resid / allowed : (s.resid > 0.0 ? 1e300 : 0.0);
        if (s.resid != s.resid) r = 1e300;
        // A system that is NaN / Inf on entry (blown-up state: the reference's direct solve just returns NaN and
        // its studies carry on, src/mms_trial_utils.py:48 "max(0, nan)") is not a convergence failure: it is
        // accepted and left out of the summary; a solve with nothing else in it reports ratio -1, from which the
        // sweep controller learns nothing.
        if (!(s.rho < 1e300) || !(s.bmax < 1e300)) continue;
        if (rho == rho) rho = (s.rho != s.rho) ? s.rho : fmax(rho, s.rho);  // a NaN ratio sticks
        ratio = fmax(ratio, r);
        res = fmax(res, s.resid);
        bound = fmax(bound, s.vmax > 0.0 ? s.resid / (gap * s.vmax) : 0.0);
    }
    s_rho[threadIdx.x] = rho; s_ratio[threadIdx.x] = ratio; s_res[threadIdx.x] = res; s_bound[threadIdx.x] = bound;
    __syncthreads();
    for (int w = blockDim.x >> 1; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            const int k = threadIdx.x + w;
            const double a = s_rho[threadIdx.x], c = s_rho[k];
            s_rho[threadIdx.x] = (a != a) ? a : ((c != c) ? c : fmax(a, c));
            s_ratio[threadIdx.x] = fmax(s_ratio[threadIdx.x], s_ratio[k]);
            s_res[threadIdx.x] = fmax(s_res[threadIdx.x], s_res[k]);
            s_bound[threadIdx.x] = fmax(s_bound[threadIdx.x], s_bound[k]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out->rho = s_rho[0]; out->ratio = s_ratio[0]; out->resid = s_res[0]; out->bound = s_bound[0];
    }
}

__global__ void k_flags_to_summary(const int* flags, const DDMember* mem, int nmem, SolveSummary* out) {
    int any = 0;
    for (int m = threadIdx.x; m < nmem; m += 32)
        if (mem[m].active && (flags[m] & 1)) any = 1;
    any = __any_sync(0xffffffffu, any);
    if (threadIdx.x == 0) {
        out->rho = any ? 1.0 : 0.0;
        out->ratio = out->resid = out->bound = 0.0;
    }
}

static int ensure_solve_slots(dd_batch* b, int n) {
    dd_ctx* ctx = b->ctx;
    if (n <= b->nsolve_cap) return DD_OK;
    cudaFree(b->d_stats);
    cudaFree(b->d_summary);
    CK(cudaMalloc((void**)&b->d_stats, sizeof(DDSolveStats) * (size_t)n * b->B));
    CK(cudaMalloc((void**)&b->d_summary, sizeof(SolveSummary) * n));
    b->nsolve_cap = n;
    return DD_OK;
}

__global__ void k_cs_arm(double* it_max, double* it_min, long long n) {
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= n) return;
    it_max[k] = 0.0;
    it_min[k] = __longlong_as_double(0x7ff0000000000000LL);
}

static int ensure_cs_buffers(dd_batch* b, int cap) {
    dd_ctx* ctx = b->ctx;
    if (!b->d_used) CK(cudaMalloc((void**)&b->d_used, sizeof(int) * b->B));
    if (cap <= b->cs_cap_alloc) return DD_OK;
    cudaFree(b->d_itmax);
    cudaFree(b->d_itmin);
    const long long n = (long long)cap * b->B;
    CK(cudaMalloc((void**)&b->d_itmax, sizeof(double) * n));
    CK(cudaMalloc((void**)&b->d_itmin, sizeof(double) * n));
    b->cs_cap_alloc = cap;
    return DD_OK;
}

// assemble + solve one Newton system.  `k` indexes the stats slot.
static int get_rows(dd_batch* b, DDRows* R) {
    int rc;
    if ((rc = get_work(b, "bb", &R->bb)) != DD_OK) return rc;
    if ((rc = get_work(b, "aW", &R->aW)) != DD_OK) return rc;
    if ((rc = get_work(b, "aE", &R->aE)) != DD_OK) return rc;
    if ((rc = get_work(b, "aS", &R->aS)) != DD_OK) return rc;
    if ((rc = get_work(b, "aN", &R->aN)) != DD_OK) return rc;
    R->ld = b->g.ld + (b->g.ld & 1);
    R->mstride = (long long)b->nrows * R->ld;
    return DD_OK;
}

// Predictor of the step; on wide grids with array (or no) sources the marching kernel, which also assembles
// the T system of the first Newton step (*fused_T = true: the caller then only solves it, solve slot 0).
static int launch_predictor(dd_batch* b, const DDStateC& s0, const DDPredictOut& po, bool* fused_T, bool need_YT) {
    dd_ctx* ctx = b->ctx;
    const DDLaunch L = launch_of(b, ROWS_STENCIL);
    *fused_T = false;
    if (dd_predict_march_ok(b->g, L, b->smode)) {
        DDRows R;
        int rc = get_rows(b, &R);
        if (rc != DD_OK) return rc;
        g_prof.launches += 1;  // the statistics reset
        CKP(PC_PREDICT, 1, dd_launch_predict_march(L, b->smode, b->g, b->d_mem, b->sF, s0, po, true, R, b->d_stats,
                                                       need_YT));
        *fused_T = true;
        return DD_OK;
    }
    CKP(PC_PREDICT, 1, dd_launch_predict(L, b->smode, b->g, b->d_mem, b->sF, s0, po));
    return DD_OK;
}

static int newton_solve(dd_batch* b, int var, const DDStateC& ustar, const double* T1, const double* cl1,
                        const double* Y, double* vnew, const dd_pc_options& opt, int k, int* sweeps_used,
                        int* passes_used, int what = 3, const double* vold = nullptr, bool summarise = true,
                        int seg_sweeps = 0, const double* seg_xin = nullptr, bool seg_finish = true,
                        const double** seg_xlast = nullptr) {
    // seg_*: one SEGMENT of a solve that the slab driver splits because the halo supports only so many sweeps
    // between two exchanges of the iterate: seg_sweeps sweeps continuing from the iterate seg_xin (null: from zero);
    // seg_finish = false leaves the iterate in a work array (*seg_xlast) instead of writing v_new and the statistics
    dd_ctx* ctx = b->ctx;
    DDRows R;
    int rc;
    if ((rc = get_rows(b, &R)) != DD_OK) return rc;
    DDSolveStats* st = b->d_stats + (size_t)k * b->B;
    // rows are assembled on every local row that has a full stencil (slabs: halo rows included,
    // so that the tile solver sees valid rows in its halo); tiles cover the owned rows only
    const DDLaunch L = launch_of(b, ROWS_OWNED);
    const DDLaunch La = launch_of(b, ROWS_ASM);
    if (what & 1)
        CKP(PC_ASM_T + (var - DD_T), 2,
            dd_launch_assemble(La, b->smode, var, b->g, b->d_mem, b->sF, ustar, T1, cl1, Y, opt.cd_band_swap, R, st));
    if (!(what & 2)) return DD_OK;
    const int vi = var - DD_T;
    int sweeps = seg_sweeps > 0 ? seg_sweeps : (opt.fixed_sweeps > 0 ? opt.fixed_sweeps : b->ctl[b->cm].sweeps[vi]);
    if (sweeps <= 0) {
        // first use: read the Gershgorin ratio back once to seed the plan
        CK(cudaMemsetAsync(b->d_summary + k, 0, sizeof(SolveSummary), ctx->stream));
        k_summarise<<<1, 256, 0, ctx->stream>>>(st, b->B, b->d_mem, opt.solve_tol, b->d_summary + k);
        SolveSummary s;
        CK(cudaMemcpyAsync(&s, b->d_summary + k, sizeof(s), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        sweeps = sweeps_for_rho(s.rho, opt.max_sweeps);
        b->ctl[b->cm].sweeps[vi] = sweeps;
    }
    const double* vstar = ustar.v[var];
    int left = sweeps, passes = 0;
    double *xa = nullptr, *xb = nullptr;
    const double* xin = seg_xin;  // nullptr = zero initial iterate, or vstar - vold on the first pass
    // rows on which the rows R and the iterate x are valid.  On a slab every pass consumes 2 rows per sweep
    // from each interior side (the halo rows are recomputed redundantly, never exchanged mid-solve); sides
    // on the physical boundary do not shrink.
    int v0 = b->asm0, v1 = b->asm1;
    const bool lo_edge = (b->row0 == 0), hi_edge = (b->row0 + b->nrows == b->N + 1);
    // wide grids, the variables DD_WAVE names (default: cl): wavefront kernel (rows marched once per pass, no row
    // halo); the previous step's increment as initial iterate is only known to the tile kernels
    const bool lanek = dd_lane_ok(b->g, L, var) && !vold;
    const bool wave = !lanek && dd_wave_ok(b->g, L, var) && !vold;
    while (left > 0) {
        DDSolvePlan P;
        memset(&P, 0, sizeof(P));
        if (lanek) {
            // lane-private marching kernel: passes of nearly equal length, at most dd_lane_max_sweeps() sweeps each
            P.sweeps = dd_lane_pass_sweeps(left);
            P.last_pass = (seg_finish && P.sweeps == left) ? 1 : 0;
            P.const_band = var == DD_T ? 1 : 0;
            P.halo = 2 * P.sweeps + 1;
        } else if (wave) {
            // as many sweeps per pass as the ring of the widest fitting kernel variant holds
            int cap_all = 0, cap_wide = 0;
            dd_wave_max_sweeps(var == DD_T, &cap_all, &cap_wide);
            P.sweeps = left <= cap_all ? left : cap_wide;
            P.last_pass = (seg_finish && P.sweeps == left) ? 1 : 0;
            P.const_band = var == DD_T ? 1 : 0;
            P.halo = 2 * P.sweeps + 1;
        } else {
            plan_pass(b, left, seg_finish, var == DD_T, &P);
        }
        P.rho_fix = b->gs_only[vi] ? 0.0 : b->relax_rho[vi];  // (ratio 0: omega = 1)
        if (P.sweeps <= 0) return fail(ctx, DD_ERR_INVALID, "no feasible solver tile");
        double* xout = nullptr;
        DDLaunch Lp = L;
        Lp.vr0 = v0;
        Lp.vr1 = v1;
        if (!P.last_pass) {
            if (!xa) {
                if ((rc = get_work(b, "xa", &xa)) != DD_OK) return rc;
                if ((rc = get_work(b, "xb", &xb)) != DD_OK) return rc;
            }
            xout = (xin == xa) ? xb : xa;
            Lp.own0 = lo_edge ? v0 : v0 + 2 * P.sweeps;
            Lp.own1 = hi_edge ? v1 : v1 - 2 * P.sweeps;
            if (Lp.own0 > b->own0 || Lp.own1 < b->own1)
                return fail(ctx, DD_ERR_INVALID, "halo too shallow for the planned SOR sweeps");
        } else {
            const int need = 2 * P.sweeps + 1;
            if ((!lo_edge && v0 + need > b->own0) || (!hi_edge && v1 - need < b->own1))
                return fail(ctx, DD_ERR_INVALID, "halo too shallow for the planned SOR sweeps");
        }
        if (P.last_pass && P.sweeps < left) return fail(ctx, DD_ERR_INVALID, "solver plan inconsistency");
        const double* vo = passes == 0 ? vold : nullptr;
        if (vo && P.rpw <= 0) {
            // shared-memory kernel: it stages x from an array
            double* x0 = nullptr;
            if ((rc = get_work(b, "x0", &x0)) != DD_OK) return rc;
            CKP(PC_SOLVE_T + (var - DD_T), 1, dd_launch_make_guess(Lp, b->g, R, vstar, vo, x0));
            xin = x0;
            vo = nullptr;
        }
        if (lanek)
            CKP(PC_SOLVE_T + (var - DD_T), 1,
                dd_launch_solve_lane(Lp, b->g, b->d_mem, R, xin, xout, vstar, vnew, var == DD_T ? 1 : 0, st,
                                     P.const_band, P.sweeps, P.last_pass, P.rho_fix));
        else if (wave)
            CKP(PC_SOLVE_T + (var - DD_T), 1,
                dd_launch_solve_wave(Lp, b->g, b->d_mem, R, xin, xout, vstar, vnew, var == DD_T ? 1 : 0, st,
                                     P.const_band, P.sweeps, P.last_pass, P.rho_fix));
        else
            CKP(PC_SOLVE_T + (var - DD_T), 1,
                dd_launch_solve_pass(Lp, b->g, b->d_mem, R, xin, vo, xout, vstar, vnew, var == DD_T ? 1 : 0, st, P));
        dd_note_solver_kernel(var, dd_last_solver_kernel());
        left -= P.sweeps;
        xin = xout;
        if (!P.last_pass) {
            v0 = Lp.own0;
            v1 = Lp.own1;
        }
        ++passes;
    }
    if (seg_xlast) *seg_xlast = xin;
    if (summarise && seg_finish) {
        ProfScope ps_(ctx->stream, PC_SUMMARISE, 1);
        k_summarise<<<1, 256, 0, ctx->stream>>>(st, b->B, b->d_mem, opt.solve_tol, b->d_summary + k);
    }
    CK(cudaGetLastError());
    *sweeps_used = sweeps;
    *passes_used = passes;
    return DD_OK;
}

struct StepIO {
    int slot_in, slot_out;
};

// Decides whether this step may start its solves from the previous step's increment, and disarms the
// record so that a retry of the same step (whose output slot then holds a rejected result) starts from zero.
static bool take_guess(dd_batch* b, int slot_in, int slot_out, const dd_pc_options& opt) {
    static const bool off = getenv("DD_NO_GUESS") != nullptr;
    bool ok = !off && opt.extrapolate_guess && b->prev_valid && b->prev_in == slot_out &&
                    b->prev_out == slot_in && slot_in != slot_out && b->prev_dt == b->cur_dt;
    b->prev_valid = false;
    b->cm = ok ? 1 : 0;
    return ok;
}
static void record_step(dd_batch* b, int slot_in, int slot_out) {
    b->prev_in = slot_in;
    b->prev_out = slot_out;
    b->prev_dt = b->cur_dt;
    b->prev_valid = true;
}

static int rec_prepare(dd_batch* b, dd_batch::StepRec& R, int nsolves) {
    dd_ctx* ctx = b->ctx;
    if (R.cap_sums < nsolves) {
        if (R.h_sums) cudaFreeHost(R.h_sums);
        R.h_sums = nullptr;
        CK(cudaMallocHost((void**)&R.h_sums, sizeof(SolveSummary) * nsolves));
        R.cap_sums = nsolves;
    }
    if (!R.h_used) CK(cudaMallocHost((void**)&R.h_used, sizeof(int) * b->B));
    if (!R.h_flags) CK(cudaMallocHost((void**)&R.h_flags, sizeof(int) * b->B));
    if (!R.done) CK(cudaEventCreateWithFlags(&R.done, cudaEventDisableTiming));
    return DD_OK;
}

// Enqueues one PC step and the read-back of its summaries into the record R (nothing is waited for).
static int pc_step_enqueue(dd_batch* b, int slot_in, int slot_out, const dd_pc_options& opt, dd_batch::StepRec& R) {
    dd_ctx* ctx = b->ctx;
    int rc;
    const int P = opt.num_pc_steps, Q = opt.num_newton_steps;
    if ((rc = ensure_solve_slots(b, 3 * P * Q)) != DD_OK) return rc;
    if ((rc = rec_prepare(b, R, 3 * P * Q)) != DD_OK) return rc;
    R.solve_sweeps.clear();
    CK(cudaMemsetAsync(b->d_flags, 0, sizeof(int) * b->B, ctx->stream));
    DDPredictOut po;
    if ((rc = get_work(b, "cp1p", &po.cp1p)) != DD_OK) return rc;
    if ((rc = get_work(b, "cs1p", &po.cs1p)) != DD_OK) return rc;
    if ((rc = get_work(b, "YT", &po.YT)) != DD_OK) return rc;
    if ((rc = get_work(b, "Ycl", &po.Ycl)) != DD_OK) return rc;
    if ((rc = get_work(b, "Ycd", &po.Ycd)) != DD_OK) return rc;
    const DDLaunch Lall = launch_of(b, ROWS_ALL);
    const DDStateC s0 = cstate(b, slot_in);
    const DDState sout = mstate(b, slot_out);
    bool fused_T = false;
    if ((rc = launch_predictor(b, s0, po, &fused_T, P * Q > 1)) != DD_OK) return rc;
    DDStateC u;
    u.v[DD_CP] = po.cp1p; u.v[DD_T] = s0.v[DD_T]; u.v[DD_CL] = s0.v[DD_CL]; u.v[DD_CD] = s0.v[DD_CD];
    u.v[DD_CS] = po.cs1p;
    const bool guess = take_guess(b, slot_in, slot_out, opt);
    static const char* tmpn[3][2] = {{"T1a", "T1b"}, {"cl1a", "cl1b"}, {"cd1a", "cd1b"}};
    int pp = 0, k = 0;
    int sweeps[3] = {0, 0, 0}, passes[3] = {0, 0, 0};
    for (int pc = 0; pc < P; ++pc) {
        for (int nw = 0; nw < Q; ++nw) {
            const bool last = (pc == P - 1) && (nw == Q - 1);
            double* dst[3];
            for (int q = 0; q < 3; ++q) {
                if (last) {
                    dst[q] = sout.v[DD_T + q];
                } else if ((rc = get_work(b, tmpn[q][pp], &dst[q])) != DD_OK) {
                    return rc;
                }
            }
            // the first Newton solve of the step starts from the previous step's increment when it is available
            const bool gs = guess && pc == 0 && nw == 0;
            b->cm = gs ? 1 : 0;
            // the marching predictor has already assembled the T system of the very first Newton step
            const int whatT = (fused_T && pc == 0 && nw == 0) ? 2 : 3;
            if ((rc = newton_solve(b, DD_T, u, nullptr, nullptr, po.YT, dst[0], opt, k++, &sweeps[0], &passes[0],
                                   whatT, gs ? sout.v[DD_T] : nullptr, false)) != DD_OK) return rc;
            if ((rc = newton_solve(b, DD_CL, u, dst[0], nullptr, po.Ycl, dst[1], opt, k++, &sweeps[1], &passes[1], 3,
                                   gs ? sout.v[DD_CL] : nullptr, false)) != DD_OK) return rc;
            if ((rc = newton_solve(b, DD_CD, u, dst[0], dst[1], po.Ycd, dst[2], opt, k++, &sweeps[2], &passes[2], 3,
                                   gs ? sout.v[DD_CD] : nullptr, false)) != DD_OK) return rc;
            for (int q = 0; q < 3; ++q) R.solve_sweeps.push_back(sweeps[q]);
            u.v[DD_T] = dst[0]; u.v[DD_CL] = dst[1]; u.v[DD_CD] = dst[2];
            pp ^= 1;
        }
        const bool lastpc = (pc == P - 1);
        double* cpd = lastpc ? sout.v[DD_CP] : po.cp1p;
        double* csd = lastpc ? sout.v[DD_CS] : po.cs1p;
        const int cap = opt.num_newton_iterations;
        const bool track = opt.consec_xs_rtol > 0.0 && cap > 0;
        if (track) {
            const int had = b->cs_cap_alloc;
            if ((rc = ensure_cs_buffers(b, cap)) != DD_OK) return rc;
            if (b->cs_cap_alloc != had) {
                const long long n = (long long)b->cs_cap_alloc * b->B;
                k_cs_arm<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(b->d_itmax, b->d_itmin, n);
            }
        }
        CKP(PC_CORRECT, 1,
            dd_launch_correct(Lall, b->smode, b->g, b->d_mem, b->sF, s0, u.v[DD_T], u.v[DD_CL], u.v[DD_CD], cpd, csd,
                              cap, track ? opt.consec_xs_rtol : 0.0, b->d_itmax, b->d_itmin, b->has_variants ? b->d_flags : nullptr));
        if (track)
            CKP(PC_CS_FINISH, 2,
                dd_launch_cs_finish(Lall, b->smode, b->g, b->d_mem, b->sF, s0, u.v[DD_CL], u.v[DD_CD], csd, cap,
                                    opt.consec_xs_rtol, b->d_itmax, b->d_itmin, b->d_used));
        u.v[DD_CP] = cpd; u.v[DD_CS] = csd;
    }
    {
        // summaries of all solves of the step in one launch (one block per solve)
        ProfScope ps_(ctx->stream, PC_SUMMARISE, 1);
        k_summarise<<<k, 256, 0, ctx->stream>>>(b->d_stats, b->B, b->d_mem, opt.solve_tol, b->d_summary);
    }
    CK(cudaGetLastError());
    // verification of every solve of this step: one small read-back, waited for in pc_step_finish
    CK(cudaMemcpyAsync(R.h_sums, b->d_summary, sizeof(SolveSummary) * k, cudaMemcpyDeviceToHost, ctx->stream));
    const bool track_used = opt.consec_xs_rtol > 0.0 && opt.num_newton_iterations > 0;
    if (track_used)
        CK(cudaMemcpyAsync(R.h_used, b->d_used, sizeof(int) * b->B, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(R.h_flags, b->d_flags, sizeof(int) * b->B, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaEventRecord(R.done, ctx->stream));
    R.active = true;
    R.slot_in = slot_in;
    R.slot_out = slot_out;
    R.nsolves = k;
    R.guess = guess;
    R.opt = opt;
    for (int q = 0; q < 3; ++q) {
        R.sweeps[q] = sweeps[q];
        R.passes[q] = passes[q];
    }
    return DD_OK;
}

// Waits for the summaries of the step in R, updates the sweep controller and fills `stats`.
static int pc_step_finish(dd_batch* b, dd_batch::StepRec& R, dd_step_stats* stats, bool* converged) {
    dd_ctx* ctx = b->ctx;
    CK(cudaEventSynchronize(R.done));
    R.active = false;
    for (int m = 0; m < b->B; ++m)
        if (b->h_mem[m].active && (R.h_flags[m] & 1))
            return fail(ctx, DD_ERR_DOMAIN,
                        "Denominator 2 - dt Kd (Sd - Cd1) (1 + Cl1) below positiveness treshold.");
    const dd_pc_options& opt = R.opt;
    const int k = R.nsolves;
    const bool guess = R.guess;
    const SolveSummary* sums = R.h_sums;
    const int* sweeps = R.sweeps;
    const int* passes = R.passes;
    *converged = true;
    for (int q = 0; q < k; ++q)
        if (!(sums[q].ratio <= 1.0)) *converged = false;
    if (opt.fixed_sweeps <= 0) {
        for (int q = 0; q < k; ++q) {
            const int vi = q % 3;
            b->cm = (guess && q < 3) ? 1 : 0;
            dd_batch::Ctl& c = b->ctl[b->cm];
            const int used = R.solve_sweeps[q];  // with deferred verification the plan may have moved on since
            if (!(sums[q].ratio <= 1.0)) {
                // not enough sweeps: remember the failing count and go (at least) to the theoretical one
                int gs_next = 0;
                if (gs_fallback(b, vi, used, sums[q].rho, opt.max_sweeps, &gs_next)) {
                    for (int m = 0; m < 2; ++m) {
                        b->ctl[m].sweeps[vi] = gs_next;
                        b->ctl[m].extra[vi] = 0;
                        b->ctl[m].floor[vi] = 1;
                    }
                    continue;
                }
                if (c.floor[vi] < used + 1) c.floor[vi] = used + 1;
                const int want = b->gs_only[vi] ? sweeps_for_gs(sums[q].rho * 1.02 + 1e-12, opt.max_sweeps)
                                                : sweeps_for_rho(sums[q].rho * 1.02 + 1e-12, opt.max_sweeps);
                if (used >= want) c.extra[vi] += (used + 1) / 2 + 1;
                int next = want + c.extra[vi];
                if (next < c.floor[vi]) next = c.floor[vi];
                if (next > opt.max_sweeps) next = opt.max_sweeps;
                if (next > c.sweeps[vi]) c.sweeps[vi] = next;
            } else if (*converged && (q >= k - 3 || q < 3) && sums[q].ratio >= 0.0) {
                // (Gauss-Seidel fallback: the plan only grows; its error factor is not the one next_plan assumes)
                int next = b->gs_only[vi] ? used : next_plan(used, sums[q].rho, sums[q].ratio, opt.max_sweeps);
                if (next < c.floor[vi]) next = c.floor[vi];
                c.sweeps[vi] = next;
            }
        }
    }
    if (stats) {
        for (int q = 0; q < 3; ++q) {
            stats->sweeps[q] = sweeps[q];
            stats->passes[q] = passes[q];
            stats->rho[q] = sums[k - 3 + q].rho;
            stats->resid[q] = sums[k - 3 + q].resid;
            stats->bound[q] = sums[k - 3 + q].bound;
        }
        stats->cs_newton_iters = opt.num_newton_iterations;
        if (opt.consec_xs_rtol > 0.0 && opt.num_newton_iterations > 0) {
            int mx = 0;
            for (int m = 0; m < b->B; ++m)
                if (b->h_mem[m].active && R.h_used[m] > mx) mx = R.h_used[m];
            stats->cs_newton_iters = mx;
        }
    }
    return DD_OK;
}

// one PC step, verified before returning; times must already be on the device (set_times or advance)
static int pc_step_once(dd_batch* b, int slot_in, int slot_out, const dd_pc_options& opt, dd_step_stats* stats,
                        bool* converged) {
    dd_batch::StepRec& R = b->rec[b->rec_cur];
    if (R.active) return fail(b->ctx, DD_ERR_INVALID, "internal: a deferred step is still pending");
    int rc = pc_step_enqueue(b, slot_in, slot_out, opt, R);
    if (rc != DD_OK) return rc;
    return pc_step_finish(b, R, stats, converged);
}

static int check_opts(dd_ctx* ctx, const dd_pc_options& o) {
    if (o.num_pc_steps < 1 || o.num_newton_steps < 1 || o.num_newton_iterations < 0 || o.max_sweeps < 2)
        return fail(ctx, DD_ERR_INVALID, "bad dd_pc_options");
    return DD_OK;
}

static int pc_step_retry(dd_batch* b, int slot_in, int slot_out, const dd_pc_options& opt, dd_step_stats* stats) {
    dd_ctx* ctx = b->ctx;
    int retries = 0;
    for (;;) {
        bool ok = false;
        int rc = pc_step_once(b, slot_in, slot_out, opt, stats, &ok);
        if (rc != DD_OK) return rc;
        if (stats) stats->retries = retries;
        if (ok) {
            record_step(b, slot_in, slot_out);
            return DD_OK;
        }
        bool can_grow = opt.fixed_sweeps <= 0;
        if (can_grow) {
            can_grow = false;
            for (int q = 0; q < 3; ++q)
                if (b->ctl[0].sweeps[q] < opt.max_sweeps) can_grow = true;  // retries start from zero
        }
        if (!can_grow || retries >= 40) {
            char msg[256];
            snprintf(msg, sizeof(msg), "linear solve did not reach the residual bound (sweeps T/cl/cd = %d/%d/%d)",
                     b->ctl[0].sweeps[0], b->ctl[0].sweeps[1], b->ctl[0].sweeps[2]);
            return fail(ctx, DD_ERR_NOT_CONVERGED, msg);
        }
        ++retries;
    }
}

extern "C" int dd_step_pc(dd_batch* b, int slot_in, int slot_out, const double* t0, const double* dt, int n_t,
                          const dd_pc_options* opt_in, dd_step_stats* stats) {
    if (!b || !slot_ok(b, slot_in) || !slot_ok(b, slot_out) || slot_in == slot_out) return DD_ERR_INVALID;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    dd_pc_options opt;
    if (opt_in) opt = *opt_in; else dd_pc_options_default(&opt);
    int rc = check_opts(ctx, opt);
    if (rc != DD_OK) return rc;
    if ((rc = set_times(b, t0, dt, n_t)) != DD_OK) return rc;
    if ((rc = stage_sources(b, t0[0], dt[0], n_t == 1, true)) != DD_OK) return rc;
    return pc_step_retry(b, slot_in, slot_out, opt, stats);
}

// ---------------------------------------------------------------------------
// deferred verification: the summaries of step n are read while step n + 1 is already running, so the
// host never drains the stream between steps.  A rejected step is redone (with more sweeps) together
// with the step enqueued after it; that needs the input of the rejected step, i.e. three rotating slots.
// ---------------------------------------------------------------------------
static int redo_step(dd_batch* b, dd_batch::StepRec& R, dd_step_stats* stats) {
    // R is inactive here; its fields stay valid because pc_step_once works on rec[rec_cur], which is R itself
    // only after the copies below have been taken
    const std::vector<double> t0 = R.t0, dt = R.dt;
    const dd_pc_options opt = R.opt;
    const int in = R.slot_in, out = R.slot_out, n_t = R.n_t;
    int rc;
    if ((rc = set_times(b, t0.data(), dt.data(), n_t)) != DD_OK) return rc;
    if ((rc = stage_sources(b, t0[0], dt[0], n_t == 1, true)) != DD_OK) return rc;
    if ((rc = pc_step_retry(b, in, out, opt, stats)) != DD_OK) return rc;
    if (stats) stats->retries += 1;
    return DD_OK;
}

static int flush_pending(dd_batch* b, dd_step_stats* stats, int* have) {
    if (have) *have = 0;
    dd_batch::StepRec& R = b->rec[b->rec_cur];
    if (!R.active) return DD_OK;
    dd_ctx* ctx = b->ctx;
    bool ok = false;
    int rc = pc_step_finish(b, R, stats, &ok);
    if (rc != DD_OK) return rc;
    if (have) *have = 1;
    if (stats) stats->retries = 0;
    if (ok) {
        record_step(b, R.slot_in, R.slot_out);
        return DD_OK;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return redo_step(b, R, stats);
}

extern "C" int dd_step_pc_flush(dd_batch* b, dd_step_stats* stats, int* have_stats) {
    if (!b) return DD_ERR_INVALID;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    return flush_pending(b, stats, have_stats);
}

extern "C" int dd_step_pc_deferred(dd_batch* b, int slot_in, int slot_out, const double* t0, const double* dt,
                                   int n_t, const dd_pc_options* opt_in, dd_step_stats* prev_stats,
                                   int* have_prev) {
    if (!b || !slot_ok(b, slot_in) || !slot_ok(b, slot_out) || slot_in == slot_out) return DD_ERR_INVALID;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    if (have_prev) *have_prev = 0;
    dd_pc_options opt;
    if (opt_in) opt = *opt_in; else dd_pc_options_default(&opt);
    int rc = check_opts(ctx, opt);
    if (rc != DD_OK) return rc;
    dd_batch::StepRec& Old = b->rec[b->rec_cur];
    dd_batch::StepRec& New = b->rec[b->rec_cur ^ 1];
    if (Old.active && (slot_in != Old.slot_out))
        return fail(ctx, DD_ERR_INVALID, "deferred step must continue from the output slot of the pending step");
    if ((rc = set_times(b, t0, dt, n_t)) != DD_OK) return rc;
    if ((rc = stage_sources(b, t0[0], dt[0], n_t == 1, true)) != DD_OK) return rc;
    New.t0.assign(t0, t0 + n_t);
    New.dt.assign(dt, dt + n_t);
    New.n_t = n_t;
    if ((rc = pc_step_enqueue(b, slot_in, slot_out, opt, New)) != DD_OK) return rc;
    if (Old.active) {
        bool ok = false;
        if ((rc = pc_step_finish(b, Old, prev_stats, &ok)) != DD_OK) return rc;
        if (have_prev) *have_prev = 1;
        if (prev_stats) prev_stats->retries = 0;
        if (ok) {
            record_step(b, Old.slot_in, Old.slot_out);
        } else {
            // the step just enqueued started from a rejected state: drop it, redo both
            if (New.slot_out == Old.slot_in)
                return fail(ctx, DD_ERR_NOT_CONVERGED,
                            "a deferred step was rejected after its input slot had been reused; rotate three slots");
            CK(cudaStreamSynchronize(ctx->stream));
            New.active = false;
            if ((rc = redo_step(b, Old, prev_stats)) != DD_OK) return rc;
            if ((rc = set_times(b, New.t0.data(), New.dt.data(), n_t)) != DD_OK) return rc;
            if ((rc = stage_sources(b, New.t0[0], New.dt[0], n_t == 1, true)) != DD_OK) return rc;
            if ((rc = pc_step_enqueue(b, slot_in, slot_out, opt, New)) != DD_OK) return rc;
        }
    }
    b->rec_cur ^= 1;
    return DD_OK;
}

// ---------------------------------------------------------------------------
// combined max-integral error norms on the device (csrc/dd_combine.cuh): running state per member, updated
// after every step -- the norm series is never stored.  out[m][6] = overall, cp, T, cl, cd, cs.
// ---------------------------------------------------------------------------
__global__ void k_combine_update(const double* __restrict__ norms, int B, const double* __restrict__ dt, int n_t,
                                 int first, double* __restrict__ state) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= B) return;
    dd_combine_fold(norms + (size_t)m * 8, dt[n_t == 1 ? 0 : m], first, state + (size_t)m * 18);
}

__global__ void k_combine_final(const double* __restrict__ state, int B, double* __restrict__ out) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= B) return;
    for (int q = 0; q < 6; ++q) out[(size_t)m * 6 + q] = __dsqrt_rn(state[(size_t)m * 18 + q]);
}

static int ensure_combine(dd_batch* b) {
    dd_ctx* ctx = b->ctx;
    if (!b->d_combine_state) CK(cudaMalloc((void**)&b->d_combine_state, sizeof(double) * 18 * b->B));
    if (!b->d_combined) CK(cudaMalloc((void**)&b->d_combined, sizeof(double) * 6 * b->B));
    return DD_OK;
}

// error norms of `slot` at the members' current time, folded into the running combination
static int norms_combine_async(dd_batch* b, int slot, int n_t, bool first) {
    dd_ctx* ctx = b->ctx;
    int rc = norms_async(b, slot, -1, nullptr, b->d_norm_out);
    if (rc != DD_OK) return rc;
    k_combine_update<<<(b->B + 127) / 128, 128, 0, ctx->stream>>>(b->d_norm_out, b->B, b->d_dt, n_t, first ? 1 : 0,
                                                                  b->d_combine_state);
    CK(cudaGetLastError());
    return DD_OK;
}

static int combine_async(dd_batch* b, double* combined_host) {
    dd_ctx* ctx = b->ctx;
    k_combine_final<<<(b->B + 127) / 128, 128, 0, ctx->stream>>>(b->d_combine_state, b->B, b->d_combined);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(combined_host, b->d_combined, sizeof(double) * 6 * b->B, cudaMemcpyDeviceToHost, ctx->stream));
    return DD_OK;
}

// Small grids (the working set of a member fits in shared memory): whole trajectories on chip, one CTA per member
// (dd_member.cu).  Possible when the forcing is evaluated on the device from tables, every member is a RegHCsTriple
// one, the step is the default p = q = 1 one and no per-step norm series is asked for.  OPT-IN (DD_MEMBER_KERNEL=1):
// measured on B200 (12 500 members of 33 x 33 nodes, profiles/README.md) it reaches 1.70e9 cell-steps/s against
// 2.12e9 of the batched mesh kernels -- 16 warps per SM and the sources re-evaluated in every phase (there is no
// room for staged source arrays next to the 19 work arrays) cost more than the HBM round trips it saves.
static bool member_path_ok(const dd_batch* b, const dd_pc_options& opt, int nsteps, const double* norms_out) {
    const char* on = getenv("DD_MEMBER_KERNEL");  // read per call: the tests compare the two paths in one process
    if (!(on && *on && *on != '0') || norms_out || nsteps < 1 || b->is_slab || b->has_variants || b->fused_sources)
        return false;
    if (b->mode != DD_FORCING_SEPARABLE && b->mode != DD_FORCING_EXPSIN) return false;
    if (opt.num_pc_steps != 1 || opt.num_newton_steps != 1 || opt.extrapolate_guess) return false;
    if (opt.num_newton_iterations > 100000) return false;
    for (int m = 0; m < b->B; ++m)
        for (int v = 0; v < DD_NVAR; ++v)
            if (b->mode == DD_FORCING_SEPARABLE && b->h_mem[m].phi_kind[v] == DD_PHI_HOST) return false;
    return dd_member_fits(b->N, b->M);
}

static int run_member_kernel(dd_batch* b, int slot_a, int slot_b, int nsteps, const dd_pc_options& opt,
                             double* combined_out, dd_step_stats* stats) {
    dd_ctx* ctx = b->ctx;
    int rc;
    if (combined_out && (rc = ensure_combine(b)) != DD_OK) return rc;
    if (!b->d_member_stats) CK(cudaMalloc((void**)&b->d_member_stats, sizeof(double) * 16 * b->B));
    DDMemberArgs A;
    memset(&A, 0, sizeof(A));
    A.g = b->g;
    A.mem = b->d_mem;
    A.mem_rw = b->d_mem;
    A.F = b->F;
    A.in = cstate(b, slot_a);
    A.out = mstate(b, nsteps % 2 == 0 ? slot_a : slot_b);
    A.nmembers = b->B;
    A.nsteps = nsteps;
    A.cap = opt.num_newton_iterations;
    A.rtol = (opt.consec_xs_rtol > 0.0 && opt.num_newton_iterations > 0) ? opt.consec_xs_rtol : 0.0;
    A.cd_swap = opt.cd_band_swap;
    A.fixed_sweeps = opt.fixed_sweeps;
    A.max_sweeps = opt.max_sweeps;
    A.solve_tol = opt.solve_tol;
    A.combined = combined_out ? b->d_combined : nullptr;
    A.stats = b->d_member_stats;
    {
        ProfScope ps_(ctx->stream, PC_OTHER, 1);
        CK(dd_launch_member_run(ctx->stream, b->mode, A, ctx->sm_count));
    }
    std::vector<double> hs((size_t)16 * b->B);
    CK(cudaMemcpyAsync(hs.data(), b->d_member_stats, sizeof(double) * hs.size(), cudaMemcpyDeviceToHost, ctx->stream));
    if (combined_out)
        CK(cudaMemcpyAsync(combined_out, b->d_combined, sizeof(double) * 6 * b->B, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    b->prev_valid = false;
    b->src_has[0] = b->src_has[1] = false;
    dd_step_stats st;
    memset(&st, 0, sizeof(st));
    bool failed = false;
    for (int m = 0; m < b->B; ++m) {
        if (!b->h_mem[m].active) continue;
        const double* q = hs.data() + (size_t)16 * m;
        for (int k = 0; k < 3; ++k) {
            if ((int)q[k] > st.sweeps[k]) st.sweeps[k] = (int)q[k];
            st.passes[k] = 1;
            if (q[3 + k] > st.rho[k] || q[3 + k] != q[3 + k]) st.rho[k] = q[3 + k];
            if (q[6 + k] > st.resid[k]) st.resid[k] = q[6 + k];
            if (q[9 + k] > st.bound[k]) st.bound[k] = q[9 + k];
        }
        if ((int)q[12] > st.cs_newton_iters) st.cs_newton_iters = (int)q[12];
        if (q[13] != 0.0) failed = true;
    }
    if (stats) *stats = st;
    if (failed) return fail(ctx, DD_ERR_NOT_CONVERGED, "linear solve of a member did not reach the residual bound");
    return DD_OK;
}

static int run_pc_loop(dd_batch* b, int slot_a, int slot_b, const double* t0, const double* dt, int n_t, int nsteps,
                       const dd_pc_options* opt_in, double* norms_out, double* combined_out, dd_step_stats* stats) {
    if (!b || !slot_ok(b, slot_a) || !slot_ok(b, slot_b) || slot_a == slot_b || nsteps < 0) return DD_ERR_INVALID;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    dd_pc_options opt;
    if (opt_in) opt = *opt_in; else dd_pc_options_default(&opt);
    int rc = check_opts(ctx, opt);
    if (rc != DD_OK) return rc;
    if ((rc = set_times(b, t0, dt, n_t)) != DD_OK) return rc;
    if (member_path_ok(b, opt, nsteps, norms_out)) return run_member_kernel(b, slot_a, slot_b, nsteps, opt, combined_out, stats);
    int cur = slot_a, nxt = slot_b;
    const size_t nstride = (size_t)8 * b->B;
    if (combined_out && (rc = ensure_combine(b)) != DD_OK) return rc;
    if (norms_out && (rc = norms_async(b, cur, -1, norms_out)) != DD_OK) return rc;
    if (combined_out && (rc = norms_combine_async(b, cur, n_t, true)) != DD_OK) return rc;
    double th = t0[0];  // host mirror of the device-side time advance (same IEEE additions)
    int carry = -1;
    for (int s = 0; s < nsteps; ++s) {
        if ((rc = stage_sources(b, th, dt[0], n_t == 1, true, &carry)) != DD_OK) return rc;
        th = th + dt[0];
        if ((rc = pc_step_retry(b, cur, nxt, opt, stats)) != DD_OK) return rc;
        CKP(PC_TIME, 1, dd_launch_time_coefs(launch_of(b), b->mode, b->d_mem, b->d_t0, b->d_dt, n_t, 1));
        if (norms_out && (rc = norms_async(b, nxt, -1, norms_out + (s + 1) * nstride)) != DD_OK) return rc;
        if (combined_out && (rc = norms_combine_async(b, nxt, n_t, false)) != DD_OK) return rc;
        const int tmp = cur; cur = nxt; nxt = tmp;
    }
    if (combined_out && (rc = combine_async(b, combined_out)) != DD_OK) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return DD_OK;
}

extern "C" int dd_run_pc(dd_batch* b, int slot_a, int slot_b, const double* t0, const double* dt, int n_t, int nsteps,
                         const dd_pc_options* opt_in, double* norms_out, dd_step_stats* stats) {
    return run_pc_loop(b, slot_a, slot_b, t0, dt, n_t, nsteps, opt_in, norms_out, nullptr, stats);
}

extern "C" int dd_run_pc_errors(dd_batch* b, int slot_a, int slot_b, const double* t0, const double* dt, int n_t,
                                int nsteps, const dd_pc_options* opt_in, double* combined_out, dd_step_stats* stats) {
    if (!combined_out) return DD_ERR_INVALID;
    return run_pc_loop(b, slot_a, slot_b, t0, dt, n_t, nsteps, opt_in, nullptr, combined_out, stats);
}

static int run_feuler_loop(dd_batch* b, int slot_a, int slot_b, const double* t0, const double* dt, int n_t,
                           int nsteps, double* norms_out, double* combined_out) {
    if (!b || !slot_ok(b, slot_a) || !slot_ok(b, slot_b) || slot_a == slot_b || nsteps < 0) return DD_ERR_INVALID;
    b->prev_valid = false;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    int rc;
    if ((rc = set_times(b, t0, dt, n_t)) != DD_OK) return rc;
    int cur = slot_a, nxt = slot_b;
    const size_t nstride = (size_t)8 * b->B;
    if (combined_out && (rc = ensure_combine(b)) != DD_OK) return rc;
    if (norms_out && (rc = norms_async(b, cur, -1, norms_out)) != DD_OK) return rc;
    if (combined_out && (rc = norms_combine_async(b, cur, n_t, true)) != DD_OK) return rc;
    double th = t0[0];
    for (int s = 0; s < nsteps; ++s) {
        if ((rc = stage_sources(b, th, dt[0], n_t == 1, false)) != DD_OK) return rc;
        th = th + dt[0];
        CKP(PC_FEULER, 1, dd_launch_feuler(launch_of(b), b->smode, b->g, b->d_mem, b->sF, cstate(b, cur), mstate(b, nxt)));
        CKP(PC_TIME, 1, dd_launch_time_coefs(launch_of(b), b->mode, b->d_mem, b->d_t0, b->d_dt, n_t, 1));
        if (norms_out && (rc = norms_async(b, nxt, -1, norms_out + (s + 1) * nstride)) != DD_OK) return rc;
        if (combined_out && (rc = norms_combine_async(b, nxt, n_t, false)) != DD_OK) return rc;
        const int tmp = cur; cur = nxt; nxt = tmp;
    }
    if (combined_out && (rc = combine_async(b, combined_out)) != DD_OK) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return DD_OK;
}

extern "C" int dd_run_feuler(dd_batch* b, int slot_a, int slot_b, const double* t0, const double* dt, int n_t,
                             int nsteps, double* norms_out) {
    return run_feuler_loop(b, slot_a, slot_b, t0, dt, n_t, nsteps, norms_out, nullptr);
}

extern "C" int dd_run_feuler_errors(dd_batch* b, int slot_a, int slot_b, const double* t0, const double* dt, int n_t,
                                    int nsteps, double* combined_out) {
    if (!combined_out) return DD_ERR_INVALID;
    return run_feuler_loop(b, slot_a, slot_b, t0, dt, n_t, nsteps, nullptr, combined_out);
}


// ---------------------------------------------------------------------------
// pieces of the step (class-level API)
// ---------------------------------------------------------------------------
extern "C" int dd_pc_predict(dd_batch* b, int slot_in, const double* t0, const double* dt, int n_t) {
    if (!b || !slot_ok(b, slot_in)) return DD_ERR_INVALID;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    int rc;
    if ((rc = set_times(b, t0, dt, n_t)) != DD_OK) return rc;
    if ((rc = stage_sources(b, t0[0], dt[0], n_t == 1, true)) != DD_OK) return rc;
    DDPredictOut po;
    if ((rc = get_work(b, "cp1p", &po.cp1p)) != DD_OK) return rc;
    if ((rc = get_work(b, "cs1p", &po.cs1p)) != DD_OK) return rc;
    if ((rc = get_work(b, "YT", &po.YT)) != DD_OK) return rc;
    if ((rc = get_work(b, "Ycl", &po.Ycl)) != DD_OK) return rc;
    if ((rc = get_work(b, "Ycd", &po.Ycd)) != DD_OK) return rc;
    CK(dd_launch_predict(launch_of(b), b->smode, b->g, b->d_mem, b->sF, cstate(b, slot_in), po));
    CK(cudaStreamSynchronize(ctx->stream));
    return DD_OK;
}

extern "C" int dd_pc_newton(dd_batch* b, int var, int slot_star, int slot_new, const double* t0, const double* dt,
                            int n_t, const dd_pc_options* opt_in, dd_step_stats* stats) {
    if (!b || !slot_ok(b, slot_star) || !slot_ok(b, slot_new) || slot_star == slot_new || var < DD_T || var > DD_CD)
        return DD_ERR_INVALID;
    b->prev_valid = false;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    dd_pc_options opt;
    if (opt_in) opt = *opt_in; else dd_pc_options_default(&opt);
    int rc = check_opts(ctx, opt);
    if (rc != DD_OK) return rc;
    if ((rc = set_times(b, t0, dt, n_t)) != DD_OK) return rc;
    if ((rc = stage_sources(b, t0[0], dt[0], n_t == 1, true)) != DD_OK) return rc;
    if ((rc = ensure_solve_slots(b, 3)) != DD_OK) return rc;
    static const char* yname[3] = {"YT", "Ycl", "Ycd"};
    double* Y;
    if ((rc = get_work(b, yname[var - DD_T], &Y)) != DD_OK) return rc;
    const DDStateC u = cstate(b, slot_star);
    const DDState nw = mstate(b, slot_new);
    b->cm = 0;
    for (int attempt = 0;; ++attempt) {
        int sw = 0, pa = 0;
        if ((rc = newton_solve(b, var, u, nw.v[DD_T], nw.v[DD_CL], Y, nw.v[var], opt, 0, &sw, &pa)) != DD_OK) return rc;
        SolveSummary s;
        CK(cudaMemcpyAsync(&s, b->d_summary, sizeof(s), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        const int vi = var - DD_T;
        if (stats) {
            stats->sweeps[vi] = sw; stats->passes[vi] = pa; stats->rho[vi] = s.rho; stats->resid[vi] = s.resid;
            stats->bound[vi] = s.bound; stats->retries = attempt;
        }
        if (s.ratio <= 1.0) return DD_OK;
        if (opt.fixed_sweeps > 0 || b->ctl[b->cm].sweeps[vi] >= opt.max_sweeps || attempt >= 40)
            return fail(ctx, DD_ERR_NOT_CONVERGED, "linear solve did not reach the residual bound");
        const int want = sweeps_for_rho(s.rho * 1.02 + 1e-12, opt.max_sweeps);
        if (b->ctl[b->cm].sweeps[vi] >= want) b->ctl[b->cm].extra[vi] += (b->ctl[b->cm].sweeps[vi] + 1) / 2 + 1;
        const int next = want + b->ctl[b->cm].extra[vi];
        b->ctl[b->cm].sweeps[vi] = next > opt.max_sweeps ? opt.max_sweeps : next;
    }
}

extern "C" int dd_pc_correct(dd_batch* b, int slot0, int slot_new, const double* t0, const double* dt, int n_t,
                             const dd_pc_options* opt_in, int* cs_iters_out) {
    if (!b || !slot_ok(b, slot0) || !slot_ok(b, slot_new) || slot0 == slot_new) return DD_ERR_INVALID;
    b->prev_valid = false;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    dd_pc_options opt;
    if (opt_in) opt = *opt_in; else dd_pc_options_default(&opt);
    int rc = check_opts(ctx, opt);
    if (rc != DD_OK) return rc;
    if ((rc = set_times(b, t0, dt, n_t)) != DD_OK) return rc;
    if ((rc = stage_sources(b, t0[0], dt[0], n_t == 1, true)) != DD_OK) return rc;
    const DDLaunch L = launch_of(b, ROWS_ALL);
    const DDStateC s0 = cstate(b, slot0);
    const DDState nw = mstate(b, slot_new);
    const int cap = opt.num_newton_iterations;
    const bool track = opt.consec_xs_rtol > 0.0 && cap > 0;
    if (track) {
        const int had = b->cs_cap_alloc;
        if ((rc = ensure_cs_buffers(b, cap)) != DD_OK) return rc;
        if (b->cs_cap_alloc != had) {
            const long long n = (long long)b->cs_cap_alloc * b->B;
            k_cs_arm<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(b->d_itmax, b->d_itmin, n);
        }
    }
    CK(cudaMemsetAsync(b->d_flags, 0, sizeof(int) * b->B, ctx->stream));
    CK(dd_launch_correct(L, b->smode, b->g, b->d_mem, b->sF, s0, nw.v[DD_T], nw.v[DD_CL], nw.v[DD_CD], nw.v[DD_CP],
                         nw.v[DD_CS], cap, track ? opt.consec_xs_rtol : 0.0, b->d_itmax, b->d_itmin, b->has_variants ? b->d_flags : nullptr));
    if (track)
        CK(dd_launch_cs_finish(L, b->smode, b->g, b->d_mem, b->sF, s0, nw.v[DD_CL], nw.v[DD_CD], nw.v[DD_CS], cap,
                               opt.consec_xs_rtol, b->d_itmax, b->d_itmin, b->d_used));
    if (cs_iters_out) {
        if (track) {
            CK(cudaMemcpyAsync(cs_iters_out, b->d_used, sizeof(int) * b->B, cudaMemcpyDeviceToHost, ctx->stream));
        } else {
            for (int m = 0; m < b->B; ++m) cs_iters_out[m] = cap;
        }
    }
    std::vector<int> flags(b->B);
    CK(cudaMemcpyAsync(flags.data(), b->d_flags, sizeof(int) * b->B, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int m = 0; m < b->B; ++m)
        if (b->h_mem[m].active && (flags[m] & 1))
            return fail(ctx, DD_ERR_DOMAIN, "Denominator 2 - dt Kd (Sd - Cd1) (1 + Cl1) below positiveness treshold.");
    return DD_OK;
}

extern "C" int dd_pc_residual(dd_batch* b, int var, int slot_state, const double* t0, const double* dt, int n_t,
                              double* out_host) {
    if (!b || !slot_ok(b, slot_state) || var < DD_T || var > DD_CD || !out_host) return DD_ERR_INVALID;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    int rc;
    if ((rc = set_times(b, t0, dt, n_t)) != DD_OK) return rc;
    if ((rc = stage_sources(b, t0[0], dt[0], n_t == 1, true)) != DD_OK) return rc;
    static const char* yname[3] = {"YT", "Ycl", "Ycd"};
    double *Y, *res;
    if ((rc = get_work(b, yname[var - DD_T], &Y)) != DD_OK) return rc;
    if ((rc = get_work(b, "resid", &res)) != DD_OK) return rc;
    CK(dd_launch_residual(launch_of(b), b->smode, var, b->g, b->d_mem, b->sF, cstate(b, slot_state), Y, res));
    CK(cudaMemcpyAsync(out_host, res, b->field_elems * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DD_OK;
}


// ---------------------------------------------------------------------------
// phased PC step for slab-decomposed meshes: the caller exchanges halo rows between phases
//   0: times + predict            (needs halos of the five input fields)
//   1/2/3: assemble + solve T/cl/cd -> slot_out (owned rows); exchange that field's halo afterwards
//   4: correctors on all local rows (needs halos of T1, cl1, cd1); accumulates the cs exit statistics
//   5: cs exit decision (after the caller reduced "cs_it_max"/"cs_it_min" over ranks) and the solve
//      summaries: summary[3][4] = rho, ratio (<= 1: bound met), resid, bound for T, cl, cd
// Only num_pc_steps = num_newton_steps = 1.
// ---------------------------------------------------------------------------
extern "C" int dd_batch_set_plan(dd_batch* b, const int sweeps[3]) {
    if (!b || !sweeps) return DD_ERR_INVALID;
    for (int q = 0; q < 3; ++q) b->ctl[0].sweeps[q] = b->ctl[1].sweeps[q] = sweeps[q];
    return DD_OK;
}

extern "C" int dd_batch_set_relax_rho(dd_batch* b, const double rho[3]) {
    if (!b) return DD_ERR_INVALID;
    for (int q = 0; q < 3; ++q) b->relax_rho[q] = (rho && rho[q] >= 0.0 && rho[q] < 1.0) ? rho[q] : -1.0;
    return DD_OK;
}

extern "C" int dd_batch_get_plan(dd_batch* b, int sweeps[3]) {
    if (!b || !sweeps) return DD_ERR_INVALID;
    for (int q = 0; q < 3; ++q) sweeps[q] = b->ctl[b->cm].sweeps[q];
    return DD_OK;
}

extern "C" int dd_sweeps_for_rho(double rho, int max_sweeps) { return sweeps_for_rho(rho, max_sweeps); }

extern "C" int dd_next_plan(int cur, double rho, double ratio, int max_sweeps) {
    return next_plan(cur, rho, ratio, max_sweeps);
}

extern "C" int dd_step_pc_phase(dd_batch* b, int phase, int slot_in, int slot_out, const double* t0, const double* dt,
                                int n_t, const dd_pc_options* opt_in, double* summary, int* cs_iters) {
    if (!b || !slot_ok(b, slot_in) || !slot_ok(b, slot_out) || slot_in == slot_out) return DD_ERR_INVALID;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    { const int rcf_ = flush_pending(b, nullptr, nullptr); if (rcf_ != DD_OK) return rcf_; }
    dd_pc_options opt;
    if (opt_in) opt = *opt_in; else dd_pc_options_default(&opt);
    int rc = check_opts(ctx, opt);
    if (rc != DD_OK) return rc;
    if (opt.num_pc_steps != 1 || opt.num_newton_steps != 1)
        return fail(ctx, DD_ERR_INVALID, "phased step supports num_pc_steps = num_newton_steps = 1");
    if ((rc = ensure_solve_slots(b, 4)) != DD_OK) return rc;  // three solves + the row the domain flag is filed in
    DDPredictOut po;
    if ((rc = get_work(b, "cp1p", &po.cp1p)) != DD_OK) return rc;
    if ((rc = get_work(b, "cs1p", &po.cs1p)) != DD_OK) return rc;
    if ((rc = get_work(b, "YT", &po.YT)) != DD_OK) return rc;
    if ((rc = get_work(b, "Ycl", &po.Ycl)) != DD_OK) return rc;
    if ((rc = get_work(b, "Ycd", &po.Ycd)) != DD_OK) return rc;
    const DDStateC s0 = cstate(b, slot_in);
    const DDState sout = mstate(b, slot_out);
    DDStateC u;
    u.v[DD_CP] = po.cp1p; u.v[DD_T] = s0.v[DD_T]; u.v[DD_CL] = s0.v[DD_CL]; u.v[DD_CD] = s0.v[DD_CD];
    u.v[DD_CS] = po.cs1p;
    const int cap = opt.num_newton_iterations;
    const bool track = opt.consec_xs_rtol > 0.0 && cap > 0;
    int sw = 0, pa = 0;
    switch (phase) {
        case 0:
            CK(cudaMemsetAsync(b->d_flags, 0, sizeof(int) * b->B, ctx->stream));
            if ((rc = set_times(b, t0, dt, n_t)) != DD_OK) return rc;
            if ((rc = stage_sources(b, t0[0], dt[0], n_t == 1, true)) != DD_OK) return rc;
            if ((rc = launch_predictor(b, s0, po, &b->phase_fused_T, false)) != DD_OK) return rc;
            b->use_guess = take_guess(b, slot_in, slot_out, opt);
            if (track) {
                const int had = b->cs_cap_alloc;
                if ((rc = ensure_cs_buffers(b, cap)) != DD_OK) return rc;
                if (b->cs_cap_alloc != had) {
                    const long long n = (long long)b->cs_cap_alloc * b->B;
                    k_cs_arm<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(b->d_itmax, b->d_itmin, n);
                }
            }
            return DD_OK;
        case 1: case 21: case 31: {
            int what = phase == 1 ? 3 : (phase == 21 ? 1 : 2);
            if (b->phase_fused_T) what &= 2;  // phase 0 assembled the T system already
            if (!what) return DD_OK;
            return newton_solve(b, DD_T, u, nullptr, nullptr, po.YT, sout.v[DD_T], opt, 0, &sw, &pa, what,
                                b->use_guess ? sout.v[DD_T] : nullptr);
        }
        case 2: case 22: case 32:
            return newton_solve(b, DD_CL, u, sout.v[DD_T], nullptr, po.Ycl, sout.v[DD_CL], opt, 1, &sw, &pa,
                                phase == 2 ? 3 : (phase == 22 ? 1 : 2), b->use_guess ? sout.v[DD_CL] : nullptr);
        case 3: case 23: case 33:
            return newton_solve(b, DD_CD, u, sout.v[DD_T], sout.v[DD_CL], po.Ycd, sout.v[DD_CD], opt, 2, &sw, &pa,
                                phase == 3 ? 3 : (phase == 23 ? 1 : 2), b->use_guess ? sout.v[DD_CD] : nullptr);
        case 4:
            CKP(PC_CORRECT, 1,
                dd_launch_correct(launch_of(b, ROWS_ALL), b->smode, b->g, b->d_mem, b->sF, s0, sout.v[DD_T],
                                  sout.v[DD_CL], sout.v[DD_CD], sout.v[DD_CP], sout.v[DD_CS], cap,
                                  track ? opt.consec_xs_rtol : 0.0, b->d_itmax, b->d_itmin, b->has_variants ? b->d_flags : nullptr));
            return DD_OK;
        case 6:
            // as 5 without the read-back: the driver reduces "summary" / "cs_used" on the device itself
            if (track)
                CKP(PC_CS_FINISH, 2,
                    dd_launch_cs_finish(launch_of(b, ROWS_ALL), b->smode, b->g, b->d_mem, b->sF, s0, sout.v[DD_CL],
                                        sout.v[DD_CD], sout.v[DD_CS], cap, opt.consec_xs_rtol, b->d_itmax,
                                        b->d_itmin, b->d_used));
            // the domain-error flags of the corrector (HCsTriple: 2 - dt R1 not positive) ride in row 3 of the
            // summaries, which the slab driver max-reduces over the ranks
            k_flags_to_summary<<<1, 32, 0, ctx->stream>>>(b->d_flags, b->d_mem, b->B, b->d_summary + 3);
            CK(cudaGetLastError());
            record_step(b, slot_in, slot_out);
            return DD_OK;
        case 5: {
            if (track)
                CKP(PC_CS_FINISH, 2,
                    dd_launch_cs_finish(launch_of(b, ROWS_ALL), b->smode, b->g, b->d_mem, b->sF, s0, sout.v[DD_CL],
                                        sout.v[DD_CD], sout.v[DD_CS], cap, opt.consec_xs_rtol, b->d_itmax,
                                        b->d_itmin, b->d_used));
            SolveSummary sums[3];
            CK(cudaMemcpyAsync(sums, b->d_summary, sizeof(sums), cudaMemcpyDeviceToHost, ctx->stream));
            std::vector<int> hflags(b->B, 0);
            CK(cudaMemcpyAsync(hflags.data(), b->d_flags, sizeof(int) * b->B, cudaMemcpyDeviceToHost, ctx->stream));
            int used0 = cap;
            if (track && cs_iters) CK(cudaMemcpyAsync(&used0, b->d_used, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            if (summary)
                for (int q = 0; q < 3; ++q) {
                    summary[q * 4 + 0] = sums[q].rho;
                    summary[q * 4 + 1] = sums[q].ratio;
                    summary[q * 4 + 2] = sums[q].resid;
                    summary[q * 4 + 3] = sums[q].bound;
                }
            if (cs_iters) *cs_iters = used0;
            for (int m = 0; m < b->B; ++m)
                if (b->h_mem[m].active && (hflags[m] & 1))
                    return fail(ctx, DD_ERR_DOMAIN,
                                "Denominator 2 - dt Kd (Sd - Cd1) (1 + Cl1) below positiveness treshold.");
            record_step(b, slot_in, slot_out);
            return DD_OK;
        }
        default:
            return fail(ctx, DD_ERR_INVALID, "phase must be 0..6, 21..23 or 31..33");
    }
}


// One segment of a Newton solve of the phased (slab) step: `sweeps` SOR sweeps of variable var (1 = T, 2 = cl,
// 3 = cd) continuing from the iterate the previous segment left (first = 1: from zero).  Unless last = 1 the
// iterate stays in a work array ("x_cur" of dd_work_dev_ptr), whose halo rows the slab driver exchanges before the
// next segment: a halo of G rows supports (G - 3) / 2 sweeps between two exchanges, and with segments any number
// of sweeps (weakly dominant matrices of large time steps) while the result stays the global red-black iteration,
// bit for bit.  The system must have been assembled (phases 21 - 23).
extern "C" int dd_pc_solve_segment(dd_batch* b, int var, int slot_in, int slot_out, const dd_pc_options* opt_in,
                                   int sweeps, int first, int last) {
    if (!b || !slot_ok(b, slot_in) || !slot_ok(b, slot_out) || slot_in == slot_out || var < DD_T || var > DD_CD ||
        sweeps < 1)
        return DD_ERR_INVALID;
    dd_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    dd_pc_options opt;
    if (opt_in) opt = *opt_in; else dd_pc_options_default(&opt);
    int rc = check_opts(ctx, opt);
    if (rc != DD_OK) return rc;
    if (!first && !b->x_cur) return fail(ctx, DD_ERR_INVALID, "solve segment without a predecessor");
    if ((rc = ensure_solve_slots(b, 4)) != DD_OK) return rc;
    DDPredictOut po;
    if ((rc = get_work(b, "cp1p", &po.cp1p)) != DD_OK) return rc;
    if ((rc = get_work(b, "cs1p", &po.cs1p)) != DD_OK) return rc;
    if ((rc = get_work(b, "YT", &po.YT)) != DD_OK) return rc;
    if ((rc = get_work(b, "Ycl", &po.Ycl)) != DD_OK) return rc;
    if ((rc = get_work(b, "Ycd", &po.Ycd)) != DD_OK) return rc;
    const DDStateC s0 = cstate(b, slot_in);
    const DDState sout = mstate(b, slot_out);
    DDStateC u;
    u.v[DD_CP] = po.cp1p; u.v[DD_T] = s0.v[DD_T]; u.v[DD_CL] = s0.v[DD_CL]; u.v[DD_CD] = s0.v[DD_CD];
    u.v[DD_CS] = po.cs1p;
    const double* Y = var == DD_T ? po.YT : (var == DD_CL ? po.Ycl : po.Ycd);
    int sw = 0, pa = 0;
    const double* xlast = nullptr;
    rc = newton_solve(b, var, u, sout.v[DD_T], sout.v[DD_CL], Y, sout.v[var], opt, var - DD_T, &sw, &pa, 2, nullptr,
                      true, sweeps, first ? nullptr : b->x_cur, last != 0, &xlast);
    b->x_cur = last ? nullptr : xlast;
    return rc;
}

// ---- halo exchange by direct peer stores (multi-process slab driver; see include/dd_b200.h) --------------------
extern "C" int dd_ipc_export(dd_ctx* ctx, const void* dev_ptr, unsigned char* handle64) {
    if (!ctx || !dev_ptr || !handle64) return DD_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles travel as 64 bytes");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
    memcpy(handle64, &h, 64);
    return DD_OK;
}
extern "C" int dd_ipc_import(dd_ctx* ctx, const unsigned char* handle64, void** dev_ptr) {
    if (!ctx || !handle64 || !dev_ptr) return DD_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DD_OK;
}
extern "C" int dd_ipc_close(dd_ctx* ctx, void* dev_ptr) {
    if (!ctx || !dev_ptr) return DD_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaIpcCloseMemHandle(dev_ptr));
    return DD_OK;
}
// flag block of a rank: 4 handshake words its neighbours write + the kernel's block counter + a status word
extern "C" int dd_halo_flags_create(dd_ctx* ctx, void** flags) {
    if (!ctx || !flags) return DD_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMalloc(flags, 64));
    CK(cudaMemsetAsync(*flags, 0, 64, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DD_OK;
}
extern "C" int dd_halo_flags_destroy(dd_ctx* ctx, void* flags) {
    if (!ctx) return DD_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (flags) CK(cudaFree(flags));
    return DD_OK;
}
extern "C" int dd_halo_push(dd_ctx* ctx, const double* src_top, double* dst_up, const double* src_bot,
                            double* dst_down, long long count, void* my_flags, void* up_flags, void* down_flags,
                            unsigned seq) {
    if (!ctx || !my_flags || count < 1 || (src_top && (!dst_up || !up_flags)) ||
        (src_bot && (!dst_down || !down_flags)))
        return DD_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    unsigned* f = static_cast<unsigned*>(my_flags);
    g_prof.launches += 1;
    CK(dd_launch_halo_push(ctx->stream, src_top, dst_up, src_bot, dst_down, count, f, static_cast<unsigned*>(up_flags),
                           static_cast<unsigned*>(down_flags), seq, f + 4, reinterpret_cast<int*>(f + 5)));
    return DD_OK;
}
// 0: every handshake so far completed; 1: a wait timed out (the neighbour never arrived)
extern "C" int dd_halo_status(dd_ctx* ctx, void* my_flags, int* status) {
    if (!ctx || !my_flags || !status) return DD_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(status, static_cast<unsigned*>(my_flags) + 5, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DD_OK;
}

// accuracy probe of the inline device exp / reciprocal (host arrays in, host arrays out)
extern "C" int dd_probe_math(dd_ctx* ctx, int n, const double* in, double* out_exp, double* out_rcp) {
    if (!ctx || n < 1 || !in || !out_exp || !out_rcp) return DD_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    double* d = nullptr;
    CK(cudaMalloc((void**)&d, 3 * sizeof(double) * (size_t)n));
    CK(cudaMemcpyAsync(d, in, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(dd_launch_probe_math(ctx->stream, d, d + n, d + 2 * (size_t)n, n));
    CK(cudaMemcpyAsync(out_exp, d + n, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out_rcp, d + 2 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    return DD_OK;
}

// fp64 roof of the device (see include/dd_b200.h)
extern "C" int dd_probe_fp64(dd_ctx* ctx, double ms_target, double* tflops) {
    if (!ctx || !tflops || !(ms_target > 0.0)) return DD_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    double* sink = nullptr;
    CK(cudaMalloc((void**)&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int sm = ctx->sm_count > 0 ? ctx->sm_count : 148;
    const int blocks = 2 * sm;  // two 1024-thread CTAs per SM: all 64 warps resident
    int iters = 2000;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0, ctx->stream));
        CK(dd_launch_probe_fp64(ctx->stream, sink, blocks, iters));
        CK(cudaEventRecord(e1, ctx->stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8.0 * 16.0 * (double)iters * 1024.0 * (double)blocks;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;  // the first round sizes the loop
        if (ms > 0.f) {
            const double scale = ms_target / ms;
            iters = (int)fmin(2e6, fmax(100.0, iters * scale));
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *tflops = best;
    return DD_OK;
}

static char g_solver_kernel[3][64] = {"", "", ""};
void dd_note_solver_kernel(int var, const char* name) {
    if (var < DD_T || var > DD_CD) return;
    snprintf(g_solver_kernel[var - DD_T], sizeof(g_solver_kernel[0]), "%s", name);
}
extern "C" const char* dd_solver_kernel_name(int var) {
    if (var < DD_T || var > DD_CD) return "";
    return g_solver_kernel[var - DD_T];
}
