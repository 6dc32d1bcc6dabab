// qspush_capi.cu — C-ABI (include/qspush.h) over the sm_100a kernels.  No CPU compute path:
// every compute entry point requires a CUDA device and fails with QSPUSH_ERR_NO_DEVICE otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <climits>
#include <vector>

#include "../../include/qspush.h"
#include "qs_kernels.cuh"
#include "qs_model.hpp"

using namespace qs;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail(QSPUSH_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));       \
    } while (0)

struct qspush_model {
    HostModel hm;
    std::mutex mu;
    std::vector<double*> d_blob;   // lazily uploaded copy per device
};

struct qspush_solver {
    int device = 0, B = 0, Bp = 0, N = 0, nmodels = 0;
    double dt = 0.0;
    qspush_opts opts;
    qspush_ctrl ctrl;
    cudaStream_t stream = nullptr;
    void* arena = nullptr;
    size_t arena_bytes = 0;
    SolverDev dev;
    double *d_Wdt = nullptr, *d_We = nullptr, *d_H = nullptr, *d_QN = nullptr, *d_models = nullptr;
    double* d_stage = nullptr;
    int* d_istage = nullptr;
    int* d_order = nullptr;       // work-queue order of the warp QP kernel (k_qp_order)
    double* d_ref_traj = nullptr; // reference trajectory kept on the device (qspush_set_reference_trajectory): [T][6], [B][6]
    double* d_ref_off = nullptr;
    int ref_T = 0;
    size_t stage_doubles = 0;
    int* h_ndone = nullptr;        // pinned
    std::vector<double> W, We;     // host copies: N x 36 (y order, column-major), 16
    bool cost_dirty = true;
    cudaEvent_t ev[6];
    cudaEvent_t evs[4] = {nullptr, nullptr, nullptr, nullptr};   // full SQP: per-iteration phase events (linearise+residuals, QP+line search)
    double acc_lin = 0.0, acc_qp = 0.0;                          // ... accumulated over the iterations of the last solve (seconds)
    bool acc_valid = false;
    double t_tot = 0, t_lin = 0, t_qp = 0, t_prep = 0;
    long long launches = 0;
    bool smem_attr_set = false;
    // qspush_step: one CUDA graph per flag combination (captured on first use, dropped when options / cost / bounds change)
    cudaGraphExec_t step_exec[4] = {nullptr, nullptr, nullptr, nullptr};
    double* d_guess = nullptr;     // qspush_snapshot_guess: copy of the u slab
    double* d_ring[2] = {nullptr, nullptr};   // qspush_closed_loop input-delay rings (plant, controller), kept between calls
    size_t ring_doubles[2] = {0, 0};
    int* d_idx = nullptr;          // reference index of the current period (read by k_step_prepare)
    int step_launches = 0;         // kernels inside one step graph
    bool capturing = false;        // the step graph is being captured
    bool order_valid = false;      // d_order holds a permutation (written by k_step_out / k_qp_order)
    bool order_frozen = false;     // development aid (QSPUSH_DEV_ORDER builds): the order was set by the host and is kept
};
// phase events: plain records outside a capture; inside the capture of the step graph they must be EXTERNAL event-record
// nodes (a plain cudaEventRecord on a capturing stream only creates a capture-internal dependency, not a timed event)
static cudaError_t rec_event(qspush_solver* s, int i) {
    return s->capturing ? cudaEventRecordWithFlags(s->ev[i], s->stream, cudaEventRecordExternal) : cudaEventRecord(s->ev[i], s->stream);
}
static void drop_step_graphs(qspush_solver* s) {
    for (auto& g : s->step_exec) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
}

extern "C" {

const char* qspush_last_error(void) { return g_err.c_str(); }
const char* qspush_version(void) { return "qspush-b200 0.1 (sm_100a)"; }
int qspush_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ------------------------------------------------------------------------------------------------ model
int qspush_model_create(const double* knots, int nknots, const double* ctrl_xy, int n, int degree,
                        double mu_sp, double c_ellipse, int single_coeffs, qspush_model** out) {
    if (!out) return fail(QSPUSH_ERR_ARG, "out is NULL");
    qspush_model* m = new qspush_model();
    std::string e = model_from_tables(knots, nknots, ctrl_xy, n, degree, mu_sp, c_ellipse, single_coeffs != 0, m->hm);
    if (!e.empty()) { delete m; return fail(QSPUSH_ERR_ARG, e); }
    *out = m;
    return QSPUSH_OK;
}

int qspush_model_create_from_ply(const char* ply_path, int flip_order, int degree, double mu_sg,
                                 double mu_sp, double mass, double tau_max, qspush_model** out) {
    if (!out || !ply_path) return fail(QSPUSH_ERR_ARG, "NULL argument");
    qspush_model* m = new qspush_model();
    std::string e = model_from_ply(ply_path, flip_order != 0, degree, mu_sg, mu_sp, mass, tau_max, m->hm);
    if (!e.empty()) { delete m; return fail(QSPUSH_ERR_IO, e); }
    *out = m;
    return QSPUSH_OK;
}

void qspush_model_free(qspush_model* m) {
    if (!m) return;
    for (size_t d = 0; d < m->d_blob.size(); ++d)
        if (m->d_blob[d]) { cudaSetDevice((int)d); cudaFree(m->d_blob[d]); }
    delete m;
}

int qspush_model_info(const qspush_model* m, int* n, int* nknots, double* b, double* c_ellipse, double* mu_sp) {
    if (!m) return fail(QSPUSH_ERR_ARG, "model is NULL");
    if (n) *n = m->hm.n;
    if (nknots) *nknots = (int)m->hm.S.size();
    if (b) *b = m->hm.b;
    if (c_ellipse) *c_ellipse = m->hm.c_ellipse;
    if (mu_sp) *mu_sp = m->hm.mu_sp;
    return QSPUSH_OK;
}
int qspush_model_tables(const qspush_model* m, double* knots, double* ctrl_xy, double* c1, double* c2) {
    if (!m) return fail(QSPUSH_ERR_ARG, "model is NULL");
    if (knots) std::memcpy(knots, m->hm.S.data(), m->hm.S.size() * 8);
    if (ctrl_xy) std::memcpy(ctrl_xy, m->hm.P.data(), m->hm.P.size() * 8);
    if (c1) std::memcpy(c1, m->hm.c1.data(), m->hm.c1.size() * 8);
    if (c2) std::memcpy(c2, m->hm.c2.data(), m->hm.c2.size() * 8);
    return QSPUSH_OK;
}
}  // extern "C"

static int need_device(int device) {
    const int n = qspush_device_count();
    if (n <= 0) return fail(QSPUSH_ERR_NO_DEVICE, "no CUDA device visible: qspush has no CPU path");
    if (device < 0 || device >= n) return fail(QSPUSH_ERR_ARG, "device index out of range");
    return QSPUSH_OK;
}

static int model_on_device(const qspush_model* cm, int device, const double** out) {
    qspush_model* m = const_cast<qspush_model*>(cm);
    std::lock_guard<std::mutex> lk(m->mu);
    if ((int)m->d_blob.size() <= device) m->d_blob.resize(device + 1, nullptr);
    if (!m->d_blob[device]) {
        CK(cudaSetDevice(device));
        double* p = nullptr;
        CK(cudaMalloc(&p, MODEL_DOUBLES * 8));
        CK(cudaMemcpy(p, m->hm.blob.data(), MODEL_DOUBLES * 8, cudaMemcpyHostToDevice));
        m->d_blob[device] = p;
    }
    *out = m->d_blob[device];
    return QSPUSH_OK;
}

// Temporary device mirrors for the stateless HOST entry points.
struct TmpBuf {
    std::vector<void*> ptrs;
    ~TmpBuf() { for (void* p : ptrs) cudaFree(p); }
    int in(const double* h, size_t n, qspush_mem mem, const double** d) {
        if (!h) { *d = nullptr; return QSPUSH_OK; }
        if (mem == QSPUSH_MEM_DEVICE) { *d = h; return QSPUSH_OK; }
        double* p = nullptr;
        CK(cudaMalloc(&p, n * 8)); ptrs.push_back(p);
        CK(cudaMemcpy(p, h, n * 8, cudaMemcpyHostToDevice));
        *d = p; return QSPUSH_OK;
    }
    int out(double* h, size_t n, qspush_mem mem, double** d) {
        if (!h) { *d = nullptr; return QSPUSH_OK; }
        if (mem == QSPUSH_MEM_DEVICE) { *d = h; return QSPUSH_OK; }
        double* p = nullptr;
        CK(cudaMalloc(&p, n * 8)); ptrs.push_back(p);
        *d = p; return QSPUSH_OK;
    }
    static int back(double* h, const double* d, size_t n, qspush_mem mem) {
        if (!h || mem == QSPUSH_MEM_DEVICE) return QSPUSH_OK;
        CK(cudaMemcpy(h, d, n * 8, cudaMemcpyDeviceToHost));
        return QSPUSH_OK;
    }
};
#define RET(call) do { int r__ = (call); if (r__ != QSPUSH_OK) return r__; } while (0)

extern "C" {

// ------------------------------------------------------------------------------------------------ stateless eval
int qspush_eval_spline(const qspush_model* m, int device, qspush_mem mem, int cnt, const double* s, int wrap,
                       double* C, double* Cd, double* Cdd, double* tvers, double* nvers, double* kappa) {
    if (!m || !s || cnt < 0) return fail(QSPUSH_ERR_ARG, "bad argument");
    RET(need_device(device));
    CK(cudaSetDevice(device));
    if (cnt == 0) return QSPUSH_OK;
    const double* dm; RET(model_on_device(m, device, &dm));
    TmpBuf tb; const double* ds; double *dC, *dCd, *dCdd, *dt_, *dn, *dk;
    RET(tb.in(s, cnt, mem, &ds));
    RET(tb.out(C, 2 * (size_t)cnt, mem, &dC)); RET(tb.out(Cd, 2 * (size_t)cnt, mem, &dCd)); RET(tb.out(Cdd, 2 * (size_t)cnt, mem, &dCdd));
    RET(tb.out(tvers, 2 * (size_t)cnt, mem, &dt_)); RET(tb.out(nvers, 2 * (size_t)cnt, mem, &dn)); RET(tb.out(kappa, cnt, mem, &dk));
    k_eval_spline<<<(cnt + 255) / 256, 256, model_smem_bytes(1)>>>(dm, cnt, ds, wrap, m->hm.single_coeffs ? 1 : 0, dC, dCd, dCdd, dt_, dn, dk);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    RET(TmpBuf::back(C, dC, 2 * (size_t)cnt, mem)); RET(TmpBuf::back(Cd, dCd, 2 * (size_t)cnt, mem)); RET(TmpBuf::back(Cdd, dCdd, 2 * (size_t)cnt, mem));
    RET(TmpBuf::back(tvers, dt_, 2 * (size_t)cnt, mem)); RET(TmpBuf::back(nvers, dn, 2 * (size_t)cnt, mem)); RET(TmpBuf::back(kappa, dk, cnt, mem));
    return QSPUSH_OK;
}

int qspush_eval_dynamics(const qspush_model* m, int device, qspush_mem mem, int cnt, const double* x, const double* u,
                         double* f, double* Jx, double* Ju) {
    if (!m || !x || !u || !f || cnt < 0) return fail(QSPUSH_ERR_ARG, "bad argument");
    RET(need_device(device));
    CK(cudaSetDevice(device));
    if (cnt == 0) return QSPUSH_OK;
    const double* dm; RET(model_on_device(m, device, &dm));
    TmpBuf tb; const double *dx, *du; double *df, *dJx, *dJu;
    RET(tb.in(x, 4 * (size_t)cnt, mem, &dx)); RET(tb.in(u, 2 * (size_t)cnt, mem, &du));
    RET(tb.out(f, 4 * (size_t)cnt, mem, &df)); RET(tb.out(Jx, 16 * (size_t)cnt, mem, &dJx)); RET(tb.out(Ju, 8 * (size_t)cnt, mem, &dJu));
    k_eval_dynamics<<<(cnt + 255) / 256, 256, model_smem_bytes(1)>>>(dm, cnt, dx, du, df, dJx, dJu);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    RET(TmpBuf::back(f, df, 4 * (size_t)cnt, mem)); RET(TmpBuf::back(Jx, dJx, 16 * (size_t)cnt, mem)); RET(TmpBuf::back(Ju, dJu, 8 * (size_t)cnt, mem));
    return QSPUSH_OK;
}

int qspush_eval_erk4_sens(const qspush_model* m, int device, qspush_mem mem, int cnt, const double* x, const double* u,
                          double dt, double* Phi, double* A, double* B) {
    if (!m || !x || !u || !Phi || cnt < 0) return fail(QSPUSH_ERR_ARG, "bad argument");
    RET(need_device(device));
    CK(cudaSetDevice(device));
    if (cnt == 0) return QSPUSH_OK;
    const double* dm; RET(model_on_device(m, device, &dm));
    TmpBuf tb; const double *dx, *du; double *dP, *dA, *dB;
    RET(tb.in(x, 4 * (size_t)cnt, mem, &dx)); RET(tb.in(u, 2 * (size_t)cnt, mem, &du));
    RET(tb.out(Phi, 4 * (size_t)cnt, mem, &dP)); RET(tb.out(A, 16 * (size_t)cnt, mem, &dA)); RET(tb.out(B, 8 * (size_t)cnt, mem, &dB));
    k_eval_erk4<<<(cnt + 127) / 128, 128, model_smem_bytes(1)>>>(dm, cnt, dx, du, dt, dP, dA, dB);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    RET(TmpBuf::back(Phi, dP, 4 * (size_t)cnt, mem)); RET(TmpBuf::back(A, dA, 16 * (size_t)cnt, mem)); RET(TmpBuf::back(B, dB, 8 * (size_t)cnt, mem));
    return QSPUSH_OK;
}

void qspush_ctrl_default(qspush_ctrl* c) {
    if (!c) return;
    c->v_alpha = 1.0; c->d_v_bound = 0.0; c->t_angle0 = 3.0;   // NMPC_controller.m:98-100
    c->u_t_ub = 0.05; c->u_n_lb = 0.0;                        // :24-25
}

int qspush_eval_v_bound(const qspush_model* m, int device, qspush_mem mem, int cnt, const double* s,
                        const qspush_ctrl* ctrl, int single_quirk, double* v_bound, double* t_angle) {
    if (!m || !s || cnt < 0) return fail(QSPUSH_ERR_ARG, "bad argument");
    RET(need_device(device));
    CK(cudaSetDevice(device));
    if (cnt == 0) return QSPUSH_OK;
    qspush_ctrl c; if (ctrl) c = *ctrl; else qspush_ctrl_default(&c);
    CtrlDev cd{c.v_alpha, c.d_v_bound, c.t_angle0, c.u_t_ub, c.u_n_lb, single_quirk};
    const double* dm; RET(model_on_device(m, device, &dm));
    TmpBuf tb; const double* ds; double *dv, *da;
    RET(tb.in(s, cnt, mem, &ds)); RET(tb.out(v_bound, cnt, mem, &dv)); RET(tb.out(t_angle, cnt, mem, &da));
    k_eval_vbound<<<(cnt + 255) / 256, 256, model_smem_bytes(1)>>>(dm, cnt, ds, cd, dv, da);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    RET(TmpBuf::back(v_bound, dv, cnt, mem)); RET(TmpBuf::back(t_angle, da, cnt, mem));
    return QSPUSH_OK;
}

// ------------------------------------------------------------------------------------------------ solver
void qspush_opts_default(qspush_opts* o) {
    if (!o) return;
    std::memset(o, 0, sizeof *o);
    o->mode = QSPUSH_MODE_RTI;
    o->max_sqp_iter = 30;                                      // NMPC_controller.m:276
    o->tol_stat = o->tol_eq = o->tol_ineq = o->tol_comp = 1e-6;
    o->qp_max_iter = 50;
    o->qp_tol = 1e-11; o->qp_mu0 = 0.1; o->qp_thr = 1e-3; o->qp_tau = 0.9995;   // mu0 tuned on config 3: K_ipm 12.3 -> 11.0
    o->qp_tol_comp = 1e-18; o->qp_t_min = 1e-12; o->qp_gamma_f = 0.01; o->qp_stall = 10;   // end game: DESIGN.md 2.1
    o->globalization = 1;                                      // merit_backtracking, :272
    o->alpha_min = 0.05; o->alpha_reduction = 0.7; o->eps_sufficient_descent = 1e-4;
    o->matlab_single_quirk = 1;
    o->problems_per_warp = 0;
    o->qp_kernel = 2;
    o->h_variant = 0;
}

// default bounds of the selected constraint set (NMPC_controller.m:251-252 with :23-26, 83-84; variant: :247-248)
static void reset_bounds(qspush_solver* s) {
    SolverDev& D = s->dev;
    if (s->opts.h_variant) {
        const double lh[3] = {s->ctrl.u_n_lb, -2.0 * s->ctrl.u_t_ub, 0.0}, uh[3] = {0.03, 0.0, 2.0 * s->ctrl.u_t_ub};
        for (int i = 0; i < 3; ++i) { D.lh[i] = lh[i]; D.uh[i] = uh[i]; }
    } else {
        const double lh[3] = {-0.06, 0.0, -0.05}, uh[3] = {0.011, 0.03, 0.05};
        for (int i = 0; i < 3; ++i) { D.lh[i] = lh[i]; D.uh[i] = uh[i]; }
    }
}
// constraint-set selection and the v_bound parameters travel with every launch
static void apply_variant(const qspush_solver* s, SolverDev& D) {
    D.h_variant = s->opts.h_variant ? 1 : 0;
    D.vbp[0] = s->ctrl.v_alpha; D.vbp[1] = s->ctrl.d_v_bound; D.vbp[2] = s->ctrl.t_angle0; D.vbp[3] = s->ctrl.u_t_ub;
}

static size_t al(size_t doubles) { return (doubles + 31) / 32 * 32; }   // 256-byte granules

int qspush_solver_create(const qspush_model* const* models, int nmodels, int N, double dt, int batch,
                         int device, const qspush_opts* opts, qspush_solver** out) {
    if (!models || !out || nmodels < 1 || nmodels > MAX_MODELS) return fail(QSPUSH_ERR_ARG, "need 1..8 models");
    if (N < 1 || N > 4096 || batch < 1 || !(dt > 0.0)) return fail(QSPUSH_ERR_ARG, "bad N / batch / dt");
    for (int i = 0; i < nmodels; ++i) if (!models[i]) return fail(QSPUSH_ERR_ARG, "NULL model");
    RET(need_device(device));
    CK(cudaSetDevice(device));
    qspush_solver* s = new qspush_solver();
    s->device = device; s->B = batch; s->Bp = (batch + 31) / 32 * 32; s->N = N; s->nmodels = nmodels; s->dt = dt;
    if (opts) s->opts = *opts; else qspush_opts_default(&s->opts);
    qspush_ctrl_default(&s->ctrl);
    CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    for (auto& e : s->ev) CK(cudaEventCreate(&e));
    for (auto& e : s->evs) CK(cudaEventCreate(&e));
    CK(cudaMallocHost(&s->h_ndone, sizeof(int)));
    const size_t Bp = (size_t)s->Bp;
    // rows of every slab
    struct Slab { double** p; size_t rows; };
    SolverDev& D = s->dev;
    std::memset(&D, 0, sizeof D);
    std::vector<Slab> slabs = {
        {&D.x, (size_t)(N + 1) * 4}, {&D.u, (size_t)N * 2}, {&D.pi, (size_t)N * 4}, {&D.lam, (size_t)N * 6},
        {&D.x0, 4}, {&D.yref, (size_t)N * 6}, {&D.yref_e, 4},
        {&D.A, (size_t)N * 8}, {&D.Bm, (size_t)N * 8}, {&D.b, (size_t)N * 4}, {&D.g, (size_t)N * 6}, {&D.qN, 4}, {&D.dx0, 4},
        {&D.hv, (size_t)N * 4},
        {&D.z, (size_t)(N + 1) * 6}, {&D.zp, (size_t)(N + 1) * 6}, {&D.zc, (size_t)N * 3}, {&D.lamq, (size_t)N * 6}, {&D.t, (size_t)N * 6},
        {&D.K, (size_t)N * 8}, {&D.Li, (size_t)N * 3}, {&D.Pb, (size_t)N * 4}, {&D.kff, (size_t)N * 2}, {&D.piq, (size_t)N * 4},
        {&D.rg, (size_t)(N + 1) * 6}, {&D.rb, (size_t)N * 4}, {&D.rgs, (size_t)N},
        {&D.cost, 1}, {&D.res, 4}, {&D.alpha, 1},
        {&D.wpi, (size_t)N * 4}, {&D.wlam, (size_t)N * 6}, {&D.wx0, 4},
    };
    size_t total = 0;
    for (auto& sl : slabs) total += al(sl.rows * Bp);
    const size_t n_const = al((size_t)N * 36) + al(16) + al((size_t)N * 21) + al(10) + al((size_t)nmodels * MODEL_DOUBLES);
    s->stage_doubles = (size_t)(N + 1) * 6 * Bp;                 // largest AoS field (x: (N+1)*4, lam: N*6)
    const size_t n_int = 10 * Bp + 32;
    s->arena_bytes = (total + n_const + al(s->stage_doubles)) * 8 + (n_int + Bp) * sizeof(int) + 1024;
    CK(cudaMalloc(&s->arena, s->arena_bytes));
    CK(cudaMemsetAsync(s->arena, 0, s->arena_bytes, s->stream));
    double* cur = (double*)s->arena;
    for (auto& sl : slabs) { *sl.p = cur; cur += al(sl.rows * Bp); }
    s->d_Wdt = cur; cur += al((size_t)N * 36);
    s->d_We = cur; cur += al(16);
    s->d_H = cur; cur += al((size_t)N * 21);
    s->d_QN = cur; cur += al(10);
    s->d_models = cur; cur += al((size_t)nmodels * MODEL_DOUBLES);
    s->d_stage = cur; cur += al(s->stage_doubles);
    int* icur = (int*)cur;
    int** ints[] = {&D.status, &D.sqp_iter, &D.qp_iter, &D.cold, &D.done, &D.qpstat};
    for (int** ip : ints) { *ip = icur; icur += Bp; }
    int* objid = icur; icur += Bp;
    s->d_istage = icur; icur += Bp;
    s->d_order = icur; icur += Bp;
    D.qp_last = icur; icur += Bp;
    D.ndone = icur; icur += 32;
    D.objid = objid;
    D.B = batch; D.Bp = s->Bp; D.N = N; D.nmodels = nmodels; D.dt = dt;
    D.models = s->d_models; D.Wdt = s->d_Wdt; D.We = s->d_We; D.H = s->d_H; D.QN = s->d_QN;
    for (int i = 0; i < nmodels; ++i)
        CK(cudaMemcpyAsync(s->d_models + (size_t)i * MODEL_DOUBLES, models[i]->hm.blob.data(), MODEL_DOUBLES * 8,
                           cudaMemcpyHostToDevice, s->stream));
    k_fill_int<<<(s->Bp + 255) / 256, 256, 0, s->stream>>>(D.cold, s->Bp, 1);
    s->launches++;
    // defaults of NMPC_controller.m:16-18 and :251-252 (with :23-26, 83-84)
    s->W.assign((size_t)N * 36, 0.0); s->We.assign(16, 0.0);
    const double wx[4] = {1.0, 1.0, 1e-3, 0.0}, wu[2] = {1e-3, 1e-3}, we[4] = {2e5, 2e5, 20.0, 0.0};
    for (int k = 0; k < N; ++k) {
        for (int i = 0; i < 4; ++i) s->W[(size_t)k * 36 + 7 * i] = wx[i];
        for (int i = 0; i < 2; ++i) s->W[(size_t)k * 36 + 7 * (4 + i)] = wu[i];
    }
    for (int i = 0; i < 4; ++i) s->We[5 * i] = we[i];
    reset_bounds(s);
    s->cost_dirty = true;
    // dynamic shared memory for the model tables
    const int smem = (int)model_smem_bytes(nmodels);
    CK(cudaFuncSetAttribute(k_prepare, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_linearise, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_linesearch, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_plant_step, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaStreamSynchronize(s->stream));
    *out = s;
    return QSPUSH_OK;
}

void qspush_solver_free(qspush_solver* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (auto& e : s->ev) if (e) cudaEventDestroy(e);
    for (auto& e : s->evs) if (e) cudaEventDestroy(e);
    if (s->arena) cudaFree(s->arena);
    if (s->d_ref_traj) cudaFree(s->d_ref_traj);
    if (s->d_ref_off) cudaFree(s->d_ref_off);
    if (s->h_ndone) cudaFreeHost(s->h_ndone);
    drop_step_graphs(s);
    if (s->d_guess) cudaFree(s->d_guess);
    for (double* r : s->d_ring) if (r) cudaFree(r);
    if (s->d_idx) cudaFree(s->d_idx);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int qspush_solver_set_opts(qspush_solver* s, const qspush_opts* o) {
    if (!s || !o) return fail(QSPUSH_ERR_ARG, "NULL argument");
    const bool changed = (o->h_variant != 0) != (s->opts.h_variant != 0);
    s->opts = *o;
    drop_step_graphs(s);
    if (changed) reset_bounds(s);
    return QSPUSH_OK;
}
int qspush_solver_set_ctrl(qspush_solver* s, const qspush_ctrl* c) {
    if (!s || !c) return fail(QSPUSH_ERR_ARG, "NULL argument");
    s->ctrl = *c; drop_step_graphs(s); return QSPUSH_OK;
}
}  // extern "C"

// upload dt*W, W_e and the packed z-order Hessians
static int flush_cost(qspush_solver* s) {
    if (!s->cost_dirty) return QSPUSH_OK;
    const int N = s->N;
    std::vector<double> Wdt((size_t)N * 36), H((size_t)N * 21), QN(10);
    auto perm = [](int zi) { return zi < 2 ? 4 + zi : zi - 2; };   // z index -> y index
    for (int k = 0; k < N; ++k) {
        for (int i = 0; i < 36; ++i) Wdt[(size_t)k * 36 + i] = s->dt * s->W[(size_t)k * 36 + i];
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j <= i; ++j) {
                // symmetrise: the reference passes symmetric W; average guards against round-off asymmetry
                const double a = s->W[(size_t)k * 36 + perm(i) + 6 * perm(j)], b = s->W[(size_t)k * 36 + perm(j) + 6 * perm(i)];
                H[(size_t)k * 21 + LT(i, j)] = s->dt * 0.5 * (a + b);
            }
    }
    for (int i = 0; i < 4; ++i) for (int j = 0; j <= i; ++j) QN[LT(i, j)] = 0.5 * (s->We[i + 4 * j] + s->We[j + 4 * i]);
    CK(cudaMemcpyAsync(s->d_Wdt, Wdt.data(), Wdt.size() * 8, cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemcpyAsync(s->d_H, H.data(), H.size() * 8, cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemcpyAsync(s->d_QN, QN.data(), QN.size() * 8, cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemcpyAsync(s->d_We, s->We.data(), 16 * 8, cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));   // the host vectors die here
    s->cost_dirty = false;
    return QSPUSH_OK;
}

struct FieldInfo { double* base; int dim; int nst; bool settable; };
static bool field_info(qspush_solver* s, qspush_field f, FieldInfo& fi) {
    SolverDev& D = s->dev; const int N = s->N;
    switch (f) {
        case QSPUSH_X0: fi = {D.x0, 4, 1, true}; return true;
        case QSPUSH_YREF: fi = {D.yref, 6, N, true}; return true;
        case QSPUSH_YREF_E: fi = {D.yref_e, 4, 1, true}; return true;
        case QSPUSH_X: fi = {D.x, 4, N + 1, true}; return true;
        case QSPUSH_U: fi = {D.u, 2, N, true}; return true;
        case QSPUSH_PI: fi = {D.pi, 4, N, true}; return true;
        case QSPUSH_LAM: fi = {D.lam, 6, N, true}; return true;
        case QSPUSH_COST: fi = {D.cost, 1, 1, false}; return true;
        case QSPUSH_RES: fi = {D.res, 4, 1, false}; return true;
        default: return false;
    }
}

extern "C" {

int qspush_set(qspush_solver* s, qspush_field f, int stage, int lo, int hi, const double* data, qspush_mem mem) {
    if (!s || !data) return fail(QSPUSH_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(s->device));
    const int N = s->N;
    if (f == QSPUSH_W) {
        if (mem != QSPUSH_MEM_HOST) return fail(QSPUSH_ERR_ARG, "cost_W must be passed from host memory");
        if (stage == N) std::memcpy(s->We.data(), data, 16 * 8);                       // NMPC_controller.m:154
        else if (stage >= 0 && stage < N) std::memcpy(&s->W[(size_t)stage * 36], data, 36 * 8);   // :157
        else if (stage == -1) for (int k = 0; k < N; ++k) std::memcpy(&s->W[(size_t)k * 36], data, 36 * 8);
        else return fail(QSPUSH_ERR_ARG, "cost_W: stage out of range");
        s->cost_dirty = true;
        drop_step_graphs(s);
        return QSPUSH_OK;
    }
    if (f == QSPUSH_LH || f == QSPUSH_UH) {
        if (mem != QSPUSH_MEM_HOST) return fail(QSPUSH_ERR_ARG, "bounds must be passed from host memory");
        for (int i = 0; i < 3; ++i) (f == QSPUSH_LH ? s->dev.lh : s->dev.uh)[i] = data[i];
        drop_step_graphs(s);                                    // the bounds travel with the kernel arguments
        return QSPUSH_OK;
    }
    FieldInfo fi;
    if (!field_info(s, f, fi) || !fi.settable) return fail(QSPUSH_ERR_ARG, "field cannot be set");
    if (lo < 0 || hi > s->B || lo >= hi) return fail(QSPUSH_ERR_ARG, "batch range out of bounds");
    if (fi.nst == 1) stage = -1;          // single-stage fields: the reference passes set('cost_y_ref_e', y, Hp) (NMPC_controller.m:348)
    if (stage < -1 || stage >= fi.nst) return fail(QSPUSH_ERR_ARG, "stage out of range");
    const int nb = hi - lo;
    const int R = (stage < 0) ? fi.nst * fi.dim : fi.dim;
    const int row0 = (stage < 0) ? 0 : stage * fi.dim;
    const double* src = data;
    if (mem == QSPUSH_MEM_HOST) {
        CK(cudaMemcpyAsync(s->d_stage, data, (size_t)nb * R * 8, cudaMemcpyHostToDevice, s->stream));
        src = s->d_stage;
    }
    dim3 grid((nb + 31) / 32, (R + 31) / 32), block(32, 8);
    k_aos_to_soa<<<grid, block, 0, s->stream>>>(src, fi.base, nb, R, row0, lo, s->Bp);
    CK(cudaGetLastError());
    s->launches++;
    return QSPUSH_OK;
}

int qspush_get(qspush_solver* s, qspush_field f, int stage, int lo, int hi, double* out, qspush_mem mem) {
    if (!s || !out) return fail(QSPUSH_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(s->device));
    const int N = s->N;
    if (f == QSPUSH_W) {
        if (stage == N) std::memcpy(out, s->We.data(), 16 * 8);
        else if (stage >= 0 && stage < N) std::memcpy(out, &s->W[(size_t)stage * 36], 36 * 8);
        else return fail(QSPUSH_ERR_ARG, "cost_W: stage out of range");
        return QSPUSH_OK;
    }
    if (f == QSPUSH_LH || f == QSPUSH_UH) { for (int i = 0; i < 3; ++i) out[i] = (f == QSPUSH_LH ? s->dev.lh : s->dev.uh)[i]; return QSPUSH_OK; }
    FieldInfo fi;
    if (!field_info(s, f, fi)) return fail(QSPUSH_ERR_ARG, "unknown field");
    if (lo < 0 || hi > s->B || lo >= hi) return fail(QSPUSH_ERR_ARG, "batch range out of bounds");
    if (fi.nst == 1) stage = -1;          // single-stage fields: the reference passes set('cost_y_ref_e', y, Hp) (NMPC_controller.m:348)
    if (stage < -1 || stage >= fi.nst) return fail(QSPUSH_ERR_ARG, "stage out of range");
    const int nb = hi - lo;
    const int R = (stage < 0) ? fi.nst * fi.dim : fi.dim;
    const int row0 = (stage < 0) ? 0 : stage * fi.dim;
    double* dst = (mem == QSPUSH_MEM_HOST) ? s->d_stage : out;
    dim3 grid((nb + 31) / 32, (R + 31) / 32), block(32, 8);
    k_soa_to_aos<<<grid, block, 0, s->stream>>>(fi.base, dst, nb, R, row0, lo, s->Bp);
    CK(cudaGetLastError());
    s->launches++;
    if (mem == QSPUSH_MEM_HOST) {
        CK(cudaMemcpyAsync(out, s->d_stage, (size_t)nb * R * 8, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
    }
    return QSPUSH_OK;
}

static int* int_field(qspush_solver* s, qspush_field f, bool& settable) {
    settable = false;
    switch (f) {
        case QSPUSH_STATUS: return s->dev.status;
        case QSPUSH_SQP_ITER: return s->dev.sqp_iter;
        case QSPUSH_QP_ITER: return s->dev.qp_iter;
        case QSPUSH_OBJECT_ID: settable = true; return const_cast<int*>(s->dev.objid);
        case QSPUSH_COLD: settable = true; return s->dev.cold;
        default: return nullptr;
    }
}

int qspush_set_int(qspush_solver* s, qspush_field f, int lo, int hi, const int* data, qspush_mem mem) {
    if (!s || !data) return fail(QSPUSH_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(s->device));
    bool settable; int* p = int_field(s, f, settable);
    if (!p || !settable) return fail(QSPUSH_ERR_ARG, "int field cannot be set");
    if (lo < 0 || hi > s->B || lo >= hi) return fail(QSPUSH_ERR_ARG, "batch range out of bounds");
    if (f == QSPUSH_OBJECT_ID && mem == QSPUSH_MEM_HOST)
        for (int i = 0; i < hi - lo; ++i) if (data[i] < 0 || data[i] >= s->nmodels) return fail(QSPUSH_ERR_ARG, "object id out of range");
    CK(cudaMemcpyAsync(p + lo, data, (size_t)(hi - lo) * sizeof(int),
                       mem == QSPUSH_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s->stream));
    if (mem == QSPUSH_MEM_HOST) CK(cudaStreamSynchronize(s->stream));
    return QSPUSH_OK;
}
int qspush_get_int(qspush_solver* s, qspush_field f, int lo, int hi, int* out, qspush_mem mem) {
    if (!s || !out) return fail(QSPUSH_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(s->device));
    bool settable; int* p = int_field(s, f, settable);
    if (!p) return fail(QSPUSH_ERR_ARG, "unknown int field");
    if (lo < 0 || hi > s->B || lo >= hi) return fail(QSPUSH_ERR_ARG, "batch range out of bounds");
    CK(cudaMemcpyAsync(out, p + lo, (size_t)(hi - lo) * sizeof(int),
                       mem == QSPUSH_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, s->stream));
    if (mem == QSPUSH_MEM_HOST) CK(cudaStreamSynchronize(s->stream));
    return QSPUSH_OK;
}

static CtrlDev ctrl_dev(const qspush_solver* s) {
    return CtrlDev{s->ctrl.v_alpha, s->ctrl.d_v_bound, s->ctrl.t_angle0, s->ctrl.u_t_ub, s->ctrl.u_n_lb, s->opts.matlab_single_quirk};
}

int qspush_prepare(qspush_solver* s) {
    if (!s) return fail(QSPUSH_ERR_ARG, "NULL solver");
    CK(cudaSetDevice(s->device));
    CK(rec_event(s, 4));
    k_prepare<<<(s->B + PREP_PROBLEMS - 1) / PREP_PROBLEMS, 2 * PREP_PROBLEMS, model_smem_bytes(s->nmodels), s->stream>>>(s->dev, ctrl_dev(s), LoopDev{}, nullptr, nullptr);
    CK(cudaGetLastError());
    CK(rec_event(s, 5));
    s->launches++;
    return QSPUSH_OK;
}

// launch the QP kernel selected by opts.qp_kernel: 1 (default when the horizon fits) = warp kernel (one or two problems per warp),
// parallel-in-time; 0 = one problem per thread (any horizon)
static int launch_qp(qspush_solver* s, const SolverDev& D, const IpmOpts& io, int ppw, int apply, bool order_ready = false) {
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, s->device);
    // throughput mapping (two problems per warp up to N = 55) unless the batch is so small that every problem gets an SM
    // of its own: then one problem per warp with the shortest chunks gives the lowest latency (B = 1, N = 40: 0.54 vs 0.59 ms)
    QwPlan plan = qp_warp_plan(s->N);
    if (s->B <= nsm && s->N + 1 > 16) plan = QwPlan{(s->N + 1 + 31) / 32, 32};
    const int C = plan.C;
    // auto (2): measured on B200 (tools/gpu_kernel_compare.py, tools/gpu_half_sweep.py) the warp kernel
    // wins at every batch size and horizon it supports (N <= 127): N = 40: 2.2M it/s at 4096, 2.6M at 16384 vs
    // 0.39M / 1.4M for one problem per thread; N = 100: 0.97M / 1.09M vs 0.19M / 1.01M; N = 10: 6.6M / 8.3M vs 2.0M / 7.1M
    const int warp_below = INT_MAX;
    // full SQP (apply == 0): always the warp kernel — its work queue skips the problems that already finished, while
    // the thread kernel keeps mostly idle warps alive (config 5 share, 32 768 x N = 100: 0.87 s vs 2.3 s)
    const bool want_warp = s->opts.qp_kernel == 1 || (s->opts.qp_kernel == 2 && (s->B < warp_below || D.h_variant || !apply));
    // resident problems (= warps) per CTA: bounded by shared memory (one CTA per SM), by the register file
    // (8 warps of 255 registers) and by TMEM (8 blocks of 32 lanes x 256 columns)
    const int pwd = (int)((qp_warp_smem_doubles(s->N, C) + 1) / 2 * 2);
    const size_t smem_cap = 227 * 1024 - 1024;                                       // static __shared__ + reserve
    // short horizons (N <= 15: at most 16 stages): two problems per warp, one per 16-lane segment
    const int ppw_seg = 32 / plan.seg;                          // problems per warp
    int W = (int)std::min<size_t>(QW_MAX_WARPS, smem_cap / ((size_t)pwd * ppw_seg * sizeof(double)));
    if (C * qw_tm_stage(C) > 128) W = std::min(W, 4);          // more than 256 TMEM columns per warp: one warp per lane quarter
    const bool warp = want_warp && C <= 4 && W >= 2;
    if (!warp && s->opts.qp_kernel == 2 && s->opts.problems_per_warp == 0) ppw = (s->B >= 12288) ? 32 : ppw;
    if (!warp) {
        k_qp<<<(unsigned)((s->B + ppw - 1) / ppw), 32, 0, s->stream>>>(D, io, ppw, apply);
        return QSPUSH_OK;
    }
    // at least half of the SM's shared memory, so that two CTAs (both wanting all TMEM columns) never share an SM
    const size_t smem = std::max((size_t)pwd * ppw_seg * W * sizeof(double), (size_t)116 * 1024);   // (sized for W warps, Wl <= W used)
    // small batches: spread the problems over the SMs first (a warp alone on an SM runs its problem fastest)
    const int slots = (s->B + ppw_seg - 1) / ppw_seg;                                  // warps' worth of work
    const int Wl = std::max(1, std::min(W, (slots + nsm - 1) / nsm));
    const unsigned blocks = (unsigned)std::min((slots + Wl - 1) / Wl, nsm);            // persistent: one CTA per SM
    CK(cudaMemsetAsync(D.ndone + 1, 0, sizeof(int), s->stream));                       // work-queue head
    SolverDev Dq = D;
    Dq.order = nullptr;
    if (s->B > nsm * Wl * ppw_seg && !std::getenv("QSPUSH_NO_ORDER")) {                // more problems than resident slots
        if (!order_ready) {                                     // (qspush_step: k_step_out of the previous period wrote the order)
            k_qp_order<<<1, 1024, 0, s->stream>>>(D, s->d_order);
            s->launches++;
            s->order_valid = true;
        }
        Dq.order = s->d_order;
    }
    // CTA-wide vote once per IPM iteration (the warps of a CTA then share their instruction fetches; every launch did this up to
    // r02 v14) only where it still pays.  Measured on the r02 v16 kernels at 4096 / 16 384 instances, vote vs none
    // (profiles/r02_v16_lockstep_sweep.txt): 4 warps per CTA (N = 40 / 47 / 100) -1.3 ... -2.4 % without, 3 warps (N = 63) -1.1 /
    // -1.7 %, 5 warps (N = 31) -4 / -7 %, 6 warps (N = 15, 24, 25) -11 ... -12 % / -1 ... -4 %, 8 warps (N = 5 ... 18) -3 ... -4 % /
    // 0 ... +2 %; only the 7-warp CTAs (N = 20 ... 23: schedulers host 2, 2, 2, 1 warps) run faster with the vote: +0.6 ... 1.8 % /
    // +4 ... 6 % without.  Full SQP (apply == 0: per-launch work queues that skip the converged problems, IPM iteration counts from 5 to
    // the limit) keeps the vote at every horizon: without it N = 10 / 31 / 40 / 63 / 100 are 7 / 14 / 4 / 2 / 5 % slower (16 384
    // instances, tools/gpu_sqp_lockstep.py; config 5 on 8 GPUs 0.975 s against 0.930 s).  QSPUSH_LOCKSTEP=0/1 forces either
    // (development aid).
    static const char* ls_env = std::getenv("QSPUSH_LOCKSTEP");
    const int lockstep = ls_env ? (std::atoi(ls_env) != 0) : (!apply || Wl == 7);
#define QW_LAUNCH(CC, HV, SEG)                                                                                     \
    CK(cudaFuncSetAttribute(k_qp_warp<CC, HV, SEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    k_qp_warp<CC, HV, SEG><<<blocks, 32 * Wl, smem, s->stream>>>(Dq, io, apply, pwd, lockstep)
    const int hvf = D.h_variant ? 1 : 0;
    switch ((plan.seg == 16 ? 16 : plan.seg == 8 ? 32 : 0) + C * 2 + hvf) {
#if QW_PLAN8_MAX >= 0
        case 32 + 2: QW_LAUNCH(1, 0, 8); break;
        case 32 + 3: QW_LAUNCH(1, 1, 8); break;
        case 32 + 4: QW_LAUNCH(2, 0, 8); break;
        case 32 + 5: QW_LAUNCH(2, 1, 8); break;
        case 32 + 6: QW_LAUNCH(3, 0, 8); break;
        case 32 + 7: QW_LAUNCH(3, 1, 8); break;
        case 32 + 8: QW_LAUNCH(4, 0, 8); break;
        case 32 + 9: QW_LAUNCH(4, 1, 8); break;
#endif
        case 16 + 2: QW_LAUNCH(1, 0, 16); break;
        case 16 + 3: QW_LAUNCH(1, 1, 16); break;
        case 16 + 4: QW_LAUNCH(2, 0, 16); break;
        case 16 + 5: QW_LAUNCH(2, 1, 16); break;
        case 16 + 6: QW_LAUNCH(3, 0, 16); break;
        case 16 + 7: QW_LAUNCH(3, 1, 16); break;
        case 16 + 8: QW_LAUNCH(4, 0, 16); break;
        case 16 + 9: QW_LAUNCH(4, 1, 16); break;
        case 2: QW_LAUNCH(1, 0, 32); break;
        case 3: QW_LAUNCH(1, 1, 32); break;
        case 4: QW_LAUNCH(2, 0, 32); break;
        case 5: QW_LAUNCH(2, 1, 32); break;
        case 6: QW_LAUNCH(3, 0, 32); break;
        case 7: QW_LAUNCH(3, 1, 32); break;
        case 8: QW_LAUNCH(4, 0, 32); break;
        case 9: QW_LAUNCH(4, 1, 32); break;
        default: return fail(QSPUSH_ERR_ARG, "no warp QP kernel for this horizon");
    }
#undef QW_LAUNCH
    return QSPUSH_OK;
}

static int pick_ppw(const qspush_solver* s) {
    int p = s->opts.problems_per_warp;
    if (p == 32 || p == 16 || p == 8 || p == 4) return p;
    // auto: keep at least ~4 warps per SM in flight (148 SMs) before packing warps fully
    if (s->B >= 32 * 148 * 4) return 32;
    if (s->B >= 16 * 148 * 4) return 16;
    return 8;
}

static int solve_impl(qspush_solver* s, bool order_ready);
int qspush_solve(qspush_solver* s) {
    if (!s) return fail(QSPUSH_ERR_ARG, "NULL solver");
    CK(cudaSetDevice(s->device));
    RET(flush_cost(s));
    return solve_impl(s, false);
}
}  // extern "C"
static int solve_impl(qspush_solver* s, bool order_ready) {
    const qspush_opts& o = s->opts;
    IpmOpts io{o.qp_max_iter, o.qp_tol, o.qp_mu0, o.qp_thr, o.qp_tau, o.qp_tol_comp, o.qp_t_min, o.qp_gamma_f, o.qp_stall};
    const int ppw = pick_ppw(s);
    const size_t smem = model_smem_bytes(s->nmodels);
    const size_t nlin = (size_t)(s->N + 1) * s->Bp;
    const unsigned lin_blocks = (unsigned)((nlin + 127) / 128);
    CK(rec_event(s, 0));
    if (o.mode == QSPUSH_MODE_RTI) {
        s->acc_valid = false;
        SolverDev D = s->dev; D.done = nullptr;
        apply_variant(s, D);
        k_linearise<<<lin_blocks, 128, smem, s->stream>>>(D);
        CK(rec_event(s, 1));
        RET(launch_qp(s, D, io, ppw, 1, order_ready));
        CK(rec_event(s, 2));
        CK(cudaGetLastError());
        s->launches += 2;
        CK(rec_event(s, 3));
        return QSPUSH_OK;
    }
    // ---- full SQP (NMPC_controller.m:271-276): host-driven loop, per-problem convergence on the device
    SqpOpts so{o.max_sqp_iter, {o.tol_stat, o.tol_eq, o.tol_ineq, o.tol_comp}, o.globalization, o.alpha_min, o.alpha_reduction, o.eps_sufficient_descent};
    SolverDev D = s->dev;
    apply_variant(s, D);
    CK(cudaMemsetAsync(D.done, 0, s->Bp * sizeof(int), s->stream));
    CK(cudaMemsetAsync(D.qp_iter, 0, s->Bp * sizeof(int), s->stream));
    CK(cudaMemsetAsync(D.ndone, 0, sizeof(int), s->stream));
    // time_lin / time_qp_sol of the reference's stat print (helper.m:264-269) are sums over the SQP iterations: event pairs per
    // iteration, read after the host synchronisation the loop needs anyway (convergence count)
    s->acc_lin = 0.0; s->acc_qp = 0.0; s->acc_valid = true;
    bool qp_pending = false;
    for (int it = 0; it <= o.max_sqp_iter; ++it) {
        CK(cudaEventRecord(s->evs[0], s->stream));
        k_linearise<<<lin_blocks, 128, smem, s->stream>>>(D);
        k_nlp_res<<<(unsigned)((s->B + 31) / 32), dim3(32, LS_CHUNKS), 0, s->stream>>>(D, so, it);
        CK(cudaEventRecord(s->evs[1], s->stream));
        s->launches += 2;
        CK(cudaMemcpyAsync(s->h_ndone, D.ndone, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        { float ms = 0.f; if (cudaEventElapsedTime(&ms, s->evs[0], s->evs[1]) == cudaSuccess) s->acc_lin += 1e-3 * ms;
          if (qp_pending && cudaEventElapsedTime(&ms, s->evs[2], s->evs[3]) == cudaSuccess) s->acc_qp += 1e-3 * ms;
          qp_pending = false; }
        if (*s->h_ndone >= s->B || it == o.max_sqp_iter) break;
        CK(cudaEventRecord(s->evs[2], s->stream));
        RET(launch_qp(s, D, io, ppw, 0));
        k_linesearch<<<(unsigned)((s->B + 31) / 32), dim3(32, LS_CHUNKS), smem, s->stream>>>(D, so, it);
        CK(cudaEventRecord(s->evs[3], s->stream));
        qp_pending = true;
        s->launches += 2;
    }
    CK(rec_event(s, 1));
    CK(rec_event(s, 2));
    k_cost<<<(s->B + 127) / 128, 128, 0, s->stream>>>(D);
    CK(cudaGetLastError());
    s->launches++;
    CK(rec_event(s, 3));
    return QSPUSH_OK;
}
extern "C" {

// DFMA issue-rate microbenchmark: best of 5 launches, CUDA-event timed; *tflops = 2 flop per DFMA
int qspush_measure_fp64_peak(int device, double* tflops) {
    if (!tflops) return fail(QSPUSH_ERR_ARG, "NULL argument");
    RET(need_device(device));
    CK(cudaSetDevice(device));
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
    const int blocks = nsm * 8, threads = 256, iters = 4096;
    double* d = nullptr;
    CK(cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(e0));
        k_fp64_peak<<<blocks, threads>>>(d, iters, 0.999999, 1e-7);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * 16.0 * 8.0 * iters * (double)blocks * threads / ((double)ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *tflops = best;
    return QSPUSH_OK;
}

#ifdef QSPUSH_DEV_ORDER
// development aid (not part of the C-ABI): host-supplied work-queue order, kept until the solver is destroyed
int qsdev_set_order(qspush_solver* s, const int* order) {
    CK(cudaSetDevice(s->device));
    CK(cudaMemcpyAsync(s->d_order, order, (size_t)s->B * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    s->order_valid = true; s->order_frozen = true;
    drop_step_graphs(s);
    return QSPUSH_OK;
}
#endif

int qspush_snapshot_guess(qspush_solver* s) {
    if (!s) return fail(QSPUSH_ERR_ARG, "NULL solver");
    CK(cudaSetDevice(s->device));
    const size_t bytes = (size_t)s->N * 2 * s->Bp * sizeof(double);
    if (!s->d_guess) CK(cudaMalloc(&s->d_guess, bytes));
    CK(cudaMemcpyAsync(s->d_guess, s->dev.u, bytes, cudaMemcpyDeviceToDevice, s->stream));
    return QSPUSH_OK;
}

// launches per period of the RTI step graph: k_prepare, k_linearise, k_qp_warp (or k_qp), k_step_out (+ k_shift)
int qspush_step(qspush_solver* s, const double* x0, int idx, unsigned flags, double* u0, int* status, qspush_mem mem) {
    if (!s || !x0 || !u0 || !status) return fail(QSPUSH_ERR_ARG, "NULL argument");
    if (flags & ~3u) return fail(QSPUSH_ERR_ARG, "qspush_step: unknown flag");
    if (s->opts.mode != QSPUSH_MODE_RTI) return fail(QSPUSH_ERR_ARG, "qspush_step is the RTI control period; full SQP goes through qspush_prepare / qspush_solve");
    if (idx < 1) return fail(QSPUSH_ERR_ARG, "qspush_step: idx is 1-based");
    if ((flags & QSPUSH_STEP_RESTORE_GUESS) && !s->d_guess) return fail(QSPUSH_ERR_ARG, "qspush_step: no snapshot (qspush_snapshot_guess)");
    CK(cudaSetDevice(s->device));
    RET(flush_cost(s));
    const size_t B = (size_t)s->B;
    if (!s->d_idx) CK(cudaMalloc(&s->d_idx, sizeof(int)));
    double* d_x0 = s->d_stage;                                   // staging: x0 [B][4] | u0 [B][2]
    double* d_u0 = s->d_stage + 4 * (size_t)s->Bp;
    const cudaMemcpyKind in_kind = mem == QSPUSH_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    CK(cudaMemcpyAsync(d_x0, x0, B * 4 * sizeof(double), in_kind, s->stream));
    CK(cudaMemcpyAsync(s->d_idx, &idx, sizeof(int), cudaMemcpyHostToDevice, s->stream));   // pageable source: staged before the call returns
    cudaGraphExec_t& exec = s->step_exec[flags & 3u];
    if (!exec) {
        if (!s->order_valid) {                                   // a valid permutation for the first period
            k_qp_order<<<1, 1024, 0, s->stream>>>(s->dev, s->d_order);
            CK(cudaGetLastError());
            s->order_valid = true;
        }
        const size_t msm = model_smem_bytes(s->nmodels);
        cudaGraph_t graph = nullptr;
        const long long launches0 = s->launches;
        CK(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeRelaxed));
        s->capturing = true;
        int rc = QSPUSH_OK;
        do {
            if (flags & QSPUSH_STEP_RESTORE_GUESS)
                if (cudaMemcpyAsync(s->dev.u, s->d_guess, (size_t)s->N * 2 * s->Bp * sizeof(double), cudaMemcpyDeviceToDevice, s->stream) != cudaSuccess) { rc = QSPUSH_ERR_CUDA; break; }
            LoopDev L{};
            L.traj = s->d_ref_traj; L.off = s->d_ref_off; L.T = s->ref_T;
            rec_event(s, 4);
            k_prepare<<<(unsigned)((B + PREP_PROBLEMS - 1) / PREP_PROBLEMS), 2 * PREP_PROBLEMS, msm, s->stream>>>(s->dev, ctrl_dev(s), L, s->d_idx, d_x0);
            rec_event(s, 5);
            s->launches++;
            rc = solve_impl(s, true);
            if (rc != QSPUSH_OK) break;
            k_step_out<<<(unsigned)((B + 1023) / 1024 + 1), 1024, 0, s->stream>>>(s->dev, d_u0, s->d_istage, s->order_frozen ? nullptr : s->d_order);
            s->launches++;
            if (flags & QSPUSH_STEP_SHIFT) { dim3 grid((unsigned)((B + 127) / 128), 16); k_shift<<<grid, 128, 0, s->stream>>>(s->dev); s->launches++; }
        } while (0);
        s->capturing = false;
        const cudaError_t ce = cudaStreamEndCapture(s->stream, &graph);
        s->step_launches = (int)(s->launches - launches0);
        s->launches = launches0;
        if (rc != QSPUSH_OK || ce != cudaSuccess || !graph) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc != QSPUSH_OK ? rc : fail(QSPUSH_ERR_CUDA, "qspush_step: graph capture failed"); }
        const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) { exec = nullptr; return fail(QSPUSH_ERR_CUDA, cudaGetErrorString(ie)); }
    }
    CK(cudaGraphLaunch(exec, s->stream));
    s->launches += s->step_launches;
    const cudaMemcpyKind out_kind = mem == QSPUSH_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    CK(cudaMemcpyAsync(u0, d_u0, B * 2 * sizeof(double), out_kind, s->stream));
    CK(cudaMemcpyAsync(status, s->d_istage, B * sizeof(int), out_kind, s->stream));
    if (mem == QSPUSH_MEM_HOST) CK(cudaStreamSynchronize(s->stream));
    return QSPUSH_OK;
}

int qspush_shift(qspush_solver* s) {
    if (!s) return fail(QSPUSH_ERR_ARG, "NULL solver");
    CK(cudaSetDevice(s->device));
    dim3 grid((s->B + 127) / 128, 16);
    k_shift<<<grid, 128, 0, s->stream>>>(s->dev);
    CK(cudaGetLastError());
    s->launches++;
    return QSPUSH_OK;
}

int qspush_plant_step(qspush_solver* s, double* x, const double* u, qspush_mem mem) {
    if (!s || !x || !u) return fail(QSPUSH_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(s->device));
    double* dx = x; const double* du = u;
    if (mem == QSPUSH_MEM_HOST) {
        dx = s->d_stage; du = s->d_stage + (size_t)4 * s->Bp;
        CK(cudaMemcpyAsync(dx, x, (size_t)s->B * 4 * 8, cudaMemcpyHostToDevice, s->stream));
        CK(cudaMemcpyAsync((void*)du, u, (size_t)s->B * 2 * 8, cudaMemcpyHostToDevice, s->stream));
    }
    k_plant_step<<<(s->B + 127) / 128, 128, model_smem_bytes(s->nmodels), s->stream>>>(s->dev, dx, du);
    CK(cudaGetLastError());
    s->launches++;
    if (mem == QSPUSH_MEM_HOST) {
        CK(cudaMemcpyAsync(x, dx, (size_t)s->B * 4 * 8, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
    }
    return QSPUSH_OK;
}

int qspush_closed_loop(qspush_solver* s, const double* traj, int T, const double* offset, double* x, int steps,
                       const qspush_loop_opts* lo, double* log_x, double* log_u, int* log_status, qspush_mem mem) {
    if (!s || !traj || !x || !lo) return fail(QSPUSH_ERR_ARG, "NULL argument");
    if (T < 1 || steps < 1 || lo->idx0 < 1) return fail(QSPUSH_ERR_ARG, "closed loop: need T >= 1, steps >= 1, idx0 >= 1");
    if (lo->delay_plant < 0 || lo->delay_comp < 0) return fail(QSPUSH_ERR_ARG, "closed loop: delays are counted in control periods, >= 0");
    CK(cudaSetDevice(s->device));
    const size_t B = (size_t)s->B;
    // device views of the caller arrays (host memory: staged once before and copied back once after the loop)
    struct Tmp { std::vector<void*> p; ~Tmp() { for (void* q : p) cudaFree(q); } } tmp;
    auto dev_in = [&](const void* h, size_t bytes, void** d) -> int {
        if (!h) { *d = nullptr; return QSPUSH_OK; }
        if (mem == QSPUSH_MEM_DEVICE) { *d = const_cast<void*>(h); return QSPUSH_OK; }
        CK(cudaMalloc(d, bytes)); tmp.p.push_back(*d);
        CK(cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, s->stream));
        return QSPUSH_OK;
    };
    auto dev_out = [&](void* h, size_t bytes, void** d) -> int {
        if (!h) { *d = nullptr; return QSPUSH_OK; }
        if (mem == QSPUSH_MEM_DEVICE) { *d = h; return QSPUSH_OK; }
        CK(cudaMalloc(d, bytes)); tmp.p.push_back(*d);
        return QSPUSH_OK;
    };
    void *d_traj, *d_off, *d_x, *d_lx, *d_lu, *d_ls;
    RET(dev_in(traj, (size_t)T * 6 * 8, &d_traj)); RET(dev_in(offset, B * 6 * 8, &d_off)); RET(dev_in(x, B * 4 * 8, &d_x));
    RET(dev_out(log_x, (size_t)steps * B * 4 * 8, &d_lx)); RET(dev_out(log_u, (size_t)steps * B * 2 * 8, &d_lu));
    RET(dev_out(log_status, (size_t)steps * B * sizeof(int), &d_ls));
    LoopDev L;
    L.traj = (const double*)d_traj; L.off = (const double*)d_off; L.T = T;
    for (int i = 0; i < 4; ++i) L.sigma[i] = lo->noise_sigma[i];
    L.seed = lo->seed; L.t_dist = lo->t_dist; L.amp = lo->amplitude_dist; L.xwidth = lo->xwidth;
    L.single = s->opts.matlab_single_quirk;
    L.dp = lo->delay_plant; L.dc = lo->delay_comp; L.ring_plant = nullptr; L.ring_contr = nullptr;
    for (int r = 0; r < 2; ++r) {                                // u_buff_plant / u_buff_contr start as zeros (helper.m:212, NMPC_controller.m:109)
        const int d = r ? L.dc : L.dp;
        if (d == 0) continue;
        const size_t need = (size_t)d * B * 2;                   // owned by the solver: the call stays asynchronous in device-memory mode
        if (s->ring_doubles[r] < need) {
            if (s->d_ring[r]) { CK(cudaStreamSynchronize(s->stream)); cudaFree(s->d_ring[r]); s->d_ring[r] = nullptr; s->ring_doubles[r] = 0; }
            CK(cudaMalloc(&s->d_ring[r], need * sizeof(double)));
            s->ring_doubles[r] = need;
        }
        CK(cudaMemsetAsync(s->d_ring[r], 0, need * sizeof(double), s->stream));
        (r ? L.ring_contr : L.ring_plant) = s->d_ring[r];
    }
    const size_t msm = model_smem_bytes(s->nmodels);
    CK(cudaFuncSetAttribute(k_loop_state, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm));
    CK(cudaFuncSetAttribute(k_loop_post, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm));
    const unsigned pb = (unsigned)((B + 127) / 128);
    const unsigned wb = (unsigned)(((size_t)s->N * s->Bp + 255) / 256);
    for (int i = 1; i <= steps; ++i) {
        double* lx = d_lx ? (double*)d_lx + (size_t)(i - 1) * B * 4 : nullptr;
        double* lu = d_lu ? (double*)d_lu + (size_t)(i - 1) * B * 2 : nullptr;
        int* ls = d_ls ? (int*)d_ls + (size_t)(i - 1) * B : nullptr;
        k_loop_state<<<pb, 128, msm, s->stream>>>(s->dev, L, i, (double*)d_x, lx);
        k_loop_window<<<wb, 256, 0, s->stream>>>(s->dev, L, lo->idx0 + i - 1);
        CK(cudaGetLastError());
        s->launches += 2;
        RET(qspush_prepare(s));
        RET(qspush_solve(s));
        k_loop_post<<<pb, 128, msm, s->stream>>>(s->dev, L, i, (double*)d_x, lu, ls);
        CK(cudaGetLastError());
        s->launches++;
        RET(qspush_shift(s));
    }
    if (mem == QSPUSH_MEM_HOST) {
        CK(cudaMemcpyAsync(x, d_x, B * 4 * 8, cudaMemcpyDeviceToHost, s->stream));
        if (log_x) CK(cudaMemcpyAsync(log_x, d_lx, (size_t)steps * B * 4 * 8, cudaMemcpyDeviceToHost, s->stream));
        if (log_u) CK(cudaMemcpyAsync(log_u, d_lu, (size_t)steps * B * 2 * 8, cudaMemcpyDeviceToHost, s->stream));
        if (log_status) CK(cudaMemcpyAsync(log_status, d_ls, (size_t)steps * B * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
    }
    return QSPUSH_OK;
}

int qspush_set_reference_trajectory(qspush_solver* s, const double* traj, int T, const double* offset, qspush_mem mem) {
    if (!s || !traj) return fail(QSPUSH_ERR_ARG, "NULL argument");
    if (T < 1) return fail(QSPUSH_ERR_ARG, "reference trajectory: need T >= 1");
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));                       // a window kernel may still read the old columns
    if (s->d_ref_traj) { cudaFree(s->d_ref_traj); s->d_ref_traj = nullptr; }
    if (s->d_ref_off) { cudaFree(s->d_ref_off); s->d_ref_off = nullptr; }
    s->ref_T = 0;
    const cudaMemcpyKind kind = mem == QSPUSH_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    CK(cudaMalloc(&s->d_ref_traj, (size_t)T * 6 * sizeof(double)));
    CK(cudaMemcpyAsync(s->d_ref_traj, traj, (size_t)T * 6 * sizeof(double), kind, s->stream));
    if (offset) {
        CK(cudaMalloc(&s->d_ref_off, (size_t)s->B * 6 * sizeof(double)));
        CK(cudaMemcpyAsync(s->d_ref_off, offset, (size_t)s->B * 6 * sizeof(double), kind, s->stream));
    }
    CK(cudaStreamSynchronize(s->stream));                       // the caller may reuse its buffers
    s->ref_T = T;
    return QSPUSH_OK;
}

int qspush_set_reference_window(qspush_solver* s, int idx) {
    if (!s) return fail(QSPUSH_ERR_ARG, "NULL solver");
    if (!s->d_ref_traj || s->ref_T < 1) return fail(QSPUSH_ERR_ARG, "no reference trajectory set (qspush_set_reference_trajectory)");
    if (idx < 1) return fail(QSPUSH_ERR_ARG, "reference window: idx is 1-based");
    CK(cudaSetDevice(s->device));
    LoopDev L{};
    L.traj = s->d_ref_traj; L.off = s->d_ref_off; L.T = s->ref_T;
    const unsigned wb = (unsigned)(((size_t)s->N * s->Bp + 255) / 256);
    k_loop_window<<<wb, 256, 0, s->stream>>>(s->dev, L, idx);
    CK(cudaGetLastError());
    s->launches++;
    return QSPUSH_OK;
}

int qspush_sync(qspush_solver* s) {
    if (!s) return fail(QSPUSH_ERR_ARG, "NULL solver");
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    return QSPUSH_OK;
}
void* qspush_stream(qspush_solver* s) { return s ? (void*)s->stream : nullptr; }

int qspush_get_stat(qspush_solver* s, qspush_stat which, double* out) {
    if (!s || !out) return fail(QSPUSH_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    float ms = 0.f;
    cudaError_t e = cudaSuccess;
    switch (which) {
        case QSPUSH_TIME_TOT: e = cudaEventElapsedTime(&ms, s->ev[0], s->ev[3]); break;
        case QSPUSH_TIME_LIN: if (s->acc_valid) { *out = s->acc_lin; return QSPUSH_OK; } e = cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]); break;
        case QSPUSH_TIME_QP: if (s->acc_valid) { *out = s->acc_qp; return QSPUSH_OK; } e = cudaEventElapsedTime(&ms, s->ev[1], s->ev[2]); break;
        case QSPUSH_TIME_PREP: e = cudaEventElapsedTime(&ms, s->ev[4], s->ev[5]); break;
        default: return fail(QSPUSH_ERR_ARG, "unknown stat");
    }
    if (e != cudaSuccess) { cudaGetLastError(); ms = 0.f; }   // events not recorded yet
    *out = (double)ms * 1e-3;
    return QSPUSH_OK;
}
long long qspush_launch_count(const qspush_solver* s) { return s ? s->launches : 0; }

#if defined(QW_PROFILE)
// development aid (not part of include/qspush.h): read and reset the per-phase cycle counters of k_qp_warp
int qspush_dev_phase_cycles(unsigned long long* out16) {
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(out16, qw_prof, sizeof(unsigned long long) * 16));
    unsigned long long z[16] = {0};
    CK(cudaMemcpyToSymbol(qw_prof, z, sizeof z));
    return QSPUSH_OK;
}
#endif

}  // extern "C"
