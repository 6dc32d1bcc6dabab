// tests/hostsim/hostsim.cpp — TEST-ONLY host simulation of the CUDA kernels.
//
// Compiles the __host__ __device__ kernel bodies (qs_device.cuh, qs_qp.cuh, qs_solver.cuh) with
// g++ and runs them thread-by-thread on the CPU over the same structure-of-arrays slabs, so the
// kernel logic can be checked against the oracle in a container without a GPU
// (`pytest -m "not gpu"`).  It is never part of the product: libqspush.so does not contain it and
// the package never loads it.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../uclv_qs_pushing_matlab_b200/csrc/qs_model.hpp"
#include "../../uclv_qs_pushing_matlab_b200/csrc/qs_solver.cuh"

using namespace qs;

static std::string g_err;


// ------------------------------------------------------------------------------------------------
// Warp emulator: the 32 lanes of a warp run as cooperative fibers (hand-rolled x86-64 context switch);
// every warp collective (shfl, reductions) is a yield point, values are exchanged through a
// double-buffered array — the lanes advance in lockstep exactly like *_sync collectives demand.
// ------------------------------------------------------------------------------------------------
#if !defined(__x86_64__)
#error "the hostsim warp emulator needs x86-64"
#endif
extern "C" void qs_ctx_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl qs_ctx_switch
.type qs_ctx_switch,@function
qs_ctx_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size qs_ctx_switch, .-qs_ctx_switch
)");

struct WarpEmu {
    static constexpr int W = 32;
    static constexpr size_t STACK = 512 * 1024;
    double xbuf[2][W];
    double tm[W][256];                            // emulated TMEM: 512 32-bit columns per lane, private to the lane
    int count[W];
    void* sp[W];
    void* main_sp = nullptr;
    bool done[W];
    int cur = 0;
    std::vector<char> stacks;
    void (*body)(int lane, void* arg) = nullptr;
    void* arg = nullptr;
};
static thread_local WarpEmu* g_emu = nullptr;

static void emu_fiber_entry() {
    WarpEmu* e = g_emu;
    const int lane = e->cur;
    e->body(lane, e->arg);
    e->done[lane] = true;
    qs_ctx_switch(&e->sp[lane], e->main_sp);     // never resumed
    __builtin_trap();
}

static void emu_run(void (*body)(int, void*), void* arg) {
    WarpEmu e;
    e.body = body; e.arg = arg;
    e.stacks.assign(WarpEmu::W * WarpEmu::STACK, 0);
    for (int l = 0; l < WarpEmu::W; ++l) {
        e.done[l] = false; e.count[l] = 0;
        for (int i = 0; i < 256; ++i) e.tm[l][i] = std::nan("");      // uninitialised TMEM must never reach the arithmetic
        char* top = e.stacks.data() + (size_t)(l + 1) * WarpEmu::STACK;
        top = (char*)((uintptr_t)top & ~(uintptr_t)15);
        void** p = (void**)top;
        *(--p) = nullptr;                          // fake return address of the entry function
        *(--p) = (void*)&emu_fiber_entry;          // `ret` target of the first switch
        for (int i = 0; i < 6; ++i) *(--p) = nullptr;   // rbp rbx r12 r13 r14 r15
        e.sp[l] = p;
    }
    g_emu = &e;
    // lane scheduling order between two collectives: ascending by default, descending with HS_EMU_REVERSE=1.  Code that
    // is correctly synchronised gives bit-identical results under both (a poor man's racecheck for the shared-memory
    // exchanges: a missing warp sync shows up as an order dependence)
    const char* rev_env = std::getenv("HS_EMU_REVERSE");
    const bool rev = rev_env && rev_env[0] == '1';
    for (;;) {
        bool any = false;
        for (int li = 0; li < WarpEmu::W; ++li) {
            const int l = rev ? WarpEmu::W - 1 - li : li;
            if (e.done[l]) continue;
            any = true; e.cur = l;
            qs_ctx_switch(&e.main_sp, e.sp[l]);
        }
        if (!any) break;
    }
    g_emu = nullptr;
}

struct WarpCtxHost {
    int lane_;
    int lane() const { return lane_; }
    // deposit v, yield until every lane has deposited, return the buffer of this collective
    const double* exchange(double v) const {
        WarpEmu* e = g_emu;
        const int p = e->count[lane_] & 1;
        e->xbuf[p][lane_] = v;
        e->count[lane_]++;
        qs_ctx_switch(&e->sp[lane_], e->main_sp);
        return e->xbuf[p];
    }
    double shfl(double v, int src) const { return exchange(v)[src & 31]; }
    template <int SEG, class F> double butterfly(double v, F f) const {
        const double* b = exchange(v);
        double t[32], u[32];
        for (int i = 0; i < 32; ++i) t[i] = b[i];
        for (int o = SEG / 2; o > 0; o >>= 1) { for (int i = 0; i < 32; ++i) u[i] = f(t[i], t[i ^ o]); for (int i = 0; i < 32; ++i) t[i] = u[i]; }
        return t[lane_];
    }
    template <int SEG = 32> double wmax(double v) const { return butterfly<SEG>(v, [](double a, double b) { return fmax(a, b); }); }
    template <int SEG = 32> double wmin(double v) const { return butterfly<SEG>(v, [](double a, double b) { return fmin(a, b); }); }
    template <int SEG = 32> double wsum(double v) const { return butterfly<SEG>(v, [](double a, double b) { return a + b; }); }
    template <int SEG = 32> int wany(int p) const {
        const double* b = exchange(p ? 1.0 : 0.0);
        const int lo = lane_ & ~(SEG - 1);
        for (int i = lo; i < lo + SEG; ++i) if (b[i] != 0.0) return 1;
        return 0;
    }
    void sync() const { exchange(0.0); }
    bool cta_all(bool pred) const { return pred; }      // one emulated warp per CTA
    bool lockstep() const { return false; }
    // tensor-memory block of the lane (tcgen05.ld / tcgen05.st on the device; warp-collective there, so the emulator
    // makes them collectives too: a lane that skips one deadlocks the emulated warp instead of passing silently)
    template <int n> void tm_ld(int off, double* v) const { sync(); for (int i = 0; i < n; ++i) v[i] = g_emu->tm[lane_][off + i]; }
    template <int n> void tm_st(int off, const double* v) const { sync(); for (int i = 0; i < n; ++i) g_emu->tm[lane_][off + i] = v[i]; }
    void tm_st16(int off, const double* v) const { sync(); for (int i = 0; i < 16; ++i) g_emu->tm[lane_][off + i] = v[i]; }
};

struct WarpJob { const SolverDev* S; const IpmOpts* io; int b[2]; int apply; double* sm; int per_problem; int C; int seg16; int handed[32]; };
static void warp_job_body(int lane, void* arg) {
    WarpJob* j = (WarpJob*)arg;
    WarpCtxHost w{lane};
    // work queue of the emulated warp: exactly one fetch (every lane of a segment sees the same sequence b, -1)
    auto next = [j, lane](int seg) -> int { return j->handed[lane]++ == 0 ? j->b[seg] : -1; };
    const int hv = j->S->h_variant ? 1 : 0;
    if (j->seg16) {
        switch (j->C * 2 + hv) {
            case 2: qp_warp_persistent<WarpCtxHost, 1, 0, 16>(w, j->sm, j->per_problem, *j->S, *j->io, j->apply, next); break;
            case 3: qp_warp_persistent<WarpCtxHost, 1, 1, 16>(w, j->sm, j->per_problem, *j->S, *j->io, j->apply, next); break;
            case 4: qp_warp_persistent<WarpCtxHost, 2, 0, 16>(w, j->sm, j->per_problem, *j->S, *j->io, j->apply, next); break;
            case 5: qp_warp_persistent<WarpCtxHost, 2, 1, 16>(w, j->sm, j->per_problem, *j->S, *j->io, j->apply, next); break;
            case 6: qp_warp_persistent<WarpCtxHost, 3, 0, 16>(w, j->sm, j->per_problem, *j->S, *j->io, j->apply, next); break;
            case 7: qp_warp_persistent<WarpCtxHost, 3, 1, 16>(w, j->sm, j->per_problem, *j->S, *j->io, j->apply, next); break;
            case 8: qp_warp_persistent<WarpCtxHost, 4, 0, 16>(w, j->sm, j->per_problem, *j->S, *j->io, j->apply, next); break;
            default: qp_warp_persistent<WarpCtxHost, 4, 1, 16>(w, j->sm, j->per_problem, *j->S, *j->io, j->apply, next); break;
        }
        return;
    }
    switch (j->C * 2 + hv) {
        case 4: qp_warp_persistent<WarpCtxHost, 2, 0, 32>(w, j->sm, j->per_problem, *j->S, *j->io, j->apply, next); break;
        case 5: qp_warp_persistent<WarpCtxHost, 2, 1, 32>(w, j->sm, j->per_problem, *j->S, *j->io, j->apply, next); break;
        case 6: qp_warp_persistent<WarpCtxHost, 3, 0, 32>(w, j->sm, j->per_problem, *j->S, *j->io, j->apply, next); break;
        case 7: qp_warp_persistent<WarpCtxHost, 3, 1, 32>(w, j->sm, j->per_problem, *j->S, *j->io, j->apply, next); break;
        case 8: qp_warp_persistent<WarpCtxHost, 4, 0, 32>(w, j->sm, j->per_problem, *j->S, *j->io, j->apply, next); break;
        default: qp_warp_persistent<WarpCtxHost, 4, 1, 32>(w, j->sm, j->per_problem, *j->S, *j->io, j->apply, next); break;
    }
}
// mirrors k_qp_warp: one emulated warp per problem, or per pair of problems on short horizons (N <= 15; b1 = -1: odd tail)
static bool qp_warp_pairs(int N) { return qp_warp_plan(N).seg == 16; }
static void qp_warp_host(const SolverDev& S, const IpmOpts& io, int b0, int b1, int apply) {
    const int pwd = (int)((qp_warp_smem_doubles(S.N) + 1) / 2 * 2);
    std::vector<double> sm((size_t)pwd * 2, 0.0);
    WarpJob j{&S, &io, {b0, b1}, apply, sm.data(), pwd, qp_warp_chunk(S.N), qp_warp_pairs(S.N) ? 1 : 0, {0}};
    emu_run(warp_job_body, &j);
}

// test hook: suffix scan of stage elements vs nothing else — returns P_k of every stage from the scan path

extern "C" {

const char* hs_last_error() { return g_err.c_str(); }

void* hs_model_create(const double* S, int nknots, const double* P, int n, int p, double mu_sp, double c_ellipse, int single) {
    HostModel* m = new HostModel();
    g_err = model_from_tables(S, nknots, P, n, p, mu_sp, c_ellipse, single != 0, *m);
    if (!g_err.empty()) { delete m; return nullptr; }
    return m;
}
void* hs_model_from_ply(const char* path, int flip, int p, double mu_sg, double mu_sp, double mass, double tau_max) {
    HostModel* m = new HostModel();
    g_err = model_from_ply(path, flip != 0, p, mu_sg, mu_sp, mass, tau_max, *m);
    if (!g_err.empty()) { delete m; return nullptr; }
    return m;
}
void hs_model_free(void* m) { delete (HostModel*)m; }
void hs_model_info(void* m_, int* n, int* nknots, double* b, double* c_ellipse, double* mu_sp) {
    HostModel* m = (HostModel*)m_;
    *n = m->n; *nknots = (int)m->S.size(); *b = m->b; *c_ellipse = m->c_ellipse; *mu_sp = m->mu_sp;
}
void hs_model_tables(void* m_, double* S, double* P, double* c1, double* c2, double* blob) {
    HostModel* m = (HostModel*)m_;
    if (S) std::memcpy(S, m->S.data(), m->S.size() * 8);
    if (P) std::memcpy(P, m->P.data(), m->P.size() * 8);
    if (c1) std::memcpy(c1, m->c1.data(), m->c1.size() * 8);
    if (c2) std::memcpy(c2, m->c2.data(), m->c2.size() * 8);
    if (blob) std::memcpy(blob, m->blob.data(), MODEL_DOUBLES * 8);
}
int hs_model_doubles() { return MODEL_DOUBLES; }

// mirrors k_eval_spline
void hs_eval_spline(void* m_, int cnt, const double* s, int wrap, double* C, double* Cd, double* Cdd, double* tv, double* nv, double* kappa) {
    HostModel* hm = (HostModel*)m_; const double* M = hm->blob.data();
    for (int i = 0; i < cnt; ++i) {
        double sg = s[i];
        if (wrap == 1) sg = matlab_mod(sg, M[1], hm->single_coeffs);
        else if (wrap == 2) sg = wrap_dyn(sg, M[1]);
        Curve c; curve_eval(M, sg, c);
        if (C) { C[2 * i] = c.cx; C[2 * i + 1] = c.cy; }
        if (Cd) { Cd[2 * i] = c.dx; Cd[2 * i + 1] = c.dy; }
        if (Cdd) { curve_dd(M, sg, Cdd[2 * i], Cdd[2 * i + 1]); }
        if (tv || nv) {
            const double nrm = sqrt(c.dx * c.dx + c.dy * c.dy);
            const double tx = c.dx / nrm, ty = c.dy / nrm;
            if (tv) { tv[2 * i] = tx; tv[2 * i + 1] = ty; }
            if (nv) { nv[2 * i] = ty; nv[2 * i + 1] = -tx; }
        }
        if (kappa) kappa[i] = (c.dx * c.hy - c.dy * c.hx) / (c.dx * c.dx + c.dy * c.dy);
    }
}
// mirrors k_eval_dynamics
void hs_eval_dynamics(void* m_, int cnt, const double* x, const double* u, double* f, double* Jx, double* Ju) {
    const double* M = ((HostModel*)m_)->blob.data();
    for (int i = 0; i < cnt; ++i) {
        Dyn d;
        dyn_eval<true>(M, x[4 * i + 2], x[4 * i + 3], u[2 * i], u[2 * i + 1], d);
        for (int r = 0; r < 4; ++r) f[4 * i + r] = d.f[r];
        if (Jx) {
            const double jt[4] = {-d.f[1], d.f[0], 0.0, 0.0};
            for (int r = 0; r < 4; ++r) { Jx[16 * i + 4 * r] = 0; Jx[16 * i + 4 * r + 1] = 0; Jx[16 * i + 4 * r + 2] = jt[r]; Jx[16 * i + 4 * r + 3] = d.fs[r]; }
        }
        if (Ju) for (int r = 0; r < 4; ++r) { Ju[8 * i + 2 * r] = d.fun[r]; Ju[8 * i + 2 * r + 1] = d.fut[r]; }
    }
}
// mirrors k_eval_erk4
void hs_eval_erk4(void* m_, int cnt, const double* x, const double* u, double dt, double* Phi, double* A, double* B) {
    const double* M = ((HostModel*)m_)->blob.data();
    for (int i = 0; i < cnt; ++i) {
        double Sm[16];
        erk4_sens(M, x + 4 * i, u[2 * i], u[2 * i + 1], dt, Phi + 4 * i, Sm);
        for (int r = 0; r < 4; ++r) {
            A[16 * i + 4 * r] = r == 0; A[16 * i + 4 * r + 1] = r == 1;
            A[16 * i + 4 * r + 2] = Sm[4 * r]; A[16 * i + 4 * r + 3] = Sm[4 * r + 1];
            B[8 * i + 2 * r] = Sm[4 * r + 2]; B[8 * i + 2 * r + 1] = Sm[4 * r + 3];
        }
    }
}
void hs_eval_vbound(void* m_, int cnt, const double* s, const double* ctrl5, int single, double* vb, double* ta) {
    const double* M = ((HostModel*)m_)->blob.data();
    for (int i = 0; i < cnt; ++i) vb[i] = v_bound_of(M, s[i], ctrl5[0], ctrl5[1], ctrl5[2], ctrl5[3], single != 0, ta ? ta + i : nullptr);
}

// Whole solver pipeline in kernel order on host slabs.
//   opts_d: [qp_tol, qp_mu0, qp_thr, qp_tau, tol_stat, tol_eq, tol_ineq, tol_comp, alpha_min, alpha_red, eps_sd, qp_tol_comp, qp_t_min, qp_gamma_f, qp_stall]
//   opts_i: [mode(0 rti,1 sqp,2 qp-only), qp_max_iter, max_sqp_iter, globalization, single_quirk, do_prepare, do_shift, qp_kernel(0 thread,1 warp), h_variant]
//   ctrl5 : [v_alpha, d_v_bound, t_angle0, u_t_ub, u_n_lb]
// AoS in/out: x0 [nb][4] (in/out: wrapped by prepare), yref [nb][N][6], yref_e [nb][4], x [nb][N+1][4], u [nb][N][2],
//             pi [nb][N][4], lam [nb][N][6], cold [nb]
// outputs : stats_i [nb][3] (status, sqp_iter, qp_iter), stats_d [nb][6] (cost, res[4], alpha),
//           step_z [nb][N+1][6] (QP solution du,dx), qp_pi [nb][N][4], qp_lam [nb][N][6], lin_* optional (A [nb][N][8] ...)
struct HostRun {
    int nb, Bp, N;
    std::vector<double> blobs;
    SolverDev S;
    std::vector<int> v_obj, v_status, v_sqp, v_qpit, v_cold, v_done, v_qpstat, v_ndone;
    std::vector<double> sx, su, spi, slam, sx0, syr, sye, sA, sB, sb, sg, sqN, sdx0, shv, sz, szp, szc, slq, st, sK, sLi, sPb, skf, spq,
        srg, srb, srgs, scost, sres, salpha, swpi, swlam, swx0, Wdt, H, QN, Wev;
    IpmOpts io; SqpOpts so; CtrlDev cp;
    int mode, use_warp, single;
    std::string err;

    HostRun(void* const* models, int nmodels, int N_, double dt, int nb_, const int* objid, const double* W, const double* We,
            const double* lh, const double* uh, const double* opts_d, const int* opts_i, const double* ctrl5, const int* cold)
        : nb(nb_), Bp((nb_ + 31) / 32 * 32), N(N_) {
        blobs.resize((size_t)nmodels * MODEL_DOUBLES);
        for (int i = 0; i < nmodels; ++i) std::memcpy(&blobs[(size_t)i * MODEL_DOUBLES], ((HostModel*)models[i])->blob.data(), MODEL_DOUBLES * 8);
        std::memset(&S, 0, sizeof S);
        S.B = nb; S.Bp = Bp; S.N = N; S.nmodels = nmodels; S.dt = dt; S.models = blobs.data();
        v_obj.assign(Bp, 0); v_status.assign(Bp, 0); v_sqp.assign(Bp, 0); v_qpit.assign(Bp, 0); v_cold.assign(Bp, 0); v_done.assign(Bp, 0);
        v_qpstat.assign(Bp, 0); v_ndone.assign(32, 0);
        for (int b = 0; b < nb; ++b) { v_obj[b] = objid ? objid[b] : 0; v_cold[b] = cold ? cold[b] : 0; }
        S.objid = v_obj.data(); S.status = v_status.data(); S.sqp_iter = v_sqp.data(); S.qp_iter = v_qpit.data();
        S.cold = v_cold.data(); S.done = v_done.data(); S.qpstat = v_qpstat.data(); S.ndone = v_ndone.data();
        auto slab = [&](std::vector<double>& v, size_t rows) { v.assign(rows * Bp, 0.0); return v.data(); };
        S.x = slab(sx, (N + 1) * 4); S.u = slab(su, N * 2); S.pi = slab(spi, N * 4); S.lam = slab(slam, N * 6);
        S.x0 = slab(sx0, 4); S.yref = slab(syr, N * 6); S.yref_e = slab(sye, 4);
        S.A = slab(sA, N * 8); S.Bm = slab(sB, N * 8); S.b = slab(sb, N * 4); S.g = slab(sg, N * 6); S.qN = slab(sqN, 4); S.dx0 = slab(sdx0, 4);
        S.hv = slab(shv, N * 4);
        S.z = slab(sz, (N + 1) * 6); S.zp = slab(szp, (N + 1) * 6); S.zc = slab(szc, N * 3); S.lamq = slab(slq, N * 6); S.t = slab(st, N * 6);
        S.K = slab(sK, N * 8); S.Li = slab(sLi, N * 3); S.Pb = slab(sPb, N * 4); S.kff = slab(skf, N * 2); S.piq = slab(spq, N * 4);
        S.rg = slab(srg, (N + 1) * 6); S.rb = slab(srb, N * 4); S.rgs = slab(srgs, N);
        S.cost = slab(scost, 1); S.res = slab(sres, 4); S.alpha = slab(salpha, 1);
        S.wpi = slab(swpi, N * 4); S.wlam = slab(swlam, N * 6); S.wx0 = slab(swx0, 4);
        // cost constants exactly like flush_cost() in qspush_capi.cu
        Wdt.resize((size_t)N * 36); H.resize((size_t)N * 21); QN.resize(10); Wev.assign(We, We + 16);
        auto perm = [](int zi) { return zi < 2 ? 4 + zi : zi - 2; };
        for (int k = 0; k < N; ++k) {
            for (int i = 0; i < 36; ++i) Wdt[(size_t)k * 36 + i] = dt * W[(size_t)k * 36 + i];
            for (int i = 0; i < 6; ++i) for (int j = 0; j <= i; ++j)
                H[(size_t)k * 21 + LT(i, j)] = dt * 0.5 * (W[(size_t)k * 36 + perm(i) + 6 * perm(j)] + W[(size_t)k * 36 + perm(j) + 6 * perm(i)]);
        }
        for (int i = 0; i < 4; ++i) for (int j = 0; j <= i; ++j) QN[LT(i, j)] = 0.5 * (We[i + 4 * j] + We[j + 4 * i]);
        S.Wdt = Wdt.data(); S.We = Wev.data(); S.H = H.data(); S.QN = QN.data();
        for (int i = 0; i < 3; ++i) { S.lh[i] = lh[i]; S.uh[i] = uh[i]; }
        S.h_variant = opts_i[8];                                     // like apply_variant() in qspush_capi.cu
        for (int i = 0; i < 4; ++i) S.vbp[i] = ctrl5[i];
        mode = opts_i[0];
        io = IpmOpts{opts_i[1], opts_d[0], opts_d[1], opts_d[2], opts_d[3], opts_d[11], opts_d[12], opts_d[13], (int)opts_d[14]};
        so = SqpOpts{opts_i[2], {opts_d[4], opts_d[5], opts_d[6], opts_d[7]}, opts_i[3], opts_d[8], opts_d[9], opts_d[10]};
        cp = CtrlDev{ctrl5[0], ctrl5[1], ctrl5[2], ctrl5[3], ctrl5[4], opts_i[4]};
        single = opts_i[4];
        use_warp = opts_i[7] && qp_warp_chunk(N) <= 4;
    }
    // AoS <-> SoA (k_aos_to_soa / k_soa_to_aos)
    void in(const double* src, std::vector<double>& dst, int R) { for (int b = 0; b < nb; ++b) for (int r = 0; r < R; ++r) dst[(size_t)r * Bp + b] = src[(size_t)b * R + r]; }
    void out(const std::vector<double>& src, double* dst, int R) const { if (!dst) return; for (int b = 0; b < nb; ++b) for (int r = 0; r < R; ++r) dst[(size_t)b * R + r] = src[(size_t)r * Bp + b]; }
    const double* Mall() const { return blobs.data(); }
    void prepare() { for (int b = 0; b < nb; ++b) prepare_one(S, cp, Mall(), b); }
    // QP of every live problem (the warp emulator takes pairs of problems on short horizons, like the kernel's queue)
    void qp_all(int apply, bool skip_done) {
        std::vector<int> todo;
        for (int b = 0; b < nb; ++b) if (!(skip_done && S.done[b])) todo.push_back(b);
        if (!use_warp) { for (int b : todo) qp_one(S, io, b, apply); return; }
        if (qp_warp_pairs(N)) { for (size_t i = 0; i < todo.size(); i += 2) qp_warp_host(S, io, todo[i], i + 1 < todo.size() ? todo[i + 1] : -1, apply); }
        else for (int b : todo) qp_warp_host(S, io, b, -1, apply);
    }
    void linearise_all() { for (int k = 0; k <= N; ++k) for (int b = 0; b < nb; ++b) if (!S.done[b]) linearise_one(S, Mall(), k, b); }
    void solve() {                                                // qspush_solve
        if (mode == 0 || mode == 2) {
            linearise_all();
            qp_all(mode == 0 ? 1 : 0, false);
        } else {
            std::fill(v_done.begin(), v_done.end(), 0); std::fill(v_qpit.begin(), v_qpit.end(), 0);
            for (int it = 0; it <= so.max_iter; ++it) {
                linearise_all();
                int nd = 0;
                for (int b = 0; b < nb; ++b) { if (!S.done[b]) nlp_res_one(S, so, it, b, ChunkSerial(), true); nd += S.done[b]; }
                if (nd >= nb || it == so.max_iter) break;
                qp_all(0, true);
                for (int b = 0; b < nb; ++b) if (!S.done[b]) linesearch_one(S, so, Mall(), it, b, ChunkSerial(), true);
            }
            for (int b = 0; b < nb; ++b) cost_one(S, b);
        }
    }
    void shift() { for (int b = 0; b < nb; ++b) for (int c = 0; c < 16; ++c) shift_one(S, c, b); }
};

int hs_solve(void* const* models, int nmodels, int N, double dt, int nb, const int* objid,
             const double* W, const double* We, const double* lh, const double* uh,
             const double* opts_d, const int* opts_i, const double* ctrl5,
             double* x0, const double* yref, const double* yref_e, double* x, double* u, double* pi, double* lam, int* cold,
             int* stats_i, double* stats_d, double* step_z, double* qp_pi, double* qp_lam,
             double* lin_A, double* lin_B, double* lin_b, double* lin_g) {
    HostRun R(models, nmodels, N, dt, nb, objid, W, We, lh, uh, opts_d, opts_i, ctrl5, cold);
    if (!R.err.empty()) { g_err = R.err; return -1; }
    SolverDev& S = R.S;
    const int Bp = R.Bp;
    R.in(x0, R.sx0, 4); R.in(yref, R.syr, N * 6); R.in(yref_e, R.sye, 4); R.in(x, R.sx, (N + 1) * 4); R.in(u, R.su, N * 2);
    if (pi) R.in(pi, R.spi, N * 4);
    if (lam) R.in(lam, R.slam, N * 6);
    const int mode = R.mode;
    if (opts_i[5]) R.prepare();
    R.solve();
    R.out(R.sz, step_z, (N + 1) * 6); R.out(mode == 0 ? R.spi : R.spq, qp_pi, N * 4); R.out(mode == 0 ? R.slam : R.slq, qp_lam, N * 6);
    R.out(R.sA, lin_A, N * 8); R.out(R.sB, lin_B, N * 8); R.out(R.sb, lin_b, N * 4); R.out(R.sg, lin_g, N * 6);
    if (opts_i[6]) R.shift();
    R.out(R.sx0, x0, 4); R.out(R.sx, x, (N + 1) * 4); R.out(R.su, u, N * 2);
    if (pi) R.out(R.spi, pi, N * 4);
    if (lam) R.out(R.slam, lam, N * 6);
    for (int b = 0; b < nb; ++b) {
        if (cold) cold[b] = S.cold[b];
        if (stats_i) { stats_i[3 * b] = S.status[b]; stats_i[3 * b + 1] = S.sqp_iter[b]; stats_i[3 * b + 2] = S.qp_iter[b]; }
        if (stats_d) { stats_d[6 * b] = R.scost[b]; for (int i = 0; i < 4; ++i) stats_d[6 * b + 1 + i] = R.sres[(size_t)i * Bp + b]; stats_d[6 * b + 5] = R.salpha[b]; }
    }
    return 0;
}

// mirrors qspush_closed_loop (qspush_capi.cu): the same kernel bodies in the same order on host slabs.
//   loop_d: [sigma0..3, amplitude_dist, xwidth]   loop_i: [idx0, t_dist, delay_plant, delay_comp]   seed: noise seed
int hs_closed_loop(void* const* models, int nmodels, int N, double dt, int nb, const int* objid,
                   const double* W, const double* We, const double* lh, const double* uh,
                   const double* opts_d, const int* opts_i, const double* ctrl5,
                   const double* traj, int T, const double* offset, double* x, int steps,
                   const double* loop_d, const int* loop_i, unsigned long long seed,
                   double* log_x, double* log_u, int* log_status) {
    std::vector<int> cold(nb, 1);
    HostRun R(models, nmodels, N, dt, nb, objid, W, We, lh, uh, opts_d, opts_i, ctrl5, cold.data());
    if (!R.err.empty()) { g_err = R.err; return -1; }
    LoopDev L;
    L.traj = traj; L.off = offset; L.T = T;
    for (int i = 0; i < 4; ++i) L.sigma[i] = loop_d[i];
    L.seed = seed; L.t_dist = loop_i[1]; L.amp = loop_d[4]; L.xwidth = loop_d[5]; L.single = R.single;
    L.dp = loop_i[2]; L.dc = loop_i[3];
    std::vector<double> ring_p((size_t)L.dp * nb * 2, 0.0), ring_c((size_t)L.dc * nb * 2, 0.0);
    L.ring_plant = ring_p.data(); L.ring_contr = ring_c.data();
    for (int i = 1; i <= steps; ++i) {
        for (int b = 0; b < nb; ++b) loop_state_one(R.S, L, R.Mall(), i, x, log_x ? log_x + (size_t)(i - 1) * nb * 4 : nullptr, b);
        for (int k = 0; k < N; ++k) for (int b = 0; b < nb; ++b) loop_window_one(R.S, L, loop_i[0] + i - 1, k, b);
        R.prepare();
        R.solve();
        for (int b = 0; b < nb; ++b)
            loop_post_one(R.S, L, R.Mall(), i, x, log_u ? log_u + (size_t)(i - 1) * nb * 2 : nullptr, log_status ? log_status + (size_t)(i - 1) * nb : nullptr, b);
        R.shift();
    }
    return 0;
}

}  // extern "C"
