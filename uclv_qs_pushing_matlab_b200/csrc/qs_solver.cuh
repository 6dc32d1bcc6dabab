// qs_solver.cuh — per-problem / per-(problem, stage) bodies of the solver kernels.
//
// Each function does the work of ONE thread of the corresponding kernel in qs_kernels.cuh on the
// structure-of-arrays slabs described there.  They are __host__ __device__ only so that
// tests/hostsim can execute the identical logic on a GPU-less CI box; the product launches them
// exclusively from the sm_100a kernels.
#pragma once
#include "qs_device.cuh"
#include "qs_qp.cuh"
#include "qs_qp_warp.cuh"

namespace qs {

// ------------------------------------------------------------------------------------------------
// device-side view of a solver (passed by value to the kernels)
// ------------------------------------------------------------------------------------------------
struct SolverDev {
    int B, Bp, N, nmodels;
    double dt;
    const double* models;   // [nmodels][MODEL_DOUBLES]
    const int* objid;       // [Bp]
    // iterate and references
    double *x, *u, *pi, *lam;            // [(N+1)*4] [N*2] [N*4] [N*6]  x Bp
    double *x0, *yref, *yref_e;          // [4] [N*6] [4]
    // linearisation
    double *A, *Bm, *b, *g, *qN, *dx0;   // [N*8] [N*8] [N*4] [N*6] [4] [4]
    double *hv;                          // [N*4] constraint function h_k (3) and v_bound'(s_k) (h_variant 1; else 0)
    // QP work
    double *z, *zp, *zc, *lamq, *t, *K, *Li, *Pb, *kff, *piq, *rg, *rb, *rgs;
    // cost / constraint constants
    const double *Wdt;      // [N][36] dt*W, y = [x;u] order, column-major
    const double *We;       // [16]
    const double *H;        // [N][21] packed, z order
    const double *QN;       // [10]
    double lh[3], uh[3];
    int h_variant;          // 0: h = [s; u_n; u_t]   1: h = [u_n; u_t - v_bound(s); u_t + v_bound(s)]  (NMPC_controller.m:237/238)
    double vbp[4];          // v_alpha, d_v_bound, t_angle0, u_t_ub of v_bound(s) (:229)
    // bookkeeping
    int *status, *sqp_iter, *qp_iter, *cold, *done, *qpstat, *ndone;   // ndone[0]: SQP finished count, ndone[1]: QP work-queue head
    int *order;             // work-queue order of the warp QP kernel (nullptr: index order), see k_qp_order
    int *qp_last;           // IPM iterations of the most recent QP of every problem (the predictor behind `order`)
    double *cost, *res, *alpha;
    // SQP merit weights
    double *wpi, *wlam, *wx0;
};

struct IpmOpts { int max_iter; double tol, mu0, thr, tau, tol_cp, t_min, gamma_f; int stall; };
struct SqpOpts { int max_iter; double tol[4]; int globalization; double alpha_min, alpha_red, eps_sd; };
struct CtrlDev { double v_alpha, d_v_bound, t_angle0, u_t_ub, u_n_lb; int single; };

#define QS_EL(p, row, b) (p)[(size_t)(row) * S.Bp + (b)]


#define QS_EL(p, row, b) (p)[(size_t)(row) * S.Bp + (b)]

// K6 prepare: NMPC_controller.solve pre-processing (NMPC_controller.m:332, 351-380)
QS_HD void prepare_one(const SolverDev& S, const CtrlDev& cp, const double* __restrict__ Mall, int b) {
    const double* M = Mall + (size_t)S.objid[b] * MODEL_DOUBLES;
    const int N = S.N;
    const double bb = M[1];
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = QS_EL(S.x0, i, b);
    {   // :332  x0(4) = mod(x0(4), b) - b*(x0(4) < 0)
        double w = matlab_mod(x[3], bb, cp.single != 0);
        if (cp.single) w = (double)((float)w - (float)bb * (x[3] < 0.0 ? 1.f : 0.f));
        else w = w - bb * (x[3] < 0.0 ? 1.0 : 0.0);
        x[3] = w;
        QS_EL(S.x0, 3, b) = w;                                   // :334 constr_x0 <- wrapped x0
    }
    const bool cold = S.cold[b] != 0;
    if (cold) {                                                  // :351-355
        for (int k = 0; k < N; ++k) {
#pragma unroll
            for (int i = 0; i < 4; ++i) QS_EL(S.pi, k * 4 + i, b) = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) QS_EL(S.lam, k * 6 + i, b) = 0.0;
        }
        S.cold[b] = 0;
    }
    double vb = v_bound_of(M, x[3], cp.v_alpha, cp.d_v_bound, cp.t_angle0, cp.u_t_ub, cp.single != 0, nullptr);   // :357
    double un_nx = cold ? cp.u_n_lb : QS_EL(S.u, 0, b), ut_nx = cold ? 0.0 : QS_EL(S.u, 1, b);
    for (int k = 0; k < N; ++k) {
        double un = un_nx, ut = ut_nx;
        if (!cold && k + 1 < N) { un_nx = QS_EL(S.u, (k + 1) * 2 + 0, b); ut_nx = QS_EL(S.u, (k + 1) * 2 + 1, b); }   // next stage's loads fly during this stage's arithmetic
        if (fabs(ut) > vb) {                                     // :358-364, :375-379
            const double old = ut;
            ut = (double)((old > 0.0) - (old < 0.0)) * vb;
            un = ut * un / old;
        }
        QS_EL(S.u, k * 2 + 0, b) = un; QS_EL(S.u, k * 2 + 1, b) = ut;
#pragma unroll
        for (int i = 0; i < 4; ++i) QS_EL(S.x, k * 4 + i, b) = x[i];
        Dyn d;
        dyn_eval<false>(M, x[2], x[3], un, ut, d);               // :369-370 forward Euler
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = fma(S.dt, d.f[i], x[i]);
        if (k + 1 < N) vb = v_bound_of(M, x[3], cp.v_alpha, cp.d_v_bound, cp.t_angle0, cp.u_t_ub, cp.single != 0, nullptr);   // :371
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) QS_EL(S.x, N * 4 + i, b) = x[i];
}

// constraint function of one stage at (s, u_n, u_t); beta = v_bound'(s) (rows 1 / 2 carry -beta / +beta on ds)
QS_HD void h_eval(const SolverDev& S, const double* __restrict__ M, double s, double un, double ut, double h[3], double* beta) {
    if (S.h_variant == 0) { h[0] = s; h[1] = un; h[2] = ut; if (beta) *beta = 0.0; return; }
    const double vb = v_bound_sym(M, s, S.vbp, beta);
    h[0] = un; h[1] = ut - vb; h[2] = ut + vb;
}

// K2 + K3 linearise, one (problem, stage) pair; stage index N does the terminal terms.
//   A_k, B_k, b_k = Phi(x_k,u_k) - x_{k+1}   (ERK4 + forward sensitivities)
//   g_k = dt * W_k ([x_k;u_k] - yref_k) in z = [u;x] order ; q_N = W_e (x_N - yref_e) ; dx0 = x0 - x_0
QS_HD void linearise_one(const SolverDev& S, const double* __restrict__ Mall, int k, int b) {
    const int N = S.N;
    if (k == N) {
        double r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            r[i] = QS_EL(S.x, N * 4 + i, b) - QS_EL(S.yref_e, i, b);
            QS_EL(S.dx0, i, b) = QS_EL(S.x0, i, b) - QS_EL(S.x, i, b);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double a = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) a = fma(S.We[i + 4 * j], r[j], a);
            QS_EL(S.qN, i, b) = a;
        }
        return;
    }
    const double* M = Mall + (size_t)S.objid[b] * MODEL_DOUBLES;
    double x[4], xn[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[i] = QS_EL(S.x, k * 4 + i, b); xn[i] = QS_EL(S.x, (k + 1) * 4 + i, b); }
    const double un = QS_EL(S.u, k * 2 + 0, b), ut = QS_EL(S.u, k * 2 + 1, b);
    double yr[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) yr[i] = QS_EL(S.yref, k * 6 + i, b);
    double Phi[4], Sm[16];
    erk4_sens(M, x, un, ut, S.dt, Phi, Sm);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        QS_EL(S.A, k * 8 + i, b) = Sm[4 * i + 0];
        QS_EL(S.A, k * 8 + 4 + i, b) = Sm[4 * i + 1];
        QS_EL(S.Bm, k * 8 + i, b) = Sm[4 * i + 2];
        QS_EL(S.Bm, k * 8 + 4 + i, b) = Sm[4 * i + 3];
        QS_EL(S.b, k * 4 + i, b) = Phi[i] - xn[i];
    }
    {
        double h[3], beta;
        h_eval(S, M, x[3], un, ut, h, &beta);
#pragma unroll
        for (int c = 0; c < 3; ++c) QS_EL(S.hv, k * 4 + c, b) = h[c];
        QS_EL(S.hv, k * 4 + 3, b) = beta;
    }
    const double r[6] = {x[0] - yr[0], x[1] - yr[1], x[2] - yr[2], x[3] - yr[3], un - yr[4], ut - yr[5]};
    const double* W = S.Wdt + (size_t)k * 36;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) a = fma(W[i + 6 * j], r[j], a);
        QS_EL(S.g, k * 6 + (i < 4 ? 2 + i : i - 4), b) = a;       // y index -> z index
    }
}

// stage cost at given (x_k, u_k): 1/2 (y - y_ref)' dt W (y - y_ref), y = [x; u]
QS_HD double stage_cost_at(const SolverDev& S, int k, int b, const double x[4], const double u[2]) {
    double r[6];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = x[i] - QS_EL(S.yref, k * 6 + i, b);
#pragma unroll
    for (int i = 0; i < 2; ++i) r[4 + i] = u[i] - QS_EL(S.yref, k * 6 + 4 + i, b);
    const double* W = S.Wdt + (size_t)k * 36;
    double q = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) a = fma(W[i + 6 * j], r[j], a);
        q = fma(r[i], a, q);
    }
    return 0.5 * q;
}
// (same, with the reference already in registers)
QS_HD double stage_cost_ref(const SolverDev& S, int k, const double x[4], const double u[2], const double yref[6]) {
    double r[6];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = x[i] - yref[i];
#pragma unroll
    for (int i = 0; i < 2; ++i) r[4 + i] = u[i] - yref[4 + i];
    const double* W = S.Wdt + (size_t)k * 36;
    double q = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) a = fma(W[i + 6 * j], r[j], a);
        q = fma(r[i], a, q);
    }
    return 0.5 * q;
}
QS_HD double terminal_cost_ref(const SolverDev& S, const double x[4], const double yref_e[4]) {
    double r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = x[i] - yref_e[i];
    double q = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) a = fma(S.We[i + 4 * j], r[j], a);
        q = fma(r[i], a, q);
    }
    return 0.5 * q;
}
QS_HD double terminal_cost_at(const SolverDev& S, int b, const double x[4]) {
    double r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = x[i] - QS_EL(S.yref_e, i, b);
    double q = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) a = fma(S.We[i + 4 * j], r[j], a);
        q = fma(r[i], a, q);
    }
    return 0.5 * q;
}
QS_HD double stage_cost(const SolverDev& S, int k, int b) {
    double r[6];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = QS_EL(S.x, k * 4 + i, b) - QS_EL(S.yref, k * 6 + i, b);
#pragma unroll
    for (int i = 0; i < 2; ++i) r[4 + i] = QS_EL(S.u, k * 2 + i, b) - QS_EL(S.yref, k * 6 + 4 + i, b);
    const double* W = S.Wdt + (size_t)k * 36;
    double q = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) a = fma(W[i + 6 * j], r[j], a);
        q = fma(r[i], a, q);
    }
    return 0.5 * q;
}
QS_HD double terminal_cost(const SolverDev& S, int b) {
    double r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = QS_EL(S.x, S.N * 4 + i, b) - QS_EL(S.yref_e, i, b);
    double q = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) a = fma(S.We[i + 4 * j], r[j], a);
        q = fma(r[i], a, q);
    }
    return 0.5 * q;
}

// K4 QP (+ K5 in RTI mode: full step, multipliers replaced, cost, status) for problem b
QS_HD void qp_one(const SolverDev& S, const IpmOpts& o, int b, int apply) {
    QpConst C;
    C.N = S.N; C.H = S.H; C.QN = S.QN;
#pragma unroll
    for (int i = 0; i < 3; ++i) { C.lh[i] = S.lh[i]; C.uh[i] = S.uh[i]; }
    C.h_variant = S.h_variant;
    C.max_iter = o.max_iter; C.tol = o.tol; C.mu0 = o.mu0; C.thr = o.thr; C.tau = o.tau;
    C.tol_cp = o.tol_cp; C.t_min = o.t_min; C.gamma_f = o.gamma_f; C.stall = o.stall;
    QpView V;
    V.stride = (size_t)S.Bp;
    V.hv = S.hv + b;
    V.A = S.A + b; V.B = S.Bm + b; V.b = S.b + b; V.g = S.g + b; V.qN = S.qN + b; V.dx0 = S.dx0 + b;
    V.x = S.x + b; V.u = S.u + b;
    V.z = S.z + b; V.zp = S.zp + b; V.zc = S.zc + b; V.t = S.t + b;
    V.K = S.K + b; V.Li = S.Li + b; V.Pb = S.Pb + b; V.kff = S.kff + b;
    V.rg = S.rg + b; V.rb = S.rb + b; V.rgs = S.rgs + b;
    V.lam = (apply ? S.lam : S.lamq) + b;      // RTI: multipliers are replaced by the QP's
    V.pi = (apply ? S.pi : S.piq) + b;
    int iters, status; double res[4];
    qp_ipm(C, V, iters, status, res);
    S.qpstat[b] = status;
    if (apply) S.qp_iter[b] = iters; else S.qp_iter[b] += iters;
    if (!apply) return;
    // ---- K5 (RTI): x += dx, u += du, cost, status
    const int N = S.N;
    bool nan = false;
    double cost = 0.0;
    for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int i = 0; i < 2; ++i) { const double v = QS_EL(S.u, k * 2 + i, b) + QS_EL(S.z, k * 6 + i, b); QS_EL(S.u, k * 2 + i, b) = v; nan = nan || !(v == v); }
#pragma unroll
        for (int i = 0; i < 4; ++i) QS_EL(S.x, k * 4 + i, b) += QS_EL(S.z, k * 6 + 2 + i, b);
        cost += stage_cost(S, k, b);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) QS_EL(S.x, N * 4 + i, b) += QS_EL(S.z, N * 6 + 2 + i, b);
    cost += terminal_cost(S, b);
    S.cost[b] = cost;
#pragma unroll
    for (int i = 0; i < 4; ++i) QS_EL(S.res, i, b) = res[i];
    S.sqp_iter[b] = 1;
    S.alpha[b] = 1.0;
    S.status[b] = nan ? 1 : (status == 2 ? 4 : 0);     // QP iteration limit is tolerated
}

// K4 (+ K5 in RTI mode), warp-per-problem version, persistent: every warp of a CTA owns one problem at a
// time and pulls the next one from a work queue when it finishes (`next` returns a warp-uniform problem
// index or -1), while one CTA-wide vote per IPM iteration keeps the warps in lockstep (shared instruction
// fetches).  sm = the warp's shared-memory state (qp_warp_smem_doubles(N) doubles).
template <class Ctx, int C, int HV, int SEG, class NextFn>
QS_HD void qp_warp_persistent(const Ctx& w, double* __restrict__ sm_warp, int per_problem_doubles, const SolverDev& S, const IpmOpts& o,
                              int apply, NextFn next) {
    QpConst Qc;
    Qc.N = S.N; Qc.H = S.H; Qc.QN = S.QN;
#pragma unroll
    for (int i = 0; i < 3; ++i) { Qc.lh[i] = S.lh[i]; Qc.uh[i] = S.uh[i]; }
    Qc.h_variant = HV;
    Qc.max_iter = o.max_iter; Qc.tol = o.tol; Qc.mu0 = o.mu0; Qc.thr = o.thr; Qc.tau = o.tau;
    Qc.tol_cp = o.tol_cp; Qc.t_min = o.t_min; Qc.gamma_f = o.gamma_f; Qc.stall = o.stall;
    // SEG = 32: one problem per warp.  SEG = 16 (N <= 15): lanes 0..15 and 16..31 carry one problem each; `b` is the
    // problem of this lane's segment (-1: none), every per-problem quantity below is per segment.
    const int seg = w.lane() / SEG;
    const int lane = w.lane() & (SEG - 1);
    double* sm = sm_warp + (size_t)seg * per_problem_doubles;
    const int N = S.N;
    QpView V;
    QwState st;
    // bind problem b to this segment: slab views + linearisation and initial point into shared memory / TMEM, right after the
    // previous problem finishes (the other warps of the CTA go on with their own problems meanwhile).
    auto bind = [&](int b_) {
        const bool live = b_ >= 0;
        const int bb = live ? b_ : 0;
        V.stride = (size_t)S.Bp;
        V.A = S.A + bb; V.B = S.Bm + bb; V.b = S.b + bb; V.g = S.g + bb; V.qN = S.qN + bb; V.dx0 = S.dx0 + bb;
        V.x = S.x + bb; V.u = S.u + bb; V.hv = S.hv + bb;
        V.z = S.z + bb; V.zp = S.zp + bb; V.zc = S.zc + bb; V.t = S.t + bb;
        V.K = S.K + bb; V.Li = S.Li + bb; V.Pb = S.Pb + bb; V.kff = S.kff + bb;
        V.rg = S.rg + bb; V.rb = S.rb + bb; V.rgs = S.rgs + bb;
        V.lam = (apply ? S.lam : S.lamq) + bb;
        V.pi = (apply ? S.pi : S.piq) + bb;
        qw_init<Ctx, C, HV, SEG>(w, sm, Qc, V, st, live);
    };
    int b = next(seg);
    {
        QW_T0();
        if (w.wany(b >= 0 ? 1 : 0)) bind(b);
        QW_TICK(9);
    }
    for (;;) {
        // The warps of a CTA run independently and leave when the queue is drained.  Up to r02 v14 they met at a CTA-wide vote once
        // per IPM iteration to share instruction fetches (on the r01 kernels voting every 2nd / 4th iteration was 9 % / 16 % slower);
        // with the instruction stream halved since then the vote costs more than it gives (N = 40: 1.8 % / 2.4 % faster without it at
        // 4096 / 16 384 instances, same results bit for bit) except in the 7-warp CTAs of N = 20 ... 23 and in full-SQP mode:
        // launch_qp sets w.lockstep() there (measurements in its comment).
        QW_T0();
        const bool idle = !w.wany(b >= 0 ? 1 : 0);         // no segment of this warp has a problem
        if (w.lockstep() ? w.cta_all(idle) : idle) break;
        QW_TICK(0);
        if (idle) continue;
        const int fin = qw_iterate<Ctx, C, HV, SEG>(w, sm, Qc, st);
        if (fin == 0) continue;
        // ---- the problems of this warp are finished: fetch the next ticket FIRST (the queue atomic then overlaps the write-back
        // and the epilogue), write back, K5 epilogue, bind.  (Up to r02 v16 the lanes also issued prefetch.global.L2 for the next
        // problem's linearisation here; k_linearise has just written those lines, they are L2-resident, and without the 114 prefetch
        // instructions per lane the control period is 1 % faster: 2.91e6 -> 2.94e6 it/s.)
        const bool live = b >= 0;
        const int b_next = next(seg);
        QW_TICK(10);
        QW_TICK(11);
        qw_writeback<Ctx, C, SEG>(w, sm, Qc, V, live, apply == 0);      // RTI: the step and the slacks stay on chip (12 of 28 doubles per stage not stored)
        QW_TICK(12);
        if (live && lane == 0) {
            S.qpstat[b] = st.status;
            if (apply) S.qp_iter[b] = st.it; else S.qp_iter[b] += st.it;
            if (S.qp_last) S.qp_last[b] = st.it;
        }
        if (apply) {
            // K5 (RTI): x += dx, u += du, cost, status — each lane updates its own stages
            w.sync();
            double cost = 0.0;
            int nan = 0;
            if (live) {
                // the step comes from shared memory and the cost is evaluated on the registers that are being stored:
                // every global load of the epilogue is independent (one latency exposure), nothing is read back
                const int Lw_ = qp_warp_lanes(N, C);
                // all loads of the lane's C stages first (the stores below could alias them as far as the compiler knows)
                double xo[C][4], uo[C][2], yr[C][6];
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    const int k = lane * C + j;
#pragma unroll
                    for (int i = 0; i < 4; ++i) xo[j][i] = (k <= N) ? QS_EL(S.x, k * 4 + i, b) : 0.0;
#pragma unroll
                    for (int i = 0; i < 2; ++i) uo[j][i] = (k < N) ? QS_EL(S.u, k * 2 + i, b) : 0.0;
#pragma unroll
                    for (int i = 0; i < 6; ++i) yr[j][i] = (k < N) ? QS_EL(S.yref, k * 6 + i, b) : ((k == N && i < 4) ? QS_EL(S.yref_e, i, b) : 0.0);
                }
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    const int k = lane * C + j;
                    if (k > N) continue;
                    double xk[4], uk[2] = {0.0, 0.0};
                    if (k < N) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) { uk[i] = uo[j][i] + QW_SM(R_Z + i, j); QS_EL(S.u, k * 2 + i, b) = uk[i]; nan |= !(uk[i] == uk[i]); }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) { xk[i] = xo[j][i] + QW_SM(R_Z + 2 + i, j); QS_EL(S.x, k * 4 + i, b) = xk[i]; }
                    cost += (k < N) ? stage_cost_ref(S, k, xk, uk, yr[j]) : terminal_cost_ref(S, xk, yr[j]);
                }
            }
            cost = w.template wsum<SEG>(cost);
            nan = w.template wany<SEG>(nan);
            if (live && lane == 0) {
                S.cost[b] = cost;
                QS_EL(S.res, 0, b) = st.r_stat; QS_EL(S.res, 1, b) = st.r_eq; QS_EL(S.res, 2, b) = st.r_in; QS_EL(S.res, 3, b) = st.r_cp;
                S.sqp_iter[b] = 1;
                S.alpha[b] = 1.0;
                S.status[b] = nan ? 1 : (st.status == 2 ? 4 : 0);
            }
        }
        b = b_next;
        QW_TICK(13);
        if (w.wany(b >= 0 ? 1 : 0)) bind(b);
        QW_TICK(9);
    }
}

// NLP residuals (inf-norms) and the convergence test of one SQP iteration; returns 1 when problem b finishes
// Stage-chunk context of the SQP-level kernels (residuals, line search): the horizon of one problem is split over
// `ny` cooperating threads (chunk y handles stages y, y + ny, ...); sum() / max() combine the chunks' partial results
// and hand the total to all of them, any() is a vote over the whole thread group.  The serial version (one thread
// per problem: host simulation) is ny = 1 with identity reductions; the kernels use a CTA of 32 problems x 8 chunks
// (qs_kernels.cuh).
struct ChunkSerial {
    int y = 0, ny = 1;
    QS_HD double sum(double v) const { return v; }
    QS_HD double max(double v) const { return v; }
    QS_HD bool any(bool p) const { return p; }
};

template <class Chunk>
QS_HD int nlp_res_one(const SolverDev& S, const SqpOpts& o, int it, int b, const Chunk& ch, bool live) {
    const int N = S.N;
    double r_stat = 0.0, r_eq = 0.0, r_in = 0.0, r_cp = 0.0;
    if (live && ch.y == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) r_eq = fmax(r_eq, fabs(QS_EL(S.dx0, i, b)));
    }
    for (int k = live ? ch.y : N; k < N; k += ch.ny) {
        double g[6], pk1[4];
#pragma unroll
        for (int i = 0; i < 6; ++i) g[i] = QS_EL(S.g, k * 6 + i, b);
#pragma unroll
        for (int i = 0; i < 4; ++i) pk1[i] = QS_EL(S.pi, k * 4 + i, b);
        double a3[4], a4[4], b1[4], b2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a3[i] = QS_EL(S.A, k * 8 + i, b); a4[i] = QS_EL(S.A, k * 8 + 4 + i, b); b1[i] = QS_EL(S.Bm, k * 8 + i, b); b2[i] = QS_EL(S.Bm, k * 8 + 4 + i, b); }
        g[0] += dot4(b1, pk1); g[1] += dot4(b2, pk1);
        g[2] += pk1[0]; g[3] += pk1[1]; g[4] += dot4(a3, pk1); g[5] += dot4(a4, pk1);
        if (k > 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) g[2 + i] -= QS_EL(S.pi, (k - 1) * 4 + i, b);
        }
        const double h[3] = {QS_EL(S.hv, k * 4, b), QS_EL(S.hv, k * 4 + 1, b), QS_EL(S.hv, k * 4 + 2, b)};
        const double beta = QS_EL(S.hv, k * 4 + 3, b);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (!h_on(S.h_variant, k, c)) continue;
            const double ll = QS_EL(S.lam, k * 6 + c, b), lu = QS_EL(S.lam, k * 6 + 3 + c, b);
            g[h_pidx(S.h_variant, c)] += lu - ll;
            g[5] += h_bcoef(S.h_variant, c, beta) * (lu - ll);
            const double sl = h[c] - S.lh[c], su = S.uh[c] - h[c];
            r_in = fmax(r_in, fmax(fmax(-sl, 0.0), fmax(-su, 0.0)));
            r_cp = fmax(r_cp, fmax(fabs(ll * sl), fabs(lu * su)));
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) { if (k == 0 && i >= 2) continue; r_stat = fmax(r_stat, fabs(g[i])); }
#pragma unroll
        for (int i = 0; i < 4; ++i) r_eq = fmax(r_eq, fabs(QS_EL(S.b, k * 4 + i, b)));
    }
    if (live && ch.y == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) r_stat = fmax(r_stat, fabs(QS_EL(S.qN, i, b) - QS_EL(S.pi, (N - 1) * 4 + i, b)));
    }
    // NaN must survive the reductions (fmax drops it): carry it as +inf markers
    const bool nan_p = !(r_stat == r_stat) || !(r_eq == r_eq) || !(r_in == r_in) || !(r_cp == r_cp);
    r_stat = ch.max(r_stat); r_eq = ch.max(r_eq); r_in = ch.max(r_in); r_cp = ch.max(r_cp);
    const bool nan = ch.max(nan_p ? 1.0 : 0.0) != 0.0;
    if (!live || ch.y != 0) return 0;
    QS_EL(S.res, 0, b) = r_stat; QS_EL(S.res, 1, b) = r_eq; QS_EL(S.res, 2, b) = r_in; QS_EL(S.res, 3, b) = r_cp;
    int fin = -1;
    if (nan) fin = 1;
    else if (r_stat < o.tol[0] && r_eq < o.tol[1] && r_in < o.tol[2] && r_cp < o.tol[3]) fin = 0;
    else if (it >= o.max_iter) fin = 2;
    if (fin >= 0) { S.status[b] = fin; S.done[b] = 1; S.sqp_iter[b] = it; return 1; }
    return 0;
}

// L1 merit of the trial point w + alpha*dw (dw in S.z), SURVEY A2.5.  `active` = this problem still needs the value
// (inactive threads only take part in the reduction).
template <class Chunk>
QS_HD double merit_at(const SolverDev& S, const double* M, int b, double alpha, const Chunk& ch, bool active) {
    const int N = S.N;
    double mval = 0.0;
    if (active) {
        if (ch.y == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double x0i = fma(alpha, QS_EL(S.z, 2 + i, b), QS_EL(S.x, i, b));
                mval += QS_EL(S.wx0, i, b) * fabs(QS_EL(S.x0, i, b) - x0i);
            }
        }
        for (int k = ch.y; k < N; k += ch.ny) {
            const double un = fma(alpha, QS_EL(S.z, k * 6 + 0, b), QS_EL(S.u, k * 2 + 0, b));
            const double ut = fma(alpha, QS_EL(S.z, k * 6 + 1, b), QS_EL(S.u, k * 2 + 1, b));
            double x[4], xn[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                x[i] = fma(alpha, QS_EL(S.z, k * 6 + 2 + i, b), QS_EL(S.x, k * 4 + i, b));
                xn[i] = fma(alpha, QS_EL(S.z, (k + 1) * 6 + 2 + i, b), QS_EL(S.x, (k + 1) * 4 + i, b));
            }
            // stage cost
            double r[6];
#pragma unroll
            for (int i = 0; i < 4; ++i) r[i] = x[i] - QS_EL(S.yref, k * 6 + i, b);
            r[4] = un - QS_EL(S.yref, k * 6 + 4, b); r[5] = ut - QS_EL(S.yref, k * 6 + 5, b);
            const double* W = S.Wdt + (size_t)k * 36;
            double q = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double a = 0.0;
#pragma unroll
                for (int j = 0; j < 6; ++j) a = fma(W[i + 6 * j], r[j], a);
                q = fma(r[i], a, q);
            }
            mval += 0.5 * q;
            double Phi[4];
            erk4_plain(M, x, un, ut, S.dt, Phi);
#pragma unroll
            for (int i = 0; i < 4; ++i) mval += QS_EL(S.wpi, k * 4 + i, b) * fabs(Phi[i] - xn[i]);
            double h[3];
            h_eval(S, M, x[3], un, ut, h, nullptr);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (!h_on(S.h_variant, k, c)) continue;
                mval += QS_EL(S.wlam, k * 6 + c, b) * fmax(0.0, S.lh[c] - h[c]);
                mval += QS_EL(S.wlam, k * 6 + 3 + c, b) * fmax(0.0, h[c] - S.uh[c]);
            }
        }
        if (ch.y == 0) {
            double r[4], q = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) r[i] = fma(alpha, QS_EL(S.z, N * 6 + 2 + i, b), QS_EL(S.x, N * 4 + i, b)) - QS_EL(S.yref_e, i, b);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double a = 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) a = fma(S.We[i + 4 * j], r[j], a);
                q = fma(r[i], a, q);
            }
            mval += 0.5 * q;
        }
    }
    return ch.sum(mval);
}

// merit-function backtracking line search and iterate update of one SQP iteration, executed by
// the `ch.ny` chunk threads of problem b together (every thread of the group calls it, `live` = problem b exists and
// is not finished); returns 1 on chunk 0 when problem b finishes here (QP failure)
template <class Chunk>
QS_HD int linesearch_one(const SolverDev& S, const SqpOpts& o, const double* __restrict__ Mall, int it, int b, const Chunk& ch, bool live) {
    const int N = S.N;
    int fin = 0;
    if (live && S.qpstat[b] == 2) {
        if (ch.y == 0) { S.status[b] = 4; S.done[b] = 1; S.sqp_iter[b] = it + 1; fin = 1; }
        live = false;
    }
    const double* M = Mall + (size_t)(live ? S.objid[b] : 0) * MODEL_DOUBLES;
    double alpha = 1.0;
    if (o.globalization == 1) {
        // merit weights: |multipliers_qp| first, then max(|m|, (w + |m|)/2); directional derivative
        double dcost = 0.0, dinf = 0.0;
        if (live) {
            for (int k = ch.y; k < N; k += ch.ny) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double a = fabs(QS_EL(S.piq, k * 4 + i, b));
                    const double w = (it == 0) ? a : fmax(a, 0.5 * (QS_EL(S.wpi, k * 4 + i, b) + a));
                    QS_EL(S.wpi, k * 4 + i, b) = w;
                    dinf += w * fabs(QS_EL(S.b, k * 4 + i, b));
                }
                const double h[3] = {QS_EL(S.hv, k * 4, b), QS_EL(S.hv, k * 4 + 1, b), QS_EL(S.hv, k * 4 + 2, b)};
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const double a = fabs(QS_EL(S.lamq, k * 6 + c, b));
                    const double w = (it == 0) ? a : fmax(a, 0.5 * (QS_EL(S.wlam, k * 6 + c, b) + a));
                    QS_EL(S.wlam, k * 6 + c, b) = w;
                    if (!h_on(S.h_variant, k, c % 3)) continue;
                    const int cc = c % 3;
                    dinf += w * (c < 3 ? fmax(0.0, S.lh[cc] - h[cc]) : fmax(0.0, h[cc] - S.uh[cc]));
                }
#pragma unroll
                for (int i = 0; i < 6; ++i) dcost = fma(QS_EL(S.g, k * 6 + i, b), QS_EL(S.z, k * 6 + i, b), dcost);
            }
            if (ch.y == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) dcost = fma(QS_EL(S.qN, i, b), QS_EL(S.z, N * 6 + 2 + i, b), dcost);
                // multiplier of the x0 equality = stage-0 costate of the QP
                double z0[6], m[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) z0[i] = QS_EL(S.z, i, b);
                double pk1[4], a3[4], a4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { pk1[i] = QS_EL(S.piq, i, b); a3[i] = QS_EL(S.A, i, b); a4[i] = QS_EL(S.A, 4 + i, b); }
#pragma unroll
                for (int i = 2; i < 6; ++i) {
                    double a = QS_EL(S.g, i, b);
#pragma unroll
                    for (int j = 0; j < 6; ++j) a = fma(S.H[LT(i, j)], z0[j], a);
                    m[i] = a;
                }
                m[2] += pk1[0]; m[3] += pk1[1]; m[4] += dot4(a3, pk1); m[5] += dot4(a4, pk1);
                if (S.h_variant) {
                    const double beta = QS_EL(S.hv, 3, b);
#pragma unroll
                    for (int c = 0; c < 3; ++c) m[5] += h_bcoef(1, c, beta) * (QS_EL(S.lamq, 3 + c, b) - QS_EL(S.lamq, c, b));
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double a = fabs(m[2 + i]);
                    const double w = (it == 0) ? a : fmax(a, 0.5 * (QS_EL(S.wx0, i, b) + a));
                    QS_EL(S.wx0, i, b) = w;
                    dinf += w * fabs(QS_EL(S.dx0, i, b));
                }
            }
        }
        dcost = ch.sum(dcost); dinf = ch.sum(dinf);
        const double dmerit = dcost - dinf;
        const double m0 = merit_at(S, M, b, 0.0, ch, live);
        bool searching = live;
        for (;;) {
            const double m1 = merit_at(S, M, b, alpha, ch, searching);
            if (searching) {
                if (m1 <= m0 + o.eps_sd * alpha * dmerit) searching = false;
                else {
                    alpha *= o.alpha_red;
                    if (alpha < o.alpha_min) { alpha = o.alpha_min; searching = false; }
                }
            }
            if (!ch.any(searching)) break;
        }
    }
    if (!live) return fin;
    // update: w += alpha dw ; pi <- (1-alpha) pi + alpha pi_qp ; lam likewise
    for (int k = ch.y; k < N; k += ch.ny) {
#pragma unroll
        for (int i = 0; i < 2; ++i) QS_EL(S.u, k * 2 + i, b) = fma(alpha, QS_EL(S.z, k * 6 + i, b), QS_EL(S.u, k * 2 + i, b));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            QS_EL(S.x, k * 4 + i, b) = fma(alpha, QS_EL(S.z, k * 6 + 2 + i, b), QS_EL(S.x, k * 4 + i, b));
            const double po = QS_EL(S.pi, k * 4 + i, b);
            QS_EL(S.pi, k * 4 + i, b) = (1.0 - alpha) * po + alpha * QS_EL(S.piq, k * 4 + i, b);
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const double lo = QS_EL(S.lam, k * 6 + i, b);
            QS_EL(S.lam, k * 6 + i, b) = (1.0 - alpha) * lo + alpha * QS_EL(S.lamq, k * 6 + i, b);
        }
    }
    if (ch.y == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) QS_EL(S.x, N * 4 + i, b) = fma(alpha, QS_EL(S.z, N * 6 + 2 + i, b), QS_EL(S.x, N * 4 + i, b));
        S.alpha[b] = alpha;
    }
    return fin;
}

QS_HD void cost_one(const SolverDev& S, int b) {
    double c = 0.0;
    for (int k = 0; k < S.N; ++k) c += stage_cost(S, k, b);
    S.cost[b] = c + terminal_cost(S, b);
}

// ------------------------------------------------------------------------------------------------
// device-resident closed loop (helper.closed_loop_matlab, helper.m:219-313, for a whole batch)
// ------------------------------------------------------------------------------------------------
struct LoopDev {
    const double* traj;     // [T][6] reference columns [x_ref; u_ref] shared by the batch (controller.y_ref, :425-431)
    const double* off;      // [B][6] per-problem offset added to every column (nullptr: none)
    int T;
    double sigma[4];        // state-noise standard deviations (helper.m:241), all 0 = off
    unsigned long long seed;
    int t_dist;             // 1-based step of the lateral shove (helper.m:222), 0 = none
    double amp, xwidth;     // amplitude_dist, slider_params.xwidth
    int single;             // MATLAB `single` rounding of mod(s, b) inside evalSpline
    // input delays (helper.m:205-212, 244, 250, 290-298; NMPC_controller.m:106-120), in control periods; rings [slot][B][2],
    // slot of the input of period j = (j - 1) mod d, zero-initialised (u_buff_plant / u_buff_contr start as zeros)
    int dp, dc;             // delay_buff_plant = ceil(plant.time_delay / dt), delay_buff_comp = ceil(delay_compensation / dt)
    double* ring_plant;     // [dp][B][2]
    double* ring_contr;     // [dc][B][2]
};

// reference window of control period `idx` (1-based, NMPC_controller.m:307-313, 343-348): stage k uses column
// min(idx + k, T); the terminal reference is the x part of the last window column
QS_HD void loop_window_one(const SolverDev& S, const LoopDev& L, int idx, int k, int b) {
    int col = idx + k; if (col > L.T) col = L.T;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const double v = L.traj[(size_t)(col - 1) * 6 + c] + (L.off ? L.off[(size_t)b * 6 + c] : 0.0);
        QS_EL(S.yref, k * 6 + c, b) = v;
        if (k == S.N - 1 && c < 4) QS_EL(S.yref_e, c, b) = v;
    }
}

// counter-based standard normal: splitmix64 of (seed, step, problem, component) -> two uniforms -> Box-Muller
QS_HD double loop_randn(unsigned long long seed, int step, int b, int c) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(((unsigned long long)step << 34) ^ ((unsigned long long)b << 3) ^ (unsigned long long)c);
    auto mix = [](unsigned long long v) { v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ull; v = (v ^ (v >> 27)) * 0x94D049BB133111EBull; return v ^ (v >> 31); };
    const unsigned long long a = mix(z), bq = mix(z + 0x9E3779B97F4A7C15ull);
    const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740993.0);      // (0, 1)
    const double u2 = (double)(bq >> 11) * (1.0 / 9007199254740992.0);              // [0, 1)
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

// argmin_s |C(s) - target|^2 (helper.m:216-232; fminunc there): scan of 2048 grid points, then Newton on the
// stationarity condition, C / C' / C'' evaluated like bspline_shape.evalSpline (argument wrapped by mod(s, b))
QS_HD double loop_reproject(const double* __restrict__ M, double tx, double ty, bool single) {
    const double bb = M[1];
    double best = 1e300, s = 0.0;
    for (int i = 0; i < 2048; ++i) {
        const double g = bb * (double)i / 2048.0;
        Curve c; curve_eval(M, matlab_mod(g, bb, single), c);
        const double d2 = (c.cx - tx) * (c.cx - tx) + (c.cy - ty) * (c.cy - ty);
        if (d2 < best) { best = d2; s = g; }
    }
    for (int it = 0; it < 20; ++it) {
        const double sg = matlab_mod(s, bb, single);
        Curve c; curve_eval(M, sg, c);
        double ex, ey; curve_dd(M, sg, ex, ey);
        const double rx = c.cx - tx, ry = c.cy - ty;
        const double g1 = 2.0 * (rx * c.dx + ry * c.dy);
        const double g2 = 2.0 * (c.dx * c.dx + c.dy * c.dy + rx * ex + ry * ey);
        if (!(g2 > 0.0)) break;
        const double step = g1 / g2;
        s -= step;
        if (fabs(step) < 1e-14) break;
    }
    // fminunc starts at s = 0 (helper.m:218) and walks to the nearby minimum on the pushed face: report the
    // representative of the minimiser closest to 0, i.e. in (-b/2, b/2]
    if (s > 0.5 * bb) s -= bb;
    return s;
}

// start of control period `step` (1-based) for problem b: disturbance (helper.m:221-236), noise (:240-242) applied to
// the plant state xs [B][4] in place; the state becomes the controller's x0 (constr_x0, NMPC_controller.m:334) and is logged
QS_HD void loop_state_one(const SolverDev& S, const LoopDev& L, const double* __restrict__ Mall, int step, double* xs, double* log_x, int b) {
    const double* M = Mall + (size_t)S.objid[b] * MODEL_DOUBLES;
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = xs[(size_t)b * 4 + i];
    if (L.t_dist > 0 && step == L.t_dist) {
        x[1] += L.amp;                                                      // :224
        Curve c; curve_eval(M, matlab_mod(x[3], M[1], L.single != 0), c);   // :226 Sp = evalSpline(FC, x(4))
        const double smin = loop_reproject(M, -0.5 * L.xwidth, c.cy - L.amp, L.single != 0);   // :227-230
        const double fm = fmod(smin, M[1]);
        double s0 = fm + ((smin < 0.0 && fm != 0.0) ? M[1] : 0.0);          // MATLAB mod(s, b)
        s0 -= M[1] * (smin < 0.0 ? 1.0 : 0.0);                              // :232
        x[3] = s0;
    }
    if (L.sigma[0] != 0.0 || L.sigma[1] != 0.0 || L.sigma[2] != 0.0 || L.sigma[3] != 0.0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] += L.sigma[i] * loop_randn(L.seed, step, b, i);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) xs[(size_t)b * 4 + i] = x[i];
    // controller.delay_buffer_sim (NMPC_controller.m:112-120): the state handed to the controller is the plant state rolled
    // forward through the dc inputs already sent but not yet applied, oldest first (u_buff_contr(:, end - k + 1))
    for (int k = 1; k <= L.dc; ++k) {
        const int j = step - L.dc + k - 1;                                   // period whose input is applied k-th (<= 0: zeros)
        const size_t slot = (size_t)(((j - 1) % L.dc + L.dc) % L.dc);
        const double un = L.ring_contr[(slot * S.B + b) * 2], ut = L.ring_contr[(slot * S.B + b) * 2 + 1];
        Dyn d;
        dyn_eval<false>(M, x[2], x[3], un, ut, d);
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = fma(S.dt, d.f[i], x[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        QS_EL(S.x0, i, b) = x[i];
        if (log_x) log_x[(size_t)b * 4 + i] = x[i];
    }
}

// end of the control period: u = get('u', 0) (NMPC_controller.m:403), forward-Euler plant step (helper.m:294, 307), logs
QS_HD void loop_post_one(const SolverDev& S, const LoopDev& L, const double* __restrict__ Mall, int step, double* xs, double* log_u, int* log_status, int b) {
    const double* M = Mall + (size_t)S.objid[b] * MODEL_DOUBLES;
    const double un = QS_EL(S.u, 0, b), ut = QS_EL(S.u, 1, b);
    if (L.dc > 0) {                                                          // helper.m:250  u_buff_contr = [u(:,i) u_buff_contr(:,1:end-1)]
        const size_t slot = (size_t)((step - 1) % L.dc);
        L.ring_contr[(slot * S.B + b) * 2] = un; L.ring_contr[(slot * S.B + b) * 2 + 1] = ut;
    }
    double ua_n = un, ua_t = ut;                                             // input that reaches the plant in this period
    if (L.dp > 0) {                                                          // helper.m:294-298: u_buff_plant(:, end), then push u(:,i)
        const size_t slot = (size_t)((step - 1) % L.dp);                     // slot of period step - dp = the oldest = the one overwritten
        ua_n = L.ring_plant[(slot * S.B + b) * 2]; ua_t = L.ring_plant[(slot * S.B + b) * 2 + 1];
        L.ring_plant[(slot * S.B + b) * 2] = un; L.ring_plant[(slot * S.B + b) * 2 + 1] = ut;
    }
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = xs[(size_t)b * 4 + i];
    Dyn d;
    dyn_eval<false>(M, x[2], x[3], ua_n, ua_t, d);
#pragma unroll
    for (int i = 0; i < 4; ++i) xs[(size_t)b * 4 + i] = fma(S.dt, d.f[i], x[i]);
    if (log_u) { log_u[(size_t)b * 2] = un; log_u[(size_t)b * 2 + 1] = ut; }
    if (log_status) log_status[b] = S.status[b];
}

// post-processing shift (NMPC_controller.m:397-399) of component c of problem b:
// c = 0..3 x, 4..5 u, 6..9 pi, 10..15 lam
QS_HD void shift_one(const SolverDev& S, int c, int b) {
    const int N = S.N;
    if (c < 4) { for (int k = 0; k < N; ++k) QS_EL(S.x, k * 4 + c, b) = QS_EL(S.x, (k + 1) * 4 + c, b); }
    else if (c < 6) { const int i = c - 4; for (int k = 0; k + 1 < N; ++k) QS_EL(S.u, k * 2 + i, b) = QS_EL(S.u, (k + 1) * 2 + i, b); }
    else if (c < 10) { const int i = c - 6; for (int k = 0; k + 1 < N; ++k) QS_EL(S.pi, k * 4 + i, b) = QS_EL(S.pi, (k + 1) * 4 + i, b); }
    else { const int i = c - 10; for (int k = 0; k + 1 < N; ++k) QS_EL(S.lam, k * 6 + i, b) = QS_EL(S.lam, (k + 1) * 6 + i, b); }
}

}  // namespace qs
