// oracle/qs_oracle.hpp
//
// TEST INFRASTRUCTURE ONLY.  CPU restatement (C++17, FP64, scalar) of the hot path of
// Vanvitelli-Robotics/uclv_qs_pushing_matlab: B-spline contact geometry, quasi-static
// pusher-slider dynamics, ERK4 + forward sensitivities, Gauss-Newton linearisation,
// stage-structured QP (Riccati primal-dual IPM), SQP-RTI / full SQP step and the
// NMPC_controller.solve pre/post-processing.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this.  The product (uclv_qs_pushing_matlab_b200 + libqspush.so) never does.
//
// PARITY UNPINNED against acados v0.2.1 / HPIPM / CasADi: the reference repository holds no
// golden vectors, tests or fixtures for this path and none of MATLAB, Octave, acados, CasADi
// is available in the build container.  What IS pinned (tests/):
//   * spline restatement vs scipy.interpolate.BSpline and the survey's probed constants,
//   * dynamics Jacobian (forward-mode dual numbers, the same rule set as CasADi's AD:
//     d(cmp)=0, d fmod/dx = 1) vs an independent sympy derivation,
//   * QP solutions vs a KKT certificate computed with dense numpy algebra,
//   * SQP solutions vs scipy.optimize on small horizons.
//
// Citations "file:line" are into /root/reference/.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace orc {

// ----------------------------------------------------------------------------------------
// Forward-mode dual number.  Rule set follows CasADi SX AD: comparisons carry no derivative,
// fmod(x, const) has derivative 1 in x.
// ----------------------------------------------------------------------------------------
template <int ND>
struct Dual {
    double v;
    double d[ND];
    Dual() : v(0.0) { for (int i = 0; i < ND; ++i) d[i] = 0.0; }
    Dual(double x) : v(x) { for (int i = 0; i < ND; ++i) d[i] = 0.0; }  // NOLINT implicit
    static Dual seed(double x, int k) { Dual r(x); r.d[k] = 1.0; return r; }
};
template <int ND> inline Dual<ND> operator+(const Dual<ND>& a, const Dual<ND>& b) {
    Dual<ND> r; r.v = a.v + b.v; for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int ND> inline Dual<ND> operator-(const Dual<ND>& a, const Dual<ND>& b) {
    Dual<ND> r; r.v = a.v - b.v; for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int ND> inline Dual<ND> operator-(const Dual<ND>& a) {
    Dual<ND> r; r.v = -a.v; for (int i = 0; i < ND; ++i) r.d[i] = -a.d[i]; return r; }
template <int ND> inline Dual<ND> operator*(const Dual<ND>& a, const Dual<ND>& b) {
    Dual<ND> r; r.v = a.v * b.v; for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int ND> inline Dual<ND> operator/(const Dual<ND>& a, const Dual<ND>& b) {
    Dual<ND> r; r.v = a.v / b.v;
    for (int i = 0; i < ND; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v; return r; }
template <int ND> inline Dual<ND> operator+(const Dual<ND>& a, double b) { return a + Dual<ND>(b); }
template <int ND> inline Dual<ND> operator+(double a, const Dual<ND>& b) { return Dual<ND>(a) + b; }
template <int ND> inline Dual<ND> operator-(const Dual<ND>& a, double b) { return a - Dual<ND>(b); }
template <int ND> inline Dual<ND> operator-(double a, const Dual<ND>& b) { return Dual<ND>(a) - b; }
template <int ND> inline Dual<ND> operator*(const Dual<ND>& a, double b) {
    Dual<ND> r; r.v = a.v * b; for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] * b; return r; }
template <int ND> inline Dual<ND> operator*(double a, const Dual<ND>& b) { return b * a; }
template <int ND> inline Dual<ND> operator/(const Dual<ND>& a, double b) {
    Dual<ND> r; r.v = a.v / b; for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] / b; return r; }
template <int ND> inline Dual<ND> operator/(double a, const Dual<ND>& b) { return Dual<ND>(a) / b; }
template <int ND> inline Dual<ND> sqrt(const Dual<ND>& a) {
    Dual<ND> r; r.v = std::sqrt(a.v); for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] / (2.0 * r.v); return r; }
template <int ND> inline Dual<ND> sin(const Dual<ND>& a) {
    Dual<ND> r; r.v = std::sin(a.v); double c = std::cos(a.v);
    for (int i = 0; i < ND; ++i) r.d[i] = c * a.d[i]; return r; }
template <int ND> inline Dual<ND> cos(const Dual<ND>& a) {
    Dual<ND> r; r.v = std::cos(a.v); double s = -std::sin(a.v);
    for (int i = 0; i < ND; ++i) r.d[i] = s * a.d[i]; return r; }
template <int ND> inline Dual<ND> atan2(const Dual<ND>& y, const Dual<ND>& x) {
    Dual<ND> r; r.v = std::atan2(y.v, x.v); double q = x.v * x.v + y.v * y.v;
    for (int i = 0; i < ND; ++i) r.d[i] = (x.v * y.d[i] - y.v * x.d[i]) / q; return r; }
template <int ND> inline Dual<ND> fmod(const Dual<ND>& a, double m) {
    Dual<ND> r = a; r.v = std::fmod(a.v, m); return r; }
// comparisons: value only, result is a plain 0/1 double (CasADi: derivative-free)
template <int ND> inline double lt(const Dual<ND>& a, double b) { return a.v < b ? 1.0 : 0.0; }
template <int ND> inline double ge(const Dual<ND>& a, double b) { return a.v >= b ? 1.0 : 0.0; }
inline double lt(double a, double b) { return a < b ? 1.0 : 0.0; }
inline double ge(double a, double b) { return a >= b ? 1.0 : 0.0; }
template <int ND> inline double val(const Dual<ND>& a) { return a.v; }

// Second-order Taylor scalar in ONE direction (value, first and second derivative): used to differentiate the
// curvature kappa(s) = d/ds atan2(C'_y, C'_x) once more through the same Cox-de Boor recursion (h_variant 1).
struct Dual2 {
    double v, d, dd;
    Dual2() : v(0.0), d(0.0), dd(0.0) {}
    Dual2(double x) : v(x), d(0.0), dd(0.0) {}  // NOLINT implicit
    static Dual2 seed(double x) { Dual2 r(x); r.d = 1.0; return r; }
};
inline Dual2 operator+(const Dual2& a, const Dual2& b) { Dual2 r; r.v = a.v + b.v; r.d = a.d + b.d; r.dd = a.dd + b.dd; return r; }
inline Dual2 operator-(const Dual2& a, const Dual2& b) { Dual2 r; r.v = a.v - b.v; r.d = a.d - b.d; r.dd = a.dd - b.dd; return r; }
inline Dual2 operator*(const Dual2& a, const Dual2& b) {
    Dual2 r; r.v = a.v * b.v; r.d = a.d * b.v + a.v * b.d; r.dd = a.dd * b.v + 2.0 * a.d * b.d + a.v * b.dd; return r; }
inline Dual2 operator*(const Dual2& a, double b) { Dual2 r; r.v = a.v * b; r.d = a.d * b; r.dd = a.dd * b; return r; }
inline Dual2 operator*(double a, const Dual2& b) { return b * a; }
inline Dual2 operator/(const Dual2& a, double b) { Dual2 r; r.v = a.v / b; r.d = a.d / b; r.dd = a.dd / b; return r; }
inline Dual2 operator-(const Dual2& a, double b) { Dual2 r = a; r.v -= b; return r; }
inline Dual2 operator-(double a, const Dual2& b) { Dual2 r; r.v = a - b.v; r.d = -b.d; r.dd = -b.dd; return r; }
inline double lt(const Dual2& a, double b) { return a.v < b ? 1.0 : 0.0; }
inline double ge(const Dual2& a, double b) { return a.v >= b ? 1.0 : 0.0; }
inline double val(const Dual2& a) { return a.v; }
inline double val(double a) { return a; }
inline double fmod(double a, double m) { return std::fmod(a, m); }
using std::sqrt; using std::sin; using std::cos; using std::atan2;

// ----------------------------------------------------------------------------------------
// Model = bspline_shape (S, P, p, b, cj_1_vect, cj_2) + slider constants.
// ----------------------------------------------------------------------------------------
struct Model {
    int p = 3;                 // degree                               (main.m:33)
    int n = 0;                 // number of control points             (bspline_shape.m:30)
    std::vector<double> S;     // knot vector, length n+p+1            (PusherSliderModel.m:123)
    std::vector<double> P;     // control points, n x 2 row-major      (PusherSliderModel.m:115)
    std::vector<double> c1;    // cj_1_vect, n x 2                     (bspline_shape.m:90-98)
    std::vector<double> c2;    // cj_2, n x 2                          (bspline_shape.m:124-131)
    double b = 0.0;            // total polygon length                 (bspline_shape.m:37)
    double mu_sp = 0.0;        // slider/pusher friction               (object_selection.m)
    double c_ellipse = 0.0;    // tau_max / f_max                      (PusherSliderModel.m:53-55)
    bool single_quirk = true;  // reproduce MATLAB `single` arithmetic inherited from pcread
    bool mod_strict = false;   // OcpOpts::sem_mod_strict
    // knot difference as the reference computes it (single - single when S is single)
    double kd(int i_hi, int i_lo) const {   // 1-based indices like the reference
        if (single_quirk) return (double)((float)S[i_hi - 1] - (float)S[i_lo - 1]);
        return S[i_hi - 1] - S[i_lo - 1];
    }
    double Sk(int i) const { return S[i - 1]; }  // 1-based
};

struct OcpOpts {
    int    max_sqp_iter = 30;          // NMPC_controller.m:276
    double tol_stat = 1e-6, tol_eq = 1e-6, tol_ineq = 1e-6, tol_comp = 1e-6;  // :276
    int    qp_max_iter = 50;           // acados default qp_solver_iter_max
    double qp_tol = 1e-11;             // QP residual tolerance on stationarity, dynamics and inequality residuals
    double qp_mu0 = 0.1;               // initial barrier parameter
    double qp_thr = 1e-3;              // lower clamp on initial slacks
    double qp_tau = 0.9995;            // fixed fraction to the boundary (used when qp_gamma_f = 0)
    double alpha_min = 0.05, alpha_reduction = 0.7, eps_sufficient_descent = 1e-4;  // acados defaults
    int    globalization = 1;          // 1 = merit backtracking (NMPC_controller.m:272), 0 = full step
    // Complementarity is driven much further than the other residuals: the multipliers of this QP are as small as 1e-9
    // (the input weight is 5e-5), so lam * t <= 1e-12 leaves the slack of such a row — and with it du — 1e-5 away from
    // the QP solution; at 1e-18 every row is within 1e-9 and the returned point no longer depends on the path of the IPM
    // (tests/test_qp_exact.py compares it with the extended-precision arbiter, oracle/qs_arbiter.cpp).
    double qp_tol_comp = 1e-18;        // tolerance on max lam * t
    double qp_t_min = 1e-12;           // slack floor: pairs with t <= 4 t_min count as converged, their centering target is lam * t_min (bounds lam / t)
    double qp_gamma_f = 0.01;          // step to the boundary: blocking pair keeps gamma_f * (predicted mu reduction) of its value
    int    qp_stall = 10;              // iterations without halving the normalised residual before a point below 1e-6 is accepted
    int    qp_split_step = 1;          // 1: separate primal / dual step lengths (r02); 0: one step length for both (r01)
    // ---- RECALLED acados v0.2.1 semantics as switches (none of it could be run here).  The defaults are
    // what the restatement believes; when golden vectors from a real acados run (tools/acados_golden.m) disagree,
    // tests/test_acados_golden.py flips these one at a time and reports which flip removes the mismatch.
    int    sem_cost_scale = 0;         // 0: stage cost x dt, terminal unscaled (A2.2); 1: nothing scaled; 2: stage and terminal x dt
    int    sem_h0_s_row = 0;           // 0: the s row of h is void at stage 0 (x_0 is fixed); 1: kept -> a violated bound makes the QP infeasible (status 4)
    int    sem_full_step_dual = 0;     // 0: pi, lam <- (1 - alpha) old + alpha qp (A2.4); 1: multipliers replaced by the QP's whatever alpha
    int    sem_merit_weights = 0;      // 0: w <- |m| first, then max(|m|, (w + |m|) / 2) (A2.5); 1: w <- |m| every iteration; 2: w <- max(w, |m|)
    int    sem_armijo = 0;             // 0: m1 <= m0 + eps alpha dmerit (line_search_use_sufficient_descent = 1); 1: plain decrease m1 < m0
    int    sem_erk_steps = 1;          // ERK4 steps per shooting interval (sim_method_num_steps, A2.1)
    int    sem_qp_maxiter_fails = 0;   // 0: a QP that hits its iteration limit is accepted (A2.4); 1: it ends the solve with status 4
    int    sem_mod_strict = 0;         // 0: MATLAB's mod as its builtin computes it (may return b); 1: result forced into [0, b)
    int    sem_qp_pivot_fails = 0;     // 0: a non-positive Cholesky pivot of the Riccati input block is dropped (inverse diagonal 0, BLASFEO dpotrf) and the IPM goes on; 1: it ends the QP (status 4 in SQP)
};

struct Ocp {
    const Model* model = nullptr;
    int N = 10;                 // horizon Hp                           (main.m:41)
    double dt = 0.05;           // sample time                          (main.m:40)
    std::vector<double> W;      // N x 36, per-stage 6x6 column-major in y=[x;u] order (NMPC_controller.m:157)
    double We[16];              // terminal 4x4 column-major            (NMPC_controller.m:154)
    double lh[3], uh[3];        // bounds on h=[s;u_n;u_t]              (NMPC_controller.m:251-252)
    // h_variant 1 = the authors' parked constraint set (NMPC_controller.m:226-238, commented expr_h at :238):
    // h = [u_n; u_t - v_bound(s); u_t + v_bound(s)] with v_bound of :229 built from the parameters below
    int    h_variant = 0;
    double vb_alpha = 1.0, vb_d = 0.0, vb_t0 = 3.0, vb_ub = 0.05;   // v_alpha :98, d_v_bound :99, t_angle0 :100, u_t_ub :24
    OcpOpts opts;
};

// Per-problem trajectories (column-major like the MATLAB arrays: x is 4x(N+1), u 2xN, pi 4xN).
struct Traj {
    std::vector<double> x, u, pi, lam;  // lam: N x 6 = [lower(s,un,ut); upper(s,un,ut)] per stage
};

struct SolveStats {
    int status = 0;       // 0 ok, 1 NaN, 2 max iter, 3 min step, 4 QP failure (acados v0.2.1 enum)
    int sqp_iter = 0;
    int qp_iter = 0;      // total IPM iterations
    double cost = 0.0;
    double res[4] = {0, 0, 0, 0};   // stat, eq, ineq, comp
    double alpha_last = 1.0;
};

}  // namespace orc
