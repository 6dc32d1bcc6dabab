// oracle/qs_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY — see the header of qs_oracle.hpp.  PARITY UNPINNED against
// acados v0.2.1 (not available anywhere in this environment); pinned against scipy / sympy /
// dense-KKT certificates by tests/.  Citations "file:line" are into /root/reference/.
//
// Build: make -C oracle   ->  oracle/libqs_oracle.so   (g++ -O2, no fast-math: IEEE semantics
// of NaN/Inf comparisons are part of the restated behaviour).

#include "qs_oracle.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <thread>

namespace orc {

// =========================================================================================
// 1. Outline ingest  (PusherSliderModel.m:84-132)
// =========================================================================================

// binary_little_endian PLY with 6 float32 per vertex (x,y,z,nx,ny,nz), no faces.
// Generic enough for any all-float32 vertex element; returns xy in float32 like pcread.
static bool ply_read_xy(const char* path, std::vector<float>& xy) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    char line[512];
    int nv = -1, nprops = 0;
    bool in_vertex = false, header_ok = false, binary_le = false;
    while (std::fgets(line, sizeof line, f)) {
        std::string s(line);
        if (s.rfind("format binary_little_endian", 0) == 0) binary_le = true;
        if (s.rfind("element vertex", 0) == 0) { nv = std::atoi(s.c_str() + 14); in_vertex = true; }
        else if (s.rfind("element", 0) == 0) in_vertex = false;
        if (in_vertex && s.rfind("property float", 0) == 0) ++nprops;
        if (s.rfind("end_header", 0) == 0) { header_ok = true; break; }
    }
    if (!header_ok || !binary_le || nv <= 0 || nprops < 2) { std::fclose(f); return false; }
    std::vector<float> row(nprops);
    xy.resize((size_t)nv * 2);
    for (int i = 0; i < nv; ++i) {
        if (std::fread(row.data(), sizeof(float), nprops, f) != (size_t)nprops) { std::fclose(f); return false; }
        xy[2 * i] = row[0]; xy[2 * i + 1] = row[1];
    }
    std::fclose(f);
    return true;
}

// sortCadPoints, PusherSliderModel.m:84-111.  All arithmetic in float32 (pcread returns single).
static std::vector<float> sort_cad_points(std::vector<float> xy, bool flip) {
    const int nv = (int)xy.size() / 2;
    const float INF = std::numeric_limits<float>::infinity();
    std::vector<float> out; out.reserve((nv + 1) * 2);
    int ind = 0;                                           // [~,ind] = min(Pxy(:,1))   :91
    for (int i = 1; i < nv; ++i) if (xy[2 * i] < xy[2 * ind]) ind = i;
    float cx = xy[2 * ind], cy = xy[2 * ind + 1];
    xy[2 * ind] = INF; xy[2 * ind + 1] = INF;              // :93
    out.push_back(cx); out.push_back(cy);
    for (int k = 1; k < nv; ++k) {                          // :98-103 greedy nearest neighbour
        int best = -1; float bd = INF;
        for (int i = 0; i < nv; ++i) {
            float dx = xy[2 * i] - cx, dy = xy[2 * i + 1] - cy;
            float d = std::sqrt(dx * dx + dy * dy);
            if (best < 0 || d < bd) { bd = d; best = i; }   // min returns the first minimum
        }
        cx = xy[2 * best]; cy = xy[2 * best + 1];
        out.push_back(cx); out.push_back(cy);
        xy[2 * best] = INF; xy[2 * best + 1] = INF;
    }
    const float scale = (float)(1.0 / 1000.0);             // :105, single .* double -> single
    for (auto& v : out) v = v * scale;
    out.push_back(out[0]); out.push_back(out[1]);          // :106 close the polygon
    if (flip) {                                            // :107-109 flipud for montana / pulirapid
        const int n = (int)out.size() / 2;
        for (int i = 0; i < n / 2; ++i) {
            std::swap(out[2 * i], out[2 * (n - 1 - i)]);
            std::swap(out[2 * i + 1], out[2 * (n - 1 - i) + 1]);
        }
    }
    return out;
}

// getSpline knots, PusherSliderModel.m:117-123 (single arithmetic; MATLAB linspace form
// y = d1 + (0:n1).*(d2-d1)./n1 with the end points overwritten).
static void make_knots(const std::vector<float>& P, int p, std::vector<float>& S, float& b) {
    const int n = (int)P.size() / 2;
    const int m = n + p + 1 - 2 * p;
    float acc = 0.f;                                        // b = sum(vecnorm(diff(P)'))  :121
    for (int i = 0; i + 1 < n; ++i) {
        float dx = P[2 * (i + 1)] - P[2 * i], dy = P[2 * (i + 1) + 1] - P[2 * i + 1];
        acc = acc + std::sqrt(dx * dx + dy * dy);
    }
    b = acc;
    S.clear();
    for (int i = 0; i < p; ++i) S.push_back(0.f);
    const float n1 = (float)(m - 1);
    for (int k = 0; k < m; ++k) {
        float v = ((float)k * b) / n1;
        if (k == 0) v = 0.f;
        if (k == m - 1) v = b;
        S.push_back(v);
    }
    for (int i = 0; i < p; ++i) S.push_back(b);
}

// bspline_shape ctor + getSymboliSplineDot/DotDot coefficient tables (bspline_shape.m:25-38,
// 90-98, 124-131).  With single_quirk the divisions happen in float32 as in MATLAB when S and
// P are `single` (single op double -> single).
static void finish_model(Model& m) {
    const int n = m.n, p = m.p;
    m.c1.assign((size_t)n * 2, 0.0);
    m.c2.assign((size_t)n * 2, 0.0);
    for (int ii = 2; ii <= n; ++ii) {                        // :92-100
        if (m.Sk(ii + p) == m.Sk(ii)) continue;
        for (int c = 0; c < 2; ++c) {
            if (m.single_quirk) {
                float num = (float)m.P[2 * (ii - 1) + c] - (float)m.P[2 * (ii - 2) + c];
                float den = (float)m.kd(ii + p, ii);
                m.c1[2 * (ii - 1) + c] = (double)((float)p * (num / den));
            } else {
                m.c1[2 * (ii - 1) + c] = p * ((m.P[2 * (ii - 1) + c] - m.P[2 * (ii - 2) + c]) / m.kd(ii + p, ii));
            }
        }
    }
    for (int ii = 3; ii <= n; ++ii) {                        // :124-132
        if (std::fabs(m.kd(ii + p - 1, ii)) < 1e-5) continue;
        for (int c = 0; c < 2; ++c) {
            double num = m.c1[2 * (ii - 1) + c] - m.c1[2 * (ii - 2) + c];   // double - double
            if (m.single_quirk) {                                           // double / single -> single
                float q = (float)num / (float)m.kd(ii + p - 1, ii);
                m.c2[2 * (ii - 1) + c] = (double)((float)(p - 1) * q);
            } else {
                m.c2[2 * (ii - 1) + c] = (p - 1) * (num / m.kd(ii + p - 1, ii));
            }
        }
    }
}

// =========================================================================================
// 2. Cox-de Boor basis and the curve functions  (eval_bspline.m, bspline_shape.m:40-144)
// =========================================================================================

// Literal recursion, eval_bspline.m:1-33 == bspline_shape.m:40-72 (1-based i).
template <class T>
static T basis_rec(const Model& m, const T& s, int i, int ord) {
    if (m.Sk(i + ord + 1) == m.Sk(i)) return T(0.0);                         // :6-9
    if (ord == 0) return T(lt(s, m.Sk(i + 1)) * ge(s, m.Sk(i)));             // :11-14
    T Na = basis_rec(m, s, i, ord - 1);                                      // :16
    T Nb = basis_rec(m, s, i + 1, ord - 1);                                  // :17
    T m1(0.0), m2(0.0);
    if (!(m.Sk(i + ord) == m.Sk(i))) m1 = (s - m.Sk(i)) / m.kd(i + ord, i);                     // :19-23
    if (!(m.Sk(i + ord + 1) == m.Sk(i + 1))) m2 = (m.Sk(i + ord + 1) - s) / m.kd(i + ord + 1, i + 1);  // :24-28
    return m1 * Na + m2 * Nb;                                                // :30
}

// The half-open span that contains s (1-based j with S(j) <= s < S(j+1)), 0 if none.
static int find_span(const Model& m, double s) {
    const int L = (int)m.S.size();
    for (int j = 1; j < L; ++j)
        if (s >= m.Sk(j) && s < m.Sk(j + 1)) return j;
    return 0;
}

// Sum_i coef_i * N_{i,ord}(s) for i in [i_lo, n].  `local` restricts the sum to the basis
// functions whose support contains s; the skipped terms are exactly +0, so both variants are
// bit-identical for finite s (asserted in tests/test_oracle_spline.py).
template <class T>
static void spline_sum(const Model& m, const T& s, int ord, int i_lo, const std::vector<double>& coef,
                       bool local, T out[2]) {
    out[0] = T(0.0); out[1] = T(0.0);
    int a = i_lo, b = m.n;
    if (local) {
        int j = find_span(m, val(s));
        if (j == 0) return;                       // every indicator is 0 -> sum of zeros
        a = std::max(i_lo, j - ord); b = std::min(m.n, j);
    }
    for (int i = a; i <= b; ++i) {
        T N = basis_rec(m, s, i, ord);
        out[0] = out[0] + N * coef[2 * (i - 1)];
        out[1] = out[1] + N * coef[2 * (i - 1) + 1];
    }
}

template <class T> static void FC(const Model& m, const T& s, bool local, T out[2]) {        // :74-83
    spline_sum(m, s, m.p, 1, m.P, local, out); }
template <class T> static void FC_dot(const Model& m, const T& s, bool local, T out[2]) {    // :85-104
    spline_sum(m, s, m.p - 1, 2, m.c1, local, out); }
template <class T> static void FC_dot_dot(const Model& m, const T& s, bool local, T out[2]) {  // :118-135
    spline_sum(m, s, m.p - 2, 3, m.c2, local, out); }

// getNormalTangentialVersors, bspline_shape.m:106-116.
template <class T>
static void frames(const Model& m, const T& s, bool local, T tv[2], T nv[2]) {
    T d[2]; FC_dot(m, s, local, d);
    T nrm = sqrt(d[0] * d[0] + d[1] * d[1]);           // norm(t_)
    tv[0] = d[0] / nrm; tv[1] = d[1] / nrm;            // :109
    nv[0] = -(-tv[1]); nv[1] = -(tv[0]);               // :110  nvers = -[-t(2) t(1)]
}

// getSymbolicAngleCurvatures, bspline_shape.m:137-144: gradient(atan2(C'_y, C'_x), s).
static double angle_dot(const Model& m, double s, bool local) {
    Dual<1> sd = Dual<1>::seed(s, 0);
    Dual<1> d[2]; FC_dot(m, sd, local, d);
    Dual<1> a = atan2(d[1], d[0]);
    return a.d[0];
}

// kappa(s) and d kappa / ds from the first and second AD derivative of FC_dot through the recursion:
// kappa = (x y' - y x') / (x^2 + y^2) with (x, y) = FC_dot(s)  (the closed form of gradient(atan2(y, x), s)).
static void angle_dot_d(const Model& m, double s, bool local, double* kappa, double* dkappa) {
    Dual2 sd = Dual2::seed(s);
    Dual2 d[2]; FC_dot(m, sd, local, d);
    const double x = d[0].v, y = d[1].v, xp = d[0].d, yp = d[1].d, xpp = d[0].dd, ypp = d[1].dd;
    const double num = x * yp - y * xp, den = x * x + y * y;
    *kappa = num / den;
    *dkappa = ((x * ypp - y * xpp) * den - num * 2.0 * (x * xp + y * yp)) / (den * den);
}

// v_bound(s) of the parked constraint variant (NMPC_controller.m:226-230) and its derivative with CasADi's AD rules
// (comparisons are constants, d fmod = 1, d|a| = sign(a), d fmin(a, c) = [a < c]):
//   s_mod = (s<0)*b + mod(s,b);  t = |kappa(s_mod)|;  v = min(v_alpha / (|t - t0| + 1e-4) + d_v_bound, u_t_ub)
static double v_bound_sym(const Model& m, const Ocp& ocp, double s, bool local, double* dv) {
    const double sw = std::fmod(s, m.b) + ((s < 0.0) ? m.b : 0.0);
    double kap, dkap;
    angle_dot_d(m, sw, local, &kap, &dkap);
    const double t = std::fabs(kap), e = std::fabs(t - ocp.vb_t0) + 0.0001;
    const double a = ocp.vb_alpha / e + ocp.vb_d;
    const double sg1 = (double)((kap > 0.0) - (kap < 0.0)), sg2 = (double)((t - ocp.vb_t0 > 0.0) - (t - ocp.vb_t0 < 0.0));
    if (dv) *dv = (a < ocp.vb_ub) ? -ocp.vb_alpha / (e * e) * sg2 * sg1 * dkap : 0.0;
    return std::fmin(a, ocp.vb_ub);
}

// MATLAB mod(s, b) for b > 0: result in [0, b).  With the single quirk the result is single
// (mod(double, single) -> single), bspline_shape.m:147,155,193; NMPC_controller.m:320,332.
// The builtin's algorithm [MATLAB-RECALL: the form MATLAB Coder emits for mod on floating-point operands]:
//   r = fmod(x, y); r is forced to 0 when x / y is an integer within eps * |x / y|; otherwise r += y when the signs differ.
// The result is NOT always below y: for a tiny negative x the sum r + y rounds to y itself (in single precision already
// for |x| < 1.5e-8 with b = 0.28) — x0(4) = mod(x0(4), b) - b (x0(4) < 0) then gives 0, not -b.
template <class F>
static F matlab_mod_t(F x, F y, F eps) {
    if (y == (F)0) return x;
    if (!(x == x) || !(y == y) || std::isinf(x)) return std::numeric_limits<F>::quiet_NaN();
    if (x == (F)0) return (F)0 / y;
    if (std::isinf(y)) return ((y < (F)0) != (x < (F)0)) ? y : x;
    F r = std::fmod(x, y);
    bool req0 = (r == (F)0);
    if (!req0 && y > std::floor(y)) {
        const F q = std::fabs(x / y);
        req0 = !(std::fabs(q - std::floor(q + (F)0.5)) > eps * q);
    }
    if (req0) r = y * (F)0;
    else if ((x < (F)0) != (y < (F)0)) r += y;
    return r;
}
static double matlab_mod(const Model& m, double s) {
    double r = m.single_quirk ? (double)matlab_mod_t<float>((float)s, (float)m.b, 1.1920929e-7f) : matlab_mod_t<double>(s, m.b, 2.220446049250313e-16);
    if (m.mod_strict && r >= m.b) r -= m.b;                 // sem_mod_strict: the variant that keeps the result inside [0, b)
    return r;
}

// =========================================================================================
// 3. Dynamics  (PusherSliderModel.m:503-603), generic in the scalar type
// =========================================================================================
template <class T>
static void dynamics(const Model& m, const T x[4], const T u[2], bool local, T f[4]) {
    const T& theta = x[2];
    const T& s = x[3];
    const T& u_n = u[0];
    const T& u_t = u[1];
    const double b = m.b;
    T s_mod = fmod(s, b) + lt(s, 0.0) * b;                           // :526
    T S_p[2]; FC(m, s_mod, local, S_p);                              // :528
    T tv[2], nv[2]; frames(m, s_mod, local, tv, nv);                 // :531  R_NT = [n' t']
    T S_p_x = nv[0] * S_p[0] + nv[1] * S_p[1];                       // :532-534  NT_p = R' * S_p'
    T S_p_y = tv[0] * S_p[0] + tv[1] * S_p[1];
    T sin_t = sin(theta), cos_t = cos(theta);                        // :540-541
    const double c = m.c_ellipse, mu = m.mu_sp;
    const double c2 = c * c;
    T factor = 1.0 / (c2 + S_p_x * S_p_x + S_p_y * S_p_y);           // :544
    T gamma_l = (mu * c2 - S_p_x * S_p_y + mu * S_p_x * S_p_x) / (c2 + S_p_y * S_p_y - mu * S_p_x * S_p_y);   // :547
    T gamma_r = (-mu * c2 - S_p_x * S_p_y - mu * S_p_x * S_p_x) / (c2 + S_p_y * S_p_y + mu * S_p_x * S_p_y);  // :548
    T u_fract = u_t / u_n;                                           // :551
    // W_R_S * S_R_NT * factor   (:554, :559 — left to right)
    T WR[2][2];
    T R[2][2] = {{nv[0], tv[0]}, {nv[1], tv[1]}};
    T Wm[2][2] = {{cos_t, -sin_t}, {sin_t, cos_t}};
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) WR[i][j] = (Wm[i][0] * R[0][j] + Wm[i][1] * R[1][j]) * factor;
    T Q[2][2] = {{c2 + S_p_x * S_p_x, S_p_x * S_p_y}, {S_p_x * S_p_y, c2 + S_p_y * S_p_y}};   // :555
    T WRQ[2][2];
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) WRQ[i][j] = WR[i][0] * Q[0][j] + WR[i][1] * Q[1][j];
    // sticking :557-560
    T xd_st[4];
    xd_st[0] = WRQ[0][0] * u_n + WRQ[0][1] * u_t;
    xd_st[1] = WRQ[1][0] * u_n + WRQ[1][1] * u_t;
    xd_st[2] = (factor * (-S_p_y)) * u_n + (factor * S_p_x) * u_t;
    xd_st[3] = T(0.0);
    // sliding left :564-573 ; sliding right :576-585.  P_s* = [v zeros(2,1)], b_s* = [-py+g*px 0]
    T xd_sl[4], xd_sr[4];
    {
        T c0 = WRQ[0][0] + WRQ[0][1] * gamma_l, c1 = WRQ[1][0] + WRQ[1][1] * gamma_l;
        xd_sl[0] = c0 * u_n + T(0.0) * u_t;
        xd_sl[1] = c1 * u_n + T(0.0) * u_t;
        xd_sl[2] = (factor * (-S_p_y + gamma_l * S_p_x)) * u_n + (factor * T(0.0)) * u_t;
        xd_sl[3] = u_t - u_n * gamma_l;                              // :570 s_dot_sl(2)
    }
    {
        T c0 = WRQ[0][0] + WRQ[0][1] * gamma_r, c1 = WRQ[1][0] + WRQ[1][1] * gamma_r;
        xd_sr[0] = c0 * u_n + T(0.0) * u_t;
        xd_sr[1] = c1 * u_n + T(0.0) * u_t;
        xd_sr[2] = (factor * (-S_p_y + gamma_r * S_p_x)) * u_n + (factor * T(0.0)) * u_t;
        xd_sr[3] = u_t - u_n * gamma_r;                              // :582
    }
    // :587-589  indicator products (IEEE: NaN compares false; 0*NaN stays NaN)
    const double r = val(u_fract), gl = val(gamma_l), gr = val(gamma_r);
    const double i_st_a = (r >= gr) ? 1.0 : 0.0, i_st_b = (r <= gl) ? 1.0 : 0.0;
    const double i_sl = (r > gl) ? 1.0 : 0.0, i_sr = (r < gr) ? 1.0 : 0.0;
    for (int i = 0; i < 4; ++i) f[i] = i_st_a * xd_st[i] * i_st_b + i_sl * xd_sl[i] + i_sr * xd_sr[i];
}

static void dynamics_num(const Model& m, const double x[4], const double u[2], bool local, double f[4]) {
    dynamics<double>(m, x, u, local, f);
}

// f and its Jacobians by forward AD (CasADi rule set).  Jx 4x4, Ju 4x2, row-major.
static void dynamics_jac(const Model& m, const double x[4], const double u[2], bool local,
                         double f[4], double Jx[16], double Ju[8]) {
    typedef Dual<4> D;  // directions: theta, s, u_n, u_t  (x and y do not enter f)
    D xd[4] = {D(x[0]), D(x[1]), D::seed(x[2], 0), D::seed(x[3], 1)};
    D ud[2] = {D::seed(u[0], 2), D::seed(u[1], 3)};
    D fd[4];
    dynamics<D>(m, xd, ud, local, fd);
    for (int i = 0; i < 4; ++i) {
        f[i] = fd[i].v;
        Jx[4 * i + 0] = 0.0; Jx[4 * i + 1] = 0.0;
        Jx[4 * i + 2] = fd[i].d[0]; Jx[4 * i + 3] = fd[i].d[1];
        Ju[2 * i + 0] = fd[i].d[2]; Ju[2 * i + 1] = fd[i].d[3];
    }
}

// =========================================================================================
// 4. ERK4 + forward sensitivities  (acados sim_erk with expl_vde_forw)
//    X = [x, Sx(4x4), Su(4x2)],  Xdot = [f, Jx*Sx, Jx*Su + Ju],  one step, 4 stages.
// =========================================================================================
static void vde_forw(const Model& m, const double* X, const double u[2], bool local, double* Xd) {
    double f[4], Jx[16], Ju[8];
    dynamics_jac(m, X, u, local, f, Jx, Ju);
    const double* Sx = X + 4; const double* Su = X + 20;
    for (int i = 0; i < 4; ++i) Xd[i] = f[i];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double a = 0.0;
            for (int k = 0; k < 4; ++k) a += Jx[4 * i + k] * Sx[4 * k + j];
            Xd[4 + 4 * i + j] = a;
        }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 2; ++j) {
            double a = 0.0;
            for (int k = 0; k < 4; ++k) a += Jx[4 * i + k] * Su[2 * k + j];
            Xd[20 + 2 * i + j] = a + Ju[2 * i + j];
        }
}

// Phi (4), A = dPhi/dx (4x4 row-major), B = dPhi/du (4x2 row-major)
static void erk4_sens(const Model& m, const double x[4], const double u[2], double dt, bool local,
                      double Phi[4], double A[16], double B[8]) {
    const int NX = 28;
    double X0[NX] = {0};
    for (int i = 0; i < 4; ++i) X0[i] = x[i];
    for (int i = 0; i < 4; ++i) X0[4 + 4 * i + i] = 1.0;
    static const double a_[4] = {0.0, 0.5, 0.5, 1.0};
    static const double b_[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
    double K[4][NX], Xs[NX], Xn[NX];
    for (int i = 0; i < NX; ++i) Xn[i] = X0[i];
    for (int st = 0; st < 4; ++st) {
        for (int i = 0; i < NX; ++i) Xs[i] = X0[i] + (st ? dt * a_[st] * K[st - 1][i] : 0.0);
        vde_forw(m, Xs, u, local, K[st]);
    }
    for (int i = 0; i < NX; ++i) {
        double acc = 0.0;
        for (int st = 0; st < 4; ++st) acc += b_[st] * K[st][i];
        Xn[i] = X0[i] + dt * acc;
    }
    for (int i = 0; i < 4; ++i) Phi[i] = Xn[i];
    for (int i = 0; i < 16; ++i) A[i] = Xn[4 + i];
    for (int i = 0; i < 8; ++i) B[i] = Xn[20 + i];
}

// num_steps ERK4 steps of dt / num_steps each with chained sensitivities (sim_method_num_steps; 1 in the reference's configuration)
static void erk4_sens_steps(const Model& m, const double x[4], const double u[2], double dt, int steps, bool local,
                            double Phi[4], double A[16], double B[8]) {
    if (steps <= 1) { erk4_sens(m, x, u, dt, local, Phi, A, B); return; }
    double xc[4] = {x[0], x[1], x[2], x[3]}, Ac[16], Bc[8];
    for (int i = 0; i < 16; ++i) Ac[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int i = 0; i < 8; ++i) Bc[i] = 0.0;
    for (int st = 0; st < steps; ++st) {
        double P1[4], A1[16], B1[8], An[16], Bn[8];
        erk4_sens(m, xc, u, dt / steps, local, P1, A1, B1);
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double a = 0; for (int l = 0; l < 4; ++l) a += A1[4 * i + l] * Ac[4 * l + j]; An[4 * i + j] = a; }
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 2; ++j) { double a = B1[2 * i + j]; for (int l = 0; l < 4; ++l) a += A1[4 * i + l] * Bc[2 * l + j]; Bn[2 * i + j] = a; }
        for (int i = 0; i < 16; ++i) Ac[i] = An[i];
        for (int i = 0; i < 8; ++i) Bc[i] = Bn[i];
        for (int i = 0; i < 4; ++i) xc[i] = P1[i];
    }
    for (int i = 0; i < 4; ++i) Phi[i] = xc[i];
    for (int i = 0; i < 16; ++i) A[i] = Ac[i];
    for (int i = 0; i < 8; ++i) B[i] = Bc[i];
}

// ERK4 without sensitivities (merit-function evaluations of the line search)
static void erk4(const Model& m, const double x[4], const double u[2], double dt, bool local, double Phi[4]) {
    double k1[4], k2[4], k3[4], k4[4], xs[4];
    dynamics_num(m, x, u, local, k1);
    for (int i = 0; i < 4; ++i) xs[i] = x[i] + 0.5 * dt * k1[i];
    dynamics_num(m, xs, u, local, k2);
    for (int i = 0; i < 4; ++i) xs[i] = x[i] + 0.5 * dt * k2[i];
    dynamics_num(m, xs, u, local, k3);
    for (int i = 0; i < 4; ++i) xs[i] = x[i] + dt * k3[i];
    dynamics_num(m, xs, u, local, k4);
    for (int i = 0; i < 4; ++i)
        Phi[i] = x[i] + dt * (k1[i] / 6.0 + k2[i] / 3.0 + k3[i] / 3.0 + k4[i] / 6.0);
}

// =========================================================================================
// 5. Stage-structured QP and a Riccati-based Mehrotra primal-dual IPM  
//    Stage variables z_k = [u_k; x_k] (acados/HPIPM order).  Inequalities are the three
//    selection rows of h = [s; u_n; u_t]:  z[5], z[0], z[1].
// =========================================================================================
struct StageQP {
    double H[6][6];   // Gauss-Newton Hessian, [u;x] order
    double g[6];
    double A[4][4], B[4][2], b[4];   // dx+ = A dx + B du + b
    double dl[3], du[3];             // lh - h , uh - h
    bool   on[3];                    // constraint present (s-bound is dropped at k=0: x0 is fixed)
    int    ci[3];                    // constraint row c = e_{ci[c]} + beta[c] * e_5 in z = [u_n,u_t,x,y,theta,s]
    double beta[3];                  // (h_variant 0: rows are selections, beta = 0; h_variant 1: beta = -+ v_bound'(s))
};
struct QP {
    int N = 0;
    std::vector<StageQP> st;
    double QN[4][4], qN[4];
    double dx0[4];
    bool infeasible = false;                 // sem_h0_s_row: the constant s row of stage 0 violates its bound
};
struct QPSol {
    std::vector<double> du, dx, pi, lam, t;   // du N*2, dx (N+1)*4, pi N*4 (pi_{k+1}), lam/t N*6 [lower;upper]
    int iters = 0, status = 0;                // status 0 ok, 1 max iter, 2 NaN / breakdown
    double res[4] = {0, 0, 0, 0};
};
static const int CIDX[3] = {5, 0, 1};   // h = [s;u_n;u_t] inside z = [u_n,u_t,x,y,theta,s]

struct RiccatiFactor {
    std::vector<double> K;     // N x (2x4)
    std::vector<double> Linv;  // N x 3 : inverse of the 2x2 Cholesky factor (l00,l10,l11 of L, stored raw)
    std::vector<double> P;     // (N+1) x 16
};

// Backward matrix recursion with the barrier-augmented Hessian diag term Hb (N x 6).
// Division by a Cholesky pivot: a dropped pivot (stored as 0) zeroes the quotient — BLASFEO's dpotrf sets the inverse
// diagonal of a non-positive pivot to 0 instead of failing [BLASFEO-RECALL], so HPIPM's Riccati never aborts on one.
static inline double pdiv(double a, double l) { return l > 0.0 ? a / l : 0.0; }

static bool riccati_factor(const QP& qp, const std::vector<double>& Hb, const std::vector<double>& Hx, RiccatiFactor& F, bool pivot_fails) {
    const int N = qp.N;
    F.K.assign((size_t)N * 8, 0.0); F.Linv.assign((size_t)N * 3, 0.0); F.P.assign((size_t)(N + 1) * 16, 0.0);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) F.P[(size_t)N * 16 + 4 * i + j] = qp.QN[i][j];
    for (int k = N - 1; k >= 0; --k) {
        const StageQP& s = qp.st[k];
        const double* Pn = &F.P[(size_t)(k + 1) * 16];
        double BA[4][6];
        for (int i = 0; i < 4; ++i) { BA[i][0] = s.B[i][0]; BA[i][1] = s.B[i][1]; for (int j = 0; j < 4; ++j) BA[i][2 + j] = s.A[i][j]; }
        double PBA[4][6];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 6; ++j) {
            double a = 0.0; for (int l = 0; l < 4; ++l) a += Pn[4 * i + l] * BA[l][j]; PBA[i][j] = a; }
        double M[6][6];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
            double a = s.H[i][j]; for (int l = 0; l < 4; ++l) a += BA[l][i] * PBA[l][j]; M[i][j] = a; }
        for (int i = 0; i < 6; ++i) M[i][i] += Hb[(size_t)k * 6 + i];
        M[5][1] += Hx[k]; M[1][5] += Hx[k];                  // barrier cross term (s, u_t) of the coupled rows
        // Cholesky of the 2x2 input block
        if (!(M[0][0] == M[0][0]) || !(M[1][1] == M[1][1]) || !(M[1][0] == M[1][0])) return false;    // NaN: no factorisation
        double l00 = M[0][0] > 0.0 ? std::sqrt(M[0][0]) : 0.0;
        if (pivot_fails && !(l00 > 0.0)) return false;
        double l10 = pdiv(M[1][0], l00);
        double d11 = M[1][1] - l10 * l10;
        if (pivot_fails && !(d11 > 0.0)) return false;
        double l11 = d11 > 0.0 ? std::sqrt(d11) : 0.0;
        F.Linv[(size_t)k * 3 + 0] = l00; F.Linv[(size_t)k * 3 + 1] = l10; F.Linv[(size_t)k * 3 + 2] = l11;
        // K = Muu^{-1} Mux  (2x4)
        double* K = &F.K[(size_t)k * 8];
        for (int j = 0; j < 4; ++j) {
            double y0 = pdiv(M[0][2 + j], l00);
            double y1 = pdiv(M[1][2 + j] - l10 * y0, l11);
            double k1 = pdiv(y1, l11);
            double k0 = pdiv(y0 - l10 * k1, l00);
            K[j] = k0; K[4 + j] = k1;
        }
        double* Pk = &F.P[(size_t)k * 16];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j)
            Pk[4 * i + j] = M[2 + i][2 + j] - (M[2 + i][0] * K[j] + M[2 + i][1] * K[4 + j]);
        for (int i = 0; i < 4; ++i) for (int j = i + 1; j < 4; ++j) {   // keep P symmetric
            double a = 0.5 * (Pk[4 * i + j] + Pk[4 * j + i]); Pk[4 * i + j] = a; Pk[4 * j + i] = a; }
    }
    return true;
}

// Solve  [Htilde, dyn'; dyn, 0] [dz; dpi] = [-rg; -rb]  with dx_0 given.
// rg: N*6 + 4 (stage gradients then terminal), rb: N*4 (residual of x_{k+1} eq.: A dx + B du - dx+ = -rb)
static void riccati_solve(const QP& qp, const std::vector<double>& Hb, const RiccatiFactor& F,
                          const std::vector<double>& rg, const std::vector<double>& rb, const double dx0[4],
                          std::vector<double>& dz, std::vector<double>& dxN, std::vector<double>& dpi) {
    const int N = qp.N;
    (void)Hb;
    std::vector<double> kff((size_t)N * 2), pvec((size_t)(N + 1) * 4);
    for (int i = 0; i < 4; ++i) pvec[(size_t)N * 4 + i] = rg[(size_t)N * 6 + i];
    for (int k = N - 1; k >= 0; --k) {
        const StageQP& s = qp.st[k];
        const double* Pn = &F.P[(size_t)(k + 1) * 16];
        const double* pn = &pvec[(size_t)(k + 1) * 4];
        double w[4];   // P b + p  with b := rb (dynamics residual of the step equation)
        for (int i = 0; i < 4; ++i) { double a = pn[i]; for (int l = 0; l < 4; ++l) a += Pn[4 * i + l] * rb[(size_t)k * 4 + l]; w[i] = a; }
        double mvec[6];
        for (int j = 0; j < 2; ++j) { double a = rg[(size_t)k * 6 + j]; for (int l = 0; l < 4; ++l) a += s.B[l][j] * w[l]; mvec[j] = a; }
        for (int j = 0; j < 4; ++j) { double a = rg[(size_t)k * 6 + 2 + j]; for (int l = 0; l < 4; ++l) a += s.A[l][j] * w[l]; mvec[2 + j] = a; }
        const double l00 = F.Linv[(size_t)k * 3], l10 = F.Linv[(size_t)k * 3 + 1], l11 = F.Linv[(size_t)k * 3 + 2];
        double y0 = pdiv(mvec[0], l00), y1 = pdiv(mvec[1] - l10 * y0, l11);
        double k1 = pdiv(y1, l11), k0 = pdiv(y0 - l10 * k1, l00);
        kff[(size_t)k * 2] = k0; kff[(size_t)k * 2 + 1] = k1;
        // p_k = m_x - Mxu kff, with Mxu = K' Muu  ->  Mxu kff = K' (Muu kff) = K' m_u
        const double* K = &F.K[(size_t)k * 8];
        for (int j = 0; j < 4; ++j) pvec[(size_t)k * 4 + j] = mvec[2 + j] - (K[j] * mvec[0] + K[4 + j] * mvec[1]);
    }
    dz.assign((size_t)N * 6, 0.0); dxN.assign(4, 0.0); dpi.assign((size_t)N * 4, 0.0);
    double xk[4] = {dx0[0], dx0[1], dx0[2], dx0[3]};
    for (int k = 0; k < N; ++k) {
        const StageQP& s = qp.st[k];
        const double* K = &F.K[(size_t)k * 8];
        double uk[2];
        for (int j = 0; j < 2; ++j) { double a = -kff[(size_t)k * 2 + j]; for (int l = 0; l < 4; ++l) a -= K[4 * j + l] * xk[l]; uk[j] = a; }
        dz[(size_t)k * 6] = uk[0]; dz[(size_t)k * 6 + 1] = uk[1];
        for (int l = 0; l < 4; ++l) dz[(size_t)k * 6 + 2 + l] = xk[l];
        double xn[4];
        for (int i = 0; i < 4; ++i) { double a = rb[(size_t)k * 4 + i]; for (int l = 0; l < 4; ++l) a += s.A[i][l] * xk[l]; a += s.B[i][0] * uk[0] + s.B[i][1] * uk[1]; xn[i] = a; }
        const double* Pn = &F.P[(size_t)(k + 1) * 16];
        for (int i = 0; i < 4; ++i) { double a = pvec[(size_t)(k + 1) * 4 + i]; for (int l = 0; l < 4; ++l) a += Pn[4 * i + l] * xn[l]; dpi[(size_t)k * 4 + i] = a; }
        for (int i = 0; i < 4; ++i) xk[i] = xn[i];
    }
    for (int i = 0; i < 4; ++i) dxN[i] = xk[i];
}

// Mehrotra predictor-corrector IPM, infeasible start, single step length for primal and dual.
static void qp_solve_ipm(const QP& qp, const OcpOpts& o, QPSol& sol) {
    const int N = qp.N;
    const size_t nz = (size_t)N * 6, nc = (size_t)N * 6;
    std::vector<double> z(nz, 0.0), xN(4, 0.0), pi((size_t)N * 4, 0.0), lam(nc, 0.0), t(nc, 1.0);
    std::vector<char> on(nc, 0);
    // z_0 carries the fixed dx0 in its x part
    for (int i = 0; i < 4; ++i) z[2 + i] = qp.dx0[i];
    int m_on = 0;
    for (int k = 0; k < N; ++k)
        for (int c = 0; c < 3; ++c) {
            const bool act = qp.st[k].on[c];
            on[(size_t)k * 6 + c] = act; on[(size_t)k * 6 + 3 + c] = act;
            if (!act) continue;
            m_on += 2;
            double v = z[(size_t)k * 6 + qp.st[k].ci[c]] + qp.st[k].beta[c] * z[(size_t)k * 6 + 5];
            double tl = v - qp.st[k].dl[c], tu = qp.st[k].du[c] - v;
            tl = std::max(tl, o.qp_thr); tu = std::max(tu, o.qp_thr);
            t[(size_t)k * 6 + c] = tl; t[(size_t)k * 6 + 3 + c] = tu;
            lam[(size_t)k * 6 + c] = o.qp_mu0 / tl; lam[(size_t)k * 6 + 3 + c] = o.qp_mu0 / tu;
        }
    std::vector<double> rg(nz + 4), rb((size_t)N * 4), rd(nc), rm(nc), Hb(nz), Hx((size_t)N), rgt(nz + 4);
    std::vector<double> dz, dxN, dpi, dlam(nc), dt_(nc), dz2, dxN2, dpi2, wgt(nc, 0.0);
    RiccatiFactor F;
    // true residuals of the current point: rg (stationarity), rb (dynamics), rd (inequality rows minus slacks)
    auto residuals = [&]() {
        for (int k = 0; k < N; ++k) {
            const StageQP& s = qp.st[k];
            const double* zk = &z[(size_t)k * 6];
            const double* xn = (k + 1 < N) ? &z[(size_t)(k + 1) * 6 + 2] : xN.data();
            const double* pk1 = &pi[(size_t)k * 4];                      // pi_{k+1}
            for (int i = 0; i < 6; ++i) {
                double a = s.g[i];
                for (int j = 0; j < 6; ++j) a += s.H[i][j] * zk[j];
                rg[(size_t)k * 6 + i] = a;
            }
            for (int j = 0; j < 2; ++j) for (int l = 0; l < 4; ++l) rg[(size_t)k * 6 + j] += s.B[l][j] * pk1[l];
            for (int j = 0; j < 4; ++j) for (int l = 0; l < 4; ++l) rg[(size_t)k * 6 + 2 + j] += s.A[l][j] * pk1[l];
            if (k > 0) for (int j = 0; j < 4; ++j) rg[(size_t)k * 6 + 2 + j] -= pi[(size_t)(k - 1) * 4 + j];
            for (int c = 0; c < 3; ++c) if (s.on[c]) {
                const double dlm = -lam[(size_t)k * 6 + c] + lam[(size_t)k * 6 + 3 + c];
                rg[(size_t)k * 6 + s.ci[c]] += dlm;
                rg[(size_t)k * 6 + 5] += s.beta[c] * dlm;
            }
            for (int i = 0; i < 4; ++i) {
                double a = s.b[i] - xn[i];
                for (int l = 0; l < 4; ++l) a += s.A[i][l] * zk[2 + l];
                a += s.B[i][0] * zk[0] + s.B[i][1] * zk[1];
                rb[(size_t)k * 4 + i] = a;
            }
            for (int c = 0; c < 3; ++c) {
                if (!s.on[c]) { rd[(size_t)k * 6 + c] = 0; rd[(size_t)k * 6 + 3 + c] = 0; continue; }
                double v = zk[s.ci[c]] + s.beta[c] * zk[5];
                rd[(size_t)k * 6 + c] = v - s.dl[c] - t[(size_t)k * 6 + c];
                rd[(size_t)k * 6 + 3 + c] = s.du[c] - v - t[(size_t)k * 6 + 3 + c];
            }
        }
        for (int i = 0; i < 4; ++i) {
            double a = qp.qN[i];
            for (int j = 0; j < 4; ++j) a += qp.QN[i][j] * xN[j];
            rg[nz + i] = a - pi[(size_t)(N - 1) * 4 + i];
        }
        // x_0 is not a variable: its stationarity row is not a residual
        for (int j = 0; j < 4; ++j) rg[2 + j] = 0.0;
    };
    sol.status = 1;
    const double tol_cp = o.qp_tol_comp, t_min = o.qp_t_min;
    int it = 0, stall = 0;
    double rmax_prev = 1e300;
    for (;; ++it) {
        residuals();
        double mu = 0.0, r_stat = 0.0, r_eq = 0.0, r_in = 0.0, r_cp = 0.0;
        for (size_t i = 0; i < nc; ++i) if (on[i]) { wgt[i] = (t[i] > 4.0 * t_min) ? 1.0 : 0.0; mu += wgt[i] * lam[i] * t[i]; r_cp = std::max(r_cp, wgt[i] * lam[i] * t[i]); r_in = std::max(r_in, std::fabs(rd[i])); }
        mu = m_on ? mu / m_on : 0.0;
        for (size_t i = 0; i < nz + 4; ++i) r_stat = std::max(r_stat, std::fabs(rg[i]));
        for (size_t i = 0; i < rb.size(); ++i) r_eq = std::max(r_eq, std::fabs(rb[i]));
        sol.res[0] = r_stat; sol.res[1] = r_eq; sol.res[2] = r_in; sol.res[3] = r_cp;
        if (!(r_stat == r_stat) || !(r_eq == r_eq) || !(mu == mu)) { sol.status = 2; break; }
        if (r_stat < o.qp_tol && r_eq < o.qp_tol && r_in < o.qp_tol && r_cp < tol_cp) { sol.status = 0; break; }
        {   // stall exit (weakly active pairs): residuals stopped moving below the reference's QP tolerance 1e-6
            const double rmax = std::max(std::max(r_stat, r_eq), std::max(r_in, r_cp));
            const double rrel = std::max(std::max(r_stat, std::max(r_eq, r_in)) / o.qp_tol, r_cp / tol_cp);   // > 1: not converged
            if (rrel < 0.5 * rmax_prev) { rmax_prev = rrel; stall = 0; } else ++stall;   // rmax_prev = best so far
            if (stall >= o.qp_stall && rmax < 1e-6) { sol.status = 0; break; }
        }
        if (it >= o.qp_max_iter) { sol.status = 1; break; }
        // ---- factorise with barrier diagonal
        std::fill(Hb.begin(), Hb.end(), 0.0); std::fill(Hx.begin(), Hx.end(), 0.0);
        for (int k = 0; k < N; ++k) for (int c = 0; c < 3; ++c) if (qp.st[k].on[c]) {
            const double D = lam[(size_t)k * 6 + c] / t[(size_t)k * 6 + c] + lam[(size_t)k * 6 + 3 + c] / t[(size_t)k * 6 + 3 + c];
            const double be = qp.st[k].beta[c];
            Hb[(size_t)k * 6 + qp.st[k].ci[c]] += D;          // D a a' with a = e_ci + beta e_5
            Hb[(size_t)k * 6 + 5] += be * be * D;
            Hx[k] += be * D;                                   // (beta != 0 only for rows on u_t: ci = 1)
        }
        if (!riccati_factor(qp, Hb, Hx, F, o.sem_qp_pivot_fails != 0)) { sol.status = 2; break; }
        const double zero4[4] = {0, 0, 0, 0};
        auto solve_with = [&](const std::vector<double>& rmv, std::vector<double>& dz_, std::vector<double>& dxN_, std::vector<double>& dpi_) {
            rgt = rg;
            for (int k = 0; k < N; ++k) for (int c = 0; c < 3; ++c) if (qp.st[k].on[c]) {
                const size_t il = (size_t)k * 6 + c, iu = il + 3;
                const double w = (rmv[il] + lam[il] * rd[il]) / t[il] - (rmv[iu] + lam[iu] * rd[iu]) / t[iu];
                rgt[(size_t)k * 6 + qp.st[k].ci[c]] += w;
                rgt[(size_t)k * 6 + 5] += qp.st[k].beta[c] * w;
            }
            riccati_solve(qp, Hb, F, rgt, rb, zero4, dz_, dxN_, dpi_);
            for (int k = 0; k < N; ++k) for (int c = 0; c < 3; ++c) {
                const size_t il = (size_t)k * 6 + c, iu = il + 3;
                if (!qp.st[k].on[c]) { dt_[il] = dt_[iu] = dlam[il] = dlam[iu] = 0.0; continue; }
                double dv = dz_[(size_t)k * 6 + qp.st[k].ci[c]] + qp.st[k].beta[c] * dz_[(size_t)k * 6 + 5];
                dt_[il] = dv + rd[il]; dt_[iu] = -dv + rd[iu];
                dlam[il] = -(rmv[il] + lam[il] * dt_[il]) / t[il];
                dlam[iu] = -(rmv[iu] + lam[iu] * dt_[iu]) / t[iu];
            }
        };
        auto max_step = [&]() {
            double a = 1.0;
            for (size_t i = 0; i < nc; ++i) if (on[i]) {
                if (dt_[i] < 0.0) a = std::min(a, -t[i] / dt_[i]);
                if (dlam[i] < 0.0) a = std::min(a, -lam[i] / dlam[i]);
            }
            return a;
        };
        // ---- predictor (affine scaling)
        for (size_t i = 0; i < nc; ++i) rm[i] = on[i] ? lam[i] * t[i] : 0.0;
        solve_with(rm, dz, dxN, dpi);
        double a_aff = max_step();
        double mu_aff = 0.0;
        for (size_t i = 0; i < nc; ++i) if (on[i]) mu_aff += wgt[i] * (lam[i] + a_aff * dlam[i]) * (t[i] + a_aff * dt_[i]);
        mu_aff = m_on ? mu_aff / m_on : 0.0;
        double sigma = (mu > 0.0) ? (mu_aff / mu) : 0.0; sigma = sigma * sigma * sigma;
        // ---- corrector
        const double smu = std::max(sigma * mu, 0.1 * tol_cp);   // centering target floor (keeps lam/t bounded at the end)
        for (size_t i = 0; i < nc; ++i) rm[i] = on[i] ? (lam[i] * t[i] + dlam[i] * dt_[i] - std::max(smu, lam[i] * t_min)) : 0.0;
        solve_with(rm, dz, dxN, dpi);
        // Step to the boundary (Mehrotra's heuristic in scalar form): the blocking pair keeps the fraction
        // gamma_f * (predicted reduction of mu) of its value, clamped to [1e-8, 0.5], instead of a fixed 1 - tau.  A blocked step
        // (a_max < 1) then leaves its blocking pair near the new central path instead of 1e-3 below it — the fixed rule made
        // inputs whose only curvature is the 5e-5 weight jump from bound to bound every iteration —, while an unblocked end-game
        // step may shrink a product by sigma (superlinear convergence; with 1 - tau = 5e-4 every iteration gains 3 digits at most).
        const double a_max_ = max_step();
        const double red = 1.0 - std::min(a_max_, 1.0) * (1.0 - smu / std::max(mu, 1e-300));
        const double tau_k = (o.qp_gamma_f > 0.0) ? 1.0 - std::min(std::max(o.qp_gamma_f * red, 1e-8), 0.5) : o.qp_tau;   // gamma_f = 0: fixed fraction qp_tau
        double alpha = std::min(1.0, tau_k * a_max_), alpha_d = alpha;
        if (o.qp_split_step) {
            // separate primal (z, t) and dual (pi, lam) step lengths, each to its own boundary with the common tau_k (HPIPM's
            // split_step [HPIPM-RECALL]): the side that is not blocked takes the longer step, the mismatch it leaves in the
            // stationarity residual is part of the next iteration's TRUE residuals.  2-7 % fewer iterations on every configuration.
            double ap = 1.0, ad = 1.0;
            for (size_t i = 0; i < nc; ++i) if (on[i]) {
                if (dt_[i] < 0.0) ap = std::min(ap, -t[i] / dt_[i]);
                if (dlam[i] < 0.0) ad = std::min(ad, -lam[i] / dlam[i]);
            }
            alpha = std::min(1.0, tau_k * ap); alpha_d = std::min(1.0, tau_k * ad);
        }
        if (m_on == 0) { alpha = 1.0; alpha_d = 1.0; }
        // ---- update
        for (int k = 0; k < N; ++k) {
            for (int i = 0; i < 6; ++i) { if (k == 0 && i >= 2) continue; z[(size_t)k * 6 + i] += alpha * dz[(size_t)k * 6 + i]; }
            for (int i = 0; i < 4; ++i) pi[(size_t)k * 4 + i] += alpha_d * dpi[(size_t)k * 4 + i];
        }
        for (int i = 0; i < 4; ++i) xN[i] += alpha * dxN[i];
        for (size_t i = 0; i < nc; ++i) if (on[i]) { lam[i] += alpha_d * dlam[i]; t[i] += alpha * dt_[i]; }
    }
    sol.iters = it;
    sol.du.resize((size_t)N * 2); sol.dx.resize((size_t)(N + 1) * 4);
    for (int k = 0; k < N; ++k) {
        sol.du[(size_t)k * 2] = z[(size_t)k * 6]; sol.du[(size_t)k * 2 + 1] = z[(size_t)k * 6 + 1];
        for (int i = 0; i < 4; ++i) sol.dx[(size_t)k * 4 + i] = z[(size_t)k * 6 + 2 + i];
    }
    for (int i = 0; i < 4; ++i) sol.dx[(size_t)N * 4 + i] = xN[i];
    sol.pi = pi; sol.lam = lam; sol.t = t;
    for (size_t i = 0; i < nc; ++i) if (!on[i]) { sol.lam[i] = 0.0; sol.t[i] = 0.0; }
}

// =========================================================================================
// 6. Gauss-Newton linearisation  (acados cost_ls + constraints_bgh + dynamics)
// =========================================================================================
struct RefData {            // per problem
    const double* x0bar;    // 4
    const double* yref;     // N x 6  ([x_ref; u_ref] per stage)
    const double* yref_e;   // 4
};

// h = [s; u_n; u_t] (NMPC_controller.m:237) or, h_variant 1, [u_n; u_t - v_bound(s); u_t + v_bound(s)] (:238).
// beta (optional) = d h_c / d s.
static double h_of(const Ocp& ocp, const Traj& tr, int k, int c, bool local, double* beta = nullptr) {
    if (beta) *beta = 0.0;
    if (ocp.h_variant == 0) return c == 0 ? tr.x[(size_t)k * 4 + 3] : tr.u[(size_t)k * 2 + (c - 1)];
    if (c == 0) return tr.u[(size_t)k * 2];
    double dv;
    const double vb = v_bound_sym(*ocp.model, ocp, tr.x[(size_t)k * 4 + 3], local, &dv);
    if (beta) *beta = (c == 1) ? -dv : dv;
    return tr.u[(size_t)k * 2 + 1] + ((c == 1) ? -vb : vb);
}

static void linearise(const Ocp& ocp, const RefData& rd, const Traj& tr, bool local, QP& qp) {
    qp.infeasible = false;
    const int N = ocp.N; const double dt = ocp.dt;
    const double cs = (ocp.opts.sem_cost_scale == 1) ? 1.0 : dt, ce = (ocp.opts.sem_cost_scale == 2) ? dt : 1.0;   // stage / terminal cost scaling
    qp.N = N; qp.st.resize(N);
    for (int k = 0; k < N; ++k) {
        StageQP& s = qp.st[k];
        const double* xk = &tr.x[(size_t)k * 4]; const double* uk = &tr.u[(size_t)k * 2];
        double Phi[4], A[16], B[8];
        erk4_sens_steps(*ocp.model, xk, uk, dt, ocp.opts.sem_erk_steps, local, Phi, A, B);
        for (int i = 0; i < 4; ++i) {
            for (int j = 0; j < 4; ++j) s.A[i][j] = A[4 * i + j];
            s.B[i][0] = B[2 * i]; s.B[i][1] = B[2 * i + 1];
            s.b[i] = Phi[i] - tr.x[(size_t)(k + 1) * 4 + i];
        }
        // y = [x;u]; H_y = dt*W ; permute to z = [u;x]
        const double* W = &ocp.W[(size_t)k * 36];
        double y[6], r[6];
        for (int i = 0; i < 4; ++i) y[i] = xk[i];
        y[4] = uk[0]; y[5] = uk[1];
        for (int i = 0; i < 6; ++i) r[i] = y[i] - rd.yref[(size_t)k * 6 + i];
        auto perm = [](int zi) { return zi < 2 ? 4 + zi : zi - 2; };   // z index -> y index
        for (int i = 0; i < 6; ++i) {
            double a = 0.0;
            for (int j = 0; j < 6; ++j) {
                s.H[i][j] = cs * W[perm(i) + 6 * perm(j)];
                a += cs * W[perm(i) + 6 * j] * r[j];
            }
            s.g[i] = a;
        }
        for (int c = 0; c < 3; ++c) {
            double h = h_of(ocp, tr, k, c, local, &s.beta[c]);
            s.dl[c] = ocp.lh[c] - h; s.du[c] = ocp.uh[c] - h;
            s.on[c] = ocp.h_variant ? true : !(k == 0 && c == 0);
            if (ocp.opts.sem_h0_s_row && !ocp.h_variant && k == 0 && c == 0) {     // the row is a constant (x_0 is fixed): it can only be (in)feasible
                const double v = rd.x0bar[3];
                if (v < ocp.lh[0] - 1e-12 || v > ocp.uh[0] + 1e-12) qp.infeasible = true;
            }
            s.ci[c] = ocp.h_variant ? (c == 0 ? 0 : 1) : CIDX[c];
        }
    }
    for (int i = 0; i < 4; ++i) {
        double a = 0.0;
        for (int j = 0; j < 4; ++j) {
            qp.QN[i][j] = ce * ocp.We[i + 4 * j];
            a += ce * ocp.We[i + 4 * j] * (tr.x[(size_t)N * 4 + j] - rd.yref_e[j]);
        }
        qp.qN[i] = a;
        qp.dx0[i] = rd.x0bar[i] - tr.x[i];
    }
}

// total cost at the iterate (acados get_cost): sum dt*0.5*||y-yref||_W^2 + 0.5*||xN - yref_e||_We^2
static double eval_cost(const Ocp& ocp, const RefData& rd, const Traj& tr) {
    const int N = ocp.N; double c = 0.0;
    for (int k = 0; k < N; ++k) {
        const double* W = &ocp.W[(size_t)k * 36];
        double r[6];
        for (int i = 0; i < 4; ++i) r[i] = tr.x[(size_t)k * 4 + i] - rd.yref[(size_t)k * 6 + i];
        for (int i = 0; i < 2; ++i) r[4 + i] = tr.u[(size_t)k * 2 + i] - rd.yref[(size_t)k * 6 + 4 + i];
        double q = 0.0;
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) q += r[i] * W[i + 6 * j] * r[j];
        c += 0.5 * ((ocp.opts.sem_cost_scale == 1) ? 1.0 : ocp.dt) * q;
    }
    double q = 0.0;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j)
        q += (tr.x[(size_t)N * 4 + i] - rd.yref_e[i]) * ocp.We[i + 4 * j] * (tr.x[(size_t)N * 4 + j] - rd.yref_e[j]);
    return c + 0.5 * ((ocp.opts.sem_cost_scale == 2) ? ocp.dt : 1.0) * q;
}

// =========================================================================================
// 7. SQP-RTI step and full SQP with L1-merit backtracking  
// =========================================================================================
static void apply_step(const Ocp& ocp, const QPSol& qs, double alpha, Traj& tr) {
    const int N = ocp.N;
    for (size_t i = 0; i < (size_t)(N + 1) * 4; ++i) tr.x[i] += alpha * qs.dx[i];
    for (size_t i = 0; i < (size_t)N * 2; ++i) tr.u[i] += alpha * qs.du[i];
    const double ad = ocp.opts.sem_full_step_dual ? 1.0 : alpha;
    for (size_t i = 0; i < (size_t)N * 4; ++i) tr.pi[i] = (1.0 - ad) * tr.pi[i] + ad * qs.pi[i];
    for (size_t i = 0; i < (size_t)N * 6; ++i) tr.lam[i] = (1.0 - ad) * tr.lam[i] + ad * qs.lam[i];
}

// One real-time iteration: linearise at (x,u), solve the QP, take the full step,
// multipliers replaced by the QP's (acados sqp_rti).
static void rti_step(const Ocp& ocp, const RefData& rd, Traj& tr, bool local, SolveStats& st) {
    QP qp; QPSol qs;
    linearise(ocp, rd, tr, local, qp);
    qp_solve_ipm(qp, ocp.opts, qs);
    apply_step(ocp, qs, 1.0, tr);
    st.sqp_iter = 1; st.qp_iter = qs.iters;
    st.status = (qs.status == 0) ? 0 : (qs.status == 1 ? (ocp.opts.sem_qp_maxiter_fails ? 4 : 0) : 4);   // QP max-iter is tolerated (A2.4)
    if (qp.infeasible) st.status = 4;
    for (size_t i = 0; i < tr.u.size(); ++i) if (!(tr.u[i] == tr.u[i])) st.status = 1;
    st.cost = eval_cost(ocp, rd, tr);
    for (int i = 0; i < 4; ++i) st.res[i] = qs.res[i];
    st.alpha_last = 1.0;
}

// NLP residuals at the iterate with the linearisation in qp (inf-norms).
static void nlp_residuals(const Ocp& ocp, const QP& qp, const Traj& tr, double res[4]) {
    const int N = ocp.N;
    double r_stat = 0.0, r_eq = 0.0, r_in = 0.0, r_cp = 0.0;
    for (int i = 0; i < 4; ++i) r_eq = std::max(r_eq, std::fabs(qp.dx0[i]));
    for (int k = 0; k < N; ++k) {
        const StageQP& s = qp.st[k];
        double g[6];
        for (int i = 0; i < 6; ++i) g[i] = s.g[i];
        for (int j = 0; j < 2; ++j) for (int l = 0; l < 4; ++l) g[j] += s.B[l][j] * tr.pi[(size_t)k * 4 + l];
        for (int j = 0; j < 4; ++j) for (int l = 0; l < 4; ++l) g[2 + j] += s.A[l][j] * tr.pi[(size_t)k * 4 + l];
        if (k > 0) for (int j = 0; j < 4; ++j) g[2 + j] -= tr.pi[(size_t)(k - 1) * 4 + j];
        for (int c = 0; c < 3; ++c) if (s.on[c]) {
            const double dlm = -tr.lam[(size_t)k * 6 + c] + tr.lam[(size_t)k * 6 + 3 + c];
            g[s.ci[c]] += dlm; g[5] += s.beta[c] * dlm;
            double sl = -s.dl[c], su = s.du[c];   // slacks h-lh, uh-h
            r_in = std::max(r_in, std::max(-sl, 0.0)); r_in = std::max(r_in, std::max(-su, 0.0));
            r_cp = std::max(r_cp, std::fabs(tr.lam[(size_t)k * 6 + c] * sl));
            r_cp = std::max(r_cp, std::fabs(tr.lam[(size_t)k * 6 + 3 + c] * su));
        }
        for (int i = 0; i < 6; ++i) { if (k == 0 && i >= 2) continue; r_stat = std::max(r_stat, std::fabs(g[i])); }
        for (int i = 0; i < 4; ++i) r_eq = std::max(r_eq, std::fabs(s.b[i]));
    }
    for (int i = 0; i < 4; ++i) r_stat = std::max(r_stat, std::fabs(qp.qN[i] - tr.pi[(size_t)(N - 1) * 4 + i]));
    res[0] = r_stat; res[1] = r_eq; res[2] = r_in; res[3] = r_cp;
}

// L1 merit: cost + sum w_pi |dyn residual| + sum w_lam * violation  (incl. the x0 constraint)
static double merit(const Ocp& ocp, const RefData& rd, const Traj& tr, bool local,
                    const std::vector<double>& wpi, const std::vector<double>& wlam, const double wx0[4]) {
    const int N = ocp.N;
    double mval = eval_cost(ocp, rd, tr);
    for (int i = 0; i < 4; ++i) mval += wx0[i] * std::fabs(rd.x0bar[i] - tr.x[i]);
    for (int k = 0; k < N; ++k) {
        double Phi[4];
        if (ocp.opts.sem_erk_steps <= 1) erk4(*ocp.model, &tr.x[(size_t)k * 4], &tr.u[(size_t)k * 2], ocp.dt, local, Phi);
        else { double A_[16], B_[8]; erk4_sens_steps(*ocp.model, &tr.x[(size_t)k * 4], &tr.u[(size_t)k * 2], ocp.dt, ocp.opts.sem_erk_steps, local, Phi, A_, B_); }
        for (int i = 0; i < 4; ++i) mval += wpi[(size_t)k * 4 + i] * std::fabs(Phi[i] - tr.x[(size_t)(k + 1) * 4 + i]);
        for (int c = 0; c < 3; ++c) {
            if (ocp.h_variant == 0 && k == 0 && c == 0) continue;
            double h = h_of(ocp, tr, k, c, local);
            mval += wlam[(size_t)k * 6 + c] * std::max(0.0, ocp.lh[c] - h);
            mval += wlam[(size_t)k * 6 + 3 + c] * std::max(0.0, h - ocp.uh[c]);
        }
    }
    return mval;
}

static void sqp_solve(const Ocp& ocp, const RefData& rd, Traj& tr, bool local, SolveStats& st) {
    const int N = ocp.N; const OcpOpts& o = ocp.opts;
    std::vector<double> wpi((size_t)N * 4, 0.0), wlam((size_t)N * 6, 0.0);
    double wx0[4] = {0, 0, 0, 0};
    st = SolveStats();
    QP qp; QPSol qs;
    int it = 0;
    for (;; ++it) {
        linearise(ocp, rd, tr, local, qp);
        nlp_residuals(ocp, qp, tr, st.res);
        bool nan = false;
        for (int i = 0; i < 4; ++i) if (!(st.res[i] == st.res[i])) nan = true;
        if (nan) { st.status = 1; break; }
        if (st.res[0] < o.tol_stat && st.res[1] < o.tol_eq && st.res[2] < o.tol_ineq && st.res[3] < o.tol_comp) { st.status = 0; break; }
        if (it >= o.max_sqp_iter) { st.status = 2; break; }
        qp_solve_ipm(qp, o, qs);
        st.qp_iter += qs.iters;
        if (qs.status == 2 || qp.infeasible || (qs.status == 1 && o.sem_qp_maxiter_fails)) { st.status = 4; ++it; break; }
        double alpha = 1.0;
        if (o.globalization == 1) {
            // merit weights: |multipliers_qp| at the first iteration, then max(|m|, (w+|m|)/2)
            auto wupd = [&](double w, double a) { return (it == 0 || o.sem_merit_weights == 1) ? a : (o.sem_merit_weights == 2 ? std::max(w, a) : std::max(a, 0.5 * (w + a))); };
            for (size_t i = 0; i < wpi.size(); ++i) wpi[i] = wupd(wpi[i], std::fabs(qs.pi[i]));
            for (size_t i = 0; i < wlam.size(); ++i) wlam[i] = wupd(wlam[i], std::fabs(qs.lam[i]));
            // multiplier of the x0 equality: costate at stage 0 (from stage-0 stationarity)
            {
                const StageQP& s = qp.st[0];
                for (int j = 0; j < 4; ++j) {
                    double a = s.g[2 + j];
                    for (int l = 0; l < 6; ++l) a += s.H[2 + j][l] * (l < 2 ? qs.du[l] : qs.dx[l - 2]);
                    for (int l = 0; l < 4; ++l) a += s.A[l][j] * qs.pi[l];
                    if (j == 3) for (int c = 0; c < 3; ++c) a += s.beta[c] * (qs.lam[3 + c] - qs.lam[c]);
                    a = std::fabs(a);
                    wx0[j] = wupd(wx0[j], a);
                }
            }
            const double m0 = merit(ocp, rd, tr, local, wpi, wlam, wx0);
            // directional derivative of the L1 merit along the QP step
            double dcost = 0.0;
            for (int k = 0; k < N; ++k) {
                const StageQP& s = qp.st[k];
                for (int i = 0; i < 6; ++i) dcost += s.g[i] * (i < 2 ? qs.du[(size_t)k * 2 + i] : qs.dx[(size_t)k * 4 + i - 2]);
            }
            for (int i = 0; i < 4; ++i) dcost += qp.qN[i] * qs.dx[(size_t)N * 4 + i];
            double dinf = 0.0;
            for (int i = 0; i < 4; ++i) dinf += wx0[i] * std::fabs(qp.dx0[i]);
            for (int k = 0; k < N; ++k) {
                for (int i = 0; i < 4; ++i) dinf += wpi[(size_t)k * 4 + i] * std::fabs(qp.st[k].b[i]);
                for (int c = 0; c < 3; ++c) if (qp.st[k].on[c]) {
                    dinf += wlam[(size_t)k * 6 + c] * std::max(0.0, qp.st[k].dl[c]);
                    dinf += wlam[(size_t)k * 6 + 3 + c] * std::max(0.0, -qp.st[k].du[c]);
                }
            }
            const double dmerit = dcost - dinf;
            Traj trial;
            for (;;) {
                trial = tr;
                for (size_t i = 0; i < trial.x.size(); ++i) trial.x[i] += alpha * qs.dx[i];
                for (size_t i = 0; i < trial.u.size(); ++i) trial.u[i] += alpha * qs.du[i];
                const double m1 = merit(ocp, rd, trial, local, wpi, wlam, wx0);
                if (o.sem_armijo ? (m1 < m0) : (m1 <= m0 + o.eps_sufficient_descent * alpha * dmerit)) break;
                if (!(m1 == m1) && alpha <= o.alpha_min) break;
                alpha *= o.alpha_reduction;
                if (alpha < o.alpha_min) { alpha = o.alpha_min; break; }
            }
        }
        apply_step(ocp, qs, alpha, tr);
        st.alpha_last = alpha;
    }
    st.sqp_iter = it;
    st.cost = eval_cost(ocp, rd, tr);
}

// =========================================================================================
// 8. NMPC_controller.solve pre/post-processing  (NMPC_controller.m:319-423)
// =========================================================================================
struct CtrlParams {
    double v_alpha = 1.0;      // NMPC_controller.m:98
    double d_v_bound = 0.0;    // :99
    double t_angle0 = 3.0;     // :100
    double u_t_ub = 0.05;      // :24
    double u_n_lb = 0.0;       // :25
};

// update_tangential_velocity_bounds, NMPC_controller.m:319-327
static double v_bound(const Model& m, const CtrlParams& cp, double s, bool local, double* t_angle_out) {
    double sm = matlab_mod(m, s);                          // :320
    sm = matlab_mod(m, sm);                                // getAngleCurvatures applies mod again, bspline_shape.m:147
    double t_angle = std::fabs(angle_dot(m, sm, local));   // :321
    if (t_angle_out) *t_angle_out = t_angle;
    return std::fmin(cp.v_alpha / (std::fabs(t_angle - cp.t_angle0) + 0.0001) + cp.d_v_bound, cp.u_t_ub);  // :322 (MATLAB min ignores NaN, like fmin)
}

static double sgn(double v) { return (v > 0.0) - (v < 0.0); }

// Steps 1, 4-6 of SURVEY A3: x0 wrap, cold start, v_bound clipping, Euler rollout.
// `cold` = 1 means utraj/xtraj are empty (first call after initial_condition_update).
static void controller_prepare(const Ocp& ocp, const CtrlParams& cp, double x0[4], int cold, Traj& tr, bool local) {
    const Model& m = *ocp.model; const int N = ocp.N;
    {   // :332  x0(4) = mod(x0(4), b) - b*(x0(4) < 0)
        double w = matlab_mod(m, x0[3]);
        if (m.single_quirk) w = (double)((float)w - (float)m.b * (x0[3] < 0 ? 1.f : 0.f));
        else w = w - m.b * (x0[3] < 0 ? 1.0 : 0.0);
        x0[3] = w;
    }
    if (cold) {                                             // :351-355
        tr.x.assign((size_t)(N + 1) * 4, 0.0);
        tr.u.assign((size_t)N * 2, 0.0);
        for (int k = 0; k < N; ++k) { tr.u[(size_t)k * 2] = cp.u_n_lb; tr.u[(size_t)k * 2 + 1] = 0.0; }
        tr.pi.assign((size_t)N * 4, 0.0);
        tr.lam.assign((size_t)N * 6, 0.0);
    }
    auto clip = [&](int k, double vb) {                     // :358-364, :375-379
        double* uk = &tr.u[(size_t)k * 2];
        if (std::fabs(uk[1]) > vb) {
            double ut_old = uk[1];
            uk[1] = sgn(ut_old) * vb;
            uk[0] = uk[1] * uk[0] / ut_old;
        }
    };
    clip(0, v_bound(m, cp, x0[3], local, nullptr));         // :357
    for (int i = 0; i < 4; ++i) tr.x[i] = x0[i];            // :366
    for (int j = 1; j <= N; ++j) {                          // :367-380
        double f[4];
        dynamics_num(m, &tr.x[(size_t)(j - 1) * 4], &tr.u[(size_t)(j - 1) * 2], local, f);
        for (int i = 0; i < 4; ++i) tr.x[(size_t)j * 4 + i] = tr.x[(size_t)(j - 1) * 4 + i] + ocp.dt * f[i];
        double vb = v_bound(m, cp, tr.x[(size_t)j * 4 + 3], local, nullptr);
        if (j == N) break;
        clip(j, vb);
    }
}

// Step 8 of SURVEY A3: shift left by one, duplicate the last column (:397-399).
static void controller_shift(const Ocp& ocp, Traj& tr) {
    const int N = ocp.N;
    for (int k = 0; k + 1 < N; ++k) for (int i = 0; i < 2; ++i) tr.u[(size_t)k * 2 + i] = tr.u[(size_t)(k + 1) * 2 + i];
    for (int k = 0; k < N; ++k) for (int i = 0; i < 4; ++i) tr.x[(size_t)k * 4 + i] = tr.x[(size_t)(k + 1) * 4 + i];
    for (int k = 0; k + 1 < N; ++k) for (int i = 0; i < 4; ++i) tr.pi[(size_t)k * 4 + i] = tr.pi[(size_t)(k + 1) * 4 + i];
    for (int k = 0; k + 1 < N; ++k) for (int i = 0; i < 6; ++i) tr.lam[(size_t)k * 6 + i] = tr.lam[(size_t)(k + 1) * 6 + i];
}

}  // namespace orc

// =========================================================================================
// C interface (ctypes).  All arrays are caller-owned, per-problem contiguous, column-major like
// the MATLAB arrays they mirror: x [nb][N+1][4], u [nb][N][2], pi [nb][N][4], lam [nb][N][6],
// yref [nb][N][6], yref_e [nb][4], x0bar [nb][4].
// =========================================================================================
using namespace orc;

struct OrcHandle { Model model; };
struct OrcOcp { Ocp ocp; CtrlParams cp; bool local = true; };

template <class F>
static void parallel_for(int n, int nthreads, F fn) {
    if (nthreads <= 1 || n <= 1) { for (int i = 0; i < n; ++i) fn(i); return; }
    nthreads = std::min(nthreads, n);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([=]() { for (int i = t; i < n; i += nthreads) fn(i); });
    for (auto& x : th) x.join();
}

extern "C" {

void* orc_model_create(const double* S, int nknots, const double* P, int n, int p,
                       double mu_sp, double c_ellipse, int single_quirk) {
    if (nknots != n + p + 1 || n < p + 1) return nullptr;
    OrcHandle* h = new OrcHandle();
    Model& m = h->model;
    m.p = p; m.n = n; m.S.assign(S, S + nknots); m.P.assign(P, P + 2 * n);
    m.mu_sp = mu_sp; m.c_ellipse = c_ellipse; m.single_quirk = single_quirk != 0;
    m.b = S[nknots - 1];
    finish_model(m);
    return h;
}

// PusherSliderModel ctor path: .ply -> sortCadPoints -> getSpline (PusherSliderModel.m:45-60, 84-132)
void* orc_model_from_ply(const char* path, int flip, int p, double mu_sg, double mu_sp, double mass, double tau_max) {
    std::vector<float> xy;
    if (!ply_read_xy(path, xy)) return nullptr;
    std::vector<float> P = sort_cad_points(xy, flip != 0);
    std::vector<float> S; float b;
    make_knots(P, p, S, b);
    OrcHandle* h = new OrcHandle();
    Model& m = h->model;
    m.p = p; m.n = (int)P.size() / 2;
    m.S.assign(S.begin(), S.end()); m.P.assign(P.begin(), P.end());
    m.b = (double)b;
    m.mu_sp = mu_sp;
    const double f_max = mu_sg * mass * 9.81;              // PusherSliderModel.m:53, helper.m:3
    m.c_ellipse = tau_max / f_max;                         // :55
    m.single_quirk = true;
    finish_model(m);
    return h;
}

void orc_model_free(void* h) { delete (OrcHandle*)h; }

void orc_model_info(void* h, int* n, int* nknots, double* b, double* c_ellipse, double* mu_sp) {
    Model& m = ((OrcHandle*)h)->model;
    *n = m.n; *nknots = (int)m.S.size(); *b = m.b; *c_ellipse = m.c_ellipse; *mu_sp = m.mu_sp;
}
void orc_model_tables(void* h, double* S, double* P, double* c1, double* c2) {
    Model& m = ((OrcHandle*)h)->model;
    if (S) std::memcpy(S, m.S.data(), m.S.size() * sizeof(double));
    if (P) std::memcpy(P, m.P.data(), m.P.size() * sizeof(double));
    if (c1) std::memcpy(c1, m.c1.data(), m.c1.size() * sizeof(double));
    if (c2) std::memcpy(c2, m.c2.data(), m.c2.size() * sizeof(double));
}

double orc_basis(void* h, double s, int i, int ord) { return basis_rec<double>(((OrcHandle*)h)->model, s, i, ord); }

// wrap: 0 = evaluate at s as given; 1 = MATLAB mod(s,b) (evalSpline, bspline_shape.m:193);
//       2 = fmod(s,b)+(s<0)*b (dynamics, PusherSliderModel.m:526)
// Outputs (each cnt x 2 unless noted, nullable): C, Cd, Cdd, t, n, kappa (cnt)
void orc_eval_spline(void* h, int cnt, const double* s, int wrap, int local,
                     double* C, double* Cd, double* Cdd, double* tv, double* nv, double* kappa) {
    const Model& m = ((OrcHandle*)h)->model;
    for (int i = 0; i < cnt; ++i) {
        double si = s[i];
        if (wrap == 1) si = matlab_mod(m, si);
        else if (wrap == 2) si = std::fmod(si, m.b) + (si < 0 ? 1.0 : 0.0) * m.b;
        double o[2];
        if (C) { FC<double>(m, si, local, o); C[2 * i] = o[0]; C[2 * i + 1] = o[1]; }
        if (Cd) { FC_dot<double>(m, si, local, o); Cd[2 * i] = o[0]; Cd[2 * i + 1] = o[1]; }
        if (Cdd) { FC_dot_dot<double>(m, si, local, o); Cdd[2 * i] = o[0]; Cdd[2 * i + 1] = o[1]; }
        if (tv || nv) {
            double t2[2], n2[2]; frames<double>(m, si, local, t2, n2);
            if (tv) { tv[2 * i] = t2[0]; tv[2 * i + 1] = t2[1]; }
            if (nv) { nv[2 * i] = n2[0]; nv[2 * i + 1] = n2[1]; }
        }
        if (kappa) kappa[i] = angle_dot(m, si, local);
    }
}

// getCurvatures, bspline_shape.m:154-179 (seam-blended |C''|)
void orc_get_curvatures(void* h, int cnt, const double* s, double* out) {
    const Model& m = ((OrcHandle*)h)->model;
    auto nrm_dd = [&](double sv) { double w = matlab_mod(m, sv); double o[2]; FC_dot_dot<double>(m, w, true, o); return std::sqrt(o[0] * o[0] + o[1] * o[1]); };
    const double d01 = 0.011, d0n = 0.011, a = 0.0;
    const double s1 = a + d01, s0 = a - d0n, sn = m.b + d01, sn_1 = m.b - d0n;
    for (int i = 0; i < cnt; ++i) {
        double sv = matlab_mod(m, s[i]);
        if (sv <= s1 && sv >= s0) { double y1 = nrm_dd(s1), y0 = nrm_dd(s0); out[i] = (y1 - y0) * (sv - s0) / (s1 - s0) + y0; }
        else if (sv <= sn && sv >= sn_1) { double yn = nrm_dd(sn), yn1 = nrm_dd(sn_1); out[i] = (yn - yn1) * (sv - sn_1) / (sn - sn_1) + yn1; }
        else out[i] = nrm_dd(sv);
    }
}

// f (cnt x 4), Jx (cnt x 16 row-major), Ju (cnt x 8 row-major); Jx/Ju nullable
void orc_dynamics(void* h, int cnt, const double* x, const double* u, int local, double* f, double* Jx, double* Ju) {
    const Model& m = ((OrcHandle*)h)->model;
    for (int i = 0; i < cnt; ++i) {
        if (Jx || Ju) {
            double jx[16], ju[8];
            dynamics_jac(m, x + 4 * i, u + 2 * i, local, f + 4 * i, jx, ju);
            if (Jx) std::memcpy(Jx + 16 * i, jx, sizeof jx);
            if (Ju) std::memcpy(Ju + 8 * i, ju, sizeof ju);
        } else dynamics_num(m, x + 4 * i, u + 2 * i, local, f + 4 * i);
    }
}

void orc_erk4_sens(void* h, int cnt, const double* x, const double* u, double dt, int local, int nthreads,
                   double* Phi, double* A, double* B) {
    const Model& m = ((OrcHandle*)h)->model;
    parallel_for(cnt, nthreads, [&](int i) { erk4_sens(m, x + 4 * i, u + 2 * i, dt, local, Phi + 4 * i, A + 16 * i, B + 8 * i); });
}

double orc_v_bound(void* h, double s, double v_alpha, double d_v_bound, double t_angle0, double u_t_ub, double* t_angle) {
    CtrlParams cp; cp.v_alpha = v_alpha; cp.d_v_bound = d_v_bound; cp.t_angle0 = t_angle0; cp.u_t_ub = u_t_ub;
    return v_bound(((OrcHandle*)h)->model, cp, s, true, t_angle);
}

// ---- OCP handle
void* orc_ocp_create(void* model, int N, double dt) {
    OrcOcp* o = new OrcOcp();
    o->ocp.model = &((OrcHandle*)model)->model; o->ocp.N = N; o->ocp.dt = dt;
    o->ocp.W.assign((size_t)N * 36, 0.0);
    // defaults of NMPC_controller.m:16-18, 251-252
    const double wx[4] = {1.0, 1.0, 1e-3, 0.0}, wu[2] = {1e-3, 1e-3}, we[4] = {2e5, 2e5, 20.0, 0.0};
    for (int k = 0; k < N; ++k) { for (int i = 0; i < 4; ++i) o->ocp.W[(size_t)k * 36 + 7 * i] = wx[i]; for (int i = 0; i < 2; ++i) o->ocp.W[(size_t)k * 36 + 7 * (4 + i)] = wu[i]; }
    std::memset(o->ocp.We, 0, sizeof o->ocp.We);
    for (int i = 0; i < 4; ++i) o->ocp.We[5 * i] = we[i];
    const double lh[3] = {-0.06, 0.0, -0.05}, uh[3] = {0.011, 0.03, 0.05};
    for (int i = 0; i < 3; ++i) { o->ocp.lh[i] = lh[i]; o->ocp.uh[i] = uh[i]; }
    return o;
}
void orc_ocp_free(void* o) { delete (OrcOcp*)o; }
void orc_ocp_set_W(void* o_, int stage, const double* W) {   // stage == N -> terminal 4x4 (NMPC_controller.m:154-157)
    OrcOcp* o = (OrcOcp*)o_;
    if (stage == o->ocp.N) std::memcpy(o->ocp.We, W, 16 * sizeof(double));
    else std::memcpy(&o->ocp.W[(size_t)stage * 36], W, 36 * sizeof(double));
}
void orc_ocp_set_bounds(void* o_, const double* lh, const double* uh) {
    OrcOcp* o = (OrcOcp*)o_;
    for (int i = 0; i < 3; ++i) { o->ocp.lh[i] = lh[i]; o->ocp.uh[i] = uh[i]; }
}
// opts: [max_sqp_iter, tol_stat, tol_eq, tol_ineq, tol_comp, qp_max_iter, qp_tol, qp_mu0, qp_thr, qp_tau,
//        alpha_min, alpha_reduction, eps_sufficient_descent, globalization, local_spline,
//        qp_tol_comp, qp_t_min, qp_gamma_f, qp_stall,
//        sem_cost_scale, sem_h0_s_row, sem_full_step_dual, sem_merit_weights, sem_armijo, sem_erk_steps, sem_qp_maxiter_fails, sem_mod_strict,
//        sem_qp_pivot_fails, qp_split_step]
void orc_ocp_set_opts(void* o_, const double* v) {
    OrcOcp* o = (OrcOcp*)o_; OcpOpts& p = o->ocp.opts;
    p.max_sqp_iter = (int)v[0]; p.tol_stat = v[1]; p.tol_eq = v[2]; p.tol_ineq = v[3]; p.tol_comp = v[4];
    p.qp_max_iter = (int)v[5]; p.qp_tol = v[6]; p.qp_mu0 = v[7]; p.qp_thr = v[8]; p.qp_tau = v[9];
    p.alpha_min = v[10]; p.alpha_reduction = v[11]; p.eps_sufficient_descent = v[12]; p.globalization = (int)v[13];
    o->local = v[14] != 0.0;
    p.qp_tol_comp = v[15]; p.qp_t_min = v[16]; p.qp_gamma_f = v[17]; p.qp_stall = (int)v[18];
    p.sem_cost_scale = (int)v[19]; p.sem_h0_s_row = (int)v[20]; p.sem_full_step_dual = (int)v[21]; p.sem_merit_weights = (int)v[22];
    p.sem_armijo = (int)v[23]; p.sem_erk_steps = (int)v[24]; p.sem_qp_maxiter_fails = (int)v[25]; p.sem_mod_strict = (int)v[26];
    p.sem_qp_pivot_fails = (int)v[27]; p.qp_split_step = (int)v[28];
    const_cast<Model*>(o->ocp.model)->mod_strict = p.sem_mod_strict != 0;
}
// h_variant 1: h = [u_n; u_t - v_bound(s); u_t + v_bound(s)]; the caller sets the matching lh / uh
// ([u_n_lb, 2 u_t_lb, 0] / [u_n_ub, 0, 2 u_t_ub], NMPC_controller.m:247-248) with orc_ocp_set_bounds.
void orc_ocp_set_h_variant(void* o_, int variant, double v_alpha, double d_v_bound, double t_angle0, double u_t_ub) {
    OrcOcp* o = (OrcOcp*)o_;
    o->ocp.h_variant = variant; o->ocp.vb_alpha = v_alpha; o->ocp.vb_d = d_v_bound; o->ocp.vb_t0 = t_angle0; o->ocp.vb_ub = u_t_ub;
}
// v_bound of the variant and its derivative (test hook)
double orc_v_bound_sym(void* o_, double s, double* dv) {
    OrcOcp* o = (OrcOcp*)o_;
    return v_bound_sym(*o->ocp.model, o->ocp, s, o->local, dv);
}
void orc_ocp_set_ctrl(void* o_, double v_alpha, double d_v_bound, double t_angle0, double u_t_ub, double u_n_lb) {
    OrcOcp* o = (OrcOcp*)o_;
    o->cp.v_alpha = v_alpha; o->cp.d_v_bound = d_v_bound; o->cp.t_angle0 = t_angle0; o->cp.u_t_ub = u_t_ub; o->cp.u_n_lb = u_n_lb;
}

static void load_traj(const OrcOcp* o, int b, const double* x, const double* u, const double* pi, const double* lam, Traj& tr) {
    const int N = o->ocp.N;
    tr.x.assign(x + (size_t)b * (N + 1) * 4, x + (size_t)(b + 1) * (N + 1) * 4);
    tr.u.assign(u + (size_t)b * N * 2, u + (size_t)(b + 1) * N * 2);
    if (pi) tr.pi.assign(pi + (size_t)b * N * 4, pi + (size_t)(b + 1) * N * 4); else tr.pi.assign((size_t)N * 4, 0.0);
    if (lam) tr.lam.assign(lam + (size_t)b * N * 6, lam + (size_t)(b + 1) * N * 6); else tr.lam.assign((size_t)N * 6, 0.0);
}
static void store_traj(const OrcOcp* o, int b, const Traj& tr, double* x, double* u, double* pi, double* lam) {
    const int N = o->ocp.N;
    std::memcpy(x + (size_t)b * (N + 1) * 4, tr.x.data(), tr.x.size() * sizeof(double));
    std::memcpy(u + (size_t)b * N * 2, tr.u.data(), tr.u.size() * sizeof(double));
    if (pi) std::memcpy(pi + (size_t)b * N * 4, tr.pi.data(), tr.pi.size() * sizeof(double));
    if (lam) std::memcpy(lam + (size_t)b * N * 6, tr.lam.data(), tr.lam.size() * sizeof(double));
}

// Linearisation only: fills A [nb][N][16], B [nb][N][8], bres [nb][N][4], g [nb][N][6] ([u;x]), qN [nb][4]
void orc_linearise_batch(void* o_, int nb, const double* x0bar, const double* yref, const double* yref_e,
                         const double* x, const double* u, double* A, double* B, double* bres, double* g, double* qN) {
    OrcOcp* o = (OrcOcp*)o_; const int N = o->ocp.N;
    for (int b = 0; b < nb; ++b) {
        Traj tr; load_traj(o, b, x, u, nullptr, nullptr, tr);
        RefData rd{x0bar + 4 * b, yref + (size_t)b * N * 6, yref_e + 4 * b};
        QP qp; linearise(o->ocp, rd, tr, o->local, qp);
        for (int k = 0; k < N; ++k) {
            for (int i = 0; i < 4; ++i) {
                for (int j = 0; j < 4; ++j) A[((size_t)b * N + k) * 16 + 4 * i + j] = qp.st[k].A[i][j];
                for (int j = 0; j < 2; ++j) B[((size_t)b * N + k) * 8 + 2 * i + j] = qp.st[k].B[i][j];
                bres[((size_t)b * N + k) * 4 + i] = qp.st[k].b[i];
            }
            for (int i = 0; i < 6; ++i) g[((size_t)b * N + k) * 6 + i] = qp.st[k].g[i];
        }
        for (int i = 0; i < 4; ++i) qN[4 * b + i] = qp.qN[i];
    }
}

// constraint function and its s-derivative at (x,u): h [nb][N][3], beta [nb][N][3] (beta = d h_c / d s beyond the
// selection entry; zero for h_variant 0)
void orc_constraints_batch(void* o_, int nb, const double* x, const double* u, double* h, double* beta) {
    OrcOcp* o = (OrcOcp*)o_; const int N = o->ocp.N;
    for (int b = 0; b < nb; ++b) {
        Traj tr; load_traj(o, b, x, u, nullptr, nullptr, tr);
        for (int k = 0; k < N; ++k) for (int c = 0; c < 3; ++c)
            h[((size_t)b * N + k) * 3 + c] = h_of(o->ocp, tr, k, c, o->local, &beta[((size_t)b * N + k) * 3 + c]);
    }
}

// QP only (linearise at (x,u) then IPM): returns the step and multipliers.
// du [nb][N][2], dx [nb][N+1][4], pi [nb][N][4], lam [nb][N][6], tslack [nb][N][6], iters [nb], status [nb], res [nb][4]
void orc_qp_batch(void* o_, int nb, const double* x0bar, const double* yref, const double* yref_e,
                  const double* x, const double* u, int nthreads,
                  double* du, double* dx, double* pi, double* lam, double* tsl, int* iters, int* status, double* res) {
    OrcOcp* o = (OrcOcp*)o_; const int N = o->ocp.N;
    parallel_for(nb, nthreads, [&](int b) {
        Traj tr; load_traj(o, b, x, u, nullptr, nullptr, tr);
        RefData rd{x0bar + 4 * b, yref + (size_t)b * N * 6, yref_e + 4 * b};
        QP qp; QPSol qs; linearise(o->ocp, rd, tr, o->local, qp);
        qp_solve_ipm(qp, o->ocp.opts, qs);
        std::memcpy(du + (size_t)b * N * 2, qs.du.data(), qs.du.size() * sizeof(double));
        std::memcpy(dx + (size_t)b * (N + 1) * 4, qs.dx.data(), qs.dx.size() * sizeof(double));
        std::memcpy(pi + (size_t)b * N * 4, qs.pi.data(), qs.pi.size() * sizeof(double));
        std::memcpy(lam + (size_t)b * N * 6, qs.lam.data(), qs.lam.size() * sizeof(double));
        if (tsl) std::memcpy(tsl + (size_t)b * N * 6, qs.t.data(), qs.t.size() * sizeof(double));
        iters[b] = qs.iters; status[b] = qs.status;
        if (res) for (int i = 0; i < 4; ++i) res[4 * b + i] = qs.res[i];
    });
}

// The QP itself (for the extended-precision arbiter, oracle/qs_arbiter.cpp): H [nb][N][36] row-major in z = [u;x] order,
// g [nb][N][6], A [nb][N][16], B [nb][N][8], b [nb][N][4], QN [16], qN [nb][4], dx0 [nb][4], dl / du / beta [nb][N][3],
// on / ci [nb][N][3].
void orc_qp_data_batch(void* o_, int nb, const double* x0bar, const double* yref, const double* yref_e,
                       const double* x, const double* u, double* H, double* g, double* A, double* B, double* bres,
                       double* QN, double* qN, double* dx0, double* dl, double* du, double* beta, int* on, int* ci) {
    OrcOcp* o = (OrcOcp*)o_; const int N = o->ocp.N;
    for (int b = 0; b < nb; ++b) {
        Traj tr; load_traj(o, b, x, u, nullptr, nullptr, tr);
        RefData rd{x0bar + 4 * b, yref + (size_t)b * N * 6, yref_e + 4 * b};
        QP qp; linearise(o->ocp, rd, tr, o->local, qp);
        for (int k = 0; k < N; ++k) {
            const size_t s = (size_t)b * N + k; const StageQP& st = qp.st[k];
            for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) H[s * 36 + 6 * i + j] = st.H[i][j]; g[s * 6 + i] = st.g[i]; }
            for (int i = 0; i < 4; ++i) {
                for (int j = 0; j < 4; ++j) A[s * 16 + 4 * i + j] = st.A[i][j];
                for (int j = 0; j < 2; ++j) B[s * 8 + 2 * i + j] = st.B[i][j];
                bres[s * 4 + i] = st.b[i];
            }
            for (int c = 0; c < 3; ++c) { dl[s * 3 + c] = st.dl[c]; du[s * 3 + c] = st.du[c]; beta[s * 3 + c] = st.beta[c]; on[s * 3 + c] = st.on[c]; ci[s * 3 + c] = st.ci[c]; }
        }
        for (int i = 0; i < 4; ++i) { for (int j = 0; j < 4; ++j) QN[4 * i + j] = qp.QN[i][j]; qN[4 * b + i] = qp.qN[i]; dx0[4 * b + i] = qp.dx0[i]; }
    }
}

// mode 0 = RTI (one iteration, full step), 1 = full SQP.  x,u,pi,lam are in/out.
// stats_i [nb][3] = status, sqp_iter, qp_iter ; stats_d [nb][6] = cost, res[4], alpha_last
void orc_solve_batch(void* o_, int mode, int nb, const double* x0bar, const double* yref, const double* yref_e,
                     double* x, double* u, double* pi, double* lam, int nthreads, int* stats_i, double* stats_d) {
    OrcOcp* o = (OrcOcp*)o_; const int N = o->ocp.N;
    parallel_for(nb, nthreads, [&](int b) {
        Traj tr; load_traj(o, b, x, u, pi, lam, tr);
        RefData rd{x0bar + 4 * b, yref + (size_t)b * N * 6, yref_e + 4 * b};
        SolveStats st;
        if (mode == 0) rti_step(o->ocp, rd, tr, o->local, st); else sqp_solve(o->ocp, rd, tr, o->local, st);
        store_traj(o, b, tr, x, u, pi, lam);
        if (stats_i) { stats_i[3 * b] = st.status; stats_i[3 * b + 1] = st.sqp_iter; stats_i[3 * b + 2] = st.qp_iter; }
        if (stats_d) { stats_d[6 * b] = st.cost; for (int i = 0; i < 4; ++i) stats_d[6 * b + 1 + i] = st.res[i]; stats_d[6 * b + 5] = st.alpha_last; }
    });
}

// NMPC_controller.solve pre-processing for a batch.  x0 [nb][4] in/out (wrapped),
// cold [nb] flags, x,u,pi,lam in/out.
void orc_prepare_batch(void* o_, int nb, double* x0, const int* cold, double* x, double* u, double* pi, double* lam, int nthreads) {
    OrcOcp* o = (OrcOcp*)o_;
    parallel_for(nb, nthreads, [&](int b) {
        Traj tr; load_traj(o, b, x, u, pi, lam, tr);
        controller_prepare(o->ocp, o->cp, x0 + 4 * b, cold ? cold[b] : 0, tr, o->local);
        store_traj(o, b, tr, x, u, pi, lam);
    });
}
void orc_shift_batch(void* o_, int nb, double* x, double* u, double* pi, double* lam) {
    OrcOcp* o = (OrcOcp*)o_;
    for (int b = 0; b < nb; ++b) { Traj tr; load_traj(o, b, x, u, pi, lam, tr); controller_shift(o->ocp, tr); store_traj(o, b, tr, x, u, pi, lam); }
}

double orc_cost(void* o_, const double* x0bar, const double* yref, const double* yref_e, const double* x, const double* u) {
    OrcOcp* o = (OrcOcp*)o_;
    Traj tr; load_traj(o, 0, x, u, nullptr, nullptr, tr);
    RefData rd{x0bar, yref, yref_e};
    return eval_cost(o->ocp, rd, tr);
}

// helper.closed_loop_matlab without noise/disturbance/delay (helper.m:195-322) around NMPC_controller.solve.
// yref_full: 6 x T column-major (the controller's y_ref), steps = number of control steps.
// Outputs: xs [steps+1][4], us [steps][2], status [steps], sqp_iters [steps], cost [steps]
void orc_closed_loop(void* o_, int mode, const double* x0_in, const double* yref_full, int T, int steps,
                     double* xs, double* us, int* status, int* sqp_iters, double* cost) {
    OrcOcp* o = (OrcOcp*)o_; const Ocp& ocp = o->ocp; const int N = ocp.N;
    Traj tr; int cold = 1;
    for (int i = 0; i < 4; ++i) xs[i] = x0_in[i];
    std::vector<double> yref((size_t)N * 6); double yref_e[4];
    for (int i = 0; i < steps; ++i) {
        double x0[4]; for (int j = 0; j < 4; ++j) x0[j] = xs[(size_t)i * 4 + j];
        const int index_time = i + 1;                                   // helper.m:248 (1-based, no delay)
        for (int k = 0; k < N; ++k) {                                   // NMPC_controller.m:343-346, 307-313
            int col = std::min(index_time + k, T) - 1;
            for (int j = 0; j < 6; ++j) yref[(size_t)k * 6 + j] = yref_full[(size_t)col * 6 + j];
        }
        for (int j = 0; j < 4; ++j) yref_e[j] = yref[(size_t)(N - 1) * 6 + j];   // :348
        controller_prepare(ocp, o->cp, x0, cold, tr, o->local); cold = 0;
        RefData rd{x0, yref.data(), yref_e};
        SolveStats st;
        if (mode == 0) rti_step(ocp, rd, tr, o->local, st); else sqp_solve(ocp, rd, tr, o->local, st);
        us[(size_t)i * 2] = tr.u[0]; us[(size_t)i * 2 + 1] = tr.u[1];   // :403
        status[i] = st.status; sqp_iters[i] = st.sqp_iter; cost[i] = st.cost;
        controller_shift(ocp, tr);                                      // :397-399
        double f[4];                                                    // helper.m:294, 307
        dynamics_num(*ocp.model, &xs[(size_t)i * 4], &us[(size_t)i * 2], o->local, f);
        for (int j = 0; j < 4; ++j) xs[(size_t)(i + 1) * 4 + j] = xs[(size_t)i * 4 + j] + ocp.dt * f[j];
    }
}

}  // extern "C"
