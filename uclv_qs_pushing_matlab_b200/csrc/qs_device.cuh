// qs_device.cuh — contact geometry, quasi-static dynamics and ERK4 forward sensitivities.
//
// Everything here is per-thread FP64 arithmetic on values held in registers; the only memory it
// touches is the per-object model blob (pp-form spline tables), which the kernels stage in shared
// memory with one bulk TMA copy per CTA.
//
// Reference behaviour being reproduced (citations into /root/reference/acados_nmpc/):
//   bspline_shape.m:40-144     Cox-de Boor curve FC, FC_dot, FC_dot_dot, frames, angle rate
//   PusherSliderModel.m:503-603  the three-mode limit-surface dynamics (one branch-free expression)
//   acados sim_erk + expl_vde_forw  ERK4 with forward sensitivities
//
// The functions are __host__ __device__ so tests/hostsim can run the identical arithmetic on the
// CPU of a GPU-less CI box.  The product never calls them on the host.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define QS_HD __host__ __device__ __forceinline__
#else
#define QS_HD inline
#endif

namespace qs {

// ---- model blob layout (doubles) --------------------------------------------------------------
// hdr: [0] nspan  [1] b  [2] nspan/b  [3] mu_sp  [4] c_ellipse^2  [5] c_ellipse  [6],[7] spare
// knots: span boundaries kn[0..nspan]
// coef : per span, stride COEF_STRIDE (odd -> conflict-light shared-memory gathers)
//        [0..3] C_x  [4..7] C_y   cubic in tau = sigma - kn[q]                (FC,      :74-83)
//        [8..10] C'_x [11..13] C'_y quadratic built from cj_1                  (FC_dot,  :85-104)
//        [14..15] C''_x [16..17] C''_y linear built from cj_2                  (FC_dot_dot, :118-135)
constexpr int MAXSPAN = 64;
constexpr int COEF_STRIDE = 19;
constexpr int HDR = 8;
constexpr int KNOT_SLOTS = 72;
constexpr int COEF_OFF = HDR + KNOT_SLOTS;
constexpr int MODEL_DOUBLES = COEF_OFF + MAXSPAN * COEF_STRIDE;   // 1296 doubles = 10368 B (16 B multiple)
constexpr int MAX_MODELS = 8;

// Half-open span lookup identical to the reference's degree-0 indicator (s<S(i+1))*(s>=S(i)),
// eval_bspline.m:12-15.  Interior knots are (nearly) uniform, so the guess is O(1) and the two
// fix-up loops make the result exact with respect to the stored knot values.
QS_HD int span_of(const double* __restrict__ M, double sg, bool& valid) {
    const int nspan = (int)M[0];
    const double* kn = M + HDR;
    valid = (sg >= kn[0]) && (sg < kn[nspan]);   // NaN -> false -> every basis function is 0
    if (!valid) return 0;
    int q = (int)(sg * M[2]);
    q = q < 0 ? 0 : (q > nspan - 1 ? nspan - 1 : q);
    while (q > 0 && sg < kn[q]) --q;
    while (q < nspan - 1 && sg >= kn[q + 1]) ++q;
    return q;
}

struct Curve {
    double cx, cy;      // FC(sigma)
    double gx, gy;      // d/dsigma of the FC polynomial (what CasADi's AD of FC yields)
    double dx, dy;      // FC_dot(sigma)
    double hx, hy;      // d/dsigma of the FC_dot polynomial (AD of FC_dot)
};

// C, C' and the AD derivatives of both at sigma (already wrapped).
QS_HD void curve_eval(const double* __restrict__ M, double sg, Curve& c) {
    bool valid;
    const int q = span_of(M, sg, valid);
    if (!valid) { c.cx = c.cy = c.gx = c.gy = c.dx = c.dy = c.hx = c.hy = 0.0; return; }
    const double tau = sg - M[HDR + q];
    const double* k = M + COEF_OFF + q * COEF_STRIDE;
    c.cx = fma(fma(fma(k[3], tau, k[2]), tau, k[1]), tau, k[0]);
    c.cy = fma(fma(fma(k[7], tau, k[6]), tau, k[5]), tau, k[4]);
    c.gx = fma(fma(3.0 * k[3], tau, 2.0 * k[2]), tau, k[1]);
    c.gy = fma(fma(3.0 * k[7], tau, 2.0 * k[6]), tau, k[5]);
    c.dx = fma(fma(k[10], tau, k[9]), tau, k[8]);
    c.dy = fma(fma(k[13], tau, k[12]), tau, k[11]);
    c.hx = fma(2.0 * k[10], tau, k[9]);
    c.hy = fma(2.0 * k[13], tau, k[12]);
}

// FC_dot_dot(sigma) from the cj_2 table.
QS_HD void curve_dd(const double* __restrict__ M, double sg, double& ex, double& ey) {
    bool valid;
    const int q = span_of(M, sg, valid);
    if (!valid) { ex = ey = 0.0; return; }
    const double tau = sg - M[HDR + q];
    const double* k = M + COEF_OFF + q * COEF_STRIDE;
    ex = fma(k[15], tau, k[14]);
    ey = fma(k[17], tau, k[16]);
}

// MATLAB mod(s,b), b>0, result in [0,b)  (bspline_shape.m:147,155,193; NMPC_controller.m:320,332).
// With `single` the operation happens in float32, as MATLAB does when b is a `single`.
// The builtin's algorithm (the form MATLAB Coder emits for mod on floating-point operands): r = fmod(x, y), forced to 0 when
// x / y is an integer within eps |x / y|, otherwise r += y when the signs differ.  The result can EQUAL y: for a tiny negative
// x the sum rounds to y (in single precision for |x| < 1.5e-8 at b = 0.28), and x0(4) = mod(x0(4), b) - b (x0(4) < 0) is 0.
QS_HD double matlab_mod(double s, double b, bool single) {
    // Fast path, bit-identical to the algorithm below: for 0 < |x| < y / 2 fmod returns x itself, x / y is further than eps from
    // every integer (the quotient cannot underflow above the 1e-30 / 1e-290 guard), so r = x (+ y when x < 0).  Every contact coordinate the controller meets lies there (|s| < 0.14 at b >= 0.22).
    if (single) {
        const float x = (float)s, y = (float)b;
        { const float ax = fabsf(x); if (ax > 1e-30f && ax < 0.5f * y) return (double)(x < 0.f ? x + y : x); }
        if (y == 0.f) return (double)x;
        if (!(x == x) || !(y == y) || fabsf(x) == (float)INFINITY) return (double)NAN;
        if (x == 0.f) return (double)(0.f / y);
        float r = fmodf(x, y);
        bool req0 = (r == 0.f);
        if (!req0 && y > floorf(y)) { const float q = fabsf(x / y); req0 = !(fabsf(q - floorf(q + 0.5f)) > 1.1920929e-7f * q); }
        if (req0) r = y * 0.f;
        else if ((x < 0.f) != (y < 0.f)) r += y;
        return (double)r;
    }
    const double x = s, y = b;
    { const double ax = fabs(x); if (ax > 1e-290 && ax < 0.5 * y) return x < 0.0 ? x + y : x; }
    if (y == 0.0) return x;
    if (!(x == x) || !(y == y) || fabs(x) == (double)INFINITY) return (double)NAN;
    if (x == 0.0) return 0.0 / y;
    double r = fmod(x, y);
    bool req0 = (r == 0.0);
    if (!req0 && y > floor(y)) { const double q = fabs(x / y); req0 = !(fabs(q - floor(q + 0.5)) > 2.220446049250313e-16 * q); }
    if (req0) r = y * 0.0;
    else if ((x < 0.0) != (y < 0.0)) r += y;
    return r;
}

// fmod(s,b) + (s<0)*b   (PusherSliderModel.m:526; CasADi mod == C fmod)
// fmod is exact, and for |s| < 2b it equals s or s -+ b — a subtraction that is itself exact (Sterbenz), so the fast
// path returns bit-identical values while skipping the iterative library routine (every contact coordinate the
// solver meets lies in (-b, 2b)); NaN and larger arguments take the library path.
QS_HD double wrap_dyn(double s, double b) {
    const double a = fabs(s);
    double r;
    if (a < b) r = s;
    else if (a < 2.0 * b) r = (s > 0.0) ? s - b : s + b;
    else r = fmod(s, b);
    return r + ((s < 0.0) ? b : 0.0);
}

// gradient(atan2(C'_y, C'_x), s)  (bspline_shape.m:137-144) at sigma.
QS_HD double angle_rate(const double* __restrict__ M, double sg) {
    Curve c; curve_eval(M, sg, c);
    return (c.dx * c.hy - c.dy * c.hx) / (c.dx * c.dx + c.dy * c.dy);
}

// update_tangential_velocity_bounds (NMPC_controller.m:319-327)
QS_HD double v_bound_of(const double* __restrict__ M, double s, double v_alpha, double d_v_bound,
                        double t_angle0, double u_t_ub, bool single, double* t_angle_out) {
    double sm = matlab_mod(s, M[1], single);
    sm = matlab_mod(sm, M[1], single);                      // getAngleCurvatures wraps again (:147)
    const double ta = fabs(angle_rate(M, sm));
    if (t_angle_out) *t_angle_out = ta;
    return fmin(v_alpha / (fabs(ta - t_angle0) + 0.0001) + d_v_bound, u_t_ub);
}

// kappa(sigma) and d kappa / d sigma from the pp-form of FC_dot: (x, y) = FC_dot, kappa = (x y' - y x') / (x^2 + y^2).
QS_HD void angle_rate_d(const double* __restrict__ M, double sg, double& kap, double& dkap) {
    bool valid;
    const int q = span_of(M, sg, valid);
    if (!valid) { kap = 0.0 / 0.0; dkap = 0.0 / 0.0; return; }     // every basis function is 0: atan2(0, 0)' = 0/0
    const double tau = sg - M[HDR + q];
    const double* k = M + COEF_OFF + q * COEF_STRIDE;
    const double x = fma(fma(k[10], tau, k[9]), tau, k[8]), y = fma(fma(k[13], tau, k[12]), tau, k[11]);
    const double xp = fma(2.0 * k[10], tau, k[9]), yp = fma(2.0 * k[13], tau, k[12]);
    const double xpp = 2.0 * k[10], ypp = 2.0 * k[13];
    const double num = x * yp - y * xp, den = x * x + y * y;
    kap = num / den;
    dkap = ((x * ypp - y * xpp) * den - num * 2.0 * (x * xp + y * yp)) / (den * den);
}

// v_bound(s) of the parked constraint variant h = [u_n; u_t -+ v_bound(s)] (NMPC_controller.m:226-230, :238) and
// its derivative with CasADi's AD rules (comparisons constant, d fmod = 1, d|a| = sign a, d fmin(a, c) = [a < c]):
//   s_mod = (s<0)*b + mod(s,b);  t = |kappa(s_mod)|;  v = min(v_alpha / (|t - t0| + 1e-4) + d_v_bound, u_t_ub)
QS_HD double v_bound_sym(const double* __restrict__ M, double s, const double vbp[4], double* dv) {
    const double sw = wrap_dyn(s, M[1]);
    double kap, dkap;
    angle_rate_d(M, sw, kap, dkap);
    const double t = fabs(kap), e = fabs(t - vbp[2]) + 0.0001;
    const double a = vbp[0] / e + vbp[1];
    const double sg1 = (double)((kap > 0.0) - (kap < 0.0)), sg2 = (double)((t - vbp[2] > 0.0) - (t - vbp[2] < 0.0));
    if (dv) *dv = (a < vbp[3]) ? -vbp[0] / (e * e) * sg2 * sg1 * dkap : 0.0;
    return fmin(a, vbp[3]);
}

// ---- dynamics ----------------------------------------------------------------------------------
struct Dyn {
    double f[4];     // xdot
    double fs[4];    // d f / d s
    double fun[4];   // d f / d u_n
    double fut[4];   // d f / d u_t
    // d f / d theta = (-f[1], f[0], 0, 0);  d f / d x = d f / d y = 0
};

// f(x,u) and, if JAC, its partial derivatives.  The three contact modes are combined through
// 0/1 indicator products exactly like PusherSliderModel.m:587-589, so IEEE corner cases behave
// like the reference: u_n = u_t = 0 gives r = NaN, all indicators 0, f = 0 and a zero Jacobian;
// u_n = 0, u_t != 0 gives r = +-Inf, sliding with v = 0 and s_dot = u_t; sigma == b (or NaN)
// makes every basis function 0 and the tangent 0/0 = NaN.
// The derivative rules are those of CasADi's AD: comparisons are constants, d fmod(s,b)/ds = 1.
template <bool JAC>
QS_HD void dyn_eval(const double* __restrict__ M, double th, double s, double un, double ut, Dyn& o) {
    const double b = M[1], mu = M[3], c2 = M[4];
    const double sg = wrap_dyn(s, b);
    Curve c; curve_eval(M, sg, c);
    const double nrm = sqrt(c.dx * c.dx + c.dy * c.dy);
    const double rn = 1.0 / nrm;
    const double tx = c.dx * rn, ty = c.dy * rn;            // t = C'/|C'| ; n = (t_y, -t_x)  (:109-110)
    const double px = ty * c.cx - tx * c.cy;                // NT_p = R_NT' * S_p'            (:532-534)
    const double py = tx * c.cx + ty * c.cy;
    double sth, cth;
#if defined(__CUDA_ARCH__)
    sincos(th, &sth, &cth);
#else
    sth = sin(th); cth = cos(th);
#endif
    const double px2 = px * px, py2 = py * py, pxy = px * py;
    const double k = 1.0 / (c2 + px2 + py2);                // factor_matrix                  (:544)
    const double Dl = c2 + py2 - mu * pxy, Dr = c2 + py2 + mu * pxy;
    const double gl = (mu * c2 - pxy + mu * px2) / Dl;      // gamma_l                        (:547)
    const double gr = (-mu * c2 - pxy - mu * px2) / Dr;     // gamma_r                        (:548)
    const double r = ut / un;                               // u_fract                        (:551)
    const double ist = ((r >= gr) && (r <= gl)) ? 1.0 : 0.0;
    const double isl = (r > gl) ? 1.0 : 0.0;
    const double isr = (r < gr) ? 1.0 : 0.0;
    const double nind = ist + isl + isr;
    // effective tangential velocity of the contact point: u_t (stick), gamma*u_n (slide)
    const double weff = ist * ut + isl * (gl * un) + isr * (gr * un);
    const double unn = nind * un;
    const double qxx = c2 + px2, qyy = c2 + py2;
    const double an = qxx * unn + pxy * weff;               // Q*[u_n; w]                     (:555)
    const double at = pxy * unn + qyy * weff;
    const double ath = -py * unn + px * weff;               // b' * [u_n; w]
    const double vn = k * an, vt = k * at;
    const double Vx = ty * vn + tx * vt;                    // S_R_NT * v, R_NT = [n t]       (:531)
    const double Vy = -tx * vn + ty * vt;
    o.f[0] = cth * Vx - sth * Vy;                           // W_R_S                          (:554)
    o.f[1] = sth * Vx + cth * Vy;
    o.f[2] = k * ath;
    o.f[3] = isl * (ut - un * gl) + isr * (ut - un * gr);   // s_dot                          (:570,:582)
    if (JAC) {
        // frame rates: t' = w n, n' = -w t, with w = (n . d(C')/ds) / |C'|
        const double w = (ty * c.hx - tx * c.hy) * rn;
        const double ndc = ty * c.gx - tx * c.gy;           // n . dC/ds
        const double tdc = tx * c.gx + ty * c.gy;           // t . dC/ds
        const double pxs = -w * py + ndc;
        const double pys = w * px + tdc;
        const double ks = -k * k * 2.0 * (px * pxs + py * pys);
        const double pxys = pxs * py + px * pys;
        const double gls = ((-pxys + 2.0 * mu * px * pxs) - gl * (2.0 * py * pys - mu * pxys)) / Dl;
        const double grs = ((-pxys - 2.0 * mu * px * pxs) - gr * (2.0 * py * pys + mu * pxys)) / Dr;
        const double ws = un * (isl * gls + isr * grs);
        const double ans = 2.0 * px * pxs * unn + pxys * weff + pxy * ws;
        const double ats = pxys * unn + 2.0 * py * pys * weff + qyy * ws;
        const double aths = -pys * unn + pxs * weff + px * ws;
        const double qn = (ks * an + k * ans) + w * vt;
        const double qt = (ks * at + k * ats) - w * vn;
        const double Vxs = ty * qn + tx * qt, Vys = -tx * qn + ty * qt;
        o.fs[0] = cth * Vxs - sth * Vys;
        o.fs[1] = sth * Vxs + cth * Vys;
        o.fs[2] = ks * ath + k * aths;
        o.fs[3] = -ws;
        const double wun = isl * gl + isr * gr;
        const double vnu = k * (qxx * nind + pxy * wun), vtu = k * (pxy * nind + qyy * wun);
        const double Vxu = ty * vnu + tx * vtu, Vyu = -tx * vnu + ty * vtu;
        o.fun[0] = cth * Vxu - sth * Vyu;
        o.fun[1] = sth * Vxu + cth * Vyu;
        o.fun[2] = k * (-py * nind + px * wun);
        o.fun[3] = -wun;
        const double vnt = k * (pxy * ist), vtt = k * (qyy * ist);
        const double Vxt = ty * vnt + tx * vtt, Vyt = -tx * vnt + ty * vtt;
        o.fut[0] = cth * Vxt - sth * Vyt;
        o.fut[1] = sth * Vxt + cth * Vyt;
        o.fut[2] = k * (px * ist);
        o.fut[3] = isl + isr;
    }
}

// ---- ERK4 + forward sensitivities ----------------------------------------------------------------
// One classical RK4 step of dt on the augmented state [x, S] (acados sim_erk).
// Because df/dx = df/dy = 0, the sensitivity columns with respect to x0 and y0 stay e1, e2
// exactly; only the columns for (theta0, s0, u_n, u_t) are integrated:  S is 4x4 row-major with
// those four columns, so  A = [e1 e2 S(:,0) S(:,1)],  B = [S(:,2) S(:,3)].
QS_HD void erk4_sens(const double* __restrict__ M, const double x[4], double un, double ut, double dt,
                     double Phi[4], double S[16]) {
    double kx[4] = {0, 0, 0, 0}, ax[4] = {0, 0, 0, 0};
    double kS[16], aS[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { kS[i] = 0.0; aS[i] = 0.0; }
#pragma unroll
    for (int st = 0; st < 4; ++st) {
        const double a = (st == 0 ? 0.0 : (st == 3 ? 1.0 : 0.5)) * dt;
        const double bw = (st == 0 || st == 3) ? (1.0 / 6.0) : (1.0 / 3.0);
        Dyn d;
        dyn_eval<true>(M, x[2] + a * kx[2], x[3] + a * kx[3], un, ut, d);
        double nS[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const double s2 = (c == 0 ? 1.0 : 0.0) + a * kS[8 + c];     // theta row of the stage sensitivity
            const double s3 = (c == 1 ? 1.0 : 0.0) + a * kS[12 + c];    // s row
            const double ju0 = (c == 2) ? d.fun[0] : (c == 3 ? d.fut[0] : 0.0);
            const double ju1 = (c == 2) ? d.fun[1] : (c == 3 ? d.fut[1] : 0.0);
            const double ju2 = (c == 2) ? d.fun[2] : (c == 3 ? d.fut[2] : 0.0);
            const double ju3 = (c == 2) ? d.fun[3] : (c == 3 ? d.fut[3] : 0.0);
            nS[0 + c] = fma(-d.f[1], s2, fma(d.fs[0], s3, ju0));
            nS[4 + c] = fma(d.f[0], s2, fma(d.fs[1], s3, ju1));
            nS[8 + c] = fma(d.fs[2], s3, ju2);
            nS[12 + c] = fma(d.fs[3], s3, ju3);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { kx[i] = d.f[i]; ax[i] = fma(bw, d.f[i], ax[i]); }
#pragma unroll
        for (int i = 0; i < 16; ++i) { kS[i] = nS[i]; aS[i] = fma(bw, nS[i], aS[i]); }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) Phi[i] = fma(dt, ax[i], x[i]);
#pragma unroll
    for (int i = 0; i < 16; ++i) S[i] = dt * aS[i];
    S[8 + 0] += 1.0;     // d theta / d theta0
    S[12 + 1] += 1.0;    // d s / d s0
}

// ERK4 without sensitivities (merit function of the SQP line search).
QS_HD void erk4_plain(const double* __restrict__ M, const double x[4], double un, double ut, double dt, double Phi[4]) {
    double kx[4] = {0, 0, 0, 0}, ax[4] = {0, 0, 0, 0};
#pragma unroll
    for (int st = 0; st < 4; ++st) {
        const double a = (st == 0 ? 0.0 : (st == 3 ? 1.0 : 0.5)) * dt;
        const double bw = (st == 0 || st == 3) ? (1.0 / 6.0) : (1.0 / 3.0);
        Dyn d;
        dyn_eval<false>(M, x[2] + a * kx[2], x[3] + a * kx[3], un, ut, d);
#pragma unroll
        for (int i = 0; i < 4; ++i) { kx[i] = d.f[i]; ax[i] = fma(bw, d.f[i], ax[i]); }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) Phi[i] = fma(dt, ax[i], x[i]);
}

}  // namespace qs
