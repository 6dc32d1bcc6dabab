// qs_model.cpp — see qs_model.hpp.  Host only; runs once per object.
#include "qs_model.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace qs {
namespace {

// ---- cubic polynomial in tau with long double coefficients -------------------------------------
struct Poly {
    long double c[4] = {0, 0, 0, 0};
};
Poly operator+(const Poly& a, const Poly& b) { Poly r; for (int i = 0; i < 4; ++i) r.c[i] = a.c[i] + b.c[i]; return r; }
Poly scale(const Poly& a, long double s) { Poly r; for (int i = 0; i < 4; ++i) r.c[i] = a.c[i] * s; return r; }
// (l0 + l1*tau) * a, truncated at degree 3 (never exceeded: basis degree <= 3)
Poly mul_lin(const Poly& a, long double l0, long double l1) {
    Poly r;
    for (int i = 0; i < 4; ++i) {
        r.c[i] += l0 * a.c[i];
        if (i + 1 < 4) r.c[i + 1] += l1 * a.c[i];
    }
    return r;
}

struct Knots {
    const std::vector<double>& S;
    bool single;
    double at(int i) const { return S[i - 1]; }                                  // 1-based like the reference
    long double diff(int hi, int lo) const {                                     // as MATLAB computes S(hi)-S(lo)
        if (single) return (long double)((float)S[hi - 1] - (float)S[lo - 1]);
        return (long double)S[hi - 1] - (long double)S[lo - 1];
    }
};

// eval_bspline.m:1-33 with s = S(j) + tau restricted to the span j (S(j) <= s < S(j+1))
Poly basis_poly(const Knots& K, int j, int i, int ord) {
    Poly zero;
    if (K.at(i + ord + 1) == K.at(i)) return zero;                               // :6-9
    if (ord == 0) { Poly r; r.c[0] = (i == j) ? 1.0L : 0.0L; return r; }         // :11-14
    Poly Na = basis_poly(K, j, i, ord - 1);                                      // :16
    Poly Nb = basis_poly(K, j, i + 1, ord - 1);                                  // :17
    Poly r;
    if (!(K.at(i + ord) == K.at(i))) {                                           // :19-23  (s - S(i)) / (S(i+p) - S(i))
        const long double den = K.diff(i + ord, i);
        r = r + mul_lin(Na, ((long double)K.at(j) - (long double)K.at(i)) / den, 1.0L / den);
    }
    if (!(K.at(i + ord + 1) == K.at(i + 1))) {                                   // :24-28  (S(i+p+1) - s) / (...)
        const long double den = K.diff(i + ord + 1, i + 1);
        r = r + mul_lin(Nb, ((long double)K.at(i + ord + 1) - (long double)K.at(j)) / den, -1.0L / den);
    }
    return r;                                                                    // :30
}

// ---- outline ingest (PusherSliderModel.m:84-111), float32 like pcread -----------------------------
bool read_ply_xy(const char* path, std::vector<float>& xy, std::string& err) {
    FILE* f = std::fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    char line[1024];
    long nv = -1;
    int nfloat = 0;
    bool vertex_block = false, le = false, done = false, only_float = true;
    while (std::fgets(line, sizeof line, f)) {
        if (!std::strncmp(line, "format", 6)) le = std::strstr(line, "binary_little_endian") != nullptr;
        else if (!std::strncmp(line, "element", 7)) {
            vertex_block = std::strstr(line, "vertex") != nullptr;
            if (vertex_block) nv = std::strtol(line + 15, nullptr, 10);
        } else if (!std::strncmp(line, "property", 8) && vertex_block) {
            if (std::strstr(line, "float") && !std::strstr(line, "list")) ++nfloat; else only_float = false;
        } else if (!std::strncmp(line, "end_header", 10)) { done = true; break; }
    }
    if (!done || !le || nv <= 0 || nfloat < 2 || !only_float) {
        std::fclose(f);
        err = "unsupported PLY (need binary_little_endian, float32 vertex properties x,y,...)";
        return false;
    }
    std::vector<float> rec(nfloat);
    xy.assign((size_t)nv * 2, 0.f);
    for (long v = 0; v < nv; ++v) {
        if (std::fread(rec.data(), sizeof(float), (size_t)nfloat, f) != (size_t)nfloat) {
            std::fclose(f); err = "truncated PLY vertex data"; return false;
        }
        xy[2 * v] = rec[0]; xy[2 * v + 1] = rec[1];
    }
    std::fclose(f);
    return true;
}

// greedy nearest-neighbour chain from the min-x vertex, mm -> m, closed, optionally reversed
std::vector<float> order_outline(const std::vector<float>& in, bool flip) {
    const int nv = (int)in.size() / 2;
    std::vector<float> px(nv), py(nv);
    for (int i = 0; i < nv; ++i) { px[i] = in[2 * i]; py[i] = in[2 * i + 1]; }
    const float inf = std::numeric_limits<float>::infinity();
    int cur = 0;
    for (int i = 1; i < nv; ++i) if (px[i] < px[cur]) cur = i;                  // :91
    std::vector<float> ox, oy;
    float qx = px[cur], qy = py[cur];
    ox.push_back(qx); oy.push_back(qy);
    px[cur] = inf; py[cur] = inf;                                               // :93
    for (int step = 1; step < nv; ++step) {                                     // :98-103
        int arg = 0; float best = inf; bool first = true;
        for (int i = 0; i < nv; ++i) {
            const float ex = px[i] - qx, ey = py[i] - qy;
            const float dist = std::sqrt(ex * ex + ey * ey);
            if (first || dist < best) { best = dist; arg = i; first = false; }
        }
        qx = px[arg]; qy = py[arg];
        ox.push_back(qx); oy.push_back(qy);
        px[arg] = inf; py[arg] = inf;
    }
    const float mm = (float)(1.0 / 1000.0);                                     // :105 (scale_factor = 1000, :72)
    for (auto& v : ox) v *= mm;
    for (auto& v : oy) v *= mm;
    ox.push_back(ox[0]); oy.push_back(oy[0]);                                   // :106
    const int n = (int)ox.size();
    std::vector<float> out((size_t)n * 2);
    for (int i = 0; i < n; ++i) {
        const int src = flip ? (n - 1 - i) : i;                                 // :107-109
        out[2 * i] = ox[src]; out[2 * i + 1] = oy[src];
    }
    return out;
}

}  // namespace

std::string model_from_tables(const double* S, int nknots, const double* P, int n, int p, double mu_sp,
                              double c_ellipse, bool single_coeffs, HostModel& M) {
    if (!S || !P) return "null table";
    if (p != 3) return "only cubic outlines (order_spline = 3, main.m:33) are supported";
    if (n < p + 1 || nknots != n + p + 1) return "need nknots == n + p + 1 and n >= p + 1";
    for (int i = 1; i < nknots; ++i) if (S[i] < S[i - 1]) return "knot vector must be non-decreasing";
    M.p = p; M.n = n; M.S.assign(S, S + nknots); M.P.assign(P, P + 2 * (size_t)n);
    M.mu_sp = mu_sp; M.c_ellipse = c_ellipse; M.single_coeffs = single_coeffs;
    M.b = S[nknots - 1];                                                        // bspline_shape.m:37 (== last knot)
    Knots K{M.S, single_coeffs};
    // cj_1 (bspline_shape.m:90-98) and cj_2 (:124-131)
    M.c1.assign((size_t)n * 2, 0.0); M.c2.assign((size_t)n * 2, 0.0);
    for (int ii = 2; ii <= n; ++ii) {
        if (K.at(ii + p) == K.at(ii)) continue;
        for (int c = 0; c < 2; ++c) {
            const double hi = M.P[2 * (ii - 1) + c], lo = M.P[2 * (ii - 2) + c];
            if (single_coeffs) M.c1[2 * (ii - 1) + c] = (double)((float)p * (((float)hi - (float)lo) / (float)K.diff(ii + p, ii)));
            else M.c1[2 * (ii - 1) + c] = p * ((hi - lo) / (double)K.diff(ii + p, ii));
        }
    }
    for (int ii = 3; ii <= n; ++ii) {
        const double den = (double)K.diff(ii + p - 1, ii);
        if (std::fabs(den) < 1e-5) continue;
        for (int c = 0; c < 2; ++c) {
            const double num = M.c1[2 * (ii - 1) + c] - M.c1[2 * (ii - 2) + c];
            if (single_coeffs) M.c2[2 * (ii - 1) + c] = (double)((float)(p - 1) * ((float)num / (float)den));
            else M.c2[2 * (ii - 1) + c] = (p - 1) * (num / den);
        }
    }
    // pp-form tables
    M.blob.assign(MODEL_DOUBLES, 0.0);
    int nspan = 0;
    for (int j = 1; j < nknots; ++j) {
        if (!(K.at(j) < K.at(j + 1))) continue;
        if (nspan >= MAXSPAN) return "too many knot spans (max 64)";
        M.blob[HDR + nspan] = K.at(j);
        double* co = &M.blob[COEF_OFF + (size_t)nspan * COEF_STRIDE];
        Poly cx, cy, dx, dy, ex, ey;
        for (int i = std::max(1, j - p); i <= std::min(n, j); ++i) {
            Poly N = basis_poly(K, j, i, p);
            cx = cx + scale(N, M.P[2 * (i - 1)]); cy = cy + scale(N, M.P[2 * (i - 1) + 1]);
        }
        for (int i = std::max(2, j - p + 1); i <= std::min(n, j); ++i) {
            Poly N = basis_poly(K, j, i, p - 1);
            dx = dx + scale(N, M.c1[2 * (i - 1)]); dy = dy + scale(N, M.c1[2 * (i - 1) + 1]);
        }
        for (int i = std::max(3, j - p + 2); i <= std::min(n, j); ++i) {
            Poly N = basis_poly(K, j, i, p - 2);
            ex = ex + scale(N, M.c2[2 * (i - 1)]); ey = ey + scale(N, M.c2[2 * (i - 1) + 1]);
        }
        for (int d = 0; d < 4; ++d) { co[d] = (double)cx.c[d]; co[4 + d] = (double)cy.c[d]; }
        for (int d = 0; d < 3; ++d) { co[8 + d] = (double)dx.c[d]; co[11 + d] = (double)dy.c[d]; }
        for (int d = 0; d < 2; ++d) { co[14 + d] = (double)ex.c[d]; co[16 + d] = (double)ey.c[d]; }
        ++nspan;
    }
    if (nspan == 0) return "degenerate knot vector";
    M.blob[HDR + nspan] = K.at(nknots);
    M.blob[0] = (double)nspan;
    M.blob[1] = M.b;
    M.blob[2] = (double)nspan / (M.blob[HDR + nspan] - M.blob[HDR]);
    M.blob[3] = mu_sp;
    M.blob[4] = c_ellipse * c_ellipse;
    M.blob[5] = c_ellipse;
    return "";
}

std::string model_from_ply(const char* path, bool flip, int p, double mu_sg, double mu_sp, double mass,
                           double tau_max, HostModel& M) {
    std::vector<float> raw;
    std::string err;
    if (!read_ply_xy(path, raw, err)) return err;
    std::vector<float> P = order_outline(raw, flip);
    const int n = (int)P.size() / 2;
    // getSpline, PusherSliderModel.m:117-123, single arithmetic
    float b = 0.f;
    for (int i = 0; i + 1 < n; ++i) {
        const float ex = P[2 * (i + 1)] - P[2 * i], ey = P[2 * (i + 1) + 1] - P[2 * i + 1];
        b += std::sqrt(ex * ex + ey * ey);
    }
    const int m = n - p + 1;
    std::vector<double> S;
    for (int i = 0; i < p; ++i) S.push_back(0.0);
    for (int k = 0; k < m; ++k) {
        float v = ((float)k * b) / (float)(m - 1);          // linspace: d1 + (0:n1).*(d2-d1)./n1
        if (k == 0) v = 0.f;
        if (k == m - 1) v = b;
        S.push_back((double)v);
    }
    for (int i = 0; i < p; ++i) S.push_back((double)b);
    std::vector<double> Pd(P.begin(), P.end());
    const double f_max = mu_sg * mass * 9.81;               // PusherSliderModel.m:53, helper.m:3
    return model_from_tables(S.data(), (int)S.size(), Pd.data(), n, p, mu_sp, tau_max / f_max, true, M);
}

}  // namespace qs
