// oracle/qs_arbiter.cpp
//
// TEST INFRASTRUCTURE ONLY (same rule as qs_oracle.hpp: only tests/, smoke() and bench.py's CPU legs may load it).
//
// Extended-precision ARBITER for the stage-structured QP of one SQP iteration (the QP that
// acados hands to HPIPM at /root/reference/acados_nmpc/NMPC_controller.m:272-276).  It answers the question
// "what IS the solution of this QP" independently of every interior-point path, so that the oracle's IPM and the
// CUDA kernels can each be compared with the exact answer instead of only with each other.
//
// Method (deliberately unrelated to the Riccati recursion of the oracle / the kernels), all in __float128
// (113-bit significand):
//   1. condensing: x_k = c_k + sum_j G_kj u_j by forward substitution of the linearised dynamics;
//      reduced Hessian Hr = sum_k Z_k' H_k Z_k + G_N' Q_N G_N (dense, 2N x 2N), Cholesky Hr = L L';
//   2. primal-dual active-set iteration on the inequality rows: with the working set pinned as equalities the KKT
//      system is solved through the Schur complement S = C Hr^-1 C' (Cholesky); rows whose multiplier has the wrong
//      sign are released, free rows that violate a bound are pinned; repeat until nothing changes.  That iteration is
//      fast from a good guess but can cycle; after 30 sweeps the Goldfarb-Idnani dual active-set method (finite for a
//      strictly convex QP) takes over from the unconstrained minimiser;
//   3. certificate: costates by the adjoint recursion, then the KKT residuals of the ORIGINAL stage-wise QP
//      (input stationarity, dynamics, primal / dual feasibility, complementarity) in __float128.  The QP is convex,
//      so KKT residuals of 1e-25 with cond ~ 1e7 pin the solution to ~1e-18: exact for every FP64 comparison.
//
// Build: make -C oracle  ->  oracle/libqs_arbiter.so   (g++; __float128 arithmetic comes from libgcc, the square root is a
// Newton iteration below, so libquadmath is not needed)
#include <cstring>
#include <vector>

namespace {

typedef __float128 Q;

inline Q qabs(Q a) { return a < 0 ? -a : a; }
inline Q qsqrt(Q a) {                                      // FP64 seed + three Newton steps (53 -> 106 -> 113 bits)
    Q x = (Q)__builtin_sqrt((double)a);
    for (int i = 0; i < 3; ++i) x = (x + a / x) / 2;
    return x;
}
inline Q qmax(Q a, Q b) { return a > b ? a : b; }

// dense Cholesky (lower, in place, row-major n x n); returns false on a non-positive pivot
bool chol(std::vector<Q>& A, int n) {
    for (int j = 0; j < n; ++j) {
        Q d = A[(size_t)j * n + j];
        for (int p = 0; p < j; ++p) d -= A[(size_t)j * n + p] * A[(size_t)j * n + p];
        if (!(d > 0)) return false;
        const Q l = qsqrt(d);
        A[(size_t)j * n + j] = l;
        for (int i = j + 1; i < n; ++i) {
            Q a = A[(size_t)i * n + j];
            for (int p = 0; p < j; ++p) a -= A[(size_t)i * n + p] * A[(size_t)j * n + p];
            A[(size_t)i * n + j] = a / l;
        }
    }
    return true;
}
void fwd(const std::vector<Q>& L, int n, Q* v) {          // L y = v
    for (int i = 0; i < n; ++i) { Q a = v[i]; for (int p = 0; p < i; ++p) a -= L[(size_t)i * n + p] * v[p]; v[i] = a / L[(size_t)i * n + i]; }
}
void bwd(const std::vector<Q>& L, int n, Q* v) {          // L' y = v
    for (int i = n - 1; i >= 0; --i) { Q a = v[i]; for (int p = i + 1; p < n; ++p) a -= L[(size_t)p * n + i] * v[p]; v[i] = a / L[(size_t)i * n + i]; }
}

struct Row { int k, c; std::vector<Q> r; Q r0; Q lo, hi; };   // value = r.u + r0, lo <= value <= hi

}  // namespace

extern "C" {

// Stage data (all row-major, doubles): H [N][36] in z = [u_n,u_t,x,y,theta,s] order, g [N][6], A [N][16], B [N][8] (4x2),
// b [N][4], QN [16], qN [4], dx0 [4]; inequality rows (k, c), c = 0..2: on [N][3], value = z[ci] + beta*z[5],
// dl <= value <= du.  act [N][3]: in = initial working set (-1 lower, 0 free, +1 upper), out = final working set.
// Outputs: du_o [N][2], dx_o [N+1][4], pi_o [N][4] (pi_o[k] = pi_{k+1}), lam_o [N][6] ([lower(3); upper(3)] per stage),
// kkt [5] = {input stationarity, dynamics, primal infeasibility, dual infeasibility, complementarity} (inf-norms,
// evaluated in __float128 on the __float128 solution), iters = active-set iterations.
// Returns 0 ok, 1 active-set iteration limit, 2 reduced Hessian not positive definite.
int arb_qp_solve(int N, const double* H, const double* g, const double* A, const double* B, const double* b,
                 const double* QN, const double* qN, const double* dx0,
                 const double* dl, const double* du, const int* on, const int* ci, const double* beta,
                 int* act, double* du_o, double* dx_o, double* pi_o, double* lam_o, double* kkt, int* iters) {
    const int nu = 2 * N, nc1 = nu + 1;
    // ---- 1. condensing: X[k] is 4 x (nu + 1), last column = constant part
    std::vector<std::vector<Q>> X(N + 1, std::vector<Q>((size_t)4 * nc1, 0));
    for (int i = 0; i < 4; ++i) X[0][(size_t)i * nc1 + nu] = dx0[i];
    for (int k = 0; k < N; ++k) {
        for (int i = 0; i < 4; ++i) {
            for (int col = 0; col < nc1; ++col) {
                Q a = 0;
                for (int l = 0; l < 4; ++l) a += (Q)A[(size_t)k * 16 + 4 * i + l] * X[k][(size_t)l * nc1 + col];
                X[k + 1][(size_t)i * nc1 + col] = a;
            }
            X[k + 1][(size_t)i * nc1 + 2 * k] += B[(size_t)k * 8 + 2 * i];
            X[k + 1][(size_t)i * nc1 + 2 * k + 1] += B[(size_t)k * 8 + 2 * i + 1];
            X[k + 1][(size_t)i * nc1 + nu] += b[(size_t)k * 4 + i];
        }
    }
    auto zrow = [&](int k, int i, int col) -> Q {                 // row i of Z_k (z_k = Z_k [u; 1])
        if (i < 2) return (col == 2 * k + i) ? (Q)1 : (Q)0;
        return X[k][(size_t)(i - 2) * nc1 + col];
    };
    std::vector<Q> Hr((size_t)nu * nu, 0), hr(nu, 0);
    {
        std::vector<Q> HZ((size_t)6 * nc1);
        for (int k = 0; k < N; ++k) {
            for (int i = 0; i < 6; ++i) for (int col = 0; col < nc1; ++col) {
                Q a = 0;
                for (int l = 0; l < 6; ++l) a += (Q)H[(size_t)k * 36 + 6 * i + l] * zrow(k, l, col);
                HZ[(size_t)i * nc1 + col] = a;
            }
            const int lim = 2 * k + 2;                               // Z_k depends on u_0 .. u_k only
            for (int r = 0; r < lim; ++r) {
                for (int col = 0; col <= r; ++col) {
                    Q a = 0;
                    for (int i = 0; i < 6; ++i) a += zrow(k, i, r) * HZ[(size_t)i * nc1 + col];
                    Hr[(size_t)r * nu + col] += a;
                }
                Q a = 0;
                for (int i = 0; i < 6; ++i) a += zrow(k, i, r) * (HZ[(size_t)i * nc1 + nu] + (Q)g[(size_t)k * 6 + i]);
                hr[r] += a;
            }
        }
        std::vector<Q> QX((size_t)4 * nc1);
        for (int i = 0; i < 4; ++i) for (int col = 0; col < nc1; ++col) {
            Q a = 0;
            for (int l = 0; l < 4; ++l) a += (Q)QN[4 * i + l] * X[N][(size_t)l * nc1 + col];
            QX[(size_t)i * nc1 + col] = a;
        }
        for (int r = 0; r < nu; ++r) {
            for (int col = 0; col <= r; ++col) {
                Q a = 0;
                for (int i = 0; i < 4; ++i) a += X[N][(size_t)i * nc1 + r] * QX[(size_t)i * nc1 + col];
                Hr[(size_t)r * nu + col] += a;
            }
            Q a = 0;
            for (int i = 0; i < 4; ++i) a += X[N][(size_t)i * nc1 + r] * (QX[(size_t)i * nc1 + nu] + (Q)qN[i]);
            hr[r] += a;
        }
        for (int r = 0; r < nu; ++r) for (int col = r + 1; col < nu; ++col) Hr[(size_t)r * nu + col] = Hr[(size_t)col * nu + r];
    }
    std::vector<Q> L = Hr;
    if (!chol(L, nu)) return 2;
    // inequality rows in terms of u
    std::vector<Row> rows;
    for (int k = 0; k < N; ++k) for (int c = 0; c < 3; ++c) {
        if (!on[k * 3 + c]) { act[k * 3 + c] = 0; continue; }
        Row r; r.k = k; r.c = c; r.r.assign(nu, 0);
        const int i0 = ci[k * 3 + c]; const Q be = beta[k * 3 + c];
        for (int col = 0; col < nu; ++col) r.r[col] = zrow(k, i0, col) + be * zrow(k, 5, col);
        r.r0 = zrow(k, i0, nu) + be * zrow(k, 5, nu);
        r.lo = dl[k * 3 + c]; r.hi = du[k * 3 + c];
        rows.push_back(r);
    }
    const int m = (int)rows.size();
    std::vector<int> ws(m);
    for (int i = 0; i < m; ++i) ws[i] = act[rows[i].k * 3 + rows[i].c];
    // unconstrained minimiser pieces: u = -Hr^-1 (hr + C' nu)
    std::vector<Q> u0(nu);
    for (int i = 0; i < nu; ++i) u0[i] = -hr[i];
    fwd(L, nu, u0.data()); bwd(L, nu, u0.data());
    std::vector<std::vector<Q>> Y(m);                                 // Y_i = L^-1 r_i
    for (int i = 0; i < m; ++i) { Y[i] = rows[i].r; fwd(L, nu, Y[i].data()); }
    std::vector<Q> u(nu), nuv(m, 0), val(m);
    const Q tol = 1e-24Q;
    // equality-constrained QP on the working set ws: fills u, nuv (signed multiplier lam_upper - lam_lower), val
    auto eqp = [&](const std::vector<int>& wset) -> bool {
        std::vector<int> idx;
        for (int i = 0; i < m; ++i) if (wset[i]) idx.push_back(i);
        const int ma = (int)idx.size();
        // Schur complement S = C Hr^-1 C' (+ tiny regularisation: dependent rows keep it factorisable), rhs = C u0 + r0 - d
        std::vector<Q> S((size_t)ma * ma), rhs(ma);
        for (int a = 0; a < ma; ++a) {
            for (int c2 = 0; c2 <= a; ++c2) {
                Q s = 0; const std::vector<Q>& ya = Y[idx[a]]; const std::vector<Q>& yb = Y[idx[c2]];
                for (int p = 0; p < nu; ++p) s += ya[p] * yb[p];
                S[(size_t)a * ma + c2] = s;
            }
            S[(size_t)a * ma + a] += 1e-28Q * (S[(size_t)a * ma + a] + 1);
            Q v = rows[idx[a]].r0;
            for (int p = 0; p < nu; ++p) v += rows[idx[a]].r[p] * u0[p];
            rhs[a] = v - (wset[idx[a]] < 0 ? rows[idx[a]].lo : rows[idx[a]].hi);
        }
        if (ma) { if (!chol(S, ma)) return false; fwd(S, ma, rhs.data()); bwd(S, ma, rhs.data()); }
        std::vector<Q> w(nu, 0);                                        // u = u0 - Hr^-1 C' nu
        for (int a = 0; a < ma; ++a) for (int p = 0; p < nu; ++p) w[p] += rows[idx[a]].r[p] * rhs[a];
        fwd(L, nu, w.data()); bwd(L, nu, w.data());
        for (int p = 0; p < nu; ++p) u[p] = u0[p] - w[p];
        std::fill(nuv.begin(), nuv.end(), (Q)0);
        for (int a = 0; a < ma; ++a) nuv[idx[a]] = rhs[a];
        for (int i = 0; i < m; ++i) { Q v = rows[i].r0; for (int p = 0; p < nu; ++p) v += rows[i].r[p] * u[p]; val[i] = v; }
        return true;
    };
    // (a) primal-dual active-set sweeps from the caller's guess: a handful of iterations when the guess is good
    int it = 0, status = 1;
    for (; it < 30; ++it) {
        if (!eqp(ws)) return 2;
        int changes = 0;
        for (int i = 0; i < m; ++i) {
            int nw = ws[i];
            if (ws[i] < 0 && nuv[i] > tol) nw = 0;                      // lower pinned needs lam_l = -nu >= 0
            else if (ws[i] > 0 && nuv[i] < -tol) nw = 0;
            else if (ws[i] == 0 && val[i] < rows[i].lo - tol) nw = -1;
            else if (ws[i] == 0 && val[i] > rows[i].hi + tol) nw = 1;
            if (nw != ws[i]) { ++changes; ws[i] = nw; }
        }
        if (!changes) { status = 0; break; }
    }
    // (b) fallback with guaranteed finite termination: Goldfarb-Idnani dual active-set method from the unconstrained
    //     minimiser, in the transformed variable w = L' u (Hessian = identity).  Constraint j < m: lower side of row j,
    //     n = +Y_j, b = lo - r0;  j >= m: upper side of row j - m, n = -Y, b = r0 - hi;  feasible when n'w - b >= 0.
    if (status != 0) {
        std::vector<Q> w(nu);
        for (int i = 0; i < nu; ++i) w[i] = -hr[i];
        fwd(L, nu, w.data());                                             // w0 = L' u0 = -L^-1 hr
        auto nvec = [&](int j, int p) -> Q { return j < m ? Y[j][p] : -Y[j - m][p]; };
        auto bval = [&](int j) -> Q { return j < m ? rows[j].lo - rows[j].r0 : rows[j - m].r0 - rows[j - m].hi; };
        std::vector<int> W; std::vector<Q> mult;                          // working set and its multipliers (>= 0)
        const int gi_max = 20 * m + 100;
        int gi = 0; bool done = false;
        while (!done && gi < gi_max) {
            // most violated constraint outside the working set
            int p = -1; Q sp = -tol;
            for (int j = 0; j < 2 * m; ++j) {
                bool inW = false; for (int q : W) if (q == j) { inW = true; break; }
                if (inW) continue;
                Q sj = -bval(j); for (int q = 0; q < nu; ++q) sj += nvec(j, q) * w[q];
                if (sj < sp) { sp = sj; p = j; }
            }
            if (p < 0) { done = true; break; }
            Q up = 0;                                                      // multiplier of the constraint being added
            for (;;) {
                ++gi; ++it;
                if (gi >= gi_max) break;
                const int ma = (int)W.size();
                std::vector<Q> r(ma, 0), z(nu);
                if (ma) {
                    std::vector<Q> S((size_t)ma * ma);
                    for (int a = 0; a < ma; ++a) {
                        for (int c2 = 0; c2 <= a; ++c2) { Q s2 = 0; for (int q = 0; q < nu; ++q) s2 += nvec(W[a], q) * nvec(W[c2], q); S[(size_t)a * ma + c2] = s2; }
                        Q s2 = 0; for (int q = 0; q < nu; ++q) s2 += nvec(W[a], q) * nvec(p, q);
                        r[a] = s2;
                    }
                    if (!chol(S, ma)) return 2;
                    fwd(S, ma, r.data()); bwd(S, ma, r.data());
                }
                Q zn2 = 0, np2 = 0, znp = 0;
                for (int q = 0; q < nu; ++q) {
                    Q v = nvec(p, q); for (int a = 0; a < ma; ++a) v -= nvec(W[a], q) * r[a];
                    z[q] = v; zn2 += v * v; np2 += nvec(p, q) * nvec(p, q); znp += v * nvec(p, q);
                }
                const bool zzero = zn2 <= 1e-44Q * np2;
                Q t1 = -1; int l = -1;                                     // dual step limit
                for (int a = 0; a < ma; ++a) if (r[a] > 0) { const Q c2 = mult[a] / r[a]; if (t1 < 0 || c2 < t1) { t1 = c2; l = a; } }
                Q sp_now = -bval(p); for (int q = 0; q < nu; ++q) sp_now += nvec(p, q) * w[q];
                const Q t2 = zzero ? (Q)-1 : -sp_now / znp;                // full primal step (negative = infinite)
                if (t1 < 0 && t2 < 0) return 1;                            // infeasible QP
                if (t2 < 0 || (t1 >= 0 && t1 < t2)) {                      // partial step: drop constraint l
                    if (!zzero) for (int q = 0; q < nu; ++q) w[q] += t1 * z[q];
                    for (int a = 0; a < ma; ++a) mult[a] -= t1 * r[a];
                    up += t1;
                    W.erase(W.begin() + l); mult.erase(mult.begin() + l);
                    continue;
                }
                for (int q = 0; q < nu; ++q) w[q] += t2 * z[q];            // full step: constraint p becomes active
                for (int a = 0; a < ma; ++a) mult[a] -= t2 * r[a];
                up += t2;
                W.push_back(p); mult.push_back(up);
                break;
            }
        }
        if (done) {
            std::fill(ws.begin(), ws.end(), 0);
            for (int q : W) { if (q < m) ws[q] = -1; else ws[q - m] = 1; }
            if (!eqp(ws)) return 2;
            status = 0;
        }
    }
    if (iters) *iters = it;
    // ---- 3. recover x, costates, multipliers; certificate in __float128
    std::vector<Q> xs((size_t)(N + 1) * 4), lam((size_t)N * 6, 0), pi((size_t)N * 4);
    for (int k = 0; k <= N; ++k) for (int i = 0; i < 4; ++i) {
        Q a = X[k][(size_t)i * nc1 + nu];
        for (int p = 0; p < nu; ++p) a += X[k][(size_t)i * nc1 + p] * u[p];
        xs[(size_t)k * 4 + i] = a;
    }
    for (int k = 0; k < N; ++k) for (int c = 0; c < 3; ++c) act[k * 3 + c] = 0;
    for (int i = 0; i < m; ++i) {
        act[rows[i].k * 3 + rows[i].c] = ws[i];
        if (ws[i] < 0) lam[(size_t)rows[i].k * 6 + rows[i].c] = -nuv[i];
        if (ws[i] > 0) lam[(size_t)rows[i].k * 6 + 3 + rows[i].c] = nuv[i];
    }
    Q r_stat = 0, r_eq = 0, r_pr = 0, r_du = 0, r_cp = 0;
    {
        Q pn[4];
        for (int i = 0; i < 4; ++i) { Q a = qN[i]; for (int l = 0; l < 4; ++l) a += (Q)QN[4 * i + l] * xs[(size_t)N * 4 + l]; pn[i] = a; }
        for (int k = N - 1; k >= 0; --k) {
            for (int i = 0; i < 4; ++i) pi[(size_t)k * 4 + i] = pn[i];
            Q z[6] = {u[2 * k], u[2 * k + 1], xs[(size_t)k * 4], xs[(size_t)k * 4 + 1], xs[(size_t)k * 4 + 2], xs[(size_t)k * 4 + 3]};
            Q gr[6];
            for (int i = 0; i < 6; ++i) { Q a = g[(size_t)k * 6 + i]; for (int l = 0; l < 6; ++l) a += (Q)H[(size_t)k * 36 + 6 * i + l] * z[l]; gr[i] = a; }
            for (int j = 0; j < 2; ++j) for (int l = 0; l < 4; ++l) gr[j] += (Q)B[(size_t)k * 8 + 2 * l + j] * pn[l];
            for (int j = 0; j < 4; ++j) for (int l = 0; l < 4; ++l) gr[2 + j] += (Q)A[(size_t)k * 16 + 4 * l + j] * pn[l];
            for (int c = 0; c < 3; ++c) if (on[k * 3 + c]) {
                const Q dlm = lam[(size_t)k * 6 + 3 + c] - lam[(size_t)k * 6 + c];
                gr[ci[k * 3 + c]] += dlm; gr[5] += (Q)beta[k * 3 + c] * dlm;
            }
            r_stat = qmax(r_stat, qmax(qabs(gr[0]), qabs(gr[1])));
            for (int i = 0; i < 4; ++i) pn[i] = gr[2 + i];               // pi_k := x-stationarity of stage k solved for pi_k
            for (int i = 0; i < 4; ++i) {
                Q a = (Q)b[(size_t)k * 4 + i] - xs[(size_t)(k + 1) * 4 + i];
                for (int l = 0; l < 4; ++l) a += (Q)A[(size_t)k * 16 + 4 * i + l] * z[2 + l];
                a += (Q)B[(size_t)k * 8 + 2 * i] * z[0] + (Q)B[(size_t)k * 8 + 2 * i + 1] * z[1];
                r_eq = qmax(r_eq, qabs(a));
            }
        }
        for (int i = 0; i < 4; ++i) r_eq = qmax(r_eq, qabs(xs[i] - (Q)dx0[i]));
    }
    for (int i = 0; i < m; ++i) {
        const Q sl = val[i] - rows[i].lo, su = rows[i].hi - val[i];
        const Q ll = lam[(size_t)rows[i].k * 6 + rows[i].c], lu = lam[(size_t)rows[i].k * 6 + 3 + rows[i].c];
        r_pr = qmax(r_pr, qmax(-sl, -su));
        r_du = qmax(r_du, qmax(-ll, -lu));
        r_cp = qmax(r_cp, qmax(qabs(ll * sl), qabs(lu * su)));
    }
    kkt[0] = (double)r_stat; kkt[1] = (double)r_eq; kkt[2] = (double)r_pr; kkt[3] = (double)r_du; kkt[4] = (double)r_cp;
    for (int k = 0; k < N; ++k) { du_o[2 * k] = (double)u[2 * k]; du_o[2 * k + 1] = (double)u[2 * k + 1]; }
    for (size_t i = 0; i < xs.size(); ++i) dx_o[i] = (double)xs[i];
    for (size_t i = 0; i < pi.size(); ++i) pi_o[i] = (double)pi[i];
    for (size_t i = 0; i < lam.size(); ++i) lam_o[i] = (double)lam[i];
    return status;
}

}  // extern "C"
