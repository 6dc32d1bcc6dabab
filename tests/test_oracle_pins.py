"""Pins of the CPU oracle (the reference ships no golden vectors: SURVEY.md section 4 / 8c).

Spline restatement vs scipy.interpolate.BSpline and the constants probed by the survey; dynamics and its
AD Jacobian vs an independent sympy derivation; closed-form invariants; IEEE corner cases; the QP solver
vs a dense numpy KKT certificate.
"""
import os

import numpy as np
import pytest
import sympy as sp
from scipy.interpolate import BSpline

from oracle import oracle as orc
from tests.conftest import REFERENCE_CAD
from tests.workloads import OBJECT_ORDER, OBJECT_TABLE, make_rti_workload, oracle_model, packaged_tables

SURVEY = {  # SURVEY.md A1.1 / 8c [PROBE]: n, len(S), b, c_ellipse
    "santal": (37, 41, 0.281899, 0.027811), "balea": (37, 41, 0.222486, 0.007141),
    "montana": (35, 39, 0.285079, 0.020867), "pulirapid": (56, 60, 0.596013, 0.023260),
}


@pytest.mark.parametrize("name", OBJECT_ORDER)
def test_survey_constants(name):
    m = oracle_model(name)
    n, ns, b, c = SURVEY[name]
    assert (m.n, m.nknots) == (n, ns)
    assert abs(m.b - b) < 5e-7 and abs(m.c_ellipse - c) < 5e-7
    # every table value is a float32-representable number (pcread returns single)
    assert np.array_equal(m.S.astype(np.float32).astype(np.float64), m.S)
    assert np.array_equal(m.P.astype(np.float32).astype(np.float64), m.P)


def test_santal_anchor_points():
    m = oracle_model("santal")
    r = m.eval_spline([0.0])
    assert np.allclose(r["C"][0], [-0.031, 0.024463], atol=5e-7)
    assert np.allclose(r["t"][0], [0.0, 1.0], atol=1e-12) and np.allclose(r["n"][0], [1.0, 0.0], atol=1e-12)
    assert np.allclose(r["C"][0], m.P[0])                            # C(0) = P_1 (clamped)


@pytest.mark.skipif(not os.path.isdir(REFERENCE_CAD), reason="reference .ply files only exist in the build container")
@pytest.mark.parametrize("name", OBJECT_ORDER)
def test_ply_ingest_matches_packaged_tables(name):
    m = orc.model_from_reference_ply(name, REFERENCE_CAD)
    t = packaged_tables()[name]
    assert np.array_equal(m.S, np.array(t["knots"])) and np.array_equal(m.P, np.array(t["ctrl_xy"]))
    assert m.c_ellipse == t["c_ellipse"] and m.b == t["b"]


@pytest.mark.parametrize("name", OBJECT_ORDER)
def test_spline_vs_scipy(name):
    t = packaged_tables()[name]
    S, P = np.array(t["knots"]), np.array(t["ctrl_xy"])
    m = orc.Model.create(S, P, 3, 0.1, 0.02, single_quirk=False)    # exact double arithmetic = textbook B-spline
    s = np.linspace(0.0, m.b * (1 - 1e-13), 4001)
    r = m.eval_spline(s)
    bs = BSpline(S, P, 3)
    assert np.abs(bs(s) - r["C"]).max() < 1e-15
    assert np.abs(bs.derivative(1)(s) - r["Cd"]).max() < 1e-12 * np.abs(r["Cd"]).max()
    assert np.abs(bs.derivative(2)(s) - r["Cdd"]).max() < 1e-11 * np.abs(r["Cdd"]).max()
    # the MATLAB-single arithmetic only perturbs the tables at float32 level
    mq = oracle_model(name)
    rq = mq.eval_spline(s)
    assert np.abs(rq["C"] - r["C"]).max() < 1e-8 and np.abs(rq["Cd"] - r["Cd"]).max() < 1e-6 * np.abs(r["Cd"]).max()


@pytest.mark.parametrize("name", OBJECT_ORDER)
def test_full_sum_equals_local_sum_bitwise(name):
    m = oracle_model(name)
    rng = np.random.default_rng(0)
    s = np.concatenate([rng.uniform(0, m.b, 300), np.unique(m.S)[:-1]])
    a, b = m.eval_spline(s, local=True), m.eval_spline(s, local=False)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    x = np.stack([np.zeros_like(s), np.zeros_like(s), rng.uniform(-3, 3, len(s)), s], 1)
    u = np.stack([rng.uniform(1e-3, 0.03, len(s)), rng.uniform(-0.05, 0.05, len(s))], 1)
    fa, Ja, Ua = m.dynamics(x, u, jac=True, local=True)
    fb, Jb, Ub = m.dynamics(x, u, jac=True, local=False)
    assert np.array_equal(fa, fb) and np.array_equal(Ja, Jb) and np.array_equal(Ua, Ub)


def test_basis_invariants():
    m = orc.Model.create(np.array(packaged_tables()["santal"]["knots"]), np.array(packaged_tables()["santal"]["ctrl_xy"]), 3, 0.19, 0.0278, False)
    for s in np.linspace(0, m.b * 0.999, 57):
        N = [m.basis(s, i, 3) for i in range(1, m.n + 1)]
        assert abs(sum(N) - 1.0) < 1e-14 and min(N) >= 0.0         # partition of unity, non-negativity
    assert all(m.basis(m.b, i, 3) == 0.0 for i in range(1, m.n + 1))  # half-open indicator: everything is 0 at s = b


@pytest.mark.parametrize("name", OBJECT_ORDER)
def test_frames_and_derivatives(name):
    m = oracle_model(name)
    s = np.linspace(1e-4, m.b - 1e-4, 997)
    r = m.eval_spline(s)
    assert np.abs((r["t"] ** 2).sum(1) - 1).max() < 1e-14 and np.abs((r["t"] * r["n"]).sum(1)).max() < 1e-15
    assert np.allclose(r["n"], np.stack([r["t"][:, 1], -r["t"][:, 0]], 1))
    h = 1e-7
    fd = (m.eval_spline(s + h)["C"] - m.eval_spline(s - h)["C"]) / (2 * h)
    assert np.abs(fd - r["Cd"]).max() < 2e-6 * np.abs(r["Cd"]).max()
    kap = (r["Cd"][:, 0] * r["Cdd"][:, 1] - r["Cd"][:, 1] * r["Cdd"][:, 0]) / (r["Cd"] ** 2).sum(1)
    assert np.abs(kap - r["kappa"]).max() < 1e-4                   # AD of FC_dot vs the single-rounded cj_2 table


def _sympy_dynamics(S, P, c1, span_lo, b, mu, c):
    """Independent derivation: span polynomials from scipy's BSpline -> sympy expression -> symbolic Jacobian."""
    th, s, un, ut = sp.symbols("theta s u_n u_t", real=True)
    bsC, bsD = BSpline(S, P, 3), BSpline(S[1:-1], c1[1:], 2)       # C and FC_dot = sum c1_i N_{i,p-1} on the inner knots
    x0 = span_lo + 1e-4                                          # expansion point strictly inside the span
    tau = s - sp.Float(x0, 30)
    def poly(bs, deg):
        # exact Taylor coefficients of the span polynomial at x0 (scipy derivatives)
        return [sum(sp.Float(float(bs.derivative(d)(x0)[k]) if d else float(bs(x0)[k]), 30) / sp.factorial(d) * tau ** d for d in range(deg + 1)) for k in range(2)]
    Cx, Cy = poly(bsC, 3)
    Dx, Dy = poly(bsD, 2)
    nrm = sp.sqrt(Dx ** 2 + Dy ** 2)
    tx, ty = Dx / nrm, Dy / nrm
    nx, ny = ty, -tx
    px, py = nx * Cx + ny * Cy, tx * Cx + ty * Cy
    k = 1 / (c ** 2 + px ** 2 + py ** 2)
    gl = (mu * c ** 2 - px * py + mu * px ** 2) / (c ** 2 + py ** 2 - mu * px * py)
    gr = (-mu * c ** 2 - px * py - mu * px ** 2) / (c ** 2 + py ** 2 + mu * px * py)
    def branch(w, sdot):
        vn = k * ((c ** 2 + px ** 2) * un + px * py * w)
        vt = k * (px * py * un + (c ** 2 + py ** 2) * w)
        Vx, Vy = nx * vn + tx * vt, ny * vn + ty * vt
        return sp.Matrix([sp.cos(th) * Vx - sp.sin(th) * Vy, sp.sin(th) * Vx + sp.cos(th) * Vy, k * (-py * un + px * w), sdot])
    f = {"st": branch(ut, 0), "sl": branch(gl * un, ut - gl * un), "sr": branch(gr * un, ut - gr * un)}
    v = (th, s, un, ut)
    return {m: (sp.lambdify(v, e, "mpmath"), sp.lambdify(v, e.jacobian(v), "mpmath")) for m, e in f.items()}, sp.lambdify(v, [gl, gr], "mpmath")


def test_dynamics_and_jacobian_vs_sympy():
    t = packaged_tables()["santal"]
    S, P = np.array(t["knots"]), np.array(t["ctrl_xy"])
    mu, c = OBJECT_TABLE["santal"]["mu_sp"], t["c_ellipse"]
    m = orc.Model.create(S, P, 3, mu, c, single_quirk=False)
    rng = np.random.default_rng(3)
    knots = np.unique(S)
    for span in (2, 20):                                     # (each span costs ~30 s of symbolic differentiation)
        lo, hi = knots[span], knots[span + 1]
        funs, cone = _sympy_dynamics(S, P, m.c1, lo, m.b, mu, c)
        for _ in range(3):
            sv = rng.uniform(lo + 1e-5, hi - 1e-5); thv = rng.uniform(-3, 3); unv = rng.uniform(2e-3, 0.03)
            gl, gr = [float(v) for v in cone(thv, sv, unv, 0.0)]
            for mode, r in (("st", gr + (gl - gr) * rng.uniform(0.1, 0.9)), ("sl", gl + rng.uniform(0.1, 1.0)), ("sr", gr - rng.uniform(0.1, 1.0))):
                utv = r * unv
                f_ref = np.array(funs[mode][0](thv, sv, unv, utv), dtype=float).reshape(4)
                J_ref = np.array(funs[mode][1](thv, sv, unv, utv), dtype=float).reshape(4, 4)
                f, Jx, Ju = m.dynamics([[0.3, -0.2, thv, sv]], [[unv, utv]], jac=True)
                sc = max(np.abs(J_ref).max(), 1e-3)
                assert np.abs(f[0] - f_ref).max() < 1e-12 * max(np.abs(f_ref).max(), 1e-3), (span, mode)
                assert np.abs(Jx[0][:, 2:] - J_ref[:, :2]).max() < 1e-10 * sc and np.abs(Ju[0] - J_ref[:, 2:]).max() < 1e-10 * sc, (span, mode)
                assert np.all(Jx[0][:, :2] == 0.0)


def test_dynamics_ieee_corner_cases_and_modes():
    m = oracle_model("santal")
    x = np.array([[0.0, 0.0, 0.3, -0.01]])
    f, Jx, Ju = m.dynamics(x, [[0.0, 0.0]], jac=True)            # u_n = u_t = 0: r = NaN -> f = 0, zero Jacobian (cold start)
    assert np.all(f == 0) and np.all(Jx == 0) and np.all(Ju == 0)
    f = m.dynamics(x, [[0.0, 0.02]])                             # u_n = 0, u_t != 0: r = +Inf -> pure sliding, s_dot = u_t
    assert np.allclose(f[0], [0, 0, 0, 0.02], atol=0) 
    f = m.dynamics(x, [[0.01, 0.0]])                             # inside the cone: sticking, s_dot = 0
    assert f[0, 3] == 0.0 and abs(f[0, 0]) > 0
    f = m.dynamics(np.array([[0, 0, 0.3, -m.b]]), [[0.01, 0.0]])  # s = -b wraps to sigma = b: every basis function 0 -> NaN tangent
    assert np.all(np.isnan(f[0, :3]))
    # continuity of f across the cone edges (mode continuity)
    rng = np.random.default_rng(5)
    for _ in range(20):
        s = rng.uniform(-0.04, 0.005); th = rng.uniform(-1, 1); un = 0.01
        # locate gamma_l by bisection on the mode switch of s_dot
        lo, hi = 0.0, 5.0
        for _ in range(60):
            mid = 0.5 * (lo + hi)
            if m.dynamics([[0, 0, th, s]], [[un, mid * un]])[0, 3] == 0.0: lo = mid
            else: hi = mid
        fa = m.dynamics([[0, 0, th, s]], [[un, lo * un]])[0]; fb = m.dynamics([[0, 0, th, s]], [[un, hi * un]])[0]
        assert np.abs(fa - fb).max() < 1e-9


def test_erk4_sens_is_derivative_of_rk4_map():
    m = oracle_model("montana")
    rng = np.random.default_rng(11)
    x = np.array([0.01, -0.02, 0.4, 0.05]); u = np.array([0.012, 0.004])
    Phi, A, B = m.erk4_sens([x], [u], 0.05)
    h = 1e-6
    for j in range(4):
        e = np.zeros(4); e[j] = h
        fd = (m.erk4_sens([x + e], [u], 0.05)[0][0] - m.erk4_sens([x - e], [u], 0.05)[0][0]) / (2 * h)
        assert np.abs(fd - A[0][:, j]).max() < 1e-6 * max(1.0, np.abs(A[0]).max())
    for j in range(2):
        e = np.zeros(2); e[j] = h
        fd = (m.erk4_sens([x], [u + e], 0.05)[0][0] - m.erk4_sens([x], [u - e], 0.05)[0][0]) / (2 * h)
        assert np.abs(fd - B[0][:, j]).max() < 1e-6 * max(1.0, np.abs(B[0]).max())
    assert np.array_equal(A[0][:, :2], np.eye(4)[:, :2])          # df/dx = df/dy = 0


T_FLOOR = 4e-12          # 4 * qp_t_min of the oracle's IPM


def _dense_kkt_check(lin, sol, W, We, dt, lh, uh, x, u, x0bar, yref_e_unused=None):
    """Independent dense-algebra KKT certificate of one QP solution (convex QP: KKT <=> optimal)."""
    N = lin["A"].shape[0]
    perm = [4, 5, 0, 1, 2, 3]
    H = dt * W[np.ix_(perm, perm)]
    r_stat = r_eq = r_in = r_cp = 0.0
    dx, du, pi, lam = sol["dx"], sol["du"], sol["pi"], sol["lam"]
    for k in range(N):
        z = np.concatenate([du[k], dx[k]])
        BA = np.concatenate([lin["B"][k], lin["A"][k]], axis=1)
        g = H @ z + lin["g"][k] + BA.T @ pi[k]
        if k > 0:
            g[2:] -= pi[k - 1]
        h = np.array([x[k, 3], u[k, 0], u[k, 1]]); idx = [5, 0, 1]
        for c in range(3):
            if k == 0 and c == 0:
                continue
            g[idx[c]] += lam[k, 3 + c] - lam[k, c]
            v = z[idx[c]]; sl, su = v - (lh[c] - h[c]), (uh[c] - h[c]) - v
            # complementarity: a slack at its floor (qp_t_min = 1e-12, "t <= 4 t_min" = converged active pair) counts as zero
            r_in = max(r_in, -min(sl, 0), -min(su, 0)); r_cp = max(r_cp, lam[k, c] * max(sl - T_FLOOR, 0), lam[k, 3 + c] * max(su - T_FLOOR, 0))
            assert lam[k, c] >= 0 and lam[k, 3 + c] >= 0
        r_stat = max(r_stat, np.abs(g[:2]).max(), np.abs(g[2:]).max() if k > 0 else 0.0)
        r_eq = max(r_eq, np.abs(lin["A"][k] @ dx[k] + lin["B"][k] @ du[k] + lin["b"][k] - dx[k + 1]).max())
    r_stat = max(r_stat, np.abs(We @ dx[N] + lin["qN"] - pi[N - 1]).max())
    r_eq = max(r_eq, np.abs(dx[0] - (x0bar - x[0])).max())
    return r_stat, r_eq, r_in, r_cp


def test_qp_solution_kkt_certificate():
    om = oracle_model("santal")
    B, N, dt = 24, 40, 0.05
    wl = make_rti_workload(None, batch=B, N=N, seed=2)
    ocp = orc.Ocp(om, N, dt)
    pr = ocp.prepare(wl["x0"], np.zeros(B, dtype=np.int32), np.zeros((B, N + 1, 4)), wl["u_init"])
    lin = ocp.linearise(pr["x0"], wl["yref"], wl["yref_e"], pr["x"], pr["u"])
    q = ocp.qp(pr["x0"], wl["yref"], wl["yref_e"], pr["x"], pr["u"])
    assert (q["status"] == 0).all()
    W = np.diag([1.0, 1.0, 1e-3, 0.0, 1e-3, 1e-3]); We = np.diag([2e5, 2e5, 20.0, 0.0])
    for b in range(B):
        lb = {k: v[b] for k, v in lin.items()}
        sb = {k: q[k][b] for k in ("du", "dx", "pi", "lam")}
        r = _dense_kkt_check(lb, sb, W, We, dt, [-0.06, 0.0, -0.05], [0.011, 0.03, 0.05], pr["x"][b], pr["u"][b], pr["x0"][b])
        assert max(r[:3]) < 1.5e-11 and r[3] < 1e-17, (b, r)


def test_qp_solution_sensitivity_to_tolerance_is_documented_behaviour():
    """the QP of this OCP (input weight 5e-5 vs terminal weight 2e5, multipliers down to 1e-9) is sensitive to
    where the IPM stops.  At the reference's own QP tolerance (1e-6 on all four residuals, NMPC_controller.m:276) du is ~1e-2
    away from the solution; with complementarity at 1e-12 (the round-1 rule) two IPM paths still differ by ~1e-5 although
    the optimal cost agrees to 1e-13; with the end game of this round (complementarity 1e-18) the paths agree to 1e-8 —
    tests/test_qp_exact.py compares the point with the exact solution."""
    om = oracle_model("santal")
    B, N = 64, 40
    wl = make_rti_workload(None, batch=B, N=N, seed=2)
    pr = orc.Ocp(om, N, 0.05).prepare(wl["x0"], np.zeros(B, dtype=np.int32), np.zeros((B, N + 1, 4)), wl["u_init"])
    args = (pr["x0"], wl["yref"], wl["yref_e"], pr["x"], pr["u"])
    r01 = dict(qp_t_min=0.0, qp_gamma_f=0.0, qp_stall=5, qp_split_step=0)
    ref = orc.Ocp(om, N, 0.05).qp(*args, nthreads=4)
    loose = orc.Ocp(om, N, 0.05, qp_tol=1e-6, qp_tol_comp=1e-6, **r01).qp(*args, nthreads=4)
    assert np.abs(loose["du"] - ref["du"]).max() > 1e-3
    a = orc.Ocp(om, N, 0.05, qp_tol=1e-12, qp_tol_comp=1e-12, qp_mu0=1.0, **r01).qp(*args, nthreads=4)
    b = orc.Ocp(om, N, 0.05, qp_tol=1e-12, qp_tol_comp=1e-12, qp_mu0=1e-2, **r01).qp(*args, nthreads=4)
    d = np.abs(a["du"] - b["du"]).reshape(B, -1).max(1)            # round-1 rule: two iterate paths, same limit
    assert 1e-6 < d.max() < 1e-4 and np.median(d) < 1e-6, (d.max(), np.median(d))
    a = orc.Ocp(om, N, 0.05, qp_mu0=1.0).qp(*args, nthreads=4)
    b = orc.Ocp(om, N, 0.05, qp_mu0=1e-2).qp(*args, nthreads=4)
    assert np.abs(a["du"] - b["du"]).max() < 1e-8 and np.abs(a["dx"] - b["dx"]).max() < 1e-8
    assert np.abs(a["du"] - ref["du"]).max() < 1e-8


def test_golden_fixtures_match_oracle():
    here = os.path.join(os.path.dirname(__file__), "golden")
    for name in OBJECT_ORDER:
        g = np.load(os.path.join(here, f"eval_{name}.npz"))
        m = oracle_model(name)
        Phi, A, B = m.erk4_sens(g["x"], g["u"], 0.05)
        assert np.array_equal(np.nan_to_num(Phi, nan=7.0), np.nan_to_num(g["Phi"], nan=7.0))
        assert np.array_equal(np.nan_to_num(A, nan=7.0), np.nan_to_num(g["A"], nan=7.0))


def test_velocity_constraint_variant_derivative_and_kkt_certificate():
    """h_variant 1 (the parked constraint set h = [u_n; u_t -+ v_bound(s)], NMPC_controller.m:226-238): v_bound'(s)
    vs central differences, the symbolic v_bound vs the numeric one of update_tangential_velocity_bounds, and a
    dense KKT certificate of QP solutions whose coupled rows (ds and du_t in one inequality) are active."""
    from tests.workloads import VARIANT_LH, VARIANT_UH, make_vbound_workload
    om = oracle_model("santal")
    B, N, dt = 16, 40, 0.05
    ocp = orc.Ocp(om, N, dt); ocp.set_h_variant(1)
    nact = 0
    for s in np.linspace(-0.02, 0.02, 81):
        v, dv = ocp.v_bound_sym(s)
        assert abs(v - om.v_bound(s)[0]) < 1e-6 * max(v, 1e-3)        # numeric path wraps in float32 (MATLAB single)
        e = 1e-7
        vp, vm = ocp.v_bound_sym(s + e)[0], ocp.v_bound_sym(s - e)[0]
        if v < 0.05 and vp < 0.05 and vm < 0.05 and abs(dv) < 50:      # away from the min() kink and the 1e-4 pole
            assert abs(dv - (vp - vm) / (2 * e)) < 1e-4 * max(1.0, abs(dv)), (s, dv)
            nact += 1
        if v == 0.05:
            assert dv == 0.0
    assert nact >= 5
    wl = make_vbound_workload(B, N)
    pr = ocp.prepare(wl["x0"], np.zeros(B, dtype=np.int32), np.zeros((B, N + 1, 4)), wl["u_init"])
    h, beta = ocp.constraints(pr["x"], pr["u"])
    assert np.abs(beta).max() > 5.0 and np.array_equal(beta[:, :, 1], -beta[:, :, 2]) and not beta[:, :, 0].any()
    lin = ocp.linearise(pr["x0"], wl["yref"], wl["yref_e"], pr["x"], pr["u"])
    q = ocp.qp(pr["x0"], wl["yref"], wl["yref_e"], pr["x"], pr["u"])
    assert (q["status"] == 0).all() and q["iters"].max() <= 25
    W = np.diag([1.0, 1.0, 1e-3, 0.0, 1e-3, 1e-3]); We = np.diag([2e5, 2e5, 20.0, 0.0])
    perm = [4, 5, 0, 1, 2, 3]; H = dt * W[np.ix_(perm, perm)]
    worst, active = 0.0, 0
    for b in range(B):
        dx, du, pi, lam = q["dx"][b], q["du"][b], q["pi"][b], q["lam"][b]
        for k in range(N):
            z = np.concatenate([du[k], dx[k]]); BA = np.concatenate([lin["B"][b, k], lin["A"][b, k]], axis=1)
            g = H @ z + lin["g"][b, k] + BA.T @ pi[k]
            if k > 0:
                g[2:] -= pi[k - 1]
            for c in range(3):
                a = np.zeros(6); a[0 if c == 0 else 1] = 1.0; a[5] = beta[b, k, c]
                g += a * (lam[k, 3 + c] - lam[k, c]); v = a @ z
                sl, su = v - (VARIANT_LH[c] - h[b, k, c]), (VARIANT_UH[c] - h[b, k, c]) - v
                worst = max(worst, -min(sl, 0), -min(su, 0), lam[k, c] * max(sl - T_FLOOR, 0), lam[k, 3 + c] * max(su - T_FLOOR, 0))
                active += int(c > 0 and max(lam[k, c], lam[k, 3 + c]) > 1e-3 and k > 0 and abs(beta[b, k, c]) > 1.0)
            worst = max(worst, np.abs(g[:2]).max(), np.abs(g[2:]).max() if k > 0 else 0.0)
            worst = max(worst, np.abs(lin["A"][b, k] @ dx[k] + lin["B"][b, k] @ du[k] + lin["b"][b, k] - dx[k + 1]).max())
        worst = max(worst, np.abs(We @ dx[N] + lin["qN"][b] - pi[N - 1]).max())
    assert worst < 1.5e-11 and active >= 10, (worst, active)


def test_feasible_start_variant_of_config5_converges():
    """Full SQP (N = 100, merit backtracking, <= 30 iterations) from the feasible start on the symmetric outline converges for
    >= 90 % of the instances; from the mixed sticking / sliding start of config 5 the three asymmetric outlines end at the iteration
    limit (the minimiser sits on the u_n = 0 kink) and, with BLASFEO's pivot rule, no instance ends as a QP failure."""
    from tests.workloads import make_feasible_start_workload
    N, B = 100, 64
    wl = make_feasible_start_workload(B, N)
    ocp = orc.Ocp(oracle_model("balea"), N, 0.05)
    pr = ocp.prepare(wl["x0"], np.zeros(B, dtype=np.int32), np.zeros((B, N + 1, 4)), wl["u_init"])
    r = ocp.solve("sqp", pr["x0"], wl["yref"], wl["yref_e"], pr["x"], pr["u"], nthreads=8)
    assert (r["status"] == 0).mean() >= 0.9 and set(np.unique(r["status"])) <= {0, 2}
    wl = make_rti_workload(None, batch=32, N=N, seed=4, mixed_modes=True)
    for name, pivot_fails, expect4 in (("pulirapid", 0, False), ("pulirapid", 1, True)):
        ocp = orc.Ocp(oracle_model(name), N, 0.05, sem_qp_pivot_fails=pivot_fails)
        pr = ocp.prepare(wl["x0"], np.zeros(32, dtype=np.int32), np.zeros((32, N + 1, 4)), wl["u_init"])
        r = ocp.solve("sqp", pr["x0"], wl["yref"], wl["yref_e"], pr["x"], pr["u"], nthreads=8)
        assert ((r["status"] == 4).mean() > 0.1) == expect4, (name, pivot_fails, np.bincount(r["status"], minlength=5))
