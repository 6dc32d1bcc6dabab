"""Loader / replayer / bisector for golden vectors recorded from the REAL reference (tools/acados_golden.m).

A golden file holds `cases`: one struct per NMPC_controller.solve call with everything the call saw (x0, index_time, y_ref,
utraj / xtraj / ptraj before the call, weights, bounds) and everything acados returned (u0, x, u, pi, status, sqp_iter, cost).
`replay` runs the CPU oracle on the recorded inputs; `compare` measures the disagreement; `bisect` flips the oracle's recalled-
semantics switches (oracle/qs_oracle.hpp `sem_*`) one at a time and ranks the flips by the mismatch they leave.

Test infrastructure only (imports oracle/).
"""
from __future__ import annotations

import numpy as np
from scipy.io import loadmat, savemat

from oracle import oracle as orc
from uclv_qs_pushing_matlab_b200.object_selection import OBJECT_TABLE
from uclv_qs_pushing_matlab_b200.workloads import packaged_tables

# harness-level recalled choice: what acados keeps in `lam` between two solves of a closed loop (the reference never sets init_lam)
LAM_CARRY = ("keep", "shift", "zero")
TOL = dict(u0=1e-6, x=1e-6, u=1e-6, pi=1e-4)          # north_star: QP solution and u0 within 1e-6 (pi scales with the 2e5 terminal weight)


def _mat(a, shape=None):
    a = np.asarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def load_cases(path):
    """-> list of dicts (numpy arrays, MATLAB column-major shapes kept: x (4, Hp+1), u (2, Hp), pi (4, Hp), y_ref (6, T))."""
    m = loadmat(path, squeeze_me=True, struct_as_record=False)
    raw = np.atleast_1d(m["cases"])
    out = []
    for c in raw:
        d = {k: getattr(c, k) for k in c._fieldnames}
        Hp = int(d["Hp"])
        d["Hp"] = Hp
        d["name"], d["object"], d["nlp"] = str(d["name"]), str(d["object"]), str(d["nlp"])
        d["first_call"] = bool(d["first_call"])
        for k, shp in (("x", (4, Hp + 1)), ("u", (2, Hp)), ("pi", (4, Hp)), ("W", (6, 6)), ("We", (4, 4))):
            d[k] = _mat(d[k], shp)
        d["y_ref"] = _mat(d["y_ref"]).reshape(6, -1)
        for k, shp in (("utraj_in", (2, Hp)), ("xtraj_in", (4, Hp + 1)), ("ptraj_in", (4, Hp))):
            d[k] = None if d["first_call"] or np.size(d[k]) == 0 else _mat(d[k], shp)
        for k in ("x0", "u0", "lh", "uh", "knots"):
            d[k] = _mat(d[k]).ravel()
        d["ctrl"] = _mat(d["ctrl"])
        out.append(d)
    return out


def save_cases(path, cases, nlp):
    """Writes the same layout tools/acados_golden.m writes (used by the harness self-test with oracle-made stand-ins)."""
    arr = np.empty(len(cases), dtype=object)
    for i, c in enumerate(cases):
        arr[i] = {k: (np.zeros((0, 0)) if v is None else v) for k, v in c.items()}
    savemat(path, {"cases": arr, "nlp": nlp}, do_compression=True)


def oracle_model(name):
    t = packaged_tables()[name]
    return orc.Model.create(t["knots"], t["ctrl_xy"], 3, OBJECT_TABLE[name]["mu_sp"], t["c_ellipse"], True)


def check_tables(case):
    """a3/a4 pin: the outline tables the reference built from its .ply equal the packaged ones bit for bit."""
    t = packaged_tables()[case["object"]]
    ctrl = case["ctrl"] if case["ctrl"].shape[1] == 2 else case["ctrl"].T
    return (np.array_equal(case["knots"], np.asarray(t["knots"], dtype=np.float64))
            and np.array_equal(ctrl, np.asarray(t["ctrl_xy"], dtype=np.float64)) and float(case["b"]) == t["b"])


def _window(case):
    """NMPC_controller.m:307-313, 343-348: columns index_time .. index_time + Hp - 1, clamped at the last one."""
    Hp, T = case["Hp"], case["y_ref"].shape[1]
    cols = np.minimum(int(case["index_time"]) + np.arange(Hp), T) - 1
    yref = case["y_ref"][:, cols].T.copy()
    return yref[None], yref[None, Hp - 1, :4].copy()


def replay(cases, lam_carry="keep", **sem):
    """Oracle answer for every recorded solve.  x / u / pi warm starts are the RECORDED ones, so a disagreement does not
    accumulate along a closed loop; only lam is carried from the oracle's own previous solve of the same loop."""
    out, ocps, lam_prev, prev_group = [], {}, None, None
    for c in cases:
        Hp, key = c["Hp"], (c["object"], c["Hp"], float(c["dt"]), c["nlp"], c["W"].tobytes(), c["We"].tobytes(), c["lh"].tobytes(), c["uh"].tobytes())
        if key not in ocps:
            o = orc.Ocp(oracle_model(c["object"]), Hp, float(c["dt"]), **sem)
            for k in range(Hp):
                o.set_W(k, c["W"])
            o.set_W(Hp, c["We"])
            o.set_bounds(c["lh"], c["uh"])
            ocps[key] = o
        o = ocps[key]
        group = c["name"].rsplit("_", 1)[0]
        in_loop = "_loop" in c["name"] and group == prev_group and not c["first_call"]
        lam = np.zeros((1, Hp, 6))
        if in_loop and lam_prev is not None and lam_carry != "zero":
            lam = lam_prev.copy()
            if lam_carry == "shift":
                lam = np.concatenate([lam[:, 1:], lam[:, -1:]], axis=1)
        cold = np.array([1 if c["first_call"] else 0], dtype=np.int32)
        x = np.zeros((1, Hp + 1, 4)) if c["xtraj_in"] is None else c["xtraj_in"].T[None].copy()
        u = np.zeros((1, Hp, 2)) if c["utraj_in"] is None else c["utraj_in"].T[None].copy()
        pi = np.zeros((1, Hp, 4)) if c["ptraj_in"] is None else c["ptraj_in"].T[None].copy()
        p = o.prepare(c["x0"][None], cold, x, u, pi, lam)
        yref, yref_e = _window(c)
        r = o.solve("rti" if c["nlp"] == "sqp_rti" else "sqp", p["x0"], yref, yref_e, p["x"], p["u"], p["pi"], p["lam"])
        lam_prev, prev_group = r["lam"], group
        out.append(dict(u0=r["u"][0, 0].copy(), x=r["x"][0].T.copy(), u=r["u"][0].T.copy(), pi=r["pi"][0].T.copy(),
                        status=int(r["status"][0]), sqp_iter=int(r["sqp_iter"][0]), cost=float(r["cost"][0])))
    return out


def replay_product(cases):
    """The same replay through the PRODUCT (C-ABI, GPU), default semantics only: one batch-1 solver per configuration, the
    recorded warm start set field by field like NMPC_controller.solve does (constr_x0, cost_y_ref per window, init_u / init_pi,
    prepare, solve).  lam is whatever the solver keeps between solves, as in acados."""
    import uclv_qs_pushing_matlab_b200 as q
    from uclv_qs_pushing_matlab_b200.workloads import packaged_model
    out, solvers, prev_group = [], {}, None
    for c in cases:
        Hp = c["Hp"]
        key = (c["object"], Hp, float(c["dt"]), c["nlp"], c["W"].tobytes(), c["We"].tobytes(), c["lh"].tobytes(), c["uh"].tobytes())
        if key not in solvers:
            s = q.Solver([packaged_model(c["object"])], Hp, float(c["dt"]), 1, mode=0 if c["nlp"] == "sqp_rti" else 1)
            for k in range(Hp):
                s.set("W", np.asfortranarray(c["W"]), stage=k)
            s.set("W", np.asfortranarray(c["We"]), stage=Hp)
            s.set("lh", c["lh"]); s.set("uh", c["uh"])
            solvers[key] = s
        s = solvers[key]
        yref, yref_e = _window(c)
        s.set("x0", c["x0"][None]); s.set("yref", yref); s.set("yref_e", yref_e)
        s.set_int("cold", np.array([1 if c["first_call"] else 0], dtype=np.int32))
        if not c["first_call"]:
            s.set("u", c["utraj_in"].T[None].copy()); s.set("pi", c["ptraj_in"].T[None].copy())
        group = c["name"].rsplit("_", 1)[0]
        if not ("_loop" in c["name"] and group == prev_group and not c["first_call"]):
            s.set("lam", np.zeros((1, Hp, 6)))                   # independent solves start from zero multipliers like replay()
        prev_group = group
        s.prepare(); s.solve()
        u, x, pi = s.get("u")[0], s.get("x")[0], s.get("pi")[0]
        out.append(dict(u0=u[0].copy(), x=x.T.copy(), u=u.T.copy(), pi=pi.T.copy(), status=int(s.get_int("status")[0]),
                        sqp_iter=int(s.get_int("sqp_iter")[0]), cost=float(s.get("cost")[0])))
    return out


def compare(cases, answers):
    """-> dict of worst-case disagreements over the cases acados solved (status 0) + integer agreement rates."""
    err = dict(u0=0.0, x=0.0, u=0.0, pi=0.0)
    err_all = dict(err)                              # every case, also the ones acados left at an iteration / step limit
    st_eq = it_eq = n_ok = 0
    for c, a in zip(cases, answers):
        st_eq += int(int(c["status"]) == a["status"])
        it_eq += int(int(c["sqp_iter"]) == a["sqp_iter"])
        n_ok += int(int(c["status"]) == 0)
        for k in err:
            e = float(np.nan_to_num(np.abs(np.asarray(c[k]).reshape(a[k].shape) - a[k]), nan=np.inf).max())
            err_all[k] = max(err_all[k], e)
            if int(c["status"]) == 0:
                err[k] = max(err[k], e)
    n = max(len(cases), 1)
    return dict(err=err, err_all=err_all, status_agree=st_eq / n, sqp_iter_agree=it_eq / n, solved=n_ok, cases=len(cases))


def score(cmp_):
    """One number to rank alternatives: worst tolerance-normalised error over ALL cases (unconverged iterates are evidence too),
    plus penalties for status / iteration-count disagreement."""
    return max(min(cmp_["err_all"][k] / TOL[k], 1e12) for k in TOL) + 1e3 * (1.0 - cmp_["status_agree"]) + (1.0 - cmp_["sqp_iter_agree"])


def passes(cmp_, need_iter_agreement=False):
    ok = all(cmp_["err"][k] <= TOL[k] for k in TOL) and cmp_["status_agree"] == 1.0
    return ok and (cmp_["sqp_iter_agree"] == 1.0 or not need_iter_agreement)


def bisect(cases):
    """Flip every recalled-semantics switch on its own (and the lam carry of the harness); return rows sorted by score:
    [(label, score, compare-dict)], the first row being the best single flip; row 'defaults' is the unflipped oracle."""
    rows = [("defaults", *(lambda c: (score(c), c))(compare(cases, replay(cases))))]
    for name, alts in orc.SEMANTIC_SWITCHES.items():
        for v in alts:
            c = compare(cases, replay(cases, **{name: v}))
            rows.append((f"{name}={v}", score(c), c))
    for lc in LAM_CARRY[1:]:
        c = compare(cases, replay(cases, lam_carry=lc))
        rows.append((f"lam_carry={lc}", score(c), c))
    rows.sort(key=lambda r: r[1])
    return rows


def report(rows):
    lines = ["flip                         score      u0        x         u         pi        status  sqp_iter   (errors over all cases)"]
    for label, sc, c in rows:
        e = c["err_all"]
        lines.append(f"{label:<28} {sc:9.3g}  {e['u0']:.2e}  {e['x']:.2e}  {e['u']:.2e}  {e['pi']:.2e}  {c['status_agree']:.3f}   {c['sqp_iter_agree']:.3f}")
    return "\n".join(lines)


def make_standin(nlp, sem, n_single=4, loop_steps=6, Hp=10, dt=0.05, obj="santal"):
    """Oracle-made file in the golden layout with the semantics `sem` — a STAND-IN for acados, used only to test that the
    loader / replayer / bisector work (a file made this way pins nothing)."""
    from uclv_qs_pushing_matlab_b200.workloads import make_rti_workload
    W = np.diag([1.0, 1.0, 1e-3, 0.0, 1e-3, 1e-3]); We = np.diag([2e5, 2e5, 20.0, 0.0])
    lh = np.array([-0.06, 0.0, -0.05]); uh = np.array([0.011, 0.03, 0.05])          # NMPC_controller.m:23-26, 251-252
    t = packaged_tables()[obj]
    base = dict(nlp=nlp, object=obj, Hp=float(Hp), dt=dt, W=W, We=We, lh=lh, uh=uh, b=t["b"], knots=np.asarray(t["knots"]),
                ctrl=np.asarray(t["ctrl_xy"]))
    cases = []
    wl = make_rti_workload(n_single, Hp, dt=dt, seed=2)
    for k in range(n_single):
        cases.append(dict(base, name=f"standin_single_{k + 1}", x0=wl["x0"][k], index_time=1.0, y_ref=wl["yref"][k].T, first_call=False,
                          utraj_in=wl["u_init"][k].T, xtraj_in=np.zeros((4, Hp + 1)), ptraj_in=np.zeros((4, Hp))))
    T = loop_steps + 4
    y = np.zeros((6, T)); y[0] = 0.01 * np.arange(T) * dt
    loop_proto = dict(base, y_ref=y)
    done = replay(_finalise(cases), **sem)
    for c, a in zip(cases, done):
        c.update(u0=a["u0"], x=a["x"], u=a["u"], pi=a["pi"], status=float(a["status"]), sqp_iter=float(a["sqp_iter"]), cost=a["cost"])
    # closed loop with the stand-in's own semantics: the recorded warm starts must be what THAT solver produced
    x = np.zeros(4); xt = ut = pt = None
    om = oracle_model(obj)
    for i in range(loop_steps):
        c = dict(loop_proto, name=f"standin_loop_{i + 1}", x0=x.copy(), index_time=float(i + 1), first_call=(i == 0),
                 utraj_in=ut, xtraj_in=xt, ptraj_in=pt)
        prior = [k for k in cases if "_loop" in k["name"]]
        a = replay(_finalise(prior + [c]), **sem)[-1]          # re-run the loop so lam is carried the stand-in's way
        c.update(u0=a["u0"], x=a["x"], u=a["u"], pi=a["pi"], status=float(a["status"]), sqp_iter=float(a["sqp_iter"]), cost=a["cost"])
        cases.append(c)
        sh = lambda m: np.concatenate([m[:, 1:], m[:, -1:]], axis=1)        # noqa: E731  NMPC_controller.m:397-399
        ut, xt, pt = sh(a["u"]), sh(a["x"]), sh(a["pi"])
        x = x + dt * om.dynamics(x[None], a["u0"][None])[0]                 # helper.m:294, 307
    return cases


def _finalise(cases):
    out = []
    for c in cases:
        d = dict(c)
        d["Hp"] = int(d["Hp"])
        d["x0"] = np.asarray(d["x0"], dtype=np.float64).ravel()
        out.append(d)
    return out
