"""Plain-C replay of the MEX call sequence of NMPC_controller.solve / helper.closed_loop_matlab against include/qspush.h
(tests/capi_replay.c, compiled by gcc in __graft_entry__.build()): batch 1, one qspush_set(QSPUSH_YREF, k, ...) per stage,
the terminal reference with stage = Hp exactly like NMPC_controller.m:343-348, qspush_opts filled from C.

  * CPU: the struct layout gcc sees (sizeof / offsets of qspush_opts, qspush_ctrl, qspush_loop_opts) equals the hand-written
    ctypes mirror of uclv_qs_pushing_matlab_b200/_lib.py;
  * GPU: config 1 (main.m: santal, x0 = 0, Hp = 10, 201 control periods) through the C program, bit-for-bit equal to the same
    call sequence issued from Python through ctypes, and within 1e-9 of the NMPC_controller mirror (whose pre-processing runs
    in k_prepare on the device instead of the host loop over qspush_eval_dynamics).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "capi_replay")


def _exe():
    if not os.path.exists(EXE) or os.path.getmtime(os.path.join(ROOT, "tests", "capi_replay.c")) > os.path.getmtime(EXE):
        lib = os.path.join(ROOT, "uclv_qs_pushing_matlab_b200")
        subprocess.check_call(["gcc", "-O1", "-std=c99", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "capi_replay.c"),
                               "-L" + lib, "-lqspush", "-Wl,-rpath," + lib, "-Wl,-rpath,$ORIGIN/../uclv_qs_pushing_matlab_b200", "-lm", "-o", EXE])
    return EXE


def test_struct_layout_seen_by_gcc_matches_ctypes_mirror():
    from uclv_qs_pushing_matlab_b200 import _lib as L
    out = subprocess.run([_exe(), "--layout"], capture_output=True, text=True, check=True).stdout.split()
    got = [int(v) for v in out]
    O = L.Opts
    want = [C.sizeof(O), O.qp_max_iter.offset, O.qp_tau.offset, O.globalization.offset, O.h_variant.offset, O.qp_tol_comp.offset,
            O.qp_stall.offset, C.sizeof(L.Ctrl), C.sizeof(L.LoopOpts)]
    assert got == want, (got, want)


def _config1():
    from uclv_qs_pushing_matlab_b200.workloads import packaged_tables
    from uclv_qs_pushing_matlab_b200.object_selection import OBJECT_TABLE
    t = packaged_tables()["santal"]
    Hp, dt, steps, T = 10, 0.05, 201, 201
    tt = np.arange(T) * dt
    yref = np.zeros((T, 6)); yref[:, 0] = np.minimum(0.01 * tt, 0.10)                     # main.m:150-178 (straight line stand-in)
    W = np.diag([1.0, 1.0, 1e-3, 0.0, 1e-3, 1e-3]); We = np.diag([2e5, 2e5, 20.0, 0.0])   # main.m:82-84
    return dict(N=Hp, dt=dt, steps=steps, T=T, knots=np.asarray(t["knots"], dtype=np.float64), ctrl=np.asarray(t["ctrl_xy"], dtype=np.float64),
                mu_sp=OBJECT_TABLE["santal"]["mu_sp"], c_ellipse=t["c_ellipse"], W=W, We=We, yref=yref, x0=np.zeros(4))


def _run_c(cfg, mode, tmp):
    fi, fo = os.path.join(tmp, f"in{mode}.bin"), os.path.join(tmp, f"out{mode}.bin")
    head = np.array([cfg["N"], cfg["dt"], mode, cfg["steps"], cfg["T"], len(cfg["knots"]), len(cfg["ctrl"]), cfg["mu_sp"], cfg["c_ellipse"]])
    np.concatenate([head, cfg["knots"], cfg["ctrl"].ravel(), np.asfortranarray(cfg["W"]).ravel(order="F"), np.asfortranarray(cfg["We"]).ravel(order="F"),
                    cfg["yref"].ravel(), cfg["x0"]]).astype(np.float64).tofile(fi)
    subprocess.run([_exe(), fi, fo], check=True)
    raw = np.fromfile(fo, dtype=np.float64)
    return raw[:8], raw[8:].reshape(cfg["steps"], 10)


def _run_py(cfg, mode):
    """The same call sequence through ctypes (numpy arithmetic on the host side is IEEE double like the C program's)."""
    import uclv_qs_pushing_matlab_b200 as q
    from uclv_qs_pushing_matlab_b200 import _lib as L
    lib = L.lib()
    N, dt, T = cfg["N"], cfg["dt"], cfg["T"]
    m = q.Model.from_tables(cfg["knots"], cfg["ctrl"], 3, cfg["mu_sp"], cfg["c_ellipse"], True)
    s = q.Solver([m], N, dt, 1, mode=mode, max_sqp_iter=30, globalization=1)
    for k in range(N):
        s.set("W", cfg["W"], stage=k)
    s.set("W", cfg["We"], stage=N)
    cc = q.default_ctrl()
    bb = np.float32(m.b)
    x = cfg["x0"].copy(); rows = []
    X = np.zeros((N + 1, 4)); U = np.zeros((N, 2)); P = np.zeros((N, 4)); cold = True
    for i in range(1, cfg["steps"] + 1):
        x0 = x.copy()
        xs, r = np.float32(x0[3]), np.float32(0.0)
        if xs != 0:
            r = np.fmod(xs, bb)
            qv = np.abs(xs / bb)
            if r == 0 or not (np.abs(qv - np.floor(qv + np.float32(0.5))) > np.float32(1.1920929e-7) * qv):
                r = np.float32(0.0)
            elif (xs < 0) != (bb < 0):
                r = np.float32(r + bb)
        x0[3] = float(np.float32(r - bb * np.float32(1.0 if x0[3] < 0 else 0.0)))
        s.set("x0", x0.reshape(1, 4))
        for k in range(N):
            s.set("yref", cfg["yref"][min(i + k, T) - 1].reshape(1, 6), stage=k)
        L.check(lib.qspush_set(s._h, L.YREF_E, N, 0, 1, np.ascontiguousarray(cfg["yref"][min(i + N - 1, T) - 1]).ctypes.data, L.MEM_HOST))
        if cold:
            X[:] = 0; P[:] = 0; U[:, 0] = cc.u_n_lb; U[:, 1] = 0.0; cold = False
        vb = m.eval_v_bound([x0[3]], cc)[0][0]
        X[0] = x0
        for k in range(N):
            if abs(U[k, 1]) > vb:
                old = U[k, 1]
                U[k, 1] = float((old > 0) - (old < 0)) * vb
                U[k, 0] = U[k, 1] * U[k, 0] / old
            f = m.eval_dynamics(X[k], U[k])[0]
            X[k + 1] = X[k] + dt * f
            if k + 1 < N:
                vb = m.eval_v_bound([X[k + 1, 3]], cc)[0][0]
        s.set("x", X.reshape(1, N + 1, 4)); s.set("u", U.reshape(1, N, 2)); s.set("pi", P.reshape(1, N, 4))
        s.solve()
        U = s.get("u")[0].copy(); X = s.get("x")[0].copy(); P = s.get("pi")[0].copy()
        u0 = s.get("u", stage=0)[0]; cost = s.get("cost")[0]
        st, it = s.get_int("status")[0], s.get_int("sqp_iter")[0]
        X[:-1] = X[1:].copy(); U[:-1] = U[1:].copy(); P[:-1] = P[1:].copy()
        rows.append([*x, *u0, st, it, cost, x0[3]])
        x = x + dt * m.eval_dynamics(x, u0)[0]
    return np.array(rows, dtype=np.float64)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1], ids=["sqp_rti", "sqp"])
def test_c_replay_of_config1_is_the_python_sequence_bit_for_bit(tmp_path, mode):
    import uclv_qs_pushing_matlab_b200 as q
    from uclv_qs_pushing_matlab_b200 import _lib as L
    cfg = _config1()
    layout, rows = _run_c(cfg, mode, str(tmp_path))
    O = L.Opts
    assert [int(v) for v in layout] == [C.sizeof(O), O.qp_max_iter.offset, O.qp_tau.offset, O.globalization.offset, O.h_variant.offset,
                                        O.qp_tol_comp.offset, O.qp_stall.offset, C.sizeof(L.Ctrl)]
    py = _run_py(cfg, mode)
    assert rows.shape == py.shape == (201, 10)
    assert np.array_equal(rows, py)                               # states, u0, status, sqp_iter, cost of all 201 periods: bit for bit
    if mode == 0:
        assert (rows[:, 6] == 0).all() and (rows[:, 7] == 1).all()
    else:
        assert set(np.unique(rows[:, 6])) <= {0.0, 2.0, 3.0, 4.0} and rows[:, 7].max() <= 30
    assert abs(rows[-1, 0] - 0.10) < 3e-3                         # the slider arrives at the end of the 0.10 m reference (main.m:150-178)
    # against the NMPC_controller mirror (pre-processing on the device, k_prepare): same control law
    if mode == 0:
        sel = q.object_selection("santal")
        p = q.PusherSliderModel("real_plant", sel, 0, sel.cad_model_path, 3, sel.pcl_path, "santal")
        p.symbolic_model_variable_shape()
        c = q.NMPC_controller("NMPC", p, cfg["dt"], cfg["N"], nlp_solver="sqp_rti")
        c.create_ocp_solver(); c.set_delay_comp(0.0); c.initial_condition_update(np.zeros(4))
        c.set_reference_trajectory(cfg["yref"].T)
        out = q.helper.closed_loop_matlab(p, c, np.zeros(4), 200 * cfg["dt"])
        assert np.abs(out[6] - rows[:, 4]).max() < 1e-9 and np.abs(out[7] - rows[:, 5]).max() < 1e-9 and np.abs(out[0] - rows[:, 0]).max() < 1e-9
