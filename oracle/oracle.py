"""ctypes binding of oracle/libqs_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.  PARITY UNPINNED against acados v0.2.1
(see qs_oracle.hpp); pinned against scipy / sympy / dense-KKT certificates by tests/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libqs_oracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    """Compile the oracle with g++ (make -C oracle)."""
    src_newer = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("qs_oracle.cpp", "qs_oracle.hpp")
    )
    if force or src_newer:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None
_build = "parity"


def select_build(which: str) -> None:
    """'parity' (default): libqs_oracle.so, -O2 without FMA contraction — the arithmetic the golden fixtures were frozen with.
    'fast': libqs_oracle_fast.so, the same source at -O3 -march=x86-64-v3 — the timed CPU baseline of bench.py.
    Must be called before the library is first used."""
    global _build
    if _lib is not None and which != _build:
        raise RuntimeError("oracle library already loaded")
    if which not in ("parity", "fast"):
        raise ValueError(which)
    _build = which


def lib():
    global _lib
    if _lib is None:
        build()
        path = _LIB_PATH
        if _build == "fast":
            path = os.path.join(_HERE, "libqs_oracle_fast.so")
            if not os.path.exists(path) or os.path.getmtime(os.path.join(_HERE, "qs_oracle.cpp")) > os.path.getmtime(path):
                subprocess.check_call(["make", "-C", _HERE, "-s", "libqs_oracle_fast.so"])
        L = C.CDLL(path)
        L.orc_model_create.restype = C.c_void_p
        L.orc_model_create.argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
        L.orc_model_from_ply.restype = C.c_void_p
        L.orc_model_from_ply.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_model_free.argtypes = [C.c_void_p]
        L.orc_model_info.argtypes = [C.c_void_p, _ip, _ip, _dp, _dp, _dp]
        L.orc_model_tables.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
        L.orc_basis.restype = C.c_double
        L.orc_basis.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int]
        L.orc_eval_spline.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, C.c_int] + [_dp] * 6
        L.orc_get_curvatures.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.orc_dynamics.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int, _dp, _dp, _dp]
        L.orc_erk4_sens.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_double, C.c_int, C.c_int, _dp, _dp, _dp]
        L.orc_v_bound.restype = C.c_double
        L.orc_v_bound.argtypes = [C.c_void_p] + [C.c_double] * 5 + [_dp]
        L.orc_ocp_create.restype = C.c_void_p
        L.orc_ocp_create.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.orc_ocp_free.argtypes = [C.c_void_p]
        L.orc_ocp_set_W.argtypes = [C.c_void_p, C.c_int, _dp]
        L.orc_ocp_set_bounds.argtypes = [C.c_void_p, _dp, _dp]
        L.orc_ocp_set_opts.argtypes = [C.c_void_p, _dp]
        L.orc_ocp_set_ctrl.argtypes = [C.c_void_p] + [C.c_double] * 5
        L.orc_ocp_set_h_variant.argtypes = [C.c_void_p, C.c_int] + [C.c_double] * 4
        L.orc_v_bound_sym.argtypes = [C.c_void_p, C.c_double, _dp]
        L.orc_v_bound_sym.restype = C.c_double
        L.orc_constraints_batch.argtypes = [C.c_void_p, C.c_int] + [_dp] * 4
        L.orc_linearise_batch.argtypes = [C.c_void_p, C.c_int] + [_dp] * 10
        L.orc_qp_batch.argtypes = [C.c_void_p, C.c_int] + [_dp] * 5 + [C.c_int] + [_dp] * 5 + [_ip, _ip, _dp]
        L.orc_qp_data_batch.argtypes = [C.c_void_p, C.c_int] + [_dp] * 5 + [_dp] * 11 + [_ip, _ip]
        L.orc_solve_batch.argtypes = [C.c_void_p, C.c_int, C.c_int] + [_dp] * 7 + [C.c_int, _ip, _dp]
        L.orc_prepare_batch.argtypes = [C.c_void_p, C.c_int, _dp, _ip] + [_dp] * 4 + [C.c_int]
        L.orc_shift_batch.argtypes = [C.c_void_p, C.c_int] + [_dp] * 4
        L.orc_cost.restype = C.c_double
        L.orc_cost.argtypes = [C.c_void_p] + [_dp] * 5
        L.orc_closed_loop.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, _dp, _ip, _ip, _dp]
        _lib = L
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _c(a, dtype=np.float64):
    return np.ascontiguousarray(a, dtype=dtype)


class Model:
    """bspline_shape + slider constants (oracle side)."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle model construction failed")
        self.h = handle
        n, nk = C.c_int(), C.c_int()
        b, ce, mu = C.c_double(), C.c_double(), C.c_double()
        lib().orc_model_info(self.h, C.byref(n), C.byref(nk), C.byref(b), C.byref(ce), C.byref(mu))
        self.n, self.nknots, self.b, self.c_ellipse, self.mu_sp = n.value, nk.value, b.value, ce.value, mu.value
        self.S = np.zeros(self.nknots)
        self.P = np.zeros((self.n, 2))
        self.c1 = np.zeros((self.n, 2))
        self.c2 = np.zeros((self.n, 2))
        lib().orc_model_tables(self.h, _d(self.S), _d(self.P), _d(self.c1), _d(self.c2))
        self.p = self.nknots - self.n - 1

    @classmethod
    def create(cls, S, P, p, mu_sp, c_ellipse, single_quirk=True):
        S = _c(S)
        P = _c(P)
        return cls(lib().orc_model_create(_d(S), len(S), _d(P), P.shape[0], p, mu_sp, c_ellipse, int(single_quirk)))

    @classmethod
    def from_ply(cls, path, flip, p, mu_sg, mu_sp, mass, tau_max):
        return cls(lib().orc_model_from_ply(str(path).encode(), int(flip), p, mu_sg, mu_sp, mass, tau_max))

    def __del__(self):
        try:
            lib().orc_model_free(self.h)
        except Exception:
            pass

    def basis(self, s, i, order):
        return lib().orc_basis(self.h, float(s), int(i), int(order))

    def eval_spline(self, s, wrap=0, local=True, want=("C", "Cd", "Cdd", "t", "n", "kappa")):
        s = _c(np.atleast_1d(s))
        k = len(s)
        out = {
            "C": np.zeros((k, 2)), "Cd": np.zeros((k, 2)), "Cdd": np.zeros((k, 2)),
            "t": np.zeros((k, 2)), "n": np.zeros((k, 2)), "kappa": np.zeros(k),
        }
        args = [_d(out[name]) if name in want else None for name in ("C", "Cd", "Cdd", "t", "n", "kappa")]
        lib().orc_eval_spline(self.h, k, _d(s), int(wrap), int(local), *args)
        return {name: out[name] for name in want}

    def get_curvatures(self, s):
        s = _c(np.atleast_1d(s))
        out = np.zeros(len(s))
        lib().orc_get_curvatures(self.h, len(s), _d(s), _d(out))
        return out

    def dynamics(self, x, u, jac=False, local=True):
        x = _c(np.atleast_2d(x))
        u = _c(np.atleast_2d(u))
        k = x.shape[0]
        f = np.zeros((k, 4))
        if not jac:
            lib().orc_dynamics(self.h, k, _d(x), _d(u), int(local), _d(f), None, None)
            return f
        Jx = np.zeros((k, 4, 4))
        Ju = np.zeros((k, 4, 2))
        lib().orc_dynamics(self.h, k, _d(x), _d(u), int(local), _d(f), _d(Jx), _d(Ju))
        return f, Jx, Ju

    def erk4_sens(self, x, u, dt, local=True, nthreads=1):
        x = _c(np.atleast_2d(x))
        u = _c(np.atleast_2d(u))
        k = x.shape[0]
        Phi = np.zeros((k, 4))
        A = np.zeros((k, 4, 4))
        B = np.zeros((k, 4, 2))
        lib().orc_erk4_sens(self.h, k, _d(x), _d(u), float(dt), int(local), int(nthreads), _d(Phi), _d(A), _d(B))
        return Phi, A, B

    def v_bound(self, s, v_alpha=1.0, d_v_bound=0.0, t_angle0=3.0, u_t_ub=0.05):
        ta = C.c_double()
        v = lib().orc_v_bound(self.h, float(s), v_alpha, d_v_bound, t_angle0, u_t_ub, C.byref(ta))
        return v, ta.value


DEFAULT_OPTS = dict(
    max_sqp_iter=30, tol_stat=1e-6, tol_eq=1e-6, tol_ineq=1e-6, tol_comp=1e-6,
    qp_max_iter=50, qp_tol=1e-11, qp_mu0=0.1, qp_thr=1e-3, qp_tau=0.9995,
    alpha_min=0.05, alpha_reduction=0.7, eps_sufficient_descent=1e-4, globalization=1, local_spline=1,
    qp_tol_comp=1e-18, qp_t_min=1e-12, qp_gamma_f=0.01, qp_stall=10,
    # recalled acados semantics as switches (qs_oracle.hpp); defaults = what the restatement believes
    sem_cost_scale=0, sem_h0_s_row=0, sem_full_step_dual=0, sem_merit_weights=0, sem_armijo=0, sem_erk_steps=1,
    sem_qp_maxiter_fails=0, sem_mod_strict=0, sem_qp_pivot_fails=0, qp_split_step=1,
)
SEMANTIC_SWITCHES = {          # name -> the alternatives to try when acados golden vectors disagree
    "sem_cost_scale": (1, 2), "sem_h0_s_row": (1,), "sem_full_step_dual": (1,), "sem_merit_weights": (1, 2), "sem_armijo": (1,),
    "sem_erk_steps": (2, 4), "sem_qp_maxiter_fails": (1,), "sem_mod_strict": (1,), "sem_qp_pivot_fails": (1,),
}
_OPT_ORDER = list(DEFAULT_OPTS.keys())


class Ocp:
    """OCP of NMPC_controller.create_ocp_model/create_ocp_opts (oracle side), batched over problems."""

    def __init__(self, model: Model, N: int, dt: float, **opts):
        self.model = model
        self.N, self.dt = int(N), float(dt)
        self.h = lib().orc_ocp_create(model.h, self.N, self.dt)
        self.opts = dict(DEFAULT_OPTS)
        self.set_opts(**opts)

    def __del__(self):
        try:
            lib().orc_ocp_free(self.h)
        except Exception:
            pass

    def set_opts(self, **opts):
        self.opts.update(opts)
        v = _c([float(self.opts[k]) for k in _OPT_ORDER])
        lib().orc_ocp_set_opts(self.h, _d(v))

    def set_W(self, stage, W):
        W = np.asfortranarray(W, dtype=np.float64)
        lib().orc_ocp_set_W(self.h, int(stage), W.ctypes.data_as(_dp))

    def set_bounds(self, lh, uh):
        lh, uh = _c(lh), _c(uh)
        lib().orc_ocp_set_bounds(self.h, _d(lh), _d(uh))

    def set_h_variant(self, variant=1, v_alpha=1.0, d_v_bound=0.0, t_angle0=3.0, u_t_ub=0.05,
                      u_n_lb=0.0, u_n_ub=0.03, u_t_lb=-0.05):
        """variant 1: h = [u_n; u_t - v_bound(s); u_t + v_bound(s)] with lh = [u_n_lb, 2 u_t_lb, 0],
        uh = [u_n_ub, 0, 2 u_t_ub] (the parked constraint set of NMPC_controller.m:226-248)."""
        lib().orc_ocp_set_h_variant(self.h, int(variant), v_alpha, d_v_bound, t_angle0, u_t_ub)
        if variant:
            self.set_bounds([u_n_lb, 2 * u_t_lb, 0.0], [u_n_ub, 0.0, 2 * u_t_ub])
        self.h_variant = int(variant)

    def v_bound_sym(self, s):
        dv = C.c_double(0.0)
        v = lib().orc_v_bound_sym(self.h, float(s), C.byref(dv))
        return v, dv.value

    def constraints(self, x, u):
        x, u = _c(x), _c(u)
        nb, N = x.shape[0], self.N
        h = np.zeros((nb, N, 3)); beta = np.zeros((nb, N, 3))
        lib().orc_constraints_batch(self.h, nb, _d(x), _d(u), _d(h), _d(beta))
        return h, beta

    def set_ctrl(self, v_alpha=1.0, d_v_bound=0.0, t_angle0=3.0, u_t_ub=0.05, u_n_lb=0.0):
        lib().orc_ocp_set_ctrl(self.h, v_alpha, d_v_bound, t_angle0, u_t_ub, u_n_lb)

    # array shapes: x0bar (nb,4), yref (nb,N,6), yref_e (nb,4), x (nb,N+1,4), u (nb,N,2), pi (nb,N,4), lam (nb,N,6)
    def linearise(self, x0bar, yref, yref_e, x, u):
        x0bar, yref, yref_e, x, u = map(_c, (x0bar, yref, yref_e, x, u))
        nb, N = x.shape[0], self.N
        A = np.zeros((nb, N, 4, 4)); B = np.zeros((nb, N, 4, 2)); b = np.zeros((nb, N, 4))
        g = np.zeros((nb, N, 6)); qN = np.zeros((nb, 4))
        lib().orc_linearise_batch(self.h, nb, _d(x0bar), _d(yref), _d(yref_e), _d(x), _d(u), _d(A), _d(B), _d(b), _d(g), _d(qN))
        return dict(A=A, B=B, b=b, g=g, qN=qN)

    def qp(self, x0bar, yref, yref_e, x, u, nthreads=1):
        x0bar, yref, yref_e, x, u = map(_c, (x0bar, yref, yref_e, x, u))
        nb, N = x.shape[0], self.N
        du = np.zeros((nb, N, 2)); dx = np.zeros((nb, N + 1, 4)); pi = np.zeros((nb, N, 4))
        lam = np.zeros((nb, N, 6)); t = np.zeros((nb, N, 6))
        iters = np.zeros(nb, dtype=np.int32); status = np.zeros(nb, dtype=np.int32); res = np.zeros((nb, 4))
        lib().orc_qp_batch(self.h, nb, _d(x0bar), _d(yref), _d(yref_e), _d(x), _d(u), int(nthreads),
                           _d(du), _d(dx), _d(pi), _d(lam), _d(t), _i(iters), _i(status), _d(res))
        return dict(du=du, dx=dx, pi=pi, lam=lam, t=t, iters=iters, status=status, res=res)

    def qp_data(self, x0bar, yref, yref_e, x, u):
        """The QP of the linearisation at (x, u) as dense per-stage data (input of the extended-precision arbiter)."""
        x0bar, yref, yref_e, x, u = map(_c, (x0bar, yref, yref_e, x, u))
        nb, N = x.shape[0], self.N
        d = dict(H=np.zeros((nb, N, 6, 6)), g=np.zeros((nb, N, 6)), A=np.zeros((nb, N, 4, 4)), B=np.zeros((nb, N, 4, 2)),
                 b=np.zeros((nb, N, 4)), QN=np.zeros((4, 4)), qN=np.zeros((nb, 4)), dx0=np.zeros((nb, 4)),
                 dl=np.zeros((nb, N, 3)), du=np.zeros((nb, N, 3)), beta=np.zeros((nb, N, 3)),
                 on=np.zeros((nb, N, 3), dtype=np.int32), ci=np.zeros((nb, N, 3), dtype=np.int32))
        lib().orc_qp_data_batch(self.h, nb, _d(x0bar), _d(yref), _d(yref_e), _d(x), _d(u),
                                *[_d(d[k]) for k in ("H", "g", "A", "B", "b", "QN", "qN", "dx0", "dl", "du", "beta")],
                                _i(d["on"]), _i(d["ci"]))
        return d

    def solve(self, mode, x0bar, yref, yref_e, x, u, pi=None, lam=None, nthreads=1):
        """mode 'rti' or 'sqp'.  Returns updated copies and stats."""
        x0bar, yref, yref_e = map(_c, (x0bar, yref, yref_e))
        x = _c(x).copy(); u = _c(u).copy()
        nb, N = x.shape[0], self.N
        pi = np.zeros((nb, N, 4)) if pi is None else _c(pi).copy()
        lam = np.zeros((nb, N, 6)) if lam is None else _c(lam).copy()
        si = np.zeros((nb, 3), dtype=np.int32); sd = np.zeros((nb, 6))
        lib().orc_solve_batch(self.h, 0 if mode == "rti" else 1, nb, _d(x0bar), _d(yref), _d(yref_e),
                              _d(x), _d(u), _d(pi), _d(lam), int(nthreads), _i(si), _d(sd))
        return dict(x=x, u=u, pi=pi, lam=lam, status=si[:, 0], sqp_iter=si[:, 1], qp_iter=si[:, 2],
                    cost=sd[:, 0], res=sd[:, 1:5], alpha=sd[:, 5])

    def prepare(self, x0, cold, x, u, pi=None, lam=None, nthreads=1):
        x0 = _c(x0).copy(); x = _c(x).copy(); u = _c(u).copy()
        nb, N = x.shape[0], self.N
        pi = np.zeros((nb, N, 4)) if pi is None else _c(pi).copy()
        lam = np.zeros((nb, N, 6)) if lam is None else _c(lam).copy()
        cold = _c(cold, np.int32)
        lib().orc_prepare_batch(self.h, nb, _d(x0), _i(cold), _d(x), _d(u), _d(pi), _d(lam), int(nthreads))
        return dict(x0=x0, x=x, u=u, pi=pi, lam=lam)

    def shift(self, x, u, pi, lam):
        x, u, pi, lam = (_c(a).copy() for a in (x, u, pi, lam))
        lib().orc_shift_batch(self.h, x.shape[0], _d(x), _d(u), _d(pi), _d(lam))
        return dict(x=x, u=u, pi=pi, lam=lam)

    def cost(self, x0bar, yref, yref_e, x, u):
        x0bar, yref, yref_e, x, u = map(_c, (x0bar, yref, yref_e, x, u))
        return lib().orc_cost(self.h, _d(x0bar), _d(yref), _d(yref_e), _d(x), _d(u))

    def closed_loop(self, mode, x0, yref_full, steps):
        """yref_full: (T,6) rows = columns of the controller's y_ref."""
        x0 = _c(x0); yref_full = _c(yref_full)
        T = yref_full.shape[0]
        xs = np.zeros((steps + 1, 4)); us = np.zeros((steps, 2))
        status = np.zeros(steps, dtype=np.int32); iters = np.zeros(steps, dtype=np.int32); cost = np.zeros(steps)
        lib().orc_closed_loop(self.h, 0 if mode == "rti" else 1, _d(x0), _d(yref_full), T, steps,
                              _d(xs), _d(us), _i(status), _i(iters), _d(cost))
        return dict(x=xs, u=us, status=status, sqp_iter=iters, cost=cost)


# objects_database/object_selection.m:3-42 restated (parameters only; paths are file names in cad_models/)
OBJECTS = {
    "santal": dict(mu_sg=0.32, mu_sp=0.19, xwidth=0.068, ywidth=0.082, m=0.2875, tau_max=0.0251,
                   pcl_path="planar_surface_santal_36_uniformed.ply", flip=False),
    "balea": dict(mu_sg=0.35, mu_sp=0.20, xwidth=0.071, ywidth=0.071, m=0.1713, tau_max=0.0042,
                  pcl_path="Balea_cad_model_planar_surface_36.ply", flip=False),
    "montana": dict(mu_sg=0.20, mu_sp=0.10, xwidth=0.057, ywidth=0.101, m=0.2467, tau_max=0.0101,
                    pcl_path="Montana_cad_model_planar_section_34.ply", flip=True),
    "pulirapid": dict(mu_sg=0.22, mu_sp=0.1, xwidth=0.13, ywidth=0.23, m=0.500, tau_max=0.0251,
                      pcl_path="pulirapid_ricarica_test_curvatura2_ply.ply", flip=True),
}


def model_from_reference_ply(name: str, cad_dir: str, p: int = 3) -> Model:
    o = OBJECTS[name]
    return Model.from_ply(os.path.join(cad_dir, o["pcl_path"]), o["flip"], p, o["mu_sg"], o["mu_sp"], o["m"], o["tau_max"])
