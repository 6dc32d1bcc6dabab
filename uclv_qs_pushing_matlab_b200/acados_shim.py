"""The subset of acados' MATLAB interface that the reference drives, over qspush.

    acados_ocp_model / acados_ocp_opts : the `set(field, value)` bags filled by
        NMPC_controller.create_ocp_model / create_ocp_opts (NMPC_controller.m:174-300)
    acados_ocp(model, opts)            : set / get / solve / get_cost / print
        (NMPC_controller.m:137-138,154-157,170,304,334-348,382-420; helper.m:253,264-269)

Shapes follow MATLAB for a single problem (column vectors, dim x stages matrices).  With batch > 1 a
MATLAB-shaped value is broadcast to every problem, and arrays with a leading batch axis in the C-ABI
layout [batch][stage][dim] address the problems individually.  Bad field names raise (the MEX layer
errors in MATLAB); solve() never raises for a non-converged problem — poll get('status').
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .capi import Solver


class _Bag:
    def __init__(self):
        self._d = {}

    def set(self, field, value):
        self._d[str(field)] = value

    def get(self, field, default=None):
        return self._d.get(str(field), default)


class acados_ocp_model(_Bag):
    pass


class acados_ocp_opts(_Bag):
    pass


_SET_FIELDS = {"constr_x0": ("x0", 4), "cost_y_ref": ("yref", 6), "cost_y_ref_e": ("yref_e", 4),
               "init_x": ("x", 4), "init_u": ("u", 2), "init_pi": ("pi", 4)}
_GET_TRAJ = {"u": ("u", 2), "x": ("x", 4), "pi": ("pi", 4), "lam": ("lam", 6)}


class acados_ocp:
    def __init__(self, model: acados_ocp_model, opts: acados_ocp_opts, batch: int = 1, device: int = 0):
        plant = model.get("dyn_expr_f")
        plants = plant if isinstance(plant, (list, tuple)) else [plant]
        if any(not hasattr(p, "_model") for p in plants):
            raise L.QspushError("model.dyn_expr_f must carry the PusherSliderModel(s) whose dynamics are compiled in")
        self.N = int(opts.get("param_scheme_N"))
        self.T = float(model.get("T"))
        self.dt = self.T / self.N
        self.batch = int(batch)
        nlp = str(opts.get("nlp_solver", "sqp"))
        if nlp not in ("sqp", "sqp_rti"):
            raise L.QspushError(f"nlp_solver {nlp!r} is not supported (sqp, sqp_rti)")
        if str(opts.get("qp_solver", "partial_condensing_hpipm")) != "partial_condensing_hpipm":
            raise L.QspushError("only the (partial-condensing-equivalent) Riccati IPM QP solver exists")
        if str(opts.get("sim_method", "erk")) != "erk":
            raise L.QspushError("only sim_method 'erk' (4 stages, 1 step) exists")
        kw = dict(
            mode=L.MODE_SQP if nlp == "sqp" else L.MODE_RTI,
            max_sqp_iter=int(opts.get("nlp_solver_max_iter", 100)),
            tol_stat=float(opts.get("nlp_solver_tol_stat", 1e-6)), tol_eq=float(opts.get("nlp_solver_tol_eq", 1e-6)),
            tol_ineq=float(opts.get("nlp_solver_tol_ineq", 1e-6)), tol_comp=float(opts.get("nlp_solver_tol_comp", 1e-6)),
            qp_max_iter=int(opts.get("qp_solver_iter_max", 50)),
            globalization=1 if str(opts.get("globalization", "fixed_step")) == "merit_backtracking" else 0,
        )
        expr_h = model.get("constr_expr_h")
        if expr_h is not None:
            expr_h = tuple(str(e).replace(" ", "") for e in expr_h)
            if expr_h == ("u_n", "u_t-v_bound(s)", "u_t+v_bound(s)"):
                kw["h_variant"] = 1                              # NMPC_controller.m:238 (the authors' parked constraint set)
            elif expr_h != ("s", "u_n", "u_t"):
                raise L.QspushError(f"constr_expr_h {expr_h!r} is not one of the two constraint sets of the reference")
        for k in ("qp_tol", "qp_mu0", "qp_thr", "qp_tau", "qp_tol_comp", "qp_t_min", "qp_gamma_f", "qp_stall", "problems_per_warp", "matlab_single_quirk", "qp_kernel"):
            if opts.get(k) is not None:
                kw[k] = opts.get(k)
        self.solver = Solver([p._model for p in plants], self.N, self.dt, self.batch, device=device, **kw)
        if model.get("constr_v_bound") is not None:
            self.solver.set_ctrl(**model.get("constr_v_bound"))
        W = model.get("cost_W")
        if W is not None:
            self.solver.set("W", np.asarray(W, dtype=np.float64), stage=-1)
        We = model.get("cost_W_e")
        if We is not None:
            self.solver.set("W", np.asarray(We, dtype=np.float64), stage=self.N)
        if model.get("constr_lh") is not None:
            self.solver.set("lh", np.asarray(model.get("constr_lh"), dtype=np.float64).reshape(-1))
        if model.get("constr_uh") is not None:
            self.solver.set("uh", np.asarray(model.get("constr_uh"), dtype=np.float64).reshape(-1))
        if model.get("constr_x0") is not None:
            self.set("constr_x0", model.get("constr_x0"))

    # ---- helpers
    def _bcast(self, v, dim, nst):
        """MATLAB-shaped (dim,), (dim,1), (dim,nst) or batch-shaped (B,nst,dim) / (B,dim) -> [B][nst][dim]."""
        a = np.asarray(v, dtype=np.float64)
        B = self.batch
        if a.ndim == 3:
            if a.shape != (B, nst, dim):
                raise L.QspushError(f"expected shape {(B, nst, dim)}, got {a.shape}")
            return np.ascontiguousarray(a)
        if nst == 1:
            if a.ndim == 2 and B > 1 and a.shape == (B, dim):
                return np.ascontiguousarray(a.reshape(B, 1, dim))
            if a.size == dim:
                return np.ascontiguousarray(np.broadcast_to(a.reshape(1, 1, dim), (B, 1, dim)))
        elif a.ndim == 2 and a.shape == (dim, nst):
            return np.ascontiguousarray(np.broadcast_to(a.T.reshape(1, nst, dim), (B, nst, dim)))
        raise L.QspushError(f"value of shape {a.shape} does not fit dim {dim} x stages {nst} (batch {B})")

    # ---- acados_ocp interface
    def set(self, field, value, stage=None):
        field = str(field)
        s = self.solver
        if field == "cost_W":
            k = -1 if stage is None else int(stage)
            s.set("W", np.asarray(value, dtype=np.float64), stage=k)
            return
        if field in ("constr_lh", "constr_uh"):
            v = np.asarray(value, dtype=np.float64).reshape(-1)
            if v.size != 3:
                raise L.QspushError(f"{field}: expected nh = 3 values, got {v.size}")
            s.set("lh" if field == "constr_lh" else "uh", v)
            return
        if field not in _SET_FIELDS:
            raise L.QspushError(f"acados_ocp.set: field {field!r} is not supported")
        name, dim = _SET_FIELDS[field]
        nst_all = {"x0": 1, "yref": self.N, "yref_e": 1, "x": self.N + 1, "u": self.N, "pi": self.N}[name]
        if stage is None or name in ("x0", "yref_e"):
            s.set(name, self._bcast(value, dim, nst_all), stage=-1)
        else:
            s.set(name, self._bcast(value, dim, 1).reshape(self.batch, dim), stage=int(stage))

    def solve(self):
        self.solver.solve()

    def get(self, field, stage=None):
        field = str(field)
        s = self.solver
        if field in _GET_TRAJ:
            name, dim = _GET_TRAJ[field]
            if stage is not None:
                v = s.get(name, stage=int(stage))                  # (B, dim)
                return v[0].copy() if self.batch == 1 else v
            v = s.get(name, stage=-1)                              # (B, nst, dim)
            return v[0].T.copy() if self.batch == 1 else v         # MATLAB: dim x stages
        if field in ("status", "sqp_iter", "qp_iter"):
            v = s.get_int(field)
            return int(v[0]) if self.batch == 1 else v
        if field in ("time_tot", "time_lin", "time_qp_sol"):
            return s.stat(field)
        if field == "residuals":
            v = s.get("res")
            return v[0] if self.batch == 1 else v
        raise L.QspushError(f"acados_ocp.get: field {field!r} is not supported")

    def get_cost(self):
        v = self.solver.get("cost")
        return float(v[0]) if self.batch == 1 else v

    def print(self, *_):
        st, it, qp = self.solver.get_int("status"), self.solver.get_int("sqp_iter"), self.solver.get_int("qp_iter")
        res = self.solver.get("res")
        print("problem\tstatus\tsqp_iter\tqp_iter\tres_stat\tres_eq\t\tres_ineq\tres_comp")
        for b in range(min(self.batch, 16)):
            print(f"{b}\t{st[b]}\t{it[b]}\t\t{qp[b]}\t{res[b,0]:e}\t{res[b,1]:e}\t{res[b,2]:e}\t{res[b,3]:e}")
