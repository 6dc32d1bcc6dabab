"""acados_nmpc/NMPC_controller.m — NMPC controller facade, batched.

Same constructor, properties and methods as the reference class.  One controller object drives
`batch` independent NMPC instances (batch = 1 reproduces the MATLAB object); the pre-processing of
`solve` (x0 wrap, cold start, tangential-velocity clipping, Euler rollout: NMPC_controller.m:332,
351-380) and the post-processing shift (:397-399) run on the GPU (kernels k_prepare, k_shift), the
solve itself is qspush_solve behind the acados_ocp interface.
"""
from __future__ import annotations

import math

import numpy as np

from .acados_shim import acados_ocp, acados_ocp_model, acados_ocp_opts


def _blkdiag(a, b):
    a, b = np.atleast_2d(a), np.atleast_2d(b)
    out = np.zeros((a.shape[0] + b.shape[0], a.shape[1] + b.shape[1]))
    out[:a.shape[0], :a.shape[1]] = a
    out[a.shape[0]:, a.shape[1]:] = b
    return out


class NMPC_controller:
    def __init__(self, name, plant, sample_time, Hp, batch=1, device=0, nlp_solver="sqp",
                 velocity_constraint_in_ocp=False, **solver_opts):
        # NMPC_controller.m:16-26
        self.W_x = 0.01 * np.diag([100, 100, 0.1, 0])
        self.W_x_e = 200 * np.diag([1000, 1000, 0.1, 0])
        self.W_u = np.diag([1e-3, 1e-3])
        self.u_n_ub, self.u_t_ub, self.u_n_lb, self.u_t_lb = 0.03, 0.05, 0.0, -0.05
        self.name = name
        self.batch, self.device = int(batch), int(device)
        self.nlp_solver = nlp_solver           # "sqp" (reference, :272) or "sqp_rti"
        # True = the constraint set the authors parked in comments (:226-238, :247-248):
        # h = [u_n; u_t - v_bound(s); u_t + v_bound(s)] instead of [s; u_n; u_t] (:237)
        self.velocity_constraint_in_ocp = bool(velocity_constraint_in_ocp)
        self.solver_opts = solver_opts
        self.plants = list(plant) if isinstance(plant, (list, tuple)) else [plant]
        self.plant = self.plants[0]
        # :76-78
        self.sym_model = dict(self.plant.sym_model)
        self.sym_model["name"] = self.plant.name
        self.sym_model.setdefault("nx", 4)
        self.sym_model.setdefault("nu", 2)
        self.initial_condition = np.zeros(4)
        # :83-84
        self.h_constr_ub = [10, self.u_n_ub, self.u_t_ub]
        self.h_constr_lb = [-10, self.u_n_lb, self.u_t_lb]
        # :87-89
        self.Hp = int(Hp)
        self.sample_time = float(sample_time)
        self.T = self.Hp * self.sample_time
        # :98-100
        self.set_v_alpha(0.002 * 500)
        self.d_v_bound = 0.0
        self.t_angle0 = 3 + 0 * 2.831
        self.ocp_solver = None
        self.cost_function_vect = []
        self.utraj = self.xtraj = self.ptraj = None
        self.y_ref = None
        self.delay_compensation = 0.0
        self.delay_buff_comp = 0
        self.u_buff_contr = np.zeros((2, 0))
        self._cold = True

    # :106-110
    def set_delay_comp(self, delay):
        self.delay_compensation = delay
        self.delay_buff_comp = int(math.ceil(self.delay_compensation / self.sample_time))
        # (nu, d) like the reference for one instance; one buffer per problem, (batch, nu, d), for a batch
        shape = (self.sym_model["nu"], self.delay_buff_comp)
        self.u_buff_contr = np.zeros(shape if getattr(self, "batch", 1) == 1 else (self.batch, *shape))

    # :112-120 (x: (4,) with a (nu, d) buffer, or (batch, 4) with a (batch, nu, d) buffer)
    def delay_buffer_sim(self, plant, x):
        xk_sim = np.array(x, dtype=np.float64)
        for k in range(1, self.delay_buff_comp + 1):
            x_dot_sim = plant.evalModelVariableShape(xk_sim, self.u_buff_contr[..., -k])
            xk_sim = xk_sim + self.sample_time * x_dot_sim
        return xk_sim

    # :122-142 (dead code path in the reference: the call in main.m:91 is commented out; the bounds it sets,
    # [u_n_lb, 2 u_t_lb, 0] / [u_n_ub, 0, 2 u_t_ub], are those of the parked constraint set
    # h = [u_n; u_t -+ v_bound(s)], i.e. of velocity_constraint_in_ocp=True)
    def update_constraints(self, u_n_ub, u_t_ub, u_n_lb, u_t_lb):
        self.u_n_lb, self.u_n_ub, self.u_t_lb, self.u_t_ub = u_n_lb, u_n_ub, u_t_lb, u_t_ub
        self.h_constr_ub = [self.h_constr_ub[0], self.u_n_ub, self.u_t_ub]
        self.h_constr_lb = [self.h_constr_lb[0], self.u_n_lb, self.u_t_lb]
        self.ocp_solver.set("constr_lh", [*self.h_constr_lb[1:-1], 2 * self.u_t_lb, -0])
        self.ocp_solver.set("constr_uh", [*self.h_constr_ub[1:-1], 0, 2 * self.u_t_ub])

    # :144-151
    def clear_variables(self):
        self.utraj = self.xtraj = self.ptraj = None
        self.y_ref = None
        self.cost_function_vect = []
        self._cold = True

    # :153-164
    def update_cost_function(self, W_x, W_u, W_x_e, initial_step, final_step):
        self.ocp_solver.set("cost_W", W_x_e, self.Hp)
        for i in range(initial_step, final_step + 1):
            self.ocp_solver.set("cost_W", _blkdiag(W_x, W_u), i)
        self.W_x, self.W_u, self.W_x_e = np.asarray(W_x), np.asarray(W_u), np.asarray(W_x_e)

    # :166-172
    def initial_condition_update(self, new_initial_condition):
        self.initial_condition = np.asarray(new_initial_condition, dtype=np.float64)
        self.ocp_solver.set("constr_x0", self.initial_condition)
        self.clear_variables()

    # :174-268
    def create_ocp_model(self):
        nx, nu = self.sym_model["nx"], self.sym_model["nu"]
        ocp_model = acados_ocp_model()
        ocp_model.set("name", self.sym_model["name"])
        ocp_model.set("sym_x", self.sym_model.get("sym_x"))
        ocp_model.set("sym_u", self.sym_model.get("sym_u"))
        ocp_model.set("cost_type", "linear_ls")
        ocp_model.set("cost_type_e", "linear_ls")
        Vx = np.zeros((nx + nu, nx)); Vx[:nx, :nx] = np.eye(nx)
        Vu = np.zeros((nx + nu, nu)); Vu[nx:, :] = np.eye(nu)
        ocp_model.set("cost_Vx", Vx)
        ocp_model.set("cost_Vu", Vu)
        ocp_model.set("cost_Vz", np.zeros((nx + nu, 0)))
        ocp_model.set("cost_W", _blkdiag(self.W_x, self.W_u))
        ocp_model.set("cost_y_ref", np.zeros(nx + nu))
        ocp_model.set("cost_Vx_e", np.eye(nx))
        ocp_model.set("cost_W_e", self.W_x_e)
        ocp_model.set("cost_y_ref_e", np.zeros(nx))
        ocp_model.set("T", self.T)
        ocp_model.set("dyn_type", "explicit")
        ocp_model.set("dyn_expr_f", self.plants)          # the compiled dynamics travel with the plant(s)
        ocp_model.set("constr_type", "bgh")
        if self.velocity_constraint_in_ocp:
            ocp_model.set("constr_expr_h", ("u_n", "u_t-v_bound(s)", "u_t+v_bound(s)"))          # :238 (commented)
            ocp_model.set("constr_lh", [*self.h_constr_lb[1:-1], 2 * self.u_t_lb, -0.0])       # :247
            ocp_model.set("constr_uh", [*self.h_constr_ub[1:-1], 0.0, 2 * self.u_t_ub])        # :248
            ocp_model.set("constr_v_bound", dict(v_alpha=self.v_alpha, d_v_bound=self.d_v_bound,
                                                 t_angle0=self.t_angle0, u_t_ub=self.u_t_ub))  # :229
        else:
            ocp_model.set("constr_expr_h", ("s", "u_n", "u_t"))                    # :237
            ocp_model.set("constr_lh", [-0.06, *self.h_constr_lb[1:]])              # :251
            ocp_model.set("constr_uh", [0.011, *self.h_constr_ub[1:]])              # :252
        ocp_model.set("constr_x0", self.initial_condition)                     # :265
        return ocp_model

    # :270-300
    def create_ocp_opts(self):
        field_s = ["nlp_solver", "qp_solver", "sim_method", "globalization", "codgen_model", "compile_model", "compile_interface"]
        values_s = [self.nlp_solver, "partial_condensing_hpipm", "erk", "merit_backtracking", "true", "true", "true"]
        field_d = ["qp_solver_cond_N", "nlp_solver_max_iter", "line_search_use_sufficient_descent", "nlp_solver_tol_stat",
                   "nlp_solver_tol_eq", "nlp_solver_tol_ineq", "nlp_solver_tol_comp"]
        values_d = [5, 30, 1, 1e-6, 1e-6, 1e-6, 1e-6]
        ocp_opts = acados_ocp_opts()
        ocp_opts.set("param_scheme_N", self.Hp)
        for f, v in zip(field_s, values_s):
            ocp_opts.set(f, v)
        for f, v in zip(field_d, values_d):
            ocp_opts.set(f, v)
        ocp_opts.set("compile_interface", "auto")
        for k, v in self.solver_opts.items():
            ocp_opts.set(k, v)
        return ocp_opts

    # :302-305
    def create_ocp_solver(self):
        self.ocp_solver = acados_ocp(self.create_ocp_model(), self.create_ocp_opts(), batch=self.batch, device=self.device)
        s = self.ocp_solver.solver
        self._ref_key = None                                                   # the new solver holds no reference trajectory yet
        s.set_ctrl(v_alpha=self.v_alpha, d_v_bound=self.d_v_bound, t_angle0=self.t_angle0, u_t_ub=self.u_t_ub, u_n_lb=self.u_n_lb)

    # :307-313  (1-based index like the reference)
    def get_y_ref(self, index_ref):
        T = self.y_ref.shape[-1]
        col = T - 1 if index_ref > T else index_ref - 1
        return self.y_ref[..., col]

    # :315-317
    def set_v_alpha(self, alpha):
        self.v_alpha = alpha
        if getattr(self, "ocp_solver", None) is not None:
            self.ocp_solver.solver.set_ctrl(v_alpha=alpha)

    # :319-327
    def update_tangential_velocity_bounds(self, s):
        s = np.atleast_1d(np.asarray(s, dtype=np.float64))
        ctrl = self.ocp_solver.solver.ctrl if self.ocp_solver is not None else None
        v, t = self.plant._model.eval_v_bound(s, ctrl=ctrl, device=self.device)
        return (float(v[0]), float(t[0])) if v.size == 1 else (v, t)

    def set_object_ids(self, ids):
        """Batch extension: which of the plants each problem uses (config 4: Monte-Carlo over shapes)."""
        self.ocp_solver.solver.set_int("object_id", np.asarray(ids, dtype=np.int32))

    # :329-423
    def solve(self, x0, index_time):
        s = self.ocp_solver.solver
        B, N = self.batch, self.Hp
        x0 = np.asarray(x0, dtype=np.float64)
        self.ocp_solver.set("constr_x0", x0)                                   # :334 (the wrap of :332 happens in k_prepare)
        # reference window (:343-348)
        if self.y_ref.ndim == 2:
            # one trajectory shared by the batch: it lives on the device like obj.y_ref lives in the controller, and the window
            # of this period is cut there (qspush_set_reference_trajectory / _window); re-sent only when the array changed
            key = hash(self.y_ref.tobytes())
            if getattr(self, "_ref_key", None) != key:
                s.set_reference_trajectory(np.ascontiguousarray(self.y_ref.T))
                self._ref_key = key
            s.set_reference_window(int(index_time))
        else:
            cols = [min(index_time + k, self.y_ref.shape[-1]) - 1 for k in range(N)]
            yref = np.ascontiguousarray(np.transpose(self.y_ref[..., cols], (0, 2, 1)))   # (B,6,N) -> (B,N,6)
            s.set("yref", yref)
            s.set("yref_e", np.ascontiguousarray(yref[:, N - 1, :4]))
        if self._cold:                                                         # :351-355
            s.set_int("cold", np.ones(B, dtype=np.int32))
            self._cold = False
        s.prepare()                                                            # :357-384
        self.ocp_solver.solve()                                                # :389
        u = self.ocp_solver.get("u", 0)                                        # :403
        cost = self.ocp_solver.get_cost()                                      # :420
        self.cost_function_vect.append(cost)
        self._status = self.ocp_solver.get("status")
        s.shift()                                                              # :397-399
        self.utraj = self.ocp_solver.get("u")                                  # shifted trajectories, :392-399
        self.xtraj = self.ocp_solver.get("x")
        self.ptraj = self.ocp_solver.get("pi")
        return u

    # :425-431
    def set_reference_trajectory(self, y_ref):
        y_ref = np.asarray(y_ref, dtype=np.float64)
        pad_shape = y_ref.shape[:-1] + (self.delay_buff_comp,)
        pad = np.zeros(pad_shape)
        self.y_ref = np.concatenate([pad, y_ref], axis=-1)
        if self.delay_buff_comp > 0:
            self.y_ref[..., -1, :self.delay_buff_comp] = self.y_ref[..., -1, self.delay_buff_comp][..., None]
