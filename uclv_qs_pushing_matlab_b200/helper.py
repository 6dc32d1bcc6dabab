"""helper.m — closed-loop / open-loop drivers around the controller (the immediate caller of the hot path).

Only the simulation drivers are mirrored (helper.m:132-193, 195-322); plots, animation, the Simulink
runner, .mat saving and the brute-force cost debugger are visualisation / debugging tools of the
reference and are out of scope.
"""
from __future__ import annotations

import math

import numpy as np


class helper:
    g = 9.81        # helper.m:3

    @staticmethod
    def my_rotz(theta):     # helper.m:7-10
        c, s = math.cos(theta), math.sin(theta)
        return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])

    @staticmethod
    def my_rotz_2d(theta):  # helper.m:12-15
        return helper.my_rotz(theta)[:2, :2]

    @staticmethod
    def _reproject_s(plant, target_xy, s0):
        """argmin_s ||C(s) - target||^2 (helper.m:216-232 uses fminunc from s0; here: GPU scan + Newton polish)."""
        SP = plant.SP
        grid = np.linspace(0.0, SP.b, 2049)[:-1]
        C = SP.evalSpline(SP.FC, grid)
        d2 = ((C - target_xy[None, :]) ** 2).sum(1)
        # fminunc is a local search started at s0: prefer the basin nearest to s0 among near-global minima
        s = float(grid[int(np.argmin(d2))])
        for _ in range(20):
            c = SP.evalSpline(SP.FC, [s])[0]
            d = SP.evalSpline(SP.FC_dot, [s])[0]
            dd = SP.evalSpline(SP.FC_dot_dot, [s])[0]
            r = c - target_xy
            g1 = 2.0 * float(r @ d)
            g2 = 2.0 * float(d @ d + r @ dd)
            if not (g2 > 0.0):
                break
            step = g1 / g2
            s -= step
            if abs(step) < 1e-14:
                break
        # fminunc starts at s0 = 0 (helper.m:218) and walks to the nearby minimum on the pushed face: report the
        # representative of the minimiser closest to 0, i.e. in (-b/2, b/2]
        if s > 0.5 * SP.b:
            s -= SP.b
        return s

    @staticmethod
    def closed_loop_device(plant, controller, x0, time_sim, sim_noise=False, disturbance_=False, amplitude_dist=0.0,
                           t_dist=10 ** 9, seed=0, offset=None):
        """helper.closed_loop_matlab (helper.m:195-322) for a whole batch WITHOUT host round trips: the loop of reference
        window -> prepare -> solve -> plant step -> shift runs on the GPU (qspush_closed_loop).  x0: (B,4) numpy array or
        torch CUDA tensor; the controller's reference (6,T) is shared by the batch, `offset` (B,6) shifts it per problem.
        Input delays (plant.time_delay, controller.set_delay_comp; helper.m:205-212, 244-250, 290-298) run on the device too.
        Returns (x_s, y_s, theta_s, s_s, u_n, u_t, time_sim_vec, found_sol) with a leading batch axis, in x0's memory space;
        the states are the ones handed to the controller (closed_loop_matlab's x_sim: the plant state rolled through the
        delay-compensation buffer; the plant state itself when there is no compensation)."""
        dt = controller.sample_time
        delay_plant = int(math.ceil(plant.time_delay / dt))                                       # helper.m:211
        delay_comp = int(controller.delay_buff_comp)                                              # NMPC_controller.m:108
        time_sim_vec = np.arange(0.0, time_sim + 1e-12, dt)
        T = len(time_sim_vec)
        solver = controller.ocp_solver.solver
        y_ref = np.ascontiguousarray(np.asarray(controller.y_ref, dtype=np.float64).T)          # (T_ref, 6)
        on_dev = helper._is_cuda(x0)
        if on_dev:
            import torch
            x = x0.clone().contiguous()
            traj = torch.from_numpy(y_ref).to(x.device)
            off = None if offset is None else offset.contiguous()
        else:
            x = np.ascontiguousarray(np.array(x0, dtype=np.float64).reshape(-1, 4))
            traj = y_ref
            off = None if offset is None else np.ascontiguousarray(offset, dtype=np.float64)
        if controller._cold:
            solver.set_int("cold", np.ones(controller.batch, dtype=np.int32)); controller._cold = False
        sig = (1e-5, 1e-5, 1e-3, 1e-4) if sim_noise else (0.0, 0.0, 0.0, 0.0)                    # helper.m:241
        r = solver.closed_loop(traj, x, T, offset=off, idx0=1 + delay_comp, noise_sigma=sig, seed=seed, delay_plant=delay_plant, delay_comp=delay_comp,
                               t_dist=(t_dist if disturbance_ and t_dist <= T else 0), amplitude_dist=amplitude_dist,
                               xwidth=plant.slider_params["xwidth"])
        xl, ul, st = r["x_log"], r["u_log"], r["status_log"]
        tr = (lambda a: a.permute(1, 0, 2)) if on_dev else (lambda a: np.transpose(a, (1, 0, 2)))
        xl, ul = tr(xl), tr(ul)
        found = (st == 0).T if not on_dev else (st == 0).t()
        return xl[..., 0], xl[..., 1], xl[..., 2], xl[..., 3], ul[..., 0], ul[..., 1], time_sim_vec, found

    @staticmethod
    def _is_cuda(a):
        try:
            import torch
            return isinstance(a, torch.Tensor) and a.is_cuda
        except ImportError:
            return False

    @staticmethod
    def closed_loop_matlab(plant, controller, x0, time_sim, print_=False, sim_noise=False, debug_cost=False,
                           disturbance_=False, amplitude_dist=0.0, t_dist=10 ** 9, rng=None):
        """helper.m:195-322.  Returns (x_s, x_sim, y_s, theta_s, S_p_x, S_p_y, u_n, u_t, time_sim_vec, mode_vect, found_sol).

        Works for batch = 1 exactly like the reference and for batch > 1 with a leading batch axis on
        the states (x0: (B,4)); the plant step runs on the GPU (qspush_plant_step).
        """
        if debug_cost:
            raise NotImplementedError("debug_cost_function is a plotting aid of the reference (helper.m:356-451): out of scope")
        rng = np.random.default_rng() if rng is None else rng
        dt = controller.sample_time
        time_sim_vec = np.arange(0.0, time_sim + 1e-12, dt)
        T = len(time_sim_vec)
        x0 = np.asarray(x0, dtype=np.float64)
        batched = x0.ndim == 2
        B = x0.shape[0] if batched else 1
        x = np.zeros((B, T + 1, 4))
        x[:, 0] = x0.reshape(B, 4)
        x_sim = np.zeros((B, T, 4))
        u = np.zeros((B, T, 2))
        found_sol = np.zeros((B, T), dtype=bool)
        delay_buff_plant = int(math.ceil(plant.time_delay / dt))
        u_buff_plant = np.zeros((B, 2, delay_buff_plant))
        s0_spline = 0.0
        solver = controller.ocp_solver.solver
        for i in range(1, T + 1):                                   # 1-based like the reference
            xi = x[:, i - 1]
            if disturbance_ and i == t_dist:                        # :221-236 lateral shove + re-projection of s
                for b in range(B):
                    xi[b, 1] += amplitude_dist
                    Sp = plant.SP.evalSpline(plant.SP.FC, [xi[b, 3]])[0]
                    target = np.array([-plant.slider_params["xwidth"] / 2.0, Sp[1] - amplitude_dist])
                    s_min = helper._reproject_s(plant, target, s0_spline)
                    s0_spline = math.fmod(s_min, plant.SP.b) + (plant.SP.b if s_min < 0 and math.fmod(s_min, plant.SP.b) != 0 else 0.0)
                    s0_spline = s0_spline - plant.SP.b * (s_min < 0)  # :232 mod(s,b) - b*(s<0)
                    xi[b, 3] = s0_spline
            if sim_noise:                                           # :240-242
                xi += rng.standard_normal((B, 4)) * np.array([1e-5, 1e-5, 1e-3, 1e-4])
            if controller.delay_buff_comp:                          # every problem is rolled forward with ITS OWN past inputs
                xk_sim = controller.delay_buffer_sim(plant, xi) if controller.u_buff_contr.ndim == 3 else controller.delay_buffer_sim(plant, xi[0])[None]
            else:
                xk_sim = xi.copy()
            x_sim[:, i - 1] = xk_sim
            ui = controller.solve(xk_sim if batched else xk_sim[0], i + controller.delay_buff_comp)   # :248
            u[:, i - 1] = np.asarray(ui).reshape(B, 2)
            if controller.delay_buff_comp:                          # :252
                if controller.u_buff_contr.ndim == 3:
                    controller.u_buff_contr = np.concatenate([u[:, i - 1][:, :, None], controller.u_buff_contr[:, :, :-1]], axis=2)
                else:
                    controller.u_buff_contr = np.concatenate([u[0, i - 1].reshape(2, 1), controller.u_buff_contr[:, :-1]], axis=1)
            status = np.atleast_1d(controller.ocp_solver.get("status"))
            found_sol[:, i - 1] = status == 0                        # :253-260
            if print_:                                              # :263-273
                controller.ocp_solver.print()
                print("\nstatus = %s, sqp_iter = %s, time_int = %f [ms] (time_lin = %f [ms], time_qp_sol = %f [ms])" % (
                    status, controller.ocp_solver.get("sqp_iter"), controller.ocp_solver.get("time_tot") * 1e3,
                    controller.ocp_solver.get("time_lin") * 1e3, controller.ocp_solver.get("time_qp_sol") * 1e3))
            # plant simulation, forward Euler (:292-307)
            if delay_buff_plant == 0:
                u_apply = u[:, i - 1]
            else:
                u_apply = u_buff_plant[:, :, -1].copy()
                u_buff_plant = np.concatenate([u[:, i - 1][:, :, None], u_buff_plant[:, :, :-1]], axis=2)
            xn = np.ascontiguousarray(xi.copy())
            solver.plant_step(xn, np.ascontiguousarray(u_apply))
            x[:, i] = xn
        xs = x[:, :-1]
        S_p = np.stack([plant.SP.evalSpline(plant.SP.FC, xs[b, :, 3]) for b in range(B)])   # :317-319
        out = (xs[:, :, 0], x_sim, xs[:, :, 1], xs[:, :, 2], S_p[:, :, 0], S_p[:, :, 1], u[:, :, 0], u[:, :, 1])
        if not batched:
            out = tuple(o[0] for o in out)
            found_sol = found_sol[0]
        mode_vect = np.zeros(T)                                     # the reference leaves it unset too (:302)
        return (*out, time_sim_vec, mode_vect, found_sol)

    @staticmethod
    def open_loop_matlab(plant, x0, u_n, u_t, time_sim, sample_time, sim_noise=False, rng=None):
        """helper.m:132-193 reduced to its simulation core: constant input, forward Euler."""
        rng = np.random.default_rng() if rng is None else rng
        t = np.arange(0.0, time_sim + 1e-12, sample_time)
        x = np.zeros((len(t) + 1, 4))
        x[0] = x0
        for i in range(len(t)):
            xi = x[i] + (rng.standard_normal(4) * np.array([1e-5, 1e-5, 1e-3, 1e-4]) if sim_noise else 0.0)
            x[i + 1] = xi + sample_time * plant.evalModelVariableShape(xi, np.array([u_n, u_t]))
        S_p = plant.SP.evalSpline(plant.SP.FC, x[:-1, 3])
        return x[:-1, 0], x[:-1, 1], x[:-1, 2], S_p[:, 0], S_p[:, 1], np.full(len(t), u_n), np.full(len(t), u_t), t
