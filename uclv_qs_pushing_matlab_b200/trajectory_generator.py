"""acados_nmpc/TrajectoryGenerator.m — reference generation."""
from __future__ import annotations

import numpy as np


class TrajectoryGenerator:
    def __init__(self, sample_time, u_n_des=0.01):
        self.sample_time = float(sample_time)
        self.u_n_des = u_n_des
        self.set_plot = False
        self.x0 = self.xf = None
        self.t0 = self.tf = 0.0
        self.waypoints_ = None
        self.waypoints_velocities = None

    def set_target(self, x0, xf, t0, tf):
        self.x0, self.xf = np.asarray(x0, dtype=np.float64).reshape(-1), np.asarray(xf, dtype=np.float64).reshape(-1)
        self.t0, self.tf = float(t0), float(tf)

    # TrajectoryGenerator.m:38-42
    def quintic_(self, time):
        tau = time / self.tf
        return 6 * tau ** 5 - 15 * tau ** 4 + 10 * tau ** 3

    # TrajectoryGenerator.m:44-79
    def straight_line(self, auto_angle):
        time = np.arange(self.t0, self.tf + 1e-12, self.sample_time)
        L = np.linalg.norm(self.xf - self.x0)
        traj = np.stack([self.x0 + self.quintic_(t) * L * (self.xf - self.x0) / L for t in time], axis=1)
        if auto_angle:
            tf_angle = self.tf / 2
            time_angle = np.arange(self.t0, tf_angle + 1e-12, self.sample_time)
            traj_angle = np.ones(len(time))
            d = self.xf[2] - self.x0[2]
            for j, t in enumerate(time_angle):
                s = self.quintic_(t) * abs(d)
                traj_angle[j] = self.x0[2] + (s * d / abs(d) if d != 0 else 0.0)
            # the reference indexes traj_angle(3,end) on a row vector (an out-of-range bug, :67); hold the last value
            traj_angle[len(time_angle) - 1:] = traj_angle[len(time_angle) - 1]
            traj = np.vstack([traj[0:2], traj_angle[None], traj[3:5]])
        return time, traj

    # TrajectoryGenerator.m:81-95
    def waypoint_gen_fixed_angle(self):
        wp = np.asarray(self.waypoints_, dtype=np.float64)
        vel = np.atleast_1d(np.asarray(self.waypoints_velocities, dtype=np.float64))
        delta_p = wp[1:] - wp[:-1]
        times = np.linalg.norm(delta_p, axis=1) / vel
        time = [0.0]
        cols = [np.concatenate([wp[0, :2], self.x0[2:4], [0.0]])]
        for i in range(len(wp) - 1):
            time_i = np.arange(time[-1] + self.sample_time, time[-1] + times[i] + 1e-12, self.sample_time)
            x_i = np.linspace(wp[i, 0], wp[i + 1, 0], len(time_i))
            y_i = np.linspace(wp[i, 1], wp[i + 1, 1], len(time_i))
            for a, b in zip(x_i, y_i):
                cols.append(np.concatenate([[a, b], self.x0[2:4], [0.0]]))
            time.extend(time_i.tolist())
        return np.asarray(time), np.stack(cols, axis=1)

    # TrajectoryGenerator.m:96-143 relies on the Navigation Toolbox `waypointTrajectory` (proprietary).
    # Documented stand-in: constant-speed straight segments between the
    # waypoints, heading = segment direction, rows [x; y; theta; v].
    def waypoints_gen(self):
        wp = np.asarray(self.waypoints_, dtype=np.float64)
        vel = np.broadcast_to(np.atleast_1d(np.asarray(self.waypoints_velocities, dtype=np.float64)), (len(wp) - 1,))
        t_knots = np.concatenate([[0.0], np.cumsum(np.linalg.norm(wp[1:, :2] - wp[:-1, :2], axis=1) / vel)])
        time = np.arange(0.0, t_knots[-1] + 1e-12, self.sample_time)
        x = np.interp(time, t_knots, wp[:, 0])
        y = np.interp(time, t_knots, wp[:, 1])
        seg = np.clip(np.searchsorted(t_knots, time, side="right") - 1, 0, len(wp) - 2)
        d = wp[1:, :2] - wp[:-1, :2]
        theta = np.arctan2(d[seg, 1], d[seg, 0])
        v = vel[seg]
        return time, np.vstack([x, y, theta, v])
