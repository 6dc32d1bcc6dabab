import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import uclv_qs_pushing_matlab_b200 as q
from tests.workloads import packaged_model_pair
gm, om = packaged_model_pair("santal")
N, dt, steps, T, B = 10, 0.05, 30, 201, 48
t = np.arange(T) * dt
traj = np.zeros((T, 6)); traj[:, 0] = np.minimum(0.01 * t, 0.10)
rng = np.random.default_rng(5)
x0s = np.stack([rng.uniform(-0.003, 0.003, B), rng.uniform(-0.003, 0.003, B), rng.uniform(-0.05, 0.05, B), rng.uniform(-0.02, 0.004, B)], 1)
off = np.zeros((B, 6)); off[:, :2] = x0s[:, :2]
def host_loop():
    s = q.Solver([gm], N, dt, B)
    x = x0s.copy(); u_host = np.zeros((steps, B, 2)); x_host = np.zeros((steps, B, 4))
    s.set_int("cold", np.ones(B, dtype=np.int32))
    for i in range(1, steps + 1):
        cols = [min(i + k, T) - 1 for k in range(N)]
        yref = np.ascontiguousarray(traj[cols][None] + off[:, None, :])
        x_host[i - 1] = x
        s.set("x0", x); s.set("yref", yref); s.set("yref_e", np.ascontiguousarray(yref[:, N - 1, :4]))
        s.prepare(); s.solve()
        u = s.get("u", stage=0); u_host[i - 1] = u
        s.shift()
        x = s.plant_step(np.ascontiguousarray(x.copy()), np.ascontiguousarray(u))
    return u_host, x_host, x
u1, x1, xf1 = host_loop(); u2, x2, xf2 = host_loop()
print("host loop repeatable:", np.array_equal(u1, u2), np.abs(u1 - u2).max())
s2 = q.Solver([gm], N, dt, B)
r = s2.closed_loop(traj, x0s.copy(), steps, offset=off)
d = np.abs(r["u_log"] - u1).max(axis=(1, 2)); dx = np.abs(r["x_log"] - x1).max(axis=(1, 2))
print("u diff per step", d[:8], "max", d.max()); print("x diff per step", dx[:8], "max", dx.max())
s3 = q.Solver([gm], N, dt, B); r3 = s3.closed_loop(traj, x0s.copy(), steps, offset=off)
print("device loop repeatable:", np.array_equal(r["u_log"], r3["u_log"]))
