"""Monte-Carlo of whole pushes on the device: 4096 closed loops x 200 control periods, N = 40, RTI (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import uclv_qs_pushing_matlab_b200 as q
from tests.workloads import gpu_model
gm = gpu_model("santal")
B, N, dt, steps, T = 4096, 40, 0.05, 200, 201
t = np.arange(T) * dt
traj = np.zeros((T, 6)); traj[:, 0] = np.minimum(0.01 * t, 0.10)
rng = np.random.default_rng(2)
x0 = np.stack([rng.uniform(-0.03, 0.03, B), rng.uniform(-0.03, 0.03, B), np.deg2rad(rng.uniform(-10, 10, B)), rng.uniform(-0.04, 0.005, B)], 1)
off = np.zeros((B, 6)); off[:, :2] = x0[:, :2]
s = q.Solver([gm], N, dt, B)
dev = torch.device("cuda:0")
tr, od = torch.from_numpy(traj).to(dev), torch.from_numpy(off).to(dev)
for rep in range(2):
    xd = torch.from_numpy(x0.copy()).to(dev)
    s.set_int("cold", np.ones(B, dtype=np.int32)); s.sync()
    t0 = time.perf_counter(); r = s.closed_loop(tr, xd, steps, offset=od); s.sync(); t1 = time.perf_counter() - t0
st = r["status_log"].cpu().numpy(); xf = xd.cpu().numpy()
print("4096 pushes x 200 periods: %.1f ms total, %.3f ms per period, %.3e controller solves/s, status ok %.4f, final |x - x_ref| median %.2e" % (
    1e3 * t1, 1e3 * t1 / steps, B * steps / t1, (st == 0).mean(), np.median(np.abs(xf[:, 0] - (x0[:, 0] + traj[-1, 0])))))
