import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import uclv_qs_pushing_matlab_b200 as q
from tests.workloads import gpu_model
from uclv_qs_pushing_matlab_b200.workloads import make_rti_workload
gm = gpu_model("santal")
for B, N in ((1, 40), (1, 10), (8, 40), (148, 40)):
    wl = make_rti_workload(B, N, seed=2)
    s = q.Solver([gm], N, 0.05, B)
    for r in range(5):
        s.set("x0", wl["x0"]); s.set("yref", wl["yref"]); s.set("yref_e", wl["yref_e"]); s.set("u", wl["u_init"]); s.set_int("cold", np.zeros(B, dtype=np.int32))
        s.sync(); t0 = time.perf_counter(); s.prepare(); s.solve(); s.sync(); t1 = time.perf_counter()
    print(B, N, "wall %.3f ms prep %.3f lin %.3f qp %.3f tot %.3f iters %s" % (1e3*(t1-t0), 1e3*s.stat("time_prep"), 1e3*s.stat("time_lin"), 1e3*s.stat("time_qp_sol"), 1e3*s.stat("time_tot"), s.get_int("qp_iter")[:4]))
