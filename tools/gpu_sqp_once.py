"""One full-SQP solve of a config-5-like batch (for a per-kernel launch list under ncu)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import uclv_qs_pushing_matlab_b200 as q
from uclv_qs_pushing_matlab_b200.workloads import make_rti_workload, OBJECT_ORDER
from tests.workloads import gpu_model
B, N = int(sys.argv[1]), int(sys.argv[2])
gms = [gpu_model(n) for n in OBJECT_ORDER]
wl = make_rti_workload(B, N, seed=4, n_objects=4, mixed_modes=True)
s = q.Solver(gms, N, 0.05, B, mode=1)
for rep in range(2):
    s.set("x0", wl["x0"]); s.set("yref", wl["yref"]); s.set("yref_e", wl["yref_e"]); s.set("u", wl["u_init"])
    s.set_int("cold", np.zeros(B, dtype=np.int32)); s.set_int("object_id", wl["object_id"])
    s.sync(); t0 = time.perf_counter(); s.prepare(); s.solve(); s.sync(); t = time.perf_counter() - t0
print("B", B, "N", N, "ms", 1e3 * t, "sqp it mean", s.get_int("sqp_iter").mean(), "launches", s.launches)
