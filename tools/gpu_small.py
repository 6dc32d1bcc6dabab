"""Small end-to-end exercise of every kernel (for compute-sanitizer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import uclv_qs_pushing_matlab_b200 as q
from tests.workloads import gpu_model, make_rti_workload
gms = [gpu_model(n) for n in ("santal", "balea", "montana", "pulirapid")]
x = np.random.default_rng(0).uniform(-0.05, 0.05, (300, 4)); u = np.abs(np.random.default_rng(1).uniform(0.001, 0.03, (300, 2)))
gms[0].eval_spline(x[:, 3], wrap=2); gms[0].eval_dynamics(x, u, jac=True); gms[0].eval_erk4_sens(x, u, 0.05); gms[0].eval_v_bound(x[:, 3])
for N, B, kern, mode in ((40, 37, 1, 0), (10, 5, 1, 0), (100, 9, 1, 0), (40, 33, 0, 0), (10, 6, 1, 1), (130, 3, 1, 0)):
    wl = make_rti_workload(None, batch=B, N=N, seed=1, n_objects=4)
    s = q.Solver(gms, N, 0.05, B, qp_kernel=kern, mode=mode, max_sqp_iter=3)
    s.set("x0", wl["x0"]); s.set("yref", wl["yref"]); s.set("yref_e", wl["yref_e"]); s.set("u", wl["u_init"])
    s.set_int("cold", np.zeros(B, dtype=np.int32)); s.set_int("object_id", wl["object_id"])
    s.prepare(); s.solve(); s.shift()
    xs = wl["x0"].copy(); s.plant_step(xs, s.get("u", stage=0))
    print(N, B, kern, mode, s.get_int("status"), float(np.abs(s.get("u")).max()))
print("done")
