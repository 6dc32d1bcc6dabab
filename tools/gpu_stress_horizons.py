"""Warp QP kernel (every mapping: 16- and 32-lane segments, C = 1..4, full and partial last lanes) against the
one-problem-per-thread kernel over many horizons (development aid / regression sweep)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import uclv_qs_pushing_matlab_b200 as q
from tests.workloads import gpu_model, make_rti_workload
gms = [gpu_model(n) for n in ("santal", "balea", "montana", "pulirapid")]
worst = 0.0
for N in [1, 2, 3, 7, 8, 14, 15, 16, 17, 30, 31, 32, 33, 46, 47, 48, 49, 62, 63, 64, 65, 80, 95, 96, 97, 126, 127]:
    for B in (7, 300):
        wl = make_rti_workload(None, batch=B, N=N, seed=N, n_objects=4)
        us = []
        for kern in (1, 0):
            s = q.Solver(gms, N, 0.05, B, qp_kernel=kern)
            s.set("x0", wl["x0"]); s.set("yref", wl["yref"]); s.set("yref_e", wl["yref_e"]); s.set("u", wl["u_init"])
            s.set_int("cold", np.zeros(B, dtype=np.int32)); s.set_int("object_id", wl["object_id"])
            for rep in range(2):                         # second pass exercises the ordered work queue
                s.set("u", wl["u_init"]); s.prepare(); s.solve()
            us.append((s.get("u"), s.get_int("status"), s.get_int("qp_iter"), s.get("res").max(1)))
        d = np.abs(us[0][0] - us[1][0]).reshape(B, -1).max(axis=1)
        ok = (us[0][1] == us[1][1]).mean()
        worst = max(worst, np.median(d))
        print("N %3d B %3d: status equal %.3f  ok(warp) %.3f  |du| median %.1e max %.1e  kipm %.1f/%.1f  res max %.1e" % (
            N, B, ok, (us[0][1] == 0).mean(), np.median(d), d.max(), us[0][2].mean(), us[1][2].mean(), us[0][3].max()), flush=True)
        assert ok > 0.97 and np.median(d) < 1e-7 and (us[0][3] < 1e-6).mean() > 0.97, (N, B)
print("STRESS OK, worst median", worst)
