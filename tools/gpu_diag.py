import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import uclv_qs_pushing_matlab_b200 as q
from tests.workloads import gpu_model, make_rti_workload
gm = gpu_model("santal"); B, N = 4096, 40
wl = make_rti_workload(None, batch=B, N=N, seed=2)
sol = {}
for kern in (1, 0):
    s = q.Solver([gm], N, 0.05, B, qp_kernel=kern)
    s.set("x0", wl["x0"]); s.set("yref", wl["yref"]); s.set("yref_e", wl["yref_e"]); s.set("u", wl["u_init"]); s.set_int("cold", np.zeros(B, dtype=np.int32))
    s.prepare(); s.solve()
    sol[kern] = dict(u=s.get("u"), x=s.get("x"), it=s.get_int("qp_iter"), st=s.get_int("status"), res=s.get("res"))
d = np.abs(sol[1]["u"] - sol[0]["u"]).max(axis=(1, 2)); dx = np.abs(sol[1]["x"] - sol[0]["x"]).max(axis=(1, 2))
print("iter diff hist", np.bincount(np.abs(sol[1]["it"] - sol[0]["it"])))
print("u diff: frac<1e-8 %.4f <1e-6 %.4f <1e-5 %.4f <1e-4 %.4f max %.2e" % ((d < 1e-8).mean(), (d < 1e-6).mean(), (d < 1e-5).mean(), (d < 1e-4).mean(), d.max()))
print("x diff: frac<1e-6 %.4f max %.2e" % ((dx < 1e-6).mean(), dx.max()))
same = sol[1]["it"] == sol[0]["it"]
print("same-iter subset: u max %.2e ; diff-iter subset: u median %.2e max %.2e" % (d[same].max(), np.median(d[~same]), d[~same].max()))
print("res max", sol[1]["res"].max(0), sol[0]["res"].max(0))
w = np.argsort(d)[-5:]; print("worst", w, d[w], sol[1]["it"][w], sol[0]["it"][w])
