"""dev aid: qspush_step vs the field-by-field control period, where do they differ?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import uclv_qs_pushing_matlab_b200 as q
from uclv_qs_pushing_matlab_b200.workloads import make_rti_workload, packaged_model

gm = packaged_model("santal")
B, N, T = 300, 40, 64
wl = make_rti_workload(B, N, seed=12)
traj = np.zeros((T, 6)); traj[:, 0] = 0.01 * 0.05 * np.arange(T)
off = np.zeros((B, 6)); off[:, :2] = wl["x0"][:, :2]
zeros = np.zeros(B, dtype=np.int32)
a = q.Solver([gm], N, 0.05, B); a.set_reference_trajectory(traj, off); a.set("u", wl["u_init"]); a.set_int("cold", zeros)
a.set("x0", wl["x0"]); a.set_reference_window(1); a.prepare()
xa, ua, ya = a.get("x"), a.get("u"), a.get("yref")
a.solve(); Ua, Xa, ita = a.get("u"), a.get("x"), a.get_int("qp_iter")
b = q.Solver([gm], N, 0.05, B); b.set_reference_trajectory(traj, off); b.set("u", wl["u_init"]); b.set_int("cold", zeros)
u0, st = b.step(wl["x0"], 1)
Ub, Xb, yb, itb = b.get("u"), b.get("x"), b.get("yref"), b.get_int("qp_iter")
print("yref equal", np.array_equal(ya, yb), "x0 equal", np.array_equal(a.get("x0"), b.get("x0")))
print("U max diff", np.abs(Ua - Ub).max(), "X max diff", np.abs(Xa - Xb).max(), "iters equal", np.array_equal(ita, itb), "n differing problems", (np.abs(Ua - Ub).max(axis=(1, 2)) > 0).sum())
print("u0 vs U[:,0]", np.array_equal(u0, Ub[:, 0]), "status", np.bincount(st))
# determinism of the field path itself
a.set("u", wl["u_init"]); a.set("x0", wl["x0"]); a.set_reference_window(1); a.prepare(); a.solve()
print("field path repeat equal", np.array_equal(a.get("u"), Ua))
b.set("u", wl["u_init"]); u0b, _ = b.step(wl["x0"], 1)
print("step path repeat equal", np.array_equal(b.get("u"), Ub))
print("stats", b.stat("time_qp_sol"), b.stat("time_lin"), b.stat("time_prep"), b.stat("time_tot"))
print("---- three periods with shift")
def fields(periods):
    s = q.Solver([gm], N, 0.05, B); s.set_reference_trajectory(traj, off); s.set("u", wl["u_init"]); s.set_int("cold", zeros)
    x0 = wl["x0"].copy(); out = []
    for i in range(1, periods + 1):
        s.set("x0", x0); s.set_reference_window(i); s.prepare()
        xp, up = s.get("x"), s.get("u")
        s.solve(); u0 = s.get("u", stage=0); st = s.get_int("status"); Uf = s.get("u"); s.shift()
        out.append(dict(u0=u0, st=st, U=s.get("u"), X=s.get("x"), xp=xp, up=up, x0=x0.copy(), Uf=Uf, pi=s.get("pi"), lam=s.get("lam"), it=s.get_int("qp_iter")))
        x0 = s.plant_step(x0.copy(), u0)
    return out
ref = fields(3)
s = q.Solver([gm], N, 0.05, B); s.set_reference_trajectory(traj, off); s.set("u", wl["u_init"]); s.set_int("cold", zeros)
x0 = wl["x0"].copy()
for i in range(1, 4):
    u0, st = s.step(x0, i, shift=True)
    r = ref[i - 1]
    print("period", i, "x0 in equal", np.array_equal(x0, r["x0"]), "u0 diff", np.abs(u0 - r["u0"]).max(), "U diff", np.abs(s.get("u") - r["U"]).max(), "X diff", np.abs(s.get("x") - r["X"]).max(),
          "pi diff", np.abs(s.get("pi") - r["pi"]).max(), "lam diff", np.abs(s.get("lam") - r["lam"]).max(), "it equal", np.array_equal(s.get_int("qp_iter"), r["it"]), "status", np.bincount(st))
    x0 = s.plant_step(x0.copy(), u0)
print("---- determinism: field path twice, 4 periods")
def fields2(periods, kern):
    s = q.Solver([gm], N, 0.05, B, qp_kernel=kern); s.set_reference_trajectory(traj, off); s.set("u", wl["u_init"]); s.set_int("cold", zeros)
    x0 = wl["x0"].copy(); out = []
    for i in range(1, periods + 1):
        s.set("x0", x0); s.set_reference_window(i); s.prepare(); s.solve(); u0 = s.get("u", stage=0); Uf = s.get("u"); s.shift()
        out.append(Uf)
        x0 = s.plant_step(x0.copy(), u0)
    return out
for kern in (1, 0):
    r1, r2 = fields2(4, kern), fields2(4, kern)
    print("kernel", kern, [float(np.abs(a - b).max()) for a, b in zip(r1, r2)], [int((np.abs(a - b).max(axis=(1,2)) > 0).sum()) for a, b in zip(r1, r2)])
