"""Where the end-to-end step spends its time: uploads, compute, download (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import uclv_qs_pushing_matlab_b200 as q
from tests.workloads import gpu_model
from uclv_qs_pushing_matlab_b200.workloads import make_rti_workload
B, N = 4096, 40
gm = gpu_model("santal"); wl = make_rti_workload(B, N, seed=2)
s = q.Solver([gm], N, 0.05, B)
h = {k: torch.from_numpy(wl[k]).pin_memory() for k in ("x0", "yref", "yref_e", "u_init")}
hc = torch.zeros(B, dtype=torch.int32).pin_memory(); hu = torch.empty(B, 2, dtype=torch.float64).pin_memory(); hs = torch.empty(B, dtype=torch.int32).pin_memory()
def up():
    s.set("x0", h["x0"]); s.set("u", h["u_init"]); s.set_int("cold", hc); s.set("yref", h["yref"]); s.set("yref_e", h["yref_e"])
def comp():
    s.prepare(); s.solve()
def down():
    s.get("u", stage=0, out=hu); s.get_int("status", out=hs)
def t(f, n=30):
    ts = []
    for _ in range(n):
        s.sync(); t0 = time.perf_counter(); f(); s.sync(); ts.append(time.perf_counter() - t0)
    ts.sort(); return 1e3 * ts[len(ts) // 2]
for _ in range(5): up(); comp(); down()
print("copy stream:", "off" if os.environ.get("QSPUSH_NO_COPY_STREAM") else "on")
print("upload only      %.3f ms" % t(up)); print("compute only     %.3f ms" % t(comp)); print("download only    %.3f ms" % t(down))
print("upload+compute   %.3f ms" % t(lambda: (up(), comp()))); print("full step        %.3f ms" % t(lambda: (up(), comp(), down())))
t0 = time.perf_counter()
for _ in range(200): up()
print("host time per upload sequence (no sync) %.3f ms" % (1e3 * (time.perf_counter() - t0) / 200)); s.sync()
