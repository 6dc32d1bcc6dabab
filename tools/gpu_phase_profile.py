"""Per-phase cycle breakdown of k_qp_warp (needs a library built with -DQW_PROFILE; development aid)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import uclv_qs_pushing_matlab_b200 as q
from tests.workloads import gpu_model, make_rti_workload
lib = C.CDLL(q._lib.LIB_PATH)
gm = gpu_model("santal"); B, N = 4096, 40
wl = make_rti_workload(None, batch=B, N=N, seed=2)
s = q.Solver([gm], N, 0.05, B, qp_kernel=1)
buf = (C.c_ulonglong * 16)()
for rep in range(2):
    s.set("x0", wl["x0"]); s.set("yref", wl["yref"]); s.set("yref_e", wl["yref_e"]); s.set("u", wl["u_init"]); s.set_int("cold", np.zeros(B, dtype=np.int32))
    s.prepare(); s.sync(); lib.qspush_dev_phase_cycles(buf)
    s.solve(); s.sync(); lib.qspush_dev_phase_cycles(buf)
v = np.array(list(buf), dtype=np.float64)
names = ["vote wait", "(1) residuals+tests", "(2) barrier+elements+local combine", "(3) element scan", "(4) local Riccati", "(5/7) rhs + affine stats", "solve (x2)", "(6) sigma etc", "(8) step+update", "init (bind)", "next ticket (queue atomic + order)", "L2 prefetch of the next problem", "write-back", "K5 epilogue (x += dx, u += du, cost, status)"]
tot = v[:14].sum(); its = s.get_int("qp_iter").sum()
print("total warp-cycles %.3e over %d IPM iterations -> %.0f cycles per iteration per warp" % (tot, its, tot / its))
for n, c in zip(names, v): print("  %-38s %6.2f %%   %8.0f cycles/iter" % (n, 100 * c / tot, c / its))
