"""Full SQP (config-5-like start, 4 shapes) over horizons: run once per QSPUSH_LOCKSTEP setting (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import uclv_qs_pushing_matlab_b200 as q
from uclv_qs_pushing_matlab_b200.workloads import make_rti_workload, OBJECT_ORDER
from tests.workloads import gpu_model
dev = torch.device("cuda:0")
gms = [gpu_model(n) for n in OBJECT_ORDER]
for N in [int(a) for a in sys.argv[1:]] or [10, 31, 40, 63, 100]:
    B = 16384
    wl = make_rti_workload(B, N, seed=4, n_objects=4, mixed_modes=True)
    order = np.argsort(wl["object_id"], kind="stable"); wl = {k: v[order] for k, v in wl.items()}
    s = q.Solver(gms, N, 0.05, B, mode=1)
    d = {k: torch.from_numpy(np.ascontiguousarray(wl[k])).to(dev) for k in ("x0", "yref", "yref_e", "u_init")}
    oid = torch.from_numpy(wl["object_id"]).to(dev); cold = torch.zeros(B, dtype=torch.int32, device=dev)
    ts = []
    for r in range(3):
        s.set("x0", d["x0"]); s.set("yref", d["yref"]); s.set("yref_e", d["yref_e"]); s.set("u", d["u_init"]); s.set_int("cold", cold); s.set_int("object_id", oid)
        s.sync(); t0 = time.perf_counter(); s.prepare(); s.solve(); s.sync(); ts.append(time.perf_counter() - t0)
    it = s.get_int("sqp_iter"); st = s.get_int("status")
    print("SQP N %3d B %d : %9.3f ms  sqp it/s %.4g  conv %.3f  lockstep=%s" % (N, B, 1e3 * min(ts), it.sum() / min(ts), (st == 0).mean(), os.environ.get("QSPUSH_LOCKSTEP", "rule")), flush=True)
    del s
