"""Quick GPU sanity + timing sweep (development aid; run under gpurun)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import uclv_qs_pushing_matlab_b200 as q
from uclv_qs_pushing_matlab_b200.workloads import make_rti_workload, make_samples_config2
from tests.workloads import gpu_model

gm = gpu_model("santal")
dev = torch.device("cuda:0")

def time_rti(B, N, ppw, tol, reps=5, **kw):
    wl = make_rti_workload(B, N, seed=2)
    s = q.Solver([gm], N, 0.05, B, qp_tol=tol, problems_per_warp=ppw, **kw)
    x0 = torch.from_numpy(wl["x0"]).to(dev); yr = torch.from_numpy(wl["yref"]).to(dev); ye = torch.from_numpy(wl["yref_e"]).to(dev)
    ui = torch.from_numpy(wl["u_init"]).to(dev); cold = torch.zeros(B, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ts = []
    for r in range(reps + 2):
        s.set("x0", x0); s.set("yref", yr); s.set("yref_e", ye); s.set("u", ui); s.set_int("cold", cold)
        s.sync(); t0 = time.perf_counter()
        s.prepare(); s.solve(); s.sync()
        t1 = time.perf_counter()
        if r >= 2: ts.append(t1 - t0)
    it = s.get_int("qp_iter"); st = s.get_int("status")
    return dict(B=B, N=N, ppw=ppw, tol=tol, ms=1e3 * min(ts), its_per_s=B / min(ts), qp_iter_mean=float(it.mean()), qp_iter_max=int(it.max()),
                status_ok=float((st == 0).mean()), t_lin_ms=1e3 * s.stat("time_lin"), t_qp_ms=1e3 * s.stat("time_qp_sol"), t_prep_ms=1e3 * s.stat("time_prep"))

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
if __name__ == "__main__":
    out = []
    for B in (4096, 16384):
        for ppw in (32, 16, 8, 4):
            r = time_rti(B, 40, ppw, 1e-12); print(r, flush=True); out.append(r)
    for tol in (1e-6, 1e-8, 1e-10):
        r = time_rti(4096, 40, 8, tol); print(r, flush=True); out.append(r)
    # config 2 kernel
    n = 1 << 20
    x, u = make_samples_config2(gm.b, n, knots=gm.S)
    xd, ud = torch.from_numpy(x).to(dev), torch.from_numpy(u).to(dev)
    Phi = torch.empty(n, 4, dtype=torch.float64, device=dev); A = torch.empty(n, 4, 4, dtype=torch.float64, device=dev); Bm = torch.empty(n, 4, 2, dtype=torch.float64, device=dev)
    for _ in range(3): gm.eval_erk4_sens_device(xd, ud, 0.05, Phi, A, Bm)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): gm.eval_erk4_sens_device(xd, ud, 0.05, Phi, A, Bm)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print("erk4_sens 1M samples: %.3f ms -> %.1f Msamples/s" % (dt * 1e3, n / dt / 1e6))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sanity.json"), "w"), indent=1)
