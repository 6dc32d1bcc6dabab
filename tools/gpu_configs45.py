"""Per-GPU shares of BASELINE configs 4 and 5 (parity-test configurations, measured for the record):
config 4: 65536 instances, 4 shapes, N = 40, RTI, over 8 GPUs -> 8192 per GPU;
config 5: 262144 instances, 4 shapes, N = 100, full SQP (<= 30 iterations), over 8 GPUs -> 32768 per GPU."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import torch
import uclv_qs_pushing_matlab_b200 as q
from uclv_qs_pushing_matlab_b200.workloads import make_rti_workload, OBJECT_ORDER
from uclv_qs_pushing_matlab_b200 import sharding
from tests.workloads import gpu_model

dev = torch.device("cuda:0")
gms = [gpu_model(n) for n in OBJECT_ORDER]
out = {}
for tag, B, N, mode, mixed, seed in (("config4_share", 8192, 40, 0, False, 3), ("config5_share", 32768, 100, 1, True, 4), ("config5_share_4096", 4096, 100, 1, True, 4)):
    wl = make_rti_workload(B, N, seed=seed, n_objects=4, mixed_modes=mixed)
    order = np.argsort(wl["object_id"], kind="stable")                    # contiguous per-object buckets per GPU
    wl = {k: v[order] for k, v in wl.items()}
    s = q.Solver(gms, N, 0.05, B, mode=mode, qp_kernel=int(sys.argv[1]) if len(sys.argv) > 1 else 2)
    d = {k: torch.from_numpy(np.ascontiguousarray(wl[k])).to(dev) for k in ("x0", "yref", "yref_e", "u_init")}
    oid = torch.from_numpy(wl["object_id"]).to(dev); cold = torch.zeros(B, dtype=torch.int32, device=dev)
    ts = []
    for r in range(3):
        s.set("x0", d["x0"]); s.set("yref", d["yref"]); s.set("yref_e", d["yref_e"]); s.set("u", d["u_init"]); s.set_int("cold", cold); s.set_int("object_id", oid)
        s.sync(); t0 = time.perf_counter(); s.prepare(); s.solve(); s.sync(); ts.append(time.perf_counter() - t0)
    st = s.get_int("status"); it = s.get_int("sqp_iter"); qi = s.get_int("qp_iter")
    t = min(ts)
    out[tag] = dict(B=B, N=N, mode="sqp" if mode else "rti", ms=1e3 * t, problems_per_s=B / t, sqp_iterations_per_s=float(it.sum()) / t,
                    converged_frac=float((st == 0).mean()), status_hist={int(k): int((st == k).sum()) for k in np.unique(st)},
                    sqp_iter_mean=float(it.mean()), qp_iter_per_sqp_iter=float(qi.sum()) / max(1, int(it.sum())))
    print(tag, out[tag], flush=True)
    del s
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs45.json"), "w"), indent=1)
