"""Full-size parity record: EVERY instance of BASELINE config 3 (4096, santal) and config 4 (65 536 over four shapes), N = 40, one
RTI control period through the C-ABI on the GPU vs the CPU oracle on the same inputs — the whole predicted trajectory (u, x), the
multipliers, status and IPM iteration count of every instance, not a strided subset.  The oracle here is the checker
(test infrastructure); writes gpurun_out/<tag>_full_parity.json."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import uclv_qs_pushing_matlab_b200 as q
from oracle import oracle as orc
from uclv_qs_pushing_matlab_b200.workloads import make_rti_workload, OBJECT_ORDER
from tests.workloads import gpu_model, oracle_model

N, DT = 40, 0.05
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
nthreads = os.cpu_count() or 8
out = {"what": __doc__.split("\n\n")[0].replace("\n", " "), "oracle_threads": nthreads, "tolerance_asserted_in_tests": 1e-8, "north_star_tolerance": 1e-6}
nseeds = int(sys.argv[2]) if len(sys.argv) > 2 else 0          # > 0: also config 4 with `nseeds` further seeds (100, 101, ...): a multi-seed sweep
cases = [("config3", 4096, ["santal"], 2), ("config4", 65536, list(OBJECT_ORDER), 3)]
cases += [("config4_seed%d" % sd, 65536, list(OBJECT_ORDER), sd) for sd in range(100, 100 + nseeds)]
for cfg, B, names, seed in cases:
    wl = make_rti_workload(B, N, seed=seed, n_objects=len(names))
    order = np.argsort(wl["object_id"], kind="stable")                    # contiguous per-object buckets
    wl = {k: np.ascontiguousarray(v[order]) for k, v in wl.items()}
    s = q.Solver([gpu_model(n) for n in names], N, DT, B)
    s.set("x0", wl["x0"]); s.set("yref", wl["yref"]); s.set("yref_e", wl["yref_e"]); s.set("u", wl["u_init"])
    s.set_int("cold", np.zeros(B, dtype=np.int32)); s.set_int("object_id", wl["object_id"])
    s.prepare(); s.solve(); s.sync()
    u, x, pi, lam = s.get("u"), s.get("x"), s.get("pi"), s.get("lam")
    st, it, res = s.get_int("status"), s.get_int("qp_iter"), s.get("res")
    rec = {"instances": B, "shapes": names, "gpu_status_ok_frac": float((st == 0).mean()),
           "gpu_kkt_residual_max": {"stat_eq_ineq": float(res[:, :3].max()), "comp": float(res[:, 3].max())}, "per_shape": {}}
    t_or = 0.0
    worst = dict(u0=0.0, u=0.0, x=0.0)
    n_it_eq = 0; n_st_eq = 0; it_diff_max = 0
    for o, name in enumerate(names):
        ia = np.where(wl["object_id"] == o)[0]
        w = {k: v[ia] for k, v in wl.items()}
        ocp = orc.Ocp(oracle_model(name), N, DT)
        t0 = time.perf_counter()
        pr = ocp.prepare(w["x0"], np.zeros(len(ia), dtype=np.int32), np.zeros((len(ia), N + 1, 4)), w["u_init"], nthreads=nthreads)
        ro = ocp.solve("rti", pr["x0"], w["yref"], w["yref_e"], pr["x"], pr["u"], nthreads=nthreads)
        t_or += time.perf_counter() - t0
        e_u0 = np.abs(u[ia][:, 0] - ro["u"][:, 0]).max(1)
        e_u = np.abs(u[ia] - ro["u"]).max(axis=(1, 2)); e_x = np.abs(x[ia] - ro["x"]).max(axis=(1, 2))
        lam_scale = max(1.0, float(np.abs(ro["lam"]).max()))
        rec["per_shape"][name] = {
            "instances": int(len(ia)),
            "u0_abs_err_max": float(e_u0.max()), "u0_abs_err_p99": float(np.quantile(e_u0, 0.99)),
            "u_abs_err_max": float(e_u.max()), "x_abs_err_max": float(e_x.max()),
            "lam_err_over_scale_max": float(np.abs(lam[ia] - ro["lam"]).max() / lam_scale),
            "instances_with_u0_err_above_1e-8": int((e_u0 >= 1e-8).sum()), "instances_with_u0_err_above_1e-6": int((e_u0 >= 1e-6).sum()),
            "status_equal": int((st[ia] == ro["status"]).sum()), "qp_iter_equal_frac": float((it[ia] == ro["qp_iter"]).mean()),
            "qp_iter_abs_diff_max": int(np.abs(it[ia] - ro["qp_iter"]).max()), "k_ipm_mean_gpu": float(it[ia].mean()), "k_ipm_mean_oracle": float(ro["qp_iter"].mean())}
        worst["u0"] = max(worst["u0"], float(e_u0.max())); worst["u"] = max(worst["u"], float(e_u.max())); worst["x"] = max(worst["x"], float(e_x.max()))
        n_it_eq += int((it[ia] == ro["qp_iter"]).sum()); n_st_eq += int((st[ia] == ro["status"]).sum()); it_diff_max = max(it_diff_max, int(np.abs(it[ia] - ro["qp_iter"]).max()))
    rec.update({"u0_abs_err_max": worst["u0"], "u_abs_err_max": worst["u"], "x_abs_err_max": worst["x"], "status_equal_frac": n_st_eq / B,
                "qp_iter_equal_frac": n_it_eq / B, "qp_iter_abs_diff_max": it_diff_max, "oracle_seconds": t_or,
                "pass_1e-8_on_every_instance": bool(worst["u"] < 1e-8 and worst["x"] < 1e-8 and n_st_eq == B)})
    out[cfg] = rec
    print(cfg, json.dumps({k: v for k, v in rec.items() if k != "per_shape"}), flush=True)
    del s
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{tag}_full_parity.json"), "w"), indent=1)
if nseeds:
    sw = [out[c] for c in out if c.startswith("config4_seed")]
    out["multi_seed_summary"] = {"seeds": nseeds, "instances": int(sum(r["instances"] for r in sw)), "u0_abs_err_max": max(r["u0_abs_err_max"] for r in sw),
                                 "u_abs_err_max": max(r["u_abs_err_max"] for r in sw), "x_abs_err_max": max(r["x_abs_err_max"] for r in sw),
                                 "status_equal_frac": min(r["status_equal_frac"] for r in sw), "gpu_status_ok_frac": min(r["gpu_status_ok_frac"] for r in sw),
                                 "qp_iter_equal_frac_min": min(r["qp_iter_equal_frac"] for r in sw), "qp_iter_abs_diff_max": max(r["qp_iter_abs_diff_max"] for r in sw),
                                 "all_pass_1e-8": bool(all(r["pass_1e-8_on_every_instance"] for r in sw))}
    print("multi_seed_summary", json.dumps(out["multi_seed_summary"]), flush=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{tag}_full_parity.json"), "w"), indent=1)
sys.exit(0 if all(out[c]["pass_1e-8_on_every_instance"] for c in out if c.startswith("config")) else 1)
