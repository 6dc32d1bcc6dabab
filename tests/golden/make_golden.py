"""Generate tests/golden/*.npz from the CPU oracle (run in the build container: python tests/golden/make_golden.py).

The reference ships no golden vectors and cannot run here (MATLAB + acados v0.2.1 + CasADi, SURVEY.md
section 8c), so these fixtures freeze the ORACLE's outputs: they guard the oracle against regressions
and let the GPU box (which has no /root/reference) check the CUDA path against committed numbers.
Inputs are seeded; the slider models come from the packaged outline tables
(uclv_qs_pushing_matlab_b200/data/objects.json, generated from the reference .ply files by
tools/make_object_tables.py).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import arbiter as arb  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests.workloads import OBJECT_ORDER, make_rti_workload, make_samples_config2, oracle_model  # noqa: E402


def main():
    for name in OBJECT_ORDER:
        om = oracle_model(name)
        x, u = make_samples_config2(om.b, 640, seed=1, n_adversarial=256, knots=om.S)
        s = x[:, 3]
        sp = om.eval_spline(s, wrap=2)
        f, Jx, Ju = om.dynamics(x, u, jac=True)
        Phi, A, B = om.erk4_sens(x, u, 0.05)
        vb = np.array([om.v_bound(v) for v in s[:256]])
        np.savez_compressed(os.path.join(HERE, f"eval_{name}.npz"), x=x, u=u, C=sp["C"], Cd=sp["Cd"], Cdd=sp["Cdd"], t=sp["t"],
                            n=sp["n"], kappa=sp["kappa"], f=f, Jx=Jx, Ju=Ju, Phi=Phi, A=A, B=B, v_bound=vb[:, 0], t_angle=vb[:, 1])
    # one small RTI batch (config 3 shape, reduced): prepare + linearise + QP + step
    om = oracle_model("santal")
    B, N = 32, 40
    wl = make_rti_workload(None, batch=B, N=N, seed=2)
    ocp = orc.Ocp(om, N, 0.05)
    pr = ocp.prepare(wl["x0"], np.zeros(B, dtype=np.int32), np.zeros((B, N + 1, 4)), wl["u_init"])
    qp = ocp.qp(pr["x0"], wl["yref"], wl["yref_e"], pr["x"], pr["u"])
    rti = ocp.solve("rti", pr["x0"], wl["yref"], wl["yref_e"], pr["x"], pr["u"])
    # the EXACT solution of the same QPs (oracle/qs_arbiter.cpp, __float128 active-set method + KKT certificate)
    d = ocp.qp_data(pr["x0"], wl["yref"], wl["yref_e"], pr["x"], pr["u"])
    ex = arb.solve_exact(d, arb.working_set_from_ipm(qp["lam"], qp["t"], d["on"]))
    assert (ex["status"] == 0).all() and ex["kkt"].max() < 1e-18
    np.savez_compressed(os.path.join(HERE, "rti_santal_N40.npz"), x0=wl["x0"], yref=wl["yref"], yref_e=wl["yref_e"], u_init=wl["u_init"],
                        x0_wrapped=pr["x0"], x_prep=pr["x"], u_prep=pr["u"], du=qp["du"], dx=qp["dx"], qp_pi=qp["pi"], qp_lam=qp["lam"],
                        qp_iters=qp["iters"], x=rti["x"], u=rti["u"], pi=rti["pi"], lam=rti["lam"], cost=rti["cost"],
                        du_exact=ex["du"], dx_exact=ex["dx"], pi_exact=ex["pi"], lam_exact=ex["lam"], kkt_exact=ex["kkt"])
    print("rti fixture: oracle vs exact du %.2e dx %.2e" % (np.abs(qp["du"] - ex["du"]).max(), np.abs(qp["dx"] - ex["dx"]).max()))
    # full SQP to convergence on a short horizon (config 1 shape: N = 10)
    B, N = 16, 10
    wl = make_rti_workload(None, batch=B, N=N, seed=4, mixed_modes=True)
    ocp = orc.Ocp(om, N, 0.05)
    pr = ocp.prepare(wl["x0"], np.zeros(B, dtype=np.int32), np.zeros((B, N + 1, 4)), wl["u_init"])
    sq = ocp.solve("sqp", pr["x0"], wl["yref"], wl["yref_e"], pr["x"], pr["u"])
    np.savez_compressed(os.path.join(HERE, "sqp_santal_N10.npz"), x0=wl["x0"], yref=wl["yref"], yref_e=wl["yref_e"], u_init=wl["u_init"],
                        x=sq["x"], u=sq["u"], status=sq["status"], sqp_iter=sq["sqp_iter"], cost=sq["cost"], res=sq["res"])
    print("golden fixtures written to", HERE)
    print("sqp status", sq["status"], "iters", sq["sqp_iter"], "res max", sq["res"].max(0))


if __name__ == "__main__":
    main()
