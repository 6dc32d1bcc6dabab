"""Seeded inputs for tools/acados_golden.m (run next to it in the reference workspace; needs numpy + scipy only).

    python tools/acados_golden_inputs.py acados_golden_inputs.mat

Cases (the same seeds and ranges as uclv_qs_pushing_matlab_b200/workloads.py):
  config1 : santal, Hp = 10, x0 = 0, straight-line reference 0 -> 0.10 m, closed loop for 201 periods from a cold start (main.m)
  config3 : santal, Hp = 40, 32 random initial poses (seed 2), initial guess u = [0.01; 0], one solve each
  config3s: santal, Hp = 10, 32 random initial poses (seed 2) — the cheap horizon
  config4 : balea / montana / pulirapid, Hp = 40, 16 poses each (seed 3)
"""
import importlib.util
import os
import sys

import numpy as np
from scipy.io import savemat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("_wl", os.path.join(ROOT, "uclv_qs_pushing_matlab_b200", "workloads.py"))
WL = importlib.util.module_from_spec(spec)
spec.loader.exec_module(WL)

W = np.diag([1.0, 1.0, 1e-3, 0.0, 1e-3, 1e-3])                 # main.m:82-84 (y = [x; u] order)
WE = np.diag([2e5, 2e5, 20.0, 0.0])


def single_solves(name, obj, Hp, count, seed, dt=0.05):
    wl = WL.make_rti_workload(count, Hp, dt=dt, seed=seed)
    T = Hp                                                       # window of period 1 = the whole reference
    y_ref = np.transpose(wl["yref"], (2, 1, 0))                  # (6, T, count)
    return dict(name=name, object=obj, Hp=Hp, dt=dt, W=W, We=WE, cold=False, x0=wl["x0"].T, y_ref=y_ref,
                u_init=np.transpose(wl["u_init"], (2, 1, 0)), closed_loop_steps=0, cl_x0=np.zeros(4), cl_y_ref=np.zeros((6, T)))


def closed_loop(name, obj, Hp, steps, dt=0.05):
    t = np.arange(steps) * dt
    y = np.zeros((6, steps)); y[0] = np.minimum(0.01 * t, 0.10)
    return dict(name=name, object=obj, Hp=Hp, dt=dt, W=W, We=WE, cold=True, x0=np.zeros((4, 0)), y_ref=np.zeros((6, Hp, 0)),
                u_init=np.zeros((2, Hp, 0)), closed_loop_steps=steps, cl_x0=np.zeros(4), cl_y_ref=y)


def main(path):
    cfg = [closed_loop("config1", "santal", 10, 201), single_solves("config3s", "santal", 10, 32, 2), single_solves("config3", "santal", 40, 32, 2)]
    cfg += [single_solves(f"config4_{o}", o, 40, 16, 3) for o in ("balea", "montana", "pulirapid")]
    arr = np.empty(len(cfg), dtype=object)
    for i, c in enumerate(cfg):
        arr[i] = c
    savemat(path, {"cfg": arr}, do_compression=True)
    print("wrote", path, "with", len(cfg), "configurations")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "acados_golden_inputs.mat")
