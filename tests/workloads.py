"""Test helpers: the same slider model built twice from the packaged tables — once for the product
(C-ABI, GPU) and once for the oracle — plus re-exports of the synthetic workloads."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from uclv_qs_pushing_matlab_b200.workloads import (OBJECT_ORDER, make_feasible_start_workload, make_rti_workload as _mk, make_samples_config2,  # noqa: E402,F401
                                                    packaged_tables, reference_line)
from uclv_qs_pushing_matlab_b200.object_selection import OBJECT_TABLE  # noqa: E402


def oracle_model(name):
    from oracle import oracle as orc
    t = packaged_tables()[name]
    return orc.Model.create(t["knots"], t["ctrl_xy"], 3, OBJECT_TABLE[name]["mu_sp"], t["c_ellipse"], True)


def hostsim_model(name):
    from tests.hostsim import hostsim as hs
    t = packaged_tables()[name]
    return hs.Model.create(t["knots"], t["ctrl_xy"], 3, OBJECT_TABLE[name]["mu_sp"], t["c_ellipse"], True)


def gpu_model(name):
    from uclv_qs_pushing_matlab_b200 import Model
    t = packaged_tables()[name]
    return Model.from_tables(t["knots"], t["ctrl_xy"], 3, OBJECT_TABLE[name]["mu_sp"], t["c_ellipse"], True)


def packaged_model_pair(name):
    return gpu_model(name), oracle_model(name)


def make_rti_workload(om=None, batch=64, N=10, seed=2, dt=0.05, **kw):
    return _mk(batch, N, dt=dt, seed=seed, **kw)


def make_vbound_workload(batch, N, seed=2):
    """Config-3-like workload that makes the parked constraint set h = [u_n; u_t -+ v_bound(s)] (NMPC_controller.m:238)
    bite: contact points spread over the high-curvature zone of the santal outline (v_bound < u_t_ub there, and
    v_bound'(s) up to +-30 1/s) and a large tangential velocity in the initial guess."""
    wl = make_rti_workload(None, batch=batch, N=N, seed=seed)
    wl["x0"][:, 3] = np.linspace(-0.004, 0.0045, batch)
    wl["u_init"][:, :, 1] = 0.03 * np.sign(np.linspace(-1.0, 1.0, batch))[:, None]
    return wl


VARIANT_LH, VARIANT_UH = (0.0, -0.1, 0.0), (0.03, 0.0, 0.1)       # [u_n_lb, 2 u_t_lb, 0] / [u_n_ub, 0, 2 u_t_ub]
